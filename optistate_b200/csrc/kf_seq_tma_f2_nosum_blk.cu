// Streamed SEQUENTIAL kernel instantiations: Real = F2, kSummary = false, decoupled groups of P (kBlock).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<F2, false, true>(const Params<typename Lanes<F2>::scalar> &, cudaStream_t);
}
