// Streamed SEQUENTIAL kernel instantiations: Real = float, kSummary = true, decoupled groups of P (kBlock).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<float, true, true>(const Params<typename Lanes<float>::scalar> &, cudaStream_t);
}
