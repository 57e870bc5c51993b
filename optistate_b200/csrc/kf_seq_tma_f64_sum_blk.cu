// Streamed SEQUENTIAL kernel instantiations: Real = double, kSummary = true, decoupled groups of P (kBlock).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<double, true, true>(const Params<typename Lanes<double>::scalar> &, cudaStream_t);
}
