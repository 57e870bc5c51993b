// Streamed SEQUENTIAL kernel, latency variant (one warp per block): Real = double, decoupled groups of P, with and without the summary.
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma_lone<double, false>(const Params<typename Lanes<double>::scalar> &, cudaStream_t);
template int launch_seq_tma_lone<double, true>(const Params<typename Lanes<double>::scalar> &, cudaStream_t);
}
