// Streamed SEQUENTIAL kernel, latency variant (one warp per block): Real = float, decoupled groups of P, with and without the summary.
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma_lone<float, false>(const Params<typename Lanes<float>::scalar> &, cudaStream_t);
template int launch_seq_tma_lone<float, true>(const Params<typename Lanes<float>::scalar> &, cudaStream_t);
}
