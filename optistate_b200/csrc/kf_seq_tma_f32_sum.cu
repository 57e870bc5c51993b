// Streamed SEQUENTIAL kernel instantiations: Real = float, kSummary = true (one translation unit per pair, built in parallel).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<float, true>(const Params<typename Lanes<float>::scalar> &, cudaStream_t);
}
