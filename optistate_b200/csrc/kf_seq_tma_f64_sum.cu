// Streamed SEQUENTIAL kernel instantiations: Real = double, kSummary = true (one translation unit per pair, built in parallel).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<double, true>(const Params<typename Lanes<double>::scalar> &, cudaStream_t);
}
