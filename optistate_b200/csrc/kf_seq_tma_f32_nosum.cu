// Streamed SEQUENTIAL kernel instantiations: Real = float, kSummary = false, full packed P (one translation unit per variant, built in parallel).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<float, false, false>(const Params<typename Lanes<float>::scalar> &, cudaStream_t);
}
