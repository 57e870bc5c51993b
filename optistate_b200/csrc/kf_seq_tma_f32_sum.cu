// Streamed SEQUENTIAL kernel instantiations: Real = float, kSummary = true, full packed P (one translation unit per variant, built in parallel).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<float, true, false>(const Params<typename Lanes<float>::scalar> &, cudaStream_t);
}
