// Streamed SEQUENTIAL kernel instantiations: Real = double, kSummary = false, full packed P (one translation unit per variant, built in parallel).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<double, false, false>(const Params<typename Lanes<double>::scalar> &, cudaStream_t);
}
