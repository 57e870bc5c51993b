// Streamed SEQUENTIAL kernel instantiations: Real = F2, kSummary = true, full packed P (one translation unit per variant, built in parallel).
#include "kf_seq_tma_host.cuh"

namespace okf {
template int launch_seq_tma<F2, true, false>(const Params<typename Lanes<F2>::scalar> &, cudaStream_t);
}
