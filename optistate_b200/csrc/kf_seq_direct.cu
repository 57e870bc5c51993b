// Direct-load SEQUENTIAL kernel instantiations (kf_seq.cuh) and their launch.
#include "kf_launch.cuh"
#include "kf_seq.cuh"

namespace okf {

template <typename Real, bool kSummary>
static void launch_seq_direct_k(const Params<Real> &p, unsigned blocks, int threads, size_t smem, cudaStream_t stream) {
    if (p.cov_model == OPTI_KF_COV_MPC) kf_seq_kernel<Real, kSummary, true, false><<<blocks, threads, smem, stream>>>(p);
    else if (p.block) kf_seq_kernel<Real, kSummary, false, true><<<blocks, threads, smem, stream>>>(p);
    else kf_seq_kernel<Real, kSummary, false, false><<<blocks, threads, smem, stream>>>(p);
}

template <typename Real>
int launch_seq_direct(const Params<Real> &p, cudaStream_t stream) {
    constexpr int kThreads = 128;
    const unsigned blocks = (unsigned)((p.N + kThreads - 1) / kThreads);
    const size_t smem = (size_t)SEQ_NOISE_ROWS * kThreads * sizeof(Real);
    if (p.summary) launch_seq_direct_k<Real, true>(p, blocks, kThreads, smem, stream);
    else launch_seq_direct_k<Real, false>(p, blocks, kThreads, smem, stream);
    return OPTI_KF_OK;
}

template int launch_seq_direct<double>(const Params<double> &, cudaStream_t);
template int launch_seq_direct<float>(const Params<float> &, cudaStream_t);

}  // namespace okf
