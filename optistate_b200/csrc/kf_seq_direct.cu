// Direct-load SEQUENTIAL kernel instantiations (kf_seq.cuh) and their launch.
#include "kf_launch.cuh"
#include "kf_seq.cuh"

namespace okf {

template <typename Real>
int launch_seq_direct(const Params<Real> &p, cudaStream_t stream) {
    constexpr int kThreads = 128;
    const unsigned blocks = (unsigned)((p.N + kThreads - 1) / kThreads);
    const size_t smem = (size_t)SEQ_NOISE_ROWS * kThreads * sizeof(Real);
    const bool mpc = p.cov_model == OPTI_KF_COV_MPC;
    if (p.summary) {
        if (mpc) kf_seq_kernel<Real, true, true><<<blocks, kThreads, smem, stream>>>(p);
        else kf_seq_kernel<Real, true, false><<<blocks, kThreads, smem, stream>>>(p);
    } else {
        if (mpc) kf_seq_kernel<Real, false, true><<<blocks, kThreads, smem, stream>>>(p);
        else kf_seq_kernel<Real, false, false><<<blocks, kThreads, smem, stream>>>(p);
    }
    return OPTI_KF_OK;
}

template int launch_seq_direct<double>(const Params<double> &, cudaStream_t);
template int launch_seq_direct<float>(const Params<float> &, cudaStream_t);

}  // namespace okf
