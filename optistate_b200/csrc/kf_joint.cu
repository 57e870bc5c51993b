// JOINT kernel instantiations (kf_joint_coop.cuh) and their launch.
#include "kf_joint_coop.cuh"
#include "kf_launch.cuh"

namespace okf {

template <typename Real>
int launch_joint(const Params<Real> &p, cudaStream_t stream) {
    const size_t smem = jc_smem_bytes<Real>();
    auto kern = kf_joint_coop_kernel<Real>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return OPTI_KF_E_CUDA;
    const unsigned blocks = (unsigned)((p.N + JC_TRAJ - 1) / JC_TRAJ);
    kern<<<blocks, JC_THREADS, smem, stream>>>(p);
    return OPTI_KF_OK;
}

template int launch_joint<double>(const Params<double> &, cudaStream_t);
template int launch_joint<float>(const Params<float> &, cudaStream_t);

}  // namespace okf
