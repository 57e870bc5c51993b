// Batched convex force MPC (kf_mpc.cuh): launch.
#include "kf_mpc.cuh"

#include "kf_launch.cuh"

namespace okf {

int launch_mpc(const MpcParams &p, cudaStream_t stream) {
    const size_t smem = mpc_smem_bytes(p.max_legs);
    if (cudaFuncSetAttribute(kf_mpc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return OPTI_KF_E_CUDA;
    const unsigned blocks = (unsigned)((p.N + MPC_WARPS - 1) / MPC_WARPS);
    kf_mpc_kernel<<<blocks, 32 * MPC_WARPS, smem, stream>>>(p);
    return OPTI_KF_OK;
}

}  // namespace okf
