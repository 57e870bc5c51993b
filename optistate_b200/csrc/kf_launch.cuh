// Host-side launchers of the filter kernels.  Every group of kernel instantiations lives in its own translation unit
// (kf_seq_tma_*.cu, kf_seq_direct.cu, kf_joint.cu) so that the in-tree build compiles them in parallel; kf_abi.cu only
// sees these declarations.
#pragma once

#include "kf_common.cuh"
#include "kf_mpc_params.cuh"

namespace okf {

// Streamed SEQUENTIAL kernel (kf_seq_tma.cuh) for Real = double | float | F2 (two FP32 trajectories per thread), with
// (kSummary) or without the per-trajectory summary, on the full packed P or its decoupled groups (kBlock, Params.block).  Picks the instantiation from the descriptor (per-step outputs,
// covariance model).  Returns OPTI_KF_OK, OPTI_KF_E_CUDA, or +1 when the tensor maps could not be built (the caller then
// falls back to the direct-load kernel).
template <typename Real, bool kSummary, bool kBlock>
int launch_seq_tma(const Params<typename Lanes<Real>::scalar> &p, cudaStream_t stream);

// The same kernel with one warp per block, for batches with no more trajectories than streams (a lone warp per scheduler: latency,
// not throughput, decides); Real = double | float, decoupled-group form and predict() covariance model only.
template <typename Real, bool kSummary>
int launch_seq_tma_lone(const Params<typename Lanes<Real>::scalar> &p, cudaStream_t stream);

// Direct-load SEQUENTIAL kernel (kf_seq.cuh).
template <typename Real>
int launch_seq_direct(const Params<Real> &p, cudaStream_t stream);

// JOINT kernel, four lanes per trajectory (kf_joint_coop.cuh).
template <typename Real>
int launch_joint(const Params<Real> &p, cudaStream_t stream);

// Batched convex force MPC, one warp per problem (kf_mpc.cuh).
int launch_mpc(const MpcParams &p, cudaStream_t stream);

}  // namespace okf
