// Step before the filter: the driver's Q / R identification pass,
// data_conversion_Kalman_to_Training.py:31-109.  For every step i of a recording the state is reset to the ground
// truth gt[i], the model is stepped once with the forces of that step (the reference obtains them from its MPC inside
// predict_mpc; here they are an input stream) and the measurement of step i+1 is formed:
//     model residual        e_x[i] = gt[i+1] - next_state(gt[i], p[i], f[i])             (12)   :81-83
//     measurement residual  e_z[i] = H gt[i+1] - z(imu, p, dp, contact)[i+1]              (10)   :94-98
//     Q = diag(var_i e_x),  R = diag(var_i e_z)   (population variance, np.var)                    :88-89,104-106
// Every (trajectory, step) pair is independent: pass 1 is one thread per pair (residuals to scratch, trajectory index
// fastest), pass 2 one thread per (trajectory, component) with a two-pass mean / variance, like np.var.
// Reference quirk kept behind a flag: the driver appends the SAME array object KF.z at every step (:74), so all stored
// measurements alias the last one; `alias_last_measurement` reproduces that, the default uses z[i+1] as intended.
#pragma once

#include "kf_common.cuh"

namespace okf {

constexpr int IDENT_ROWS = NX + NZ;  // residual components per (trajectory, step)

template <typename Real>
__global__ void __launch_bounds__(128) kf_identify_residuals_kernel(Params<Real> prm, const Real *__restrict__ gt, int alias_last,
                                                                     Real *__restrict__ resid /* [T-1][22][N] */) {
    const long long N = prm.N, S = prm.S, T = prm.T;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (step, trajectory), trajectory fastest
    if (idx >= (T - 1) * N) return;
    const long long t = idx / N, i = idx % N;
    const long long s = stream_of(prm, i);
    Real x[NX], pf[12], ff[12], Rm[9];
#pragma unroll
    for (int c = 0; c < NX; ++c) x[c] = gt[(t * NX + c) * S + s];
#pragma unroll
    for (int c = 0; c < 12; ++c) { pf[c] = prm.p[(t * 12 + c) * S + s]; ff[c] = prm.f[(t * 12 + c) * S + s]; }
    propagate_mean(prm, x, pf, ff, Rm);
    const long long tz = alias_last ? T - 1 : t + 1;
    Real imu[6], pz[12], dp[12], contact[4], z[NZ];
#pragma unroll
    for (int c = 0; c < 6; ++c) imu[c] = prm.imu[(tz * 6 + c) * S + s];
#pragma unroll
    for (int c = 0; c < 12; ++c) { pz[c] = prm.p[(tz * 12 + c) * S + s]; dp[c] = prm.dp[(tz * 12 + c) * S + s]; }
#pragma unroll
    for (int c = 0; c < 4; ++c) contact[c] = prm.contact[(tz * 4 + c) * S + s];
    const bool all_swing = form_measurement(imu, pz, dp, contact, z);
    if (all_swing && prm.status) atomicOr(prm.status + i, (uint32_t)OPTI_KF_ST_ALL_SWING);
    Real *out = resid + (t * IDENT_ROWS) * N + i;
#pragma unroll
    for (int c = 0; c < NX; ++c) out[(long long)c * N] = gt[((t + 1) * NX + c) * S + s] - x[c];
#pragma unroll
    for (int j = 0; j < NZ; ++j) out[(long long)(NX + j) * N] = gt[((t + 1) * NX + sel(j)) * S + s] - z[j];
}

template <typename Real>
__global__ void __launch_bounds__(128) kf_identify_variance_kernel(long long N, long long n_steps, const Real *__restrict__ resid,
                                                                    Real *__restrict__ q_diag /* [12][N] */, Real *__restrict__ r_diag /* [10][N] */) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (component, trajectory), trajectory fastest
    if (idx >= IDENT_ROWS * N) return;
    const long long c = idx / N, i = idx % N;
    double mean = 0.0;
    for (long long t = 0; t < n_steps; ++t) mean += (double)resid[(t * IDENT_ROWS + c) * N + i];
    mean /= (double)n_steps;
    double ss = 0.0;
    for (long long t = 0; t < n_steps; ++t) {
        const double d = (double)resid[(t * IDENT_ROWS + c) * N + i] - mean;
        ss += d * d;
    }
    const Real var = (Real)(ss / (double)n_steps);
    if (c < NX) q_diag[c * N + i] = var;
    else r_diag[(c - NX) * N + i] = var;
}

}  // namespace okf
