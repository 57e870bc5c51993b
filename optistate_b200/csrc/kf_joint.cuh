// General kernel: the reference's joint update in the reference's operand order, on a full (never symmetrised)
// 12x12 P, with dense Q / R / P0, the predict_mpc covariance model, selectable phases and the gain matrix K as
// an output.  One thread per trajectory; P, Q, R, the Cholesky factor and K live in per-thread arrays (local
// memory, L1-resident).  This is the semantic workhorse behind the Kalman_Filter class and the fallback of the
// batched entry for anything the register-resident sequential kernel (kf_seq.cuh) does not cover.
//
//   predict   kalman_filter.py:119-138 (model 0) / :153-161 (model 1), mean by force_controller.py:269-291
//   update    kalman_filter.py:164-174:  y = z - H x ; S = H P H^T + R ; K = (P H^T) S^-1 ; x += K y ;
//             P <- (I - K H) P = P - K P[sel,:]   (columns of P for K, rows of P for the product - SURVEY 0.3)
//   S^-1 is applied through a Cholesky factorisation S = L L^T (rows of K solved by forward + back substitution).
#pragma once

#include "kf_common.cuh"

namespace okf {

template <typename Real>
__device__ __noinline__ void cov_predict_dense(Real *P, const Real *Rm, const Real *Q, Real dt, int model) {
    Real Fd[NX * NX], W[NX * NX];
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NX; ++j) {
            Real fij = Real(0);
            if (i < 3 && j >= 6 && j < 9) fij = Rm[3 * (j - 6) + i];  // F[0:3,6:9] = R^T
            if (i >= 3 && i < 6 && j == i + 6) fij = Real(1);         // F[3:6,9:12] = I
            Fd[i * NX + j] = model == OPTI_KF_COV_MPC ? (Real)exp((double)(dt * fij)) : ((i == j) ? Real(1) : Real(0)) + dt * fij;
        }
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NX; ++j) {
            Real s = Real(0);
            for (int k = 0; k < NX; ++k) s += Fd[i * NX + k] * P[k * NX + j];
            W[i * NX + j] = s;
        }
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NX; ++j) {
            Real s = Real(0);
            for (int k = 0; k < NX; ++k) s += W[i * NX + k] * Fd[j * NX + k];
            P[i * NX + j] = s + Q[i * NX + j];
        }
}

template <typename Real>
__device__ __noinline__ void joint_update(Real *x, Real *P, const Real *z, const Real *Rn, Real *K, Real &ptrace,
                                          Real &kgain, Real &nis, uint32_t &status) {
    Real L[NZ * NZ], y[NZ];
    for (int i = 0; i < NZ; ++i) {
        y[i] = z[i] - x[sel(i)];
        for (int j = 0; j < NZ; ++j) L[i * NZ + j] = P[sel(i) * NX + sel(j)] + Rn[i * NZ + j];
    }
    // Visible asymmetry of S (a user-supplied non-symmetric P0 / Q / R; rounding-level asymmetry is ~1e-16): a
    // Cholesky factorisation would silently symmetrise it, so that rare case takes the pivoted-LU inverse the
    // reference itself uses (np.linalg.inv, kalman_filter.py:168) and is flagged in the status word.
    bool asym = false;
    for (int i = 0; i < NZ; ++i)
        for (int j = 0; j < i; ++j) {
            const Real d = fabs(L[i * NZ + j] - L[j * NZ + i]);
            if (d > Real(sizeof(Real) == 8 ? 1e-12 : 1e-5) * sqrt(fabs(L[i * NZ + i] * L[j * NZ + j]))) asym = true;
        }
    if (asym) {
        status |= OPTI_KF_ST_ASYMMETRIC;
        Real A[NZ * 2 * NZ];  // [S | I] -> Gauss-Jordan with partial pivoting -> [U | L^-1 P], then back substitution
        for (int i = 0; i < NZ; ++i)
            for (int j = 0; j < NZ; ++j) { A[i * 2 * NZ + j] = L[i * NZ + j]; A[i * 2 * NZ + NZ + j] = (i == j) ? Real(1) : Real(0); }
        for (int c = 0; c < NZ; ++c) {
            int piv = c;
            for (int rr = c + 1; rr < NZ; ++rr)
                if (fabs(A[rr * 2 * NZ + c]) > fabs(A[piv * 2 * NZ + c])) piv = rr;
            if (A[piv * 2 * NZ + c] == Real(0) || !isfinite(A[piv * 2 * NZ + c])) status |= OPTI_KF_ST_NOT_PD;
            if (piv != c)
                for (int j = 0; j < 2 * NZ; ++j) { const Real tmp = A[c * 2 * NZ + j]; A[c * 2 * NZ + j] = A[piv * 2 * NZ + j]; A[piv * 2 * NZ + j] = tmp; }
            for (int rr = c + 1; rr < NZ; ++rr) {
                const Real m = A[rr * 2 * NZ + c] / A[c * 2 * NZ + c];
                for (int j = c; j < 2 * NZ; ++j) A[rr * 2 * NZ + j] -= m * A[c * 2 * NZ + j];
            }
        }
        Real *Sinv = L;  // S itself is no longer needed
        for (int c = NZ - 1; c >= 0; --c)
            for (int j = 0; j < NZ; ++j) {
                Real s = A[c * 2 * NZ + NZ + j];
                for (int k = c + 1; k < NZ; ++k) s -= A[c * 2 * NZ + k] * Sinv[k * NZ + j];
                Sinv[c * NZ + j] = s / A[c * 2 * NZ + c];
            }
        nis = Real(0);
        for (int i = 0; i < NZ; ++i) {
            Real s = Real(0);
            for (int j = 0; j < NZ; ++j) s += Sinv[i * NZ + j] * y[j];
            nis += y[i] * s;
        }
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NZ; ++j) {
                Real s = Real(0);
                for (int k = 0; k < NZ; ++k) s += P[i * NX + sel(k)] * Sinv[k * NZ + j];
                K[i * NZ + j] = s;
            }
    } else {
    // Cholesky, in place in the lower triangle; Linv_d[j] = 1 / L[j][j]
    Real dinv[NZ];
    for (int j = 0; j < NZ; ++j) {
        Real d = L[j * NZ + j];
        for (int k = 0; k < j; ++k) d -= L[j * NZ + k] * L[j * NZ + k];
        if (!(d > Real(0)) || !(d < Real(3e38))) status |= OPTI_KF_ST_NOT_PD;
        const Real ljj = sqrt(d);
        L[j * NZ + j] = ljj;
        dinv[j] = Real(1) / ljj;
        for (int i = j + 1; i < NZ; ++i) {
            Real s = L[i * NZ + j];
            for (int k = 0; k < j; ++k) s -= L[i * NZ + k] * L[j * NZ + k];
            L[i * NZ + j] = s * dinv[j];
        }
    }
    // NIS = y^T S^-1 y = |L^-1 y|^2
    {
        Real u[NZ];
        nis = Real(0);
        for (int j = 0; j < NZ; ++j) {
            Real s = y[j];
            for (int k = 0; k < j; ++k) s -= L[j * NZ + k] * u[k];
            u[j] = s * dinv[j];
            nis += u[j] * u[j];
        }
    }
    // K[i,:] = P[i,sel] S^-1  (S symmetric): solve L u = P[i,sel]^T, then L^T k = u
    for (int i = 0; i < NX; ++i) {
        Real u[NZ];
        for (int j = 0; j < NZ; ++j) {
            Real s = P[i * NX + sel(j)];
            for (int k = 0; k < j; ++k) s -= L[j * NZ + k] * u[k];
            u[j] = s * dinv[j];
        }
        for (int j = NZ - 1; j >= 0; --j) {
            Real s = u[j];
            for (int k = j + 1; k < NZ; ++k) s -= L[k * NZ + j] * K[i * NZ + k];
            K[i * NZ + j] = s * dinv[j];
        }
    }
    }  // Cholesky path
    for (int i = 0; i < NX; ++i) {
        Real s = Real(0);
        for (int j = 0; j < NZ; ++j) s += K[i * NZ + j] * y[j];
        x[i] += s;
    }
    // P <- P - K P[sel,:]  (old rows of P: copy them first)
    Real Ps[NZ * NX];
    for (int j = 0; j < NZ; ++j)
        for (int c = 0; c < NX; ++c) Ps[j * NX + c] = P[sel(j) * NX + c];
    ptrace = Real(0);
    for (int i = 0; i < NX; ++i)
        for (int c = 0; c < NX; ++c) {
            Real s = Real(0);
            for (int j = 0; j < NZ; ++j) s += K[i * NZ + j] * Ps[j * NX + c];
            P[i * NX + c] -= s;
            if (i == c) ptrace += P[i * NX + c];
        }
    kgain = Real(0);
    for (int j = 0; j < NZ; ++j) kgain += K[j * NZ + j];
}

template <typename Real>
__global__ void __launch_bounds__(64) kf_joint_kernel(const __grid_constant__ Params<Real> prm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.N) return;
    const long long N = prm.N, S = prm.S;
    const long long s = stream_of(prm, i);

    Real x[NX], P[NX * NX], Q[NX * NX], Rn[NZ * NZ], K[NX * NZ];
    for (int c = 0; c < NX * NZ; ++c) K[c] = Real(0);
    for (int a = 0; a < NX; ++a)
        for (int b = 0; b < NX; ++b) {
            Real v;
            switch (prm.q_kind) {
                case OPTI_KF_MAT_DIAG: v = (a == b) ? prm.Q[a] : Real(0); break;
                case OPTI_KF_MAT_DIAG_PER: v = (a == b) ? prm.Q[a * N + i] : Real(0); break;
                case OPTI_KF_MAT_DENSE: v = prm.Q[a * NX + b]; break;
                default: v = prm.Q[(long long)(a * NX + b) * N + i]; break;
            }
            Q[a * NX + b] = v;
        }
    for (int a = 0; a < NZ; ++a)
        for (int b = 0; b < NZ; ++b) {
            Real v;
            switch (prm.r_kind) {
                case OPTI_KF_MAT_DIAG: v = (a == b) ? prm.R[a] : Real(0); break;
                case OPTI_KF_MAT_DIAG_PER: v = (a == b) ? prm.R[a * N + i] : Real(0); break;
                case OPTI_KF_MAT_DENSE: v = prm.R[a * NZ + b]; break;
                default: v = prm.R[(long long)(a * NZ + b) * N + i]; break;
            }
            Rn[a * NZ + b] = v;
        }
    for (int a = 0; a < NX; ++a)
        for (int b = 0; b < NX; ++b) {
            Real v;
            switch (prm.p0_kind) {
                case OPTI_KF_MAT_NONE: v = Q[a * NX + b]; break;
                case OPTI_KF_MAT_DIAG: v = (a == b) ? prm.P0[a] : Real(0); break;
                case OPTI_KF_MAT_DIAG_PER: v = (a == b) ? prm.P0[a * N + i] : Real(0); break;
                case OPTI_KF_MAT_DENSE: v = prm.P0[a * NX + b]; break;
                default: v = prm.P0[(long long)(a * NX + b) * N + i]; break;
            }
            P[a * NX + b] = v;
        }
#pragma unroll
    for (int c = 0; c < NX; ++c) x[c] = prm.x0[c * prm.x0_ld + i * prm.x0_inc];

    uint32_t status = 0;
    Real ptrace = Real(0), kgain = Real(0), ymax = Real(0);
    for (int c = 0; c < NX; ++c) ptrace += P[c * NX + c];
    double acc_truth[NX], acc_nom[NX], acc_nis = 0.0;
    for (int c = 0; c < NX; ++c) { acc_truth[c] = 0.0; acc_nom[c] = 0.0; }

    for (long long t = 0; t < prm.T; ++t) {
        Real z[NZ];
        Real pf[12], ff[12];
        const bool need_p = (prm.phases & (OPTI_KF_PHASE_MEASURE | OPTI_KF_PHASE_PREDICT)) != 0;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            pf[c] = need_p ? ld_stream(prm.p + (t * 12 + c) * S + s) : Real(0);
            ff[c] = (prm.phases & OPTI_KF_PHASE_PREDICT) ? ld_stream(prm.f + (t * 12 + c) * S + s) : Real(0);
        }
        if (prm.phases & OPTI_KF_PHASE_MEASURE) {
            Real imu[6], dp[12], contact[4];
#pragma unroll
            for (int c = 0; c < 6; ++c) imu[c] = ld_stream(prm.imu + (t * 6 + c) * S + s);
#pragma unroll
            for (int c = 0; c < 12; ++c) dp[c] = ld_stream(prm.dp + (t * 12 + c) * S + s);
#pragma unroll
            for (int c = 0; c < 4; ++c) contact[c] = ld_stream(prm.contact + (t * 4 + c) * S + s);
            if (form_measurement(imu, pf, dp, contact, z)) status |= OPTI_KF_ST_ALL_SWING;
        } else if (prm.z_in) {
#pragma unroll
            for (int c = 0; c < NZ; ++c) z[c] = ld_stream(prm.z_in + (t * NZ + c) * S + s);
        } else {
#pragma unroll
            for (int c = 0; c < NZ; ++c) z[c] = Real(0);
        }
        if (prm.z_steps) {
#pragma unroll
            for (int c = 0; c < NZ; ++c) st_stream(prm.z_steps + (t * NZ + c) * N + i, z[c]);
        }

        if (prm.phases & OPTI_KF_PHASE_PREDICT) {
            Real Rm[9];
            if (prm.cov_model == OPTI_KF_COV_MPC) {
                // predict_mpc propagates the covariance first, with R from the reference body angles
                // (kalman_filter.py:153-158), then the mean with R from the state (:161)
                Real Rb[9];
                rot_zyx(ld_stream(prm.body_ref + (t * 12 + 0) * S + s), ld_stream(prm.body_ref + (t * 12 + 1) * S + s),
                        ld_stream(prm.body_ref + (t * 12 + 2) * S + s), Rb);
                cov_predict_dense(P, Rb, Q, prm.dt, OPTI_KF_COV_MPC);
                propagate_mean(prm, x, pf, ff, Rm);
            } else {
                propagate_mean(prm, x, pf, ff, Rm);
                cov_predict_dense(P, Rm, Q, prm.dt, OPTI_KF_COV_PREDICT);
            }
            ptrace = Real(0);
            for (int c = 0; c < NX; ++c) ptrace += P[c * NX + c];
            if (prm.x_model_steps) {
#pragma unroll
                for (int c = 0; c < NX; ++c) st_stream(prm.x_model_steps + (t * NX + c) * N + i, x[c]);
            }
            if (prm.p_world_steps) {
#pragma unroll
                for (int c = 0; c < 12; ++c) st_stream(prm.p_world_steps + (t * 12 + c) * N + i, pf[c]);
            }
        }

        Real nis = Real(0);
        if (prm.phases & OPTI_KF_PHASE_UPDATE) {
            joint_update(x, P, z, Rn, K, ptrace, kgain, nis, status);
            ymax = fmax(ymax, nis);
        }
        bool fin = true;
#pragma unroll
        for (int c = 0; c < NX; ++c) fin &= isfinite(x[c]);
        if (!fin) status |= OPTI_KF_ST_NONFINITE;

        if (prm.x_steps) {
#pragma unroll
            for (int c = 0; c < NX; ++c) st_stream(prm.x_steps + (t * NX + c) * N + i, x[c]);
        }
        if (prm.p_trace_steps) st_stream(prm.p_trace_steps + t * N + i, ptrace);
        if (prm.k_gain_steps) st_stream(prm.k_gain_steps + t * N + i, kgain);
        if (prm.nis_steps) st_stream(prm.nis_steps + t * N + i, nis);
        if (prm.summary) {
            acc_nis += (double)nis;
            if (prm.truth)
                for (int c = 0; c < NX; ++c) {
                    const double e = (double)x[c] - (double)ld_stream(prm.truth + (t * NX + c) * S + s);
                    acc_truth[c] += e * e;
                }
            if (prm.nominal)
                for (int c = 0; c < NX; ++c) {
                    const double e = (double)x[c] - (double)ld_stream(prm.nominal + (t * NX + c) * S + s);
                    acc_nom[c] += e * e;
                }
        }
        if (prm.P_ckpt && prm.ckpt_every > 0 && (t + 1) % prm.ckpt_every == 0) {
            Real *dst = prm.P_ckpt + ((t + 1) / prm.ckpt_every - 1) * (long long)(NX * NX) * N + i;
            for (int c = 0; c < NX * NX; ++c) dst[(long long)c * N] = P[c];
        }
    }

    if (prm.x_final)
        for (int c = 0; c < NX; ++c) prm.x_final[c * N + i] = x[c];
    if (prm.P_final)
        for (int c = 0; c < NX * NX; ++c) prm.P_final[(long long)c * N + i] = P[c];
    if (prm.K_final)
        for (int c = 0; c < NX * NZ; ++c) prm.K_final[(long long)c * N + i] = K[c];
    if (prm.summary) {
        Real *sm = prm.summary + i;
        const double invT = prm.T > 0 ? 1.0 / (double)prm.T : 0.0;
        for (int c = 0; c < NX; ++c) {
            sm[(long long)c * N] = x[c];
            sm[(long long)(12 + c) * N] = P[c * NX + c];
            sm[(long long)(24 + c) * N] = (Real)sqrt(acc_truth[c] * invT);
            sm[(long long)(36 + c) * N] = (Real)sqrt(acc_nom[c] * invT);
        }
        sm[48LL * N] = (Real)(acc_nis * invT);
        sm[49LL * N] = ptrace;
        sm[50LL * N] = kgain;
        sm[51LL * N] = (Real)sqrt((double)ymax);
    }
    if (prm.status) prm.status[i] = status;
}

}  // namespace okf
