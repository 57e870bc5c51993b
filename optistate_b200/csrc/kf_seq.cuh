// Throughput kernel: one thread per trajectory, the whole filter state in registers.
//
//   * P is kept as a packed symmetric matrix (78 scalars) in registers for the entire time loop; x (12) too.
//   * The covariance transition F_d P F_d^T + Q (kalman_filter.py:125-135) is evaluated block-wise on the packed
//     form in the reference's W = F_d P, P' = W F_d^T order, skipping the structural zeros of F_d (168 FMAs).
//   * The update (kalman_filter.py:164-174) folds the 10 measurements in one at a time.  For diagonal R this is
//     the same Schur complement the reference evaluates jointly as P - (P H^T) S^-1 (H P) (block elimination ==
//     successive scalar eliminations), so it needs no 10x10 factorisation, no square roots and no gain matrix:
//     per measurement 1 reciprocal + 101 FMAs.  K_gain = sum_i K[i][i] is recovered from the identity
//     K = P'[:, sel] R^-1.  Parity against the reference's joint form is <= 1e-12 (tests/test_parity_gpu.py).
//   * Independent trajectories => no shuffles, no shared-memory traffic in the recursion, no redundant work.
//
// Per-step inputs are read with coalesced loads from the [T][C][S] layout (stream index fastest).
#pragma once

#include "kf_common.cuh"

namespace okf {

__host__ __device__ constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
constexpr int NP = 78;
constexpr int SEQ_NOISE_ROWS = 32;  // shared memory rows per thread: q[12] r[10] 1/r[10]

// P <- F_d P F_d^T + diag(q), F_d = I + dt N, N[a,c] = R^T, N[b,d] = I   (blocks a=0..2 b=3..5 c=6..8 d=9..11)
// Written with explicit fma_ so that double, float and the packed F2 type run the same operation sequence.
template <typename Real, typename Scalar>
__device__ __forceinline__ void cov_predict_sym(Real (&P)[NP], const Real (&R)[9], Scalar dt_s, const Real *q, int qs) {
    constexpr int a = 0, b = 3, c = 6, d = 9;
    const Real dt = Real(dt_s);
    Real A[9];  // A = dt R^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) A[3 * i + k] = dt * R[3 * k + i];
        // W rows that feed P'[a,a], P'[b,a], P'[b,b] use the OLD c- and d-rows: do them first.
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            Real s = P[tri(a + i, a + j)];
#pragma unroll
            for (int k = 0; k < 3; ++k) s = fma_(A[3 * i + k], P[tri(c + k, a + j)], s);
            P[tri(a + i, a + j)] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) P[tri(b + i, a + j)] = fma_(dt, P[tri(d + i, a + j)], P[tri(b + i, a + j)]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) P[tri(b + i, b + j)] = fma_(dt, P[tri(d + i, b + j)], P[tri(b + i, b + j)]);
        // P'[c,a] = P[c,a] + P[c,c] A^T ; P'[d,a] = P[d,a] + P[d,c] A^T ; P'[c,b] = P[c,b] + dt P[c,d] ; P'[d,b] = P[d,b] + dt P[d,d]
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            Real s = P[tri(c + i, a + j)], u = P[tri(d + i, a + j)];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                s = fma_(P[tri(c + i, c + k)], A[3 * j + k], s);
                u = fma_(P[tri(d + i, c + k)], A[3 * j + k], u);
            }
            P[tri(c + i, a + j)] = s;
            P[tri(d + i, a + j)] = u;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            P[tri(c + i, b + j)] = fma_(dt, P[tri(d + j, c + i)], P[tri(c + i, b + j)]);
            P[tri(d + i, b + j)] = fma_(dt, P[tri(d + i, d + j)], P[tri(d + i, b + j)]);
        }
        // second factor: + W[a,c] A^T, + W[b,c] A^T, + dt W[b,d], with W[.,c] = P'[c,.]^T and W[b,d] = P'[d,b]^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            Real s = P[tri(a + i, a + j)];
#pragma unroll
            for (int k = 0; k < 3; ++k) s = fma_(P[tri(c + k, a + i)], A[3 * j + k], s);
            P[tri(a + i, a + j)] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            Real s = P[tri(b + i, a + j)];
#pragma unroll
            for (int k = 0; k < 3; ++k) s = fma_(P[tri(c + k, b + i)], A[3 * j + k], s);
            P[tri(b + i, a + j)] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) P[tri(b + i, b + j)] = fma_(dt, P[tri(d + j, b + i)], P[tri(b + i, b + j)]);
#pragma unroll
    for (int i = 0; i < NX; ++i) P[tri(i, i)] += q[i * qs];
}

// One scalar measurement z_j of state k = sel(j):  s = P_kk + r,  x += P[:,k] (z_j - x_k)/s,  P -= P[:,k] P[k,:]/s.
// Row/column k of the result is P[:,k] * (r/s), so the old column stays in place until the end (no copy).
template <int J, typename Real>
__device__ __forceinline__ void fold_measurement(Real (&P)[NP], Real (&x)[NX], Real zj, Real rj, Real &nis,
                                                 uint32_t &status) {
    constexpr int k = sel(J);
    const Real s = P[tri(k, k)] + rj;
    if (!(s > Real(0)) || !(s < Real(3e38))) status |= OPTI_KF_ST_NOT_PD;
    const Real inv = Real(1) / s;
    const Real y = zj - x[k];
    const Real g = inv * y;
    nis += y * g;
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] += P[tri(i, k)] * g;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        if (i == k) continue;
        const Real w = P[tri(i, k)] * inv;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            if (j == k) continue;
            P[tri(i, j)] -= w * P[tri(j, k)];
        }
    }
    const Real cfac = rj * inv;
#pragma unroll
    for (int i = 0; i < NX; ++i) P[tri(i, k)] *= cfac;
}

template <typename Real, bool kSummary>
__global__ void __launch_bounds__(128) kf_seq_kernel(const __grid_constant__ Params<Real> prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Real *noise = reinterpret_cast<Real *>(smem_raw);  // [SEQ_NOISE_ROWS][blockDim.x]
    const int nt = blockDim.x;
    const long long i = (long long)blockIdx.x * nt + threadIdx.x;
    if (i >= prm.N) return;
    const long long N = prm.N, S = prm.S;
    const long long s = stream_of(prm, i);
    Real *q = noise + threadIdx.x, *r = noise + 12 * nt + threadIdx.x, *rinv = noise + 22 * nt + threadIdx.x;
#pragma unroll
    for (int c = 0; c < NX; ++c) q[c * nt] = prm.q_kind == OPTI_KF_MAT_DIAG ? prm.Q[c] : prm.Q[c * N + i];
#pragma unroll
    for (int c = 0; c < NZ; ++c) {
        const Real rv = prm.r_kind == OPTI_KF_MAT_DIAG ? prm.R[c] : prm.R[c * N + i];
        r[c * nt] = rv;
        rinv[c * nt] = Real(1) / rv;
    }

    Real x[NX], P[NP];
#pragma unroll
    for (int c = 0; c < NX; ++c) x[c] = prm.x0[c * prm.x0_ld + i * prm.x0_inc];
#pragma unroll
    for (int a = 0; a < NX; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            Real v;
            switch (prm.p0_kind) {
                case OPTI_KF_MAT_NONE: v = (a == b) ? q[a * nt] : Real(0); break;
                case OPTI_KF_MAT_DIAG: v = (a == b) ? prm.P0[a] : Real(0); break;
                case OPTI_KF_MAT_DIAG_PER: v = (a == b) ? prm.P0[a * N + i] : Real(0); break;
                case OPTI_KF_MAT_DENSE: v = prm.P0[a * NX + b]; break;
                default: v = prm.P0[(long long)(a * NX + b) * N + i]; break;
            }
            P[tri(a, b)] = v;
        }

    uint32_t status = 0;
    Real ptrace = Real(0), kgain = Real(0), ymax = Real(0);
    double acc_truth[kSummary ? NX : 1], acc_nom[kSummary ? NX : 1], acc_nis = 0.0;
    if (kSummary) {
#pragma unroll
        for (int c = 0; c < NX; ++c) { acc_truth[c] = 0.0; acc_nom[c] = 0.0; }
    }
    const bool want_gain = prm.k_gain_steps != nullptr || prm.summary != nullptr;

    for (long long t = 0; t < prm.T; ++t) {
        Real z[NZ], pf[12], ff[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            pf[c] = ld_stream(prm.p + (t * 12 + c) * S + s);
            ff[c] = ld_stream(prm.f + (t * 12 + c) * S + s);
        }
        if (prm.z_in) {
#pragma unroll
            for (int c = 0; c < NZ; ++c) z[c] = ld_stream(prm.z_in + (t * NZ + c) * S + s);
        } else {
            Real imu[6], dp[12], contact[4];
#pragma unroll
            for (int c = 0; c < 6; ++c) imu[c] = ld_stream(prm.imu + (t * 6 + c) * S + s);
#pragma unroll
            for (int c = 0; c < 12; ++c) dp[c] = ld_stream(prm.dp + (t * 12 + c) * S + s);
#pragma unroll
            for (int c = 0; c < 4; ++c) contact[c] = ld_stream(prm.contact + (t * 4 + c) * S + s);
            if (form_measurement(imu, pf, dp, contact, z)) status |= OPTI_KF_ST_ALL_SWING;
        }

        Real Rm[9];
        propagate_mean(prm, x, pf, ff, Rm);
        if (prm.x_model_steps) {
#pragma unroll
            for (int c = 0; c < NX; ++c) st_stream(prm.x_model_steps + (t * NX + c) * N + i, x[c]);
        }
        if (prm.p_world_steps) {
#pragma unroll
            for (int c = 0; c < 12; ++c) st_stream(prm.p_world_steps + (t * 12 + c) * N + i, pf[c]);
        }
        if (prm.z_steps) {
#pragma unroll
            for (int c = 0; c < NZ; ++c) st_stream(prm.z_steps + (t * NZ + c) * N + i, z[c]);
        }

        cov_predict_sym(P, Rm, prm.dt, q, nt);

        Real nis = Real(0);
        fold_measurement<0>(P, x, z[0], r[0 * nt], nis, status);
        fold_measurement<1>(P, x, z[1], r[1 * nt], nis, status);
        fold_measurement<2>(P, x, z[2], r[2 * nt], nis, status);
        fold_measurement<3>(P, x, z[3], r[3 * nt], nis, status);
        fold_measurement<4>(P, x, z[4], r[4 * nt], nis, status);
        fold_measurement<5>(P, x, z[5], r[5 * nt], nis, status);
        fold_measurement<6>(P, x, z[6], r[6 * nt], nis, status);
        fold_measurement<7>(P, x, z[7], r[7 * nt], nis, status);
        fold_measurement<8>(P, x, z[8], r[8 * nt], nis, status);
        fold_measurement<9>(P, x, z[9], r[9 * nt], nis, status);

        ymax = fmax(ymax, nis);
        ptrace = Real(0);
#pragma unroll
        for (int c = 0; c < NX; ++c) ptrace += P[tri(c, c)];
        if (want_gain) {  // K = P'[:, sel] R^-1  =>  K[j][j] = P'[j][sel(j)] / r_j
            kgain = Real(0);
#pragma unroll
            for (int j = 0; j < NZ; ++j) kgain += P[tri(j, sel(j))] * rinv[j * nt];
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
        if (kSummary) {
            acc_nis += (double)nis;
            if (prm.truth) {
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    const double e = (double)x[c] - (double)ld_stream(prm.truth + (t * NX + c) * S + s);
                    acc_truth[c] += e * e;
                }
            }
            if (prm.nominal) {
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    const double e = (double)x[c] - (double)ld_stream(prm.nominal + (t * NX + c) * S + s);
                    acc_nom[c] += e * e;
                }
            }
        }
        if (prm.P_ckpt && prm.ckpt_every > 0 && (t + 1) % prm.ckpt_every == 0) {
            Real *dst = prm.P_ckpt + ((t + 1) / prm.ckpt_every - 1) * (long long)(NX * NX) * N + i;
#pragma unroll
            for (int a = 0; a < NX; ++a)
#pragma unroll
                for (int b = 0; b < NX; ++b) dst[(long long)(a * NX + b) * N] = P[tri(a, b)];
        }
    }

    if (prm.x_final) {
#pragma unroll
        for (int c = 0; c < NX; ++c) prm.x_final[c * N + i] = x[c];
    }
    if (prm.P_final) {
#pragma unroll
        for (int a = 0; a < NX; ++a)
#pragma unroll
            for (int b = 0; b < NX; ++b) prm.P_final[(long long)(a * NX + b) * N + i] = P[tri(a, b)];
    }
    if (kSummary && prm.summary) {
        Real *sm = prm.summary + i;
        const double invT = prm.T > 0 ? 1.0 / (double)prm.T : 0.0;
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            sm[(long long)c * N] = x[c];
            sm[(long long)(12 + c) * N] = P[tri(c, c)];
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
