// General kernel, cooperative version: the reference's joint update in the reference's operand order on a full (never
// symmetrised) 12x12 P, dense Q / R / P0, both covariance models, selectable phases, K as an output - with FOUR lanes
// per trajectory.
//
//   * Lane q of a trajectory's quad owns rows 3q..3q+2 of P (36 scalars) and of K (30 scalars) in registers; the state x
//     is replicated in the four lanes.  32 trajectories per 128-thread block; a quad never spans two warps, so all
//     synchronisation is quad-wide __syncwarp(mask) and shuffles.
//   * Row operations are local, column/row exchanges go through shuffles (F_d P needs rows 6..11 in lanes 0/1, the state
//     update is an all-gather of 3 entries per lane) or through a per-trajectory shared-memory copy of P (the old rows
//     P[sel,:] that P - K (H P) needs, S = H P H^T + R, and the dense predict_mpc transition).
//   * S = L L^T is factorised cooperatively in shared memory, column by column (each lane takes every fourth row below the
//     diagonal); the rows of K = (P H^T) S^-1 are then solved per lane by forward / back substitution with every L entry
//     loaded once for the lane's three rows.  A visibly asymmetric S (user-supplied non-symmetric P0 / Q / R) takes the
//     pivoted Gauss-Jordan inverse instead - what np.linalg.inv does in the reference (kalman_filter.py:168).
//
//   predict   kalman_filter.py:119-138 (model 0) / :153-161 (model 1), mean by force_controller.py:269-291
//   update    kalman_filter.py:164-174:  y = z - H x ; S = H P H^T + R ; K = (P H^T) S^-1 ; x += K y ;
//             P <- (I - K H) P = P - K P[sel,:]   (columns of P for K, rows of P for the product - SURVEY 0.3)
#pragma once

#include "kf_common.cuh"

namespace okf {

constexpr int JC_THREADS = 128;
constexpr int JC_TRAJ = JC_THREADS / 4;  // trajectories per block
constexpr int JC_LD = JC_TRAJ + 1;       // padded trajectory stride of the shared arrays: the four lanes of a quad often touch
                                         // different elements of the same trajectory, which an unpadded stride maps to one bank
constexpr int JC_VEC = 12;               // per-trajectory scratch: dinv[10]
template <typename Real>
constexpr size_t jc_smem_bytes() { return (size_t)(NX * NX + NZ * NZ + JC_VEC) * JC_LD * sizeof(Real); }

// Quads are independent of each other (one may take the pivoted-LU path while its neighbours do not), so every
// shuffle and every barrier names only the four lanes of the quad.
__device__ __forceinline__ unsigned quad_mask() { return 0xFu << ((threadIdx.x & 31) & ~3); }
template <typename T> __device__ __forceinline__ T quad_get(T v, int lane, int src_q) { return __shfl_sync(quad_mask(), v, (lane & ~3) | src_q); }
template <typename T> __device__ __forceinline__ T quad_sum(T v) {
    v += __shfl_xor_sync(quad_mask(), v, 1);
    v += __shfl_xor_sync(quad_mask(), v, 2);
    return v;
}
__device__ __forceinline__ uint32_t quad_or(uint32_t v) {
    v |= __shfl_xor_sync(quad_mask(), v, 1);
    v |= __shfl_xor_sync(quad_mask(), v, 2);
    return v;
}
__device__ __forceinline__ void quad_sync() { __syncwarp(quad_mask()); }

template <typename Real>
__device__ __forceinline__ Real mat_at(const Real *M, int kind, int n, int a, int b, long long N, long long i) {
    switch (kind) {
        case OPTI_KF_MAT_DIAG: return a == b ? M[a] : Real(0);
        case OPTI_KF_MAT_DIAG_PER: return a == b ? M[a * N + i] : Real(0);
        case OPTI_KF_MAT_DENSE: return M[a * n + b];
        default: return M[(long long)(a * n + b) * N + i];
    }
}

// lane q's three entries 3q..3q+2 of a replicated 12-vector, without dynamic register indexing
template <typename Real>
__device__ __forceinline__ void own3(const Real (&v)[NX], int q, Real (&o)[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = q == 0 ? v[a] : q == 1 ? v[3 + a] : q == 2 ? v[6 + a] : v[9 + a];
}

// Diagonal entry P[3q+a][3q+a] of lane q's row a.  Written as a select chain over compile-time indices: a loop of the
// form "sum_c (c == 3q+a) ? Pr[a][c] : 0" is recognised by the compiler as Pr[a][3q+a], a dynamic register index that
// would push the whole array into local memory.
template <typename Real>
__device__ __forceinline__ Real own_diag(const Real (&Pr)[3][NX], int q, int a) {
    return q == 0 ? Pr[a][a] : q == 1 ? Pr[a][3 + a] : q == 2 ? Pr[a][6 + a] : Pr[a][9 + a];
}
template <typename Real>
__device__ __forceinline__ Real own_trace(const Real (&Pr)[3][NX], int q) {
    return own_diag(Pr, q, 0) + own_diag(Pr, q, 1) + own_diag(Pr, q, 2);
}
template <typename Real>
__device__ __forceinline__ void add_own_diag(Real (&Pr)[3][NX], int q, int a, Real v) {
    Pr[a][a] += q == 0 ? v : Real(0);
    Pr[a][3 + a] += q == 1 ? v : Real(0);
    Pr[a][6 + a] += q == 2 ? v : Real(0);
    Pr[a][9 + a] += q == 3 ? v : Real(0);
}

// noise matrices without per-element switches: element e of a dense matrix / entry a of a diagonal lives at
// base[e * stride + offset] with (stride, offset) = (1, 0) for a shared array and (N, i) for a per-trajectory one
template <typename Real>
struct NoiseView {
    const Real *base;
    long long stride, offset;
    bool dense;
    __device__ __forceinline__ NoiseView(const Real *m, int kind, long long N, long long i)
        : base(m), stride((kind == OPTI_KF_MAT_DIAG_PER || kind == OPTI_KF_MAT_DENSE_PER) ? N : 1),
          offset((kind == OPTI_KF_MAT_DIAG_PER || kind == OPTI_KF_MAT_DENSE_PER) ? i : 0),
          dense(kind == OPTI_KF_MAT_DENSE || kind == OPTI_KF_MAT_DENSE_PER) {}
    __device__ __forceinline__ Real at(int e) const { return __ldg(base + e * stride + offset); }
};

// in-place inverse of the 10x10 matrix at lm_ (element (i,j) at lm_[(i*10+j)*JC_LD]) by Gauss-Jordan with partial pivoting
template <typename Real>
__device__ __noinline__ bool invert10_inplace(Real *lm_) {
    int piv[NZ];
    bool singular = false;
    for (int c = 0; c < NZ; ++c) {
        int p = c;
        Real best = fabs(lm_[(c * NZ + c) * JC_LD]);
        for (int r = c + 1; r < NZ; ++r) {
            const Real v = fabs(lm_[(r * NZ + c) * JC_LD]);
            if (v > best) { best = v; p = r; }
        }
        piv[c] = p;
        if (!(best > Real(0)) || !isfinite(best)) singular = true;
        if (p != c)
            for (int j = 0; j < NZ; ++j) {
                const Real t = lm_[(c * NZ + j) * JC_LD];
                lm_[(c * NZ + j) * JC_LD] = lm_[(p * NZ + j) * JC_LD];
                lm_[(p * NZ + j) * JC_LD] = t;
            }
        const Real pinv = Real(1) / lm_[(c * NZ + c) * JC_LD];
        lm_[(c * NZ + c) * JC_LD] = Real(1);
        for (int j = 0; j < NZ; ++j) lm_[(c * NZ + j) * JC_LD] *= pinv;
        for (int r = 0; r < NZ; ++r) {
            if (r == c) continue;
            const Real f = lm_[(r * NZ + c) * JC_LD];
            lm_[(r * NZ + c) * JC_LD] = Real(0);
            for (int j = 0; j < NZ; ++j) lm_[(r * NZ + j) * JC_LD] -= f * lm_[(c * NZ + j) * JC_LD];
        }
    }
    for (int c = NZ - 1; c >= 0; --c)
        if (piv[c] != c)
            for (int r = 0; r < NZ; ++r) {
                const Real t = lm_[(r * NZ + c) * JC_LD];
                lm_[(r * NZ + c) * JC_LD] = lm_[(r * NZ + piv[c]) * JC_LD];
                lm_[(r * NZ + piv[c]) * JC_LD] = t;
            }
    return singular;
}

template <typename Real>
__global__ void __launch_bounds__(JC_THREADS, sizeof(Real) == 4 ? 3 : 2) kf_joint_coop_kernel(const __grid_constant__ Params<Real> prm) {
    extern __shared__ __align__(16) unsigned char jc_raw[];
    const int tid = threadIdx.x, lane = tid & 31, q = tid & 3, tj = tid >> 2;
    const int r0 = 3 * q;  // first row owned by this lane
    const long long N = prm.N, S = prm.S;
    const long long i_raw = (long long)blockIdx.x * JC_TRAJ + tj;
    const bool active = i_raw < N;
    const long long i = active ? i_raw : N - 1;
    const long long s = stream_of(prm, i);
    Real *pf_ = reinterpret_cast<Real *>(jc_raw) + tj;   // full P of this trajectory: element e at pf_[e * JC_LD]
    Real *lm_ = pf_ + NX * NX * JC_LD;                    // S, then its Cholesky factor / inverse (10x10)
    Real *vc_ = lm_ + NZ * NZ * JC_LD;                    // dinv[10]
#define PF(e) pf_[(e) * JC_LD]
#define LM(e) lm_[(e) * JC_LD]
#define VC(e) vc_[(e) * JC_LD]

    const NoiseView<Real> Qv(prm.Q, prm.q_kind, N, i), Rv(prm.R, prm.r_kind, N, i);
    Real Pr[3][NX], x[NX];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = 0; b < NX; ++b) {
            const int row = r0 + a;
            Pr[a][b] = prm.p0_kind == OPTI_KF_MAT_NONE ? mat_at(prm.Q, prm.q_kind, NX, row, b, N, i)
                                                        : mat_at(prm.P0, prm.p0_kind, NX, row, b, N, i);
        }
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) x[c] = prm.x0[c * prm.x0_ld + i * prm.x0_inc];

    uint32_t status = 0;
    Real ptrace = Real(0), kgain = Real(0), ymax = Real(0);
    ptrace = quad_sum(own_trace(Pr, q));
    double acc_truth[3] = {0.0, 0.0, 0.0}, acc_nom[3] = {0.0, 0.0, 0.0}, acc_nis = 0.0;

    for (long long t = 0; t < prm.T; ++t) {
        Real z[NZ], pf[12], ff[12];
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
        if (active && prm.z_steps) {
#pragma unroll
            for (int c = 0; c < NZ; ++c)
                if ((c & 3) == q) st_stream(prm.z_steps + (t * NZ + c) * N + i, z[c]);
        }

        if (prm.phases & OPTI_KF_PHASE_PREDICT) {
            Real Rm[9];
            if (prm.cov_model == OPTI_KF_COV_MPC) {
                // predict_mpc: covariance first, F_d = exp(dt F) element-wise (dense, all entries ~1) with R from the
                // reference body angles (kalman_filter.py:153-158); then the mean with R from the state (:161)
                Real Rb[9];
                rot_zyx(ld_stream(prm.body_ref + (t * 12 + 0) * S + s), ld_stream(prm.body_ref + (t * 12 + 1) * S + s),
                        ld_stream(prm.body_ref + (t * 12 + 2) * S + s), Rb);
                Real E[9];  // E[3*i + k] = F_d[i][6+k] = exp(dt R^T[i][k]) = exp(dt Rb[k][i])
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int k = 0; k < 3; ++k) E[3 * a + k] = (Real)exp((double)(prm.dt * Rb[3 * k + a]));
                const Real e1 = (Real)exp((double)prm.dt);
                auto fd = [&](int r, int m) -> Real {  // F_d[r][m]
                    if (r < 3 && m >= 6 && m < 9) return E[3 * r + (m - 6)];
                    if (r >= 3 && r < 6 && m == r + 6) return e1;
                    return Real(1);
                };
                // F_d[3q + a][k] of this lane's own rows, with compile-time indices into E (rows 0..2 live in lane 0,
                // rows 3..5 in lane 1, the other rows of F_d are all ones)
                auto fd_own = [&](int a, int k) -> Real {
                    Real v = Real(1);
                    if (k >= 6 && k < 9) v = q == 0 ? E[3 * a + (k - 6)] : Real(1);
                    if (k == 9 + a) v = q == 1 ? e1 : v;
                    return v;
                };
                quad_sync();
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < NX; ++c) PF((r0 + a) * NX + c) = Pr[a][c];
                quad_sync();
                Real W[3][NX];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < NX; ++c) {
                        Real acc = Real(0);
#pragma unroll
                        for (int k = 0; k < NX; ++k) acc += fd_own(a, k) * PF(k * NX + c);
                        W[a][c] = acc;
                    }
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < NX; ++c) {
                        Real acc = Real(0);
#pragma unroll
                        for (int m = 0; m < NX; ++m) acc += W[a][m] * fd(c, m);
                        Pr[a][c] = acc + (Qv.dense ? Qv.at((r0 + a) * NX + c) : Real(0));
                    }
                if (!Qv.dense) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) add_own_diag(Pr, q, a, Qv.at(r0 + a));
                }
                propagate_mean(prm, x, pf, ff, Rm);
            } else {
                propagate_mean(prm, x, pf, ff, Rm);
                // W = F_d P: rows 0..2 += dt R^T rows 6..8 (lane 0 <- lane 2), rows 3..5 += dt rows 9..11 (lane 1 <- lane 3)
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    const Real o0 = quad_get(Pr[0][c], lane, (q + 2) & 3), o1 = quad_get(Pr[1][c], lane, (q + 2) & 3),
                               o2 = quad_get(Pr[2][c], lane, (q + 2) & 3);
                    if (q == 0) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) Pr[a][c] += prm.dt * (Rm[a] * o0 + Rm[3 + a] * o1 + Rm[6 + a] * o2);
                    } else if (q == 1) {
                        Pr[0][c] += prm.dt * o0; Pr[1][c] += prm.dt * o1; Pr[2][c] += prm.dt * o2;
                    }
                }
                // P' = W F_d^T: columns 0..2 += dt W[:,6..8] R, columns 3..5 += dt W[:,9..11]; then + Q (own rows)
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const Real w6 = Pr[a][6], w7 = Pr[a][7], w8 = Pr[a][8];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        Pr[a][j] += prm.dt * (w6 * Rm[j] + w7 * Rm[3 + j] + w8 * Rm[6 + j]);
                        Pr[a][3 + j] += prm.dt * Pr[a][9 + j];
                    }
                    if (Qv.dense) {
#pragma unroll
                        for (int c = 0; c < NX; ++c) Pr[a][c] += Qv.at((r0 + a) * NX + c);
                    } else {
                        add_own_diag(Pr, q, a, Qv.at(r0 + a));
                    }
                }
            }
            ptrace = quad_sum(own_trace(Pr, q));
            if (active && prm.x_model_steps) {
                Real xo[3];
                own3(x, q, xo);
#pragma unroll
                for (int a = 0; a < 3; ++a) st_stream(prm.x_model_steps + (t * NX + r0 + a) * N + i, xo[a]);
            }
            if (active && prm.p_world_steps) {  // lane q stores foot q
                Real po[3];
                own3(pf, q, po);
#pragma unroll
                for (int a = 0; a < 3; ++a) st_stream(prm.p_world_steps + (t * 12 + r0 + a) * N + i, po[a]);
            }
        }

        Real nis = Real(0);
        if (prm.phases & OPTI_KF_PHASE_UPDATE) {
            // all-gather of P (pre-update): S and the old rows P[sel,:] are read from this copy
            quad_sync();
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < NX; ++c) PF((r0 + a) * NX + c) = Pr[a][c];
            quad_sync();
            // S = P[sel,sel] + R: lane q forms rows q, q+4, q+8
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const int a = q + 4 * rr;
                if (a < NZ) {
                    const int sa = a < 3 ? a : a + 2;
#pragma unroll
                    for (int b = 0; b < NZ; ++b) {
                        const Real rn = Rv.dense ? Rv.at(a * NZ + b) : (a == b ? Rv.at(a) : Real(0));
                        LM(a * NZ + b) = PF(sa * NX + sel(b)) + rn;
                    }
                }
            }
            quad_sync();
            uint32_t asym = 0;
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const int a = q + 4 * rr;
                if (a < NZ) {
#pragma unroll
                    for (int b = 0; b < NZ; ++b) {
                        const Real d = fabs(LM(a * NZ + b) - LM(b * NZ + a));
                        if (d > Real(sizeof(Real) == 8 ? 5e-13 : 5e-6) * (fabs(LM(a * NZ + a)) + fabs(LM(b * NZ + b)))) asym = 1;
                    }
                }
            }
            asym = quad_or(asym);
            Real Kr[3][NZ];  // this lane's rows of the gain
            Real y[NZ];
#pragma unroll
            for (int j = 0; j < NZ; ++j) y[j] = z[j] - x[sel(j)];

            // A symmetric S is factorised by Cholesky.  If that breaks down (S indefinite: only a user-supplied non-PSD P / R gets
            // there) the strict lower triangle, which the factorisation has overwritten, is restored from the upper one and the
            // trajectory takes the pivoted inverse like a visibly asymmetric S does - the reference's np.linalg.inv raises for a
            // SINGULAR S only and carries on with an indefinite one (kalman_filter.py:168).
            bool lu = asym != 0;
            if (!lu) {
                // cooperative Cholesky, in place in the lower triangle of LM; dinv[j] = 1 / L[j][j] in VC
                bool broke = false;
#pragma unroll
                for (int j = 0; j < NZ; ++j) {
                    Real d = LM(j * NZ + j);
#pragma unroll
                    for (int k = 0; k < j; ++k) d -= LM(j * NZ + k) * LM(j * NZ + k);
                    if (!(d > Real(0)) || !(d < Real(3e38))) broke = true;  // -> OPTI_KF_ST_NOT_PD, and the pivoted inverse below
                    const Real dinv = rsqrt(d);  // 1 / L[j][j]
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) {
                        const int r = j + 1 + q + 4 * rr;
                        if (r < NZ) {
                            Real v = LM(r * NZ + j);
#pragma unroll
                            for (int k = 0; k < j; ++k) v -= LM(r * NZ + k) * LM(j * NZ + k);
                            LM(r * NZ + j) = v * dinv;
                        }
                    }
                    if (q == 0) VC(j) = dinv;
                    quad_sync();
                }
                if (broke) {  // quad-uniform: every lane evaluates the same pivots
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) {
                        const int a = q + 4 * rr;
                        if (a < NZ) {
#pragma unroll
                            for (int b = 0; b < NZ; ++b)
                                if (b < a) LM(a * NZ + b) = LM(b * NZ + a);
                        }
                    }
                    quad_sync();
                    lu = true;
                    status |= OPTI_KF_ST_NOT_PD;
                }
            }
            if (!lu) {
                // NIS = |L^-1 y|^2 (replicated) and the three K rows of this lane: L u = P[i,sel]^T, then L^T k = u
                {
                    Real u[NZ];
#pragma unroll
                    for (int j = 0; j < NZ; ++j) {
                        Real v = y[j];
#pragma unroll
                        for (int k = 0; k < j; ++k) v -= LM(j * NZ + k) * u[k];
                        u[j] = v * VC(j);
                        nis += u[j] * u[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < NZ; ++j) {
                    const Real dj = VC(j);
                    Real v0 = Pr[0][sel(j)], v1 = Pr[1][sel(j)], v2 = Pr[2][sel(j)];
#pragma unroll
                    for (int k = 0; k < j; ++k) {
                        const Real l = LM(j * NZ + k);
                        v0 -= l * Kr[0][k]; v1 -= l * Kr[1][k]; v2 -= l * Kr[2][k];
                    }
                    Kr[0][j] = v0 * dj; Kr[1][j] = v1 * dj; Kr[2][j] = v2 * dj;
                }
#pragma unroll
                for (int j = NZ - 1; j >= 0; --j) {
                    const Real dj = VC(j);
                    Real v0 = Kr[0][j], v1 = Kr[1][j], v2 = Kr[2][j];
#pragma unroll
                    for (int k = j + 1; k < NZ; ++k) {
                        const Real l = LM(k * NZ + j);
                        v0 -= l * Kr[0][k]; v1 -= l * Kr[1][k]; v2 -= l * Kr[2][k];
                    }
                    Kr[0][j] = v0 * dj; Kr[1][j] = v1 * dj; Kr[2][j] = v2 * dj;
                }
            } else {
                if (asym) status |= OPTI_KF_ST_ASYMMETRIC;
                if (q == 0 && invert10_inplace(lm_)) status |= OPTI_KF_ST_NOT_PD | OPTI_KF_ST_SINGULAR;
                quad_sync();
#pragma unroll
                for (int a = 0; a < NZ; ++a) {
                    Real v = Real(0);
#pragma unroll
                    for (int b = 0; b < NZ; ++b) v += LM(a * NZ + b) * y[b];
                    nis += y[a] * v;
                }
#pragma unroll
                for (int j = 0; j < NZ; ++j) {
                    Real v0 = Real(0), v1 = Real(0), v2 = Real(0);
#pragma unroll
                    for (int k = 0; k < NZ; ++k) {
                        const Real si = LM(k * NZ + j);
                        v0 += Pr[0][sel(k)] * si; v1 += Pr[1][sel(k)] * si; v2 += Pr[2][sel(k)] * si;
                    }
                    Kr[0][j] = v0; Kr[1][j] = v1; Kr[2][j] = v2;
                }
            }
            // x <- x + K y: three entries per lane, then an all-gather over the quad
            {
                Real xo[3];
                own3(x, q, xo);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    Real v = Real(0);
#pragma unroll
                    for (int j = 0; j < NZ; ++j) v += Kr[a][j] * y[j];
                    xo[a] += v;
                }
#pragma unroll
                for (int src = 0; src < 4; ++src)
#pragma unroll
                    for (int a = 0; a < 3; ++a) x[3 * src + a] = quad_get(xo[a], lane, src);
            }
            // P <- P - K P[sel,:]  (old rows from the shared copy)
#pragma unroll
            for (int j = 0; j < NZ; ++j)
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    const Real ps = PF(sel(j) * NX + c);
                    Pr[0][c] -= Kr[0][j] * ps; Pr[1][c] -= Kr[1][j] * ps; Pr[2][c] -= Kr[2][j] * ps;
                }
            // np.trace of the 12x10 gain: K[j][j], j < 10 (select chain over compile-time indices, see own_diag)
            const Real g = q == 0 ? Kr[0][0] + Kr[1][1] + Kr[2][2]
                         : q == 1 ? Kr[0][3] + Kr[1][4] + Kr[2][5]
                         : q == 2 ? Kr[0][6] + Kr[1][7] + Kr[2][8] : Kr[0][9];
            ptrace = quad_sum(own_trace(Pr, q));
            kgain = quad_sum(g);
            ymax = fmax(ymax, nis);
            if (active && prm.K_final && t + 1 == prm.T) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int j = 0; j < NZ; ++j) prm.K_final[(long long)((r0 + a) * NZ + j) * N + i] = Kr[a][j];
            }
        }
        Real xq[3];
        own3(x, q, xq);
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (!isfinite(xq[a])) status |= OPTI_KF_ST_NONFINITE;

        if (active) {
            if (prm.x_steps) {
#pragma unroll
                for (int a = 0; a < 3; ++a) st_stream(prm.x_steps + (t * NX + r0 + a) * N + i, xq[a]);
            }
            if (q == 0) {
                if (prm.p_trace_steps) st_stream(prm.p_trace_steps + t * N + i, ptrace);
                if (prm.k_gain_steps) st_stream(prm.k_gain_steps + t * N + i, kgain);
                if (prm.nis_steps) st_stream(prm.nis_steps + t * N + i, nis);
            }
            if (prm.P_ckpt && prm.ckpt_every > 0 && (t + 1) % prm.ckpt_every == 0) {
                Real *dst = prm.P_ckpt + ((t + 1) / prm.ckpt_every - 1) * (long long)(NX * NX) * N + i;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < NX; ++c) dst[(long long)((r0 + a) * NX + c) * N] = Pr[a][c];
            }
        }
        if (prm.summary) {
            acc_nis += (double)nis;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (prm.truth) {
                    const double e = (double)xq[a] - (double)ld_stream(prm.truth + (t * NX + r0 + a) * S + s);
                    acc_truth[a] += e * e;
                }
                if (prm.nominal) {
                    const double e = (double)xq[a] - (double)ld_stream(prm.nominal + (t * NX + r0 + a) * S + s);
                    acc_nom[a] += e * e;
                }
            }
        }
    }

    status = quad_or(status);
    if (!active) return;
    if (prm.K_final && !(prm.phases & OPTI_KF_PHASE_UPDATE)) {  // no update ran: the gain is defined as zero
        for (int e = 0; e < 3 * NZ; ++e) prm.K_final[(long long)(r0 * NZ + e) * N + i] = Real(0);
    }
    Real xf[3];
    own3(x, q, xf);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int row = r0 + a;
        if (prm.x_final) prm.x_final[row * N + i] = xf[a];
        if (prm.P_final) {
#pragma unroll
            for (int c = 0; c < NX; ++c) prm.P_final[(long long)(row * NX + c) * N + i] = Pr[a][c];
        }
        if (prm.summary) {
            const double invT = prm.T > 0 ? 1.0 / (double)prm.T : 0.0;
            st_summary(prm, row, i, xf[a]);
            st_summary(prm, 12 + row, i, own_diag(Pr, q, a));
            st_summary(prm, 24 + row, i, (Real)sqrt(acc_truth[a] * invT));
            st_summary(prm, 36 + row, i, (Real)sqrt(acc_nom[a] * invT));
        }
    }
    if (q == 0) {
        if (prm.summary) {
            const double invT = prm.T > 0 ? 1.0 / (double)prm.T : 0.0;
            st_summary(prm, 48, i, (Real)(acc_nis * invT));
            st_summary(prm, 49, i, ptrace);
            st_summary(prm, 50, i, kgain);
            st_summary(prm, 51, i, (Real)sqrt((double)ymax));
        }
        if (prm.status) prm.status[i] = status;
    }
#undef PF
#undef LM
#undef VC
}

}  // namespace okf
