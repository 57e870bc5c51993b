// Throughput kernel, streamed variant: same arithmetic as kf_seq.cuh (one thread per trajectory, packed symmetric
// P and x in registers, sequential scalar updates) with the per-step inputs delivered by the TMA engine.
//
//   * A block owns 128 consecutive trajectories that read 128 consecutive base streams, so every input channel
//     of one step is one contiguous 128-element run of the [T][C][S] layout.  Lanes of warp 0 issue one
//     cp.async.bulk (global -> shared, completion on an mbarrier with complete_tx) per channel: no per-thread
//     address arithmetic, no LDG in the recursion, no registers tied up by loads in flight.
//   * Two stages: the copy for step t+2 is issued as soon as every thread has finished reading step t, i.e. a
//     whole filter step (~3 us) ahead of its use, which hides HBM/L2 latency behind the FMA work even at
//     8 warps per SM (the FP64 state needs ~250 registers per thread).
//   * Inputs are consumed straight from shared memory (conflict-free: consecutive threads, consecutive words);
//     z_j is read when measurement j is folded in, so nothing but x and P stays live across the step.
//   * Channels per step: p[12] f[12] z[10] (+ truth[12], nominal[12] label streams when the summary wants them);
//     z comes from the measurement pre-pass (optistate_kf_measure) because it is state-independent and, with
//     shared base streams, identical for every Monte-Carlo member of a stream.
#pragma once

#include "kf_seq.cuh"

namespace okf {

constexpr int TMA_THREADS = 128;
constexpr int TMA_STAGES = 2;
constexpr int TMA_CH_BASE = 34;  // p[12] f[12] z[10]

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename Real>
struct TmaSmem {
    // byte offsets inside dynamic shared memory, all 128-byte aligned
    static __host__ __device__ constexpr size_t stage_bytes(int ch) { return (size_t)ch * TMA_THREADS * sizeof(Real); }
    static __host__ __device__ constexpr size_t off_stage() { return 128; }  // mbarriers live in the first 128 bytes
    static __host__ __device__ constexpr size_t off_noise(int ch) { return off_stage() + TMA_STAGES * stage_bytes(ch); }
    static __host__ __device__ constexpr size_t off_acc(int ch) { return off_noise(ch) + (size_t)SEQ_NOISE_ROWS * TMA_THREADS * sizeof(Real); }
    static __host__ __device__ constexpr size_t total(int ch, bool acc_in_smem) {
        return off_acc(ch) + (acc_in_smem ? (size_t)25 * TMA_THREADS * sizeof(double) : 0);
    }
};

// warp 0: one bulk copy per channel of step t into stage buffer `dst`
template <typename Real>
__device__ __forceinline__ void issue_step_loads(const Params<Real> &prm, long long t, long long s0, int n_ch, Real *dst,
                                                 uint64_t *bar, int lane) {
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(n_ch * TMA_THREADS * sizeof(Real)));
    __syncwarp();
    const Real *lab0 = prm.truth ? prm.truth : prm.nominal;
    const Real *lab1 = prm.nominal;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int ch = lane + 32 * k;
        if (ch < n_ch) {
            const Real *arr;
            int C, c;
            if (ch < 12) { arr = prm.p; C = 12; c = ch; }
            else if (ch < 24) { arr = prm.f; C = 12; c = ch - 12; }
            else if (ch < 34) { arr = prm.z_in; C = 10; c = ch - 24; }
            else if (ch < 46) { arr = lab0; C = 12; c = ch - 34; }
            else { arr = lab1; C = 12; c = ch - 46; }
            bulk_g2s(dst + ch * TMA_THREADS, arr + (t * C + c) * prm.S + s0, (uint32_t)(TMA_THREADS * sizeof(Real)), bar);
        }
    }
}

template <typename Real, bool kSummary>
__global__ void __launch_bounds__(TMA_THREADS) kf_seq_tma_kernel(const __grid_constant__ Params<Real> prm) {
    constexpr bool kAccSmem = kSummary && sizeof(Real) == 8;  // FP64: the 25 running sums do not fit next to P in registers
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int nt = TMA_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long N = prm.N, S = prm.S;
    const long long i0 = (long long)blockIdx.x * nt;
    const long long i = i0 + tid;
    const bool active = i < N;
    const long long ic = active ? i : N - 1;  // clamped index for per-trajectory parameter loads
    const long long s0 = (i0 + prm.stream_offset) % S;
    const int n_lab = (prm.truth ? 1 : 0) + (prm.nominal ? 1 : 0);
    const int n_ch = TMA_CH_BASE + (kSummary ? 12 * n_lab : 0);

    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    Real *stage = reinterpret_cast<Real *>(smem_raw + TmaSmem<Real>::off_stage());
    Real *noise = reinterpret_cast<Real *>(smem_raw + TmaSmem<Real>::off_noise(n_ch));
    double *acc_s = reinterpret_cast<double *>(smem_raw + TmaSmem<Real>::off_acc(n_ch)) + tid;
    const int stage_elems = n_ch * nt;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (prm.T > 0) issue_step_loads(prm, 0, s0, n_ch, stage, &bars[0], lane);
        if (prm.T > 1) issue_step_loads(prm, 1, s0, n_ch, stage + stage_elems, &bars[1], lane);
    }

    Real *q = noise + tid, *r = noise + 12 * nt + tid, *rinv = noise + 22 * nt + tid;
#pragma unroll
    for (int c = 0; c < NX; ++c) q[c * nt] = prm.q_kind == OPTI_KF_MAT_DIAG ? prm.Q[c] : prm.Q[c * N + ic];
#pragma unroll
    for (int c = 0; c < NZ; ++c) {
        const Real rv = prm.r_kind == OPTI_KF_MAT_DIAG ? prm.R[c] : prm.R[c * N + ic];
        r[c * nt] = rv;
        rinv[c * nt] = Real(1) / rv;
    }
    Real x[NX], P[NP];
#pragma unroll
    for (int c = 0; c < NX; ++c) x[c] = prm.x0[c * prm.x0_ld + ic * prm.x0_inc];
#pragma unroll
    for (int a = 0; a < NX; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            Real v;
            switch (prm.p0_kind) {
                case OPTI_KF_MAT_NONE: v = (a == b) ? q[a * nt] : Real(0); break;
                case OPTI_KF_MAT_DIAG: v = (a == b) ? prm.P0[a] : Real(0); break;
                case OPTI_KF_MAT_DIAG_PER: v = (a == b) ? prm.P0[a * N + ic] : Real(0); break;
                case OPTI_KF_MAT_DENSE: v = prm.P0[a * NX + b]; break;
                default: v = prm.P0[(long long)(a * NX + b) * N + ic]; break;
            }
            P[tri(a, b)] = v;
        }

    uint32_t status = 0;
    Real ptrace = Real(0), kgain = Real(0), ymax = Real(0);
    constexpr int kAccRegs = (kSummary && !kAccSmem) ? 25 : 1;
    double acc_r[kAccRegs];
    auto acc_add = [&](int idx, double v) {  // running sums: 0-11 truth, 12-23 nominal, 24 NIS
        if constexpr (kAccSmem) acc_s[idx * nt] += v;
        else if constexpr (kSummary) acc_r[idx] += v;
    };
    auto acc_get = [&](int idx) -> double {
        if constexpr (kAccSmem) return acc_s[idx * nt];
        else if constexpr (kSummary) return acc_r[idx];
        else return 0.0;
    };
    if constexpr (kSummary) {
#pragma unroll
        for (int c = 0; c < 25; ++c) {
            if constexpr (kAccSmem) acc_s[c * nt] = 0.0; else acc_r[c] = 0.0;
        }
    }
    const bool want_gain = prm.k_gain_steps != nullptr || prm.summary != nullptr;

    for (long long t = 0; t < prm.T; ++t) {
        const int st = (int)(t & 1);
        mbar_wait(&bars[st], (uint32_t)((t >> 1) & 1));
        const Real *in = stage + st * stage_elems + tid;

        Real pf[12], ff[12], Rm[9];
#pragma unroll
        for (int c = 0; c < 12; ++c) { pf[c] = in[c * nt]; ff[c] = in[(12 + c) * nt]; }
        propagate_mean(prm, x, pf, ff, Rm);
        if (active) {
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
                for (int c = 0; c < NZ; ++c) st_stream(prm.z_steps + (t * NZ + c) * N + i, in[(24 + c) * nt]);
            }
        }

        cov_predict_sym(P, Rm, prm.dt, q, nt);

        Real nis = Real(0);
        fold_measurement<0>(P, x, in[24 * nt], r[0 * nt], nis, status);
        fold_measurement<1>(P, x, in[25 * nt], r[1 * nt], nis, status);
        fold_measurement<2>(P, x, in[26 * nt], r[2 * nt], nis, status);
        fold_measurement<3>(P, x, in[27 * nt], r[3 * nt], nis, status);
        fold_measurement<4>(P, x, in[28 * nt], r[4 * nt], nis, status);
        fold_measurement<5>(P, x, in[29 * nt], r[5 * nt], nis, status);
        fold_measurement<6>(P, x, in[30 * nt], r[6 * nt], nis, status);
        fold_measurement<7>(P, x, in[31 * nt], r[7 * nt], nis, status);
        fold_measurement<8>(P, x, in[32 * nt], r[8 * nt], nis, status);
        fold_measurement<9>(P, x, in[33 * nt], r[9 * nt], nis, status);

        ymax = fmax(ymax, nis);
        ptrace = Real(0);
#pragma unroll
        for (int c = 0; c < NX; ++c) ptrace += P[tri(c, c)];
        if (want_gain) {
            kgain = Real(0);
#pragma unroll
            for (int j = 0; j < NZ; ++j) kgain += P[tri(j, sel(j))] * rinv[j * nt];
        }
        bool fin = true;
#pragma unroll
        for (int c = 0; c < NX; ++c) fin &= isfinite(x[c]);
        if (!fin) status |= OPTI_KF_ST_NONFINITE;

        if (active) {
            if (prm.x_steps) {
#pragma unroll
                for (int c = 0; c < NX; ++c) st_stream(prm.x_steps + (t * NX + c) * N + i, x[c]);
            }
            if (prm.p_trace_steps) st_stream(prm.p_trace_steps + t * N + i, ptrace);
            if (prm.k_gain_steps) st_stream(prm.k_gain_steps + t * N + i, kgain);
            if (prm.nis_steps) st_stream(prm.nis_steps + t * N + i, nis);
            if (prm.P_ckpt && prm.ckpt_every > 0 && (t + 1) % prm.ckpt_every == 0) {
                Real *dst = prm.P_ckpt + ((t + 1) / prm.ckpt_every - 1) * (long long)(NX * NX) * N + i;
#pragma unroll
                for (int a = 0; a < NX; ++a)
#pragma unroll
                    for (int b = 0; b < NX; ++b) dst[(long long)(a * NX + b) * N] = P[tri(a, b)];
            }
        }
        if constexpr (kSummary) {
            acc_add(24, (double)nis);
            if (prm.truth) {
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    const double e = (double)x[c] - (double)in[(34 + c) * nt];
                    acc_add(c, e * e);
                }
            }
            if (prm.nominal) {
                const int base = prm.truth ? 46 : 34;
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    const double e = (double)x[c] - (double)in[(base + c) * nt];
                    acc_add(12 + c, e * e);
                }
            }
        }

        __syncthreads();  // every thread is done with stage `st`: refill it with step t + 2
        if (warp == 0 && t + TMA_STAGES < prm.T) issue_step_loads(prm, t + TMA_STAGES, s0, n_ch, stage + st * stage_elems, &bars[st], lane);
    }

    if (!active) return;
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
            sm[(long long)(24 + c) * N] = (Real)sqrt(acc_get(c) * invT);
            sm[(long long)(36 + c) * N] = (Real)sqrt(acc_get(12 + c) * invT);
        }
        sm[48LL * N] = (Real)(acc_get(24) * invT);
        sm[49LL * N] = ptrace;
        sm[50LL * N] = kgain;
        sm[51LL * N] = (Real)sqrt((double)ymax);
    }
    if (prm.status) prm.status[i] = status | (prm.stream_status ? prm.stream_status[(s0 + tid) % S] : 0u);
}

}  // namespace okf
