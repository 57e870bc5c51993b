// Batched force MPC, dual active-set solver (Goldfarb & Idnani 1983) for problems with at most two legs out of swing - the
// primary path of optistate_kf_mpc_forces for a trot.  Same QP and same minimiser as kf_mpc.cuh / kf_mpc_rows.cuh (the QP is
// strictly convex: the minimiser is unique and the method terminates at it); what changes is the work per problem:
//
//   * the interior point refactorises the n x n normal equations every iteration (~13 factorisations + ~30 triangular solves
//     per QP, each a chain of n dependent columns).  But H is FIXED per problem and every constraint row touches one leg of one
//     stage (three unknowns), so here H is inverted ONCE - symmetric sweeps, lane i owns row i in registers, one broadcast
//     through a shared-memory strip per column - and everything after that is small:
//   * the method starts at the unconstrained minimiser -H^-1 g and adds the most violated constraint p at a time.  With N the
//     normals of the working set: y = H^-1 n_p is THREE columns of H^-1 combined; d = N^T y three products per working
//     constraint; r = S^-1 d a q x q product with S^-1 = (N^T H^-1 N)^-1 KEPT EXPLICITLY and updated by the bordering formula
//     when a constraint enters and by a rank-one downdate when one leaves (no factorisation, no dependent chain); the primal
//     direction z = y - H^-1 N r three columns per working constraint.  One iteration is a few hundred instructions on
//     shared memory instead of a factorisation, and the ~14 iterations of a saturated trot problem cost about what ONE
//     interior-point iteration costs.
//   * a partial step (a multiplier reaches zero first) drops that constraint and repeats with the same p; a normal that is
//     linearly dependent on the working set (the apex of the friction pyramid: four faces, three unknowns) takes the dual
//     step alone, exactly as in the paper.
//   * the loop is rolled and compact (it fits the 32 KB instruction cache); only the inversion is straight-line code, run once.
//   * a problem the method gives up on (iteration cap, working set full, a breakdown of S^-1) is flagged and solved by the
//     interior-point kernel (kf_mpc_rows.cuh) in a second launch that only touches flagged problems.
// Measured against the independent active-set solve of the test infrastructure: see tests/test_mpc_gpu.py.
#pragma once

#include "kf_mpc_rows.cuh"

namespace okf {

constexpr int MPCG_WARPS = 5;       // problems per block: 3 blocks = 15 warps per SM fit the shared memory (14.7 KB per problem)
constexpr int MPCG_SLOTS = 24;      // capacity of the working set (linear independence bounds it by n = 30; the tests peak at 21)
constexpr int MPCG_LDP = MPCG_SLOTS + 1;
constexpr int MPCG_MAX_IT = 200;    // constraints added + dropped

// per-warp shared memory (doubles): H^-1 [n][n|1] (Su [12][n] while H is built) | S^-1 [SLOTS][SLOTS+1] | x g | strip 2 x 64 (after the
// inversion: y d r) | normals [3][32] | block of a slot (32 ints)
__host__ __device__ constexpr int mpcg_warp_doubles() {
    constexpr int n = 15 * MPCR_MAX_LEGS;
    return mpcr_even(n * (n | 1)) + MPCG_SLOTS * MPCG_LDP + 2 * 32 + 128 + 3 * 32 + 16;
}
__host__ __device__ constexpr size_t mpcg_smem_bytes() { return (size_t)MPCG_WARPS * mpcg_warp_doubles() * sizeof(double); }

// coefficient c of constraint row r of a stance block:  r0: fz <= fz_max;  r1: fx - mu fz <= 0;  r2: -fx - mu fz <= 0;
// r3: fy - mu fz <= 0;  r4: -fy - mu fz <= 0
__device__ __forceinline__ double row_coef(int r, int c, double mu) {
    if (c == 0) return r == 1 ? 1.0 : (r == 2 ? -1.0 : 0.0);
    if (c == 1) return r == 3 ? 1.0 : (r == 4 ? -1.0 : 0.0);
    return r == 0 ? 1.0 : -mu;
}

// smallest (value, index) of the warp; ties go to the smaller index; index < 0 = no candidate
__device__ __forceinline__ void warp_argmin(double &v, int &idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (oi >= 0 && (idx < 0 || ov < v || (ov == v && oi < idx))) { v = ov; idx = oi; }
    }
}

// In place: m = row `lane` of the symmetric positive definite H  ->  row `lane` of -H^-1 (symmetric sweeps on every pivot).
// strip: 2 x 64 doubles.  Returns false (warp-uniform) on a non-positive pivot.
template <int N>
__device__ __forceinline__ bool sweep_invert(double (&m)[N], double *strip, int lane) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double *cb = strip + (k & 1) * 64;
        if (lane < N) cb[lane] = m[k];  // column k = row k (symmetry is kept by the sweeps)
        __syncwarp();
        const double d = cb[k];
        ok = ok && (d > 0.0);
        const double inv = rcp2_(d);
        const double l = lane == k ? 1.0 - inv : m[k] * inv;  // the pivot row is scaled by 1 / d, the others eliminated
        const double2 *cp = reinterpret_cast<const double2 *>(cb);
#pragma unroll
        for (int q = 0; q < N / 2; ++q) {
            const double2 v = cp[q];
            m[2 * q] = fma(-l, v.x, m[2 * q]);
            m[2 * q + 1] = fma(-l, v.y, m[2 * q + 1]);
        }
        m[k] = lane == k ? -inv : l;
    }
    return ok;
}

template <int NFL>
__device__ __forceinline__ void mpc_solve_gi(const MpcParams &prm, long long prob, int lane, double *base, const int (&kind_leg)[4],
                                             const int (&free_leg)[4]) {
    constexpr int n = 15 * NFL, nb = 5 * NFL, ldg = n | 1;
    static_assert(n % 2 == 0 && 12 * mpcr_even(n) <= n * ldg, "layout");
    const long long N = prm.N;
    double *Ginv = base, *Su = base, *P = Ginv + mpcr_even(n * ldg), *xs = P + MPCG_SLOTS * MPCG_LDP, *g = xs + 32, *strip = g + 32;
    double *y = strip, *dslot = y + 32, *rslot = dslot + 32, *ncoef = strip + 128;  // y, d, r reuse the strip once H is inverted
    int *sblk = reinterpret_cast<int *>(ncoef + 96);

    int my_leg = 0;  // leg of the block this lane owns (blocks: stage-major, NFL legs per stage)
#pragma unroll
    for (int r = 0; r < NFL; ++r)
        if (r == lane % NFL) my_leg = free_leg[r];
    int my_kind = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l)
        if (l == my_leg) my_kind = kind_leg[l];
    const bool my_act = lane < nb && my_kind == 1;
    const double mu_f = prm.mu;
    const int row = lane < n ? lane : n - 1;
    const double bvec[5] = {prm.fz_max, 0.0, 0.0, 0.0, 0.0};

    // ---- H, g; H^-1; the unconstrained minimiser ----------------------------------------------------------------------------
    for (int e = lane; e < 128; e += 32) strip[e] = 0.0;
    double m[n], grow;
    mpc_condense<NFL>(prm, prob, lane, kind_leg, free_leg, Su, m, grow);
#pragma unroll
    for (int j = 0; j < n; ++j)
        if (j == row) m[j] += 2.0 * prm.w_force;
    if (lane < n) g[lane] = grow;
    __syncwarp();  // Su is dead from here on: its storage becomes H^-1
    bool ok = sweep_invert<n>(m, strip, lane);
    double xi = 0.0;  // entry `lane` of the iterate
    {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < n; j += 2) {
            a0 = fma(m[j], g[j], a0);  // m = -H^-1
            a1 = fma(m[j + 1], g[j + 1], a1);
        }
        xi = a0 + a1;
    }
    __syncwarp();  // every lane has read g: from here on g[] holds the unconstrained minimiser x0 (the warm start needs it)
    if (lane < n) {
#pragma unroll
        for (int j = 0; j < n; ++j) Ginv[lane * ldg + j] = -m[j];
        xs[lane] = xi;
        g[lane] = xi;
    }

    // ---- dual active set -------------------------------------------------------------------------------------------------------
    // slot role (lane j < MPCG_SLOTS): multiplier uj, normal (nj0, nj1, nj2) and block bj of working constraint j, its id (5 block + row)
    // block role (lane b < nb): inA = rows of block b that are in the working set
    uint32_t valid = 0u, inA = 0u;
    double uj = 0.0, nj0 = 0.0, nj1 = 0.0, nj2 = 0.0;
    int bj = 0, myid = -1;
    int it = 0;
    uint32_t status = 0u;
    const double feas_tol = 1e-9 * prm.fz_max;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double *gr = Ginv + row * ldg;  // this lane's row of H^-1
    const auto is_slot = [&]() { return ((valid >> lane) & 1u) != 0u; };
    // r_j = sum_k S^-1[j][k] v[k] over the working slots (v in shared memory)
    const auto times_sinv = [&](const double *v) {
        double acc = 0.0;
        if (is_slot()) {
            const double *pr = P + lane * MPCG_LDP;
            for (uint32_t mk = valid; mk; mk &= mk - 1u) {
                const int k = __ffs(mk) - 1;
                acc = fma(pr[k], v[k], acc);
            }
        }
        return acc;
    };
    // entry `lane` of sum_k c[k] H^-1 n_k over the working slots (three columns of H^-1 per constraint)
    const auto times_hinv_n = [&](const double *c) {
        double acc = 0.0;
        for (uint32_t mk = valid; mk; mk &= mk - 1u) {
            const int k = __ffs(mk) - 1;
            const int bk = sblk[k];
            const double hk = fma(gr[3 * bk + 2], ncoef[64 + k], fma(gr[3 * bk + 1], ncoef[32 + k], gr[3 * bk] * ncoef[k]));
            acc = fma(c[k], hk, acc);
        }
        return acc;
    };
    // constraint p = 5 bp + rp with normal np enters slot s:  S^-1 <- [[S^-1 + r r^T / delta, -r / delta], [-r^T / delta, 1 / delta]]
    // (r in rslot and, per lane, rj;  delta = n^T H^-1 n - d^T r)
    const auto enter = [&](int s, double rj, double delta, int p, int bp, int rp, double np0, double np1, double np2, double u_new) {
        const double idel = rcp2_(delta);
        if (is_slot()) {
            double *pr = P + lane * MPCG_LDP;
            const double f = rj * idel;
            for (uint32_t mk = valid; mk; mk &= mk - 1u) {
                const int k = __ffs(mk) - 1;
                pr[k] = fma(f, rslot[k], pr[k]);
            }
            pr[s] = -f;
        }
        if (lane == s) {
            double *pr = P + s * MPCG_LDP;
            for (uint32_t mk = valid; mk; mk &= mk - 1u) {
                const int k = __ffs(mk) - 1;
                pr[k] = -rslot[k] * idel;
            }
            pr[s] = idel;
            uj = u_new; nj0 = np0; nj1 = np1; nj2 = np2; bj = bp; myid = p;
            sblk[s] = bp;
            ncoef[s] = np0; ncoef[32 + s] = np1; ncoef[64 + s] = np2;
        }
        valid |= 1u << s;
        if (lane == bp) inA |= 1u << rp;
    };
    // working constraint k leaves:  S^-1 <- S^-1 - S^-1[:, k] S^-1[k, :] / S^-1[k][k] on the remaining slots
    const auto leave = [&](int k) {
        const double ipkk = rcp2_(P[k * MPCG_LDP + k]);
        const uint32_t rest = valid & ~(1u << k);
        if ((rest >> lane) & 1u) {
            double *pr = P + lane * MPCG_LDP;
            const double f = pr[k] * ipkk;
            const double *pk = P + k * MPCG_LDP;
            for (uint32_t mk = rest; mk; mk &= mk - 1u) {
                const int kk = __ffs(mk) - 1;
                pr[kk] = fma(-f, pk[kk], pr[kk]);
            }
        }
        const int idk = __shfl_sync(0xffffffffu, myid, k);
        if (lane == idk / 5) inA &= ~(1u << (idk % 5));
        if (lane == k) { uj = 0.0; myid = -1; }
        valid = rest;
    };

    uint32_t pattern = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) pattern |= kind_leg[l] != 0 ? (1u << (24 + l)) : 0u;
    // ---- warm start: the working set of the previous solve of this problem (closed loops) -----------------------------------
    // Its constraints enter by bordering alone (no step, no ratio test), the minimiser on that set and its multipliers follow from
    // one product with S^-1, and constraints whose multiplier comes out negative leave one at a time: what remains is a point the
    // dual method may start from (the minimiser on its working set, all multipliers >= 0), usually the answer or one step from it.
    if (ok && prm.warm_set) {
        const uint32_t word = lane < nb ? prm.warm_set[(long long)(lane / NFL) * N + prob] : 0u;
        const uint32_t word0 = __shfl_sync(0xffffffffu, word, 0);
        if ((word0 >> 31) && (word0 & 0x0f000000u) == pattern) {
            uint32_t cand = my_act ? ((word >> (5 * my_leg)) & 31u) : 0u;
            const double gs = fmax(warp_max(lane < n ? fabs(grow) : 0.0), 1e-300);
#pragma unroll 1
            while (true) {
                const unsigned who = __ballot_sync(0xffffffffu, cand != 0u);
                const uint32_t freeslots = ~valid & ((1u << MPCG_SLOTS) - 1u);
                if (!who || !freeslots) break;
                const int bp = __ffs(who) - 1;
                const int rp = __ffs(__shfl_sync(0xffffffffu, cand, bp)) - 1;
                if (lane == bp) cand &= cand - 1u;
                const double np0 = -row_coef(rp, 0, mu_f), np1 = -row_coef(rp, 1, mu_f), np2 = -row_coef(rp, 2, mu_f);
                const double yi = fma(gr[3 * bp + 2], np2, fma(gr[3 * bp + 1], np1, gr[3 * bp] * np0));
                __syncwarp();
                if (lane < n) y[lane] = yi;
                __syncwarp();
                const double nGn = fma(np2, y[3 * bp + 2], fma(np1, y[3 * bp + 1], np0 * y[3 * bp]));
                const double dj = is_slot() ? fma(nj2, y[3 * bj + 2], fma(nj1, y[3 * bj + 1], nj0 * y[3 * bj])) : 0.0;
                dslot[lane] = dj;
                __syncwarp();
                const double rj = times_sinv(dslot);
                rslot[lane] = rj;
                const double delta = nGn - warp_sum(dj * rj);
                __syncwarp();
                if (delta > 1e-10 * nGn) enter(__ffs(freeslots) - 1, rj, delta, 5 * bp + rp, bp, rp, np0, np1, np2, 0.0);  // else: dependent on the set so far
            }
            if (valid) {
                status |= 8u;
#pragma unroll 1
                while (valid) {
                    __syncwarp();
                    // u = S^-1 (b' - N^T x0), x = x0 + H^-1 N u   (x0 = the unconstrained minimiser, kept in g[])
                    const int rj_row = myid - 5 * bj;
                    double rhs = 0.0;
#pragma unroll
                    for (int r = 0; r < 5; ++r)
                        if (r == rj_row) rhs = -bvec[r];
                    dslot[lane] = is_slot() ? rhs - fma(nj2, g[3 * bj + 2], fma(nj1, g[3 * bj + 1], nj0 * g[3 * bj])) : 0.0;
                    __syncwarp();
                    uj = times_sinv(dslot);
                    rslot[lane] = uj;
                    __syncwarp();
                    xi = (lane < n ? g[lane] : 0.0) + times_hinv_n(rslot);
                    double worst = is_slot() ? uj : inf;
                    int kw = is_slot() ? lane : -1;
                    warp_argmin(worst, kw);
                    if (!(worst < -1e-12 * gs)) break;
                    leave(kw);
                    ++it;
                }
                __syncwarp();
                if (lane < n) xs[lane] = valid ? xi : g[lane];
                if (!valid) xi = lane < n ? g[lane] : 0.0;
            }
        }
    }

    bool done = false;
#pragma unroll 1
    while (ok && !done) {
        __syncwarp();  // xs is current
        // most violated constraint outside the working set
        double best = inf;
        int p = -1;
        if (my_act) {
            const double x0 = xs[3 * lane], x1 = xs[3 * lane + 1], x2 = xs[3 * lane + 2];
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const double sl = bvec[r] - (row_coef(r, 0, mu_f) * x0 + row_coef(r, 1, mu_f) * x1 + row_coef(r, 2, mu_f) * x2);
                if (!((inA >> r) & 1u) && sl < best) { best = sl; p = 5 * lane + r; }
            }
        }
        warp_argmin(best, p);
        if (p < 0 || best >= -feas_tol) { done = true; break; }
        const int bp = p / 5, rp = p - 5 * bp;
        const double np0 = -row_coef(rp, 0, mu_f), np1 = -row_coef(rp, 1, mu_f), np2 = -row_coef(rp, 2, mu_f);  // normal of n^T x >= b'
        double sp = best, up = 0.0;
        // y = H^-1 n_p, d = N^T y: they do not change while constraints are dropped for this p
        const double yi = fma(gr[3 * bp + 2], np2, fma(gr[3 * bp + 1], np1, gr[3 * bp] * np0));
        if (lane < n) y[lane] = yi;
        __syncwarp();
        const double nGn = fma(np2, y[3 * bp + 2], fma(np1, y[3 * bp + 1], np0 * y[3 * bp]));
        const double dj = is_slot() ? fma(nj2, y[3 * bj + 2], fma(nj1, y[3 * bj + 1], nj0 * y[3 * bj])) : 0.0;
        dslot[lane] = dj;
        bool added = false;
#pragma unroll 1
        while (!added) {
            __syncwarp();  // dslot (first pass) / the downdated S^-1 (later passes) is visible
            // r = S^-1 d, z = y - H^-1 N r
            const double rj = is_slot() ? times_sinv(dslot) : 0.0;
            rslot[lane] = rj;
            const double zn = nGn - warp_sum(dj * rj);
            __syncwarp();
            const double zi = yi - times_hinv_n(rslot);
            const bool dep = !(zn > 1e-12 * nGn);  // n_p is (numerically) a combination of the working normals: dual step only
            // step lengths: t1 keeps the multipliers non-negative, t2 makes constraint p hold
            double t1 = (is_slot() && rj > 1e-300) ? uj * rcp2_(rj) : inf;
            int k1 = (is_slot() && rj > 1e-300) ? lane : -1;
            warp_argmin(t1, k1);
            if (k1 < 0) t1 = inf;
            const double t2 = dep ? inf : -sp * rcp2_(zn);
            const double t = fmin(t1, t2);
            if (!(t < inf)) { ok = false; break; }  // neither step exists: cannot happen for a feasible problem in exact arithmetic
            if (!dep) {
                xi = fma(t, zi, xi);
                sp = fma(t, zn, sp);
            }
            uj = is_slot() ? fma(-t, rj, uj) : 0.0;
            up += t;
            ++it;
            if (it > MPCG_MAX_IT) { ok = false; break; }
            if (!dep && t2 <= t1) {  // full step: p enters
                const uint32_t freeslots = ~valid & ((1u << MPCG_SLOTS) - 1u);
                if (!freeslots) { ok = false; break; }
                enter(__ffs(freeslots) - 1, rj, zn, p, bp, rp, np0, np1, np2, up);
                if (lane < n) xs[lane] = xi;
                added = true;
            } else {
                leave(k1);  // partial (or purely dual) step: the multiplier of working constraint k1 has reached zero
            }
        }
    }

    // ---- results -----------------------------------------------------------------------------------------------------------------
    for (int e = lane; e < MPC_N; e += 32) prm.forces[(long long)e * N + prob] = ok ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
    __syncwarp();
    if (ok && lane < n) {  // compact -> (stage, leg, component)
        const int stage = lane / (3 * NFL), within = lane % (3 * NFL);
        int l = 0;
#pragma unroll
        for (int r = 0; r < NFL; ++r)
            if (r == within / 3) l = free_leg[r];
        prm.forces[(long long)(12 * stage + 3 * l + within % 3) * N + prob] = xi;
    }
    if (prm.warm_set) {
        // the working set of this solve, by absolute (stage, leg, row), and its multipliers
        uint32_t bits = ok ? (inA << (5 * my_leg)) : 0u;
#pragma unroll
        for (int o = 1; o < NFL; o <<= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);  // the legs of one stage sit in adjacent lanes
        if (lane < nb && lane % NFL == 0) prm.warm_set[(long long)(lane / NFL) * N + prob] = ok ? (bits | pattern | 0x80000000u) : 0u;
        if (ok && is_slot()) {
            int l = 0;
#pragma unroll
            for (int r = 0; r < NFL; ++r)
                if (r == bj % NFL) l = free_leg[r];
            prm.warm_mult[(long long)(((bj / NFL) * 4 + l) * 5 + (myid - 5 * bj)) * N + prob] = uj;
        }
    }
    if (lane == 0) prm.status[prob] = ok ? (status | ((uint32_t)it << 8)) : MPC_ST_GIVEN_UP;
}

// One warp per problem; prm.status is required (the second launch reads the flags).
__global__ void __launch_bounds__(32 * MPCG_WARPS, 3) kf_mpc_gi_kernel(const __grid_constant__ MpcParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long prob = (long long)blockIdx.x * MPCG_WARPS + warp;
    if (prob >= prm.N) return;  // whole warp
    double *base = reinterpret_cast<double *>(smem_raw) + (size_t)warp * mpcg_warp_doubles();
    const long long N = prm.N;
    int kind_leg[4];  // 0 pinned (swing), 1 pyramid (stance), 2 free
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const double c = prm.contact[l * N + prob];
        kind_leg[l] = c == 0.0 ? 0 : (c == 1.0 ? 1 : 2);
    }
    int free_leg[4] = {0, 0, 0, 0}, nfl = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l)
        if (kind_leg[l] != 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (r == nfl) free_leg[r] = l;
            ++nfl;
        }
    if (nfl > prm.max_legs || nfl > MPCR_MAX_LEGS) {  // the caller's bound on the legs out of swing is wrong for this problem: no answer
        for (int e = lane; e < MPC_N; e += 32) prm.forces[(long long)e * N + prob] = __longlong_as_double(0x7ff8000000000000LL);
        if (lane == 0) prm.status[prob] = 4u;
        return;
    }
    if (nfl == 0) {
        for (int e = lane; e < MPC_N; e += 32) prm.forces[(long long)e * N + prob] = 0.0;
        if (lane == 0) prm.status[prob] = 0u;
        return;
    }
    if (nfl == 1) {  // the lone leg is paired with a phantom that decouples exactly (mpc_condense)
#pragma unroll
        for (int l = 3; l >= 0; --l)
            if (kind_leg[l] == 0) free_leg[1] = l;
    }
    mpc_solve_gi<2>(prm, prob, lane, base, kind_leg, free_leg);
}

}  // namespace okf
