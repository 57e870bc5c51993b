// Batched force MPC, register-resident variant for problems with at most two legs out of swing (a trot: order n <= 30).
// Same QP, same interior-point + active-set-polish algorithm and the same answers as kf_mpc.cuh; what differs is where the
// linear algebra lives:
//
//   * ONE WARP per QP, LANE i OWNS ROW i of the normal-equations matrix M = H + A^T D A - all n entries of the row in
//     registers, symmetric storage.  The factorisation M = L D L^T is right-looking: per column the lanes publish their entry
//     of the column through a double-buffered 32-double strip of shared memory (one STS, one __syncwarp, broadcast LDS), take
//     the reciprocal of the pivot themselves and update their own row with n - k - 1 independent FMAs - no shared-memory
//     read-modify-write, no second barrier, the next column's entry is the first FMA so its round trip overlaps the rest.
//     After the last column the part of row i to the right of the diagonal is scaled into row i of L^T, so both triangular
//     solves run with the vector in registers too: per column one 64-bit shuffle + one FMA (no barrier, no shared memory).
//     kf_mpc.cuh's shared-memory forms (two barriers per column, two loads per FMA) remain for three and four stance legs.
//   * H (packed) stays in shared memory - it is re-read once per factorisation - and is accumulated lane-by-row in
//     registers: the stage loop of the condensed Hessian has 12 independent FMAs per entry instead of a serial sweep.
//   * WARM START (optional, for the closed loop estimate_state_mpc runs - consecutive QPs of a trajectory differ by one
//     filter step): the active set and multipliers of the previous solve are read back, the polish phase (method of
//     multipliers on that set + verification of primal and dual feasibility) is tried first, and the interior point only
//     runs when the verification fails within MPCR_WARM_ROUNDS rounds or the contact pattern has changed.
//   * The multiplier iteration stops as soon as the multipliers move by less than 1e-12 |g| (kf_mpc.cuh always runs 8).
#pragma once

#include "kf_mpc.cuh"

namespace okf {

constexpr int MPCR_WARPS = 4;         // problems per block (independent warps)
constexpr int MPCR_MAX_LEGS = 2;      // legs out of swing this kernel is instantiated for
constexpr int MPCR_WARM_ROUNDS = 3;   // polish rounds granted to a warm start before the interior point takes over

__host__ __device__ constexpr int mpcr_even(int v) { return v + (v & 1); }
// per-warp shared memory (doubles): H packed | factor -L [n][n|1] (Su [12][n] while H is built) | u g rv dv ukeep | column strip 2 x 64 |
// block terms [5 nfl][6]
__host__ __device__ constexpr int mpcr_warp_doubles(int max_legs) {
    return mpcr_even((15 * max_legs) * (15 * max_legs + 1) / 2) + mpcr_even((15 * max_legs) * ((15 * max_legs) | 1)) + 5 * mpcr_even(15 * max_legs) + 128 +
           6 * 5 * max_legs;
}
__host__ __device__ constexpr size_t mpcr_smem_bytes() { return (size_t)MPCR_WARPS * mpcr_warp_doubles(MPCR_MAX_LEGS) * sizeof(double); }

// Reciprocal of a positive normal number, straight-line (20-bit seed + two Newton steps: rounding-limited): the IEEE division is a
// call into a ~70-instruction slow-path routine, and the interior point takes ~45 of them per iteration.
__device__ __forceinline__ double rcp2_(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}

// Row-owned symmetric matrix of order N <= 32, N a multiple of MPCR_BC (see the header).  Surplus lanes (>= N) mirror the last
// row and never store.
//
// CODE SIZE is what bounds this kernel: a fully unrolled factorisation + solves + interior-point bookkeeping came to ~70 KB of
// SASS per iteration, twice the 32 KB L1.5 instruction cache - every pass streamed its code from L2 (ncu: SM i-cache hit rate
// 79 % = the sequential hits inside a line, GPC instruction-cache requests at 93 % of peak, stall_no_inst 15 %, and no gain from
// more resident warps).  So the loops that can be rolled are rolled:
//   * factor(): the lane's row lives in a register WINDOW whose position 0 is the first column of the current block of MPCR_BC
//     columns; the block is unrolled (static register indices), then the window is shifted by MPCR_BC and the block loop
//     repeats - one body of ~400 instructions instead of N columns x ~55.  The price: the tail of the window multiplies
//     zeros (+20 % FMAs) and 2 (N - MPCR_BC) register moves per block.
//   * the finished columns of -L go to shared memory ([N][ld], ld odd, explicit zeros on and above the diagonal), so both
//     triangular solves are rolled loops with a run-time column index: shuffle (broadcast of the entry that has just become
//     final) -> FMA with a coefficient that was loaded ahead of the chain; lanes the column does not touch see a zero.
constexpr int MPCR_BC = 6;

template <int N>
struct MpcRows {
    static_assert(N % MPCR_BC == 0 && MPCR_BC % 2 == 0 && N <= 32, "window blocking");
    static constexpr int ld = N | 1;  // odd row stride of the stored factor: conflict-free by rows and by columns
    double m[N];
    double dinv;

    // M <- H (packed lower triangle in shared memory) + 3x3 diagonal blocks (shared, six doubles per block: xx yx yy zx zy zz)
    __device__ __forceinline__ void form(const double *H, const double *blk, int lane) {
        const int i = lane < N ? lane : N - 1;
        const int bi = i / 3, ci = i - 3 * bi;
        const double *hb = H + i * (i + 1) / 2;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double v = j <= i ? hb[j] : H[j * (j + 1) / 2 + i];
            if (j / 3 == bi) {
                const int cj = j % 3;
                const int hi = ci > cj ? ci : cj, lo = ci > cj ? cj : ci;
                v += blk[6 * bi + hi * (hi + 1) / 2 + lo];
            }
            m[j] = v;
        }
    }

    // M = L D L^T: -L into Ls, 1 / d_i into dinv; m is consumed.  strip: 2 x 64 doubles whose entries >= N are zero.
    // Returns false (warp-uniform) on a non-positive pivot.
    __device__ __forceinline__ bool factor(double *strip, double *Ls, int lane) {
        bool ok = true;
#pragma unroll 1
        for (int K0 = 0; K0 < N; K0 += MPCR_BC) {
#pragma unroll
            for (int c = 0; c < MPCR_BC; ++c) {
                const int k = K0 + c;
                double *cb = strip + (c & 1) * 64;
                if (lane < N) cb[lane] = m[c];  // column k as it stands (rows <= k are not read)
                __syncwarp();
                const double d = cb[k];
                ok = ok && (d > 0.0);
                const double inv = rcp2_(d);
                if (lane == k) dinv = inv;
                const double nl = lane > k ? -(m[c] * inv) : 0.0;  // -L_ik; rows on and above the pivot take no part
                if (lane < N) Ls[k * ld + lane] = nl;
                const double2 *cp = reinterpret_cast<const double2 *>(cb + K0);  // K0 even: 16-byte aligned pairs
#pragma unroll
                for (int q = (c + 1) / 2; q < N / 2; ++q) {
                    const double2 v = cp[q];
                    if (2 * q > c) m[2 * q] = fma(nl, v.x, m[2 * q]);
                    m[2 * q + 1] = fma(nl, v.y, m[2 * q + 1]);
                }
            }
#pragma unroll
            for (int w = 0; w < N - MPCR_BC; ++w) m[w] = m[w + MPCR_BC];
#pragma unroll
            for (int w = N - MPCR_BC; w < N; ++w) m[w] = 0.0;
        }
        return ok;
    }

    // x = M^-1 b, lane i holds entry i
    __device__ __forceinline__ double solve(double b, const double *Ls, int lane) const {
        const int i = lane < N ? lane : N - 1;
        const double *col = Ls + i;       // forward: -L_ij = col[j ld], zero for i <= j
        const double *rowp = Ls + i * ld;  // backward: -L_ji = rowp[j], zero for j <= i
#pragma unroll 6
        for (int j = 0; j < N - 1; ++j) {
            const double yj = __shfl_sync(0xffffffffu, b, j);
            b = fma(col[j * ld], yj, b);
        }
        b *= dinv;
#pragma unroll 6
        for (int j = N - 1; j > 0; --j) {
            const double xj = __shfl_sync(0xffffffffu, b, j);
            b = fma(rowp[j], xj, b);
        }
        return b;
    }
};

// The threads that work on one problem: one warp, or - for the orders 45 and 60 of three and four legs out of swing in
// kf_mpc_gi.cuh - the two warps of a 64-thread block.  Thread t of the group owns row t / slot t / constraint block t.
// `scr`: eight doubles of shared memory per problem (only the two-warp form touches it).
__device__ __forceinline__ void warp_argmin(double &v, int &idx) {  // smallest (value, index); ties to the smaller index; index < 0 = none
    // Three warp-wide integer reductions (REDUX) on an order-preserving 64-bit key instead of five rounds of three shuffles and a
    // compare chain: the high word, the low word among the lanes that hold the smallest high word, then the index among those.
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);  // ascending as unsigned: negative values below positive ones
    if (idx < 0) u = ~0ull;
    const unsigned hi = (unsigned)(u >> 32), lo = (unsigned)u;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const bool mine = idx >= 0 && hi == mhi && lo == mlo;
    const unsigned midx = __reduce_min_sync(0xffffffffu, mine ? (unsigned)idx : 0xffffffffu);
    if (midx == 0xffffffffu) { idx = -1; return; }  // no candidate (v is left as it is)
    unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
    m = (m >> 63) ? (m & 0x7fffffffffffffffull) : ~m;
    v = __longlong_as_double((long long)m);
    idx = (int)midx;
}
template <int NW>
struct MpcGroup;
template <>
struct MpcGroup<1> {
    static constexpr int threads = 32;
    static __device__ __forceinline__ int tid() { return threadIdx.x & 31; }
    static __device__ __forceinline__ void sync() { __syncwarp(); }
    static __device__ __forceinline__ double sum(double v, double *) { return warp_sum(v); }
    static __device__ __forceinline__ double max(double v, double *) { return warp_max(v); }
    static __device__ __forceinline__ void argmin(double &v, int &idx, double *) { warp_argmin(v, idx); }
};
template <>
struct MpcGroup<2> {  // one problem per 64-thread block
    static constexpr int threads = 64;
    static __device__ __forceinline__ int tid() { return threadIdx.x; }
    static __device__ __forceinline__ void sync() { __syncthreads(); }
    static __device__ __forceinline__ double sum(double v, double *scr) {
        v = warp_sum(v);
        if ((threadIdx.x & 31) == 0) scr[threadIdx.x >> 5] = v;
        __syncthreads();
        v = scr[0] + scr[1];
        __syncthreads();
        return v;
    }
    static __device__ __forceinline__ double max(double v, double *scr) {
        v = warp_max(v);
        if ((threadIdx.x & 31) == 0) scr[threadIdx.x >> 5] = v;
        __syncthreads();
        v = fmax(scr[0], scr[1]);
        __syncthreads();
        return v;
    }
    static __device__ __forceinline__ void argmin(double &v, int &idx, double *scr) {
        warp_argmin(v, idx);
        if ((threadIdx.x & 31) == 0) {
            scr[2 + 2 * (threadIdx.x >> 5)] = v;
            scr[3 + 2 * (threadIdx.x >> 5)] = (double)idx;
        }
        __syncthreads();
        const double v0 = scr[2], v1 = scr[4];
        const int i0 = (int)scr[3], i1 = (int)scr[5];
        const bool second = i1 >= 0 && (i0 < 0 || v1 < v0 || (v1 == v0 && i1 < i0));
        v = second ? v1 : v0;
        idx = second ? i1 : i0;
        __syncthreads();
    }
};

// Condensed QP, thread by row in registers: hrow <- row `lane` of 2 sum_i Su_i^T W Su_i (WITHOUT the 2 w_force of the diagonal),
// grow <- entry `lane` of g = 2 sum_i Su_i^T W (sc_i - ref_i).  `lane` is the thread's index in its group (MpcGroup<NW>).
// Unknown 3 (NFL i + r) + c = component c of free_leg[r] at stage i; a free_leg entry that is a swing leg is a phantom (no
// dynamics: its row and column are zero).
// scratch: shared memory, mpc_condense_scratch(NFL) doubles: the sensitivities Su [12][ld] of the stage state to the forces, the five
// rotation matrices of the horizon and its 60 reference values.
//   * the rotations (three sin / cos pairs each) are evaluated ONCE, stage i by thread i, instead of by every thread in every
//     stage; the reference states are fetched once, coalesced, instead of one dependent global load per use;
//   * rows 3..5 and 9..11 of Su (position and velocity sensitivities) are known in closed form - a force component c of stage j
//     changes velocity c by dt / m from stage j on and position c by (i - j) dt^2 / m at stage i, nothing else - so their part of
//     the Gram matrix, sum_i over the common stages of w_pos (i - j)(i - j') (dt^2 / m)^2 + w_vel (dt / m)^2 for equal components,
//     is added once at the end, and the per-stage accumulation runs over the six attitude / body-rate rows only (pairs of columns
//     per 16-byte load): half the FMAs and a third of the shared-memory loads of the straightforward form.
__host__ __device__ constexpr int mpc_condense_scratch(int nfl) { return 12 * mpcr_even(15 * nfl) + 48 + 64; }

template <int NFL, int NW = 1>
__device__ __forceinline__ void mpc_condense(const MpcParams &prm, long long prob, int lane, const int (&kind_leg)[4], const int (&free_leg)[4],
                                             double *scratch, double (&hrow)[15 * NFL], double &grow) {
    using G = MpcGroup<NW>;
    constexpr int n = 15 * NFL, ld = mpcr_even(n), GT = G::threads;
    const long long N = prm.N;
    const int row = lane < n ? lane : n - 1;
    double *Su = scratch, *Rs = scratch + 12 * ld, *bref = Rs + 48;  // Rs [5][9] (padded), bref [5][12]
    for (int e = lane; e < 12 * ld; e += GT) Su[e] = 0.0;
    for (int e = lane; e < MPC_NH * 12; e += GT) bref[e] = prm.body_ref[(long long)e * N + prob];
    grow = 0.0;
#pragma unroll
    for (int j = 0; j < n; ++j) hrow[j] = 0.0;
    double sc[12], pf[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) sc[k] = prm.x[k * N + prob];
#pragma unroll
    for (int k = 0; k < 12; ++k) pf[k] = prm.p[k * N + prob];
    bool real[NFL];  // leg r of the unknowns carries a force (not a phantom)
#pragma unroll
    for (int r = 0; r < NFL; ++r) {
        real[r] = false;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q == free_leg[r]) real[r] = kind_leg[q] != 0;
    }
    G::sync();
    if (lane < MPC_NH) {  // linearisation attitude of stage i: the current state for stage 0, body_ref[:, i - 1] after
        double R[9];
        if (lane == 0) rot_zyx(sc[0], sc[1], sc[2], R);
        else rot_zyx(bref[(lane - 1) * 12], bref[(lane - 1) * 12 + 1], bref[(lane - 1) * 12 + 2], R);
#pragma unroll
        for (int k = 0; k < 9; ++k) Rs[9 * lane + k] = R[k];
    }
    G::sync();
#pragma unroll 1
    for (int i = 0; i < MPC_NH; ++i) {
        double R[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = Rs[9 * i + k];
        // Su <- (I + dt A) Su: rows 0..2 += dt R^T rows 6..8, rows 3..5 += dt rows 9..11 (rows 6..11 unchanged)
        if (lane < n) {
            const int c = lane;
            const double w0 = Su[6 * ld + c], w1 = Su[7 * ld + c], w2 = Su[8 * ld + c];
#pragma unroll
            for (int a = 0; a < 3; ++a) Su[a * ld + c] += prm.dt * (R[0 * 3 + a] * w0 + R[1 * 3 + a] * w1 + R[2 * 3 + a] * w2);
#pragma unroll
            for (int a = 0; a < 3; ++a) Su[(3 + a) * ld + c] += prm.dt * Su[(9 + a) * ld + c];
        }
        G::sync();
        // Su[:, 3 NFL i + 3 r + c] += dt B: rows 6..8 = Ihat^-1 skew(R p_l), rows 9..11 = I / m;  Ihat^-1 = R diag(1/I) R^T
        if (lane < 3 * NFL) {
            const int c = lane % 3;
            int l = 0;
            bool real_leg = false;
#pragma unroll
            for (int r = 0; r < NFL; ++r)
                if (r == lane / 3) { l = free_leg[r]; real_leg = real[r]; }
            double pl[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q == l) { pl[0] = pf[3 * q]; pl[1] = pf[3 * q + 1]; pl[2] = pf[3 * q + 2]; }
            double pw[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) pw[a] = R[3 * a] * pl[0] + R[3 * a + 1] * pl[1] + R[3 * a + 2] * pl[2];
            double sk[3];  // column c of skew(pw) = [[0,-z,y],[z,0,-x],[-y,x,0]]
            sk[0] = c == 0 ? 0.0 : (c == 1 ? -pw[2] : pw[1]);
            sk[1] = c == 0 ? pw[2] : (c == 1 ? 0.0 : -pw[0]);
            sk[2] = c == 0 ? -pw[1] : (c == 1 ? pw[0] : 0.0);
            double t3[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) t3[a] = prm.inv_inertia[a] * (R[a] * sk[0] + R[3 + a] * sk[1] + R[6 + a] * sk[2]);
            const int col = 3 * NFL * i + lane;
            if (real_leg) {
#pragma unroll
                for (int a = 0; a < 3; ++a) Su[(6 + a) * ld + col] += prm.dt * (R[3 * a] * t3[0] + R[3 * a + 1] * t3[1] + R[3 * a + 2] * t3[2]);
                Su[(9 + c) * ld + col] += prm.dt * prm.inv_mass;
            }
        }
        // free response: sc <- (I + dt A) sc + dt g
        {
            const double w0 = sc[6], w1 = sc[7], w2 = sc[8];
#pragma unroll
            for (int a = 0; a < 3; ++a) sc[a] += prm.dt * (R[a] * w0 + R[3 + a] * w1 + R[6 + a] * w2);
#pragma unroll
            for (int a = 0; a < 3; ++a) sc[3 + a] += prm.dt * sc[9 + a];
            sc[11] += prm.dt * prm.gravity;
        }
        G::sync();
        // gradient: all twelve rows of this thread's column; Gram matrix: the attitude (0..2) and body-rate (6..8) rows, columns in pairs
        double sa[6], gacc = 0.0;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const double s_k = Su[k * ld + row];
            if (k < 3) sa[k] = prm.w_state[k] * s_k;
            if (k >= 6 && k < 9) sa[k - 3] = prm.w_state[k] * s_k;
            gacc = fma(s_k, prm.w_state[k] * (sc[k] - bref[i * 12 + k]), gacc);
        }
        grow = fma(2.0, gacc, grow);
        const int ncol = 3 * NFL * (i + 1);  // the later columns of Su are still zero
#pragma unroll
        for (int b = 0; b + 1 < n; b += 2) {
            if (b < ncol) {  // uniform; ncol is even or the odd last column is handled below
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const double2 v = *reinterpret_cast<const double2 *>(Su + (k < 3 ? k : k + 3) * ld + b);
                    a0 = fma(sa[k], v.x, a0);
                    a1 = fma(sa[k], v.y, a1);
                }
                hrow[b] = fma(2.0, a0, hrow[b]);
                hrow[b + 1] = fma(2.0, a1, hrow[b + 1]);
            }
        }
        if (n & 1) {
            double a0 = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) a0 = fma(sa[k], Su[(k < 3 ? k : k + 3) * ld + n - 1], a0);
            hrow[n - 1] = fma(2.0, a0, hrow[n - 1]);
        }
        G::sync();
    }
    // closed-form part: position and velocity rows (see the header)
    {
        const int jr = row / (3 * NFL), rr = (row / 3) % NFL, cr = row % 3;
        bool real_row = false;
#pragma unroll
        for (int r = 0; r < NFL; ++r)
            if (r == rr) real_row = real[r];
        const double d2m = prm.dt * prm.dt * prm.inv_mass, d1m = prm.dt * prm.inv_mass;
        double wp = 0.0, wv = 0.0;  // 2 w_pos (dt^2 / m)^2 and 2 w_vel (dt / m)^2 of this row's component
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (c == cr) { wp = 2.0 * prm.w_state[3 + c] * d2m * d2m; wv = 2.0 * prm.w_state[9 + c] * d1m * d1m; }
#pragma unroll
        for (int b = 0; b < n; ++b) {
            const int jb = b / (3 * NFL), rb = (b / 3) % NFL, cb = b % 3;  // compile-time after unrolling
            int s1 = 0, cnt = 0;  // sum over the common stages i >= max(jr, jb) of (i - jr)(i - jb), and their number
#pragma unroll
            for (int i = 0; i < MPC_NH; ++i)
                if (i >= jb && i >= jr) { s1 += (i - jr) * (i - jb); ++cnt; }
            if (cb == cr && real_row && real[rb]) hrow[b] += wp * (double)s1 + wv * (double)cnt;
        }
    }
}

// bit of constraint row r of leg l in a warm-start word (one word per stage); bits 24..27: legs out of swing; bit 31: valid
__device__ __forceinline__ uint32_t warm_bit(int leg, int r) { return 1u << (5 * leg + r); }

template <int NFL>
__device__ __forceinline__ void mpc_solve_rows(const MpcParams &prm, long long prob, int lane, double *base, const int (&kind_leg)[4],
                                               const int (&free_leg)[4]) {
    constexpr int n = 15 * NFL, nb = 5 * NFL, ntri = n * (n + 1) / 2, ld = mpcr_even(n);
    const long long N = prm.N;
    static_assert(mpc_condense_scratch(NFL) <= n * (n | 1), "the scratch of the condensation fits the factor's storage");
    double *H = base, *Ls = H + mpcr_even(ntri), *Su = Ls, *u = Ls + mpcr_even(n * (n | 1)), *g = u + ld, *rv = g + ld, *dv = rv + ld, *ukeep = dv + ld;
    double *strip = ukeep + ld, *blk = strip + 128;

    int my_leg = 0;  // leg of the block this lane owns (blocks: stage-major, NFL legs per stage)
#pragma unroll
    for (int r = 0; r < NFL; ++r)
        if (r == lane % NFL) my_leg = free_leg[r];
    int my_kind = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l)
        if (l == my_leg) my_kind = kind_leg[l];
    if (lane >= nb) my_kind = 0;
    const bool my_act = my_kind == 1;
    const int my_stage = lane / NFL;
    const double mu_f = prm.mu;
    const int row = lane < n ? lane : n - 1;

    // ---- condensed QP (mpc_condense): this lane's row of H in M.m, its entry of g in grow ----------------------------------
    if (lane < n) u[lane] = ukeep[lane] = 0.0;
    for (int e = lane; e < 128; e += 32) strip[e] = 0.0;  // the entries beyond n stay zero: the window's tail multiplies them
    MpcRows<n> M;  // its row doubles as the accumulator of this lane's row of H
    double grow;
    mpc_condense<NFL>(prm, prob, lane, kind_leg, free_leg, Su, M.m, grow);
    if (lane < n) {
#pragma unroll
        for (int j = 0; j < n; ++j)
            if (j <= lane) H[lane * (lane + 1) / 2 + j] = M.m[j] + (j == lane ? 2.0 * prm.w_force : 0.0);
        g[lane] = grow;
    }
    __syncwarp();
    double hmax = 0.0, gmax = 0.0;
    if (lane < n) {
        hmax = H[tri_idx(lane, lane)];
        gmax = fabs(grow);
    }
    hmax = warp_max(hmax);
    const double gs = fmax(warp_max(gmax), 1e-300);
    const double rho = 1e2 * hmax;
    const double bvec[5] = {prm.fz_max, 0.0, 0.0, 0.0, 0.0};
    const double m_act = warp_sum(my_act ? 5.0 : 0.0);
    const double inv_m_act = m_act > 0.0 ? 1.0 / m_act : 0.0;
    const double inv_gs = 1.0 / gs, inv_fz = 1.0 / prm.fz_max, gs_over_fz = gs * inv_fz;

    // ---- solver state ---------------------------------------------------------------------------------------------------
    // phase 0: polish from the warm-start set; 1: interior point; 2: polish from the interior point's set; 3: done
    double s[5], lam[5], lw[5];
    bool W[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) { s[r] = 1.0; lam[r] = 0.0; lw[r] = 0.0; W[r] = false; }
    uint32_t pattern = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) pattern |= kind_leg[l] != 0 ? (1u << (24 + l)) : 0u;
    int phase = m_act > 0.0 ? 1 : 2;  // nothing constrained: the empty active set is the answer (one factorisation, one solve)
    if (prm.warm_set && m_act > 0.0) {
        const uint32_t word = lane < nb ? prm.warm_set[(long long)my_stage * N + prob] : 0u;
        const uint32_t word0 = __shfl_sync(0xffffffffu, word, 0);
        if ((word0 >> 31) && (word0 & 0x0f000000u) == pattern) {
            phase = 0;
            if (my_act) {
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    W[r] = (word & warm_bit(my_leg, r)) != 0;
                    lw[r] = W[r] ? prm.warm_mult[(long long)((my_stage * 4 + my_leg) * 5 + r) * N + prob] : 0.0;
                }
            }
        }
    }
    uint32_t status = phase == 0 ? 8u : 0u;
    int it = 0, rounds = 0;
    bool start_ipm = phase == 1;

#pragma unroll 1
    while (phase != 3) {
        double d[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, rp[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        double mu_c = 0.0;
        if (phase == 1) {
            if (start_ipm) {  // strictly inside every pyramid
                start_ipm = false;
                if (lane < n) u[lane] = 0.0;
                __syncwarp();
                if (my_act) u[3 * lane + 2] = fmin(10.0, 0.5 * prm.fz_max);
                __syncwarp();
                double au[5];
                rows_times(u + 3 * (lane < nb ? lane : 0), mu_f, au);
#pragma unroll
                for (int r = 0; r < 5; ++r) { s[r] = my_act ? bvec[r] - au[r] : 1.0; lam[r] = my_act ? rcp2_(s[r]) : 0.0; }
            }
            // residuals: rd = H u + g + A^T lam, rp = A u + s - b
            double acc0 = 0.0, acc1 = 0.0;
            {
                const double *hb = H + row * (row + 1) / 2;
#pragma unroll 2
                for (int j = 0; j < n; j += 2) {
                    acc0 = fma(j <= row ? hb[j] : H[j * (j + 1) / 2 + row], u[j], acc0);
                    acc1 = fma(j + 1 <= row ? hb[j + 1] : H[(j + 1) * (j + 2) / 2 + row], u[j + 1], acc1);
                }
            }
            if (lane < n) rv[lane] = (acc0 + acc1) + g[lane];
            __syncwarp();
            if (my_act) {
                double au[5], atl[3];
                rows_times(u + 3 * lane, mu_f, au);
#pragma unroll
                for (int r = 0; r < 5; ++r) rp[r] = au[r] + s[r] - bvec[r];
                rows_transpose_times(lam, mu_f, atl);
#pragma unroll
                for (int c = 0; c < 3; ++c) rv[3 * lane + c] += atl[c];
            }
            __syncwarp();
            double rdmax = lane < n ? fabs(rv[lane]) : 0.0, rpmax = 0.0, comp = 0.0;
#pragma unroll
            for (int r = 0; r < 5; ++r) { rpmax = fmax(rpmax, fabs(rp[r])); comp += my_act ? s[r] * lam[r] : 0.0; }
            rdmax = warp_max(rdmax) * inv_gs;
            rpmax = warp_max(rpmax) * inv_fz;
            mu_c = warp_sum(comp) * inv_m_act;
            const bool converged = it > 0 && mu_c < 1e-9 && fmax(rdmax, rpmax) < 1e-8;
            if (converged || it >= MPC_MAX_IPM) {
                if (!converged) status |= 1u;
                phase = 2;
                rounds = 0;
                if (lane < n) ukeep[lane] = u[lane];
#pragma unroll
                for (int r = 0; r < 5; ++r) { W[r] = my_act && s[r] * gs_over_fz < lam[r]; lw[r] = W[r] ? lam[r] : 0.0; }
            }
        }
        // ---- M = H + A^T diag(d) A, factorised ------------------------------------------------------------------------
        double rs[5];  // 1 / slack (interior point only; slacks stay positive)
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            rs[r] = phase == 1 ? rcp2_(s[r]) : 0.0;
            d[r] = phase == 1 ? (my_act ? lam[r] * rs[r] : 0.0) : (W[r] ? rho : 0.0);
        }
        if (lane < nb) {
            double gm[6];
            rows_gram(d, mu_f, gm);
#pragma unroll
            for (int q = 0; q < 6; ++q) blk[6 * lane + q] = gm[q];
        }
        __syncwarp();
        M.form(H, blk, lane);
        const bool pd = M.factor(strip, Ls, lane);
        if (!pd) {
            if (phase == 1) {  // hand what there is to the polish phase, flagged
                status |= 1u;
                phase = 2;
                rounds = 0;
                if (lane < n) ukeep[lane] = u[lane];
#pragma unroll
                for (int r = 0; r < 5; ++r) { W[r] = my_act && s[r] * gs_over_fz < lam[r]; lw[r] = W[r] ? lam[r] : 0.0; }
                continue;
            }
            if (phase == 0) { phase = 1; start_ipm = true; status &= ~8u; continue; }
            if (lane < n) u[lane] = ukeep[lane];
            status |= 2u;
            __syncwarp();
            break;
        }
        if (phase == 1) {
            // Mehrotra predictor-corrector: pass 0 affine direction (rc = s lam), pass 1 corrector (rc = s lam + ds dl - sigma mu)
            double ds[5] = {0, 0, 0, 0, 0}, dl[5] = {0, 0, 0, 0, 0}, rc[5];
            double alpha_aff_p = 1.0, alpha_aff_d = 1.0;
            const double rvi = lane < n ? rv[lane] : 0.0;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                double sigma_mu = 0.0;
                if (pass == 1) {
                    double comp_aff = 0.0;
#pragma unroll
                    for (int r = 0; r < 5; ++r) comp_aff += my_act ? (s[r] + alpha_aff_p * ds[r]) * (lam[r] + alpha_aff_d * dl[r]) : 0.0;
                    const double ratio = warp_sum(comp_aff) * inv_m_act * rcp2_(mu_c);
                    sigma_mu = ratio * ratio * ratio * mu_c;
                }
#pragma unroll
                for (int r = 0; r < 5; ++r) rc[r] = s[r] * lam[r] + (pass == 1 ? ds[r] * dl[r] - sigma_mu : 0.0);
                if (lane < n) dv[lane] = -rvi;
                __syncwarp();
                if (my_act) {
                    double t[5], att[3];
#pragma unroll
                    for (int r = 0; r < 5; ++r) t[r] = (-rc[r] + lam[r] * rp[r]) * rs[r];
                    rows_transpose_times(t, mu_f, att);
#pragma unroll
                    for (int c = 0; c < 3; ++c) dv[3 * lane + c] -= att[c];
                }
                __syncwarp();
                const double x = M.solve(lane < n ? dv[lane] : 0.0, Ls, lane);  // (surplus lanes carry a dummy entry)
                if (lane < n) dv[lane] = x;
                __syncwarp();
                if (my_act) {
                    double adu[5];
                    rows_times(dv + 3 * lane, mu_f, adu);
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        ds[r] = -rp[r] - adu[r];
                        dl[r] = (-rc[r] - lam[r] * ds[r]) * rs[r];
                    }
                }
                // largest steps that keep s and lam positive: the smallest ratio s / -ds (lam / -dl) as a fraction, one
                // reciprocal per lane instead of ten divisions
                double pn = 1.0, pd_ = 1.0, dn = 1.0, dd = 1.0;
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    if (my_act && ds[r] < 0.0 && s[r] * pd_ < -ds[r] * pn) { pn = s[r]; pd_ = -ds[r]; }
                    if (my_act && dl[r] < 0.0 && lam[r] * dd < -dl[r] * dn) { dn = lam[r]; dd = -dl[r]; }
                }
                const double ap = warp_min(pn * rcp2_(pd_)), ad = warp_min(dn * rcp2_(dd));
                if (pass == 0) { alpha_aff_p = ap; alpha_aff_d = ad; }
                else {
                    const double a = fmin(fmin(1.0, 0.995 * ap), fmin(1.0, 0.995 * ad));
                    if (lane < n) u[lane] = fma(a, x, u[lane]);
#pragma unroll
                    for (int r = 0; r < 5; ++r) { s[r] = fma(a, ds[r], s[r]); lam[r] = fma(a, dl[r], lam[r]); }
                }
                __syncwarp();
            }
            ++it;
            continue;
        }
        // ---- polish round: method of multipliers on the active set W, then verification / correction of the set -------
        double au[5] = {0, 0, 0, 0, 0};
        const double gi = lane < n ? g[lane] : 0.0;
#pragma unroll 1
        for (int k = 0; k < MPC_MOM_ITERS; ++k) {
            if (lane < n) dv[lane] = -gi;
            __syncwarp();
            if (my_act) {
                double t[5], att[3];
#pragma unroll
                for (int r = 0; r < 5; ++r) t[r] = W[r] ? -lw[r] + rho * bvec[r] : 0.0;
                rows_transpose_times(t, mu_f, att);
#pragma unroll
                for (int c = 0; c < 3; ++c) dv[3 * lane + c] += att[c];
            }
            __syncwarp();
            const double x = M.solve(lane < n ? dv[lane] : 0.0, Ls, lane);  // (surplus lanes carry a dummy entry)
            if (lane < n) dv[lane] = x;
            __syncwarp();
            double moved = 0.0;
            if (my_act) {
                rows_times(dv + 3 * lane, mu_f, au);
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    const double dl_r = W[r] ? rho * (au[r] - bvec[r]) : 0.0;
                    lw[r] += dl_r;
                    moved = fmax(moved, fabs(dl_r));
                }
            }
            __syncwarp();
            if (warp_max(moved) <= 1e-12 * gs) break;  // multipliers settled (always after one pass when W is empty)
        }
        bool infeas[5], neg[5], any_inf = false, any_neg = false;
        double worst_inf = 0.0, worst_neg = 0.0;
        int r_inf = -1, r_neg = -1;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            infeas[r] = my_act && !W[r] && au[r] - bvec[r] > 1e-9 * prm.fz_max;
            neg[r] = W[r] && lw[r] < -1e-9 * gs;
            any_inf |= infeas[r];
            any_neg |= neg[r];
            if (infeas[r] && au[r] - bvec[r] > worst_inf) { worst_inf = au[r] - bvec[r]; r_inf = r; }
            if (neg[r] && lw[r] < worst_neg) { worst_neg = lw[r]; r_neg = r; }
        }
        const bool warp_inf = __any_sync(0xffffffffu, any_inf), warp_neg = __any_sync(0xffffffffu, any_neg);
        if (!warp_inf && !warp_neg) {
            if (lane < n) u[lane] = dv[lane];
            __syncwarp();
            phase = 3;
            break;
        }
        ++rounds;
        if (phase == 0 && rounds >= (prm.warm_rounds > 0 ? prm.warm_rounds : MPCR_WARM_ROUNDS)) {  // the previous set does not carry over: interior point from scratch
            phase = 1;
            start_ipm = true;
            status &= ~8u;
            continue;
        }
        if (phase == 2 && rounds >= MPC_POLISH_ROUNDS) {
            if (lane < n) u[lane] = ukeep[lane];
            status |= 2u;
            __syncwarp();
            break;
        }
        // The first two rounds take every violated constraint in and every negative multiplier out at once; after that ONE
        // change per round - the most violated constraint, or else the most negative multiplier - because simultaneous
        // changes can cycle.
        if (rounds <= 2) {
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                if (infeas[r]) { W[r] = true; lw[r] = 0.0; }
                if (neg[r]) { W[r] = false; lw[r] = 0.0; }
            }
        } else {
            const double mine = warp_inf ? worst_inf : -worst_neg;
            const double worst = warp_max(mine);
            const unsigned cand = __ballot_sync(0xffffffffu, (warp_inf ? r_inf : r_neg) >= 0 && mine == worst);
            if (lane == __ffs(cand) - 1) {
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    if (warp_inf && r == r_inf) { W[r] = true; lw[r] = 0.0; }
                    if (!warp_inf && r == r_neg) { W[r] = false; lw[r] = 0.0; }
                }
            }
        }
    }

    // ---- results ---------------------------------------------------------------------------------------------------------
    for (int e = lane; e < MPC_N; e += 32) prm.forces[(long long)e * N + prob] = 0.0;
    __syncwarp();
    if (lane < n) {  // compact -> (stage, leg, component)
        const int stage = lane / (3 * NFL), within = lane % (3 * NFL);
        int l = 0;
#pragma unroll
        for (int r = 0; r < NFL; ++r)
            if (r == within / 3) l = free_leg[r];
        prm.forces[(long long)(12 * stage + 3 * l + within % 3) * N + prob] = u[lane];
    }
    if (prm.warm_set) {
        // active set and multipliers of this solve, by absolute (stage, leg) so that they survive a change of numbering
        const bool good = phase == 3 && m_act > 0.0;
        uint32_t bits = 0;
#pragma unroll
        for (int r = 0; r < 5; ++r) bits |= (good && W[r]) ? warm_bit(my_leg, r) : 0u;
#pragma unroll
        for (int o = 1; o < NFL; o <<= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);  // the legs of one stage sit in adjacent lanes
        if (lane < nb && lane % NFL == 0) prm.warm_set[(long long)my_stage * N + prob] = good ? (bits | pattern | 0x80000000u) : 0u;
        if (lane < nb && good && my_act) {
#pragma unroll
            for (int r = 0; r < 5; ++r) prm.warm_mult[(long long)((my_stage * 4 + my_leg) * 5 + r) * N + prob] = W[r] ? lw[r] : 0.0;
        }
    }
    if (prm.status && lane == 0) prm.status[prob] = status | ((uint32_t)it << 8);
}

__global__ void __launch_bounds__(32 * MPCR_WARPS, 3) kf_mpc_rows_kernel(const __grid_constant__ MpcParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long prob = (long long)blockIdx.x * MPCR_WARPS + warp;
    if (prob >= prm.N) return;  // whole warp
    if (prm.only_flagged && prm.status[prob] != MPC_ST_GIVEN_UP) return;  // second launch behind the dual active-set kernel
    double *base = reinterpret_cast<double *>(smem_raw) + (size_t)warp * mpcr_warp_doubles(MPCR_MAX_LEGS);
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
        if (prm.status && lane == 0) prm.status[prob] = 4u;
        if (prm.warm_set && lane < MPC_NH) prm.warm_set[(long long)lane * N + prob] = 0u;
        return;
    }
    if (nfl == 0) {
        for (int e = lane; e < MPC_N; e += 32) prm.forces[(long long)e * N + prob] = 0.0;
        if (prm.status && lane == 0) prm.status[prob] = 0u;
        if (prm.warm_set && lane < MPC_NH) prm.warm_set[(long long)lane * N + prob] = 0u;
        return;
    }
    if (nfl == 1) {  // one instantiation serves both: the lone leg is paired with a phantom that decouples exactly (see the B columns)
#pragma unroll
        for (int l = 3; l >= 0; --l)
            if (kind_leg[l] == 0) free_leg[1] = l;
    }
    mpc_solve_rows<2>(prm, prob, lane, base, kind_leg, free_leg);
}

}  // namespace okf
