// Scalar abstraction of the filter kernels: the same source runs on `double`, `float` and `F2` - two FP32
// trajectories packed into one thread, computed with the sm_100 packed instructions FFMA2 / FADD2 / FMUL2
// (one issue slot, two FMAs).  The hot loops are written with explicit fma_/fnma_ so that all three types
// execute the same operation sequence per trajectory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace okf {

struct F2 {
    float2 v;
    __device__ __forceinline__ F2() {}
    __device__ __forceinline__ F2(float s) : v(make_float2(s, s)) {}
    __device__ __forceinline__ F2(float a, float b) : v(make_float2(a, b)) {}
    __device__ __forceinline__ explicit F2(float2 f) : v(f) {}
};

template <typename Real> struct Lanes { static constexpr int n = 1; using scalar = Real; };
template <> struct Lanes<F2> { static constexpr int n = 2; using scalar = float; };

// ---- F2 operators (not fused: generic code outside the hot loops) ------------------------------------------
__device__ __forceinline__ F2 operator-(F2 a) { return F2(-a.v.x, -a.v.y); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return F2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return F2(__fadd2_rn(a.v, (-b).v)); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return F2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 &operator+=(F2 &a, F2 b) { a = a + b; return a; }
__device__ __forceinline__ F2 &operator-=(F2 &a, F2 b) { a = a - b; return a; }
__device__ __forceinline__ F2 &operator*=(F2 &a, F2 b) { a = a * b; return a; }

// ---- fused / elementary operations ------------------------------------------------------------------------------
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ F2 fma_(F2 a, F2 b, F2 c) { return F2(__ffma2_rn(a.v, b.v, c.v)); }
// c - a*b
__device__ __forceinline__ double fnma_(double a, double b, double c) { return fma(-a, b, c); }
__device__ __forceinline__ float fnma_(float a, float b, float c) { return fmaf(-a, b, c); }
__device__ __forceinline__ F2 fnma_(F2 a, F2 b, F2 c) { return F2(__ffma2_rn((-a).v, b.v, c.v)); }

// Reciprocal of a pivot (positive, normal; anything else is flagged NOT_PD by the caller): hardware seed (MUFU.RCP64H /
// MUFU.RCP) refined by Newton steps, straight-line.  The IEEE division `1.0 / a` costs the same FMAs plus a
// special-case branch with a slow-path call, and that branch splits the time loop into basic blocks across which the
// scheduler cannot overlap the reciprocal chain with the rank-1 FMAs.  Relative error <= 2 ulp.
__device__ __forceinline__ double rcp_(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));  // ~20 good bits
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);  // ~40 bits
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);  // ~80 bits -> rounding-limited
    e = fma(-a, r, 1.0);
    return fma(r, e, r);  // one more step removes the residual of the seed's worst case
}
__device__ __forceinline__ float rcp_(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    const float e = fmaf(-a, r, 1.0f);
    return fmaf(r, e, r);
}
__device__ __forceinline__ F2 rcp_(F2 a) {
    F2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.v.x) : "f"(a.v.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.v.y) : "f"(a.v.y));
    const F2 e = F2(__ffma2_rn((-a).v, r.v, make_float2(1.0f, 1.0f)));
    return F2(__ffma2_rn(r.v, e.v, r.v));
}
__device__ __forceinline__ double div_(double a, double b) { return a / b; }
__device__ __forceinline__ float div_(float a, float b) { return a / b; }
__device__ __forceinline__ F2 div_(F2 a, F2 b) { return F2(a.v.x / b.v.x, a.v.y / b.v.y); }
__device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ F2 max_(F2 a, F2 b) { return F2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y)); }
__device__ __forceinline__ double sqrt_(double a) { return sqrt(a); }
__device__ __forceinline__ float sqrt_(float a) { return sqrtf(a); }
__device__ __forceinline__ F2 sqrt_(F2 a) { return F2(sqrtf(a.v.x), sqrtf(a.v.y)); }

// per-lane access
__device__ __forceinline__ double lane_get(double a, int) { return a; }
__device__ __forceinline__ float lane_get(float a, int) { return a; }
__device__ __forceinline__ float lane_get(F2 a, int l) { return l == 0 ? a.v.x : a.v.y; }
__device__ __forceinline__ void lane_set(double &a, int, double s) { a = s; }
__device__ __forceinline__ void lane_set(float &a, int, float s) { a = s; }
__device__ __forceinline__ void lane_set(F2 &a, int l, float s) { if (l == 0) a.v.x = s; else a.v.y = s; }

// Magnitude of a value as an ordered integer (high word of a double, the word of a float, sign cleared; the larger of the
// two lanes of an F2): |a| >= |b| <=> abs_bits(a) >= abs_bits(b) up to the low word of a double.  The status checks and
// the trunc pre-test compare these on the integer pipe instead of spending FP64-pipe slots on DSETP.
__device__ __forceinline__ int abs_bits(double a) { return __double2hiint(a) & 0x7fffffff; }
__device__ __forceinline__ int abs_bits(float a) { return __float_as_int(a) & 0x7fffffff; }
__device__ __forceinline__ int abs_bits(F2 a) { return max(abs_bits(a.v.x), abs_bits(a.v.y)); }

// pivot check: s must be positive and finite (OPTI_KF_ST_NOT_PD otherwise), per lane.  On the bit pattern: the (high) word
// of a positive finite number lies in [1, 0x7fefffff] (double) / [1, 0x7f7fffff] (float).
__device__ __forceinline__ bool bad_pivot_scalar(double s) { return (unsigned)(__double2hiint(s) - 1) >= 0x7fefffffu; }
__device__ __forceinline__ bool bad_pivot_scalar(float s) { return (unsigned)(__float_as_int(s) - 1) >= 0x7f7fffffu; }
__device__ __forceinline__ void note_bad_pivot(double s, uint32_t (&st)[1], uint32_t bit) { if (bad_pivot_scalar(s)) st[0] |= bit; }
__device__ __forceinline__ void note_bad_pivot(float s, uint32_t (&st)[1], uint32_t bit) { if (bad_pivot_scalar(s)) st[0] |= bit; }
__device__ __forceinline__ void note_bad_pivot(F2 s, uint32_t (&st)[2], uint32_t bit) {
    if (bad_pivot_scalar(s.v.x)) st[0] |= bit;
    if (bad_pivot_scalar(s.v.y)) st[1] |= bit;
}
__device__ __forceinline__ void note_nonfinite(double x, uint32_t (&st)[1], uint32_t bit) { if (abs_bits(x) >= 0x7ff00000) st[0] |= bit; }
__device__ __forceinline__ void note_nonfinite(float x, uint32_t (&st)[1], uint32_t bit) { if (abs_bits(x) >= 0x7f800000) st[0] |= bit; }
__device__ __forceinline__ void note_nonfinite(F2 x, uint32_t (&st)[2], uint32_t bit) {
    if (abs_bits(x.v.x) >= 0x7f800000) st[0] |= bit;
    if (abs_bits(x.v.y) >= 0x7f800000) st[1] |= bit;
}

// ---- running sums of the summary ---------------------------------------------------------------------------------
// double kernel: FP64 sums; float kernel: FP64 sums (the FP64 pipe is idle there); F2 kernel: packed FP32 sums
template <typename Real> struct Acc { using type = double; };
template <> struct Acc<F2> { using type = F2; };
__device__ __forceinline__ double acc_zero(double) { return 0.0; }
__device__ __forceinline__ F2 acc_zero(F2) { return F2(0.f); }
__device__ __forceinline__ double to_acc(double v) { return v; }
__device__ __forceinline__ double to_acc(float v) { return (double)v; }
__device__ __forceinline__ F2 to_acc(F2 v) { return v; }
__device__ __forceinline__ double err_sq(double x, double lab) { const double e = x - lab; return e * e; }
__device__ __forceinline__ double err_sq(float x, float lab) { const double e = (double)x - (double)lab; return e * e; }
__device__ __forceinline__ F2 err_sq(F2 x, F2 lab) { const F2 e = x - lab; return e * e; }
// acc + (x - lab)^2, fused
__device__ __forceinline__ double err_acc(double acc, double x, double lab) { const double e = x - lab; return fma(e, e, acc); }
__device__ __forceinline__ double err_acc(double acc, float x, float lab) { const double e = (double)x - (double)lab; return fma(e, e, acc); }
__device__ __forceinline__ F2 err_acc(F2 acc, F2 x, F2 lab) { const F2 e = x - lab; return fma_(e, e, acc); }
__device__ __forceinline__ double rms_out(double acc, double inv_t, double) { return sqrt(acc * inv_t); }
__device__ __forceinline__ float rms_out(double acc, double inv_t, float) { return (float)sqrt(acc * inv_t); }
__device__ __forceinline__ F2 rms_out(F2 acc, double inv_t, F2) { return F2(sqrtf(acc.v.x * (float)inv_t), sqrtf(acc.v.y * (float)inv_t)); }
__device__ __forceinline__ double mean_out(double acc, double inv_t, double) { return acc * inv_t; }
__device__ __forceinline__ float mean_out(double acc, double inv_t, float) { return (float)(acc * inv_t); }
__device__ __forceinline__ F2 mean_out(F2 acc, double inv_t, F2) { return acc * F2((float)inv_t); }

// ---- per-trajectory global arrays [C][N]: thread `first` = index of its first trajectory --------------------------
__device__ __forceinline__ double ld_traj(const double *base, long long idx, double) { return base[idx]; }
__device__ __forceinline__ float ld_traj(const float *base, long long idx, float) { return base[idx]; }
__device__ __forceinline__ F2 ld_traj(const float *base, long long idx, F2) { return F2(*reinterpret_cast<const float2 *>(base + idx)); }
__device__ __forceinline__ void st_traj(double *base, long long idx, double v) { __stcs(base + idx, v); }
__device__ __forceinline__ void st_traj(float *base, long long idx, float v) { __stcs(base + idx, v); }
__device__ __forceinline__ void st_traj(float *base, long long idx, F2 v) { __stcs(reinterpret_cast<float2 *>(base + idx), v.v); }

}  // namespace okf
