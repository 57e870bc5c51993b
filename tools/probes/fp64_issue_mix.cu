// Developer probe: do other instructions issue underneath FP64 instructions, or does a warp DFMA hold the issue port of its scheduler?
// Per loop trip every thread runs 16 independent DFMAs and kOther independent instructions of another pipe (integer IMAD, FP32 FFMA, or
// shared-memory loads), with 1, 2 or 3 warps per scheduler.  If the time per trip stays at 16 x the DFMA interval while kOther grows, the
// other instructions are free (the bound of a kernel is its FP64 instruction count alone); if it grows by ~1 cycle per other
// instruction, a kernel's bound is  n_fp64 x interval + n_other.
#include <cstdio>
#include <cuda_runtime.h>

template <int kOther, int kKind>
__global__ void k(long long iters, double seed, int iseed, double *sink) {
    __shared__ int sh[1024];
    sh[threadIdx.x] = iseed + threadIdx.x;
    __syncthreads();
    double a[16];
    int n[16];
    float f[16];
    const double m = 0.999 + seed * 1e-9, c = 1e-3 * (threadIdx.x & 7);
    const float fm = 0.999f + (float)seed * 1e-7f, fc = 1e-3f * (threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed + i; n[i] = iseed + i + threadIdx.x; f[i] = (float)seed + i + threadIdx.x; }
    int ld = 0;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            a[i] = a[i] * m + c;
#pragma unroll
            for (int j = 0; j < kOther / 16; ++j) {
                if (kKind == 0) n[(i + j) & 15] = n[(i + j) & 15] * n[(i + j + 1) & 15] + 7;
                else if (kKind == 1) f[(i + j) & 15] = f[(i + j) & 15] * fm + fc;
                else ld += sh[(threadIdx.x + 32 * (i + j) + (int)it) & 1023];
            }
        }
    }
    double s = (double)ld;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + n[i] + f[i];
    if (s == -12345.678) sink[threadIdx.x] = s;
}

template <int kOther, int kKind>
void run(const char *kind) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps_per_sched : {1, 2, 3}) {
        const int threads = 32 * 4 * warps_per_sched;
        const long long iters = 100000;
        k<kOther, kKind><<<148, threads>>>(1000, 1.0, 3, nullptr);
        cudaEventRecord(e0);
        k<kOther, kKind><<<148, threads>>>(iters, 1.0, 3, nullptr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * warps_per_sched);
        printf("16 DFMA + %2d %-5s per trip, %d warp(s) per scheduler: %7.2f cycles per warp-trip  (%.2f per DFMA if the rest were free; %.2f left per other instruction at 2.18 per DFMA)\n",
               kOther, kind, warps_per_sched, cyc, cyc / 16, kOther ? (cyc - 16 * 2.18) / kOther : 0.0);
    }
}

int main() {
    run<0, 0>("none");
    run<16, 0>("IMAD"); run<32, 0>("IMAD"); run<64, 0>("IMAD");
    run<16, 1>("FFMA"); run<32, 1>("FFMA"); run<64, 1>("FFMA");
    run<16, 2>("LDS"); run<32, 2>("LDS");
    return 0;
}
