// Developer probe: does the FP64 pipe's time per warp instruction depend on the number of active lanes?
// One warp per SM sub-partition (latency-free: 16 independent chains), lanes >= `active` exit at once.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long iters, int active, double seed, double *sink) {
    if ((threadIdx.x & 31) >= active) return;
    double a[16];
    const double m = 0.999 + seed * 1e-9, c = 1e-3 * (threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = a[i] * m + c;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == -12345.678) sink[threadIdx.x] = s;
}
int main() {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps_per_block : {4, 16}) {
        for (int active : {32, 16, 8, 4, 1}) {
            k<<<148, 32 * warps_per_block>>>(1000, active, 1.0, nullptr);
            cudaEventRecord(e0);
            k<<<148, 32 * warps_per_block>>>(200000, active, 1.0, nullptr);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("warps/SM %2d active lanes %2d: %.3f ms  (%.2f cycles per warp FMA per scheduler at 1.965 GHz)\n", warps_per_block, active, ms,
                   ms * 1e-3 * 1.965e9 / (200000.0 * 16 * warps_per_block / 4));
        }
    }
    return 0;
}
