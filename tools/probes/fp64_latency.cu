// Developer probe: dependent-issue latency of DFMA / DMUL / DADD, of a 64-bit shuffle and of a shared-memory round trip (one warp).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(long long iters, double seed, double *sink, long long *cycles) {
    __shared__ double sh[64];
    double a = seed, m = 0.999 + seed * 1e-9, c = 1e-3;
    sh[threadIdx.x] = seed;
    __syncwarp();
    long long t0 = clock64();
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a = fma(a, m, c);
            else if (MODE == 1) a = a * m;
            else if (MODE == 2) a = a + c;
            else if (MODE == 3) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
            else { sh[threadIdx.x] = a; __syncwarp(); a = sh[(threadIdx.x + 1) & 31] + c; __syncwarp(); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    if (a == -12345.678) *sink = a;
}
template <int MODE>
void run(const char *name, double *sink, long long *cyc) {
    k<MODE><<<1, 32>>>(100, 1.0, sink, cyc); cudaDeviceSynchronize();
    k<MODE><<<1, 32>>>(10000, 1.0, sink, cyc); cudaDeviceSynchronize();
    printf("%-42s %.1f cycles per dependent op\n", name, (double)*cyc / (10000.0 * 16));
}
int main() {
    long long *cyc; double *sink;
    cudaMallocManaged(&cyc, 8); cudaMallocManaged(&sink, 8);
    run<0>("DFMA", sink, cyc); run<1>("DMUL", sink, cyc); run<2>("DADD", sink, cyc); run<3>("64-bit SHFL", sink, cyc);
    run<4>("STS + syncwarp + LDS + DADD + syncwarp", sink, cyc);
    return 0;
}
