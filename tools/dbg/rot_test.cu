#include <cstdio>
#include "../../optistate_b200/csrc/kf_common.cuh"
using namespace okf;
__global__ void k(const float* ang, const float* p, float* out) {
    int i = threadIdx.x;
    float a = ang[3*i], b = ang[3*i+1], c = ang[3*i+2];
    float R[9]; rot_zyx(a, b, c, R);
    F2 R2[9]; rot_zyx(F2(a, a), F2(b, b), F2(c, c), R2);
    float px = p[0], py = p[1], pz = p[2];
    float pw = fma_(R[2], pz, fma_(R[1], py, R[0] * px));
    F2 pw2 = fma_(R2[2], F2(pz), fma_(R2[1], F2(py), R2[0] * F2(px)));
    int nd = 0;
    for (int k2 = 0; k2 < 9; ++k2) nd += (R[k2] != R2[k2].v.x) + (R[k2] != R2[k2].v.y);
    out[3*i] = nd; out[3*i+1] = pw; out[3*i+2] = pw2.v.x;
}
int main() {
    float ha[96], hp[3] = {0.2025f, 0.1479f, -0.2913f}, ho[96];
    for (int i = 0; i < 96; ++i) ha[i] = 0.01f * sinf(1.7f * i) + 1e-3f * i;
    float *da, *dp, *dout; cudaMalloc(&da, sizeof ha); cudaMalloc(&dp, sizeof hp); cudaMalloc(&dout, sizeof ho);
    cudaMemcpy(da, ha, sizeof ha, cudaMemcpyHostToDevice); cudaMemcpy(dp, hp, sizeof hp, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(da, dp, dout); cudaMemcpy(ho, dout, sizeof ho, cudaMemcpyDeviceToHost);
    int bad = 0, badp = 0;
    for (int i = 0; i < 32; ++i) { bad += (int)ho[3*i]; badp += ho[3*i+1] != ho[3*i+2]; }
    printf("R mismatches %d, p_world mismatches %d\n", bad, badp);
    return 0;
}
