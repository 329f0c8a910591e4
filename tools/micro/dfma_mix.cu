// Does an FP64 instruction block the scheduler's issue port for its second pipe cycle?  Mix DFMA with independent
// FFMA / IMAD / LDS work and measure the achieved per-SMSP rates (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NF, int NI, int NL> __global__ void k(int iters, double a, double b, float fa, float fb, int ia, double *out, long long *cyc) {
    __shared__ float sm[1024];
    sm[threadIdx.x & 1023] = threadIdx.x;
    __syncthreads();
    double v[ND > 0 ? ND : 1];
    float f[NF > 0 ? NF : 1];
    int n[NI > 0 ? NI : 1];
    float l[NL > 0 ? NL : 1];
    for (int u = 0; u < ND; u++) v[u] = threadIdx.x * 1e-9 + u;
    for (int u = 0; u < NF; u++) f[u] = threadIdx.x * 1e-3f + u;
    for (int u = 0; u < NI; u++) n[u] = threadIdx.x + u;
    for (int u = 0; u < NL; u++) l[u] = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (u < ND) v[u] = fma(v[u], a, b);
            if (u < NF) f[u] = fmaf(f[u], fa, fb);
            if (u < NI) n[u] = n[u] * ia + 12345;
            if (u < NL) l[u] += sm[(threadIdx.x + i + u * 32) & 1023];
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int u = 0; u < ND; u++) s += v[u];
    for (int u = 0; u < NF; u++) s += f[u];
    for (int u = 0; u < NI; u++) s += n[u];
    for (int u = 0; u < NL; u++) s += l[u];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ND, int NF, int NI, int NL> void run(double *out, long long *cyc) {
    const int iters = 2048, warps = 32;
    for (int r = 0; r < 2; r++) { k<ND, NF, NI, NL><<<148, warps * 32>>>(iters, 0.999, 1e-9, 0.999f, 1e-3f, 3, out, cyc); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / iters; // cycles per iteration for 8 warps/SMSP
    printf("per iter: %d DFMA + %d FFMA + %d IMAD + %d LDS per warp: %.1f cycles/iter -> per SMSP/cycle: DFMA %.3f FFMA %.3f IMAD %.3f LDS %.3f total %.3f\n", ND, NF, NI, NL, per,
           8.0 * ND / per, 8.0 * NF / per, 8.0 * NI / per, 8.0 * NL / per, 8.0 * (ND + NF + NI + NL) / per);
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    run<4, 0, 0, 0>(out, cyc); run<0, 8, 0, 0>(out, cyc); run<0, 0, 8, 0>(out, cyc);
    run<4, 4, 0, 0>(out, cyc); run<4, 8, 0, 0>(out, cyc); run<4, 0, 4, 0>(out, cyc); run<4, 0, 8, 0>(out, cyc);
    run<4, 4, 4, 0>(out, cyc); run<4, 2, 2, 0>(out, cyc); run<4, 2, 2, 1>(out, cyc); run<4, 0, 0, 2>(out, cyc); run<2, 4, 4, 0>(out, cyc);
    return 0;
}
