// FP64 pipe micro-benchmark: dependent-chain latency and throughput vs ILP and warps per SM sub-partition (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(int iters, double a, double b, double *out, long long *cyc) {
    double v[ILP];
    for (int u = 0; u < ILP; u++) v[u] = threadIdx.x * 1e-9 + u;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < ILP; u++) v[u] = fma(v[u], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int u = 0; u < ILP; u++) s += v[u];
    if (s == 1.2345) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps_per_sm, double *out, long long *cyc) {
    const int iters = 4096;
    k<ILP><<<148, warps_per_sm * 32>>>(iters, 0.999, 1e-9, out, cyc);
    cudaDeviceSynchronize();
    k<ILP><<<148, warps_per_sm * 32>>>(iters, 0.999, 1e-9, out, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / iters;
    printf("warps/SM %2d (%.1f/SMSP) ILP %d: %.2f cycles per loop iteration, %.2f cycles per DFMA per warp, SMSP DFMA rate %.3f /cycle\n", warps_per_sm,
           warps_per_sm / 4.0, ILP, per, per / ILP, (warps_per_sm / 4.0) * ILP / per);
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    for (int w : {4, 8, 16, 32}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
    return 0;
}
