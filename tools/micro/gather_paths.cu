// Micro-benchmark: cost of a warp-wide random 32-byte gather (one double4 per lane from a window of ~1500 sorted
// entries, the access pattern of pair_kernel's phase 2) through the LSU path (ld.global.nc 256-bit), through the
// texture path (two 128-bit point fetches) and with the two paths interleaved.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_paths gather_paths.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ double4 ldg256(const double4 *p) {
    double4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double4 tex256(cudaTextureObject_t t, int j) {
    const int4 a = tex1Dfetch<int4>(t, 2 * j), b = tex1Dfetch<int4>(t, 2 * j + 1);
    return make_double4(__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z));
}
__device__ __forceinline__ unsigned hash(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// window sweep (how many distinct 128 B lines a warp-wide gather touches) and a shared-memory source
template <int SRC> // 0 = global 256-bit, 1 = global 2 x 128-bit, 2 = shared 2 x LDS.128 (window <= 1536 entries)
__global__ void __launch_bounds__(256, 2) window_kernel(const double4 *pos, int n, int iters, unsigned window, double *out) {
    extern __shared__ double4 sm[];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int base = (int)(((long long)warp * 1237) % (n - 2048));
    if (SRC == 2) {
        for (int k = threadIdx.x; k < 1536; k += 256) sm[k] = pos[base + k];
        __syncthreads();
    }
    double ax = 0, ay = 0, az = 0, aw = 0;
    unsigned s = hash(warp * 32 + lane + 1);
#pragma unroll 4
    for (int t = 0; t < iters; t++) {
        s = s * 1664525u + 1013904223u;
        const int j = (int)((s >> 8) % window);
        double4 v;
        if (SRC == 0) v = ldg256(pos + base + j);
        else if (SRC == 1) {
            const double2 a = __ldg(reinterpret_cast<const double2 *>(pos + base + j)), b = __ldg(reinterpret_cast<const double2 *>(pos + base + j) + 1);
            v = make_double4(a.x, a.y, b.x, b.y);
        } else v = sm[j];
        ax += v.x; ay += v.y; az += v.z; aw += v.w;
    }
    if (ax + ay + az + aw == 123.456) out[0] = ax;
}

template <int MODE> // 0 = LSU, 1 = TEX, 2 = alternate
__global__ void __launch_bounds__(256, 2) gather_kernel(const double4 *pos, cudaTextureObject_t tex, int n, int iters, double *out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int base = (int)(((long long)warp * 1237) % (n - 2048));
    double ax = 0, ay = 0, az = 0, aw = 0;
    unsigned s = hash(warp * 32 + lane + 1);
#pragma unroll 4
    for (int t = 0; t < iters; t++) {
        s = s * 1664525u + 1013904223u;
        const int j = base + (int)((s >> 8) % 1536u);
        double4 v;
        if (MODE == 0) v = ldg256(pos + j);
        else if (MODE == 1) v = tex256(tex, j);
        else v = (t & 1) ? tex256(tex, j) : ldg256(pos + j);
        ax += v.x; ay += v.y; az += v.z; aw += v.w;
    }
    if (ax + ay + az + aw == 123.456) out[0] = ax;
}

int main() {
    const int n = 1300000;
    double4 *pos;
    cudaMalloc(&pos, sizeof(double4) * n);
    cudaMemset(pos, 0, sizeof(double4) * n);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = pos;
    rd.res.linear.desc = cudaCreateChannelDesc<int4>();
    rd.res.linear.sizeInBytes = sizeof(double4) * n;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { printf("texture object failed\n"); return 1; }
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 2048, blocks = 148 * 2 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) gather_kernel<0><<<blocks, 256>>>(pos, tex, n, iters, out);
            if (mode == 1) gather_kernel<1><<<blocks, 256>>>(pos, tex, n, iters, out);
            if (mode == 2) gather_kernel<2><<<blocks, 256>>>(pos, tex, n, iters, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double gathers = (double)blocks * 8 * iters; // warp-wide gathers
        const double cyc = best * 1e-3 * 1.965e9 * 148 / gathers;
        printf("mode %d (%s): %.3f ms, %.1f SM-cycles per warp-wide 32 B gather, %.0f GB/s useful\n", mode,
               mode == 0 ? "LSU ld.global.nc.v4.f64" : mode == 1 ? "TEX 2 x tex1Dfetch<int4>" : "alternating LSU / TEX", best, cyc,
               gathers * 1024 / (best * 1e-3) / 1e9);
    }
    cudaFuncSetAttribute(window_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1536 * 32);
    const unsigned windows[] = {8, 32, 64, 128, 256, 512, 1536};
    for (int src = 0; src < 3; src++)
        for (unsigned w : windows) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (src == 0) window_kernel<0><<<blocks, 256>>>(pos, n, iters, w, out);
                if (src == 1) window_kernel<1><<<blocks, 256>>>(pos, n, iters, w, out);
                if (src == 2) window_kernel<2><<<blocks, 256, 1536 * 32>>>(pos, n, iters, w, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            const double gathers = (double)blocks * 8 * iters;
            printf("src %d (%s) window %4u entries: %.1f SM-cycles per warp-wide gather\n", src,
                   src == 0 ? "global 256-bit" : src == 1 ? "global 2x128-bit" : "shared 2xLDS.128", w, best * 1e-3 * 1.965e9 * 148 / gathers);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
