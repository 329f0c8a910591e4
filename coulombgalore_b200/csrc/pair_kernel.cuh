// pair_kernel.cuh -- the sm_100a pair-interaction kernel (FP64-pipe bound; no tensor cores: the work is
// elementwise transcendental/polynomial math per pair, not a contraction).
//
// Work decomposition
//   * one warp per work item = (i-group, j-split slice); an i-group is <= 32 spatially adjacent particles of one
//     cell row, one particle per lane, held in registers for the whole item (positions, charge, dipole, accumulators);
//   * the candidate j's are the particles of every cell that intersects the group's bounding box grown by the
//     cut-off; cells are stored x-fastest so each (y,z) cell row contributes ONE contiguous run of sorted entries.
//     The run table is built 32 rows at a time, one row per lane, with rounded-corner culling in the y-z plane;
//   * candidates stream through shared memory 32 at a time (one coalesced 16 B load per lane, then 32 broadcast
//     LDS.128) and are tested against the lane's own particle in FP32 on the FMA pipe -- a conservative prefilter
//     (cut-off grown by the fp32 rounding margin), so the FP64 pipe never sees the ~80 % of candidates that are out
//     of range;
//   * a candidate that passes is pushed on the lane's private queue in shared memory (deferred evaluation).  The
//     warp evaluates queued pairs only when every active lane has one (or a queue is about to overflow), so the
//     expensive FP64 pair arithmetic runs with all lanes busy instead of at the ~15-25 % in-cut-off hit rate a
//     branch-per-candidate loop would give;
//   * the FP64 evaluation re-tests the exact reference gate r2 < cutoff^2 on the FP64 coordinates, gathers j through
//     L1/L2 (sorted order keeps the neighbourhood resident) and accumulates into the lane's registers: no atomics,
//     no cross-lane reduction for per-particle outputs; energies/pair counts are reduced once per item.
#pragma once
#include "kparams.cuh"

namespace cgb {

template <class F> inline F make_functor(const cg_scheme_desc &d, const double *dev_table) {
    F f(d);
    if constexpr (F::kHasTable) f.tab = dev_table;
    return f;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <class F, bool YUK, int MODE, int WHAT>
__global__ void __launch_bounds__(kThreads, 2)
pair_kernel(const __grid_constant__ KParams p, const __grid_constant__ F f, int table_doubles, int items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_table = reinterpret_cast<double *>(smem_raw);
    const int table_bytes = (table_doubles * 8 + 15) & ~15;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *stage = reinterpret_cast<float4 *>(smem_raw + table_bytes) + warp * 32;
    unsigned int *queue = reinterpret_cast<unsigned int *>(smem_raw + table_bytes + kWarpsPerCta * 32 * sizeof(float4)) +
                          warp * (kQueueCap * 32);
    const double *tab = nullptr;
    if constexpr (F::kHasTable) { // Splined: stage the four Andrea tables (<= a few KB) in shared memory
        for (int k = threadIdx.x; k < table_doubles; k += kThreads) s_table[k] = f.tab[k];
        __syncthreads();
        tab = s_table;
    }
    const int item = blockIdx.x * kWarpsPerCta + warp;
    if (item >= items) return;
    const unsigned full = 0xffffffffu;
    const int split = p.split;
    const int g = p.group_begin + item / split, slice = item - (item / split) * split;
    const int2 grp = p.groups[g];
    const bool active = lane < grp.y;
    const int i = grp.x + (active ? lane : 0);

    // the lane's own particle
    const double4 pi = p.pos[i];
    cgd::D3 mui = {0.0, 0.0, 0.0};
    if constexpr (MODE != cgd::kModeIon) {
        const double4 m = p.mu[i];
        mui = {m.x, m.y, m.z};
    }
    const float4 pf = p.pos32[i];
    // bounding box of the group (coordinates relative to the grid origin are >= 0: float order == int order)
    const float xmin = __int_as_float(__reduce_min_sync(full, __float_as_int(pf.x)));
    const float xmax = __int_as_float(__reduce_max_sync(full, __float_as_int(pf.x)));
    const float ymin = __int_as_float(__reduce_min_sync(full, __float_as_int(pf.y)));
    const float ymax = __int_as_float(__reduce_max_sync(full, __float_as_int(pf.y)));
    const float zmin = __int_as_float(__reduce_min_sync(full, __float_as_int(pf.z)));
    const float zmax = __int_as_float(__reduce_max_sync(full, __float_as_int(pf.z)));
    // inactive lanes sit infinitely far away: they never pass the prefilter
    const float xi = active ? pf.x : 1e30f, yi = pf.y, zi = pf.z;

    const float rcm = p.rc_m, rc2m = p.rc2_m;
    const int cy0 = clampi((int)floorf((ymin - rcm) * p.inv_cs_y - p.slack), 0, p.ny - 1);
    const int cy1 = clampi((int)floorf((ymax + rcm) * p.inv_cs_y + p.slack), 0, p.ny - 1);
    const int cz0 = clampi((int)floorf((zmin - rcm) * p.inv_cs_z - p.slack), 0, p.nz - 1);
    const int cz1 = clampi((int)floorf((zmax + rcm) * p.inv_cs_z + p.slack), 0, p.nz - 1);
    const int nyr = cy1 - cy0 + 1;
    const int nrows = nyr * (cz1 - cz0 + 1);

    cgd::Accum acc;
    unsigned int npairs = 0;
    int qlen = 0;

    // iterator over the candidate runs (all warp-uniform except jb/je, which hold one row's run per lane)
    int row_next = 0;       // next chunk of 32 rows to expand
    unsigned have = 0;      // lanes whose run of the current chunk is still unvisited
    int jb = 0, je = 0;     // this lane's run in the current chunk
    int cur_j = 0, cur_end = 0;

    for (;;) {
        // ---- advance: fetch and prefilter one batch of <= 32 candidates, if any are left ----
        bool more = true;
        while (cur_j >= cur_end) {
            if (have == 0) {
                if (row_next >= nrows) {
                    more = false;
                    break;
                }
                const int rr = row_next + lane;
                row_next += 32;
                jb = je = 0;
                if (rr < nrows) {
                    const int rzi = rr / nyr;
                    const int ry = cy0 + (rr - rzi * nyr), rz = cz0 + rzi;
                    const float ylo = (float)ry * p.cs_y, zlo = (float)rz * p.cs_z;
                    const float dy = fmaxf(0.0f, fmaxf(ylo - ymax, ymin - (ylo + p.cs_y)));
                    const float dz = fmaxf(0.0f, fmaxf(zlo - zmax, zmin - (zlo + p.cs_z)));
                    const float w2 = rcm * rcm - (dy * dy + dz * dz);
                    if (w2 > 0.0f) { // rows farther than the cut-off from the box in the y-z plane hold no neighbour
                        const float w = sqrtf(w2);
                        const int cxa = clampi((int)floorf((xmin - w) * p.inv_cs_x - p.slack), 0, p.nx - 1);
                        const int cxb = clampi((int)floorf((xmax + w) * p.inv_cs_x + p.slack), 0, p.nx - 1);
                        const int rowbase = (rz * p.ny + ry) * p.nx;
                        jb = p.cell_start[rowbase + cxa];
                        je = p.cell_start[rowbase + cxb + 1];
                        if (split > 1) { // j-split: this slice's share of the run
                            const long long len = je - jb;
                            const int a = jb + (int)(len * slice / split), b = jb + (int)(len * (slice + 1) / split);
                            jb = a;
                            je = b;
                        }
                    }
                }
                have = __ballot_sync(full, je > jb);
                continue;
            }
            const int l = __ffs(have) - 1;
            have &= have - 1;
            cur_j = __shfl_sync(full, jb, l);
            cur_end = __shfl_sync(full, je, l);
        }
        if (more) {
            const int j0 = cur_j;
            const int n = min(32, cur_end - j0);
            cur_j += 32;
            float4 c = make_float4(3e30f, 0.0f, 0.0f, 0.0f); // sentinel: far away
            if (lane < n) c = p.pos32[j0 + lane];
            __syncwarp();
            stage[lane] = c;
            __syncwarp();
            const int rel_self = i - j0; // candidate slot that is the lane's own particle (if in range)
            const int n4 = (n + 3) & ~3;
#pragma unroll 1
            for (int t = 0; t < n4; t += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float4 cj = stage[t + u];
                    const float dx = cj.x - xi, dy = cj.y - yi, dz = cj.z - zi;
                    const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    if (r2 < rc2m && (t + u) != rel_self) {
                        queue[qlen * 32 + lane] = (unsigned int)(j0 + t + u);
                        qlen++;
                    }
                }
            }
        }
        // ---- evaluate queued pairs while the warp can do so with every active lane busy (or must) ----
        for (;;) {
            bool go;
            if (more) go = __any_sync(full, qlen > kQueueCap - 33) || __all_sync(full, qlen > 0 || !active);
            else go = __any_sync(full, qlen > 0);
            if (!go) break;
            if (qlen > 0) {
                qlen--;
                const int j = (int)queue[qlen * 32 + lane];
                const double4 pj = p.pos[j];
                const cgd::D3 d = {pi.x - pj.x, pi.y - pj.y, pi.z - pj.z};
                const double r2 = d.x * d.x + d.y * d.y + d.z * d.z;
                if (r2 < p.rc2) { // the exact, strict reference gate
                    npairs++;
                    cgd::D3 muj = {0.0, 0.0, 0.0};
                    if constexpr (MODE != cgd::kModeIon) {
                        const double4 m = p.mu[j];
                        muj = {m.x, m.y, m.z};
                    }
                    cgd::pair_accumulate<F, YUK, MODE, WHAT>(f, p.core, tab, d, r2, pi.w, pj.w, mui, muj, acc);
                }
            }
        }
        if (!more) break;
    }

    // ---- epilogue: per-particle outputs straight from registers, one reduction per item for the scalars ----
    const size_t part = (size_t)slice * (size_t)p.n;
    if (active) {
        const size_t o = part + (size_t)p.orig[i];
        if constexpr ((WHAT & cgd::kWhatForce) != 0) {
            if (p.force) {
                p.force[3 * o + 0] = acc.Fx;
                p.force[3 * o + 1] = acc.Fy;
                p.force[3 * o + 2] = acc.Fz;
            }
        }
        if constexpr ((WHAT & cgd::kWhatField) != 0) {
            if (p.field) {
                p.field[3 * o + 0] = acc.Ex;
                p.field[3 * o + 1] = acc.Ey;
                p.field[3 * o + 2] = acc.Ez;
            }
        }
        if constexpr ((WHAT & cgd::kWhatPotential) != 0) {
            if (p.potential) p.potential[o] = acc.phi;
        }
    }
    double u = ((WHAT & cgd::kWhatEnergy) != 0 && active) ? acc.U : 0.0;
    unsigned int np = active ? npairs : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        u += __shfl_down_sync(full, u, off);
        np += __shfl_down_sync(full, np, off);
    }
    if (lane == 0) {
        p.item_energy[item] = u;
        p.item_pairs[item] = np;
    }
}

template <class F, bool YUK, int MODE, int WHAT>
cudaError_t launch_pair(const KParams &p, const cg_scheme_desc &desc, const double *dev_table, int table_doubles,
                        int items, cudaStream_t stream) {
    const F f = make_functor<F>(desc, dev_table);
    const int table_bytes = (table_doubles * 8 + 15) & ~15;
    const size_t smem = (size_t)table_bytes + kWarpsPerCta * (32 * sizeof(float4) + kQueueCap * 32 * sizeof(unsigned int));
    auto kern = pair_kernel<F, YUK, MODE, WHAT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (items + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid <= 0) return cudaSuccess;
    kern<<<grid, kThreads, smem, stream>>>(p, f, table_doubles, items);
    return cudaGetLastError();
}

template <class F> __global__ void srf_kernel(const __grid_constant__ F f, int64_t n, const double *q, double *out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s[4];
    f.template eval<3>(q[k], s, nullptr);
    out[k] = s[0];
    out[n + k] = s[1];
    out[2 * n + k] = s[2];
    out[3 * n + k] = s[3];
}

template <class F>
cudaError_t launch_srf(const cg_scheme_desc &desc, const double *dev_table, int, int64_t n, const double *q, double *out,
                       cudaStream_t stream) {
    const F f = make_functor<F>(desc, dev_table);
    if (n <= 0) return cudaSuccess;
    srf_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(f, n, q, out);
    return cudaGetLastError();
}

// table of launchers for one functor: the eight instantiated (mode, what) combinations
#define CGB_DEFINE_FUNCTOR_ENTRY(NAME, FUNCTOR, YUK)                                                                  \
    namespace cgb {                                                                                                   \
    const FunctorEntry *NAME() {                                                                                      \
        static const FunctorEntry e = {                                                                               \
            {launch_pair<FUNCTOR, YUK, combo(0).mode, combo(0).what>, launch_pair<FUNCTOR, YUK, combo(1).mode, combo(1).what>, \
             launch_pair<FUNCTOR, YUK, combo(2).mode, combo(2).what>, launch_pair<FUNCTOR, YUK, combo(3).mode, combo(3).what>, \
             launch_pair<FUNCTOR, YUK, combo(4).mode, combo(4).what>, launch_pair<FUNCTOR, YUK, combo(5).mode, combo(5).what>, \
             launch_pair<FUNCTOR, YUK, combo(6).mode, combo(6).what>, launch_pair<FUNCTOR, YUK, combo(7).mode, combo(7).what>}, \
            launch_srf<FUNCTOR>};                                                                                     \
        return &e;                                                                                                    \
    }                                                                                                                 \
    }

} // namespace cgb
