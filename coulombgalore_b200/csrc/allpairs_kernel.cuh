// allpairs_kernel.cuh -- the all-pairs tile kernel for small systems (a few thousand particles: BASELINE config 1).
//
// A cell list has nothing to cull when the whole system lies within a cut-off or two of every particle, and below ~10^4
// particles the sort phase costs as much as the pair arithmetic.  This kernel therefore reads the caller's SoA arrays as they
// are -- no binning, no sorted copy, no group table:
//   * work item = (block of i-groups, j-slice).  A warp keeps 32 consecutive i-particles in registers, one per lane, for its
//     whole life; the CTA's warps share one j-slice, and the slices of an i-group are summed afterwards in slice order by
//     the epilogue kernel (deterministic, like the j-split of the cell-list kernel);
//   * the j-slice streams through shared memory in tiles of up to kApTile particles, TWO STAGES: one elected thread arms an
//     mbarrier with the byte count and issues one TMA bulk copy (cp.async.bulk.shared.global, completion on the mbarrier) per
//     SoA array of the tile after next while the warps work on the current one.  The copies need 16-byte granules; an odd
//     last element is stored by the issuing thread before it arrives on the barrier;
//   * PHASE 1 (fp32, conservative): the tile is converted once per CTA into {x', y', z', |c'|^2} relative to particle 0, and
//     every lane tests its particle against the 32 candidates of a batch by broadcast reads:  |c'|^2 - 2 x'_i.c' < thr_i
//     (three FFMA and a compare per candidate; the margin in thr_i covers the fp32 roundings, see prefilter_margin).  The hit
//     masks go to a per-lane list in shared memory, closed by a sentinel that points at far-away dummy entries;
//   * PHASE 2 (FP64 pipe): exactly the deferred evaluation of pair_kernel.cuh -- every lane walks its own hits, all lanes for
//     as many steps as the busiest one needs -- with the gathers served by the tile in shared memory (four 8-byte LDS per
//     pair) instead of L2.  The exact reference gate r2 < cutoff^2 is applied in fp64.
// Open boundaries only (the engine routes periodic systems to the cell-list kernel), fp64 arithmetic, no quadrupoles.
#pragma once
#include <cstdio>

#include "pair_kernel.cuh"

namespace cgb {

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "AP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra AP_DONE;\n"
        "bra AP_WAIT;\n"
        "AP_DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared (1-D, 16-byte granules), completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// launch shape of the all-pairs kernel (tuning builds: -DCG_AP_WARPS_ION=.. etc.); defaults = the cell-list kernel's
#ifndef CG_AP_WARPS_ION
#define CG_AP_WARPS_ION CG_WARPS_ION
#endif
#ifndef CG_AP_MINB_ION
#define CG_AP_MINB_ION CG_MINB_ION
#endif
#ifndef CG_AP_WARPS_DIP
#define CG_AP_WARPS_DIP CG_WARPS_DIP
#endif
#ifndef CG_AP_MINB_DIP
#define CG_AP_MINB_DIP CG_MINB_DIP
#endif
#ifndef CG_AP_SETS
#define CG_AP_SETS 2
#endif
#ifndef CG_AP_ILP_ION
#define CG_AP_ILP_ION CG_ILP
#endif
template <int MODE> constexpr int ap_warps() { return MODE == cgd::kModeIon ? CG_AP_WARPS_ION : CG_AP_WARPS_DIP; }
template <int MODE> constexpr int ap_minb() { return MODE == cgd::kModeIon ? CG_AP_MINB_ION : CG_AP_MINB_DIP; }
template <int MODE> constexpr int ap_ilp() { return MODE == cgd::kModeIon ? CG_AP_ILP_ION : pairs_in_flight<MODE>(); }

constexpr int kApStride = kApTile + kDummies; // doubles per SoA component and stage: the tile, then the dummy entries
template <int MODE> constexpr int ap_arrays() { return MODE == cgd::kModeIon ? 4 : 7; }

template <class F, bool YUK, int MODE, int WHAT>
__global__ void __launch_bounds__(32 * ap_warps<MODE>(), ap_minb<MODE>())
allpairs_kernel(const __grid_constant__ APParams p, const __grid_constant__ F f, const double *__restrict__ dev_table, int table_doubles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kWarps = ap_warps<MODE>();
    constexpr int kArr = ap_arrays<MODE>();
    constexpr int kIlp = ap_ilp<MODE>();
    constexpr int kSets = CG_AP_SETS;
    constexpr int kNS = F::kSchemes;
    typedef double T;
    typedef cgd::Vec3T<T> V3;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned full = 0xffffffffu;

    const size_t o_stage = ((size_t)table_doubles * 8 + 127) & ~(size_t)127; // = ap_layout(table_doubles, kArr, kWarps)
    const size_t o_f32 = o_stage + (size_t)2 * kArr * kApStride * sizeof(double);
    const size_t o_lists = o_f32 + (size_t)2 * kApTile * sizeof(float4);
    const size_t o_bars = o_lists + (size_t)kWarps * (kApTile / 32 + 1) * 32 * sizeof(uint2);
    double *s_table = reinterpret_cast<double *>(smem_raw);
    double *s_stage = reinterpret_cast<double *>(smem_raw + o_stage);
    float4 *s_f32 = reinterpret_cast<float4 *>(smem_raw + o_f32);
    uint2 *hits = reinterpret_cast<uint2 *>(smem_raw + o_lists) + (size_t)warp * (kApTile / 32 + 1) * 32 + lane;
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem_raw + o_bars);
    int *s_wmax = reinterpret_cast<int *>(s_bar + 2);

#ifdef CG_AP_PROF
    long long apt[8];
    int apk = 0;
#define AP_MARK() apt[apk++] = clock64()
#else
#define AP_MARK()
#endif
    AP_MARK();
    // slice s of the j-range = the batches of 32 consecutive particles s, s + S, s + 2 S, ...: every slice samples the whole
    // system, so every (i-block, slice) item sees the same density of neighbours (contiguous slices of a spatially ordered
    // input would give the items next to the diagonal several times the work of the others)
    const int slice = blockIdx.y, S = p.n_slices;
    const int nbatch = (p.n + 31) >> 5;                               // batches of the system
    const int nbs = slice < nbatch ? (nbatch - 1 - slice) / S + 1 : 0; // batches of this slice
    constexpr int kTileBatches = kApTile / 32;
    const int ntiles = (nbs + kTileBatches - 1) / kTileBatches;
    const double *src[7] = {p.x, p.y, p.z, p.q, p.mx, p.my, p.mz};
    const unsigned present = (p.x ? 1u : 0u) | (p.y ? 2u : 0u) | (p.z ? 4u : 0u) | (p.q ? 8u : 0u) | (p.mx ? 16u : 0u) | (p.my ? 32u : 0u) |
                             (p.mz ? 64u : 0u);

    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) { // what the epilogue kernel reads: i-groups, no overflow
        p.counts[0] = p.n;
        p.counts[1] = (p.n + 31) >> 5;
        p.counts[2] = 0;
        p.counts[3] = 0;
    }
    const double *tab = nullptr;
    if constexpr (F::kTable != cgd::kTableNone) {
        for (int k = tid; k < table_doubles; k += 32 * kWarps) s_table[k] = dev_table[k];
        tab = s_table;
    }
    // dummy entries behind both stages (far away: they fail the gate), zeros for an absent array
    for (int k = tid; k < 2 * kArr * kDummies; k += 32 * kWarps) {
        const int sc = k / kDummies, c = sc % kArr, e = k - sc * kDummies;
        s_stage[sc * kApStride + kApTile + e] = c < 3 ? 1e30 : 0.0;
    }
    if ((present & ((1u << kArr) - 1u)) != (1u << kArr) - 1u) {
        for (int k = tid; k < 2 * kArr * kApTile; k += 32 * kWarps) {
            const int sc = k / kApTile, c = sc % kArr, e = k - sc * kApTile;
            if (!((present >> c) & 1u)) s_stage[sc * kApStride + e] = 0.0;
        }
    }
    if (tid == 0) {
        mbar_init(smem_addr(s_bar), 1);
        mbar_init(smem_addr(s_bar + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // issue tile t into stage t & 1 (warp 0): one bulk copy per batch and SoA array (256 bytes; the system's last batch may be
    // shorter, and an odd last element is stored by hand: the copies move 16-byte granules).  Lane 0 arms the barrier with the
    // byte count of the whole tile, then the lanes share the copies.
    auto issue = [&](int t) {
        const int s = t & 1, k0 = t * kTileBatches, nbt = min(kTileBatches, nbs - k0);
        double *st = s_stage + (size_t)s * kArr * kApStride;
        const unsigned bar = smem_addr(s_bar + s);
        const int narr = __popc(present & ((1u << kArr) - 1u));
        if (lane == 0) {
            s_wmax[s] = 0;
            const int last = slice + (k0 + nbt - 1) * S;                     // the tile's last batch
            const int cnt_last = min(32, p.n - 32 * last), ce_last = cnt_last & ~1;
            if (cnt_last & 1) {
#pragma unroll
                for (int c = 0; c < kArr; c++)
                    if ((present >> c) & 1u) st[c * kApStride + 32 * (nbt - 1) + cnt_last - 1] = src[c][32 * last + cnt_last - 1];
            }
            mbar_arrive_expect_tx(bar, (unsigned)((32 * (nbt - 1) + ce_last) * 8 * narr));
        }
        __syncwarp();
        for (int w = lane; w < nbt * kArr; w += 32) {
            const int k = w / kArr, c = w - k * kArr;
            if (!((present >> c) & 1u)) continue;
            const int b = slice + (k0 + k) * S;
            const int ce = min(32, p.n - 32 * b) & ~1;
            const double *g = c == 0 ? p.x : c == 1 ? p.y : c == 2 ? p.z : c == 3 ? p.q : c == 4 ? p.mx : c == 5 ? p.my : p.mz;
            if (ce > 0) tma_bulk_g2s(smem_addr(st + c * kApStride + 32 * k), g + 32 * b, (unsigned)(ce * 8), bar);
        }
    };
    if (warp == 0 && ntiles > 0) {
        issue(0);
        if (ntiles > 1) issue(1);
    }

    // the lane's own particle
    // i-groups are dealt to the CTAs round robin (warp w of i-block b: group b + w * i-blocks): an open system has half the
    // neighbours at its faces, and every SM should get the same mix of light and heavy groups
    const int ig = blockIdx.x + warp * gridDim.x;
    const int i = ig * 32 + lane;
    const bool act = i < p.n;
    const int il = act ? i : 0;
    const double xi = p.x[il], yi = p.y[il], zi_ = p.z[il];
    const T zi = p.q ? p.q[il] : 0.0;
    V3 mui = {T(0), T(0), T(0)};
    if constexpr (MODE != cgd::kModeIon) mui = {p.mx[il], p.my[il], p.mz[il]};
    // fp32 frame of the prefilter: coordinates relative to particle 0
    const double Ox = p.x[0], Oy = p.y[0], Oz = p.z[0];
    const float xr = (float)(xi - Ox), yr = (float)(yi - Oy), zr = (float)(zi_ - Oz);
    const float ax = -2.0f * xr, ay = -2.0f * yr, az = -2.0f * zr;
    const double xn2 = (double)xr * xr + (double)yr * yr + (double)zr * zr;
    const float xnorm = sqrtf((float)xn2) * 1.000001f;
    const float thr0 = (float)(p.rc2 < 1e300 ? p.rc2 - xn2 : 1e300); // +inf for an unbounded cut-off

    cgd::Accum acc[kNS];
    unsigned int npairs = 0;

    for (int t = 0; t < ntiles; t++) {
        const int s = t & 1, k0 = t * kTileBatches, nbt = min(kTileBatches, nbs - k0);
        const double *st = s_stage + (size_t)s * kArr * kApStride;
        float4 *c32 = s_f32 + s * kApTile;
        if (t == 0) AP_MARK();
        mbar_wait(smem_addr(s_bar + s), (unsigned)((t >> 1) & 1));
        if (t == 0) AP_MARK();
        // ---- fp32 copy of the tile for the prefilter, and the largest |c'|^2 of the tile (for the margin) ----
        {
            float wm = 0.0f;
            for (int k = tid; k < 32 * nbt; k += 32 * kWarps) {
                float4 v = make_float4(0.0f, 0.0f, 0.0f, INFINITY);
                if (32 * (slice + (k0 + (k >> 5)) * S) + (k & 31) < p.n) { // only the system's last batch can be short
                    v.x = (float)(st[k] - Ox);
                    v.y = (float)(st[kApStride + k] - Oy);
                    v.z = (float)(st[2 * kApStride + k] - Oz);
                    v.w = v.x * v.x + v.y * v.y + v.z * v.z;
                    wm = fmaxf(wm, v.w);
                }
                c32[k] = v;
            }
            wm = __int_as_float(__reduce_max_sync(full, __float_as_int(wm))); // >= 0: float order == int order
            if (lane == 0) atomicMax(s_wmax + s, __float_as_int(wm));
        }
        __syncthreads();
        if (t == 0) AP_MARK();
        // ---- PHASE 1 ----
        // Error budget of the fp32 test against the exact r2 < rc2, in units of 2^-24 (S = |x'_i| + max |c'| of the tile): both
        // positions rounded to fp32: 2 S^2; |c'|^2 in three roundings: 3 S^2; three FFMA partial sums <= S^2 each: 3 S^2; the
        // rounding of thr0 and of thr0 + margin: 2 (rc2 + S^2).  Needed: 10 S^2 + 2 rc2; the margin takes 1.5 x that.
        float thr = -INFINITY;
        if (act) {
            const float S = (xnorm + sqrtf(__int_as_float(s_wmax[s]))) * 1.0001f;
            thr = thr0 + 1.5f * 5.9604645e-8f * (10.0f * S * S + 2.0f * (fabsf(thr0) + (float)xn2)); // |thr0| + |x'|^2 >= rc2
        }
        int qn = 0, total = 0;
        for (int b = 0; b < nbt; b++) {
            const float4 *c4 = c32 + 32 * b;
            unsigned m = 0u;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const float4 v = c4[c];
                float d = fmaf(ax, v.x, v.w);
                d = fmaf(ay, v.y, d);
                d = fmaf(az, v.z, d);
                m |= d < thr ? (1u << c) : 0u;
            }
            const unsigned rel_self = (unsigned)(i - 32 * (slice + (k0 + b) * S)); // the lane's own particle is not a neighbour
            if (rel_self < 32u) m &= ~(1u << rel_self);
            if (m != 0u) hits[(qn++) * 32] = make_uint2(m, (unsigned)(32 * b));
            total += __popc(m);
        }
        hits[qn * 32] = make_uint2(0xffffffffu, (unsigned)kApTile); // sentinel: the dummy entries behind the tile
        const int rounds = (__reduce_max_sync(full, total) + kSets * kIlp - 1) / (kSets * kIlp);
        __syncwarp();
        if (t == 0) AP_MARK();
        // ---- PHASE 2 ----
        if (rounds > 0) {
            const unsigned list = smem_addr(hits);
            const unsigned sent = list + 256u * (unsigned)qn;
            unsigned ptr = list, cur, base;
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(cur), "=r"(base) : "r"(ptr));
            auto pop = [&]() -> int {
                const int b = top_bit(cur);
                const int j = (int)base + b;
                cur &= ~(1u << b);
                if (cur == 0u) {
                    ptr = min(ptr + 256u, sent);
                    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(cur), "=r"(base) : "r"(ptr));
                }
                return j;
            };
            int jj[kSets][kIlp];
            double px[kSets][kIlp], py[kSets][kIlp], pz[kSets][kIlp], pq[kSets][kIlp];
            double ma[kSets][kIlp], mb[kSets][kIlp], mc[kSets][kIlp];
            auto fetch = [&](int k) {
#pragma unroll
                for (int u = 0; u < kIlp; u++) {
                    const int j = pop();
                    jj[k][u] = j;
                    px[k][u] = st[j];
                    py[k][u] = st[kApStride + j];
                    pz[k][u] = st[2 * kApStride + j];
                    pq[k][u] = st[3 * kApStride + j];
                    if constexpr (MODE != cgd::kModeIon) {
                        ma[k][u] = st[4 * kApStride + j];
                        mb[k][u] = st[5 * kApStride + j];
                        mc[k][u] = st[6 * kApStride + j];
                    }
                }
            };
            auto compute = [&](int k) {
#pragma unroll
                for (int u = 0; u < kIlp; u++) {
                    const double dx = xi - px[k][u], dy = yi - py[k][u], dz = zi_ - pz[k][u];
                    const double r2 = dx * dx + dy * dy + dz * dz;
                    const bool pass = r2 < p.rc2 && jj[k][u] < kApTile; // the exact, strict reference gate; dummies never pass
                    if (pass) npairs++;
                    const double r2c = __hiloint2double(min(__double2hiint(r2), p.r2cap_hi), __double2loint(r2));
                    const V3 d = {dx, dy, dz};
                    V3 muj = {T(0), T(0), T(0)};
                    if constexpr (MODE != cgd::kModeIon) muj = {ma[k][u], mb[k][u], mc[k][u]};
                    if constexpr (kNS > 1)
                        cgd::pair_accumulate_multi<F, WHAT, T>(f, p.core, tab, d, r2c, zi, pq[k][u], acc, pass);
                    else
                        cgd::pair_accumulate<F, YUK, MODE, WHAT, T>(f, p.core, tab, d, r2c, zi, pq[k][u], mui, muj, acc[0], pass);
                }
            };
#pragma unroll
            for (int k = 0; k < kSets - 1; k++) fetch(k);
#pragma unroll 1
            for (int r = 0; r < rounds; r++) {
#pragma unroll
                for (int k = 0; k < kSets; k++) {
                    fetch((k + kSets - 1) % kSets);
                    compute(k);
                }
            }
        }
        if (t == 0) AP_MARK();
        __syncthreads(); // every warp is done with stage s: refill it
        if (t == 0) AP_MARK();
        if (warp == 0 && t + 2 < ntiles) issue(t + 2);
    }

    // ---- per-particle partial outputs of this slice, one reduction per (i-group, slice) for the scalars ----
    if (ig >= ((p.n + 31) >> 5)) return;
    const int item = ig * p.n_slices + slice;
    unsigned int np = npairs;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) np += __shfl_down_sync(full, np, off);
    if (lane == 0) p.item_pairs[item] = np;
#pragma unroll
    for (int k = 0; k < kNS; k++) {
        double *o_force = k == 0 ? p.force : p.force_k[k > 0 ? k - 1 : 0], *o_field = k == 0 ? p.field : p.field_k[k > 0 ? k - 1 : 0];
        double *o_pot = k == 0 ? p.potential : p.potential_k[k > 0 ? k - 1 : 0];
        double *o_energy = k == 0 ? p.item_energy : p.item_energy_k[k > 0 ? k - 1 : 0];
        double u = 0.0;
        if (act) {
            const size_t o = (size_t)slice * (size_t)p.n + (size_t)i;
            if constexpr ((WHAT & cgd::kWhatForce) != 0) {
                if (o_force) {
                    o_force[3 * o + 0] = acc[k].Fx;
                    o_force[3 * o + 1] = acc[k].Fy;
                    o_force[3 * o + 2] = acc[k].Fz;
                }
            }
            if constexpr ((WHAT & cgd::kWhatField) != 0) {
                if (o_field) {
                    o_field[3 * o + 0] = acc[k].Ex;
                    o_field[3 * o + 1] = acc[k].Ey;
                    o_field[3 * o + 2] = acc[k].Ez;
                }
            }
            if constexpr ((WHAT & cgd::kWhatPotential) != 0) {
                if (o_pot) o_pot[o] = acc[k].phi;
            }
            if ((WHAT & cgd::kWhatEnergy) != 0) u = acc[k].U;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) u += __shfl_down_sync(full, u, off);
        if (lane == 0) o_energy[item] = u;
    }
#ifdef CG_AP_PROF
    AP_MARK();
    if (lane == 0 && (warp == 0 || warp == 5) && (blockIdx.x == 0 || blockIdx.x == 5) && blockIdx.y < 2)
        printf("ap prof cta (%d,%d) warp %d: init %lld wait %lld convert %lld phase1 %lld phase2 %lld sync %lld out %lld total %lld cycles\n",
               blockIdx.x, blockIdx.y, warp, apt[1] - apt[0], apt[2] - apt[1], apt[3] - apt[2], apt[4] - apt[3], apt[5] - apt[4],
               apt[6] - apt[5], apt[7] - apt[6], apt[7] - apt[0]);
#endif
}

template <class F, bool YUK, int MODE, int WHAT>
cudaError_t launch_allpairs(const APParams &p, const cg_scheme_desc &desc, const double *dev_table, int table_doubles, int sms,
                            cudaStream_t stream) {
    const F f(desc);
    if (F::kTable == cgd::kTableNone) table_doubles = 0;
    constexpr int warps = ap_warps<MODE>();
    const ApLayout L = ap_layout(table_doubles, ap_arrays<MODE>(), warps);
    auto kern = allpairs_kernel<F, YUK, MODE, WHAT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return e;
    (void)sms;
    const int n_groups = (p.n + 31) / 32;
    const dim3 grid((unsigned)((n_groups + warps - 1) / warps), (unsigned)p.n_slices);
    kern<<<grid, 32 * warps, L.total, stream>>>(p, f, dev_table, table_doubles);
    return cudaGetLastError();
}

template <int MODE> void ap_dims_of(int *warps, int *ctas_per_sm) {
    *warps = ap_warps<MODE>();
    *ctas_per_sm = ap_minb<MODE>();
}
inline void ap_dims(int ci, int *warps, int *ctas_per_sm) {
    const int mode = combo(ci).mode;
    if (mode == cgd::kModeIon) ap_dims_of<cgd::kModeIon>(warps, ctas_per_sm);
    else if (mode == cgd::kModeDipole) ap_dims_of<cgd::kModeDipole>(warps, ctas_per_sm);
    else ap_dims_of<cgd::kModeMultipole>(warps, ctas_per_sm);
}

#define CGB_AP_ROW(FUNCTOR, YUK)                                                                                              \
    {launch_allpairs<FUNCTOR, YUK, combo(0).mode, combo(0).what>, launch_allpairs<FUNCTOR, YUK, combo(1).mode, combo(1).what>, \
     launch_allpairs<FUNCTOR, YUK, combo(2).mode, combo(2).what>, launch_allpairs<FUNCTOR, YUK, combo(3).mode, combo(3).what>, \
     launch_allpairs<FUNCTOR, YUK, combo(4).mode, combo(4).what>, launch_allpairs<FUNCTOR, YUK, combo(5).mode, combo(5).what>, \
     launch_allpairs<FUNCTOR, YUK, combo(6).mode, combo(6).what>, launch_allpairs<FUNCTOR, YUK, combo(7).mode, combo(7).what>, \
     nullptr, nullptr}
#define CGB_AP_ROW_ION(FUNCTOR)                                                                                          \
    {launch_allpairs<FUNCTOR, false, combo(0).mode, combo(0).what>, launch_allpairs<FUNCTOR, false, combo(1).mode, combo(1).what>, \
     launch_allpairs<FUNCTOR, false, combo(2).mode, combo(2).what>, nullptr, nullptr, nullptr, nullptr,                  \
     launch_allpairs<FUNCTOR, false, combo(7).mode, combo(7).what>, nullptr, nullptr}

} // namespace cgb
