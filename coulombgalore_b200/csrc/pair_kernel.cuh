// pair_kernel.cuh -- the sm_100a pair-interaction kernel.  The pair arithmetic is elementwise transcendental / polynomial FP64
// math (FP64 pipe); the one dense product of the path, the 32 x 32 distance test of the prefilter, runs on the tensor cores.
//
// Work decomposition
//   * persistent grid: every warp fetches work items (i-group, j-split slice) from a global counter until none is left; an
//     i-group is <= 32 spatially adjacent particles of one cell row, one particle per lane, held in registers for the whole
//     item (position, charge, dipole, accumulators);
//   * the candidate j's are the particles of every cell that intersects the group's bounding box grown by the cut-off; cells
//     are stored x-fastest so each (y,z) cell row contributes ONE contiguous run of sorted entries.  The run table is built
//     32 rows at a time, one row per lane, with rounded-corner culling in the y-z plane.  Rows are visited in a strided order
//     (row k S mod n, S coprime to n), so that any stretch of the walk samples the whole neighbourhood: the hits of one chunk
//     (below) are then spread evenly over the lanes;
//   * PHASE 1 (prefilter; no FP64 instruction): 32 candidates at a time are tested against the 32 particles of the group,
//           D[i][c] = |c'|^2 - 2 x'_i . c' - thr'_i  <  0         (coordinates relative to the group centre)
//     as a 32 x 32 x 5 product on the tensor cores (eight TF32 mma.sync.m16n8k8; the operand rounding is covered by a per-group
//     margin, so the test is conservative).  The B fragments are loaded so that the packed sign bits come out in candidate
//     order: bit b of a particle's mask <-> candidate j0 + b.  Non-empty (mask, j0) words of up to `slots` batches are parked
//     in a per-lane list in shared memory, closed by a sentinel entry (all bits set, pointing at 32 far-away dummy entries);
//   * PHASE 2 (FP64 pipe): every lane walks the set bits of its own masks.  All lanes run the same number of steps -- what the
//     lane with the most hits needs -- and lanes that run dry pop dummy entries from the sentinel, which fail the exact gate:
//     the expensive FP64 pair arithmetic runs with (nearly) all lanes busy instead of at the ~15-25 % in-cut-off hit rate a
//     branch-per-candidate loop would give, and the walk itself costs nine instructions per pair.  Per pair: one 256-bit gather
//     of j through L1 (two for dipoles), the exact reference gate r2 < cutoff^2 as a select, the scheme's radial terms and the
//     accumulation into the lane's registers.  kIlp pairs per lane are in flight and the next step's gathers are issued before
//     the current step's arithmetic.  No atomics, no cross-lane reduction for per-particle outputs; energies and pair counts
//     are reduced once per item.
#pragma once
#include <algorithm>
#include <type_traits>

#include "kparams.cuh"

namespace cgb {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// one 256-bit load of a sorted entry (one 32-byte sector); LDG.E...256 on sm_100a.  POLICY: 0 = non-coherent path allocating
// in L1 (neighbouring lanes gather from the same few cells, and what hits L1 does not wait for L2), 1 = L1-allocating with
// evict_last, 3 = non-coherent without L1 allocation (the gathers stream through L2), 4 = ld.global.cg.  Measured on B200
// (gather_policy below): the ion kernels, one gather per pair, are fastest with 0, the dipole / multipole kernels with 3.
template <int POLICY> __device__ __forceinline__ double4 ldg256(const double4 *p) {
    double4 v;
    if constexpr (POLICY == 0)
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    else if constexpr (POLICY == 1)
        asm volatile("ld.global.L1::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    else if constexpr (POLICY == 4)
        asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}

// Launch shape and gather policy per mode (tuning builds: -DCG_WARPS / -DCG_MINB / -DCG_GATHER override all modes, the _ION / _DIP /
// _MP forms one mode).  Every shape keeps 12 warps on an SM, i.e. up to 168 registers per thread: at 16 warps (128 registers) all
// kernels but plain Wolf spill their gather sets to local memory, and the spill traffic shares the LSU data pipe the gathers are
// bound by.  Measured (B200, full-size BASELINE configs, pair kernel only; 16-warp shapes in brackets): Wolf 1.75 ms (1.88),
// Wolf + Fennell in one pass 1.93 (2.07), ReactionField dipoles 0.57 (0.63), qPotential multipoles 2.71 (2.93), splined Poisson
// multipoles 8.7 (9.7).  Ion kernels run one 12-warp CTA per SM (the longer hit lists of a single CTA hold a whole item) with
// L1-allocating gathers, dipole and multipole kernels two 6-warp CTAs with streaming gathers.  10 and 14 warps per SM are slower.
#ifndef CG_WARPS_ION
#define CG_WARPS_ION 12
#endif
#ifndef CG_MINB_ION
#define CG_MINB_ION 1
#endif
#ifndef CG_WARPS_DIP
#define CG_WARPS_DIP 6
#endif
#ifndef CG_MINB_DIP
#define CG_MINB_DIP 2
#endif
#ifndef CG_WARPS_MP
#define CG_WARPS_MP 6
#endif
#ifndef CG_MINB_MP
#define CG_MINB_MP 2
#endif
#if defined(CG_WARPS) || defined(CG_MINB)
#ifndef CG_WARPS
#define CG_WARPS 12
#endif
#ifndef CG_MINB
#define CG_MINB 1
#endif
template <int MODE> constexpr int tuned_warps() { return CG_WARPS; }
template <int MODE> constexpr int tuned_minb() { return CG_MINB; }
#else
template <int MODE> constexpr int tuned_warps() { return MODE == cgd::kModeIon ? CG_WARPS_ION : (MODE == cgd::kModeDipole ? CG_WARPS_DIP : CG_WARPS_MP); }
template <int MODE> constexpr int tuned_minb() { return MODE == cgd::kModeIon ? CG_MINB_ION : (MODE == cgd::kModeDipole ? CG_MINB_DIP : CG_MINB_MP); }
#endif
template <int MODE> constexpr int gather_policy() {
#ifdef CG_GATHER
    return CG_GATHER;
#else
    return MODE == cgd::kModeIon ? 0 : 3;
#endif
}
// the quadrupole mode carries two tensors per pair in registers: half the warps (255 registers each) instead of spilling at 128
#ifndef CG_WARPS_QUAD
#define CG_WARPS_QUAD 8
#endif
#ifndef CG_MINB_QUAD
#define CG_MINB_QUAD 1
#endif
template <int MODE> constexpr int min_blocks_per_sm() { return MODE == cgd::kModeMultipoleQuad ? CG_MINB_QUAD : tuned_minb<MODE>(); }
template <int MODE> constexpr int warps_per_cta() { return MODE == cgd::kModeMultipoleQuad ? CG_WARPS_QUAD : tuned_warps<MODE>(); }
// pairs in flight per consumer lane and step
#ifndef CG_ILP_DIP
#define CG_ILP_DIP 2
#endif
#ifndef CG_ILP_MP
#define CG_ILP_MP 1
#endif
template <int MODE> constexpr int pairs_in_flight() {
    return MODE == cgd::kModeIon ? CG_ILP : (MODE == cgd::kModeDipole ? CG_ILP_DIP : (MODE == cgd::kModeMultipole ? CG_ILP_MP : 1));
}

// register sets of gathered j data per consumer lane (one being evaluated, the others in flight)
#ifndef CG_SETS
#define CG_SETS 2
#endif
#ifndef CG_SETS_DIP
#define CG_SETS_DIP 2
#endif
#ifndef CG_SETS_MP
#define CG_SETS_MP 2
#endif
template <int MODE> constexpr int gather_sets() {
    return MODE == cgd::kModeIon ? CG_SETS : (MODE == cgd::kModeDipole ? CG_SETS_DIP : (MODE == cgd::kModeMultipole ? CG_SETS_MP : 2));
}

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned to_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}
__device__ __forceinline__ int top_bit(unsigned m) { // position of the highest set bit (m != 0)
    int b;
    asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(m));
    return b;
}


// tuning builds (-DCG_KPROF): per-warp cycle counters, summed by the engine after the launch (CG_KPROF=1 in the environment)
#ifdef CG_KPROF
#define KPROF_DECL long long kp_t0 = clock64(), kp_mark = kp_t0, kp_c[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define KPROF_ADD(k)                     \
    do {                                 \
        const long long now = clock64(); \
        kp_c[k] += now - kp_mark;        \
        kp_mark = now;                   \
    } while (0)
#define KPROF_COUNT(k, v) kp_c[k] += (v)
#define KPROF_FLUSH()                                                                               \
    do {                                                                                            \
        if (lane == 0 && p.prof) {                                                                  \
            kp_c[7] = clock64() - kp_t0;                                                            \
            for (int k = 0; k < 8; k++) atomicAdd(p.prof + k, (unsigned long long)kp_c[k]);          \
        }                                                                                           \
    } while (0)
#else
#define KPROF_DECL
#define KPROF_ADD(k)
#define KPROF_COUNT(k, v)
#define KPROF_FLUSH()
#endif

// shared memory: the functor's table, then per warp one list buffer of (slots + 1) entries per lane
inline size_t pair_smem_bytes(int table_doubles, int slots, int warps) {
    const size_t table_bytes = ((size_t)table_doubles * 8 + 15) & ~(size_t)15;
    return table_bytes + (size_t)warps * (size_t)(slots + 1) * 32 * sizeof(uint2);
}

// T = double: the fp64 path (parity 1e-12).  T = float (CG_FP32): positions, the distance vector and the exact gate stay
// fp64, the pair arithmetic incl. S(q) runs in fp32, per-chunk fp32 partial sums are folded into fp64 accumulators.
template <class F, bool YUK, int MODE, int WHAT, class T>
__global__ void __launch_bounds__(32 * warps_per_cta<MODE>(), min_blocks_per_sm<MODE>())
pair_kernel(const __grid_constant__ KParams p, const __grid_constant__ F f, const double *__restrict__ dev_table,
            int table_doubles, int items, int slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_table = reinterpret_cast<double *>(smem_raw);
    const int table_bytes = (table_doubles * 8 + 15) & ~15;
    constexpr int kWarps = warps_per_cta<MODE>();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-lane lists of (hit mask, first candidate index) of the batches in which the lane had at least one hit
    uint2 *hits = reinterpret_cast<uint2 *>(smem_raw + table_bytes) + (size_t)warp * (size_t)(slots + 1) * 32 + lane;
    // `items` is only the launch bound (the sync-free path does not know the group count on the host); CTAs beyond the
    // real work leave before they touch anything.  After an overflow the cell table points past the sorted arrays.
    // tail split (kparams.cuh: tail_plan; launches with split == 1 only): the groups beyond `n_whole` become `tsplit` items each
    int real_items;
    {
        int n_whole = p.counts[p.count_index], tsplit = 1;
        if (p.tail_wave > 0 && p.split == 1) tail_plan(p.counts[p.count_index], p.tail_wave, n_whole, tsplit);
        real_items = p.counts[2] != 0 ? 0 : min(items, p.split > 1 ? p.counts[p.count_index] * p.split : n_whole + (p.counts[p.count_index] - n_whole) * tsplit);
    }
    if ((int)blockIdx.x * kWarps >= real_items) return;
    const double *tab = nullptr;
    if constexpr (F::kTable != cgd::kTableNone) { // Splined: the four Andrea tables; erfc family: the erfc/exp table
        if (F::kTable == cgd::kTableErfc && p.erfc_rep > 1) { // eight interleaved copies of the entries [0, erfc_kmax]
            for (int k = threadIdx.x; k < 8 * (p.erfc_kmax + 1); k += 32 * kWarps) {
                s_table[2 * k] = dev_table[2 * (k >> 3)];
                s_table[2 * k + 1] = dev_table[2 * (k >> 3) + 1];
            }
        } else {
            for (int k = threadIdx.x; k < table_doubles; k += 32 * kWarps) s_table[k] = dev_table[k];
        }
        __syncthreads();
        tab = s_table;
    }
    // what the functors get as their table: the pointer, or for the erfc family in fp64 the (replicated) table's addressing
    auto tabx = [&] {
        if constexpr (F::kTable == cgd::kTableErfc && std::is_same<T, double>::value) {
            const unsigned base = smem_addr(s_table);
            return p.erfc_rep > 1 ? cgd::ErfcTabRep{base + 16u * (threadIdx.x & 7u), 128u, (unsigned)p.erfc_kmax}
                                  : cgd::ErfcTabRep{base, 16u, (unsigned)p.erfc_kmax};
        } else {
            return tab;
        }
    }();
    const unsigned full = 0xffffffffu;
    const int split = p.split;
    const int dummy0 = min(p.counts[0], p.cap_sorted - kDummies); // first of the far-away dummy entries
    constexpr int kIlp = pairs_in_flight<MODE>();
    constexpr int kSets = gather_sets<MODE>();
    constexpr int kGather = gather_policy<MODE>();
    constexpr int kNS = F::kSchemes;
    typedef cgd::Vec3T<T> V3;
    constexpr bool kF32 = !std::is_same<T, double>::value;
    const int fg = lane >> 2, ft = lane & 3;
    // candidate <-> MMA column: column n of n-tile nt holds candidate 8 (n >> 1) + 2 nt + (n & 1) of the batch, so that the
    // packed sign bits come out in candidate order (bit b of the mask <-> candidate j0 + b).  Lane (g, t) loads component t
    // of the candidate of column g: four 32-byte pieces per n-tile, every sector of the batch exactly once.
    const int ccol = 8 * (fg >> 1) + (fg & 1);
    const int foff = 4 * ccol + ft;
    // Batches per chunk.  An item that does not fit the list buffer is cut into EQUAL chunks (the lanes' hit counts of a short
    // remainder chunk are uneven).  The batch count of an item is only known once it has been walked, so the cut follows the
    // previous item of this warp -- neighbouring groups of a sorted system see nearly the same number of candidates.
    int chunk_len = slots;
    KPROF_DECL;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1);
        item = __shfl_sync(full, item, 0);
        if (item >= real_items) break;
        int g = p.group_begin + item / split;
        int slice = item - (item / split) * split;
        // tail items (launches with split == 1 only): row slice `slice` of `rs` of a group of the last round -- every rs-th
        // row of the walk, where a j-split slice takes every split-th batch of every row
        int tail = 1; // (slot of the item's partial outputs) << 3 | row slices of the group (1: a whole group)
        if (p.tail_wave > 0 && split == 1) {
            int n_whole, tsplit;
            tail_plan(p.counts[p.count_index], p.tail_wave, n_whole, tsplit);
            if (item >= n_whole && tsplit > 1) {
                const int r = item - n_whole;
                g = p.group_begin + n_whole + r / tsplit;
                slice = r - (r / tsplit) * tsplit;
                tail = (r << 3) | tsplit;
            }
        }
        const int2 grp = p.groups[g];
        // the lane's own particle: sorted entry grp.x + lane
        const bool act = lane < grp.y;
        const int i = grp.x + (act ? lane : 0);
        const double4 pi = ldg256<kGather>(p.pos + i);
        const T zi = (T)pi.w;
        const float4 pf = p.pos32[i];
        V3 mui = {T(0), T(0), T(0)};
        if constexpr (MODE != cgd::kModeIon) {
            const double4 m = ldg256<kGather>(p.mu + i);
            mui = {(T)m.x, (T)m.y, (T)m.z};
        }
        cgd::QuadT<T> quadi = {};
        if constexpr (MODE == cgd::kModeMultipoleQuad) {
            const double4 a = ldg256<kGather>(p.quad + 2 * i), b = ldg256<kGather>(p.quad + 2 * i + 1);
            quadi = {(T)a.x, (T)a.y, (T)a.z, (T)a.w, (T)b.x, (T)b.y, (T)b.z};
        }
        // bounding box of the group (coordinates relative to the grid origin are >= 0: float order == int order)
        const float xmin = __int_as_float(__reduce_min_sync(full, __float_as_int(pf.x)));
        const float xmax = __int_as_float(__reduce_max_sync(full, __float_as_int(pf.x)));
        const float ymin = __int_as_float(__reduce_min_sync(full, __float_as_int(pf.y)));
        const float ymax = __int_as_float(__reduce_max_sync(full, __float_as_int(pf.y)));
        const float zmin = __int_as_float(__reduce_min_sync(full, __float_as_int(pf.z)));
        const float zmax = __int_as_float(__reduce_max_sync(full, __float_as_int(pf.z)));
        // A fragments: row = i-particle, k = (-2x', -2y', -2z', 1, -thr', 0, 0, 0), coordinates relative to the group centre
        const float Ox = 0.5f * (xmin + xmax), Oy = 0.5f * (ymin + ymax), Oz = 0.5f * (zmin + zmax);
        unsigned afrag[2][4];
        {
            const float xr = pf.x - Ox, yr = pf.y - Oy, zr = pf.z - Oz;
            // operand rounding (2^-11 relative each, cvt.rna): |c'|^2, the three products twice, thr'
            const float hx = 0.5f * (xmax - xmin), hy = 0.5f * (ymax - ymin), hz = 0.5f * (zmax - zmin);
            const float X = sqrtf(hx * hx + hy * hy + hz * hz), R = X + p.rc_m + 1e-3f;
            const float tf_margin = 1.25f * 4.8828125e-4f * (R * R + 4.0f * R * X + (float)p.rc2_m + X * X);
            // pass <=> |c'|^2 - 2 x'.c' < thr'; inactive slots never pass
            const float thrp = act ? (float)(p.rc2_m - ((double)xr * xr + (double)yr * yr + (double)zr * zr)) + tf_margin : -INFINITY;
            const float comp[4] = {-2.0f * xr, -2.0f * yr, -2.0f * zr, 1.0f};
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                for (int hh_ = 0; hh_ < 2; hh_++) {
                    const int src = 16 * mt + 8 * hh_ + fg;
                    const float v0 = __shfl_sync(full, comp[0], src), v1 = __shfl_sync(full, comp[1], src);
                    const float v2 = __shfl_sync(full, comp[2], src), v3 = __shfl_sync(full, comp[3], src);
                    const float th = __shfl_sync(full, thrp, src);
                    const float vt = ft == 0 ? v0 : (ft == 1 ? v1 : (ft == 2 ? v2 : v3));
                    afrag[mt][hh_] = to_tf32(vt);                     // a0 / a1: (row g [+8], col t)
                    afrag[mt][2 + hh_] = ft == 0 ? to_tf32(-th) : 0u; // a2 / a3: (row g [+8], col t + 4)
                }
            }
        }
        const float Ot = ft == 0 ? Ox : (ft == 1 ? Oy : (ft == 2 ? Oz : 0.0f));
        const unsigned bone = ft == 0 ? to_tf32(1.0f) : 0u; // B row 4: the column of ones that picks up -thr'
        const float rcm = p.rc_m;
        const int cy0 = clampi((int)floorf((ymin - rcm) * p.inv_cs_y - p.slack), 0, p.ny - 1);
        const int cy1 = clampi((int)floorf((ymax + rcm) * p.inv_cs_y + p.slack), 0, p.ny - 1);
        const int cz0 = clampi((int)floorf((zmin - rcm) * p.inv_cs_z - p.slack) - p.cz_off, 0, p.nz - 1);
        const int cz1 = clampi((int)floorf((zmax + rcm) * p.inv_cs_z + p.slack) - p.cz_off, 0, p.nz - 1);
        const int nyr = cy1 - cy0 + 1;
        const int nrows = nyr * (cz1 - cz0 + 1);
        // stride of the row walk: a prime of about nrows / 4 that does not divide nrows
        int rstride = 1;
#ifndef CG_ROWSTRIDE
#define CG_ROWSTRIDE 1
#endif
        if (CG_ROWSTRIDE) {
            const int primes[14] = {3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47};
#pragma unroll
            for (int k = 13; k >= 0; k--)
                if ((4 * primes[k] > nrows || k == 13) && primes[k] < nrows && nrows % primes[k] != 0) rstride = primes[k];
        }

        cgd::AccumT<T> acc[kNS]; // what the pair arithmetic accumulates into (one set per scheme of a multi-scheme functor)
        cgd::Accum acc64[kNS];   // fp32 path only: running fp64 totals
        unsigned int npairs = 0;

        // ---- iterator over the candidate batches (warp-uniform; jb/je hold one cell row's run per lane) ----
        int row_next = 0;   // next chunk of 32 rows to expand
        unsigned have = 0;  // lanes whose run of the current chunk is still unvisited
        int jb = 0, je = 0; // this lane's run in the current chunk
        int cur_j = 0, cur_end = 0;
        // j-split: batch k of the run of the item's r-th row belongs to slice (r + k) mod split.  Every slice gets candidates from
        // all rows and, over `split` consecutive runs, from every position along x, i.e. hits for every lane; cutting the
        // runs themselves (or a plain round robin, which resonates with the batches per run) would hand the left part of
        // each run to one slice and starve the lanes on the right.
        auto next_batch = [&](int &j0, int &n) -> bool {
            while (cur_j >= cur_end) { // next run
                if (have == 0) {       // next 32 rows
                    const int rs = tail & 7, ro = rs > 1 ? slice : 0;
                    if (row_next * rs + ro >= nrows) return false;
                    const int rr = (row_next + lane) * rs + ro;
                    const bool mine = rr < nrows;
                    row_next += 32;
                    jb = je = 0;
                    if (mine) {
                        const int rp = (int)(((unsigned)rr * (unsigned)rstride) % (unsigned)nrows); // the rr-th row of the walk
                        const int rzi = rp / nyr;
                        const int ry = cy0 + (rp - rzi * nyr), rz = cz0 + rzi;
                        const float ylo = (float)ry * p.cs_y, zlo = (float)(rz + p.cz_off) * p.cs_z;
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
                            if (split > 1) { // this slice's first batch of the run (see above); rr = the row's index in the walk
                                int first = (slice - rr) % split;
                                first += first < 0 ? split : 0;
                                jb += 32 * first;
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
            j0 = cur_j;
            n = min(32, cur_end - cur_j);
            cur_j += 32 * split;
            return true;
        };
        // software pipeline: the coordinates of the NEXT batch are in flight while the current one is tested
        int nj0 = 0, nn = 0;
        bool nvalid = next_batch(nj0, nn);
        float bpre[4];
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
            bpre[nt] = (nvalid && ccol + 2 * nt < nn && ft < 3) ? reinterpret_cast<const float *>(p.pos32 + nj0)[foff + 8 * nt] : 0.0f;
        int nb_item = 0;
        KPROF_ADD(0); // [0] item set-up, epilogue

        do {
            // ================= PHASE 1: prefilter up to chunk_len batches of 32 candidates into bit masks =================
            int nb = 0, cnt = 0, qn = 0;
            while (nb < chunk_len && nvalid) {
                const int j0 = nj0, n = nn;
                float braw[4];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) braw[nt] = bpre[nt];
                nvalid = next_batch(nj0, nn);
#pragma unroll
                for (int nt = 0; nt < 4; nt++)
                    bpre[nt] = (nvalid && ccol + 2 * nt < nn && ft < 3) ? reinterpret_cast<const float *>(p.pos32 + nj0)[foff + 8 * nt] : 0.0f;
                float dres[2][4][4];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const float cr = braw[nt] - Ot; // recentred component (t < 3)
                    float sq = ft < 3 ? cr * cr : 0.0f;
                    sq += __shfl_xor_sync(full, sq, 1);
                    sq += __shfl_xor_sync(full, sq, 2); // |c'|^2 in all four lanes of the quad
                    const bool live = ccol + 2 * nt < n;
                    const unsigned b0 = to_tf32(ft == 3 ? (live ? sq : INFINITY) : cr);
                    mma_tf32_16x8x8(dres[0][nt], afrag[0], b0, bone);
                    mma_tf32_16x8x8(dres[1][nt], afrag[1], b0, bone);
                }
                // sign bits -> one word per lane: byte k = 2 mt + h (row 16 mt + 8 h + g), bit 2 nt + cp (column 2 t + cp of n-tile
                // nt); four independent funnel-shift chains (one per byte), then one merge
                unsigned wb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int nt = 3; nt >= 0; nt--)
#pragma unroll
                    for (int cp = 1; cp >= 0; cp--)
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) wb[kk] = __funnelshift_l(__float_as_uint(dres[kk >> 1][nt][2 * (kk & 1) + cp]), wb[kk], 1);
                const unsigned W = wb[0] | (wb[1] << 8) | (wb[2] << 16) | (wb[3] << 24);
                // particle `lane` lives in row byte lane / 8 of the four lanes of quad lane % 8; the byte of lane t lands at
                // bits 8 t .. 8 t + 7, i.e. bit 8 t + 2 nt + cp = the candidate loaded into that column
                unsigned m = 0u;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const unsigned w = __shfl_sync(full, W, 4 * (lane & 7) + t);
                    m |= ((w >> (8 * (lane >> 3))) & 255u) << (8 * t);
                }
                const unsigned rel_self = (unsigned)(i - j0); // the lane's own entry is not a neighbour
                if (rel_self < 32u) m &= ~(1u << rel_self);
                if (m != 0u) hits[(qn++) * 32] = make_uint2(m, (unsigned)j0);
                cnt += __popc(m);
                nb++;
            }
            // close the list: the sentinel feeds a lane that runs dry with far-away dummy entries, for as long as it takes
            hits[qn * 32] = make_uint2(0xffffffffu, (unsigned)dummy0);
            nb_item += nb;
            const int rounds = (__reduce_max_sync(full, cnt) + kSets * kIlp - 1) / (kSets * kIlp);
            __syncwarp();
            KPROF_ADD(1); // [1] prefilter
            KPROF_COUNT(3, nb);
            KPROF_COUNT(4, rounds);
            KPROF_COUNT(5, 1);

            // ================= PHASE 2: evaluate the recorded pairs in FP64 =================
            {
                // The lane's cursor: `cur` = unvisited hits of the current list entry, `base` = its first candidate index.  The next
                // entry is fetched as soon as the current one is exhausted, one pop ahead of its first use; past the last real
                // entry the cursor stays on the sentinel (32 dummy entries, re-read as often as needed).
                const unsigned list = smem_addr(hits);          // entry e of this lane at list + 256 e
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
                // kSets register sets in a ring: the gathers of the next kSets - 1 steps are in flight while one step is evaluated
                int jj[kSets][kIlp]; // quadrupole mode: the tensors of j are fetched in the compute step (no room for prefetched sets)
                double4 pp[kSets][kIlp], mm[kSets][kIlp];
                auto fetch = [&](int s) {
#pragma unroll
                    for (int u = 0; u < kIlp; u++) {
                        const int j = pop();
                        jj[s][u] = j;
                        pp[s][u] = ldg256<kGather>(p.pos + j);
                        if constexpr (MODE != cgd::kModeIon) mm[s][u] = ldg256<kGather>(p.mu + j);
                    }
                };
                auto compute = [&](int s) {
#pragma unroll
                    for (int u = 0; u < kIlp; u++) {
                        const double dx = pi.x - pp[s][u].x, dy = pi.y - pp[s][u].y, dz = pi.z - pp[s][u].z;
                        const double r2 = dx * dx + dy * dy + dz * dz;
                        // the exact, strict reference gate (fp64 in both precisions).  Dummy entries lie far outside: they fail it
                        // (an unbounded cut-off -- Plain, plain or splined -- has no "beyond": the index tells a dummy from a neighbour)
                        const bool pass = r2 < p.rc2 && jj[s][u] < dummy0;
                        if (pass) npairs++;
                        // The arithmetic is branch-free and must stay finite for what fails the gate: prefilter false positives lie
                        // within the margin of the cut-off, dummies are clamped to a few cut-offs (the high word of a positive
                        // double orders like an integer), and `pass` zeroes the contribution.
                        const double r2c = __hiloint2double(min(__double2hiint(r2), p.r2cap_hi), __double2loint(r2));
                        const V3 d = {(T)dx, (T)dy, (T)dz};
                        V3 muj = {T(0), T(0), T(0)};
                        if constexpr (MODE != cgd::kModeIon) muj = {(T)mm[s][u].x, (T)mm[s][u].y, (T)mm[s][u].z};
                        if constexpr (kNS > 1) {
                            cgd::pair_accumulate_multi<F, WHAT, T>(f, p.core, tabx, d, (T)r2c, zi, (T)pp[s][u].w, acc, pass);
                        } else if constexpr (MODE == cgd::kModeMultipoleQuad) {
                            const double4 qa = ldg256<kGather>(p.quad + 2 * jj[s][u]), qb = ldg256<kGather>(p.quad + 2 * jj[s][u] + 1);
                            const cgd::QuadT<T> quadj = {(T)qa.x, (T)qa.y, (T)qa.z, (T)qa.w, (T)qb.x, (T)qb.y, (T)qb.z};
                            cgd::pair_accumulate<F, YUK, MODE, WHAT, T>(f, p.core, tabx, d, (T)r2c, zi, (T)pp[s][u].w, mui, muj, acc[0], pass, &quadi, &quadj);
                        } else {
                            cgd::pair_accumulate<F, YUK, MODE, WHAT, T>(f, p.core, tabx, d, (T)r2c, zi, (T)pp[s][u].w, mui, muj, acc[0], pass);
                        }
                    }
                };
                // every lane pops kSets kIlp entries per round (dry lanes: dummies)
#pragma unroll
                for (int k = 0; k < kSets - 1; k++) fetch(k);
#pragma unroll 1
                for (int r = 0; r < rounds; r++) {
#pragma unroll
                    for (int k = 0; k < kSets; k++) {
                        fetch((k + kSets - 1) % kSets); // the set evaluated in the previous step
                        compute(k);
                    }
                }
            }
            if constexpr (kF32) { // fold this chunk's fp32 partial sums into the fp64 totals
#pragma unroll
                for (int k = 0; k < kNS; k++) {
                    acc64[k].U += (double)acc[k].U;
                    acc64[k].Fx += (double)acc[k].Fx;
                    acc64[k].Fy += (double)acc[k].Fy;
                    acc64[k].Fz += (double)acc[k].Fz;
                    acc64[k].Ex += (double)acc[k].Ex;
                    acc64[k].Ey += (double)acc[k].Ey;
                    acc64[k].Ez += (double)acc[k].Ez;
                    acc64[k].phi += (double)acc[k].phi;
                    acc[k] = cgd::AccumT<T>();
                }
            }
            __syncwarp();
            KPROF_ADD(2); // [2] pair loop
        } while (nvalid);
        {
            const int nchunks = max(1, (nb_item + slots - 1) / slots);
            chunk_len = min(slots, (nb_item + nchunks - 1) / nchunks + 1); // + 1: small variations stay inside nchunks chunks
        }

        // ---- epilogue: per-particle outputs straight from registers, one reduction per item for the scalars ----
        unsigned int np = npairs; // inactive lanes only ever pop dummies
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) np += __shfl_down_sync(full, np, off);
        if (lane == 0) p.item_pairs[item] = np;
#pragma unroll
        for (int k = 0; k < kNS; k++) {
            if constexpr (!kF32) {
                acc64[k].U = (double)acc[k].U;
                acc64[k].Fx = (double)acc[k].Fx;
                acc64[k].Fy = (double)acc[k].Fy;
                acc64[k].Fz = (double)acc[k].Fz;
                acc64[k].Ex = (double)acc[k].Ex;
                acc64[k].Ey = (double)acc[k].Ey;
                acc64[k].Ez = (double)acc[k].Ez;
                acc64[k].phi = (double)acc[k].phi;
            }
            // scheme 0 writes the launch's main outputs, the further schemes of a multi-scheme functor their own
            double *o_force = k == 0 ? p.force : p.force_k[k > 0 ? k - 1 : 0], *o_field = k == 0 ? p.field : p.field_k[k > 0 ? k - 1 : 0];
            double *o_pot = k == 0 ? p.potential : p.potential_k[k > 0 ? k - 1 : 0];
            double *o_energy = k == 0 ? p.item_energy : p.item_energy_k[k > 0 ? k - 1 : 0];
            if (tail > 1) { // a tail item's partial outputs
                o_force = p.tail_out[k][0];
                o_field = p.tail_out[k][1];
                o_pot = p.tail_out[k][2];
            }
            double u = 0.0;
            if (act) {
                // tail items write slot (item - n_whole) * 32 + lane of the tail buffers, the others their particle's entry
                const size_t o = tail > 1 ? (size_t)(tail >> 3) * 32 + (size_t)lane : (size_t)slice * (size_t)p.n + (size_t)p.orig[i];
                if constexpr ((WHAT & cgd::kWhatForce) != 0) {
                    if (o_force) {
                        o_force[3 * o + 0] = acc64[k].Fx;
                        o_force[3 * o + 1] = acc64[k].Fy;
                        o_force[3 * o + 2] = acc64[k].Fz;
                    }
                }
                if constexpr ((WHAT & cgd::kWhatField) != 0) {
                    if (o_field) {
                        o_field[3 * o + 0] = acc64[k].Ex;
                        o_field[3 * o + 1] = acc64[k].Ey;
                        o_field[3 * o + 2] = acc64[k].Ez;
                    }
                }
                if constexpr ((WHAT & cgd::kWhatPotential) != 0) {
                    if (o_pot) o_pot[o] = acc64[k].phi;
                }
                if ((WHAT & cgd::kWhatEnergy) != 0) u = acc64[k].U;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) u += __shfl_down_sync(full, u, off);
            if (lane == 0) o_energy[item] = u;
        }
    }
    KPROF_FLUSH();
}

// list entries per lane that fit beside the table: all the shared memory an SM offers, shared by its resident warps
inline int pair_list_slots(int table_doubles, int ctas_per_sm, int warps) {
    const size_t table_bytes = ((size_t)table_doubles * 8 + 15) & ~(size_t)15;
    const size_t per_cta = (size_t)227 * 1024 / (size_t)ctas_per_sm - 1024; // 1 KB per CTA is the system's
    const long s = ((long)per_cta - (long)table_bytes) / (long)(warps * 32 * sizeof(uint2)) - 1;
    return (int)std::max(8L, std::min((long)CG_SLOTS, s));
}

template <class F, bool YUK, int MODE, int WHAT, class T>
cudaError_t launch_pair(const KParams &p, const cg_scheme_desc &desc, const double *dev_table, int table_doubles,
                        int items, int sms, cudaStream_t stream) {
    const F f(desc);
    if (F::kTable == cgd::kTableNone) table_doubles = 0;
    constexpr int ctas_per_sm = min_blocks_per_sm<MODE>(), warps = warps_per_cta<MODE>();
    const int slots = pair_list_slots(table_doubles, ctas_per_sm, warps);
    const size_t smem = pair_smem_bytes(table_doubles, slots, warps);
    auto kern = pair_kernel<F, YUK, MODE, WHAT, T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // persistent grid: one CTA per resident slot, never more than the work needs
    const int grid = std::min((items + warps - 1) / warps, sms * ctas_per_sm);
    if (grid <= 0) return cudaSuccess;
    kern<<<grid, 32 * warps, smem, stream>>>(p, f, dev_table, table_doubles, items, slots);
    return cudaGetLastError();
}

// per-functor probe of S..S''' (tests/test_gpu_functors.py): same functor object, same table staging as the pair kernel
template <class F> __global__ void srf_kernel(const __grid_constant__ F f, const double *__restrict__ dev_table, int table_doubles,
                                              int64_t n, const double *q, double *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_table = reinterpret_cast<double *>(smem_raw);
    const double *tab = nullptr;
    if constexpr (F::kTable != cgd::kTableNone) {
        for (int k = threadIdx.x; k < table_doubles; k += blockDim.x) s_table[k] = dev_table[k];
        __syncthreads();
        tab = s_table;
    }
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s[4];
    f.template eval<3>(q[k], s, tab);
    out[k] = s[0];
    out[n + k] = s[1];
    out[2 * n + k] = s[2];
    out[3 * n + k] = s[3];
}

template <class F>
cudaError_t launch_srf(const cg_scheme_desc &desc, const double *dev_table, int table_doubles, int64_t n, const double *q,
                       double *out, cudaStream_t stream) {
    const F f(desc);
    if (n <= 0) return cudaSuccess;
    if (F::kTable == cgd::kTableNone) table_doubles = 0;
    const size_t smem = (size_t)((table_doubles * 8 + 15) & ~15);
    srf_kernel<F><<<(unsigned)((n + 255) / 256), 256, smem, stream>>>(f, dev_table, table_doubles, n, q, out);
    return cudaGetLastError();
}

// table of launchers for one functor: the ten instantiated (mode, what) combinations, in fp64 and fp32
// (the two quadrupole combinations exist in fp64 only), and the all-pairs kernel's (allpairs_kernel.cuh, which the
// instantiating translation units include)
#define CGB_PAIR_ROW(FUNCTOR, YUK, T)                                                                                          \
    {launch_pair<FUNCTOR, YUK, combo(0).mode, combo(0).what, T>, launch_pair<FUNCTOR, YUK, combo(1).mode, combo(1).what, T>,   \
     launch_pair<FUNCTOR, YUK, combo(2).mode, combo(2).what, T>, launch_pair<FUNCTOR, YUK, combo(3).mode, combo(3).what, T>,   \
     launch_pair<FUNCTOR, YUK, combo(4).mode, combo(4).what, T>, launch_pair<FUNCTOR, YUK, combo(5).mode, combo(5).what, T>,   \
     launch_pair<FUNCTOR, YUK, combo(6).mode, combo(6).what, T>, launch_pair<FUNCTOR, YUK, combo(7).mode, combo(7).what, T>,   \
     launch_pair<FUNCTOR, YUK, combo(8).mode, combo(8).what, double>, launch_pair<FUNCTOR, YUK, combo(9).mode, combo(9).what, double>}
// ion modes only (combos 0, 1, 2, 7); the other slots stay empty
#define CGB_PAIR_ROW_ION(FUNCTOR, T)                                                                                       \
    {launch_pair<FUNCTOR, false, combo(0).mode, combo(0).what, T>, launch_pair<FUNCTOR, false, combo(1).mode, combo(1).what, T>, \
     launch_pair<FUNCTOR, false, combo(2).mode, combo(2).what, T>, nullptr, nullptr, nullptr, nullptr,                     \
     launch_pair<FUNCTOR, false, combo(7).mode, combo(7).what, T>, nullptr, nullptr}
#define CGB_DEFINE_FUNCTOR_ENTRY_ION(NAME, FUNCTOR)                                                                        \
    namespace cgb {                                                                                                       \
    const FunctorEntry *NAME() {                                                                                          \
        static const FunctorEntry e = {CGB_PAIR_ROW_ION(FUNCTOR, double), CGB_PAIR_ROW_ION(FUNCTOR, float), CGB_AP_ROW_ION(FUNCTOR), \
                                       ap_dims, launch_srf<FUNCTOR>, FUNCTOR::kTable};                                   \
        return &e;                                                                                                        \
    }                                                                                                                     \
    }
#define CGB_DEFINE_FUNCTOR_ENTRY(NAME, FUNCTOR, YUK)                                                                  \
    namespace cgb {                                                                                                   \
    const FunctorEntry *NAME() {                                                                                      \
        static const FunctorEntry e = {CGB_PAIR_ROW(FUNCTOR, YUK, double), CGB_PAIR_ROW(FUNCTOR, YUK, float), CGB_AP_ROW(FUNCTOR, YUK), \
                                       ap_dims, launch_srf<FUNCTOR>, FUNCTOR::kTable};                               \
        return &e;                                                                                                    \
    }                                                                                                                 \
    }

} // namespace cgb
