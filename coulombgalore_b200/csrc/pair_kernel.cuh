// pair_kernel.cuh -- the sm_100a pair-interaction kernel.  The pair arithmetic is elementwise transcendental / polynomial FP64
// math (FP64 pipe); the one dense product of the path, the 32 x 32 distance test of the prefilter, runs on the tensor cores in
// the ion and dipole kernels.
//
// Work decomposition
//   * one warp per work item = (i-group, j-split slice); an i-group is <= 32 spatially adjacent particles of one
//     cell row, one particle per lane, held in registers for the whole item (positions, charge, dipole, accumulators);
//   * the candidate j's are the particles of every cell that intersects the group's bounding box grown by the
//     cut-off; cells are stored x-fastest so each (y,z) cell row contributes ONE contiguous run of sorted entries.
//     The run table is built 32 rows at a time, one row per lane, with rounded-corner culling in the y-z plane;
//   * PHASE 1 (prefilter): candidates are tested 32 at a time against all particles of the group,
//           |c|^2 - 2 c.x_i  <  rc^2 + margin - |x_i|^2      (margin covers the rounding of the test's arithmetic)
//     either as eight TF32 mma.sync.m16n8k8 per batch (operands relative to the group centre, B fragments loaded coalesced
//     from the float4 array, results packed and transposed to one 32-bit mask per particle) or, in the register-bound
//     multipole kernels, with three FFMAs per test on candidates broadcast from a shared-memory stage.  Either way the test
//     is conservative, so the FP64 pipe never sees the ~80 % of candidates that are out of range.  Non-empty masks of up to
//     kSlots batches are parked in a per-lane list in shared memory (deferred evaluation);
//   * PHASE 2 (FP64 pipe): every lane walks the set bits of its own masks.  All lanes hold work until the lane
//     with the fewest hits runs dry, so the expensive FP64 pair arithmetic runs with (nearly) all lanes busy instead
//     of at the ~15-25 % in-cut-off hit rate a branch-per-candidate loop would give.  The evaluation re-tests the
//     exact reference gate r2 < cutoff^2 on the FP64 coordinates, gathers j with one 256-bit load through L1/L2
//     (sorted order keeps the neighbourhood resident) and accumulates into the lane's registers: no atomics, no
//     cross-lane reduction for per-particle outputs; energies/pair counts are reduced once per item.
#pragma once
#include <type_traits>

#include "kparams.cuh"

namespace cgb {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// one 256-bit load of a sorted entry (one 32-byte sector); LDG.E...256 on sm_100a.  CG_GATHER selects the cache
// policy for tuning runs: 0 = non-coherent path, 1 = L1-allocating with evict_last, 2 = two 128-bit loads, 3 = non-coherent
// without L1 allocation (default: the gathers stream through L2, measured 1-5 % faster), 4 = ld.global.cg.
#ifndef CG_GATHER
#define CG_GATHER 3
#endif
__device__ __forceinline__ double4 ldg256(const double4 *p) {
    double4 v;
#if CG_GATHER == 0
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#elif CG_GATHER == 1
    asm volatile("ld.global.L1::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#elif CG_GATHER == 3
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#elif CG_GATHER == 4
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
#else
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    v = make_double4(a.x, a.y, b.x, b.y);
#endif
    return v;
}

// the quadrupole mode carries two tensors per pair in registers: one CTA per SM (255 registers) instead of spilling at 128
template <int MODE> constexpr int min_blocks_per_sm() { return MODE == cgd::kModeMultipoleQuad ? 1 : CG_MINB; }
#ifndef CG_ILP_DIP
#define CG_ILP_DIP 1
#endif
// i-particles per lane (CG_IPL, default 1).  With 2, a lane of an ion-mode kernel owns two adjacent sorted entries (same or
// neighbouring cell, neighbour sets overlapping by ~75 %), groups are 64 entries long and one gathered j serves two pairs:
// 40 % fewer gathers and half the prefilter broadcasts per useful pair, for ~25 % more FP64 work (pairs only one of the
// two particles needs) and ~18 % more instructions in total.  Measured on B200 (Wolf force, N = 10^6): 2.47 ms against
// 1.95 ms with one particle per lane -- the kernel follows its instruction count (it issues ~0.6 instructions per cycle
// and scheduler at 4 warps per scheduler), not the L1/LSU wavefront count alone.  Kept as a tuning build only.
template <int MODE> constexpr int particles_per_lane() { return MODE == cgd::kModeIon ? CG_IPL : 1; }
template <int MODE> constexpr int pairs_in_flight() {
    return MODE == cgd::kModeIon ? (particles_per_lane<MODE>() > 1 ? 1 : CG_ILP) : CG_ILP_DIP;
}

// ---- tensor-core prefilter (CG_MMA_PREFILTER, default on; ion and dipole kernels) ----
// The distance test of 32 i-particles against 32 candidates is a 32 x 32 x 5 product: D[i][c] = |c'|^2 - 2 x'_i . c' - thr'_i
// (coordinates relative to the group centre), negative <=> candidate inside the margin-widened cut-off.  Eight TF32
// mma.sync.m16n8k8 per batch replace 32 broadcast shared-memory loads and 128 FFMA/FADD; the 2^-11 operand rounding is covered by
// a per-group margin (false positives, ~1 % more phase-2 steps, fall to the exact FP64 gate).  Measured: Wolf N = 10^6 1.955 ->
// 1.865 ms, ReactionField dipoles 0.653 -> 0.625 ms; the multipole kernels are register-bound and keep the FFMA prefilter.
#ifndef CG_MMA_PREFILTER
#define CG_MMA_PREFILTER 1
#endif
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
// bit b of a hit mask <-> candidate index within the batch
template <bool MMA> __device__ __forceinline__ int mask_bit_to_candidate(int b) {
    if (MMA) return ((b & 6) << 2) | ((b >> 3) << 1) | (b & 1); // b = 8 t + 2 nt + cp  <->  candidate 8 nt + 2 t + cp
    return b ^ 7;
}
template <bool MMA> __device__ __forceinline__ unsigned candidate_to_mask_bit(unsigned c) {
    if (MMA) return (((c >> 1) & 3u) << 3) | ((c >> 3) << 1) | (c & 1u);
    return c ^ 7u;
}
// which kernels use it: the light ones (ion and dipole modes), where the prefilter is a visible share of the instructions; the
// multipole kernels are register-bound and lose more to the fragment registers than they gain (measured)
template <int MODE> constexpr bool mma_prefilter() {
#ifndef CG_MMA_ALL_MODES
#define CG_MMA_ALL_MODES 0
#endif
    return CG_MMA_PREFILTER != 0 && particles_per_lane<MODE>() == 1 &&
           (CG_MMA_ALL_MODES != 0 || MODE == cgd::kModeIon || MODE == cgd::kModeDipole);
}

// T = double: the fp64 path (parity 1e-12).  T = float (CG_FP32): positions, the distance vector and the exact gate stay
// fp64, the pair arithmetic incl. S(q) runs in fp32, per-chunk fp32 partial sums are folded into fp64 accumulators.
template <class F, bool YUK, int MODE, int WHAT, class T>
__global__ void __launch_bounds__(kThreads, min_blocks_per_sm<MODE>())
pair_kernel(const __grid_constant__ KParams p, const __grid_constant__ F f, const double *__restrict__ dev_table,
            int table_doubles, int items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_table = reinterpret_cast<double *>(smem_raw);
    const int table_bytes = (table_doubles * 8 + 15) & ~15;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *stage = reinterpret_cast<float4 *>(smem_raw + table_bytes) + warp * 32;
    // per-lane lists of (hit mask, first candidate index) of the batches in which the lane had at least one hit
    uint2 *hits = reinterpret_cast<uint2 *>(smem_raw + table_bytes + kWarpsPerCta * 32 * sizeof(float4)) + warp * (kSlots * 32) + lane;
    constexpr int kIlp = pairs_in_flight<MODE>();
    // `items` is only the launch bound (the sync-free path does not know the group count on the host); CTAs beyond the
    // real work leave before they touch anything.  After an overflow the cell table points past the sorted arrays.
    const int real_items = p.counts[2] != 0 ? 0 : min(items, p.counts[p.count_index] * p.split);
    if ((int)blockIdx.x * kWarpsPerCta >= real_items) return;
    const double *tab = nullptr;
    if constexpr (F::kTable != cgd::kTableNone) { // Splined: the four Andrea tables; erfc family: the erfc/exp table
        for (int k = threadIdx.x; k < table_doubles; k += kThreads) s_table[k] = dev_table[k];
        __syncthreads();
        tab = s_table;
    }
    const int item = blockIdx.x * kWarpsPerCta + warp;
    const int split = p.split;
    if (item >= real_items) return;
    const int dummy = min(p.counts[0], p.cap_sorted - 1);
    const unsigned full = 0xffffffffu;
    const int g = p.group_begin + item / split, slice = item - (item / split) * split;
    const int2 grp = p.groups[g];
    constexpr int kIPL = particles_per_lane<MODE>();
    typedef cgd::Vec3T<T> V3;
    constexpr bool kF32 = !std::is_same<T, double>::value;

    // the lane's own particle(s): sorted entries grp.x + kIPL * lane + a
    bool act[kIPL];
    int iN[kIPL];
    double4 pi[kIPL];
    T zi[kIPL];
    float m2x[kIPL], m2y[kIPL], m2z[kIPL], thr[kIPL];
    float xmin = INFINITY, ymin = INFINITY, zmin = INFINITY, xmax = -INFINITY, ymax = -INFINITY, zmax = -INFINITY;
#pragma unroll
    for (int a = 0; a < kIPL; a++) {
        act[a] = kIPL * lane + a < grp.y;
        iN[a] = grp.x + (act[a] ? kIPL * lane + a : 0);
        pi[a] = ldg256(p.pos + iN[a]);
        zi[a] = (T)pi[a].w;
        const float4 pf = p.pos32[iN[a]];
        xmin = fminf(xmin, pf.x); xmax = fmaxf(xmax, pf.x);
        ymin = fminf(ymin, pf.y); ymax = fmaxf(ymax, pf.y);
        zmin = fminf(zmin, pf.z); zmax = fmaxf(zmax, pf.z);
        // prefilter constants: pass  <=>  |c|^2 + m2 . c < thr ; inactive slots never pass
        m2x[a] = -2.0f * pf.x;
        m2y[a] = -2.0f * pf.y;
        m2z[a] = -2.0f * pf.z;
        thr[a] = act[a] ? (float)(p.rc2_m - ((double)pf.x * pf.x + (double)pf.y * pf.y + (double)pf.z * pf.z)) : -INFINITY;
    }
    const int i = iN[0];
    V3 mui = {T(0), T(0), T(0)};
    if constexpr (MODE != cgd::kModeIon) {
        const double4 m = ldg256(p.mu + i);
        mui = {(T)m.x, (T)m.y, (T)m.z};
    }
    cgd::QuadT<T> quadi = {};
    if constexpr (MODE == cgd::kModeMultipoleQuad) {
        const double4 a = ldg256(p.quad + 2 * i), b = ldg256(p.quad + 2 * i + 1);
        quadi = {(T)a.x, (T)a.y, (T)a.z, (T)a.w, (T)b.x, (T)b.y, (T)b.z};
    }
    // bounding box of the group (coordinates relative to the grid origin are >= 0: float order == int order)
    xmin = __int_as_float(__reduce_min_sync(full, __float_as_int(xmin)));
    xmax = __int_as_float(__reduce_max_sync(full, __float_as_int(xmax)));
    ymin = __int_as_float(__reduce_min_sync(full, __float_as_int(ymin)));
    ymax = __int_as_float(__reduce_max_sync(full, __float_as_int(ymax)));
    zmin = __int_as_float(__reduce_min_sync(full, __float_as_int(zmin)));
    zmax = __int_as_float(__reduce_max_sync(full, __float_as_int(zmax)));

    constexpr bool kMma = mma_prefilter<MODE>();
    // A fragments: row = i-particle, k = (-2x', -2y', -2z', 1, -thr', 0, 0, 0), coordinates relative to the group centre
    const int fg = lane >> 2, ft = lane & 3;
    const float Ox = 0.5f * (xmin + xmax), Oy = 0.5f * (ymin + ymax), Oz = 0.5f * (zmin + zmax);
    unsigned afrag[2][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};
    float Ot = 0.0f;
    if constexpr (kMma) {
        const float4 pf0 = p.pos32[iN[0]];
        const float xr = pf0.x - Ox, yr = pf0.y - Oy, zr = pf0.z - Oz;
        // operand rounding (2^-11 relative each, cvt.rna): |c'|^2, the three products twice, thr'
        const float hx = 0.5f * (xmax - xmin), hy = 0.5f * (ymax - ymin), hz = 0.5f * (zmax - zmin);
        const float X = sqrtf(hx * hx + hy * hy + hz * hz), R = X + p.rc_m + 1e-3f;
        const float tf_margin = 1.25f * 4.8828125e-4f * (R * R + 4.0f * R * X + (float)p.rc2_m + X * X);
        const float thrp = act[0] ? (float)(p.rc2_m - ((double)xr * xr + (double)yr * yr + (double)zr * zr)) + tf_margin : -INFINITY;
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
                afrag[mt][hh_] = to_tf32(vt);                       // a0 / a1: (row g [+8], col t)
                afrag[mt][2 + hh_] = ft == 0 ? to_tf32(-th) : 0u;   // a2 / a3: (row g [+8], col t + 4)
            }
        }
        Ot = ft == 0 ? Ox : (ft == 1 ? Oy : (ft == 2 ? Oz : 0.0f));
    }
    const unsigned bone = ft == 0 ? to_tf32(1.0f) : 0u; // B row 4: the column of ones that picks up -thr'
    const float rcm = p.rc_m;
    const int cy0 = clampi((int)floorf((ymin - rcm) * p.inv_cs_y - p.slack), 0, p.ny - 1);
    const int cy1 = clampi((int)floorf((ymax + rcm) * p.inv_cs_y + p.slack), 0, p.ny - 1);
    const int cz0 = clampi((int)floorf((zmin - rcm) * p.inv_cs_z - p.slack) - p.cz_off, 0, p.nz - 1);
    const int cz1 = clampi((int)floorf((zmax + rcm) * p.inv_cs_z + p.slack) - p.cz_off, 0, p.nz - 1);
    const int nyr = cy1 - cy0 + 1;
    const int nrows = nyr * (cz1 - cz0 + 1);

    cgd::AccumT<T> acc[kIPL];  // what the pair arithmetic accumulates into
    cgd::Accum acc64[kIPL];    // fp32 path only: running fp64 totals
    unsigned int npairs = 0;

    // ---- iterator over the candidate batches (warp-uniform; jb/je hold one cell row's run per lane) ----
    int row_next = 0;       // next chunk of 32 rows to expand
    unsigned have = 0;      // lanes whose run of the current chunk is still unvisited
    int jb = 0, je = 0;     // this lane's run in the current chunk
    int cur_j = 0, cur_end = 0;
    // j-split: batch k of the run of the item's r-th row belongs to slice (r + k) mod split.  Every slice gets candidates from
    // all rows and, over `split` consecutive runs, from every position along x, i.e. hits for every lane; cutting the
    // runs themselves (or a plain round robin, which resonates with the batches per run) would hand the left part of
    // each run to one slice and starve the lanes on the right.
    auto next_batch = [&](int &j0, int &n) -> bool {
        while (cur_j >= cur_end) { // next run
            if (have == 0) {       // next 32 rows
                if (row_next >= nrows) return false;
                const int rr = row_next + lane;
                row_next += 32;
                jb = je = 0;
                if (rr < nrows) {
                    const int rzi = rr / nyr;
                    const int ry = cy0 + (rr - rzi * nyr), rz = cz0 + rzi;
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
                        if (split > 1) { // this slice's first batch of the run (see above); rr = the row's index in the item
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
        cur_j += 32 * p.split;
        return true;
    };
    const float4 sentinel = make_float4(0.0f, 0.0f, 0.0f, INFINITY); // |c|^2 = inf never passes
    // software pipeline: the coordinates of the NEXT batch are in flight while the current one is tested
    int nj0 = 0, nn = 0;
    bool nvalid = next_batch(nj0, nn);
    float bpre[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float4 cpre = sentinel;
    if constexpr (kMma) {
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
            bpre[nt] = (nvalid && 8 * nt + fg < nn && ft < 3) ? reinterpret_cast<const float *>(p.pos32 + nj0)[32 * nt + lane] : 0.0f;
    } else {
        if (nvalid && lane < nn) cpre = p.pos32[nj0 + lane];
    }

    while (nvalid) {
        // ================= PHASE 1: prefilter up to kSlots batches of 32 candidates into bit masks =================
        int nb = 0, cnt = 0, qn = 0;
        while (nb < kSlots && nvalid) {
            const int j0 = nj0, n = nn;
            unsigned m = 0u;
            if constexpr (kMma) {
            // B fragments straight from global memory: lane (g, t) holds component t of candidate 8 nt + g; the float4
            // array read as floats makes that one coalesced 128-byte load per n-tile
            float braw[4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) braw[nt] = bpre[nt];
            nvalid = next_batch(nj0, nn);
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
                bpre[nt] = (nvalid && 8 * nt + fg < nn && ft < 3) ? reinterpret_cast<const float *>(p.pos32 + nj0)[32 * nt + lane] : 0.0f;
            unsigned W = 0u;
            float dres[2][4][4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const float cr = braw[nt] - Ot;              // recentred component (t < 3)
                float sq = ft < 3 ? cr * cr : 0.0f;
                sq += __shfl_xor_sync(full, sq, 1);
                sq += __shfl_xor_sync(full, sq, 2);          // |c'|^2 in all four lanes of the quad
                const bool live = 8 * nt + fg < n;
                const unsigned b0 = to_tf32(ft == 3 ? (live ? sq : INFINITY) : cr);
                mma_tf32_16x8x8(dres[0][nt], afrag[0], b0, bone);
                mma_tf32_16x8x8(dres[1][nt], afrag[1], b0, bone);
            }
            // sign bits -> one word per lane: byte k = 2 mt + h (row 16 mt + 8 h + g), bit 2 nt + cp (candidate 8 nt + 2 t + cp);
            // four independent funnel-shift chains (one per byte), then one merge
            unsigned wb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int nt = 3; nt >= 0; nt--)
#pragma unroll
                for (int cp = 1; cp >= 0; cp--)
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) wb[kk] = __funnelshift_l(__float_as_uint(dres[kk >> 1][nt][2 * (kk & 1) + cp]), wb[kk], 1);
            W = wb[0] | (wb[1] << 8) | (wb[2] << 16) | (wb[3] << 24);
            // particle `lane` lives in row byte lane / 8 of the four lanes of quad lane % 8
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const unsigned w = __shfl_sync(full, W, 4 * (lane & 7) + t);
                m |= ((w >> (8 * (lane >> 3))) & 255u) << (8 * t);
            }
            } else {
            const float4 c = cpre;
            nvalid = next_batch(nj0, nn);
            cpre = sentinel;
            if (nvalid && lane < nn) cpre = p.pos32[nj0 + lane];
            __syncwarp();
            stage[lane] = c;
            __syncwarp();
            // candidate t0+u lands in bit t0 + 7 - u: the sign bit of (v - thr) is funnel-shifted into the mask
            // (one ALU instruction per test; NaN from the inf sentinel has a clear sign bit, i.e. does not pass)
#pragma unroll 1
            for (int t0 = 0; t0 < n; t0 += 8) {
                unsigned m8 = 0;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float4 cj = stage[t0 + u];
                    unsigned sb = 0u; // sign bit set <=> the candidate passes for at least one of the lane's particles
#pragma unroll
                    for (int a = 0; a < kIPL; a++)
                        sb |= __float_as_uint(fmaf(m2z[a], cj.z, fmaf(m2y[a], cj.y, fmaf(m2x[a], cj.x, cj.w))) - thr[a]);
                    m8 = __funnelshift_l(sb, m8, 1);
                }
                m |= m8 << t0;
            }
            }
            if constexpr (kIPL == 1) { // the lane's own entry is not a neighbour (with two particles per lane each is the
                const unsigned rel_self = (unsigned)(i - j0); // other's neighbour: the index check moves to phase 2)
                if (rel_self < 32u) m &= ~(1u << candidate_to_mask_bit<kMma>(rel_self));
            }
            if (m != 0u) hits[(qn++) * 32] = make_uint2(m, (unsigned)j0);
            cnt += __popc(m);
            nb++;
        }
        __syncwarp();
        // ================= PHASE 2: evaluate the recorded pairs in FP64 =================
        // kIlp independent pairs per lane and step, the gathers of the NEXT step already in flight (two register
        // sets used in ping-pong), and no branch inside the arithmetic (the exact gate becomes a select).
        // The lane's cursor: `cur` = unvisited hits of the current list entry, `base` = its first candidate index.
        // The next entry is fetched as soon as the current one is exhausted, one pop ahead of its first use, so the
        // shared-memory latency hides behind the pair arithmetic; pop() itself is straight-line code.  A lane that
        // has run dry gathers the far-away dummy entry p.dummy (finite arithmetic, contribution zeroed).
        int slot = 0;
        unsigned cur = 0u;
        int base = 0;
        if (qn > 0) {
            const uint2 e0 = hits[0];
            cur = e0.x;
            base = (int)e0.y;
        }
        auto pop = [&](bool &has) -> int {
            has = cnt > 0;
            const int b = 31 - __clz((int)cur); // highest set bit (-1 when empty)
            const int j = has ? base + mask_bit_to_candidate<kMma>(b) : dummy;
            cur &= ~(has ? (1u << b) : 0u);
            cnt -= has ? 1 : 0;
            if (cur == 0u && cnt > 0) {
                const uint2 e = hits[(++slot) * 32];
                cur = e.x;
                base = (int)e.y;
            }
            return j;
        };
        bool hh[2][kIlp];
        int jj[2][kIlp]; // quadrupole mode: the tensors of j are fetched in the compute step (no room for a prefetched set)
        double4 pp[2][kIlp], mm[2][kIlp];
        auto fetch = [&](auto S) {
            constexpr int s = decltype(S)::value;
#pragma unroll
            for (int u = 0; u < kIlp; u++) {
                const int j = pop(hh[s][u]);
                jj[s][u] = j;
                pp[s][u] = ldg256(p.pos + j);
                if constexpr (MODE != cgd::kModeIon) mm[s][u] = ldg256(p.mu + j);
            }
        };
        auto compute = [&](auto S) {
            constexpr int s = decltype(S)::value;
#pragma unroll
            for (int u = 0; u < kIlp; u++) {
#pragma unroll
                for (int a = 0; a < kIPL; a++) {
                    const double dx = pi[a].x - pp[s][u].x, dy = pi[a].y - pp[s][u].y, dz = pi[a].z - pp[s][u].z;
                    const double r2 = dx * dx + dy * dy + dz * dz;
                    // the exact, strict reference gate (fp64 in both precisions); a hit of the lane's other particle may be
                    // this particle itself or lie outside its own cut-off
                    const bool mine = hh[s][u] && (kIPL == 1 || (act[a] && jj[s][u] != iN[a]));
                    const bool pass = mine && r2 < p.rc2;
                    npairs += pass ? 1u : 0u;
                    const V3 d = {(T)dx, (T)dy, (T)dz};
                    V3 muj = {T(0), T(0), T(0)};
                    if constexpr (MODE != cgd::kModeIon) muj = {(T)mm[s][u].x, (T)mm[s][u].y, (T)mm[s][u].z};
                    // a lane that ran dry (or looks at itself) evaluates a harmless in-range distance, its contribution is
                    // zeroed by `pass`; rejected real pairs of a single-particle lane lie within the prefilter margin of the
                    // cut-off, those of the second particle anywhere below ~3 cut-offs: finite arithmetic in all schemes
                    const T r2e = mine ? (T)r2 : (T)p.rc2_safe;
                    if constexpr (MODE == cgd::kModeMultipoleQuad) {
                        const double4 qa = ldg256(p.quad + 2 * jj[s][u]), qb = ldg256(p.quad + 2 * jj[s][u] + 1);
                        const cgd::QuadT<T> quadj = {(T)qa.x, (T)qa.y, (T)qa.z, (T)qa.w, (T)qb.x, (T)qb.y, (T)qb.z};
                        cgd::pair_accumulate<F, YUK, MODE, WHAT, T>(f, p.core, tab, d, r2e, zi[a], (T)pp[s][u].w, mui, muj, acc[a], pass, &quadi,
                                                                    &quadj);
                    } else {
                        cgd::pair_accumulate<F, YUK, MODE, WHAT, T>(f, p.core, tab, d, r2e, zi[a], (T)pp[s][u].w, mui, muj, acc[a], pass);
                    }
                }
            }
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        fetch(I0{});
        for (;;) {
            if (!__any_sync(full, hh[0][0])) break;
            fetch(I1{});
            compute(I0{});
            if (!__any_sync(full, hh[1][0])) break;
            fetch(I0{});
            compute(I1{});
        }
        if constexpr (kF32) { // fold this chunk's fp32 partial sums into the fp64 totals
#pragma unroll
            for (int a = 0; a < kIPL; a++) {
                acc64[a].U += (double)acc[a].U;
                acc64[a].Fx += (double)acc[a].Fx;
                acc64[a].Fy += (double)acc[a].Fy;
                acc64[a].Fz += (double)acc[a].Fz;
                acc64[a].Ex += (double)acc[a].Ex;
                acc64[a].Ey += (double)acc[a].Ey;
                acc64[a].Ez += (double)acc[a].Ez;
                acc64[a].phi += (double)acc[a].phi;
                acc[a] = cgd::AccumT<T>();
            }
        }
        __syncwarp();
    }
    if constexpr (!kF32) {
#pragma unroll
        for (int a = 0; a < kIPL; a++) {
            acc64[a].U = (double)acc[a].U;
            acc64[a].Fx = (double)acc[a].Fx;
            acc64[a].Fy = (double)acc[a].Fy;
            acc64[a].Fz = (double)acc[a].Fz;
            acc64[a].Ex = (double)acc[a].Ex;
            acc64[a].Ey = (double)acc[a].Ey;
            acc64[a].Ez = (double)acc[a].Ez;
            acc64[a].phi = (double)acc[a].phi;
        }
    }

    // ---- epilogue: per-particle outputs straight from registers, one reduction per item for the scalars ----
    const size_t part = (size_t)slice * (size_t)p.n;
    double u = 0.0;
#pragma unroll
    for (int a = 0; a < kIPL; a++) {
        if (!act[a]) continue;
        const size_t o = part + (size_t)p.orig[iN[a]];
        if constexpr ((WHAT & cgd::kWhatForce) != 0) {
            if (p.force) {
                p.force[3 * o + 0] = acc64[a].Fx;
                p.force[3 * o + 1] = acc64[a].Fy;
                p.force[3 * o + 2] = acc64[a].Fz;
            }
        }
        if constexpr ((WHAT & cgd::kWhatField) != 0) {
            if (p.field) {
                p.field[3 * o + 0] = acc64[a].Ex;
                p.field[3 * o + 1] = acc64[a].Ey;
                p.field[3 * o + 2] = acc64[a].Ez;
            }
        }
        if constexpr ((WHAT & cgd::kWhatPotential) != 0) {
            if (p.potential) p.potential[o] = acc64[a].phi;
        }
        if ((WHAT & cgd::kWhatEnergy) != 0) u += acc64[a].U;
    }
    unsigned int np = npairs; // inactive slots never pass the gate
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

inline size_t pair_smem_bytes(int table_doubles) {
    const int table_bytes = (table_doubles * 8 + 15) & ~15;
    return (size_t)table_bytes + kWarpsPerCta * (32 * sizeof(float4) + kSlots * 32 * sizeof(uint2));
}

template <class F, bool YUK, int MODE, int WHAT, class T>
cudaError_t launch_pair(const KParams &p, const cg_scheme_desc &desc, const double *dev_table, int table_doubles,
                        int items, cudaStream_t stream) {
    const F f(desc);
    if (F::kTable == cgd::kTableNone) table_doubles = 0;
    const size_t smem = pair_smem_bytes(table_doubles);
    auto kern = pair_kernel<F, YUK, MODE, WHAT, T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (items + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid <= 0) return cudaSuccess;
    kern<<<grid, kThreads, smem, stream>>>(p, f, dev_table, table_doubles, items);
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
// (the two quadrupole combinations exist in fp64 only)
#define CGB_PAIR_ROW(FUNCTOR, YUK, T)                                                                                          \
    {launch_pair<FUNCTOR, YUK, combo(0).mode, combo(0).what, T>, launch_pair<FUNCTOR, YUK, combo(1).mode, combo(1).what, T>,   \
     launch_pair<FUNCTOR, YUK, combo(2).mode, combo(2).what, T>, launch_pair<FUNCTOR, YUK, combo(3).mode, combo(3).what, T>,   \
     launch_pair<FUNCTOR, YUK, combo(4).mode, combo(4).what, T>, launch_pair<FUNCTOR, YUK, combo(5).mode, combo(5).what, T>,   \
     launch_pair<FUNCTOR, YUK, combo(6).mode, combo(6).what, T>, launch_pair<FUNCTOR, YUK, combo(7).mode, combo(7).what, T>,   \
     launch_pair<FUNCTOR, YUK, combo(8).mode, combo(8).what, double>, launch_pair<FUNCTOR, YUK, combo(9).mode, combo(9).what, double>}
#define CGB_DEFINE_FUNCTOR_ENTRY(NAME, FUNCTOR, YUK)                                                                  \
    namespace cgb {                                                                                                   \
    const FunctorEntry *NAME() {                                                                                      \
        static const FunctorEntry e = {CGB_PAIR_ROW(FUNCTOR, YUK, double), CGB_PAIR_ROW(FUNCTOR, YUK, float), launch_srf<FUNCTOR>, \
                                       FUNCTOR::kTable};                                                             \
        return &e;                                                                                                    \
    }                                                                                                                 \
    }

} // namespace cgb
