// kparams.cuh -- parameter blocks shared by the host engine and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "coulombgalore/detail/pair.hpp"
#include "coulombgalore/detail/srf.hpp"

namespace cgb {

namespace cgd = CoulombGalore::detail;

// tuning builds: make EXTRA_NVFLAGS="-DCG_WARPS=.. -DCG_MINB=.. -DCG_ILP=.. -DCG_SLOTS=.. -DCG_GATHER=.." (pair_kernel.cuh)
#ifndef CG_ILP
#define CG_ILP 3
#endif
#ifndef CG_SLOTS
#define CG_SLOTS 64
#endif
// CG_SLOTS: most candidate batches (of 32) whose per-lane hit masks are parked before the FP64 phase; the launch takes what the
// shared memory of an SM allows (pair_kernel.cuh: pair_list_slots)
constexpr int kMaxWarpsPerSm = 12; // resident warps per SM of every launch shape of the pair kernel (pair_kernel.cuh): the j-split fills them
constexpr int kDummies = 32;     // far-away entries behind the sorted arrays: what a lane without work evaluates

// Spatially sorted system.  With periodic boundaries the arrays also hold the periodic images that lie within one
// cut-off of the box ("ghost" entries, coordinates already shifted), so the pair kernel never applies a minimum
// image: d = r_i - r_j on the stored coordinates.
struct KParams {
    const double4 *pos;   // x, y, z, charge            [n_sorted + kDummies]
    const double4 *mu;    // mu_x, mu_y, mu_z, 0        [n_sorted + kDummies]  (may be null when no dipoles are needed)
    const double4 *quad;  // two per entry: {S_xx, S_yy, S_zz, trace}, {S_xy, S_xz, S_yz, 0}, S = Q + Q^T  [2 (n_sorted + kDummies)] (quadrupole mode)
    const float4 *pos32;  // position relative to the grid origin in fp32, w = |xyz|^2   [n_sorted]
    const int *orig;      // original particle index of a sorted entry               [n_sorted]
    const int *cell_start; // first sorted entry of each cell, x fastest             [ncell + 1]
    const int2 *groups;   // i-groups: (first sorted entry, count <= 32); a group never crosses a cell row
    int count_index;      // which of the device counts is the number of these groups (1: groups of 32, 3: groups of 64)
    int group_begin, group_end; // shard of groups this launch evaluates
    int split;            // j-split factor: work item = (group, s), s in [0, split)
    int nx, ny, nz;       // cells per dimension (ghost layers included; nz = the shard's layer window)
    int cz_off;           // first cell layer of the window in the full grid
    float inv_cs_x, inv_cs_y, inv_cs_z, cs_x, cs_y, cs_z;
    float rc_m;           // cut-off plus fp32 rounding margin, for the candidate cell ranges
    double rc2_m;         // squared cut-off plus the rounding margin of the three-FFMA prefilter test
    float slack;          // slack in cell units for fp32 cell-range rounding
    double rc2;           // exact gate: r2 < cutoff_squared (strict, reference core.hpp:354 etc.)
    int r2cap_hi;         // high word of the r2 above which the branch-free arithmetic is fed a clamped value (dummy entries)
    int *work_counter;    // next work item of this launch (zeroed by the host before the kernel)
    unsigned long long *prof; // tuning builds (-DCG_KPROF): 2 x 8 cycle counters, producers then consumers; else null
    cgd::CoreParams core;
    int n;                // number of (original) particles
    const int *counts;    // device counts of this evaluation: [0] sorted entries, [1] i-groups of 32, [2] overflow flag, [3] i-groups of 64
    int cap_sorted;       // capacity of the sorted arrays (the dummy entries start at min(counts[0], cap_sorted - kDummies))
    // outputs, indexed by original particle index; with split > 1 they are [split][n] partial buffers
    double *force, *field, *potential;
    // schemes 1.. of a multi-scheme functor (cg_eval_multi_device): their own outputs and item energies
    double *force_k[3], *field_k[3], *potential_k[3], *item_energy_k[3];
    double *item_energy;              // [items] sum_j u_ij over the item's pairs (not yet halved)
    unsigned long long *item_pairs;   // [items]
    // Tail split (split == 1 only; 0 = off): tail_wave = resident warps of the machine.  The groups beyond the last full
    // round of whole-group items are cut into `ts` row slices each (tail_plan), so that the last round of the persistent
    // grid is a round of short items; their partial outputs go to tail_out[scheme][force, field, potential]
    // (slot (item - n_whole) * 32 + lane: at most tail_wave * 32 slots) and are summed in slice order by the epilogue kernel.
    int tail_wave;
    double *tail_out[4][3];
    // erfc table in shared memory (specfun.hpp: ErfcTabRep): copies (1 or 8) and the last entry they hold
    int erfc_rep, erfc_kmax;
};
constexpr int kMaxTailSplit = 4;
// groups evaluated whole, and the number of row slices of each of the others
__host__ __device__ inline void tail_plan(int ng, int wave, int &n_whole, int &ts) {
    n_whole = ng;
    ts = 1;
    if (wave <= 0 || ng <= wave) return;
    const int whole = (ng / wave) * wave, rem = ng - whole;
    if (rem == 0) return;
    const int t = wave / rem < kMaxTailSplit ? wave / rem : kMaxTailSplit;
    if (t < 2) return;
    n_whole = whole;
    ts = t;
}
__host__ __device__ inline int tail_items(int ng, int wave) {
    int n_whole, ts;
    tail_plan(ng, wave, n_whole, ts);
    return n_whole + (ng - n_whole) * ts;
}

// All-pairs path for small systems (allpairs_kernel.cuh): no sort, no cell list -- the kernel reads the caller's SoA arrays
// as they are.  Work item = (block of i-groups, j-slice); every CTA streams its j-slice through shared memory in tiles
// (TMA bulk copies, two stages) and every warp keeps 32 i-particles in registers.
constexpr int kApTile = 512;      // most j's of one tile (a multiple of 32)
constexpr int kApMaxSlices = 64;  // j-slices per i-block: what the epilogue's slice sum is sized for
struct APParams {
    int n;                         // particles
    const double *x, *y, *z, *q;   // SoA inputs, 16-byte aligned (q may be null: no charges)
    const double *mx, *my, *mz;    // dipole moments (null in ion mode)
    int n_slices;                  // j-slices: slice s = the batches of 32 consecutive particles s, s + n_slices, ...
    double box[3];                 // periodic box (minimum image, box >= 2 cut-offs per dimension); 0 = open
    double rc2;                    // exact gate: r2 < cutoff^2 (strict)
    float rc2f;                    // the same for the fp32 prefilter (before the margin); +inf for an unbounded cut-off
    int r2cap_hi;                  // see KParams
    cgd::CoreParams core;
    int *counts;                   // [4]: written by the kernel for the epilogue: {n, i-groups, 0, 0}
    double *force, *field, *potential; // [n_slices][n] partial outputs (the caller's arrays when n_slices == 1)
    double *force_k[3], *field_k[3], *potential_k[3], *item_energy_k[3];
    double *item_energy;           // [i-groups * n_slices]
    unsigned long long *item_pairs;
};
typedef cudaError_t (*ap_launch_fn)(const APParams &p, const cg_scheme_desc &desc, const double *dev_table, int table_doubles,
                                    int sms, cudaStream_t stream);
// shared-memory layout of the all-pairs kernel (byte offsets): functor table | stages [2][arrays][kApTile + kDummies] doubles |
// fp32 tile [2][kApTile] float4 | hit lists [warps][(kApTile / 32 + 1) * 32] uint2 | mbarriers [2] | tile maxima [2]
struct ApLayout {
    size_t stage, f32, lists, bars, total;
};
inline ApLayout ap_layout(int table_doubles, int arrays, int warps) {
    ApLayout L;
    L.stage = ((size_t)table_doubles * 8 + 127) & ~(size_t)127;
    L.f32 = L.stage + (size_t)2 * arrays * (kApTile + kDummies) * sizeof(double);
    L.lists = L.f32 + (size_t)2 * kApTile * sizeof(float4);
    L.bars = L.lists + (size_t)warps * (kApTile / 32 + 1) * 32 * sizeof(uint2);
    L.total = L.bars + 2 * sizeof(unsigned long long) + 2 * sizeof(int) + 8;
    return L;
}

typedef cudaError_t (*pair_launch_fn)(const KParams &p, const cg_scheme_desc &desc, const double *dev_table,
                                      int table_doubles, int items, int sms, cudaStream_t stream);
typedef cudaError_t (*srf_launch_fn)(const cg_scheme_desc &desc, const double *dev_table, int table_doubles, int64_t n,
                                     const double *q, double *out, cudaStream_t stream);

// the (mode, what) combinations that are instantiated; a request maps to the smallest superset
struct Combo {
    int mode, what;
};
constexpr int kNumCombos = 10;
__host__ __device__ constexpr Combo combo(int i) {
    return i == 0   ? Combo{cgd::kModeIon, 2}
           : i == 1 ? Combo{cgd::kModeIon, 3}
           : i == 2 ? Combo{cgd::kModeIon, 15}
           : i == 3 ? Combo{cgd::kModeDipole, 7}
           : i == 4 ? Combo{cgd::kModeDipole, 15}
           : i == 5 ? Combo{cgd::kModeMultipole, 3}
           : i == 6 ? Combo{cgd::kModeMultipole, 15}
           : i == 7 ? Combo{cgd::kModeIon, 1}
           : i == 8 ? Combo{cgd::kModeMultipoleQuad, 3}
                    : Combo{cgd::kModeMultipoleQuad, 15};
}

struct FunctorEntry {
    pair_launch_fn pair[kNumCombos];   // fp64 pair arithmetic
    pair_launch_fn pair32[kNumCombos]; // fp32 pair arithmetic (CG_FP32); the quadrupole combinations are fp64 in both tables
    ap_launch_fn ap[kNumCombos];       // all-pairs path (fp64; no quadrupole combinations)
    void (*ap_dims)(int combo, int *warps, int *ctas_per_sm);
    srf_launch_fn srf;
    int table_kind; // cgd::kTableNone / kTableSpline / kTableErfc: which table the engine must pass
};

// one per functor family, defined in pair_inst_*.cu
enum FunctorId {
    kFnPlain = 0,
    kFnPlainYukawa,
    kFnReactionField,
    kFnPoisson,
    kFnPoissonYukawa,
    kFnQPotential,  // run-time order
    kFnQPotential3, // compile-time order 3 (BASELINE config 1)
    kFnFanourgakis,
    kFnWolf,
    kFnFennell,
    kFnZeroDipole,
    kFnZahn,
    kFnEwald,
    kFnEwaldScreened,
    kFnEwaldT,
    kFnSpline,
    kFnSplineYukawa,
    kFnErfcMulti2, // two schemes of the erfc family with one eta, ion modes (cg_eval_multi_device)
    kNumFunctors
};
const FunctorEntry *functor_entry(int id);

} // namespace cgb
