// kparams.cuh -- parameter blocks shared by the host engine and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "coulombgalore/detail/pair.hpp"
#include "coulombgalore/detail/srf.hpp"

namespace cgb {

namespace cgd = CoulombGalore::detail;

// launch shape (overridable for tuning builds: make EXTRA_NVFLAGS="-DCG_WARPS=.. -DCG_MINB=.. -DCG_ILP=.. -DCG_SLOTS=..")
#ifndef CG_WARPS
#define CG_WARPS 8
#endif
#ifndef CG_MINB
#define CG_MINB 2
#endif
#ifndef CG_ILP
#define CG_ILP 2
#endif
#ifndef CG_SLOTS
#define CG_SLOTS 40
#endif
// i-particles per lane in the ion modes (pair_kernel.cuh: particles_per_lane); decides which i-group table they use.
// 2 was measured and rejected (see pair_kernel.cuh); the code path stays as a tuning build (-DCG_IPL=2).
#ifndef CG_IPL
#define CG_IPL 1
#endif
constexpr int kWarpsPerCta = CG_WARPS;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kSlots = CG_SLOTS; // candidate batches (of 32) whose per-lane hit masks are parked before the FP64 phase

// Spatially sorted system.  With periodic boundaries the arrays also hold the periodic images that lie within one
// cut-off of the box ("ghost" entries, coordinates already shifted), so the pair kernel never applies a minimum
// image: d = r_i - r_j on the stored coordinates.
struct KParams {
    const double4 *pos;   // x, y, z, charge            [n_sorted + 1]
    const double4 *mu;    // mu_x, mu_y, mu_z, 0        [n_sorted + 1]  (may be null when no dipoles are needed)
    const double4 *quad;  // two per entry: {S_xx, S_yy, S_zz, trace}, {S_xy, S_xz, S_yz, 0}, S = Q + Q^T  [2 (n_sorted + 1)] (quadrupole mode)
    const float4 *pos32;  // position relative to the grid origin in fp32, w = |xyz|^2   [n_sorted]
    const int *orig;      // original particle index of a sorted entry               [n_sorted]
    const int *cell_start; // first sorted entry of each cell, x fastest             [ncell + 1]
    const int2 *groups;   // i-groups: (first sorted entry, count <= 32 * particles per lane); a group never crosses a cell row
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
    double rc2_safe;      // cutoff_squared / 4 (1 for Plain): the r2 a lane without work feeds to the branch-free arithmetic
    cgd::CoreParams core;
    int n;                // number of (original) particles
    const int *counts;    // device counts of this evaluation: [0] sorted entries, [1] i-groups of 32, [2] overflow flag, [3] i-groups of 64
    int cap_sorted;       // capacity of the sorted arrays (the far-away dummy entry sits at min(counts[0], cap_sorted - 1))
    // outputs, indexed by original particle index; with split > 1 they are [split][n] partial buffers
    double *force, *field, *potential;
    double *item_energy;              // [items] sum_j u_ij over the item's pairs (not yet halved)
    unsigned long long *item_pairs;   // [items]
};

typedef cudaError_t (*pair_launch_fn)(const KParams &p, const cg_scheme_desc &desc, const double *dev_table,
                                      int table_doubles, int items, cudaStream_t stream);
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
    kNumFunctors
};
const FunctorEntry *functor_entry(int id);

} // namespace cgb
