/*
 * cg_b200.h -- C ABI of the B200 batched pair-interaction engine (libcg_b200.so).
 *
 * The reference (mlund/coulombgalore) is a header-only C++ library whose public surface is the scheme classes
 * (reference include/coulombgalore/core.hpp:110-244 `SchemeBase`, :251-843 `EnergyImplementation<T>`); every call
 * evaluates ONE pair on the CPU and the caller owns the loop over pairs.  This ABI is the boundary behind which
 * that loop now runs on the GPU: plain pointers and sizes, int status codes, no C++ or torch types.
 *
 *   reference call (one pair)                                   batched replacement (all pairs within the cut-off)
 *   ----------------------------------------------------------  ---------------------------------------------------
 *   T::T(cutoff, ...)             e.g. wolf.hpp:44              cg_scheme_wolf(...) / include/coulombgalore/wolf.hpp
 *   pot.ion_potential(z,r)        core.hpp:268                  cg_eval(mode=CG_MODE_ION,       what=CG_POTENTIAL)
 *   pot.ion_field(z,r)            core.hpp:352                  cg_eval(mode=CG_MODE_ION,       what=CG_FIELD)
 *   pot.ion_ion_energy(zA,zB,r)   core.hpp:501                  cg_eval(mode=CG_MODE_ION,       what=CG_ENERGY)
 *   pot.ion_ion_force(zA,zB,r)    core.hpp:630                  cg_eval(mode=CG_MODE_ION,       what=CG_FORCE)
 *   pot.dipole_potential(mu,r)    core.hpp:294                  cg_eval(mode=CG_MODE_DIPOLE,    what=CG_POTENTIAL)
 *   pot.dipole_field(mu,r)        core.hpp:380                  cg_eval(mode=CG_MODE_DIPOLE,    what=CG_FIELD)
 *   pot.dipole_dipole_energy      core.hpp:543                  cg_eval(mode=CG_MODE_DIPOLE,    what=CG_ENERGY)
 *   pot.dipole_dipole_force       core.hpp:677                  cg_eval(mode=CG_MODE_DIPOLE,    what=CG_FORCE)
 *   pot.multipole_field           core.hpp:457                  cg_eval(mode=CG_MODE_MULTIPOLE, what=CG_FIELD)
 *   pot.multipole_multipole_energy core.hpp:581                 cg_eval(mode=CG_MODE_MULTIPOLE, what=CG_ENERGY)
 *   pot.multipole_multipole_force core.hpp:737                  cg_eval(mode=CG_MODE_MULTIPOLE, what=CG_FORCE)
 *   pot.short_range_function*(q)  core.hpp:181-184              cg_srf_device(...) (per-functor parity probe)
 *
 * Batched semantics (identical to oracle/harness.hpp): for every particle i, sum over j != i with
 * |d| < cutoff, d = r_i - r_j (minimum image when pbc): potential/field use d; the ion force on i is
 * ion_ion_force(z_j, z_i, d); the dipole force on i is dipole_dipole_force(mu_i, mu_j, -d); multipole energy and
 * force are the reference functions called with A = i, B = j, r = r_j - r_i and zero quadrupoles (the given ones with
 * CG_MODE_MULTIPOLE_QUAD, where the potential also adds quadrupole_potential(quad_j, d)); U = 1/2 sum u_ij.
 *
 * All entry points return CG_OK (0) or a negative cg_status; none throws.  A handle owns its device buffers and one
 * CUDA stream; it is not safe to call one handle from two threads at once, separate handles are independent.
 * There is NO CPU fallback: without a CUDA device cg_create returns CG_ERR_NO_DEVICE.
 */
#ifndef CG_B200_H
#define CG_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum cg_status {
    CG_OK = 0,
    CG_ERR_INVALID = -1,   /* bad argument / scheme parameters rejected (mirrors the reference ctor throws) */
    CG_ERR_NO_DEVICE = -2, /* no CUDA device / driver */
    CG_ERR_CUDA = -3,      /* a CUDA call failed; see cg_last_error */
    CG_ERR_ALLOC = -4,
    CG_ERR_UNSUPPORTED = -5, /* valid request outside what the kernels cover (e.g. pbc box shorter than cutoff) */
    CG_ERR_STATE = -6        /* e.g. cg_eval before cg_set_system */
} cg_status;

/* same order as the reference's `enum class Scheme` (core.hpp:56-70) */
typedef enum cg_scheme_id {
    CG_PLAIN = 0,
    CG_EWALD,
    CG_EWALDT,
    CG_REACTIONFIELD,
    CG_WOLF,
    CG_POISSON,
    CG_QPOTENTIAL,
    CG_FANOURGAKIS,
    CG_ZERODIPOLE,
    CG_ZAHN,
    CG_FENNELL,
    CG_QPOTENTIAL5,
    CG_SPLINE
} cg_scheme_id;

#define CG_MAX_POISSON_C 12
#define CG_SPLINE_MAX_KNOTS 64 /* per table; the reference default utol=1e-3 needs <= 9 */

/*
 * POD scheme descriptor: what SURVEY.md 8b calls the "scheme -> device parameter export".  The reference keeps
 * these constants in private members (wolf.hpp:35, poisson.hpp:64-70, ewald.hpp:115-117, splined.hpp:323-325);
 * the new header classes (the headers under include/coulombgalore/) expose them through `descriptor()`, and the cg_scheme_*
 * functions below fill the same struct for C / ctypes callers.  Every constant is computed with the expression the
 * reference constructor uses.
 */
typedef struct cg_scheme_desc {
    int32_t functor;  /* cg_scheme_id of the S(q) functor that is evaluated (CG_SPLINE for tabulated schemes) */
    int32_t scheme;   /* value of the reference's public `scheme` member (Splined copies the wrapped id) */
    int32_t order;    /* qPotential: order */
    int32_t C, D;     /* Poisson */
    int32_t yukawa_map; /* Poisson: use_yukawa_screening (poisson.hpp:95-102) */
    int32_t shifted;  /* ReactionField */
    int32_t reserved0;
    double cutoff, inverse_cutoff, cutoff_squared; /* core.hpp:116-117,151 */
    double debye_length, inverse_debye_length;     /* core.hpp:118,152 (kappa of the core formulas) */
    double eta, zeta;        /* erfc family: alpha*cutoff ; Ewald: cutoff/debye_length (0 when unscreened) */
    double erfc_eta;         /* std::erfc(eta) */
    double exp_eta2;         /* std::exp(-eta*eta) */
    double F0;               /* EwaldTruncated scaling (ewald_truncated.hpp:59) */
    double rf_a, rf_b;       /* ReactionField: (epsRF-epsr)/(2epsRF+epsr), 3epsRF/(2epsRF+epsr)*shifted */
    double eps_rf, eps_r;    /* ReactionField constructor arguments (kept for the literal host evaluation) */
    double reduced_kappa, yukawa_denom, binomCDC;  /* Poisson (poisson.hpp:92-107) */
    double poisson_w[CG_MAX_POISSON_C];            /* binomial(D-1+c,c)*(C-c)/C, c = 0..C-1 */
    double self_energy_prefactor[2], self_field_prefactor[2], T0, chi; /* O(N) terms, core.hpp:112-121 */
    /* CG_SPLINE: four independent tables (S, S', S'', S'''), knots ascending in q, 6 coefficients per interval */
    int32_t spline_knots[4];
    const double *spline_r2[4];
    const double *spline_c[4];
} cg_scheme_desc;

/* Constructors mirroring the reference constructors; they validate like the reference does. */
int cg_scheme_plain(double debye_length, cg_scheme_desc *out);                                  /* plain.hpp:37 */
int cg_scheme_reactionfield(double cutoff, double epsRF, double epsr, int shifted, cg_scheme_desc *out); /* reactionfield.hpp:47 */
int cg_scheme_poisson(double cutoff, int C, int D, double debye_length, cg_scheme_desc *out);   /* poisson.hpp:79 */
int cg_scheme_qpotential(double cutoff, int order, cg_scheme_desc *out);                        /* qpotential.hpp:253 */
int cg_scheme_qpotential5(double cutoff, cg_scheme_desc *out);                                  /* qpotential.hpp:205, order 5 */
int cg_scheme_fanourgakis(double cutoff, cg_scheme_desc *out);                                  /* fanourgakis.hpp:39 */
int cg_scheme_fennell(double cutoff, double alpha, cg_scheme_desc *out);                        /* fennell.hpp:44 */
int cg_scheme_zerodipole(double cutoff, double alpha, cg_scheme_desc *out);                     /* zerodipole.hpp:44 */
int cg_scheme_zahn(double cutoff, double alpha, cg_scheme_desc *out);                           /* zahn.hpp:45 */
int cg_scheme_wolf(double cutoff, double alpha, cg_scheme_desc *out);                           /* wolf.hpp:44 */
int cg_scheme_ewald(double cutoff, double alpha, double eps_sur, double debye_length, cg_scheme_desc *out); /* ewald.hpp:129 */
int cg_scheme_ewaldt(double cutoff, double alpha, double eps_sur, double debye_length, cg_scheme_desc *out); /* ewald_truncated.hpp:47 */
/* Run-time factory: the reference's createScheme(json) (all.hpp:42-104) with the JSON object spelled out as a type
 * keyword ("plain", "wolf", "zahn", "fennell", "zerodipole", "fanourgakis", "qpotential", "ewald", "ewaldt",
 * "poisson", "reactionfield") and (key, value) pairs with the reference's keys: cutoff, alpha, order, C, D, epsRF, epsr,
 * shifted (0/1), epss, debyelength (the last two optional, default infinity).  CG_ERR_INVALID: unknown keyword, a
 * mandatory key missing, or parameters the constructor rejects; CG_ERR_UNSUPPORTED: "spline" (the reference's factory
 * returns a null pointer for it). */
int cg_scheme_create(const char *type, int nkeys, const char *const *keys, const double *values, cg_scheme_desc *out);
/*
 * Splined().spline<T>(...) (splined.hpp:340-367).  `inner` is the analytic scheme; the four tables are generated on
 * the host with the reference's Andrea algorithm (utol<=0 -> 1e-3).  The tables live in `storage`, which the caller
 * provides: at least cg_spline_storage_doubles() doubles, alive until cg_create has copied them.
 */
int64_t cg_spline_storage_doubles(void);
int cg_scheme_splined(const cg_scheme_desc *inner, double utol, double *storage, cg_scheme_desc *out);
/* Same, but with tables supplied by the caller (parity runs feed the oracle's tables to both sides) */
int cg_scheme_splined_tables(const cg_scheme_desc *inner, const int32_t n_knots[4], const double *const r2[4],
                             const double *const c[4], cg_scheme_desc *out);

/* host evaluation of S..S''' from a descriptor (same code path as the header classes): out[k*n+i] */
int cg_srf_host(const cg_scheme_desc *d, int64_t n, const double *q, double *out);

/* CG_MODE_MULTIPOLE_QUAD: CG_MODE_MULTIPOLE with the quadrupole tensors of cg_set_quadrupoles* (the reference's
 * quadrupole_potential core.hpp:320, the quad branches of multipole_field :477-481, multipole_multipole_energy :603-609 and
 * multipole_multipole_force :767-773, i.e. the ion-quadrupole terms; fp64 arithmetic in both precisions) */
typedef enum cg_mode { CG_MODE_ION = 0, CG_MODE_DIPOLE = 1, CG_MODE_MULTIPOLE = 2, CG_MODE_MULTIPOLE_QUAD = 3 } cg_mode;
enum cg_what { CG_ENERGY = 1, CG_FORCE = 2, CG_FIELD = 4, CG_POTENTIAL = 8 };
enum cg_precision { CG_FP64 = 0, CG_FP32 = 1 }; /* CG_FP32: fp32 pair arithmetic, fp64 positions/accumulators */

typedef struct cg_handle cg_handle;

int cg_device_count(void);
/* device < 0 -> current device */
int cg_create(const cg_scheme_desc *scheme, int device, cg_handle **out);
int cg_destroy(cg_handle *h);
const char *cg_last_error(const cg_handle *h); /* h may be NULL: last error of cg_create on this thread */
const char *cg_version(void);

/* run on a caller-owned CUDA stream (cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream); NULL -> own stream */
int cg_set_stream(cg_handle *h, void *cuda_stream);
int cg_set_precision(cg_handle *h, int precision);
/* Evaluate this handle's scheme on the system of `source` (same device, same cut-off): `source` uploads, bins and
 * sorts the particles once per configuration (its cg_set_system* + cg_eval*), every borrower reuses the sorted arrays
 * -- the way to compute several schemes (e.g. Wolf and Fennell) on one configuration.  Evaluate `source` first after
 * each cg_set_system*; the streams are ordered by events, the host calls must come from one thread.  NULL detaches. */
int cg_share_system(cg_handle *h, cg_handle *source);
/* evaluate only the i-particles of shard `rank` out of `nranks` (contiguous blocks of the spatially sorted order);
 * every rank still needs all N particles as j.  Default 0 of 1. */
int cg_set_shard(cg_handle *h, int rank, int nranks);

/* Positions (and charges / dipoles, any of which may be NULL = all zero) as SoA buffers.
 * box[3]: periodic box lengths when pbc != 0 (positions must lie in [0, box)); ignored when pbc == 0 (open). */
int cg_set_system(cg_handle *h, int64_t n, const double *x, const double *y, const double *z, const double *q,
                  const double *mux, const double *muy, const double *muz, const double *box, int pbc);
/* same with DEVICE pointers: zero-copy, the buffers must stay valid and unchanged until the next cg_eval returns */
int cg_set_system_device(cg_handle *h, int64_t n, const double *x, const double *y, const double *z, const double *q,
                         const double *mux, const double *muy, const double *muz, const double *box, int pbc);

/* Evaluate.  HOST output buffers, any may be NULL: force[3n], field[3n] (AoS xyz per particle), potential[n],
 * *energy (total U), *pairs (ordered in-cutoff pair count).  Blocks until the results are in the buffers. */
int cg_eval(cg_handle *h, int mode, int what, double *force, double *field, double *potential, double *energy,
            int64_t *pairs);
/* The same call in two halves, so that several handles (one per scheme, each on its own stream) overlap their
 * uploads, kernels and downloads: cg_eval_begin enqueues the evaluation and the copies into the HOST buffers (pinned
 * buffers make the copies truly asynchronous; pageable ones work, staged by the driver) and returns at once;
 * cg_eval_end waits, and hands out U and the pair count.  The buffers must stay valid until cg_eval_end returns. */
int cg_eval_begin(cg_handle *h, int mode, int what, double *force, double *field, double *potential);
int cg_eval_end(cg_handle *h, double *energy, int64_t *pairs);
/* Same with DEVICE output buffers; scalars[0] = U, scalars[1] = pair count (as double), on the device.
 * Asynchronous on the handle's stream.  From the second evaluation of a system shape (same n, box, shard) on, array sizes and
 * launch bounds come from the previous evaluation and no host round trip takes place.  Should the new configuration not fit
 * them -- or should a device-side particle count (cg_set_count_device) exceed the arrays given to cg_set_system_device --
 * the evaluation is INVALID and says so: scalars[0] = NaN, scalars[1] = -1, per-particle outputs partial.  A caller of this
 * entry point checks scalars[1] < 0 and evaluates again (the next call sizes exactly); cg_eval / cg_eval_end do that
 * themselves.
 * Two kernels serve the evaluation.  Open systems of up to 8192 particles (environment: CG_ALLPAIRS_MAX) whose arrays are
 * 16-byte aligned are evaluated straight from the SoA arrays by the all-pairs tile kernel (no sort; j-tiles staged through
 * shared memory by TMA bulk copies); everything else is binned into cells and evaluated by the cell-list kernel.  The results
 * agree to rounding, the pair counts exactly.  Handles that lend or borrow a system (cg_share_system), shards, device-side
 * counts, fp32 arithmetic and quadrupoles always take the cell-list kernel. */
int cg_eval_device(cg_handle *h, int mode, int what, double *force, double *field, double *potential,
                   double *scalars);

/* Several schemes in ONE pass over the pairs.  hs[0..n) evaluate the same system (hs[0] owns it or borrows it, the others
 * borrow the same one through cg_share_system) with the same mode and outputs; force / field / potential / scalars are arrays
 * of n DEVICE pointers (NULL arrays or entries allowed), one per scheme.  The binning, the prefilter, the walk over the hit
 * lists, the gathers, the distance and 1/r are done once per pair; schemes of the erfc family that share alpha also share
 * erfc(eta q) and exp(-eta^2 q^2) and differ by a cubic polynomial each (detail/srf.hpp: ErfcMultiSrf).  Supported: n = 2,
 * CG_MODE_ION, unscreened Wolf / Fennell / ZeroDipole / Zahn / Ewald / EwaldTruncated with one alpha and one cut-off;
 * CG_ERR_UNSUPPORTED otherwise (evaluate one by one then).  n = 1 is cg_eval_device.
 * The results agree with the one-by-one evaluation to rounding (1e-15 relative), the pair counts exactly.  Timings and
 * counters of the launch are reported for every handle. */
int cg_eval_multi_device(cg_handle *const *hs, int n, int mode, int what, double *const *force, double *const *field,
                         double *const *potential, double *const *scalars);
/* The same with HOST output buffers, in two halves like cg_eval_begin / cg_eval_end: force / field / potential are arrays of n
 * host pointers (NULL arrays or entries allowed; pinned buffers make the copies asynchronous), energy[n] and pairs[n] receive
 * every scheme's U and pair count.  Everything runs on hs[0]'s stream.  Two sets of handles used in turn -- set B's
 * cg_set_system + cg_eval_multi_begin issued before set A's cg_eval_multi_end -- overlap the upload of one configuration and
 * the download of another with the pair kernel of a third (bench.py's `e2e` figure). */
int cg_eval_multi_begin(cg_handle *const *hs, int n, int mode, int what, double *const *force, double *const *field,
                        double *const *potential);
int cg_eval_multi_end(cg_handle *const *hs, int n, double *energy, int64_t *pairs);
/* two scalars written by cg_eval_device (a DEVICE pointer) -> host; waits for the handle's stream */
int cg_read_scalars(cg_handle *h, const double *scalars_dev, double host2[2]);

/* Quadrupole tensors of the current system for CG_MODE_MULTIPOLE_QUAD: quad[n][9], row-major 3x3 per particle, not
 * necessarily symmetric or traceless (reference mat33 arguments).  Call after every cg_set_system* (which clears them);
 * NULL clears.  The _device form keeps the caller's device pointer (zero copy). */
int cg_set_quadrupoles(cg_handle *h, const double *quad);
int cg_set_quadrupoles_device(cg_handle *h, const double *quad);
/* O(N) terms of the current system, per particle as the reference's scheme functions give them
 * (core.hpp:124-145, 809-842): out[0] = sum_i self_energy({z_i^2, |mu_i|^2}), out[1] = neutralization_energy(charges,
 * box volume) (0 for an open system), out[2] = sum_i z_i; self_field[3n] (may be NULL) = self_field({0, mu_i}).
 * The _device forms take device pointers and are asynchronous on the handle's stream. */
int cg_self_terms(cg_handle *h, double out[3], double *self_field);
int cg_self_terms_device(cg_handle *h, double *out3, double *self_field);
/* Total moments of the current system, the two sums EwaldTruncated::surface_energy is built from (reference
 * ewald_truncated.hpp:91-104): out[0..2] = sum_i z_i r_i, out[3..5] = sum_i mu_i, accumulated separately in a fixed order.
 * The surface term is 2 pi / (2 eps_sur + 1) / V * (P.P + 2 P.D + D.D); the header class EwaldTruncated evaluates it. */
int cg_dipole_moments(cg_handle *h, double out[6]);
int cg_dipole_moments_device(cg_handle *h, double *out6);
/* torque[3n] = mu_i x field_i (reference core.hpp:794); field as written by cg_eval* with CG_FIELD */
int cg_dipole_torque(cg_handle *h, const double *field, double *torque);
int cg_dipole_torque_device(cg_handle *h, const double *field, double *torque);
/* per-functor probe: S..S''' evaluated by the device functor, q and out are DEVICE pointers, out[k*n+i] */
int cg_srf_device(cg_handle *h, int64_t n, const double *q, double *out);

/* timings of the last evaluation in milliseconds (CUDA events on the handle's stream):
 * ms[0] binning+sort+reorder, ms[1] pair kernel(s), ms[2] reductions/epilogue, ms[3] whole step */
int cg_last_timings(cg_handle *h, double ms[4]);
/* the same four numbers for each of the last k evaluations (k <= 64, oldest first, ms[4 k]): for callers that enqueue many
 * steps without waiting for any of them; waits for the newest */
int cg_timings_history(cg_handle *h, int k, double *ms);
/* counters of the last evaluation: [0] sorted entries incl. periodic images, [1] cells, [2] i-groups,
 * [3] pair-kernel launches, [4] total kernel launches, [5] j-split factor, [6] 1 if the evaluation took the sync-free path
 * (sizes from the previous evaluation, no host round trip), [7] work items of the pair kernel.  After an evaluation by the
 * all-pairs kernel: [0] = n, [1] = 0 (no cells), [2] = i-groups of 32 consecutive particles, [5] = j-slices */
int cg_last_counters(cg_handle *h, int64_t c[8]);

/* Exchange step over peer memory, one process per GPU: each rank keeps its owned particles in a block
 * obtained from cg_peer_alloc (layout: narrays arrays of count doubles, array-major), publishes the 64-byte handle
 * through its process group, and maps the other ranks' blocks with cg_peer_open.  cg_peer_gather then PULLS all
 * blocks over NVLink / NVSwitch with one kernel on the handle's stream into this rank's full SoA arrays
 * (full[a][offset_r .. offset_r + counts[r]) = block r, array a) -- the all-gather of the inputs without a collective
 * library call.  Ordering between ranks (blocks written before they are read, not rewritten while being read) is the
 * caller's: coulombgalore_b200.distributed.PeerExchange uses two alternating blocks and one small all-reduce per step.
 * blocks[rank] may be the local pointer.  device < 0: the current device. */
int cg_peer_alloc(int device, size_t bytes, void **ptr, void *handle64);
int cg_peer_free(void *ptr);
int cg_peer_open(int device, const void *handle64, void **ptr);
int cg_peer_close(void *ptr);
int cg_peer_gather(cg_handle *h, int nranks, const void *const *blocks, const int64_t *counts, int narrays,
                   double *const *full);

/* Halo exchange -- the same pull, but only of the particles a rank needs.  A shard evaluates a block of cell rows and
 * touches only the cell layers within one cut-off of it; cg_shard_window gives that z-interval [window[0], window[1]) for
 * any rank (box coordinates; periodic images count, i.e. test z - L, z, z + L).  Per step every owner sorts its particles
 * into one segment per destination rank inside its peer block (cg_halo_publish: three small kernels, stable order,
 * counts in the block header; block size cg_halo_block_bytes), and every rank pulls just its segments from all blocks
 * (cg_halo_pull: one kernel over NVLink; n_out[0] = particles received, n_out[1] = 1 if dst_cap was too small).  The
 * received subset is then evaluated as usual: cg_set_system_device(h, dst_cap, ...) + cg_set_count_device(h, n_out)
 * (the valid count stays on the device) with cg_set_grid_count(h, N_global) so that all ranks lay out the same cell
 * grid.  Arrays are SoA of doubles; carry the global particle index as one more array to map the outputs back.
 * zindex = which of the arrays is z.  counts[r] = particles owned by rank r (= its segment capacity). */
int cg_set_grid_count(cg_handle *h, int64_t n_global);
int cg_set_count_device(cg_handle *h, const int *n_dev);
int cg_shard_window(cg_handle *h, const double *box, int rank, int nranks, int64_t n_global, double window[2]);
size_t cg_halo_block_bytes(int nranks, int narrays, int64_t count);
/* Host-only probe of the pair kernel's work plan (csrc/kparams.cuh: tail_plan), for tests: with `groups` i-groups and
 * `resident_warps` warps on the machine, out[0] = groups evaluated whole, out[1] = row slices of each of the other groups
 * (1 = no tail split), out[2] = work items of the launch. */
int cg_tail_plan_probe(int groups, int resident_warps, int out[3]);

/* The cell layout and shard window as plain numbers, computed on the host without a handle or a GPU: out_i = cells per
 * dimension [3], ghost layers [3], the shard's primary cell rows [row0, row1) (row = z_cell * cells_y + y_cell), its first
 * cell layer and layer count in the ghost-extended grid; out_d = cell edge [3], z-window [2] (as cg_shard_window). */
int cg_layout_probe(double cutoff, const double *box, int pbc, int rank, int nranks, int64_t n_global, int out_i[10],
                    double out_d[5]);
int cg_halo_publish(cg_handle *h, int nranks, int narrays, int zindex, int64_t count, const double *const *owned,
                    const double *zlo, const double *zhi, double period, void *block);
int cg_halo_pull(cg_handle *h, int nranks, int rank, const void *const *blocks, const int64_t *counts, int narrays,
                 double *const *dst, int64_t dst_cap, int *n_out);
/* The same two calls with the ranks ordered ON THE DEVICE instead of by a collective of the caller: `step` = 0, 1, 2, ...
 * counts the exchange steps (the same on every rank), block = the rank's block number step % 2.  The owner's last publishing
 * CTA sets a `published` word in the block header (st.release.sys), cg_halo_pull_step spins on the owners' words
 * (ld.acquire.sys over NVLink) before it reads; the pullers acknowledge at the start of their next publish -- prev_blocks =
 * all ranks' blocks of step - 1, peer-mapped, needed from step 1 on -- and an owner waits for the acknowledgements of step - 2
 * before it rewrites that block.  A peer that does not show up within ~2 s sets *err_dev = 2 / n_out[1] = 2 instead of
 * hanging the GPU.  One process per GPU only: kernels of two ranks that wait on each other must run at the same time.
 * New blocks must be zero-filled in their first 256 bytes. */
int cg_halo_publish_step(cg_handle *h, int nranks, int narrays, int zindex, int64_t count, const double *const *owned,
                         const double *zlo, const double *zhi, double period, void *block, int64_t step, int rank,
                         void *const *prev_blocks, int *err_dev);
int cg_halo_pull_step(cg_handle *h, int nranks, int rank, const void *const *blocks, const int64_t *counts, int narrays,
                      double *const *dst, int64_t dst_cap, int *n_out, int64_t step);
/* Optional: a device-visible pinned HOST location (two ints) that every later cg_halo_pull[_step] of this handle also writes
 * n_out[0..1] to, straight from the kernel.  A caller that follows the received count on the host then needs no copy operation
 * in the stream (which would queue behind its own downloads on the copy engine).  NULL switches it off. */
int cg_halo_set_count_mirror(cg_handle *h, int *host_mapped);
/* One exchange step in one call: cg_halo_publish_step + cg_halo_pull_step + cg_set_system_device(h, n_bound, dst...) +
 * cg_set_count_device(h, n_out).  The arrays are x, y, z[, q][, mux, muy, muz] then anything else (e.g. the global index);
 * n_out = two ints on the device (count received; error word, also used by the publish); blocks = all ranks' blocks of this
 * step, prev_blocks = those of the previous step (NULL at step 0).  For small shards the host time between the launches of
 * an exchange step is what the GPUs wait for: this entry point issues all of them back to back. */
int cg_halo_exchange_step(cg_handle *h, int nranks, int rank, int narrays, int zindex, int64_t count, const double *const *owned,
                          const double *zlo, const double *zhi, const double *box, void *own_block, void *const *blocks,
                          void *const *prev_blocks, const int64_t *counts, double *const *dst, int64_t dst_cap, int64_t n_bound,
                          int *n_out, int64_t step, int has_q, int has_mu);
/* Per-particle results of an evaluation on a halo window come in received order (row r belongs to the particle whose global
 * index the exchanged index array holds).  This brings the rows of the rank's OWN particles into the order of the `owned`
 * arrays it published -- complete when ownership is spatial (every rank owns the particles of its block of cell rows);
 * owned particles that another rank evaluates are not touched (see cg_collect_evaluated for arbitrary ownership): dst[k][(ids[r] - lo) * width + c] = src[k][r * width + c] for
 * ids[r] in [lo, lo + count), r < min(*n_dev, bound).  ids, n_dev, src and dst are DEVICE pointers; narrays <= 8 output arrays
 * of `width` doubles per row (3 for force / field, 1 for the potential) are handled by one launch on the handle's stream. */
int cg_halo_collect_owned(cg_handle *h, const double *ids, const int *n_dev, int64_t bound, int64_t lo, int64_t count, int narrays,
                          int width, const double *const *src, double *const *dst);

/* The per-particle results a shard (or a whole system) evaluated, compacted for the way back: slot g * 32 + l of dst / ids_out
 * belongs to particle l of i-group g of the last cell-list evaluation of this handle.  ids_out[slot] = ids[row] (the global
 * index that travelled with the data; ids == NULL: the row in the arrays given to cg_set_system_device), -1 for unused
 * slots; dst[k][slot * width + c] = src[k][row * width + c].  Every particle of the shard's block of cell rows appears exactly
 * once, whoever owns it, so the ranks' outputs together hold every particle of the system once.  cap_rows = rows of dst /
 * ids_out (at least *rows_bound = 32 x the group bound of the last evaluation; cap_rows = 0 only queries it).  All pointers
 * but rows_bound are DEVICE pointers; one launch on the handle's stream. */
int cg_collect_evaluated(cg_handle *h, const double *ids, int narrays, int width, const double *const *src, double *const *dst,
                         double *ids_out, int64_t cap_rows, int64_t *rows_bound);

/* synthetic jittered lattice of SURVEY.md 8d generated on the device (n^3 particles) into DEVICE SoA buffers */
int cg_generate_lattice_device(cg_handle *h, int32_t n, double a, uint64_t seed, double *x, double *y, double *z,
                               double *q, double *mux, double *muy, double *muz);
/* FP64-pipe peak probe: a register-resident DFMA chain; returns measured TFLOP/s (2 flops per DFMA) */
int cg_fp64_peak_probe(cg_handle *h, double *tflops, double *ms);

/* ---- Reciprocal-space Ewald (reference ewald.hpp:44-94, 214-385: ReciprocalEwaldState + ReciprocalEwaldGaussian) ----
 * cg_recip_create takes the state fields (cutoff, alpha, surface_dielectric_constant, kappa, reciprocal_cutoff <= 24,
 * box_length) and generates the k-vectors and A_k as generateKVectors / setAks do.  The handle owns Q_mp (device):
 * cg_recip_update* = updateComplex (sign +1 add, -1 subtract, 0 = zero first), cg_recip_energy = reciprocal_energy,
 * cg_recip_force_field* = reciprocal_force(r_p, z_p, mu_p) and reciprocal_field(r_p) for every particle (either output
 * may be NULL), cg_recip_surface = surface_energy and surface_field.  *_device forms take device pointers and are
 * asynchronous on the handle's stream.  Multi-GPU: update with the rank's own particles, all-reduce the 2K doubles at
 * cg_recip_q_device (re, im interleaved), then evaluate the rank's own particles. */
typedef struct cg_recip cg_recip;
int cg_recip_create(double cutoff, double alpha, double eps_sur, double kappa, int reciprocal_cutoff, const double *box,
                    int device, cg_recip **out);
int cg_recip_destroy(cg_recip *r);
const char *cg_recip_last_error(const cg_recip *r);
int cg_recip_set_stream(cg_recip *r, void *cuda_stream);
int cg_recip_num_kvectors(const cg_recip *r);
int cg_recip_kvectors(const cg_recip *r, double *k3, double *aks); /* host copies: 3K and K doubles */
double *cg_recip_q_device(cg_recip *r);
int cg_recip_get_q(cg_recip *r, double *q2k);
int cg_recip_set_q(cg_recip *r, const double *q2k);
int cg_recip_update(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q,
                    const double *mux, const double *muy, const double *muz, int sign);
int cg_recip_update_device(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q,
                           const double *mux, const double *muy, const double *muz, int sign);
int cg_recip_energy(cg_recip *r, double *energy);
int cg_recip_energy_device(cg_recip *r, double *energy_dev);
int cg_recip_force_field(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q,
                         const double *mux, const double *muy, const double *muz, double *force, double *field);
int cg_recip_force_field_device(cg_recip *r, int64_t n, const double *x, const double *y, const double *z,
                                const double *q, const double *mux, const double *muy, const double *muz,
                                double *force, double *field);
int cg_recip_surface(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q,
                     const double *mux, const double *muy, const double *muz, double *energy, double *field3);
int cg_recip_surface_device(cg_recip *r, int64_t n, const double *x, const double *y, const double *z, const double *q,
                            const double *mux, const double *muy, const double *muz, double *dipole3_dev);

#ifdef __cplusplus
}
#endif
#endif
