/*
 * cgo_api.h -- C ABI of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * Two shared objects export exactly this interface:
 *   oracle/_ref/libcg_ref.so   the UNMODIFIED reference headers (/root/reference/include) compiled against
 *                              oracle/eigen_shim (kind "reference"); built by oracle/Makefile target `ref`.
 *   oracle/libcg_port.so       an independent plain-C++ restatement (oracle/port_schemes.hpp) that needs no
 *                              reference sources (kind "port"); pinned against libcg_ref.so and tests/golden.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load them.
 * The product (coulombgalore_b200/, include/) never does.
 */
#ifndef CGO_API_H
#define CGO_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* same order as the reference's `enum class Scheme` (core.hpp:56-70) */
enum cgo_scheme_id {
    CGO_PLAIN = 0,
    CGO_EWALD,
    CGO_EWALDT,
    CGO_REACTIONFIELD,
    CGO_WOLF,
    CGO_POISSON,
    CGO_QPOTENTIAL,
    CGO_FANOURGAKIS,
    CGO_ZERODIPOLE,
    CGO_ZAHN,
    CGO_FENNELL,
    CGO_QPOTENTIAL5, /* qPotentialFixedOrder<5> (qpotential.hpp:195) */
    CGO_SPLINE
};

/* Constructor arguments, exactly as the reference constructors take them. */
typedef struct cgo_params {
    int32_t scheme;      /* cgo_scheme_id of the (wrapped) analytic scheme */
    int32_t splined;     /* !=0: wrap in Splined().spline<T>(args...) (splined.hpp:364) */
    int32_t order;       /* qPotential order */
    int32_t C, D;        /* Poisson */
    int32_t shifted;     /* ReactionField */
    double cutoff;       /* ignored by Plain */
    double alpha;        /* Wolf/Fennell/Zahn/ZeroDipole/Ewald/EwaldT */
    double eps_sur;      /* Ewald/EwaldT surface dielectric constant (inf allowed) */
    double debye_length; /* Plain/Poisson/Ewald (inf = unscreened) */
    double eps_rf, eps_r;/* ReactionField */
    double spline_utol;  /* Splined::setTolerance; <=0 -> reference default 1e-3 */
} cgo_params;

enum cgo_mode { CGO_MODE_ION = 0, CGO_MODE_DIPOLE = 1, CGO_MODE_MULTIPOLE = 2 };
enum cgo_what { CGO_ENERGY = 1, CGO_FORCE = 2, CGO_FIELD = 4, CGO_POTENTIAL = 8 };

const char *cgo_kind(void); /* "reference" or "port" */

/* returns 0 on success; <0 when the reference constructor throws / parameters are rejected */
int cgo_create(const cgo_params *p, void **handle);
void cgo_destroy(void *handle);

/* info[0..9] = cutoff, debye_length, self_energy({1,0}), self_energy({0,1}), |self_field({0,ex})|,
 *              calc_dielectric(1.0), neutralization_energy({1},1.0), scheme id, T0-free spare, spare */
int cgo_info(void *handle, double *info);

/* S, S', S'', S''' on a grid: out[k*n + i] = k-th derivative at q[i] */
int cgo_srf(void *handle, int64_t n, const double *q, double *out);

/* Every per-pair function of the reference for one pair (quadrupoles row-major 3x3, may be NULL = zero).
 * out[ 0] ion_potential(zA,|r|)            out[ 1] dipole_potential(muA,r)      out[ 2] quadrupole_potential(quadA,r)
 * out[ 3.. 5] ion_field(zA,r)              out[ 6.. 8] dipole_field(muA,r)      out[ 9..11] quadrupole_field(quadA,r)
 * out[12..14] multipole_field(zA,muA,quadA,r)
 * out[15] ion_ion_energy(zA,zB,|r|)        out[16] ion_dipole_energy(zA,muB,r)  out[17] dipole_dipole_energy(muA,muB,r)
 * out[18] ion_quadrupole_energy(zB,quadA,r) out[19] multipole_multipole_energy(zA,zB,muA,muB,quadA,quadB,r)
 * out[20..22] ion_ion_force(zA,zB,r)       out[23..25] ion_dipole_force(zB,muA,r) out[26..28] dipole_dipole_force(muA,muB,r)
 * out[29..31] ion_quadrupole_force(zB,quadA,r) out[32..34] multipole_multipole_force(zA,zB,muA,muB,quadA,quadB,r)
 */
#define CGO_PAIR_OUT 35
int cgo_pair(void *handle, double zA, double zB, const double *muA, const double *muB, const double *quadA,
             const double *quadB, const double *r, double *out);

/* Splined only: number of knots of table k (0..3); copies knots/coefficients when pointers are non-NULL */
int cgo_spline_table(void *handle, int k, int32_t *n_knots, double *r2, double *c);

/* Synthetic jittered lattice of SURVEY.md 8d: n^3 sites, spacing a, splitmix64 random-access draws. */
void cgo_generate(int32_t n, double a, uint64_t seed, double *x, double *y, double *z, double *q, double *mux,
                  double *muy, double *muz);

/*
 * Batched reference loop (the loop the reference itself lacks; conventions of SURVEY.md 8b):
 * for i in [i_begin,i_end), sum over j != i of the reference's per-pair calls with d = r_i - (r_j + shift),
 * accumulating in long double.  box[3]; pbc!=0 -> periodic minimum image via cell shifts.
 * Outputs (any may be NULL), indexed by i - i_begin:
 *   force[3*n], field[3*n], potential[n]; energy_i[n] = 1/2 sum_j u_ij
 *   *_scale = sum_j of the term norms (the denominators of the parity tolerance)
 * totals[0] = sum_i energy_i, totals[1] = sum_i sum_j |u_ij|/2, totals[2] = in-cutoff ordered pair count
 */
int cgo_batched(void *handle, int64_t N, const double *x, const double *y, const double *z, const double *q,
                const double *mux, const double *muy, const double *muz, const double *box, int32_t pbc,
                int32_t mode, int32_t what, int64_t i_begin, int64_t i_end, int32_t nthreads, int32_t want_scales,
                double *force, double *field, double *potential, double *energy_i, double *force_scale,
                double *field_scale, double *potential_scale, double *totals);

/* cgo_batched in multipole mode with quadrupole tensors quad[N][9] (row-major, not necessarily symmetric or traceless):
 * potential += quadrupole_potential(quad_j, d); field = multipole_field(z_j, mu_j, quad_j, d); energy / force =
 * multipole_multipole_energy / _force(z_i, z_j, mu_i, mu_j, quad_i, quad_j, -d)  (core.hpp:320, 457, 581, 737). */
int cgo_batched_quad(void *handle, int64_t N, const double *x, const double *y, const double *z, const double *q,
                     const double *mux, const double *muy, const double *muz, const double *quad, const double *box,
                     int32_t pbc, int32_t what, int64_t i_begin, int64_t i_end, int32_t nthreads, int32_t want_scales,
                     double *force, double *field, double *potential, double *energy_i, double *force_scale,
                     double *field_scale, double *potential_scale, double *totals);

/*
 * O(N) terms of a configuration, per particle through the scheme's own functions (reference core.hpp:794, 809-842):
 * scalars[0] = sum_i self_energy({z_i^2, |mu_i|^2}), scalars[1] = sum of the absolute terms (tolerance scale),
 * scalars[2] = neutralization_energy(charges, volume) (0 when volume <= 0), scalars[3] = sum_i z_i;
 * self_field[3N] = self_field({0, mu_i}); torque[3N] = dipole_torque(mu_i, field_i).  Output pointers may be NULL.
 */
int cgo_self_terms(void *handle, int64_t N, const double *q, const double *mux, const double *muy, const double *muz,
                   double volume, const double *field, double *scalars, double *self_field, double *torque);

/*
 * Reciprocal-space Ewald (reference ewald.hpp:44-94, 214-385: ReciprocalEwaldState / ReciprocalEwaldGaussian).
 * create: the state fields + generateKVectors(box).  update: updateComplex with sign = +1 (add), -1 (subtract) or
 * 0 (zero Q first, then add).  q: Q_mp as 2K doubles (re, im interleaved); kvectors: 3K doubles + K Aks.
 * force_field: reciprocal_force(r_p, z_p, mu_p) and reciprocal_field(r_p) per particle (outputs may be NULL);
 * scales (may be NULL): per particle sum_k |term| of the force (tolerance denominator).
 * surface: surface_energy and surface_field (3 doubles) of the particle set.
 */
int cgo_recip_create(double cutoff, double alpha, double eps_sur, double kappa, int32_t reciprocal_cutoff, const double *box,
                     void **handle);
void cgo_recip_destroy(void *handle);
int32_t cgo_recip_numk(void *handle);
int cgo_recip_kvectors(void *handle, double *k3, double *aks);
int cgo_recip_update(void *handle, int64_t N, const double *x, const double *y, const double *z, const double *q,
                     const double *mux, const double *muy, const double *muz, int32_t sign);
int cgo_recip_q(void *handle, double *q2k);
int cgo_recip_energy(void *handle, double *energy);
int cgo_recip_force_field(void *handle, int64_t N, const double *x, const double *y, const double *z, const double *q,
                          const double *mux, const double *muy, const double *muz, double *force, double *field,
                          double *force_scale);
int cgo_recip_surface(void *handle, int64_t N, const double *x, const double *y, const double *z, const double *q,
                      const double *mux, const double *muy, const double *muz, double *energy, double *field3);

/* EwaldTruncated::surface_energy(positions, charges, dipoles, volume) (reference ewald_truncated.hpp:91-104) of a scheme
 * constructed as EwaldTruncated(cutoff, alpha, eps_sur); q / mu pointers may be NULL (all zero). */
int cgo_ewaldt_surface(double cutoff, double alpha, double eps_sur, int64_t N, const double *x, const double *y, const double *z,
                       const double *q, const double *mux, const double *muy, const double *muz, double volume, double *energy);

/* out[0..3] = qPochhammerSymbol, ...Derivative, ...SecondDerivative, ...ThirdDerivative (q, l, P) (qpotential.hpp:50-192) */
int cgo_qpochhammer(double q, int32_t l, int32_t P, double *out4);

#ifdef __cplusplus
}
#endif
#endif
