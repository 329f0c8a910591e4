// oracle/harness.hpp -- TEST INFRASTRUCTURE ONLY (see cgo_api.h).
//
// The loop code shared by both oracle builds.  It is templated over the scheme class so that the *same*
// harness drives (a) the unmodified reference classes (ref_main.cpp) and (b) the independent restatement
// (port_main.cpp): any numerical difference between the two builds can then only come from the scheme
// arithmetic, never from the loop.  The reference has no batched loop of its own; the
// conventions used here are the "literal-to-reference" ones of SURVEY.md 8b:
//   potential/field at i use d = r_i - r_j ("distance vector from the source", core.hpp:285,344,370),
//   ion force on i   = ion_ion_force(z_j, z_i, d)            (force on B, core.hpp:621-630)
//   dipole force on i= dipole_dipole_force(mu_i, mu_j, -d)   (force on A for r = r_B - r_A, core.hpp:654-677)
//   multipole energy/force are called with A=i, B=j, r = r_j - r_i (core.hpp:574,730), quadrupoles zero
//   U = 1/2 sum_i sum_{j!=i} u_ij
#pragma once
#include "cgo_api.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace cgo {

// ---------------------------------------------------------------------------------------------------
// splitmix64 random-access generator of SURVEY.md 8d
// ---------------------------------------------------------------------------------------------------
inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
inline uint64_t draw(uint64_t seed, uint64_t k) { return mix64(seed + (k + 1) * 0x9E3779B97F4A7C15ULL); }
inline double unit(uint64_t d) { return (double)(d >> 11) * (1.0 / 9007199254740992.0); }

inline void generate(int32_t n, double a, uint64_t seed, double *x, double *y, double *z, double *q, double *mux,
                     double *muy, double *muz) {
    const int64_t N = (int64_t)n * n * n;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < N; p++) {
        const int64_t k = p % n, j = (p / n) % n, i = p / ((int64_t)n * n);
        const uint64_t b = 7 * (uint64_t)p;
        const double ux = unit(draw(seed, b + 0)), uy = unit(draw(seed, b + 1)), uz = unit(draw(seed, b + 2));
        if (x) x[p] = ((double)i + 0.5 + 0.5 * (ux - 0.5)) * a;
        if (y) y[p] = ((double)j + 0.5 + 0.5 * (uy - 0.5)) * a;
        if (z) z[p] = ((double)k + 0.5 + 0.5 * (uz - 0.5)) * a;
        if (q) q[p] = (draw(seed, b + 3) & 1ULL) ? 1.0 : -1.0;
        double m0 = unit(draw(seed, b + 4)) - 0.5, m1 = unit(draw(seed, b + 5)) - 0.5, m2 = unit(draw(seed, b + 6)) - 0.5;
        const double nrm = std::sqrt(m0 * m0 + m1 * m1 + m2 * m2);
        if (mux) mux[p] = m0 / nrm;
        if (muy) muy[p] = m1 / nrm;
        if (muz) muz[p] = m2 / nrm;
    }
}

// ---------------------------------------------------------------------------------------------------
// type-erased interface held behind the C handle
// ---------------------------------------------------------------------------------------------------
struct Batch {
    int64_t N;
    const double *x, *y, *z, *q, *mux, *muy, *muz;
    const double *quad; // [N][9] row-major quadrupole tensors (multipole mode only), or null = zero quadrupoles
    double box[3];
    int pbc, mode, what;
    int64_t i_begin, i_end;
    int nthreads, want_scales;
    double *force, *field, *potential, *energy_i, *force_scale, *field_scale, *potential_scale, *totals;
};

struct Iface {
    virtual ~Iface() {}
    virtual int info(double *out) = 0;
    virtual int srf(int64_t n, const double *q, double *out) = 0;
    virtual int pair(double zA, double zB, const double *muA, const double *muB, const double *quadA,
                     const double *quadB, const double *r, double *out) = 0;
    virtual int spline_table(int k, int32_t *n_knots, double *r2, double *c) = 0;
    virtual int batched(const Batch &b) = 0;
    virtual int self_terms(int64_t N, const double *q, const double *mx, const double *my, const double *mz, double volume,
                           const double *field, double *scalars, double *self_field, double *torque) = 0;
};

// simple CPU cell list (cells >= cutoff, 27-cell stencil, periodic shifts or open walls)
struct CellList {
    int nc[3];
    double origin[3], cs[3], box[3];
    std::vector<int64_t> start; // ncell+1
    std::vector<int64_t> idx;   // particle indices sorted by cell
    int64_t ncell() const { return (int64_t)nc[0] * nc[1] * nc[2]; }
    int coord(double v, int d) const {
        int c = (int)std::floor((v - origin[d]) / cs[d]);
        if (c < 0) c = 0;
        if (c >= nc[d]) c = nc[d] - 1;
        return c;
    }
    int64_t cell_of(double px, double py, double pz) const {
        return ((int64_t)coord(pz, 2) * nc[1] + coord(py, 1)) * nc[0] + coord(px, 0);
    }
    void build(int64_t N, const double *x, const double *y, const double *z, const double *boxin, int pbc,
               double cutoff) {
        double lo[3] = {0, 0, 0}, ext[3] = {boxin[0], boxin[1], boxin[2]};
        if (!pbc) { // open walls: bounding box of the particles
            double hi[3];
            for (int d = 0; d < 3; d++) {
                lo[d] = std::numeric_limits<double>::infinity();
                hi[d] = -lo[d];
            }
            for (int64_t p = 0; p < N; p++) {
                const double v[3] = {x[p], y[p], z[p]};
                for (int d = 0; d < 3; d++) {
                    lo[d] = std::min(lo[d], v[d]);
                    hi[d] = std::max(hi[d], v[d]);
                }
            }
            for (int d = 0; d < 3; d++)
                ext[d] = std::max(hi[d] - lo[d], 1e-300);
        }
        for (int d = 0; d < 3; d++) {
            origin[d] = lo[d];
            box[d] = ext[d];
            double f = ext[d] / cutoff;
            int n = (f >= 1.0 && f < 1e6) ? (int)std::floor(f) : 1;
            if (!(cutoff < 1e100)) n = 1;
            nc[d] = std::max(1, std::min(n, 512));
            cs[d] = ext[d] / nc[d];
        }
        start.assign((size_t)ncell() + 1, 0);
        std::vector<int64_t> cell((size_t)N);
        for (int64_t p = 0; p < N; p++) {
            cell[(size_t)p] = cell_of(x[p], y[p], z[p]);
            start[(size_t)cell[(size_t)p] + 1]++;
        }
        for (int64_t c = 0; c < ncell(); c++)
            start[(size_t)c + 1] += start[(size_t)c];
        idx.resize((size_t)N);
        std::vector<int64_t> fill(start.begin(), start.end() - 1);
        for (int64_t p = 0; p < N; p++)
            idx[(size_t)fill[(size_t)cell[(size_t)p]]++] = p;
    }
};

template <class V> inline double norm3(const V &v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// T: scheme class with the reference's per-pair API; VEC/MAT: its 3-vector / 3x3 types.
template <class T, class VEC, class MAT> struct Impl : Iface {
    T pot;
    template <class... A> explicit Impl(A &&...a) : pot(std::forward<A>(a)...) {}

    int info(double *o) override {
        o[0] = pot.cutoff;
        o[1] = pot.debye_length;
        o[2] = pot.self_energy({1.0, 0.0});
        o[3] = pot.self_energy({0.0, 1.0});
        VEC zero(0, 0, 0), ex(1, 0, 0);
        o[4] = pot.self_field({zero, ex})[0];
        o[5] = pot.calc_dielectric(1.0);
        o[6] = pot.neutralization_energy({1.0}, 1.0);
        o[7] = (double)(int)pot.scheme;
        o[8] = o[9] = 0.0;
        return 0;
    }
    int srf(int64_t n, const double *q, double *out) override {
        for (int64_t i = 0; i < n; i++) {
            out[0 * n + i] = pot.short_range_function(q[i]);
            out[1 * n + i] = pot.short_range_function_derivative(q[i]);
            out[2 * n + i] = pot.short_range_function_second_derivative(q[i]);
            out[3 * n + i] = pot.short_range_function_third_derivative(q[i]);
        }
        return 0;
    }
    static MAT mat(const double *m) {
        MAT r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                r(i, j) = m ? m[3 * i + j] : 0.0;
        return r;
    }
    static void put(double *o, const VEC &v) {
        o[0] = v[0];
        o[1] = v[1];
        o[2] = v[2];
    }
    int pair(double zA, double zB, const double *a, const double *b, const double *qa, const double *qb,
             const double *rr, double *o) override {
        const VEC muA(a[0], a[1], a[2]), muB(b[0], b[1], b[2]), r(rr[0], rr[1], rr[2]);
        const MAT quadA = mat(qa), quadB = mat(qb);
        const double r1 = r.norm();
        o[0] = pot.ion_potential(zA, r1);
        o[1] = pot.dipole_potential(muA, r);
        o[2] = pot.quadrupole_potential(quadA, r);
        put(o + 3, pot.ion_field(zA, r));
        put(o + 6, pot.dipole_field(muA, r));
        put(o + 9, pot.quadrupole_field(quadA, r));
        put(o + 12, pot.multipole_field(zA, muA, quadA, r));
        o[15] = pot.ion_ion_energy(zA, zB, r1);
        o[16] = pot.ion_dipole_energy(zA, muB, r);
        o[17] = pot.dipole_dipole_energy(muA, muB, r);
        o[18] = pot.ion_quadrupole_energy(zB, quadA, r);
        o[19] = pot.multipole_multipole_energy(zA, zB, muA, muB, quadA, quadB, r);
        put(o + 20, pot.ion_ion_force(zA, zB, r));
        put(o + 23, pot.ion_dipole_force(zB, muA, r));
        put(o + 26, pot.dipole_dipole_force(muA, muB, r));
        put(o + 29, pot.ion_quadrupole_force(zB, quadA, r));
        put(o + 32, pot.multipole_multipole_force(zA, zB, muA, muB, quadA, quadB, r));
        return 0;
    }
    int spline_table(int, int32_t *, double *, double *) override { return -1; }

    // O(N) terms per particle through the scheme's own functions (reference core.hpp:794, 809-842); long double sums
    int self_terms(int64_t N, const double *q, const double *mx, const double *my, const double *mz, double volume,
                   const double *field, double *scalars, double *self_field, double *torque) override {
        long double us = 0.0L, usa = 0.0L, zs = 0.0L;
        std::vector<double> charges((size_t)N, 0.0);
        const VEC zero(0, 0, 0);
        for (int64_t i = 0; i < N; i++) {
            const double z = q ? q[i] : 0.0;
            const VEC mu = mx ? VEC(mx[i], my[i], mz[i]) : zero;
            const double m2 = mu[0] * mu[0] + mu[1] * mu[1] + mu[2] * mu[2];
            charges[(size_t)i] = z;
            zs += z;
            us += (long double)pot.self_energy({z * z, m2});
            usa += (long double)(std::fabs(pot.self_energy({z * z, 0.0})) + std::fabs(pot.self_energy({0.0, m2})));
            if (self_field) {
                const VEC sf = pot.self_field({zero, mu});
                for (int d = 0; d < 3; d++) self_field[3 * i + d] = sf[d];
            }
            if (torque && field) {
                const VEC t = pot.dipole_torque(mu, VEC(field[3 * i], field[3 * i + 1], field[3 * i + 2]));
                for (int d = 0; d < 3; d++) torque[3 * i + d] = t[d];
            }
        }
        scalars[0] = (double)us;
        scalars[1] = (double)usa;
        scalars[2] = volume > 0.0 ? pot.neutralization_energy(charges, volume) : 0.0;
        scalars[3] = (double)zs;
        return 0;
    }

    int batched(const Batch &b) override {
        CellList cl;
        cl.build(b.N, b.x, b.y, b.z, b.box, b.pbc, pot.cutoff);
        const MAT zeroquad = mat(nullptr);
        const int64_t n = b.i_end - b.i_begin;
        long double U = 0, Uabs = 0;
        long long npairs = 0;
        const double rc2 = pot.cutoff * pot.cutoff;
#ifdef _OPENMP
        const int nthr = b.nthreads > 0 ? b.nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthr) reduction(+ : U, Uabs, npairs)
#endif
        for (int64_t ii = 0; ii < n; ii++) {
            const int64_t i = b.i_begin + ii;
            const VEC ri(b.x[i], b.y[i], b.z[i]);
            const double zi = b.q ? b.q[i] : 0.0;
            const VEC mui = b.mux ? VEC(b.mux[i], b.muy[i], b.muz[i]) : VEC(0, 0, 0);
            long double F[3] = {0, 0, 0}, E[3] = {0, 0, 0}, phi = 0, u = 0, uabs = 0, Fs = 0, Es = 0, ps = 0;
            const int c[3] = {cl.coord(b.x[i], 0), cl.coord(b.y[i], 1), cl.coord(b.z[i], 2)};
            for (int oz = -1; oz <= 1; oz++)
                for (int oy = -1; oy <= 1; oy++)
                    for (int ox = -1; ox <= 1; ox++) {
                        int nb[3] = {c[0] + ox, c[1] + oy, c[2] + oz};
                        double sh[3] = {0, 0, 0};
                        bool skip = false;
                        for (int d = 0; d < 3; d++) {
                            if (nb[d] < 0) {
                                if (!b.pbc) skip = true;
                                nb[d] += cl.nc[d];
                                sh[d] = -cl.box[d];
                            } else if (nb[d] >= cl.nc[d]) {
                                if (!b.pbc) skip = true;
                                nb[d] -= cl.nc[d];
                                sh[d] = cl.box[d];
                            }
                        }
                        if (skip) continue;
                        const int64_t cell = ((int64_t)nb[2] * cl.nc[1] + nb[1]) * cl.nc[0] + nb[0];
                        for (int64_t k = cl.start[(size_t)cell]; k < cl.start[(size_t)cell + 1]; k++) {
                            const int64_t j = cl.idx[(size_t)k];
                            if (j == i) continue;
                            // image of j, then the distance vector from the source j to i
                            const VEC d(ri[0] - (b.x[j] + sh[0]), ri[1] - (b.y[j] + sh[1]), ri[2] - (b.z[j] + sh[2]));
                            const double r2 = d.squaredNorm();
                            if (!(r2 < rc2)) continue; // same strict gate as core.hpp:354 etc.
                            npairs++;
                            const double zj = b.q ? b.q[j] : 0.0;
                            const VEC muj = b.mux ? VEC(b.mux[j], b.muy[j], b.muz[j]) : VEC(0, 0, 0);
                            const VEC md = -d; // r_j - r_i
                            const MAT quadi = b.quad ? mat(b.quad + 9 * i) : zeroquad, quadj = b.quad ? mat(b.quad + 9 * j) : zeroquad;
                            if (b.mode == CGO_MODE_ION) {
                                const double r1 = std::sqrt(r2);
                                if (b.what & CGO_POTENTIAL) {
                                    const double t = pot.ion_potential(zj, r1);
                                    phi += t;
                                    ps += std::fabs(t);
                                }
                                if (b.what & CGO_FIELD) {
                                    const VEC t = pot.ion_field(zj, d);
                                    for (int a = 0; a < 3; a++) E[a] += t[a];
                                    if (b.want_scales) Es += norm3(t);
                                }
                                if (b.what & CGO_FORCE) {
                                    const VEC t = pot.ion_ion_force(zj, zi, d);
                                    for (int a = 0; a < 3; a++) F[a] += t[a];
                                    if (b.want_scales) Fs += norm3(t);
                                }
                                if (b.what & CGO_ENERGY) {
                                    const double t = pot.ion_ion_energy(zj, zi, r1);
                                    u += t;
                                    uabs += std::fabs(t);
                                }
                            } else if (b.mode == CGO_MODE_DIPOLE) {
                                if (b.what & CGO_POTENTIAL) {
                                    const double t = pot.dipole_potential(muj, d);
                                    phi += t;
                                    ps += std::fabs(t);
                                }
                                if (b.what & CGO_FIELD) {
                                    const VEC t = pot.dipole_field(muj, d);
                                    for (int a = 0; a < 3; a++) E[a] += t[a];
                                    if (b.want_scales) Es += norm3(t);
                                }
                                if (b.what & CGO_FORCE) {
                                    const VEC t = pot.dipole_dipole_force(mui, muj, md);
                                    for (int a = 0; a < 3; a++) F[a] += t[a];
                                    if (b.want_scales) Fs += norm3(t);
                                }
                                if (b.what & CGO_ENERGY) {
                                    const double t = pot.dipole_dipole_energy(mui, muj, md);
                                    u += t;
                                    uabs += std::fabs(t);
                                }
                            } else {
                                if (b.what & CGO_POTENTIAL) {
                                    const double t = pot.ion_potential(zj, std::sqrt(r2)) + pot.dipole_potential(muj, d) +
                                                     (b.quad ? pot.quadrupole_potential(quadj, d) : 0.0);
                                    phi += t;
                                    ps += std::fabs(t);
                                }
                                if (b.what & CGO_FIELD) {
                                    const VEC t = pot.multipole_field(zj, muj, quadj, d);
                                    for (int a = 0; a < 3; a++) E[a] += t[a];
                                    if (b.want_scales) Es += norm3(t);
                                }
                                if (b.what & CGO_FORCE) {
                                    const VEC t = pot.multipole_multipole_force(zi, zj, mui, muj, quadi, quadj, md);
                                    for (int a = 0; a < 3; a++) F[a] += t[a];
                                    if (b.want_scales) Fs += norm3(t);
                                }
                                if (b.what & CGO_ENERGY) {
                                    const double t = pot.multipole_multipole_energy(zi, zj, mui, muj, quadi, quadj, md);
                                    u += t;
                                    uabs += std::fabs(t);
                                }
                            }
                        }
                    }
            if (b.force) for (int a = 0; a < 3; a++) b.force[3 * ii + a] = (double)F[a];
            if (b.field) for (int a = 0; a < 3; a++) b.field[3 * ii + a] = (double)E[a];
            if (b.potential) b.potential[ii] = (double)phi;
            if (b.energy_i) b.energy_i[ii] = (double)(0.5L * u);
            if (b.force_scale) b.force_scale[ii] = (double)Fs;
            if (b.field_scale) b.field_scale[ii] = (double)Es;
            if (b.potential_scale) b.potential_scale[ii] = (double)ps;
            U += 0.5L * u;
            Uabs += 0.5L * uabs;
        }
        if (b.totals) {
            b.totals[0] = (double)U;
            b.totals[1] = (double)Uabs;
            b.totals[2] = (double)npairs;
        }
        return 0;
    }
};


// reciprocal-space Ewald behind the C ABI: each build supplies cgo::make_recip
struct RecipIface {
    virtual ~RecipIface() {}
    virtual int numk() = 0;
    virtual int kvectors(double *k3, double *aks) = 0;
    virtual int update(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx, const double *my,
                       const double *mz, int sign) = 0;
    virtual int getq(double *q2k) = 0;
    virtual int energy(double *e) = 0;
    virtual int force_field(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx,
                            const double *my, const double *mz, double *force, double *field, double *fscale) = 0;
    virtual int surface(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx,
                        const double *my, const double *mz, double *energy, double *field3) = 0;
};
RecipIface *make_recip(double cutoff, double alpha, double eps_sur, double kappa, int kc, const double *box);

} // namespace cgo

// ---------------------------------------------------------------------------------------------------
// C ABI glue; each build supplies  cgo::Iface *cgo_make(const cgo_params &)  and  cgo_kind().
// ---------------------------------------------------------------------------------------------------
namespace cgo {
Iface *make(const cgo_params &p); // throws on invalid parameters
}

#define CGO_DEFINE_C_ABI(KIND)                                                                                    \
    extern "C" {                                                                                                   \
    const char *cgo_kind(void) { return KIND; }                                                                    \
    int cgo_create(const cgo_params *p, void **h) {                                                                \
        try {                                                                                                      \
            *h = cgo::make(*p);                                                                                    \
            return *h ? 0 : -2;                                                                                    \
        } catch (...) {                                                                                            \
            *h = nullptr;                                                                                          \
            return -1;                                                                                             \
        }                                                                                                          \
    }                                                                                                              \
    void cgo_destroy(void *h) { delete static_cast<cgo::Iface *>(h); }                                             \
    int cgo_info(void *h, double *o) { return static_cast<cgo::Iface *>(h)->info(o); }                             \
    int cgo_srf(void *h, int64_t n, const double *q, double *o) { return static_cast<cgo::Iface *>(h)->srf(n, q, o); } \
    int cgo_pair(void *h, double zA, double zB, const double *a, const double *b, const double *qa,               \
                 const double *qb, const double *r, double *o) {                                                   \
        return static_cast<cgo::Iface *>(h)->pair(zA, zB, a, b, qa, qb, r, o);                                     \
    }                                                                                                              \
    int cgo_spline_table(void *h, int k, int32_t *n, double *r2, double *c) {                                      \
        return static_cast<cgo::Iface *>(h)->spline_table(k, n, r2, c);                                            \
    }                                                                                                              \
    void cgo_generate(int32_t n, double a, uint64_t seed, double *x, double *y, double *z, double *q,             \
                      double *mx, double *my, double *mz) {                                                        \
        cgo::generate(n, a, seed, x, y, z, q, mx, my, mz);                                                         \
    }                                                                                                              \
    int cgo_recip_create(double cutoff, double alpha, double eps_sur, double kappa, int32_t kc, const double *box,  \
                         void **h) {                                                                               \
        try {                                                                                                      \
            *h = cgo::make_recip(cutoff, alpha, eps_sur, kappa, kc, box);                                          \
            return *h ? 0 : -2;                                                                                    \
        } catch (...) {                                                                                            \
            *h = nullptr;                                                                                          \
            return -1;                                                                                             \
        }                                                                                                          \
    }                                                                                                              \
    void cgo_recip_destroy(void *h) { delete static_cast<cgo::RecipIface *>(h); }                                  \
    int32_t cgo_recip_numk(void *h) { return static_cast<cgo::RecipIface *>(h)->numk(); }                          \
    int cgo_recip_kvectors(void *h, double *k3, double *a) { return static_cast<cgo::RecipIface *>(h)->kvectors(k3, a); } \
    int cgo_recip_update(void *h, int64_t N, const double *x, const double *y, const double *z, const double *q,  \
                         const double *mx, const double *my, const double *mz, int32_t sign) {                     \
        return static_cast<cgo::RecipIface *>(h)->update(N, x, y, z, q, mx, my, mz, sign);                         \
    }                                                                                                              \
    int cgo_recip_q(void *h, double *q2k) { return static_cast<cgo::RecipIface *>(h)->getq(q2k); }                 \
    int cgo_recip_energy(void *h, double *e) { return static_cast<cgo::RecipIface *>(h)->energy(e); }              \
    int cgo_recip_force_field(void *h, int64_t N, const double *x, const double *y, const double *z,               \
                              const double *q, const double *mx, const double *my, const double *mz,               \
                              double *force, double *field, double *fs) {                                          \
        return static_cast<cgo::RecipIface *>(h)->force_field(N, x, y, z, q, mx, my, mz, force, field, fs);        \
    }                                                                                                              \
    int cgo_recip_surface(void *h, int64_t N, const double *x, const double *y, const double *z, const double *q, \
                          const double *mx, const double *my, const double *mz, double *e, double *f3) {           \
        return static_cast<cgo::RecipIface *>(h)->surface(N, x, y, z, q, mx, my, mz, e, f3);                       \
    }                                                                                                              \
    int cgo_self_terms(void *h, int64_t N, const double *q, const double *mx, const double *my, const double *mz,  \
                       double volume, const double *field, double *scalars, double *self_field, double *torque) {  \
        return static_cast<cgo::Iface *>(h)->self_terms(N, q, mx, my, mz, volume, field, scalars, self_field, torque); \
    }                                                                                                              \
    int cgo_batched(void *h, int64_t N, const double *x, const double *y, const double *z, const double *q,       \
                    const double *mx, const double *my, const double *mz, const double *box, int32_t pbc,          \
                    int32_t mode, int32_t what, int64_t i0, int64_t i1, int32_t nthr, int32_t scales,              \
                    double *force, double *field, double *pot, double *en, double *fs, double *es, double *ps,     \
                    double *totals) {                                                                              \
        cgo::Batch b{N, x, y, z, q, mx, my, mz, nullptr, {box[0], box[1], box[2]}, pbc, mode, what, i0, i1, nthr,   \
                     scales, force, field, pot, en, fs, es, ps, totals};                                           \
        return static_cast<cgo::Iface *>(h)->batched(b);                                                           \
    }                                                                                                              \
    int cgo_batched_quad(void *h, int64_t N, const double *x, const double *y, const double *z, const double *q,  \
                         const double *mx, const double *my, const double *mz, const double *quad,                 \
                         const double *box, int32_t pbc, int32_t what, int64_t i0, int64_t i1, int32_t nthr,       \
                         int32_t scales, double *force, double *field, double *pot, double *en, double *fs,        \
                         double *es, double *ps, double *totals) {                                                 \
        cgo::Batch b{N, x, y, z, q, mx, my, mz, quad, {box[0], box[1], box[2]}, pbc, CGO_MODE_MULTIPOLE, what, i0,  \
                     i1, nthr, scales, force, field, pot, en, fs, es, ps, totals};                                 \
        return static_cast<cgo::Iface *>(h)->batched(b);                                                           \
    }                                                                                                              \
    }
