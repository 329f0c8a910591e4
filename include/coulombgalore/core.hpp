// coulombgalore/core.hpp -- drop-in class surface of CoulombGalore for the B200 engine (host side).
//
// Same namespace, type names, public data members and per-pair method signatures as the reference
// (reference include/coulombgalore/core.hpp:46-244 `SchemeBase`, :251-843 `EnergyImplementation<T>`), so code written
// against the reference compiles against these headers unchanged.  The arithmetic is NOT the reference's: every method
// is a thin call into coulombgalore/detail/pair.hpp, the single host/device formula set that the sm_100a kernels of
// libcg_b200.so are instantiated from.  New relative to the reference: `SchemeBase::descriptor()`, the POD parameter
// block the batched GPU path consumes (include/cg_b200.h, coulombgalore/batched.hpp).
//
// vec3 / mat33 are Eigen::Vector3d / Eigen::Matrix3d when Eigen is on the include path (as in the reference,
// core.hpp:49-50) and small built-in types otherwise, so the headers also work stand-alone.
#pragma once
#include <array>
#include <cmath>
#include <functional>
#include <limits>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "../cg_b200.h"
#include "detail/descriptor.hpp"
#include "detail/pair.hpp"

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && !defined(COULOMBGALORE_NO_EIGEN)
#include <Eigen/Core>
#include <Eigen/Dense>
#define COULOMBGALORE_HAVE_EIGEN 1
#endif
#endif

namespace CoulombGalore {

#ifdef COULOMBGALORE_HAVE_EIGEN
using vec3 = Eigen::Vector3d;
using mat33 = Eigen::Matrix3d;
#else
/** Minimal 3-vector used when Eigen is not available */
struct vec3 {
    double v[3];
    vec3() : v{0.0, 0.0, 0.0} {}
    vec3(double x, double y, double z) : v{x, y, z} {}
    static vec3 Zero() { return vec3(); }
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double &operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double dot(const vec3 &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    vec3 cross(const vec3 &o) const {
        return vec3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
    vec3 &operator+=(const vec3 &o) {
        for (int i = 0; i < 3; i++) v[i] += o.v[i];
        return *this;
    }
    vec3 &operator-=(const vec3 &o) {
        for (int i = 0; i < 3; i++) v[i] -= o.v[i];
        return *this;
    }
    vec3 &operator*=(double s) {
        for (int i = 0; i < 3; i++) v[i] *= s;
        return *this;
    }
};
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline vec3 operator-(const vec3 &a) { return vec3(-a[0], -a[1], -a[2]); }
inline vec3 operator*(const vec3 &a, double s) { return vec3(a[0] * s, a[1] * s, a[2] * s); }
inline vec3 operator*(double s, const vec3 &a) { return a * s; }
inline vec3 operator/(const vec3 &a, double s) { return vec3(a[0] / s, a[1] / s, a[2] / s); }
/** Minimal 3x3 matrix (row, column) used when Eigen is not available */
struct mat33 {
    double m[3][3];
    mat33() : m{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}} {}
    static mat33 Zero() { return mat33(); }
    double &operator()(int i, int j) { return m[i][j]; }
    double operator()(int i, int j) const { return m[i][j]; }
    double trace() const { return m[0][0] + m[1][1] + m[2][2]; }
};
#endif

constexpr double infinity = std::numeric_limits<double>::infinity(); //!< Numerical infinity

/** All schemes; numbering as in the reference (core.hpp:56-70) and `cg_scheme_id` */
enum class Scheme { plain, ewald, ewaldt, reactionfield, wolf, poisson, qpotential, fanourgakis, zerodipole, zahn, fennell, qpotential5, spline };

inline double powi(double x, int n) { return detail::powi(x, n); }                                            //!< x^n, n integer
inline constexpr unsigned int factorial(unsigned int n) { return detail::factorial(n); }                      //!< n! (32 bit)
inline constexpr unsigned int binomial(signed int n, signed int k) { return detail::binomial(n, k); }         //!< n over k (32 bit)

/**
 * @brief Dynamic interface of all truncation schemes (reference core.hpp:110-244)
 */
class SchemeBase {
  protected:
    cg_scheme_desc desc_; //!< every constant of the scheme; what the GPU path consumes

    void adopt(const cg_scheme_desc &d) { //!< install a descriptor and mirror it in the public members
        desc_ = d;
        scheme = static_cast<Scheme>(d.scheme);
        cutoff = d.cutoff;
        debye_length = d.debye_length;
        const double ic = d.inverse_cutoff;
        const std::array<double, 2> se = {d.self_energy_prefactor[0], d.self_energy_prefactor[1]};
        const std::array<double, 2> sf = {d.self_field_prefactor[0], d.self_field_prefactor[1]};
        selfEnergyFunctor = [ic, se](const std::array<double, 2> &m2) {
            double e = 0.0;
            for (int i = 0; i < 2; i++) e += se[i] * m2[i] * powi(ic, 2 * i + 1);
            return e;
        };
        selfFieldFunctor = [ic, sf](const std::array<vec3, 2> &m) {
            vec3 f(0.0, 0.0, 0.0);
            for (int i = 0; i < 2; i++) f += sf[i] * powi(ic, 2 * i + 1) * m[i];
            return f;
        };
    }

  public:
    std::string doi;     //!< DOI for original citation
    std::string name;    //!< Descriptive name
    Scheme scheme;       //!< Truncation scheme
    double cutoff;       //!< Cut-off distance
    double debye_length; //!< Debye-length

    std::function<double(const std::array<double, 2> &)> selfEnergyFunctor = nullptr; //!< self-energy functor
    std::function<vec3(const std::array<vec3, 2> &)> selfFieldFunctor = nullptr;      //!< self-field functor

    SchemeBase() : scheme(Scheme::plain), cutoff(0.0), debye_length(infinity) { desc_ = detail::desc_base(CG_PLAIN, 1.0, infinity); }
    virtual ~SchemeBase() = default;

    /** POD parameter block for the batched GPU path (cg_create); valid as long as this object lives */
    const cg_scheme_desc &descriptor() const { return desc_; }

    virtual double neutralization_energy(const std::vector<double> &, double) const { return 0.0; }
    double calc_dielectric(double M2V) const { return (M2V * desc_.T0 + 2.0 * M2V + 1.0) / (M2V * desc_.T0 - M2V + 1.0); }

    virtual double short_range_function(double q) const = 0;
    virtual double short_range_function_derivative(double q) const = 0;
    virtual double short_range_function_second_derivative(double q) const = 0;
    virtual double short_range_function_third_derivative(double q) const = 0;

    virtual double ion_potential(double, double) const = 0;
    virtual double dipole_potential(const vec3 &, const vec3 &) const = 0;
    virtual double quadrupole_potential(const mat33 &, const vec3 &) const = 0;

    virtual double ion_ion_energy(double, double, double) const = 0;
    virtual double ion_dipole_energy(double, const vec3 &, const vec3 &) const = 0;
    virtual double dipole_dipole_energy(const vec3 &, const vec3 &, const vec3 &) const = 0;
    virtual double ion_quadrupole_energy(double, const mat33 &, const vec3 &) const = 0;
    virtual double multipole_multipole_energy(double, double, const vec3 &, const vec3 &, const mat33 &, const mat33 &, const vec3 &) const = 0;

    virtual vec3 ion_field(double, const vec3 &) const = 0;
    virtual vec3 dipole_field(const vec3 &, const vec3 &) const = 0;
    virtual vec3 quadrupole_field(const mat33 &, const vec3 &) const = 0;
    virtual vec3 multipole_field(double, const vec3 &, const mat33 &, const vec3 &) const = 0;

    virtual vec3 ion_ion_force(double, double, const vec3 &) const = 0;
    virtual vec3 ion_dipole_force(double, const vec3 &, const vec3 &) const = 0;
    virtual vec3 dipole_dipole_force(const vec3 &, const vec3 &, const vec3 &) const = 0;
    virtual vec3 ion_quadrupole_force(double, const mat33 &, const vec3 &) const = 0;
    virtual vec3 multipole_multipole_force(double, double, const vec3 &, const vec3 &, const mat33 &, const mat33 &, const vec3 &) const = 0;
};

/**
 * @brief Implements every pair function on top of a derived class' short_range_function*(q) (CRTP), like the
 * reference's class of the same name (core.hpp:251-843).  `debyehuckel` keeps the reference's meaning: the kappa
 * terms of the core formulas are evaluated (kappa = 1/debye_length; infinity gives exactly the unscreened values).
 */
template <class T, bool debyehuckel = true> class EnergyImplementation : public SchemeBase {
    // adapts the derived class' four S-functions to the functor interface of detail/pair.hpp
    struct Functor {
        const T *self;
        template <int ND, class R = double> void eval(double q, double *s, const double *) const {
            s[0] = self->short_range_function(q);
            if (ND >= 1) s[1] = self->short_range_function_derivative(q);
            if (ND >= 2) s[2] = self->short_range_function_second_derivative(q);
            if (ND >= 3) s[3] = self->short_range_function_third_derivative(q);
        }
        static constexpr bool kHasAngcor = false;
        static constexpr int kTable = detail::kTableNone;
    };
    Functor functor() const { return Functor{static_cast<const T *>(this)}; }
    detail::CoreParams core() const { return detail::CoreParams{desc_.inverse_cutoff, desc_.cutoff_squared, desc_.inverse_debye_length}; }
    static detail::D3 d3(const vec3 &v) { return detail::D3{v[0], v[1], v[2]}; }

    template <int MODE, int WHAT>
    detail::Accum pair(const vec3 &d, double zi, double zj, const vec3 &mui, const vec3 &muj) const {
        detail::Accum acc;
        const detail::D3 dd = d3(d);
        const double r2 = dd.x * dd.x + dd.y * dd.y + dd.z * dd.z;
        if (r2 < desc_.cutoff_squared) // strict gate on r^2, as in the reference (e.g. core.hpp:354)
            detail::pair_accumulate<Functor, debyehuckel, MODE, WHAT>(functor(), core(), nullptr, dd, r2, zi, zj, d3(mui), d3(muj), acc);
        return acc;
    }
    // radial factors for the quadrupole terms, which only exist on the host (reference core.hpp:320-339, 412-439)
    struct Quad {
        bool in;
        double r1, r2, ek, a, b, c; // a = totcor, b = unicor, c = r3corr
    };
    Quad quad_terms(const vec3 &r) const {
        Quad Q;
        Q.r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        Q.in = Q.r2 < desc_.cutoff_squared;
        if (!Q.in) return Q;
        const detail::Radial R = detail::radial_terms<Functor, debyehuckel, 3>(functor(), core(), Q.r2, nullptr);
        Q.r1 = R.r1;
        Q.ek = R.ek;
        Q.a = R.totcor;
        Q.b = R.unicor;
        Q.c = R.r3corr;
        return Q;
    }
    static double quadratic(const mat33 &m, const vec3 &r) { // r^T M r
        double s = 0.0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) s += r[i] * m(i, j) * r[j];
        return s;
    }
    static vec3 sym_apply(const mat33 &m, const vec3 &r) { // (M + M^T) r
        vec3 o(0.0, 0.0, 0.0);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) o[i] += (m(i, j) + m(j, i)) * r[j];
        return o;
    }

  public:
    EnergyImplementation() = default;

    // ---- potentials ----
    double ion_potential(double z, double r) const override { // gate on the scalar distance (core.hpp:269)
        if (!(r < cutoff)) return 0.0;
        double s[4];
        functor().template eval<0>(r * desc_.inverse_cutoff, s, nullptr);
        return debyehuckel ? z / r * s[0] * std::exp(-desc_.inverse_debye_length * r) : z / r * s[0];
    }
    double dipole_potential(const vec3 &mu, const vec3 &r) const override {
        return pair<detail::kModeDipole, detail::kWhatPotential>(r, 0.0, 0.0, vec3(0, 0, 0), mu).phi;
    }
    double quadrupole_potential(const mat33 &quad, const vec3 &r) const override {
        const Quad Q = quad_terms(r);
        if (!Q.in) return 0.0;
        const double tr = quad.trace();
        return 0.5 * ((3.0 / Q.r2 * quadratic(quad, r) - tr) * Q.a + tr * Q.b) / Q.r2 / Q.r1 * Q.ek;
    }
    // ---- fields ----
    vec3 ion_field(double z, const vec3 &r) const override {
        const detail::Accum a = pair<detail::kModeIon, detail::kWhatField>(r, 0.0, z, vec3(0, 0, 0), vec3(0, 0, 0));
        return vec3(a.Ex, a.Ey, a.Ez);
    }
    vec3 dipole_field(const vec3 &mu, const vec3 &r) const override {
        const detail::Accum a = pair<detail::kModeDipole, detail::kWhatField>(r, 0.0, 0.0, vec3(0, 0, 0), mu);
        return vec3(a.Ex, a.Ey, a.Ez);
    }
    vec3 quadrupole_field(const mat33 &quad, const vec3 &r) const override {
        const Quad Q = quad_terms(r);
        if (!Q.in) return vec3(0, 0, 0);
        const vec3 rh = r / Q.r1;
        const double qf = quadratic(quad, r) / Q.r2, r4 = Q.r2 * Q.r2;
        const vec3 fieldD = 3.0 * ((5.0 * qf - quad.trace()) * rh - sym_apply(quad, rh)) / r4 * Q.a;
        const vec3 fieldI = qf * rh / r4 * Q.c;
        return 0.5 * (fieldD + fieldI) * Q.ek;
    }
    vec3 multipole_field(double z, const vec3 &mu, const mat33 &quad, const vec3 &r) const override {
        const detail::Accum a = pair<detail::kModeMultipole, detail::kWhatField>(r, 0.0, z, vec3(0, 0, 0), mu);
        return vec3(a.Ex, a.Ey, a.Ez) + quadrupole_field(quad, r);
    }
    // ---- energies ----
    double ion_ion_energy(double zA, double zB, double r) const override { return zB * ion_potential(zA, r); }
    double ion_dipole_energy(double charge, const vec3 &mu, const vec3 &r) const override { return charge * dipole_potential(mu, -r); }
    double dipole_dipole_energy(const vec3 &muA, const vec3 &muB, const vec3 &r) const override { return -muA.dot(dipole_field(muB, r)); }
    double ion_quadrupole_energy(double charge, const mat33 &quad, const vec3 &r) const override {
        return charge * quadrupole_potential(quad, -r);
    }
    double multipole_multipole_energy(double zA, double zB, const vec3 &muA, const vec3 &muB, const mat33 &quadA, const mat33 &quadB,
                                      const vec3 &r) const override {
        // kernel convention: A = i, B = j, d = r_i - r_j = -r
        const double u = pair<detail::kModeMultipole, detail::kWhatEnergy>(-r, zA, zB, muA, muB).U;
        return u + zA * quadrupole_potential(quadB, r) + zB * quadrupole_potential(quadA, r); // ion-quadrupole terms (core.hpp:604-609)
    }
    // ---- forces ----
    vec3 ion_ion_force(double zA, double zB, const vec3 &r) const override { return zB * ion_field(zA, r); }
    vec3 ion_dipole_force(double charge, const vec3 &mu, const vec3 &r) const override { return charge * dipole_field(mu, r); }
    vec3 dipole_dipole_force(const vec3 &muA, const vec3 &muB, const vec3 &r) const override {
        const detail::Accum a = pair<detail::kModeDipole, detail::kWhatForce>(-r, 0.0, 0.0, muA, muB);
        return vec3(a.Fx, a.Fy, a.Fz);
    }
    vec3 ion_quadrupole_force(double charge, const mat33 &quad, const vec3 &r) const override { return charge * quadrupole_field(quad, r); }
    vec3 multipole_multipole_force(double zA, double zB, const vec3 &muA, const vec3 &muB, const mat33 &quadA, const mat33 &quadB,
                                   const vec3 &r) const override {
        const detail::Accum a = pair<detail::kModeMultipole, detail::kWhatForce>(-r, zA, zB, muA, muB);
        // ion-quadrupole terms exactly as the reference signs them (core.hpp:767-773)
        return vec3(a.Fx, a.Fy, a.Fz) - zA * quadrupole_field(quadB, r) + zB * quadrupole_field(quadA, r);
    }
    /** torque on a dipole in a field (core.hpp:794) */
    vec3 dipole_torque(const vec3 &mu, const vec3 &E) const { return mu.cross(E); }
    /** self-energy from squared moments {z^2, |mu|^2} (core.hpp:809-812) */
    double self_energy(const std::array<double, 2> &squared_moments) const { return selfEnergyFunctor(squared_moments); }
    /** self-field from moments {z, mu} (core.hpp:827-830) */
    vec3 self_field(const std::array<vec3, 2> &moments) const { return selfFieldFunctor(moments); }
    /** compensating term for non-neutral systems (core.hpp:839-842) */
    double neutralization_energy(const std::vector<double> &charges, double volume) const override {
        const double s = std::accumulate(charges.begin(), charges.end(), 0.0);
        return desc_.chi / 2.0 / volume * s * s;
    }
};

namespace detail {
/** Schemes whose S(q) comes from one of the descriptor-constructed functors of detail/srf.hpp */
template <class Derived, class Srf> class FunctorScheme : public EnergyImplementation<Derived> {
  protected:
    Srf srf_;
    void install(const cg_scheme_desc &d, const char *scheme_name, const char *scheme_doi) {
        this->adopt(d);
        srf_ = Srf(d);
        this->name = scheme_name;
        this->doi = scheme_doi;
    }
    template <int K> double s(double q) const {
        double v[4];
        srf_.template eval<K>(q, v, nullptr);
        return v[K];
    }

  public:
    double short_range_function(double q) const override { return s<0>(q); }
    double short_range_function_derivative(double q) const override { return s<1>(q); }
    double short_range_function_second_derivative(double q) const override { return s<2>(q); }
    double short_range_function_third_derivative(double q) const override { return s<3>(q); }
};
inline void require_ok(int rc, const char *what) {
    if (rc != CG_OK) throw std::runtime_error(what);
}
} // namespace detail

} // namespace CoulombGalore
