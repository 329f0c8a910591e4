// coulombgalore/detail/descriptor.hpp -- host-side construction of cg_scheme_desc.
//
// One function per reference constructor.  Each computes the scheme constants with the expressions
// the reference constructor uses (cited per function), validates like the reference does (Poisson), and fills the
// O(N)-term prefactors (self energy / self field / T0 / chi) that ride along in the descriptor.
#pragma once
#include "../../cg_b200.h"
#include "srf.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <vector>

namespace CoulombGalore {
namespace detail {

inline cg_scheme_desc desc_base(int functor, double cutoff, double debye_length) {
    cg_scheme_desc d;
    std::memset(&d, 0, sizeof(d));
    d.functor = functor;
    d.scheme = functor;
    d.cutoff = cutoff;
    d.inverse_cutoff = 1.0 / cutoff; // core.hpp:157-159
    d.cutoff_squared = cutoff * cutoff;
    d.debye_length = debye_length;
    d.inverse_debye_length = 1.0 / debye_length;
    d.F0 = 1.0;
    return d;
}

// S..S''' on the host from a descriptor: the single dispatch point used by the scheme classes, the C ABI
// (cg_srf_host) and the spline generator
inline void srf_host(const cg_scheme_desc &d, double q, double *s) {
    switch (d.functor) {
    case CG_PLAIN: PlainSrf(d).eval<3>(q, s); break;
    case CG_REACTIONFIELD: ReactionFieldSrf(d).eval<3>(q, s); break;
    case CG_POISSON: PoissonSrf(d).eval<3>(q, s); break;
    case CG_QPOTENTIAL:
    case CG_QPOTENTIAL5: QPotentialSrf<0>(d).eval<3>(q, s); break;
    case CG_FANOURGAKIS: FanourgakisSrf(d).eval<3>(q, s); break;
    case CG_WOLF: ErfcFamilySrf<kWolf>(d).eval<3>(q, s); break;
    case CG_FENNELL: ErfcFamilySrf<kFennell>(d).eval<3>(q, s); break;
    case CG_ZERODIPOLE: ErfcFamilySrf<kZeroDipole>(d).eval<3>(q, s); break;
    case CG_ZAHN: ErfcFamilySrf<kZahn>(d).eval<3>(q, s); break;
    case CG_EWALDT: ErfcFamilySrf<kEwaldT>(d).eval<3>(q, s); break;
    case CG_EWALD:
        if (d.zeta != 0.0) EwaldScreenedSrf(d).eval<3>(q, s);
        else ErfcFamilySrf<kEwaldPlain>(d).eval<3>(q, s);
        break;
    case CG_SPLINE: {
        SplineSrf f;
        std::vector<double> buf((size_t)SplineSrf::packed_size(d));
        f.pack_from(d, buf.data());
        f.eval<3>(q, s);
        break;
    }
    default: s[0] = s[1] = s[2] = s[3] = std::numeric_limits<double>::quiet_NaN();
    }
}

inline void finish_T0(cg_scheme_desc &d) { // T0 = S'(1) - S(1) + S(0), e.g. plain.hpp:44
    double s1[4], s0[4];
    srf_host(d, 1.0, s1);
    srf_host(d, 0.0, s0);
    d.T0 = s1[1] - s1[0] + s0[0];
}

constexpr double kPi = 3.14159265358979323846;
inline double pi_sqrt_ref() { return 2.0 * std::sqrt(std::atan(1.0)); } // wolf.hpp:36 etc.

inline int make_plain(double debye_length, cg_scheme_desc &d) { // plain.hpp:37-46
    d = desc_base(CG_PLAIN, std::sqrt(std::numeric_limits<double>::max()), debye_length);
    finish_T0(d);
    d.chi = -2.0 * std::acos(-1.0) * d.cutoff_squared;
    return CG_OK;
}

inline int make_reactionfield(double cutoff, double epsRF, double epsr, bool shifted, cg_scheme_desc &d) { // reactionfield.hpp:47-58
    d = desc_base(CG_REACTIONFIELD, cutoff, kInfinity);
    d.shifted = shifted ? 1 : 0;
    d.eps_rf = epsRF;
    d.eps_r = epsr;
    d.rf_a = (epsRF - epsr) / (2.0 * epsRF + epsr);
    d.rf_b = 3.0 * epsRF / (2.0 * epsRF + epsr) * double(shifted);
    d.self_energy_prefactor[0] = -3.0 * epsRF * double(shifted) / (4.0 * epsRF + 2.0 * epsr);
    d.self_energy_prefactor[1] = -(2.0 * epsRF - 2.0 * epsr) / (2.0 * (2.0 * epsRF + epsr));
    finish_T0(d);
    const double pi = 4.0 * std::atan(1.0);
    d.chi = -6.0 * cutoff * cutoff * pi * ((-10.0 * double(shifted) / 3.0 + 4.0) * epsRF + epsr) / ((5.0 * (2.0 * epsRF + epsr)));
    return CG_OK;
}

inline int make_poisson(double cutoff, int C, int D, double debye_length, cg_scheme_desc &d) { // poisson.hpp:79-117
    if (C < 1) return CG_ERR_INVALID;                       // "`C` must be larger than zero"
    if ((D < -1) && (D != -C)) return CG_ERR_INVALID;       // "If `D` is less than negative one, ..."
    if ((D == 0) && (C != 1)) return CG_ERR_INVALID;        // "If `D` is zero, then `C` has to equal one"
    if (D == -1 && C > 1) return CG_ERR_INVALID;            // the reference recurses without bound here
    if (C > CG_MAX_POISSON_C || (D != -C && C + D > 12)) return CG_ERR_INVALID; // 32-bit factorial overflow
    d = desc_base(CG_POISSON, cutoff, debye_length);
    d.C = C;
    d.D = D;
    double a1 = -double(C + D) / double(C);
    if (!std::isinf(debye_length)) {
        d.reduced_kappa = cutoff / debye_length;
        if (std::fabs(d.reduced_kappa) > 1e-6) {
            d.yukawa_map = 1;
            d.yukawa_denom = 1.0 / (1.0 - std::exp(2.0 * d.reduced_kappa));
            a1 *= -2.0 * d.reduced_kappa * d.yukawa_denom;
        } else {
            d.reduced_kappa = 0.0;
        }
    }
    if (D != -C) {
        d.binomCDC = double(binomial(C + D, C) * (unsigned int)D);
        // (C, D) = (1, 0) returns before the polynomial in the reference (poisson.hpp:128-130); its weights would
        // need binomial(-1, 0), an unbounded recursion
        if (!(D == 0 && C == 1))
            for (int c = 0; c < C; c++) d.poisson_w[c] = double(binomial(D - 1 + c, c)) * double(C - c) / double(C);
    }
    d.self_energy_prefactor[0] = 0.5 * a1;
    d.self_energy_prefactor[1] = (C == 2) ? -double(D) * (double(D * D) + 3.0 * double(D) + 2.0) / 12.0 : 0.0;
    finish_T0(d);
    d.chi = -2.0 * std::acos(-1.0) * cutoff * cutoff * (1.0 + double(C)) * (2.0 + double(C)) / (3.0 * double(D + 1 + C) * double(D + 2 + C));
    return CG_OK;
}

// fixed_order_chi: qPotentialFixedOrder<order> sets the order-5 value of chi whatever its order (qpotential.hpp:212),
// qPotential(cutoff, order) leaves chi = 0 (qpotential.hpp:260)
inline int make_qpotential(double cutoff, int order, bool fixed_order_chi, cg_scheme_desc &d) { // qpotential.hpp:205-213, 253-261
    if (order < 0) return CG_ERR_INVALID;
    d = desc_base(CG_QPOTENTIAL, cutoff, kInfinity);
    d.order = order;
    d.self_energy_prefactor[0] = -0.5;
    d.self_energy_prefactor[1] = -0.5;
    finish_T0(d);
    d.chi = fixed_order_chi ? -86459.0 * std::acos(-1.0) * cutoff * cutoff / 235620.0 : 0.0;
    return CG_OK;
}

inline int make_fanourgakis(double cutoff, cg_scheme_desc &d) { // fanourgakis.hpp:39-47
    d = desc_base(CG_FANOURGAKIS, cutoff, kInfinity);
    d.self_energy_prefactor[0] = -0.875;
    finish_T0(d);
    d.chi = -5.0 * std::acos(-1.0) * cutoff * cutoff / 18.0;
    return CG_OK;
}

inline void erfc_constants(cg_scheme_desc &d, double alpha) {
    d.eta = alpha * d.cutoff;
    d.erfc_eta = std::erfc(d.eta);
    d.exp_eta2 = std::exp(-(d.eta * d.eta));
}

inline int make_wolf(double cutoff, double alpha, cg_scheme_desc &d) { // wolf.hpp:44-58
    d = desc_base(CG_WOLF, cutoff, kInfinity);
    erfc_constants(d, alpha);
    const double eta = d.eta, eta2 = eta * eta, ps = pi_sqrt_ref(), pi = 4.0 * std::atan(1.0);
    d.self_energy_prefactor[0] = -eta / ps - std::erfc(eta) / 2.0;
    d.self_energy_prefactor[1] = -powi(eta, 3) * 2.0 / 3.0 / ps;
    finish_T0(d);
    d.chi = 2.0 * std::sqrt(pi) * cutoff * cutoff *
            (3.0 * std::exp(-eta2) * eta - std::sqrt(pi) * (std::erfc(eta) * eta2 + 3.0 * std::erf(eta) * 0.5)) / (3.0 * eta2);
    return CG_OK;
}

inline int make_fennell(double cutoff, double alpha, cg_scheme_desc &d) { // fennell.hpp:44-57
    d = desc_base(CG_FENNELL, cutoff, kInfinity);
    erfc_constants(d, alpha);
    const double eta = d.eta, eta2 = eta * eta, ps = pi_sqrt_ref(), pi = 4.0 * std::atan(1.0);
    d.self_energy_prefactor[0] = -eta * (1.0 + std::exp(-eta2)) / ps - std::erfc(eta);
    finish_T0(d);
    d.chi = (2.0 * ((eta2 + 3.0) * eta * std::exp(-eta2) + 0.5 * (std::erf(eta) * eta2 - eta2 - 3.0 * std::erf(eta)) * std::sqrt(pi))) *
            std::sqrt(pi) * cutoff * cutoff / (3.0 * eta2);
    return CG_OK;
}

inline int make_zerodipole(double cutoff, double alpha, cg_scheme_desc &d) { // zerodipole.hpp:44-59
    d = desc_base(CG_ZERODIPOLE, cutoff, kInfinity);
    erfc_constants(d, alpha);
    const double eta = d.eta, eta2 = eta * eta, ps = pi_sqrt_ref(), pi = 4.0 * std::atan(1.0);
    d.self_energy_prefactor[0] = -eta * (1.0 + 0.5 * std::exp(-eta2)) / ps - 0.75 * std::erfc(eta);
    d.self_energy_prefactor[1] = -eta * (2.0 * eta2 * (1.0 / 3.0) + std::exp(-eta2)) / ps - 0.5 * std::erfc(eta);
    finish_T0(d);
    d.chi = cutoff * cutoff *
            ((6.0 * eta2 - 15.0) * std::erf(eta) * pi - 6.0 * pi * eta2 + (8.0 * eta2 + 30.0) * eta * std::exp(-eta2) * std::sqrt(pi)) /
            (15.0 * eta2);
    return CG_OK;
}

inline int make_zahn(double cutoff, double alpha, cg_scheme_desc &d) { // zahn.hpp:45-58
    d = desc_base(CG_ZAHN, cutoff, kInfinity);
    erfc_constants(d, alpha);
    const double eta = d.eta, eta2 = eta * eta, ps = pi_sqrt_ref(), pi = 4.0 * std::atan(1.0);
    d.self_energy_prefactor[0] = -eta * (1.0 - std::exp(-eta2)) / ps + 0.5 * std::erfc(eta);
    finish_T0(d);
    d.chi = -(2.0 * (eta * (eta2 - 3.0) * std::exp(-eta2) - 0.5 * std::sqrt(pi) * ((7.0 - 3.0 / eta2) * std::erf(eta) - 7.0) * eta2)) *
            cutoff * cutoff * std::sqrt(pi) / (3.0 * eta2);
    return CG_OK;
}

inline int make_ewald(double cutoff, double alpha, double eps_sur, double debye_length, cg_scheme_desc &d) { // ewald.hpp:129-173
    // the base class is constructed without the Debye length (ewald.hpp:131): kappa of the core formulas stays 0 and
    // the public debye_length stays infinite; screening lives only in S(q) through zeta
    d = desc_base(CG_EWALD, cutoff, kInfinity);
    erfc_constants(d, alpha);
    const double eta = d.eta, eta2 = eta * eta, eta3 = eta2 * eta, ps = pi_sqrt_ref(), pi = 4.0 * std::atan(1.0);
    const double eps = (eps_sur < 1.0) ? kInfinity : eps_sur;
    const double Q = 1.0 - std::erfc(eta) - 2.0 * eta / ps * std::exp(-eta2);
    d.T0 = std::isinf(eps) ? Q : Q - 1.0 + 2.0 * (eps - 1.0) / (2.0 * eps + 1.0);
    const double zeta = cutoff / debye_length, zeta2 = zeta * zeta, zeta3 = zeta2 * zeta;
    d.zeta = zeta;
    if (zeta < 1e-6) {
        d.chi = -pi * d.cutoff_squared * (1.0 - std::erfc(eta) * (1.0 - 2.0 * eta2) - 2.0 * eta * std::exp(-eta2) / ps) / eta2;
    } else {
        d.chi = 4.0 *
                (0.5 * (1.0 - zeta) * std::erfc(eta + zeta / (2.0 * eta)) * std::exp(zeta) + std::erf(eta) * std::exp(-zeta2 / (4.0 * eta2)) +
                 0.5 * (1.0 + zeta) * std::erfc(eta - zeta / (2.0 * eta)) * std::exp(-zeta) - 1.0) *
                pi * d.cutoff_squared / zeta2;
    }
    d.self_energy_prefactor[0] = -eta / ps * (std::exp(-zeta2 / 4.0 / eta2) - ps * zeta / (2.0 * eta) * std::erfc(zeta / (2.0 * eta)));
    d.self_energy_prefactor[1] = -eta3 / ps * 2.0 / 3.0 *
                                 (ps * zeta3 / 4.0 / eta3 * std::erfc(zeta / (2.0 * eta)) + (1.0 - zeta2 / 2.0 / eta2) * std::exp(-zeta2 / 4.0 / eta2));
    d.self_field_prefactor[1] = 4.0 / 3.0 * eta3 / ps;
    return CG_OK;
}

inline int make_ewaldt(double cutoff, double alpha, double eps_sur, double /*debye_length: ignored, ewald_truncated.hpp:48*/,
                       cg_scheme_desc &d) { // ewald_truncated.hpp:47-69
    d = desc_base(CG_EWALDT, cutoff, kInfinity);
    erfc_constants(d, alpha);
    const double eta = d.eta, eta2 = eta * eta, eta3 = eta2 * eta, ps = pi_sqrt_ref(), pi = 4.0 * std::atan(1.0);
    const double eps = (eps_sur < 1.0) ? kInfinity : eps_sur;
    d.F0 = 1.0 - std::erfc(eta) - 2.0 * eta / ps * std::exp(-eta2);
    d.T0 = std::isinf(eps) ? 1.0 : 2.0 * (eps - 1.0) / (2.0 * eps + 1.0);
    d.chi = -(1.0 - 4.0 * eta3 * std::exp(-eta2) / (3.0 * ps * d.F0)) * d.cutoff_squared * pi / eta2;
    d.self_energy_prefactor[0] = -eta / ps * (1.0 - std::exp(-eta2)) / d.F0;
    d.self_energy_prefactor[1] = -eta3 * 2.0 / 3.0 / (std::erf(eta) * ps - 2.0 * eta * std::exp(-eta2));
    return CG_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Tabulate::Andrea::generate(f, 0, 1) as Splined drives it (splined.hpp:195-298):
// walk from q = 1 downwards, intervals equidistant in sqrt(q), quintic Hermite from value / first / second central
// differences (numdr = 1e-4) at both ends, accepted when 11 sample points agree with f to `utol`; the trial interval
// shrinks by 0.9 up to 100 times.  Returns ascending knots and 6 coefficients per interval.
struct SplineTable {
    std::vector<double> r2, c;
};

inline int andrea_generate(const std::function<double(double)> &f, double utol, SplineTable &td) {
    const double numdr = 0.0001, drfrac = 0.9;
    const int mngrid = 1200, ndr = 100;
    auto f1 = [&](double x) { return (f(x + numdr * 0.5) - f(x - numdr * 0.5)) / (numdr); };
    auto f2 = [&](double x) { return (f1(x + numdr * 0.5) - f1(x - numdr * 0.5)) / (numdr); };
    const double rmin = 0.0, rmax = 1.0;
    td.r2.clear();
    td.c.clear();
    double rupp = rmax;
    td.r2.push_back(rmax * rmax);
    int i = 0;
    for (; i < mngrid; i++) {
        double rlow = rupp, zlow = 0.0, coeff[6] = {0, 0, 0, 0, 0, 0};
        double dr = rupp - rmin;
        int j = 0;
        for (; j < ndr; j++) {
            const double zupp = rupp * rupp;
            rlow = std::max(rupp - dr, rmin);
            zlow = rlow * rlow;
            const double u0l = f(zlow), u1l = f1(zlow), u2l = f2(zlow), u0u = f(zupp), u1u = f1(zupp), u2u = f2(zupp);
            if (std::fabs(u0l) < 1e-9 && std::fabs(u1l) < 1e-9) { // "zero potential and force return no coefficients"
                for (double &v : coeff) v = 0.0;
            } else {
                const double dz1 = zupp - zlow, dz2 = dz1 * dz1, dz3 = dz2 * dz1;
                const double c0 = u0l, c1 = u1l, c2 = u2l * 0.5;
                const double a = 6 * (u0u - c0 - c1 * dz1 - c2 * dz2) / dz3;
                const double b = 2 * (u1u - c1 - 2 * c2 * dz1) / dz2;
                const double c = (u2u - 2 * c2) / dz1;
                coeff[0] = c0;
                coeff[1] = c1;
                coeff[2] = c2;
                coeff[3] = (10 * a - 12 * b + 3 * c) / 6;
                coeff[4] = (-15 * a + 21 * b - 6 * c) / (6 * dz1);
                coeff[5] = (2 * a - 3 * b + c) / (2 * dz2);
            }
            bool accepted = true;
            const double step = (rupp - rlow) / 10;
            for (int k = 0; k < 11 && accepted; k++) {
                const double r1 = rlow + step * ((double)k), z = r1 * r1, dz = z - rlow * rlow;
                const double u = coeff[0] + dz * (coeff[1] + dz * (coeff[2] + dz * (coeff[3] + dz * (coeff[4] + dz * coeff[5]))));
                if (std::fabs(u - f(z)) > utol) accepted = false;
            }
            if (accepted) {
                rupp = rlow;
                break;
            }
            dr *= drfrac;
        }
        if (j >= ndr) return CG_ERR_INVALID; // "Andrea spline: try to increase utol/ftol"
        td.r2.push_back(zlow);
        for (double v : coeff) td.c.push_back(v);
        if (rlow <= rmin) break;
    }
    if (i >= mngrid) return CG_ERR_INVALID;
    std::reverse(td.r2.begin(), td.r2.end());
    for (size_t k = 0; k < td.c.size() / 2; k += 6) std::swap_ranges(td.c.begin() + k, td.c.begin() + k + 6, td.c.end() - k - 6);
    return CG_OK;
}

// Splined::spline<T>(...) (splined.hpp:327-337, 364-367): base data (cutoff, kappa, prefactors, scheme id) copied from
// the wrapped scheme, functor switched to the tables.  The table pointers of `out` point into `tables`.
inline int make_splined_from_tables(const cg_scheme_desc &inner, const SplineTable tables[4], cg_scheme_desc &out) {
    out = inner;
    out.functor = CG_SPLINE;
    for (int k = 0; k < 4; k++) {
        const int n = (int)tables[k].r2.size();
        if (n < 2 || n > CG_SPLINE_MAX_KNOTS || (int)tables[k].c.size() != 6 * (n - 1)) return CG_ERR_INVALID;
        out.spline_knots[k] = n;
        out.spline_r2[k] = tables[k].r2.data();
        out.spline_c[k] = tables[k].c.data();
    }
    return CG_OK;
}

inline int generate_spline_tables(const cg_scheme_desc &inner, double utol, SplineTable tables[4]) {
    if (inner.functor == CG_SPLINE) return CG_ERR_INVALID;
    if (!(utol > 0)) utol = 1e-3; // Splined() default, splined.hpp:340
    for (int k = 0; k < 4; k++) {
        const int rc = andrea_generate(
            [&inner, k](double q) {
                double s[4];
                srf_host(inner, q, s);
                return s[k];
            },
            utol, tables[k]);
        if (rc != CG_OK) return rc;
    }
    return CG_OK;
}

} // namespace detail
} // namespace CoulombGalore
