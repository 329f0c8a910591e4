// coulombgalore/detail/srf.hpp -- short-range-function functors S(q), S'(q), S''(q), S'''(q), q = r/R_c.
//
// One POD functor per truncation scheme, constructed from a cg_scheme_desc and usable on the host (scheme classes)
// and inside the sm_100a pair kernels (passed by value as a __grid_constant__ kernel parameter, so every member is
// a constant-bank operand).  `eval<ND>(q, s, tab)` fills s[0..ND]; sub-expressions shared between the derivatives
// (exp(-eta^2 q^2), erfc, powers of q) are evaluated once, where the reference re-evaluates them in each of its
// four virtual functions.  Constants that depend only on the scheme parameters are hoisted into the descriptor
// with the same libm calls and arguments the reference uses, so hoisting changes no value.
//
// Reference formulas: plain.hpp:47-50, reactionfield.hpp:60-73, poisson.hpp:119-220, qpotential.hpp:50-192,
// fanourgakis.hpp:49-61, wolf.hpp:60-71, fennell.hpp:59-73, zerodipole.hpp:61-80, zahn.hpp:60-76,
// ewald.hpp:175-195, ewald_truncated.hpp:71-82, splined.hpp:167-176.
#pragma once
#include "../../cg_b200.h"
#include "hd.hpp"
#include "specfun.hpp"

namespace CoulombGalore {
namespace detail {

// what `tab` points to in eval(): nothing, the packed Andrea tables of a Splined scheme, or the erfc/exp table of
// specfun.hpp (device only; the host evaluates libm directly)
enum { kTableNone = 0, kTableSpline = 1, kTableErfc = 2 };

// defaults every functor inherits
struct SrfDefaults {
    // true: the functor offers angcor(q, tab) = S - q S' in closed form (cheaper than S and S' separately)
    static constexpr bool kHasAngcor = false;
    // schemes evaluated per pair (ErfcMultiSrf: several, each with its own accumulators)
    static constexpr int kSchemes = 1;
    static constexpr bool kErfcArgIsEtaQ = false; // erfc family: the only erfc / Gaussian argument is eta q (<= eta inside the cut-off)
};

// 2/sqrt(pi); the reference forms it as 2/(2*sqrt(atan(1))) (e.g. wolf.hpp:36), equal to within one ulp
constexpr double kTwoOverSqrtPi = 1.1283791670955125739;

// ---------------------------------------------------------------------------------------------------------------
struct PlainSrf : SrfDefaults { // plain.hpp:47-50
    static constexpr int kTable = kTableNone;
    CG_HD PlainSrf() {}
    CG_HD explicit PlainSrf(const cg_scheme_desc &) {}
    template <int ND, class T = double> CG_HD void eval(T, T *s, const double * = nullptr) const {
        s[0] = T(1);
        if (ND >= 1) s[1] = T(0);
        if (ND >= 2) s[2] = T(0);
        if (ND >= 3) s[3] = T(0);
    }
};

struct ReactionFieldSrf : SrfDefaults { // reactionfield.hpp:60-73: S = 1 + a q^3 - b q
    static constexpr int kTable = kTableNone;
    double a, b;
    CG_HD ReactionFieldSrf() : a(0), b(0) {}
    CG_HD explicit ReactionFieldSrf(const cg_scheme_desc &d) : a(d.rf_a), b(d.rf_b) {}
    template <int ND, class T = double> CG_HD void eval(T q, T *s, const double * = nullptr) const {
        const T A = (T)a, B = (T)b;
        const T aq = A * q, aq2 = aq * q;
        s[0] = T(1) + (aq2 - B) * q;
        if (ND >= 1) s[1] = T(3) * aq2 - B;
        if (ND >= 2) s[2] = T(6) * aq;
        if (ND >= 3) s[3] = T(6) * A;
    }
};

struct FanourgakisSrf : SrfDefaults { // fanourgakis.hpp:49-61
    static constexpr int kTable = kTableNone;
    CG_HD FanourgakisSrf() {}
    CG_HD explicit FanourgakisSrf(const cg_scheme_desc &) {}
    template <int ND, class T = double> CG_HD void eval(T q, T *s, const double * = nullptr) const {
        const T q2 = q * q, a = T(1) - q, a2 = a * a;
        s[0] = (a2 * a2) * (T(1) + q * (T(2.25) + q * (T(3) + T(2.5) * q)));
        if (ND >= 1) s[1] = T(-1.75) + (q2 * q2) * (T(26.25) + q * (T(-42) + T(17.5) * q));
        if (ND >= 2) s[2] = T(105) * (q2 * q) * a2;
        if (ND >= 3) s[3] = T(525) * q2 * (q - T(0.6)) * (q - T(1));
    }
};

// ---------------------------------------------------------------------------------------------------------------
// qPotential: S = prod_{n=1}^{P} (1 - q^n) = (1-q)^P prod_{n=2}^{P} g_n,  g_n = sum_{k<n} q^k  (qpotential.hpp:50-61).
// The factored form is kept (all g_n are positive, no cancellation as q -> 1); derivatives by Leibniz' rule over
// the factors with the recurrence g_n = 1 + q g_{n-1}, O(P) work instead of the reference's O(P^2) powi loops.
// ORDER > 0: compile-time order (fully unrolled); ORDER == 0: run-time order.
template <int ND, class T> CG_HD void jet_mul(T *c, const T *g) { // c <- c * g as Taylor jets up to order ND
    if (ND >= 3) c[3] = c[3] * g[0] + T(3) * (c[2] * g[1] + c[1] * g[2]) + c[0] * g[3];
    if (ND >= 2) c[2] = c[2] * g[0] + T(2) * c[1] * g[1] + c[0] * g[2];
    if (ND >= 1) c[1] = c[1] * g[0] + c[0] * g[1];
    c[0] = c[0] * g[0];
}

template <int ORDER> struct QPotentialSrf : SrfDefaults {
    static constexpr int kTable = kTableNone;
    int order;
    CG_HD QPotentialSrf() : order(ORDER) {}
    CG_HD explicit QPotentialSrf(const cg_scheme_desc &d) : order(d.order) {}
    template <int ND, class T = double> CG_HD void eval(T q, T *s, const double * = nullptr) const {
        const int P = ORDER > 0 ? ORDER : order;
        T g[4] = {T(1), T(0), T(0), T(0)}; // jet of g_1
        T c[4] = {T(1), T(0), T(0), T(0)}; // jet of prod g_n
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int n = 2; n <= P; n++) {
            if (ND >= 3) g[3] = T(3) * g[2] + q * g[3];
            if (ND >= 2) g[2] = T(2) * g[1] + q * g[2];
            if (ND >= 1) g[1] = g[0] + q * g[1];
            g[0] = T(1) + q * g[0];
            jet_mul<ND, T>(c, g);
        }
        // jet of (1-q)^P
        const T a = T(1) - q;
        T d[4];
        T ap = (P > 3) ? powi(a, P - 3) : T(1); // a^(max(P-3,0))
        d[3] = (P >= 3) ? -(T)(P * (P - 1) * (P - 2)) * ap : T(0);
        if (P >= 3) ap *= a;
        d[2] = (P >= 2) ? (T)(P * (P - 1)) * ap : T(0);
        if (P >= 2) ap *= a;
        d[1] = (P >= 1) ? -(T)P * ap : T(0);
        if (P >= 1) ap *= a;
        d[0] = ap;
        jet_mul<ND, T>(c, d);
        s[0] = c[0];
        if (ND >= 1) s[1] = c[1];
        if (ND >= 2) s[2] = c[2];
        if (ND >= 3) s[3] = c[3];
    }
};

// ---------------------------------------------------------------------------------------------------------------
// erfc family.  x = eta q, E = exp(-x^2), k1 = 2 eta / sqrt(pi); c = erfc(eta), e = exp(-eta^2), c2 = c + k1 e.
enum ErfcKind { kWolf, kFennell, kZeroDipole, kZahn, kEwaldT, kEwaldPlain };

template <int KIND> struct ErfcFamilySrf : SrfDefaults {
    static constexpr int kTable = kTableErfc;
    static constexpr bool kErfcArgIsEtaQ = true;
    static constexpr bool kHasAngcor = true;
    double eta, eta2, k1, c, e, c2, invF0;
    CG_HD ErfcFamilySrf() : eta(0), eta2(0), k1(0), c(0), e(0), c2(0), invF0(1) {}
    CG_HD explicit ErfcFamilySrf(const cg_scheme_desc &d)
        : eta(d.eta), eta2(d.eta * d.eta), k1(d.eta * kTwoOverSqrtPi), c(d.erfc_eta), e(d.exp_eta2),
          c2(d.erfc_eta + d.eta * kTwoOverSqrtPi * d.exp_eta2), invF0(KIND == kEwaldT ? 1.0 / d.F0 : 1.0) {}
    template <int ND, class T = double, class TAB = const double *> CG_HD void eval(T q, T *s, const TAB &tab = TAB()) const {
        const T x = (T)eta * q;
        T E, ec;
        erfc_and_gauss(x, ec, E, tab); // ec = erfc(x), E = exp(-x^2)
        const T K1 = (T)k1, C = (T)c, C2 = (T)c2, ETA2 = (T)eta2;
        const T k1E = K1 * E;
        const T g2 = T(2) * ETA2 * q * k1E;                  // 4 eta^3 q E / sqrt(pi)
        const T g3 = T(2) * ETA2 * k1E * (T(1) - T(2) * x * x); // 4 eta^3 (1 - 2 x^2) E / sqrt(pi)
        if (KIND == kWolf) { // wolf.hpp:60-71
            s[0] = ec - q * C;
            if (ND >= 1) s[1] = -k1E - C;
            if (ND >= 2) s[2] = g2;
            if (ND >= 3) s[3] = g3;
        } else if (KIND == kFennell) { // fennell.hpp:59-73
            s[0] = ec - q * C + (q - T(1)) * q * C2;
            if (ND >= 1) s[1] = -k1E + T(2) * q * C2 - (C2 + C);
            if (ND >= 2) s[2] = g2 + T(2) * C2;
            if (ND >= 3) s[3] = g3;
        } else if (KIND == kZeroDipole) { // zerodipole.hpp:61-80
            const T q2 = q * q;
            s[0] = ec - q * C + T(0.5) * (q2 - T(1)) * q * C2;
            if (ND >= 1) s[1] = -k1E + T(1.5) * q2 * C2 - (T(0.5) * C2 + C);
            if (ND >= 2) s[2] = g2 + T(3) * q * C2;
            if (ND >= 3) s[3] = g3 + T(3) * C2;
        } else if (KIND == kZahn) { // zahn.hpp:60-76
            s[0] = ec - (q - T(1)) * q * C2;
            if (ND >= 1) s[1] = -k1E - (T(2) * q - T(1)) * C2;
            if (ND >= 2) s[2] = g2 - T(2) * C2;
            if (ND >= 3) s[3] = g3;
        } else if (KIND == kEwaldT) { // ewald_truncated.hpp:71-82
            const T IF0 = (T)invF0, k1e = K1 * (T)e;
            s[0] = (ec - C - (T(1) - q) * k1e) * IF0;
            if (ND >= 1) s[1] = -(k1E - k1e) * IF0;
            if (ND >= 2) s[2] = g2 * IF0;
            if (ND >= 3) s[3] = g3 * IF0;
        } else { // Ewald without screening, ewald.hpp:175-195 with zeta = 0
            s[0] = ec;
            if (ND >= 1) s[1] = -k1E;
            if (ND >= 2) s[2] = g2;
            if (ND >= 3) s[3] = g3;
        }
    }
    // S - q S' with the q-linear and q-polynomial terms of S and q S' cancelled analytically: what ion fields and
    // forces need (core.hpp:352-365 with kappa = 0).  erfc(x) + (2/sqrt(pi)) x exp(-x^2) is common to all kinds.
    template <class T = double, class TAB = const double *> CG_HD T angcor(T q, const TAB &tab) const {
        const T x = (T)eta * q;
        T E, ec;
        erfc_and_gauss(x, ec, E, tab);
        const T g = ec + (q * (T)k1) * E;
        if (KIND == kWolf || KIND == kEwaldPlain) return g;
        if (KIND == kFennell) return g - (q * q) * (T)c2;
        if (KIND == kZeroDipole) return g - (q * q) * (q * (T)c2);
        if (KIND == kZahn) return g + (q * q) * (T)c2;
        return (g - (T)(c + k1 * e)) * (T)invF0; // kEwaldT
    }
};

// Several schemes of the erfc family on ONE pair (batched path only; the reference evaluates one scheme per call): schemes that
// share eta = alpha R_c and the cut-off differ only in a cubic polynomial added to erfc(eta q), so the expensive part --
// erfc(x) and exp(-x^2) -- is evaluated once per pair and every scheme costs a short Horner on top:
//     S_k      = m_k ( erfc(x) + a0_k + a1_k q + a2_k q^2 + a3_k q^3 )
//     S_k-qS_k' = m_k ( erfc(x) + k1 q exp(-x^2) + b0_k + b2_k q^2 + b3_k q^3 )        (what ion fields and forces need)
// with the coefficients of wolf.hpp:60-71, fennell.hpp:59-73, zerodipole.hpp:61-80, zahn.hpp:60-76, ewald_truncated.hpp:71-82
// and ewald.hpp:175-195 (zeta = 0).  NS schemes, NS accumulators: cg_eval_multi_device (include/cg_b200.h).
template <int NS> struct ErfcMultiSrf : SrfDefaults {
    static constexpr int kTable = kTableErfc;
    static constexpr bool kErfcArgIsEtaQ = true;
    static constexpr int kSchemes = NS;
    double eta, k1;
    double m[NS], a0[NS], a1[NS], a2[NS], a3[NS], b0[NS], b2[NS], b3[NS];
    CG_HD ErfcMultiSrf() : eta(0), k1(0) {}
    // d: the descriptor of the first scheme; the functor ids of all NS schemes are packed into d.reserved0, 8 bits each
    CG_HD explicit ErfcMultiSrf(const cg_scheme_desc &d) : eta(d.eta), k1(d.eta * kTwoOverSqrtPi) {
        const double c = d.erfc_eta, k1e = k1 * d.exp_eta2, c2 = c + k1e;
        for (int k = 0; k < NS; k++) {
            const int id = (d.reserved0 >> (8 * k)) & 255;
            m[k] = 1.0;
            a0[k] = a1[k] = a2[k] = a3[k] = b0[k] = b2[k] = b3[k] = 0.0;
            if (id == CG_WOLF) {
                a1[k] = -c;
            } else if (id == CG_FENNELL) {
                a1[k] = -c - c2, a2[k] = c2, b2[k] = -c2;
            } else if (id == CG_ZERODIPOLE) {
                a1[k] = -c - 0.5 * c2, a3[k] = 0.5 * c2, b3[k] = -c2;
            } else if (id == CG_ZAHN) {
                a1[k] = c2, a2[k] = -c2, b2[k] = c2;
            } else if (id == CG_EWALDT) {
                m[k] = 1.0 / d.F0, a0[k] = -c - k1e, a1[k] = k1e, b0[k] = -(c + k1e);
            } // CG_EWALD (unscreened): erfc alone
        }
    }
    // single-scheme interface (the per-functor probe): scheme 0
    template <int ND, class T = double, class TAB = const double *> CG_HD void eval(T q, T *s, const TAB &tab = TAB()) const {
        const T x = (T)eta * q;
        T E, ec;
        erfc_and_gauss(x, ec, E, tab);
        const T k1E = (T)k1 * E, e2 = (T)eta * (T)eta;
        s[0] = (T)m[0] * (ec + (T)a0[0] + q * ((T)a1[0] + q * ((T)a2[0] + q * (T)a3[0])));
        if (ND >= 1) s[1] = (T)m[0] * (-k1E + (T)a1[0] + q * (T(2) * (T)a2[0] + q * T(3) * (T)a3[0]));
        if (ND >= 2) s[2] = (T)m[0] * (T(2) * e2 * q * k1E + T(2) * (T)a2[0] + T(6) * q * (T)a3[0]);
        if (ND >= 3) s[3] = (T)m[0] * (T(2) * e2 * k1E * (T(1) - T(2) * x * x) + T(6) * (T)a3[0]);
    }
};

// Ewald with Debye screening (zeta = R_c / lambda_D > 0), ewald.hpp:175-195.
//   S = 1/2 [ erfc(x+) e^{2 zeta q} + erfc(x-) ],  x+- = eta q +- zeta/(2 eta)
struct EwaldScreenedSrf : SrfDefaults {
    static constexpr int kTable = kTableErfc;
    double eta, eta2, k1, zeta, zeta2, zeta3, half_z_over_eta, z_over_eta, z2_over_eta2;
    CG_HD EwaldScreenedSrf() : eta(0), eta2(0), k1(0), zeta(0), zeta2(0), zeta3(0), half_z_over_eta(0), z_over_eta(0), z2_over_eta2(0) {}
    CG_HD explicit EwaldScreenedSrf(const cg_scheme_desc &d)
        : eta(d.eta), eta2(d.eta * d.eta), k1(d.eta * kTwoOverSqrtPi), zeta(d.zeta), zeta2(d.zeta * d.zeta),
          zeta3(d.zeta * d.zeta * d.zeta), half_z_over_eta(d.zeta / (2.0 * d.eta)), z_over_eta(d.zeta / d.eta),
          z2_over_eta2((d.zeta * d.zeta) / (d.eta * d.eta)) {}
    template <int ND, class T = double, class TAB = const double *> CG_HD void eval(T q, T *s, const TAB &tab = TAB()) const {
        const T xq = (T)eta * q, H = (T)half_z_over_eta, ZE = (T)z_over_eta, Z = (T)zeta, K1 = (T)k1;
        const T xm = xq - H, xp = xq + H;
        T ecm, Em, ecp, Ep; // erfc(x-), exp(-x-^2), erfc(x+)
        erfc_and_gauss<true>(xm, ecm, Em, tab);
        erfc_and_gauss<false>(xp, ecp, Ep, tab);
        // P = erfc(x+) exp(2 zeta q).  Since x+^2 - x-^2 = 2 zeta q, exp(2 zeta q) = exp(-x-^2) / exp(-x+^2): on the device the
        // two Gaussians are already there, a division replaces the exponential (inside the table range, where both are
        // accurate to a few ulp; beyond it the exponential is evaluated as the reference writes it).
        T P;
#ifdef __CUDA_ARCH__
        if (xp < T(7.5)) P = ecp * (Em / Ep);
        else P = ecp * exp_any(T(2) * Z * q);
#else
        (void)Ep;
        P = ecp * exp_any(T(2) * Z * q);
#endif
        s[0] = T(0.5) * (P + ecm);
        if (ND >= 1) s[1] = -K1 * Em + Z * P;
        if (ND >= 2) s[2] = T(2) * K1 * (T)eta * (xq - ZE) * Em + T(2) * (T)zeta2 * P;
        if (ND >= 3) s[3] = T(2) * K1 * (T)eta2 * (T(1) - T(2) * (xq - ZE) * xm - (T)z2_over_eta2) * Em + T(4) * (T)zeta3 * P;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Poisson (poisson.hpp:119-220), run-time C and D.  q~ = q, or (1 - e^{2 kappa* q}) * yukawa_denom when screened.
struct PoissonSrf : SrfDefaults {
    static constexpr int kTable = kTableNone;
    int C, D, yukawa;
    double rk, denom, binomCDC;
    double w[CG_MAX_POISSON_C];
    CG_HD PoissonSrf() : C(1), D(-1), yukawa(0), rk(0), denom(0), binomCDC(0) {}
    CG_HD explicit PoissonSrf(const cg_scheme_desc &d)
        : C(d.C), D(d.D), yukawa(d.yukawa_map), rk(d.reduced_kappa), denom(d.yukawa_denom), binomCDC(d.binomCDC) {
        for (int i = 0; i < CG_MAX_POISSON_C; i++) w[i] = d.poisson_w[i];
    }
    template <int ND, class T = double> CG_HD void eval(T q, T *s, const double * = nullptr) const {
        if (D == -C) { // poisson.hpp:120-122 etc.
            s[0] = T(1);
            if (ND >= 1) s[1] = T(0);
            if (ND >= 2) s[2] = T(0);
            if (ND >= 3) s[3] = T(0);
            return;
        }
        T qp = q, d1 = T(1), d2 = T(0), d3 = T(0);
        if (yukawa) { // :126, :147-149, :174-176, :204-207
            const T RK = (T)rk, DEN = (T)denom;
            const T ex = exp_any(T(2) * RK * q);
            qp = (T(1) - ex) * DEN;
            d1 = T(-2) * RK * ex * DEN;
            d2 = T(2) * RK * d1;
            d3 = T(2) * RK * d2;
        }
        if (D == 0 && C == 1) { // :128-130, :141-143: the derivatives are returned as zero
            s[0] = T(1) - qp;
            if (ND >= 1) s[1] = T(0);
            if (ND >= 2) s[2] = T(0);
            if (ND >= 3) s[3] = T(0);
            return;
        }
        // tmp1 = sum_c w_c qp^c, tmp2 = d tmp1 / d qp  (Horner)
        T t1 = T(0), t2 = T(0);
        for (int c = C - 1; c >= 0; c--) {
            t2 = t2 * qp + t1;
            t1 = t1 * qp + (T)w[c];
        }
        const T a = T(1) - qp;
        const T aD = powi(a, D);
        s[0] = aD * a * t1;
        if (ND >= 1) {
            const T dS = -(T)(D + 1) * aD * t1 + aD * a * t2;
            s[1] = dS * d1;
            if (ND >= 2) {
                const T d2S = (T)binomCDC * powi(a, D - 1) * powi(qp, C - 1);
                // the reference adds the chain-rule terms only inside `if (use_yukawa_screening)`; d2 = d3 = 0 otherwise
                s[2] = d2S * d1 * d1 + dS * d2;
                if (ND >= 3) {
                    const T d3S = (T)binomCDC * powi(a, D - 2) * powi(qp, C - 2) * ((T(2) - (T)(C + D)) * qp + (T)C - T(1));
                    s[3] = d3S * d1 * d1 * d1 + T(3) * d2S * d1 * d2 + dS * d3;
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Splined (splined.hpp:321-386): four independent Andrea tables; eval = lower_bound over the knots, then a quintic
// Horner in dz = q - knot (splined.hpp:167-176).  Packed layout in `tab` (global memory on the device, staged to
// shared memory by the kernels; a host pointer for the scheme classes), all offsets in doubles:
//   for k in 0..3:  knots[k][0..n_k)  at koff[k],   coefficients[k][6*(n_k-1)]  at coff[k] (16-byte aligned)
//   lookup grid at lut_off: kSplineGrid 32-bit words, byte k of word c = the interval of table k that contains the
//   left edge of grid cell c, i.e. a lower bound for every q in [c, c+1)/kSplineGrid; the device then advances at
//   most `kmax` intervals with exact FP64 comparisons, which reproduces std::lower_bound bit for bit.
constexpr int kSplineGrid = 256;

struct SplineSrf : SrfDefaults {
    static constexpr int kTable = kTableSpline;
    int n[4], koff[4], coff[4], lut_off, kmax, total;
    const double *tab;
    CG_HD SplineSrf() : lut_off(0), kmax(0), total(0), tab(nullptr) {
        for (int k = 0; k < 4; k++) n[k] = koff[k] = coff[k] = 0;
    }
    // layout only (tab is attached afterwards: the device copy for kernels, pack_from() on the host)
    inline explicit SplineSrf(const cg_scheme_desc &d) : lut_off(0), kmax(0), total(0), tab(nullptr) {
        int t = 0;
        for (int k = 0; k < 4; k++) {
            n[k] = d.spline_knots[k];
            koff[k] = t;
            t += n[k] + (n[k] & 1); // keep the coefficients 16-byte aligned
            coff[k] = t;
            t += 6 * (n[k] - 1);
        }
        lut_off = t;
        t += kSplineGrid / 2; // kSplineGrid 32-bit words
        total = t;
        kmax = max_advance(d);
    }
    // number of doubles of the packed table
    static inline int packed_size(const cg_scheme_desc &d) { return SplineSrf(d).total; }
    static inline int interval_of(const double *knots, int nk, double q) { // (#knots strictly below q) - 1, clamped
        int pos = -1;
        for (int i = 0; i < nk; i++) pos += (knots[i] < q) ? 1 : 0;
        return pos < 0 ? 0 : (pos > nk - 2 ? nk - 2 : pos);
    }
    static inline int max_advance(const cg_scheme_desc &d) {
        int m = 0;
        for (int k = 0; k < 4; k++)
            for (int c = 0; c < kSplineGrid; c++) {
                const int lo = interval_of(d.spline_r2[k], d.spline_knots[k], (double)c / kSplineGrid);
                const int hi = interval_of(d.spline_r2[k], d.spline_knots[k], (double)(c + 1) / kSplineGrid);
                m = (hi - lo > m) ? hi - lo : m;
            }
        return m;
    }
    // host-side packing; `dst` must hold packed_size(d) doubles
    inline void pack_from(const cg_scheme_desc &d, double *dst) {
        *this = SplineSrf(d);
        for (int i = 0; i < total; i++) dst[i] = 0.0;
        for (int k = 0; k < 4; k++) {
            for (int i = 0; i < n[k]; i++) dst[koff[k] + i] = d.spline_r2[k][i];
            for (int i = 0; i < 6 * (n[k] - 1); i++) dst[coff[k] + i] = d.spline_c[k][i];
        }
        unsigned int *lut = reinterpret_cast<unsigned int *>(dst + lut_off);
        for (int c = 0; c < kSplineGrid; c++) {
            unsigned int w = 0;
            for (int k = 0; k < 4; k++) w |= (unsigned int)interval_of(d.spline_r2[k], n[k], (double)c / kSplineGrid) << (8 * k);
            lut[c] = w;
        }
        tab = dst;
    }
#ifdef __CUDA_ARCH__
    // device: grid lookup + at most kmax exact comparisons, coefficients fetched as three 16-byte shared-memory loads.
    // The four tables are searched and evaluated side by side (independent loads and Horner chains in flight together):
    // the kernel is latency-bound at its occupancy, not throughput-bound.
    // (fp32 kernels evaluate the tables in fp64 and round the results: the tables are the scheme)
    template <int ND, class T = double> __device__ __forceinline__ void eval(T qin, T *s, const double *t = nullptr) const {
        const double q = (double)qin;
        const unsigned base = (unsigned)__cvta_generic_to_shared(t);
        int c = __double2int_rd(q * (double)kSplineGrid);
        c = min(max(c, 0), kSplineGrid - 1);
        unsigned int w;
        asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(base + 8u * (unsigned)lut_off + 4u * (unsigned)c));
        int pos[ND + 1];
#pragma unroll
        for (int k = 0; k <= ND; k++) pos[k] = (int)((w >> (8 * k)) & 255u);
        for (int st = 0; st < kmax; st++) {
#pragma unroll
            for (int k = 0; k <= ND; k++) {
                const int nx = pos[k] + 1;
                double kn;
                asm("ld.shared.f64 %0, [%1];" : "=d"(kn) : "r"(base + 8u * (unsigned)(koff[k] + nx)));
                pos[k] = (nx <= n[k] - 2 && kn < q) ? nx : pos[k];
            }
        }
        double k0[ND + 1], cf[ND + 1][6];
#pragma unroll
        for (int k = 0; k <= ND; k++) {
            asm("ld.shared.f64 %0, [%1];" : "=d"(k0[k]) : "r"(base + 8u * (unsigned)(koff[k] + pos[k])));
            const unsigned ca = base + 8u * (unsigned)(coff[k] + 6 * pos[k]);
            asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(cf[k][0]), "=d"(cf[k][1]) : "r"(ca));
            asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(cf[k][2]), "=d"(cf[k][3]) : "r"(ca + 16u));
            asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(cf[k][4]), "=d"(cf[k][5]) : "r"(ca + 32u));
        }
        double dz[ND + 1], v[ND + 1];
#pragma unroll
        for (int k = 0; k <= ND; k++) {
            dz[k] = q - k0[k];
            v[k] = cf[k][5];
        }
#pragma unroll
        for (int j = 4; j >= 0; j--) {
#pragma unroll
            for (int k = 0; k <= ND; k++) v[k] = cf[k][j] + dz[k] * v[k]; // same Horner form as the reference, all tables in step
        }
#pragma unroll
        for (int k = 0; k <= ND; k++) s[k] = (T)v[k];
    }
#else
    inline double table(int k, double q, const double *t) const {
        const double *kn = t + koff[k];
        const int pos = interval_of(kn, n[k], q); // q == 0 asserts in the reference; clamped here
        const double dz = q - kn[pos];
        const double *c = t + coff[k] + 6 * pos;
        return c[0] + dz * (c[1] + dz * (c[2] + dz * (c[3] + dz * (c[4] + dz * (c[5])))));
    }
    template <int ND, class T = double> inline void eval(T qin, T *s, const double *t = nullptr) const {
        const double *p = t ? t : tab;
        const double q = (double)qin;
        s[0] = (T)table(0, q, p);
        if (ND >= 1) s[1] = (T)table(1, q, p);
        if (ND >= 2) s[2] = (T)table(2, q, p);
        if (ND >= 3) s[3] = (T)table(3, q, p);
    }
#endif
};

} // namespace detail
} // namespace CoulombGalore
