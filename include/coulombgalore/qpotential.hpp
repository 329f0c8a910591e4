// coulombgalore/qpotential.hpp -- qPotential and qPotentialFixedOrder<order>: same class name and constructor as the reference (qpotential.hpp:195-282); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class qPotential : public detail::FunctorScheme<qPotential, detail::QPotentialSrf<0>> {
  public:
    inline qPotential(double cutoff, int order) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_qpotential(cutoff, order, false, d), "qPotential: invalid parameters");
        install(d, "qpotential", "10.1039/c9cp03875b");
    }
};

/** compile-time order: the Leibniz product over the q-Pochhammer factors is fully unrolled */
template <int order> class qPotentialFixedOrder : public detail::FunctorScheme<qPotentialFixedOrder<order>, detail::QPotentialSrf<order>> {
  public:
    inline qPotentialFixedOrder(double cutoff) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_qpotential(cutoff, order, true, d), "qPotentialFixedOrder: invalid parameters");
        this->install(d, "qpotential", "10.1039/c9cp03875b");
    }
};

namespace detail {
// Value and first three derivatives of what the reference's qPochhammerSymbol* functions evaluate (qpotential.hpp:50-192):
//     (1 - q)^P  prod_{n=1}^{P} g_{n+l}(q),      g_m(q) = sum_{k=0}^{m-1} q^k   (an empty sum, m <= 0, is zero)
// i.e. prod_{n=1}^{P} (1 - q^{n+l}).  Same Leibniz product over the factors as the device functor (detail/srf.hpp,
// QPotentialSrf), with the recurrence g_m = 1 + q g_{m-1} started at m = l + 1: O(P + l) work where the reference spends
// O(P (P + l)) powi calls per function, and no division by the factors.
inline void qpochhammer_jet(double q, int l, int P, double s[4]) {
    double c[4] = {1.0, 0.0, 0.0, 0.0}; // jet of the product
    if (P > 0 && 1 + l <= 0) {          // the first factor is an empty sum
        s[0] = s[1] = s[2] = s[3] = 0.0;
        return;
    }
    double g[4] = {1.0, 0.0, 0.0, 0.0}; // jet of g_1
    for (int m = 2; m <= l + P; m++) {
        g[3] = 3.0 * g[2] + q * g[3];
        g[2] = 2.0 * g[1] + q * g[2];
        g[1] = g[0] + q * g[1];
        g[0] = 1.0 + q * g[0];
        if (m >= l + 1) jet_mul<3, double>(c, g);
    }
    const double a = 1.0 - q;
    double d[4];
    d[0] = powi(a, P);
    d[1] = P > 0 ? -(double)P * powi(a, P - 1) : 0.0;
    d[2] = P > 1 ? (double)P * (double)(P - 1) * powi(a, P - 2) : 0.0;
    d[3] = P > 2 ? -(double)P * (double)(P - 1) * (double)(P - 2) * powi(a, P - 3) : 0.0;
    jet_mul<3, double>(c, d);
    for (int k = 0; k < 4; k++) s[k] = c[k];
}
} // namespace detail

/** q-Pochhammer symbol and its derivatives as free functions, reference qpotential.hpp:50, 66, 97, 138 */
inline double qPochhammerSymbol(double q, int l = 0, int P = 300) {
    double s[4];
    detail::qpochhammer_jet(q, l, P, s);
    return s[0];
}
inline double qPochhammerSymbolDerivative(double q, int l = 0, int P = 300) {
    double s[4];
    detail::qpochhammer_jet(q, l, P, s);
    return s[1];
}
inline double qPochhammerSymbolSecondDerivative(double q, int l = 0, int P = 300) {
    double s[4];
    detail::qpochhammer_jet(q, l, P, s);
    return s[2];
}
inline double qPochhammerSymbolThirdDerivative(double q, int l = 0, int P = 300) {
    double s[4];
    detail::qpochhammer_jet(q, l, P, s);
    return s[3];
}

} // namespace CoulombGalore
