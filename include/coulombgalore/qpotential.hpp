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

// the bare q-Pochhammer symbol (reference qpotential.hpp:50-61, l = 0 only)
inline double qPochhammerSymbol(double q, int l = 0, int P = 300) {
    if (l != 0) throw std::runtime_error("qPochhammerSymbol: only l = 0 is supported");
    cg_scheme_desc d = detail::desc_base(CG_QPOTENTIAL, 1.0, infinity);
    d.order = P;
    double s[4];
    detail::QPotentialSrf<0>(d).eval<0>(q, s);
    return s[0];
}

} // namespace CoulombGalore
