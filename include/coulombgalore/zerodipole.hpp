// coulombgalore/zerodipole.hpp -- ZeroDipole: same class name and constructor as the reference (zerodipole.hpp:33-90); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class ZeroDipole : public detail::FunctorScheme<ZeroDipole, detail::ErfcFamilySrf<detail::kZeroDipole>> {
  public:
    inline ZeroDipole(double cutoff, double alpha) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_zerodipole(cutoff, alpha, d), "ZeroDipole: invalid parameters");
        install(d, "ZeroDipole", "10.1063/1.3582791");
    }
};

} // namespace CoulombGalore
