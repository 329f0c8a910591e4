// coulombgalore/fennell.hpp -- Fennell: same class name and constructor as the reference (fennell.hpp:33-82); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class Fennell : public detail::FunctorScheme<Fennell, detail::ErfcFamilySrf<detail::kFennell>> {
  public:
    inline Fennell(double cutoff, double alpha) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_fennell(cutoff, alpha, d), "Fennell: invalid parameters");
        install(d, "Fennell", "10.1063/1.2206581");
    }
};

} // namespace CoulombGalore
