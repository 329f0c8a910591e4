// coulombgalore/wolf.hpp -- Wolf: same class name and constructor as the reference (wolf.hpp:33-80); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class Wolf : public detail::FunctorScheme<Wolf, detail::ErfcFamilySrf<detail::kWolf>> {
  public:
    inline Wolf(double cutoff, double alpha) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_wolf(cutoff, alpha, d), "Wolf: invalid parameters");
        install(d, "Wolf", "10.1063/1.478738");
    }
};

} // namespace CoulombGalore
