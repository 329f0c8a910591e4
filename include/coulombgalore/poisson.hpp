// coulombgalore/poisson.hpp -- Poisson: same class name and constructor as the reference (poisson.hpp:62-232; the constructor throws std::runtime_error on the same conditions, :81-86); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class Poisson : public detail::FunctorScheme<Poisson, detail::PoissonSrf> {
  public:
    inline Poisson(double cutoff, signed int C, signed int D, double debye_length = infinity) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_poisson(cutoff, C, D, debye_length, d), "Poisson: invalid parameters");
        install(d, "poisson", "10/c5fr");
    }
};

} // namespace CoulombGalore
