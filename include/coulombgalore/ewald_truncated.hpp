// coulombgalore/ewald_truncated.hpp -- EwaldTruncated: same class name and constructor as the reference (ewald_truncated.hpp:37-120; the O(N) surface_energy member is outside the pair path); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class EwaldTruncated : public detail::FunctorScheme<EwaldTruncated, detail::ErfcFamilySrf<detail::kEwaldT>> {
  public:
    inline EwaldTruncated(double cutoff, double alpha, double surface_dielectric_constant = infinity, double debye_length = infinity) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_ewaldt(cutoff, alpha, surface_dielectric_constant, debye_length, d), "EwaldTruncated: invalid parameters");
        install(d, "EwaldT real-space", "XYZ");
    }
};

} // namespace CoulombGalore
