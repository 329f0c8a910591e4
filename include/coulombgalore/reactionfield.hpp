// coulombgalore/reactionfield.hpp -- ReactionField: same class name and constructor as the reference (reactionfield.hpp:39-86); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class ReactionField : public detail::FunctorScheme<ReactionField, detail::ReactionFieldSrf> {
  public:
    inline ReactionField(double cutoff, double epsRF, double epsr, bool shifted) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_reactionfield(cutoff, epsRF, epsr, shifted, d), "ReactionField: invalid parameters");
        install(d, "Reaction-field", "10.1080/00268977300102101");
    }
};

} // namespace CoulombGalore
