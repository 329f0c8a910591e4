// coulombgalore/plain.hpp -- Plain: same class name and constructor as the reference (plain.hpp:35-57); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class Plain : public detail::FunctorScheme<Plain, detail::PlainSrf> {
  public:
    inline Plain(double debye_length = infinity) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_plain(debye_length, d), "Plain: invalid parameters");
        install(d, "plain", "Premier memoire sur l electricite et le magnetisme by Charles-Augustin de Coulomb");
    }
};

} // namespace CoulombGalore
