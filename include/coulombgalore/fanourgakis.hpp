// coulombgalore/fanourgakis.hpp -- Fanourgakis: same class name and constructor as the reference (fanourgakis.hpp:37-69); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class Fanourgakis : public detail::FunctorScheme<Fanourgakis, detail::FanourgakisSrf> {
  public:
    inline Fanourgakis(double cutoff) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_fanourgakis(cutoff, d), "Fanourgakis: invalid parameters");
        install(d, "fanourgakis", "10.1063/1.3216520");
    }
};

} // namespace CoulombGalore
