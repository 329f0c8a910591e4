// coulombgalore/zahn.hpp -- Zahn: same class name and constructor as the reference (zahn.hpp:34-85); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

class Zahn : public detail::FunctorScheme<Zahn, detail::ErfcFamilySrf<detail::kZahn>> {
  public:
    inline Zahn(double cutoff, double alpha) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_zahn(cutoff, alpha, d), "Zahn: invalid parameters");
        install(d, "Zahn", "10.1021/jp025949h");
    }
};

} // namespace CoulombGalore
