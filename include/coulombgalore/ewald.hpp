// coulombgalore/ewald.hpp -- Ewald (real space): same class name and constructor as the reference (ewald.hpp:114-212); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

// The reciprocal-space classes of the reference's ewald.hpp (ReciprocalEwaldState / ReciprocalEwaldGaussian, :44-94,
// :214-385) live in reciprocal.hpp (a k-space sum on the GPU, not a pair scheme).
class Ewald : public EnergyImplementation<Ewald> {
    detail::ErfcFamilySrf<detail::kEwaldPlain> plain_;
    detail::EwaldScreenedSrf screened_;
    bool is_screened_ = false;
    template <int K> double s(double q) const {
        double v[4];
        if (is_screened_) screened_.eval<K>(q, v, nullptr);
        else plain_.eval<K>(q, v, nullptr);
        return v[K];
    }

  public:
    inline Ewald(double cutoff, double alpha, double surface_dielectric_constant = infinity, double debye_length = infinity) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_ewald(cutoff, alpha, surface_dielectric_constant, debye_length, d), "Ewald: invalid parameters");
        adopt(d);
        is_screened_ = d.zeta != 0.0;
        plain_ = detail::ErfcFamilySrf<detail::kEwaldPlain>(d);
        screened_ = detail::EwaldScreenedSrf(d);
        name = "Ewald real-space";
        doi = "10.1002/andp.19213690304";
    }
    double short_range_function(double q) const override { return s<0>(q); }
    double short_range_function_derivative(double q) const override { return s<1>(q); }
    double short_range_function_second_derivative(double q) const override { return s<2>(q); }
    double short_range_function_third_derivative(double q) const override { return s<3>(q); }
};

} // namespace CoulombGalore
