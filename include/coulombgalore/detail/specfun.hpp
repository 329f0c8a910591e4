// coulombgalore/detail/specfun.hpp -- the transcendental functions of the pair path.
//
// Host: glibc libm, exactly the calls the reference makes (std::erfc, std::exp, std::sqrt).
// Device: CUDA's double-precision erfc/erfcx/exp/rsqrt (<= 2-5 ulp), far inside the 1e-12 parity budget.
#pragma once
#include "hd.hpp"

namespace CoulombGalore {
namespace detail {

// 1/sqrt(x) for x > 0
CG_HD double rsqrt_pos(double x) {
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / std::sqrt(x);
#endif
}

// exp(y), any sign
CG_HD double exp_any(double y) {
#ifdef __CUDA_ARCH__
    return exp(y);
#else
    return std::exp(y);
#endif
}

// ec = erfc(x), E = exp(-x^2); x may be negative (screened Ewald at small q)
CG_HD void erfc_and_gauss(double x, double &ec, double &E) {
#ifdef __CUDA_ARCH__
    E = exp(-x * x);
    ec = erfc(x);
#else
    E = std::exp(-x * x);
    ec = std::erfc(x);
#endif
}

// erfc(xp) * exp(a) where exp(a) = exp(xp^2) * Em, i.e. erfcx(xp) * Em (Em = exp(-xm^2), xp^2 - xm^2 = a)
CG_HD double erfc_times_exp(double xp, double a, double Em) {
#ifdef __CUDA_ARCH__
    (void)a;
    return erfcx(xp) * Em;
#else
    (void)Em;
    return std::erfc(xp) * std::exp(a);
#endif
}

} // namespace detail
} // namespace CoulombGalore
