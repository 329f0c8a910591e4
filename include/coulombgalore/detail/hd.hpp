// coulombgalore/detail/hd.hpp -- host/device portability layer of the B200 CoulombGalore engine.
//
// Everything under detail/ compiles both as plain C++17 (g++, for the header-only scheme classes that mirror the
// reference's per-pair API on the host) and as CUDA device code (nvcc, for the sm_100a kernels).  One source of
// truth for the arithmetic keeps the host classes and the kernels consistent by construction.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

#if defined(__CUDACC__)
#define CG_HD __host__ __device__ __forceinline__
#define CG_HD_NOINLINE __host__ __device__
#else
#define CG_HD inline
#define CG_HD_NOINLINE inline
#endif

namespace CoulombGalore {
namespace detail {

constexpr double kInfinity = std::numeric_limits<double>::infinity();

// integer power by repeated squaring; negative exponents give the reciprocal.  Same multiplication order as
// libgcc's __powidf2, which is what the reference's powi() resolves to on GCC (reference core.hpp:77-83).
template <class T> CG_HD T powi(T x, int m) {
    unsigned int n = m < 0 ? 0u - (unsigned int)m : (unsigned int)m;
    T y = (n & 1u) ? x : T(1);
    while (n >>= 1) {
        x = x * x;
        if (n & 1u) y *= x;
    }
    return m < 0 ? T(1) / y : y;
}

// 32-bit unsigned factorial/binomial: overflow for n >= 13 is part of the reference's behaviour
// (reference core.hpp:89,97), the Poisson constructor therefore rejects C + D > 12 here.
inline constexpr unsigned int factorial(unsigned int n) { return n <= 1 ? 1 : n * factorial(n - 1); }
inline constexpr unsigned int binomial(signed int n, signed int k) {
    return factorial((unsigned int)n) / (factorial((unsigned int)k) * factorial((unsigned int)(n - k)));
}

template <class T> struct Vec3T {
    T x, y, z;
};
typedef Vec3T<double> D3;
template <class T> CG_HD T dot(const Vec3T<T> &a, const Vec3T<T> &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

} // namespace detail
} // namespace CoulombGalore
