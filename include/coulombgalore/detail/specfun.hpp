// coulombgalore/detail/specfun.hpp -- the transcendental functions of the pair path.
//
// Host: glibc libm, exactly the calls the reference makes (std::erfc, std::exp, std::sqrt).
//
// Device (sm_100a): the FP64 pipe issues one warp instruction every two cycles, so the pair kernels are bound by
// the NUMBER of FP64 instructions per pair.  CUDA's libm erfc + exp cost ~85 FP64 instructions plus ~70 constant
// moves per call; the versions below cost 28 FP64 instructions and one 16-byte shared-memory load for BOTH
// erfc(x) and exp(-x^2):
//   * a table {erfc(x_k), exp(-x_k^2)} at x_k = k/128, k < 1024 (16 KB, staged in shared memory by the kernels,
//     built on the host with glibc -- include/coulombgalore/detail/erfc_table.hpp);
//   * exp(-x^2) = exp(-x_k^2) exp(-w), w = (x - x_k)(x + x_k), |w| <= 1/16: degree-7 Taylor polynomial;
//   * erfc(x) = erfc(x_k) - 2/sqrt(pi) exp(-x_k^2) I,  I = int_0^d exp(-2 x_k t - t^2) dt, d = x - x_k, by the
//     two-point Hermite (osculatory) rule  I = d/2 (f0+f1) + d^2/10 (f0'-f1') + d^3/120 (f0''+f1'') + d^7 f^(6)/100800,
//     whose end-point values all follow from the exp(-w) just computed: f1 = e, f' = -(2 x_k + 2 t) f, ...
// Measured against mpmath (tests/test_specfun.py, on the GPU): <= 4 ulp for x < 4, <= 1e-14 relative for x < 6, <= 1e-13 up to
// x = 7.99, i.e. at least an order inside the 1e-12 parity budget where the values matter at all (erfc(6) = 2e-17).  From
// x = 8 - 1/256 on both functions are only known to be < 2e-28 in magnitude.
#pragma once
#include "hd.hpp"

namespace CoulombGalore {
namespace detail {

constexpr int kErfcTableEntries = 1024;      // x_k = k / 128, x < 8
constexpr double kErfcTableInvStep = 128.0;
constexpr double kErfcTableStep = 1.0 / 128.0;

#ifdef __CUDACC__
// Polynomial coefficients live in constant memory: an FP64 instruction can take a constant-bank operand directly,
// whereas a 64-bit literal costs two extra uniform moves per use (measured: 12 % of all issued instructions).
static __device__ __constant__ double cg_specfun_c[10] = {-1.0 / 5040.0, 1.0 / 720.0, -1.0 / 120.0, 1.0 / 24.0,
                                                          -1.0 / 6.0,    1.0 / 6.0,   0.4,          -0.56418958354775628695, // -1/sqrt(pi)
                                                          0.375,         128.0}; // Halley coefficient, table steps per unit
#endif

// What the device erfc needs besides its argument: the shared-memory address of the table and the ten coefficients of
// cg_specfun_c.  The pair kernel builds one ErfcCtx per thread with the coefficients held in registers: left to itself the
// compiler re-loads them from the constant bank in every iteration of the pair loop, and a constant load that misses the
// small per-SM constant cache costs as much as a global load.
struct ErfcCtx {
    unsigned tab;
    double c[10];
};
// The table as the cell-list pair kernel addresses it: entry k at shared address tab + stride k, k clamped to kmax.  Eight
// interleaved copies (stride 128, tab = base + 16 (lane & 7)) put the eight lanes of a quarter-warp on eight different
// 16-byte bank groups whatever their arguments: four wavefronts per warp-wide lookup where a single copy costs about nine
// (random bank conflicts).  Schemes whose argument is eta q <= eta use it (the copies hold the entries up to eta only); an
// index beyond kmax can only belong to a pair that failed the cut-off gate, whose value is discarded.
struct ErfcTabRep {
    unsigned tab, stride, kmax;
};

#ifdef __CUDA_ARCH__
// 1/sqrt(x), x > 0 normal: MUFU.RSQ64H seed (2^-22) + one third-order (Halley) step -> full double precision
__device__ __forceinline__ double rsqrt_device(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y); // 0.375, 0.5: immediates (the low word of the literal is zero)
}

// ec = erfc(x), E = exp(-x^2) for x >= 0 from the shared-memory table at address `tab` (interleaved {erfc(x_k), exp(-x_k^2)}).
// Branch-free: beyond the table (x >= 8, where erfc < 1.2e-29 and exp(-x^2) < 1.7e-28) the last entry is expanded
// around the argument's own grid point, which returns values of that same negligible magnitude.
__device__ __forceinline__ void erfc_gauss_table(double x, unsigned tab, const double (&c)[10], double &ec, double &E, unsigned stride = 16u,
                                                 unsigned kmax = (unsigned)(kErfcTableEntries - 1)) {
    const double magic = 6755399441055744.0; // 2^52 + 2^51: the low word of (x*128 + magic) is round(x*128)
    const double t = fma(x, kErfcTableInvStep, magic);
    // x >= 0: one unsigned min clamps the index (a negative low word would read as huge and clamp too)
    const int k = (int)min((unsigned)__double2loint(t), kmax);
    const double kf = t - magic;
    const double d = fma(kf, -kErfcTableStep, x); // x - x_k, exact
    const double xk = kf * kErfcTableStep;
    double2 T; // the table lives in shared memory: address it as such (a generic load would cost an aperture check)
    asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(T.x), "=d"(T.y) : "r"(tab + stride * (unsigned)k));
    const double w = d * (x + xk);
    double e = fma(w, c[0], c[1]);
    e = fma(e, w, c[2]);
    e = fma(e, w, c[3]);
    e = fma(e, w, c[4]);
    e = fma(e, w, 0.5);
    e = fma(e, w, -1.0);
    e = fma(e, w, 1.0); // exp(-w)
    // Hermite rule, with a2/2 = (x_k^2 - 1/2)(1 + e) + w e  (x^2 = x_k^2 + w) and the factor 1/2 moved into the constant:
    //   2 I = d [ a0 + 0.4 d ( a1 + d/6 * a2/2 ) ],  a0 = 1 + e,  a1 = x e - x_k
    const double a0 = 1.0 + e;
    const double a1 = fma(x, e, -xk);
    const double a2h = fma(fma(xk, xk, -0.5), a0, w * e);
    const double u = fma(d * c[5], a2h, a1);
    const double I2 = d * fma(d * c[6], u, a0);
    E = T.y * e;
    ec = fma(c[7] * T.y, I2, T.x);
}
__device__ __forceinline__ void erfc_gauss_table(double x, const ErfcCtx &cx, double &ec, double &E) { erfc_gauss_table(x, cx.tab, cx.c, ec, E); }
__device__ __forceinline__ void erfc_gauss_table(double x, const ErfcTabRep &r, double &ec, double &E) {
    erfc_gauss_table(x, r.tab, cg_specfun_c, ec, E, r.stride, r.kmax);
}
__device__ __forceinline__ void erfc_gauss_table(double x, const double *tab, double &ec, double &E) {
    erfc_gauss_table(x, (unsigned)__cvta_generic_to_shared(tab), cg_specfun_c, ec, E);
}
#endif

// a double from a table that lives in shared memory on the device (plain memory on the host)
CG_HD double table_load(const double *p) {
#ifdef __CUDA_ARCH__
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
#else
    return *p;
#endif
}

// 1/sqrt(x) for x > 0
CG_HD double rsqrt_pos(double x) {
#ifdef __CUDA_ARCH__
    return rsqrt_device(x);
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

// ec = erfc(x), E = exp(-x^2); `tab` is the device erfc table -- a pointer into shared memory or an ErfcCtx -- and is
// ignored on the host.  SIGNED: x may be negative (screened Ewald at small q): erfc(-x) = 2 - erfc(x).
template <bool SIGNED = false, class TAB = const double *> CG_HD void erfc_and_gauss(double x, double &ec, double &E, const TAB &tab) {
#ifdef __CUDA_ARCH__
    if (SIGNED) {
        erfc_gauss_table(fabs(x), tab, ec, E); // the kernels always stage the table
        ec = x < 0.0 ? 2.0 - ec : ec;
    } else {
        erfc_gauss_table(x, tab, ec, E);
    }
#else
    (void)tab;
    E = std::exp(-x * x);
    ec = std::erfc(x);
#endif
}

// ---- fp32 variants (CG_FP32 kernels): CUDA / libm single-precision functions, no tables ----
CG_HD float rsqrt_pos(float x) {
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / std::sqrt(x);
#endif
}
CG_HD float exp_any(float y) {
#ifdef __CUDA_ARCH__
    return expf(y);
#else
    return std::exp(y);
#endif
}
template <bool SIGNED = false, class TAB = const double *> CG_HD void erfc_and_gauss(float x, float &ec, float &E, const TAB &) {
#ifdef __CUDA_ARCH__
    E = expf(-x * x);
    ec = erfcf(x);
#else
    E = std::exp(-x * x);
    ec = std::erfc(x);
#endif
}

} // namespace detail
} // namespace CoulombGalore
