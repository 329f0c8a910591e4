// coulombgalore/detail/pair.hpp -- the pair-interaction formulas of reference core.hpp:268-779, restated once for
// host and device on raw doubles.
//
// Every formula of the reference is of the form  (geometry) x (radial factor built from S..S''' and kappa r) x
// exp(-kappa r).  `radial_terms` evaluates the radial factors once per pair:
//     angcor = S (1 + kr) - q S'                                  (core.hpp:596, 754;  potential/field of an ion)
//     unicor = (S kr - 2 q S') kr / 3 + q^2 S'' / 3               (core.hpp:597, 755)
//     totcor = angcor + unicor                                    (the dipole "A" factor, core.hpp:393)
//     r3corr = angcor kr^2 - 2 q S' (1+kr) kr + q^2 S'' (1+3kr) - q^3 S'''     (core.hpp:696-697, 757-758)
// and the per-mode functions below combine them with the geometry exactly as the reference does, with the
// divisions replaced by multiplications with 1/r (one rsqrt per pair).  YUK=false removes every kappa term at
// compile time (kappa = 0 gives bit-identical values to evaluating them: exp(-0) = 1, kr = 0).
#pragma once
#include "hd.hpp"
#include "specfun.hpp"
#include "srf.hpp"

namespace CoulombGalore {
namespace detail {

enum { kWhatEnergy = 1, kWhatForce = 2, kWhatField = 4, kWhatPotential = 8 };
enum { kModeIon = 0, kModeDipole = 1, kModeMultipole = 2, kModeMultipoleQuad = 3 };

struct CoreParams {
    double inverse_cutoff, cutoff_squared, kappa; // core.hpp:116-118
};

// highest S-derivative a (mode, what) combination needs
constexpr int derivatives_needed(int mode, int what) {
    return mode == kModeMultipoleQuad ? ((what & (kWhatForce | kWhatField)) ? 3 : 2) // quadrupole field / force need r3corr
           : mode == kModeIon ? ((what & (kWhatForce | kWhatField)) ? 1 : 0)
                            : ((what & kWhatForce) ? 3 : ((what & (kWhatField | kWhatEnergy)) ? 2 : 1));
}

template <class T> struct RadialT {
    T rinv, r1, q, s0, ek, angcor, unicor, totcor, r3corr;
};
typedef RadialT<double> Radial;

// S..S^(ND) through the functor; the erfc family takes the table argument as it comes (pointer or ErfcCtx)
template <int ND, class T, class F, class TAB> CG_HD void eval_srf(const F &f, T q, T *s, const TAB &tab) {
    if constexpr (F::kTable == kTableErfc) f.template eval<ND, T, TAB>(q, s, tab);
    else f.template eval<ND, T>(q, s, tab);
}

// ANGCOR_ONLY: the caller needs nothing but angcor (ion field / force without screening) and the functor has a closed form.
// T: arithmetic type of the pair math (double; float for the CG_FP32 kernels)
// TAB: what the functor's table argument is (a pointer, nullptr, or the kernels' ErfcCtx)
template <class F, bool YUK, int ND, bool ANGCOR_ONLY = false, class T = double, class TAB = const double *>
CG_HD RadialT<T> radial_terms(const F &f, const CoreParams &cp, T r2, const TAB &tab) {
    RadialT<T> R;
    R.rinv = rsqrt_pos(r2);
    R.r1 = r2 * R.rinv;
    R.q = R.r1 * (T)cp.inverse_cutoff;
    if constexpr (ANGCOR_ONLY) {
        R.s0 = T(0);
        R.ek = T(1);
        R.angcor = f.template angcor<T, TAB>(R.q, tab);
        R.unicor = R.totcor = R.r3corr = T(0);
        return R;
    } else {
        T s[4] = {T(0), T(0), T(0), T(0)};
        eval_srf<ND, T>(f, R.q, s, tab);
        R.s0 = s[0];
        const T third = T(1.0 / 3.0);
        const T qs1 = R.q * s[1];
        const T q2s2 = (ND >= 2) ? (R.q * R.q) * s[2] : T(0);
        const T q3s3 = (ND >= 3) ? (R.q * R.q) * (R.q * s[3]) : T(0);
        if (YUK) {
            const T kr = (T)cp.kappa * R.r1;
            R.ek = exp_any(-kr);
            R.angcor = s[0] * (T(1) + kr) - qs1;
            R.unicor = (s[0] * kr - T(2) * qs1) * kr * third + q2s2 * third;
            R.totcor = R.angcor + R.unicor;
            R.r3corr = R.angcor * kr * kr - qs1 * T(2) * (T(1) + kr) * kr + q2s2 * (T(1) + T(3) * kr) - q3s3;
        } else {
            R.ek = T(1);
            R.angcor = s[0] - qs1;
            R.unicor = q2s2 * third;
            R.totcor = R.angcor + R.unicor;
            R.r3corr = q2s2 - q3s3;
        }
        return R;
    }
}

// per-particle accumulators; members that a (mode, what) instantiation never touches stay dead and cost nothing
template <class T> struct AccumT {
    T U = T(0), Fx = T(0), Fy = T(0), Fz = T(0), Ex = T(0), Ey = T(0), Ez = T(0), phi = T(0);
};
typedef AccumT<double> Accum;

// A quadrupole tensor Q as the pair formulas use it: only r^T Q r, (Q + Q^T) r and trace(Q) occur
// (core.hpp:320-336, 412-436, 603-609, 767-773), so the symmetrised S = Q + Q^T and the trace are stored.
template <class T> struct QuadT {
    T xx, yy, zz, tr; // S_xx, S_yy, S_zz, trace(Q)
    T xy, xz, yz;     // S_xy, S_xz, S_yz
};
template <class T> CG_HD Vec3T<T> quad_times(const QuadT<T> &s, const Vec3T<T> &d) { // (Q + Q^T) d
    return {s.xx * d.x + s.xy * d.y + s.xz * d.z, s.xy * d.x + s.yy * d.y + s.yz * d.z, s.xz * d.x + s.yz * d.y + s.zz * d.z};
}

// One in-cutoff ordered pair (i <- j), d = r_i - r_j, r2 = |d|^2 < cutoff^2 already established by the caller.
// Conventions: SURVEY.md 8b / oracle/harness.hpp.
// `pass` = the pair is inside the cut-off.  The kernels evaluate branch-free (the radial arithmetic must then stay
// finite for rejected / absent pairs): every contribution carries a scalar factor that is zeroed by a select.
template <class F, bool YUK, int MODE, int WHAT, class T = double, class TAB = const double *>
CG_HD void pair_accumulate(const F &f, const CoreParams &cp, const TAB &tab, const Vec3T<T> &d, T r2, T zi, T zj,
                           const Vec3T<T> &mui, const Vec3T<T> &muj, AccumT<T> &acc, bool pass = true,
                           const QuadT<T> *quadi = nullptr, const QuadT<T> *quadj = nullptr) {
    constexpr int ND = derivatives_needed(MODE, WHAT);
    constexpr bool kAngcorOnly = F::kHasAngcor && !YUK && MODE == kModeIon && (WHAT & (kWhatEnergy | kWhatPotential)) == 0;
    const RadialT<T> R = radial_terms<F, YUK, ND, kAngcorOnly, T, TAB>(f, cp, r2, tab);
    const T rinv2 = R.rinv * R.rinv;
    const T rinv3 = rinv2 * R.rinv;
    const T ek3 = pass ? R.ek * rinv3 : T(0); // exp(-kr) / r^3

    if (MODE == kModeIon) {
        // ion_potential core.hpp:268-280, ion_field :352-365, ion_ion_energy :501, ion_ion_force :630
        if (WHAT & (kWhatPotential | kWhatEnergy)) {
            const T p = pass ? zj * R.rinv * R.s0 * R.ek : T(0);
            if (WHAT & kWhatPotential) acc.phi += p;
            if (WHAT & kWhatEnergy) acc.U += zi * p;
        }
        if (WHAT & (kWhatField | kWhatForce)) {
            const T c = zj * ek3 * R.angcor;
            if (WHAT & kWhatField) {
                acc.Ex += c * d.x;
                acc.Ey += c * d.y;
                acc.Ez += c * d.z;
            }
            if (WHAT & kWhatForce) {
                const T cf = zi * c;
                acc.Fx += cf * d.x;
                acc.Fy += cf * d.y;
                acc.Fz += cf * d.z;
            }
        }
        return;
    }

    const T mjd = dot(muj, d);
    if (MODE == kModeDipole) {
        // dipole_potential :294-307, dipole_field :380-399, dipole_dipole_energy :543-546, dipole_dipole_force :677-702
        if (WHAT & kWhatPotential) acc.phi += mjd * ek3 * R.angcor;
        // Shared factors: Z = angcor e / r^3 and B = 3 totcor e / r^5.  The field is  B (mu_j.d) d - Z mu_j  (B - A = -angcor), the
        // energy  -mu_i.E = Z (mu_i.mu_j) - B (mu_i.d)(mu_j.d), and B (mu_j.d), B (mu_i.d) are also the force's coefficients of
        // mu_i and mu_j: every product is formed once and every sum is an FMA into its accumulator.
        const T Z = R.angcor * ek3;
        const T B = (T(3) * R.totcor) * (ek3 * rinv2);
        const T ci = B * mjd;
        if (WHAT & kWhatField) {
            acc.Ex += ci * d.x; // one FMA per statement
            acc.Ex -= Z * muj.x;
            acc.Ey += ci * d.y;
            acc.Ey -= Z * muj.y;
            acc.Ez += ci * d.z;
            acc.Ez -= Z * muj.z;
        }
        if (WHAT & (kWhatEnergy | kWhatForce)) {
            const T mid = dot(mui, d), mm = dot(mui, muj);
            if (WHAT & kWhatEnergy) {
                acc.U += Z * mm;
                acc.U -= ci * mid;
            }
            if (WHAT & kWhatForce) {
                const T ab = mid * mjd * rinv2;
                const T cj = B * mid;
                const T cd = -(B * (T(5) * ab - mm) + (ab * R.r3corr) * (ek3 * rinv2));
                acc.Fx += cd * d.x;
                acc.Fx += ci * mui.x;
                acc.Fx += cj * muj.x;
                acc.Fy += cd * d.y;
                acc.Fy += ci * mui.y;
                acc.Fy += cj * muj.y;
                acc.Fz += cd * d.z;
                acc.Fz += ci * mui.z;
                acc.Fz += cj * muj.z;
            }
        }
        return;
    }

    // MODE == kModeMultipole: potential = ion_potential + dipole_potential, field = multipole_field :457-486 (quad = 0),
    // energy = multipole_multipole_energy :581-615, force = multipole_multipole_force :737-779, both with A = i,
    // B = j, r = r_j - r_i = -d and zero quadrupoles.
    if (WHAT & kWhatPotential) acc.phi += (pass ? zj * R.rinv * R.s0 * R.ek : T(0)) + mjd * ek3 * R.angcor;
    if (WHAT & kWhatField) {
        const T cd = (zj * R.angcor + T(3) * mjd * rinv2 * R.totcor) * ek3;
        const T cm = -R.angcor * ek3;
        acc.Ex += cd * d.x + cm * muj.x;
        acc.Ey += cd * d.y + cm * muj.y;
        acc.Ez += cd * d.z + cm * muj.z;
    }
    if (WHAT & (kWhatEnergy | kWhatForce)) {
        const T mid = dot(mui, d), mm = dot(mui, muj);
        const T ab = mid * mjd * rinv2;
        if (WHAT & kWhatEnergy) {
            const T ion_ion = zi * zj * R.s0 * r2;
            const T ion_dipole = (zi * mjd - zj * mid) * R.angcor;
            const T dipole_dipole = -(T(3) * ab - mm) * R.totcor - mm * R.unicor;
            acc.U += (ion_ion + ion_dipole + dipole_dipole) * ek3;
        }
        if (WHAT & kWhatForce) {
            const T t3 = T(3) * R.totcor;
            const T e4 = ek3 * R.rinv; // exp(-kr)/r^4
            const T ar = R.angcor * R.r1;
            const T cd = (-zi * zj * ar + R.rinv * (t3 * (zi * mjd + zj * mid) - t3 * (T(5) * ab - mm) - ab * R.r3corr)) * e4;
            const T ci = (-zj * ar + t3 * mjd * R.rinv) * e4;
            const T cj = (-zi * ar + t3 * mid * R.rinv) * e4;
            acc.Fx += cd * d.x + ci * mui.x + cj * muj.x;
            acc.Fy += cd * d.y + ci * mui.y + cj * muj.y;
            acc.Fz += cd * d.z + ci * mui.z + cj * muj.z;
        }
    }
    if (MODE == kModeMultipoleQuad) {
        // the ion-quadrupole terms of quadrupole_potential :320-336, multipole_field :477-481, multipole_multipole_energy
        // :603-609 and multipole_multipole_force :767-773 (A = i, B = j, r = -d).  With qf = rhat^T Q rhat:
        //   potential   1/2 ((3 qf_j - tr_j) totcor + tr_j unicor) e^{-kr} / r^3
        //   field       1/2 ((3 (5 qf_j - tr_j) totcor + qf_j r3corr) d - 3 totcor S_j d) e^{-kr} / r^5
        //   energy      z_i * (potential term of j) + z_j * (potential term of i)
        //   force       1/2 (z_i (3 (5 qf_j - tr_j) totcor + qf_j r3corr) - z_j (3 (5 qf_i - tr_i) totcor + qf_i r3corr)) d e^{-kr} / r^5
        //               + 3/2 totcor (z_j S_i d - z_i S_j d) e^{-kr} / r^5
        const QuadT<T> &Qi = *quadi, &Qj = *quadj;
        const Vec3T<T> sdj = quad_times(Qj, d);
        const T half = T(0.5);
        const T qfj = half * dot(d, sdj) * rinv2;
        const T e5 = ek3 * rinv2; // exp(-kr) / r^5
        if (WHAT & (kWhatPotential | kWhatEnergy)) {
            const T pj = half * ((T(3) * qfj - Qj.tr) * R.totcor + Qj.tr * R.unicor) * ek3;
            if (WHAT & kWhatPotential) acc.phi += pj;
            if (WHAT & kWhatEnergy) {
                const T qfi = half * dot(d, quad_times(Qi, d)) * rinv2;
                acc.U += zi * pj + zj * half * ((T(3) * qfi - Qi.tr) * R.totcor + Qi.tr * R.unicor) * ek3;
            }
        }
        if (WHAT & (kWhatField | kWhatForce)) {
            const T t3 = T(3) * R.totcor;
            const T gj = t3 * (T(5) * qfj - Qj.tr) + qfj * R.r3corr;
            if (WHAT & kWhatField) {
                const T cd = half * gj * e5, cs = -half * t3 * e5;
                acc.Ex += cd * d.x + cs * sdj.x;
                acc.Ey += cd * d.y + cs * sdj.y;
                acc.Ez += cd * d.z + cs * sdj.z;
            }
            if (WHAT & kWhatForce) {
                const Vec3T<T> sdi = quad_times(Qi, d);
                const T qfi = half * dot(d, sdi) * rinv2;
                const T gi = t3 * (T(5) * qfi - Qi.tr) + qfi * R.r3corr;
                const T cd = half * (zi * gj - zj * gi) * e5;
                const T ci = half * t3 * zj * e5, cj = -half * t3 * zi * e5;
                acc.Fx += cd * d.x + ci * sdi.x + cj * sdj.x;
                acc.Fy += cd * d.y + ci * sdi.y + cj * sdj.y;
                acc.Fz += cd * d.z + ci * sdi.z + cj * sdj.z;
            }
        }
    }
}

// One in-cutoff ordered pair (i <- j) for the NS schemes of an ErfcMultiSrf, ion mode (ion_potential core.hpp:268-280, ion_field
// :352-365, ion_ion_energy :501, ion_ion_force :630, all with kappa = 0): the geometry, 1/r, erfc and the Gaussian once,
// then per scheme its polynomial part and its own accumulators.
template <class F, int WHAT, class T, class TAB>
CG_HD void pair_accumulate_multi(const F &f, const CoreParams &cp, const TAB &tab, const Vec3T<T> &d, T r2, T zi, T zj,
                                 AccumT<T> (&acc)[F::kSchemes], bool pass) {
    const T rinv = rsqrt_pos(r2);
    const T r1 = r2 * rinv;
    const T q = r1 * (T)cp.inverse_cutoff;
    const T x = (T)f.eta * q;
    T E, ec;
    erfc_and_gauss(x, ec, E, tab);
    const T rinv3 = rinv * rinv * rinv;
    const T zr1 = pass ? zj * rinv : T(0);  // z_j / r
    const T zr3 = pass ? zj * rinv3 : T(0); // z_j / r^3
    const T g = ec + (q * (T)f.k1) * E;     // erfc(x) + 2 x exp(-x^2) / sqrt(pi)
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < F::kSchemes; k++) {
        if (WHAT & (kWhatPotential | kWhatEnergy)) {
            const T s0 = (T)f.m[k] * (ec + (T)f.a0[k] + q * ((T)f.a1[k] + q * ((T)f.a2[k] + q * (T)f.a3[k])));
            const T p = zr1 * s0;
            if (WHAT & kWhatPotential) acc[k].phi += p;
            if (WHAT & kWhatEnergy) acc[k].U += zi * p;
        }
        if (WHAT & (kWhatField | kWhatForce)) {
            const T ang = (T)f.m[k] * (g + (T)f.b0[k] + (q * q) * ((T)f.b2[k] + q * (T)f.b3[k]));
            const T c = zr3 * ang;
            if (WHAT & kWhatField) {
                acc[k].Ex += c * d.x;
                acc[k].Ey += c * d.y;
                acc[k].Ez += c * d.z;
            }
            if (WHAT & kWhatForce) {
                const T cf = zi * c;
                acc[k].Fx += cf * d.x;
                acc[k].Fy += cf * d.y;
                acc[k].Fz += cf * d.z;
            }
        }
    }
}

} // namespace detail
} // namespace CoulombGalore
