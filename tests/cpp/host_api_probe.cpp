// tests/cpp/host_api_probe.cpp -- exercises the header-only scheme classes the way reference user code does
// (cf. reference test/example.cpp:29-52, test/unittests.cpp) and exposes the results through a tiny C interface for
// pytest: every per-pair function of one pair, S..S''' on a grid, and the O(N) terms.  Built by tests/test_host_api.py
// with g++ -std=c++17 -Iinclude (no CUDA needed); a second build adds -Ioracle/eigen_shim to cover the Eigen path.
#include <coulombgalore/all.hpp>

#include <cstring>
#include <memory>

using namespace CoulombGalore;

namespace {
std::shared_ptr<SchemeBase> make(int scheme, int splined, double cutoff, double alpha, double eps_sur, double debye_length, double eps_rf,
                                 double eps_r, int shifted, int C, int D, int order, double utol) {
    if (splined) {
        auto s = std::make_shared<Splined>();
        if (utol > 0) s->setTolerance(utol);
        switch (scheme) {
        case CG_PLAIN: s->spline<Plain>(debye_length); break;
        case CG_EWALD: s->spline<Ewald>(cutoff, alpha, eps_sur, debye_length); break;
        case CG_EWALDT: s->spline<EwaldTruncated>(cutoff, alpha, eps_sur, debye_length); break;
        case CG_REACTIONFIELD: s->spline<ReactionField>(cutoff, eps_rf, eps_r, shifted != 0); break;
        case CG_WOLF: s->spline<Wolf>(cutoff, alpha); break;
        case CG_POISSON: s->spline<Poisson>(cutoff, C, D, debye_length); break;
        case CG_QPOTENTIAL: s->spline<qPotential>(cutoff, order); break;
        case CG_FANOURGAKIS: s->spline<Fanourgakis>(cutoff); break;
        case CG_ZERODIPOLE: s->spline<ZeroDipole>(cutoff, alpha); break;
        case CG_ZAHN: s->spline<Zahn>(cutoff, alpha); break;
        case CG_FENNELL: s->spline<Fennell>(cutoff, alpha); break;
        case CG_QPOTENTIAL5: s->spline<qPotentialFixedOrder<5>>(cutoff); break;
        default: return nullptr;
        }
        return s;
    }
    switch (scheme) {
    case CG_PLAIN: return std::make_shared<Plain>(debye_length);
    case CG_EWALD: return std::make_shared<Ewald>(cutoff, alpha, eps_sur, debye_length);
    case CG_EWALDT: return std::make_shared<EwaldTruncated>(cutoff, alpha, eps_sur, debye_length);
    case CG_REACTIONFIELD: return std::make_shared<ReactionField>(cutoff, eps_rf, eps_r, shifted != 0);
    case CG_WOLF: return std::make_shared<Wolf>(cutoff, alpha);
    case CG_POISSON: return std::make_shared<Poisson>(cutoff, C, D, debye_length);
    case CG_QPOTENTIAL: return std::make_shared<qPotential>(cutoff, order);
    case CG_FANOURGAKIS: return std::make_shared<Fanourgakis>(cutoff);
    case CG_ZERODIPOLE: return std::make_shared<ZeroDipole>(cutoff, alpha);
    case CG_ZAHN: return std::make_shared<Zahn>(cutoff, alpha);
    case CG_FENNELL: return std::make_shared<Fennell>(cutoff, alpha);
    case CG_QPOTENTIAL5: return std::make_shared<qPotentialFixedOrder<5>>(cutoff);
    default: return nullptr;
    }
}
mat33 mat(const double *m) {
    mat33 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r(i, j) = m ? m[3 * i + j] : 0.0;
    return r;
}
void put(double *o, const vec3 &v) {
    o[0] = v[0];
    o[1] = v[1];
    o[2] = v[2];
}
} // namespace

extern "C" {
// createScheme(type, {key: value}) -> 0 and the scheme, 1 = null pointer, 2 = unknown keyword, 3 = missing key
int probe_factory(const char *type, int n, const char *const *keys, const double *values, void **out) {
    std::map<std::string, double> j;
    for (int i = 0; i < n; i++) j[keys[i]] = values[i];
    try {
        auto s = createScheme(type, j);
        if (!s) return 1;
        *out = new std::shared_ptr<SchemeBase>(s);
        return 0;
    } catch (const std::out_of_range &) {
        return 3;
    } catch (const std::runtime_error &) {
        return 2;
    }
}


const char *probe_vector_backend() {
#ifdef COULOMBGALORE_HAVE_EIGEN
    return "eigen";
#else
    return "builtin";
#endif
}

// returns 0, or -1 when the constructor throws (as the reference constructors do on invalid parameters)
int probe_create(int scheme, int splined, double cutoff, double alpha, double eps_sur, double debye_length, double eps_rf, double eps_r,
                 int shifted, int C, int D, int order, double utol, void **out) {
    try {
        auto p = make(scheme, splined, cutoff, alpha, eps_sur, debye_length, eps_rf, eps_r, shifted, C, D, order, utol);
        if (!p) return -2;
        *out = new std::shared_ptr<SchemeBase>(p);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
void probe_destroy(void *h) { delete static_cast<std::shared_ptr<SchemeBase> *>(h); }

// same slots as cgo_info (oracle/cgo_api.h); slot 8 = descriptor().functor
int probe_info(void *h, double *o) {
    SchemeBase &s = **static_cast<std::shared_ptr<SchemeBase> *>(h);
    o[0] = s.cutoff;
    o[1] = s.debye_length;
    o[2] = s.selfEnergyFunctor({1.0, 0.0});
    o[3] = s.selfEnergyFunctor({0.0, 1.0});
    o[4] = s.selfFieldFunctor({vec3(0, 0, 0), vec3(1, 0, 0)})[0];
    o[5] = s.calc_dielectric(1.0);
    o[6] = s.neutralization_energy({1.0}, 1.0);
    o[7] = (double)(int)s.scheme;
    o[8] = (double)s.descriptor().functor;
    o[9] = 0.0;
    return 0;
}
int probe_name(void *h, char *name, int n) {
    SchemeBase &s = **static_cast<std::shared_ptr<SchemeBase> *>(h);
    std::strncpy(name, s.name.c_str(), (size_t)n - 1);
    name[n - 1] = 0;
    return 0;
}
int probe_srf(void *h, int64_t n, const double *q, double *out) {
    SchemeBase &s = **static_cast<std::shared_ptr<SchemeBase> *>(h);
    for (int64_t i = 0; i < n; i++) {
        out[0 * n + i] = s.short_range_function(q[i]);
        out[1 * n + i] = s.short_range_function_derivative(q[i]);
        out[2 * n + i] = s.short_range_function_second_derivative(q[i]);
        out[3 * n + i] = s.short_range_function_third_derivative(q[i]);
    }
    return 0;
}
// same output layout as cgo_pair (oracle/cgo_api.h), through the virtual SchemeBase interface
int probe_pair(void *h, double zA, double zB, const double *a, const double *b, const double *qa, const double *qb, const double *rr, double *o) {
    SchemeBase &s = **static_cast<std::shared_ptr<SchemeBase> *>(h);
    const vec3 muA(a[0], a[1], a[2]), muB(b[0], b[1], b[2]), r(rr[0], rr[1], rr[2]);
    const mat33 quadA = mat(qa), quadB = mat(qb);
    const double r1 = r.norm();
    o[0] = s.ion_potential(zA, r1);
    o[1] = s.dipole_potential(muA, r);
    o[2] = s.quadrupole_potential(quadA, r);
    put(o + 3, s.ion_field(zA, r));
    put(o + 6, s.dipole_field(muA, r));
    put(o + 9, s.quadrupole_field(quadA, r));
    put(o + 12, s.multipole_field(zA, muA, quadA, r));
    o[15] = s.ion_ion_energy(zA, zB, r1);
    o[16] = s.ion_dipole_energy(zA, muB, r);
    o[17] = s.dipole_dipole_energy(muA, muB, r);
    o[18] = s.ion_quadrupole_energy(zB, quadA, r);
    o[19] = s.multipole_multipole_energy(zA, zB, muA, muB, quadA, quadB, r);
    put(o + 20, s.ion_ion_force(zA, zB, r));
    put(o + 23, s.ion_dipole_force(zB, muA, r));
    put(o + 26, s.dipole_dipole_force(muA, muB, r));
    put(o + 29, s.ion_quadrupole_force(zB, quadA, r));
    put(o + 32, s.multipole_multipole_force(zA, zB, muA, muB, quadA, quadB, r));
    return 0;
}
// the q-Pochhammer free functions of the reference (qpotential.hpp:50, 66, 97, 138)
void probe_qpochhammer(double q, int l, int P, double *o) {
    o[0] = qPochhammerSymbol(q, l, P);
    o[1] = qPochhammerSymbolDerivative(q, l, P);
    o[2] = qPochhammerSymbolSecondDerivative(q, l, P);
    o[3] = qPochhammerSymbolThirdDerivative(q, l, P);
}
// statically typed use, as in the reference's README: no virtual dispatch
double probe_static_wolf_energy(double cutoff, double alpha, double zA, double zB, double r) {
    Wolf pot(cutoff, alpha);
    return pot.ion_ion_energy(zA, zB, r) + pot.self_energy({zA * zB, 0.0}) + pot.dipole_torque(vec3(1, 0, 0), vec3(0, 1, 0))[2] * 0.0;
}
}
