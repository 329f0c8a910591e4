// oracle/ref_main.cpp -- TEST INFRASTRUCTURE ONLY.
// Builds oracle/_ref/libcg_ref.so: the UNMODIFIED reference headers (/root/reference/include/coulombgalore/*.hpp,
// included from where they lie -- nothing is copied) behind the cgo_api.h C ABI, with oracle/eigen_shim standing
// in for the Eigen 3.4.0 dependency the reference fetches over the network (reference CMakeLists.txt:28-29).
// Flags (oracle/Makefile): -std=c++17 -O3 -DNDEBUG -ffp-contract=off, no -march, no -ffast-math.
#include <coulombgalore/all.hpp>

#include "harness.hpp"

using namespace CoulombGalore;

namespace {

typedef Tabulate::Andrea<double> Andrea;
typedef Tabulate::TabulatorBase<double>::data TableData;

// Splined keeps its tables private (splined.hpp:325).  Tabulate::Andrea is public, so the same tables are
// regenerated here by the calls generate_spline_data makes (splined.hpp:330-336), in this same binary.
template <class T> struct SplinedImpl : cgo::Impl<Splined, vec3, mat33> {
    std::shared_ptr<T> inner;
    TableData tables[4];
    template <class... A> explicit SplinedImpl(double utol, A... a) {
        if (utol > 0) pot.setTolerance(utol);
        pot.template spline<T>(a...);
        inner = std::make_shared<T>(a...);
        Andrea tab;
        tab.setTolerance(utol > 0 ? utol : 1e-3);
        auto p = inner;
        tables[0] = tab.generate([p](double q) { return p->short_range_function(q); }, 0, 1);
        tables[1] = tab.generate([p](double q) { return p->short_range_function_derivative(q); }, 0, 1);
        tables[2] = tab.generate([p](double q) { return p->short_range_function_second_derivative(q); }, 0, 1);
        tables[3] = tab.generate([p](double q) { return p->short_range_function_third_derivative(q); }, 0, 1);
    }
    int spline_table(int k, int32_t *n, double *r2, double *c) override {
        if (k < 0 || k > 3) return -1;
        const TableData &t = tables[k];
        if (n) *n = (int32_t)t.r2.size();
        if (r2) std::copy(t.r2.begin(), t.r2.end(), r2);
        if (c) std::copy(t.c.begin(), t.c.end(), c);
        return 0;
    }
};

template <class T, class... A> cgo::Iface *build(const cgo_params &p, A... a) {
    if (p.splined) return new SplinedImpl<T>(p.spline_utol, a...);
    return new cgo::Impl<T, vec3, mat33>(a...);
}

} // namespace

cgo::Iface *cgo::make(const cgo_params &p) {
    switch (p.scheme) {
    case CGO_PLAIN: return build<Plain>(p, p.debye_length);
    case CGO_EWALD: return build<Ewald>(p, p.cutoff, p.alpha, p.eps_sur, p.debye_length);
    case CGO_EWALDT: return build<EwaldTruncated>(p, p.cutoff, p.alpha, p.eps_sur, p.debye_length);
    case CGO_REACTIONFIELD: return build<ReactionField>(p, p.cutoff, p.eps_rf, p.eps_r, p.shifted != 0);
    case CGO_WOLF: return build<Wolf>(p, p.cutoff, p.alpha);
    case CGO_POISSON: return build<Poisson>(p, p.cutoff, (signed int)p.C, (signed int)p.D, p.debye_length);
    case CGO_QPOTENTIAL: return build<qPotential>(p, p.cutoff, (int)p.order);
    case CGO_FANOURGAKIS: return build<Fanourgakis>(p, p.cutoff);
    case CGO_ZERODIPOLE: return build<ZeroDipole>(p, p.cutoff, p.alpha);
    case CGO_ZAHN: return build<Zahn>(p, p.cutoff, p.alpha);
    case CGO_FENNELL: return build<Fennell>(p, p.cutoff, p.alpha);
    case CGO_QPOTENTIAL5: return build<qPotentialFixedOrder<5>>(p, p.cutoff);
    default: return nullptr;
    }
}

CGO_DEFINE_C_ABI("reference")
