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

// ---- reciprocal-space Ewald: the reference's own ReciprocalEwaldGaussian (ewald.hpp:214-385) ----
namespace {
struct RefRecip : cgo::RecipIface {
    std::unique_ptr<ReciprocalEwaldGaussian> pot;
    struct KAccess : ReciprocalEwaldGaussian { // k_vectors / Aks / Q_mp are protected members of the state class
        using ReciprocalEwaldGaussian::Aks;
        using ReciprocalEwaldGaussian::k_vectors;
        using ReciprocalEwaldGaussian::Q_mp;
    };
    const KAccess &acc() const { return static_cast<const KAccess &>(*pot); }
    RefRecip(double cutoff, double alpha, double eps_sur, double kappa, int kc, const double *box) {
        ReciprocalEwaldState st;
        st.box_length = {box[0], box[1], box[2]};
        st.cutoff = cutoff;
        st.reciprocal_cutoff = kc;
        st.alpha = alpha;
        st.kappa = kappa;
        st.surface_dielectric_constant = eps_sur;
        pot.reset(new ReciprocalEwaldGaussian(st));
    }
    static void load(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx, const double *my,
                     const double *mz, std::vector<vec3> &pos, std::vector<double> &ch, std::vector<vec3> &dip) {
        pos.resize((size_t)N);
        ch.resize((size_t)N);
        dip.resize((size_t)N);
        for (int64_t i = 0; i < N; i++) {
            pos[(size_t)i] = vec3(x[i], y[i], z[i]);
            ch[(size_t)i] = q ? q[i] : 0.0;
            dip[(size_t)i] = mx ? vec3(mx[i], my[i], mz[i]) : vec3(0, 0, 0);
        }
    }
    int numk() override { return pot->numKVectors(); }
    int kvectors(double *k3, double *aks) override {
        for (int i = 0; i < pot->numKVectors(); i++) {
            const vec3 k = acc().k_vectors.col(i);
            if (k3)
                for (int d = 0; d < 3; d++) k3[3 * i + d] = k[d];
            if (aks) aks[i] = acc().Aks[i];
        }
        return 0;
    }
    int update(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx, const double *my,
               const double *mz, int sign) override {
        std::vector<vec3> pos, dip;
        std::vector<double> ch;
        load(N, x, y, z, q, mx, my, mz, pos, ch, dip);
        if (sign == 0) pot->setZeroComplex();
        if (sign < 0) pot->updateComplex(pos, ch, dip, std::minus<>());
        else pot->updateComplex(pos, ch, dip);
        return 0;
    }
    int getq(double *q2k) override {
        for (int i = 0; i < pot->numKVectors(); i++) {
            q2k[2 * i] = acc().Q_mp[i].real();
            q2k[2 * i + 1] = acc().Q_mp[i].imag();
        }
        return 0;
    }
    int energy(double *e) override {
        *e = pot->reciprocal_energy();
        return 0;
    }
    int force_field(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx, const double *my,
                    const double *mz, double *force, double *field, double *fs) override {
        const double pi = 4.0 * std::atan(1.0);
        for (int64_t i = 0; i < N; i++) {
            const vec3 r(x[i], y[i], z[i]), mu = mx ? vec3(mx[i], my[i], mz[i]) : vec3(0, 0, 0);
            const double ch = q ? q[i] : 0.0;
            if (force) {
                const vec3 f = pot->reciprocal_force(r, ch, mu);
                for (int d = 0; d < 3; d++) force[3 * i + d] = f[d];
            }
            if (field) {
                const vec3 f = pot->reciprocal_field(r);
                for (int d = 0; d < 3; d++) field[3 * i + d] = f[d];
            }
            if (fs) { // sum_k |term_k| of reciprocal_force: the scale rounding errors are measured against
                long double s = 0;
                for (int k = 0; k < pot->numKVectors(); k++) {
                    const vec3 kv = acc().k_vectors.col(k);
                    s += (std::fabs(mu.dot(kv)) + std::fabs(ch)) * std::abs(acc().Q_mp[k]) * kv.norm() * acc().Aks[k];
                }
                fs[i] = (double)(4.0 * pi / pot->getVolume() * s);
            }
        }
        return 0;
    }
    int surface(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx, const double *my,
                const double *mz, double *energy, double *field3) override {
        std::vector<vec3> pos, dip;
        std::vector<double> ch;
        load(N, x, y, z, q, mx, my, mz, pos, ch, dip);
        if (energy) *energy = pot->surface_energy(pos, ch, dip);
        if (field3) {
            const vec3 f = pot->surface_field(pos, ch, dip);
            for (int d = 0; d < 3; d++) field3[d] = f[d];
        }
        return 0;
    }
};
} // namespace
cgo::RecipIface *cgo::make_recip(double cutoff, double alpha, double eps_sur, double kappa, int kc, const double *box) {
    return new RefRecip(cutoff, alpha, eps_sur, kappa, kc, box);
}

CGO_DEFINE_C_ABI("reference")

extern "C" {
int cgo_ewaldt_surface(double cutoff, double alpha, double eps_sur, int64_t N, const double *x, const double *y, const double *z,
                       const double *q, const double *mx, const double *my, const double *mz, double volume, double *energy) {
    try {
        const EwaldTruncated pot(cutoff, alpha, eps_sur);
        std::vector<vec3> pos, dip;
        std::vector<double> ch;
        RefRecip::load(N, x, y, z, q, mx, my, mz, pos, ch, dip);
        *energy = pot.surface_energy(pos, ch, dip, volume);
        return 0;
    } catch (...) {
        return -1;
    }
}
int cgo_qpochhammer(double q, int32_t l, int32_t P, double *out4) {
    out4[0] = qPochhammerSymbol(q, l, P);
    out4[1] = qPochhammerSymbolDerivative(q, l, P);
    out4[2] = qPochhammerSymbolSecondDerivative(q, l, P);
    out4[3] = qPochhammerSymbolThirdDerivative(q, l, P);
    return 0;
}
}
