// coulombgalore/splined.hpp -- Splined and Tabulate::Andrea: same class name and constructor as the reference (splined.hpp:87-386); constants and S(q) come
// from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a kernels are instantiated from.
#pragma once
#include "core.hpp"

namespace CoulombGalore {

namespace Tabulate {
/** The reference's tabulator, reduced to what Splined uses: generate(f, 0, 1) and eval (splined.hpp:167-176, 195-298) */
template <typename T = double> class Andrea {
    T utol = 1e-5;

  public:
    struct data {
        std::vector<T> r2, c;
        bool empty() const { return r2.empty() && c.empty(); }
        size_t numKnots() const { return r2.size(); }
    };
    void setTolerance(T tol, T = -1, T = -1, T = -1) { utol = tol; }
    data generate(std::function<T(T)> f, double rmin, double rmax) {
        if (rmin != 0.0 || rmax != 1.0) throw std::runtime_error("Andrea: only the unit interval used by Splined is supported");
        detail::SplineTable t;
        detail::require_ok(detail::andrea_generate(f, utol, t), "Andrea spline: try to increase utol/ftol");
        data d;
        d.r2 = t.r2;
        d.c = t.c;
        return d;
    }
    T eval(const data &d, T x) const {
        const size_t pos = std::lower_bound(d.r2.begin(), d.r2.end(), x) - d.r2.begin() - 1;
        const T dz = x - d.r2[pos];
        const T *c = &d.c[6 * pos];
        return c[0] + dz * (c[1] + dz * (c[2] + dz * (c[3] + dz * (c[4] + dz * (c[5])))));
    }
};
} // namespace Tabulate

class Splined : public EnergyImplementation<Splined> {
    std::shared_ptr<SchemeBase> pot;
    double tolerance = 1e-3; // splined.hpp:340
    detail::SplineTable tables[4];
    std::vector<double> packed;
    detail::SplineSrf srf_;

    void install_tables() {
        cg_scheme_desc d = pot->descriptor(); // cutoff, kappa, prefactors, scheme id of the wrapped scheme (splined.hpp:329)
        cg_scheme_desc out;
        detail::require_ok(detail::make_splined_from_tables(d, tables, out), "Splined: table too large");
        adopt(out);
        name = pot->name;
        doi = pot->doi;
        packed.assign((size_t)detail::SplineSrf::packed_size(out), 0.0);
        srf_.pack_from(out, packed.data());
    }
    template <int K> double s(double q) const {
        double v[4];
        srf_.eval<K>(q, v, nullptr);
        return v[K];
    }

  public:
    inline Splined() { adopt(detail::desc_base(CG_SPLINE, infinity, infinity)); }
    inline void setTolerance(double tol) { tolerance = tol; }
    inline std::vector<size_t> numKnots() const {
        std::vector<size_t> n;
        for (const auto &t : tables) n.push_back(t.r2.size());
        return n;
    }
    /** pot.spline<Wolf>(cutoff, alpha): tabulate S and its three derivatives of T(args...) with the reference's Andrea algorithm */
    template <class T, class... Args> void spline(Args &&...args) {
        pot = std::make_shared<T>(args...);
        auto p = pot;
        const std::function<double(double)> f[4] = {[p](double q) { return p->short_range_function(q); },
                                                    [p](double q) { return p->short_range_function_derivative(q); },
                                                    [p](double q) { return p->short_range_function_second_derivative(q); },
                                                    [p](double q) { return p->short_range_function_third_derivative(q); }};
        for (int k = 0; k < 4; k++) detail::require_ok(detail::andrea_generate(f[k], tolerance, tables[k]), "Andrea spline: try to increase utol/ftol");
        install_tables();
    }
    /** same, with tables generated elsewhere (knots ascending, 6 coefficients per interval) */
    template <class T, class... Args> void spline_from_tables(const detail::SplineTable (&t)[4], Args &&...args) {
        pot = std::make_shared<T>(args...);
        for (int k = 0; k < 4; k++) tables[k] = t[k];
        install_tables();
    }
    double short_range_function(double q) const override { return s<0>(q); }
    double short_range_function_derivative(double q) const override { return s<1>(q); }
    double short_range_function_second_derivative(double q) const override { return s<2>(q); }
    double short_range_function_third_derivative(double q) const override { return s<3>(q); }
};

} // namespace CoulombGalore
