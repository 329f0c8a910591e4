// scheme_api.cpp -- the host-only part of the C ABI: descriptor constructors (cg_scheme_*) and cg_srf_host.
// Thin wrappers over include/coulombgalore/detail/descriptor.hpp, the same code the header classes use.
#include <cstring>
#include <string>
#include <vector>

#include "cg_b200.h"
#include "coulombgalore/detail/descriptor.hpp"

namespace cgd = CoulombGalore::detail;

extern "C" {

int cg_scheme_plain(double debye_length, cg_scheme_desc *out) { return out ? cgd::make_plain(debye_length, *out) : CG_ERR_INVALID; }
int cg_scheme_reactionfield(double cutoff, double epsRF, double epsr, int shifted, cg_scheme_desc *out) {
    return out ? cgd::make_reactionfield(cutoff, epsRF, epsr, shifted != 0, *out) : CG_ERR_INVALID;
}
int cg_scheme_poisson(double cutoff, int C, int D, double debye_length, cg_scheme_desc *out) {
    return out ? cgd::make_poisson(cutoff, C, D, debye_length, *out) : CG_ERR_INVALID;
}
int cg_scheme_qpotential(double cutoff, int order, cg_scheme_desc *out) {
    return out ? cgd::make_qpotential(cutoff, order, false, *out) : CG_ERR_INVALID;
}
int cg_scheme_qpotential5(double cutoff, cg_scheme_desc *out) { return out ? cgd::make_qpotential(cutoff, 5, true, *out) : CG_ERR_INVALID; }
int cg_scheme_fanourgakis(double cutoff, cg_scheme_desc *out) { return out ? cgd::make_fanourgakis(cutoff, *out) : CG_ERR_INVALID; }
int cg_scheme_fennell(double cutoff, double alpha, cg_scheme_desc *out) { return out ? cgd::make_fennell(cutoff, alpha, *out) : CG_ERR_INVALID; }
int cg_scheme_zerodipole(double cutoff, double alpha, cg_scheme_desc *out) {
    return out ? cgd::make_zerodipole(cutoff, alpha, *out) : CG_ERR_INVALID;
}
int cg_scheme_zahn(double cutoff, double alpha, cg_scheme_desc *out) { return out ? cgd::make_zahn(cutoff, alpha, *out) : CG_ERR_INVALID; }
int cg_scheme_wolf(double cutoff, double alpha, cg_scheme_desc *out) { return out ? cgd::make_wolf(cutoff, alpha, *out) : CG_ERR_INVALID; }
int cg_scheme_ewald(double cutoff, double alpha, double eps_sur, double debye_length, cg_scheme_desc *out) {
    return out ? cgd::make_ewald(cutoff, alpha, eps_sur, debye_length, *out) : CG_ERR_INVALID;
}
int cg_scheme_ewaldt(double cutoff, double alpha, double eps_sur, double debye_length, cg_scheme_desc *out) {
    return out ? cgd::make_ewaldt(cutoff, alpha, eps_sur, debye_length, *out) : CG_ERR_INVALID;
}

// The reference's run-time factory (createScheme, all.hpp:42-104) without the JSON dependency: `type` and the
// (key, value) pairs are what the JSON object holds; keys and defaults are those of the reference's JSON constructors.
int cg_scheme_create(const char *type, int nkeys, const char *const *keys, const double *values, cg_scheme_desc *out) {
    if (!type || !out || nkeys < 0 || (nkeys > 0 && (!keys || !values))) return CG_ERR_INVALID;
    bool missing = false;
    auto find = [&](const char *k, double *v) {
        for (int i = 0; i < nkeys; i++)
            if (keys[i] && std::strcmp(keys[i], k) == 0) {
                *v = values[i];
                return true;
            }
        return false;
    };
    auto at = [&](const char *k) { // j.at(k): mandatory
        double v = 0.0;
        if (!find(k, &v)) missing = true;
        return v;
    };
    auto value = [&](const char *k, double dflt) { // j.value(k, default)
        double v = dflt;
        find(k, &v);
        return v;
    };
    const double inf = cgd::kInfinity;
    const std::string t(type);
    int rc = CG_ERR_INVALID;
    if (t == "plain") rc = cgd::make_plain(value("debyelength", inf), *out);
    else if (t == "wolf") { const double c = at("cutoff"), a = at("alpha"); if (!missing) rc = cgd::make_wolf(c, a, *out); }
    else if (t == "zahn") { const double c = at("cutoff"), a = at("alpha"); if (!missing) rc = cgd::make_zahn(c, a, *out); }
    else if (t == "fennell") { const double c = at("cutoff"), a = at("alpha"); if (!missing) rc = cgd::make_fennell(c, a, *out); }
    else if (t == "zerodipole") { const double c = at("cutoff"), a = at("alpha"); if (!missing) rc = cgd::make_zerodipole(c, a, *out); }
    else if (t == "fanourgakis") { const double c = at("cutoff"); if (!missing) rc = cgd::make_fanourgakis(c, *out); }
    else if (t == "qpotential") { const double c = at("cutoff"), o = at("order"); if (!missing) rc = cgd::make_qpotential(c, (int)o, false, *out); }
    else if (t == "ewald") {
        const double c = at("cutoff"), a = at("alpha");
        if (!missing) rc = cgd::make_ewald(c, a, value("epss", inf), value("debyelength", inf), *out);
    } else if (t == "ewaldt") {
        const double c = at("cutoff"), a = at("alpha");
        if (!missing) rc = cgd::make_ewaldt(c, a, value("epss", inf), value("debyelength", inf), *out);
    } else if (t == "poisson") {
        const double c = at("cutoff"), C = at("C"), D = at("D");
        if (!missing) rc = cgd::make_poisson(c, (int)C, (int)D, value("debyelength", inf), *out);
    } else if (t == "reactionfield") {
        const double c = at("cutoff"), erf = at("epsRF"), er = at("epsr"), sh = at("shifted");
        if (!missing) rc = cgd::make_reactionfield(c, erf, er, sh != 0.0, *out);
    } else if (t == "spline") {
        rc = CG_ERR_UNSUPPORTED; // known keyword, but the reference's factory builds nothing for it either (all.hpp:98-100)
    }
    return rc;
}

int64_t cg_spline_storage_doubles(void) { return 4 * (CG_SPLINE_MAX_KNOTS + 6 * (CG_SPLINE_MAX_KNOTS - 1)); }

static int splined_into(const cg_scheme_desc *inner, const cgd::SplineTable t[4], double *storage, cg_scheme_desc *out) {
    *out = *inner;
    out->functor = CG_SPLINE;
    double *w = storage;
    for (int k = 0; k < 4; k++) {
        const int n = (int)t[k].r2.size();
        if (n < 2 || n > CG_SPLINE_MAX_KNOTS || (int)t[k].c.size() != 6 * (n - 1)) return CG_ERR_INVALID;
        out->spline_knots[k] = n;
        std::memcpy(w, t[k].r2.data(), sizeof(double) * (size_t)n);
        out->spline_r2[k] = w;
        w += n;
        std::memcpy(w, t[k].c.data(), sizeof(double) * (size_t)(6 * (n - 1)));
        out->spline_c[k] = w;
        w += 6 * (n - 1);
    }
    return CG_OK;
}

int cg_scheme_splined(const cg_scheme_desc *inner, double utol, double *storage, cg_scheme_desc *out) {
    if (!inner || !storage || !out) return CG_ERR_INVALID;
    cgd::SplineTable t[4];
    const int rc = cgd::generate_spline_tables(*inner, utol, t);
    if (rc != CG_OK) return rc;
    return splined_into(inner, t, storage, out);
}

int cg_scheme_splined_tables(const cg_scheme_desc *inner, const int32_t n_knots[4], const double *const r2[4], const double *const c[4],
                             cg_scheme_desc *out) {
    if (!inner || !n_knots || !r2 || !c || !out || inner->functor == CG_SPLINE) return CG_ERR_INVALID;
    *out = *inner;
    out->functor = CG_SPLINE;
    for (int k = 0; k < 4; k++) {
        if (n_knots[k] < 2 || n_knots[k] > CG_SPLINE_MAX_KNOTS || !r2[k] || !c[k]) return CG_ERR_INVALID;
        for (int i = 1; i < n_knots[k]; i++)
            if (!(r2[k][i] > r2[k][i - 1])) return CG_ERR_INVALID; // knots must ascend (lower_bound, splined.hpp:169)
        out->spline_knots[k] = n_knots[k];
        out->spline_r2[k] = r2[k];
        out->spline_c[k] = c[k];
    }
    return CG_OK;
}

int cg_srf_host(const cg_scheme_desc *d, int64_t n, const double *q, double *out) {
    if (!d || n < 0 || (n > 0 && (!q || !out))) return CG_ERR_INVALID;
    for (int64_t i = 0; i < n; i++) {
        double s[4];
        cgd::srf_host(*d, q[i], s);
        for (int k = 0; k < 4; k++) out[k * n + i] = s[k];
    }
    return CG_OK;
}

} // extern "C"
