// oracle/port_main.cpp -- TEST INFRASTRUCTURE ONLY.
// Builds oracle/libcg_port.so: the independent restatement (port_schemes.hpp) behind the cgo_api.h C ABI.
#include <complex>

#include "port_schemes.hpp"

#include "harness.hpp"

namespace {
struct PortImpl : cgo::Impl<cgport::Scheme, cgport::V3, cgport::M33> {
    explicit PortImpl(const cgo_params &p) : cgo::Impl<cgport::Scheme, cgport::V3, cgport::M33>(p) {}
    int spline_table(int k, int32_t *n, double *r2, double *c) override {
        if (!pot.is_splined() || k < 0 || k > 3) return -1;
        const cgport::Table &t = pot.table(k);
        if (n) *n = (int32_t)t.r2.size();
        if (r2) std::copy(t.r2.begin(), t.r2.end(), r2);
        if (c) std::copy(t.c.begin(), t.c.end(), c);
        return 0;
    }
};
} // namespace

cgo::Iface *cgo::make(const cgo_params &p) { return new PortImpl(p); }

// ---- reciprocal-space Ewald, restated from the reference (ewald.hpp:44-94, 214-385) with plain arrays ----
namespace {
struct PortRecip : cgo::RecipIface {
    double cutoff, alpha, eps_sur, kappa, box[3];
    int kc;
    std::vector<double> k;             // 3K
    std::vector<double> aks;           // K
    std::vector<std::complex<double>> Q; // K
    const double pi = 4.0 * std::atan(1.0);
    PortRecip(double cutoff_, double alpha_, double eps_, double kappa_, int kc_, const double *b)
        : cutoff(cutoff_), alpha(alpha_), eps_sur(eps_), kappa(kappa_), kc(kc_) {
        for (int d = 0; d < 3; d++) box[d] = b[d];
        // generateKVectors :257-286: all integer triples with 0 < n^2 <= kc^2, nx outermost, nz innermost
        for (int nx = -kc; nx <= kc; nx++)
            for (int ny = -kc; ny <= kc; ny++)
                for (int nz = -kc; nz <= kc; nz++) {
                    const int r = nx * nx + ny * ny + nz * nz;
                    if (r > 0 && r <= kc * kc) {
                        k.push_back(2.0 * pi * (1.0 / box[0]) * nx);
                        k.push_back(2.0 * pi * (1.0 / box[1]) * ny);
                        k.push_back(2.0 * pi * (1.0 / box[2]) * nz);
                    }
                }
        if (k.empty()) { // resize(0) special case :68-72: one dummy vector with zero weight
            k = {1.0, 0.0, 0.0};
            aks = {0.0};
        } else {
            // setAks :216-225 (kappa^2 enters twice, as in the reference)
            const double c2 = cutoff * cutoff, rd2 = alpha * alpha * c2, rk2 = kappa * kappa * c2;
            aks.resize(k.size() / 3);
            for (size_t i = 0; i < aks.size(); i++) {
                const double k2 = k[3 * i] * k[3 * i] + k[3 * i + 1] * k[3 * i + 1] + k[3 * i + 2] * k[3 * i + 2] + kappa * kappa;
                aks[i] = std::exp(-(k2 * c2 + rk2) / (4.0 * rd2)) / k2;
            }
        }
        Q.assign(aks.size(), std::complex<double>(0.0, 0.0));
    }
    double volume() const { return box[0] * box[1] * box[2]; }
    int numk() override { return (int)aks.size(); }
    int kvectors(double *k3, double *a) override {
        if (k3) std::copy(k.begin(), k.end(), k3);
        if (a) std::copy(aks.begin(), aks.end(), a);
        return 0;
    }
    int update(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx, const double *my,
               const double *mz, int sign) override { // calcQ :236-250, updateComplex :295-300
        if (sign == 0) std::fill(Q.begin(), Q.end(), std::complex<double>(0.0, 0.0));
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)aks.size(); i++) {
            std::complex<double> s(0.0, 0.0);
            const double kx = k[3 * i], ky = k[3 * i + 1], kz = k[3 * i + 2];
            for (int64_t p = 0; p < N; p++) {
                const double kr = kx * x[p] + ky * y[p] + kz * z[p];
                const double c = std::cos(kr), sn = std::sin(kr);
                const double ch = q ? q[p] : 0.0;
                const double mk = mx ? mx[p] * kx + my[p] * ky + mz[p] * kz : 0.0;
                s += ch * std::complex<double>(c, sn) + mk * std::complex<double>(-sn, c);
            }
            Q[(size_t)i] = sign < 0 ? Q[(size_t)i] - s : Q[(size_t)i] + s;
        }
        return 0;
    }
    int getq(double *q2k) override {
        for (size_t i = 0; i < Q.size(); i++) {
            q2k[2 * i] = Q[i].real();
            q2k[2 * i + 1] = Q[i].imag();
        }
        return 0;
    }
    int energy(double *e) override { // reciprocal_energy :305-317
        double s = 0.0;
        for (size_t i = 0; i < Q.size(); i++) s += std::norm(Q[i]) * aks[i];
        *e = 2.0 * pi / volume() * s;
        return 0;
    }
    int force_field(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx, const double *my,
                    const double *mz, double *force, double *field, double *fs) override { // :325-351
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < N; p++) {
            double F[3] = {0, 0, 0}, E[3] = {0, 0, 0};
            long double sc = 0;
            const double ch = q ? q[p] : 0.0;
            for (size_t i = 0; i < aks.size(); i++) {
                const double kx = k[3 * i], ky = k[3 * i + 1], kz = k[3 * i + 2];
                const double kr = kx * x[p] + ky * y[p] + kz * z[p];
                const double mk = mx ? mx[p] * kx + my[p] * ky + mz[p] * kz : 0.0;
                const std::complex<double> e(std::cos(kr), std::sin(kr));
                const double rf = std::real(e * std::complex<double>(-mk, ch) * std::conj(Q[i]));
                const double re = std::real(std::complex<double>(-e.imag(), e.real()) * std::conj(Q[i]));
                const double kk[3] = {kx, ky, kz};
                for (int d = 0; d < 3; d++) {
                    F[d] += rf * kk[d] * aks[i];
                    E[d] += re * kk[d] * aks[i];
                }
                sc += (std::fabs(mk) + std::fabs(ch)) * std::abs(Q[i]) * std::sqrt(kx * kx + ky * ky + kz * kz) * aks[i];
            }
            for (int d = 0; d < 3; d++) {
                if (force) force[3 * p + d] = -4.0 * pi / volume() * F[d];
                if (field) field[3 * p + d] = -4.0 * pi / volume() * E[d];
            }
            if (fs) fs[p] = (double)(4.0 * pi / volume() * sc);
        }
        return 0;
    }
    int surface(int64_t N, const double *x, const double *y, const double *z, const double *q, const double *mx, const double *my,
                const double *mz, double *energy, double *field3) override { // dipoleMoment :77-88, :360-374
        double M[3] = {0, 0, 0};
        for (int64_t p = 0; p < N; p++) {
            const double ch = q ? q[p] : 0.0;
            M[0] += x[p] * ch;
            M[1] += y[p] * ch;
            M[2] += z[p] * ch;
        }
        if (mx)
            for (int64_t p = 0; p < N; p++) {
                M[0] += mx[p];
                M[1] += my[p];
                M[2] += mz[p];
            }
        if (energy) *energy = 2.0 * pi / (2.0 * eps_sur + 1.0) / volume() * (M[0] * M[0] + M[1] * M[1] + M[2] * M[2]);
        if (field3)
            for (int d = 0; d < 3; d++) field3[d] = -4.0 * pi / (2.0 * eps_sur + 1.0) / volume() * M[d];
        return 0;
    }
};
} // namespace
cgo::RecipIface *cgo::make_recip(double cutoff, double alpha, double eps_sur, double kappa, int kc, const double *box) {
    return new PortRecip(cutoff, alpha, eps_sur, kappa, kc, box);
}

CGO_DEFINE_C_ABI("port")

extern "C" {
// ewald_truncated.hpp:91-104: the charge-weighted position sum and the dipole sum are accumulated separately (particle
// order), squared as P.P + 2 P.D + D.D; the prefactor uses the constructor argument as given (the constructor's
// `surface_dielectric_constant < 1 -> infinity` repair at :56-58 assigns the parameter, not the member)
int cgo_ewaldt_surface(double cutoff, double alpha, double eps_sur, int64_t N, const double *x, const double *y, const double *z,
                       const double *q, const double *mx, const double *my, const double *mz, double volume, double *energy) {
    (void)cutoff;
    (void)alpha;
    const double pi = 4.0 * std::atan(1.0);
    double P[3] = {0, 0, 0}, D[3] = {0, 0, 0};
    for (int64_t p = 0; p < N; p++) {
        const double ch = q ? q[p] : 0.0;
        P[0] += x[p] * ch;
        P[1] += y[p] * ch;
        P[2] += z[p] * ch;
        if (mx) {
            D[0] += mx[p];
            D[1] += my[p];
            D[2] += mz[p];
        }
    }
    const double pp = P[0] * P[0] + P[1] * P[1] + P[2] * P[2], pd = P[0] * D[0] + P[1] * D[1] + P[2] * D[2],
                 dd = D[0] * D[0] + D[1] * D[1] + D[2] * D[2];
    *energy = 2.0 * pi / (2.0 * eps_sur + 1.0) / volume * (pp + 2.0 * pd + dd);
    return 0;
}
int cgo_qpochhammer(double q, int32_t l, int32_t P, double *out4) {
    cgport::qpochhammer(q, P, out4, l);
    return 0;
}
}
