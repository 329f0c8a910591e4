// tests/cpp/surface_probe.cpp -- user-style call of EwaldTruncated::surface_energy (reference ewald_truncated.hpp:91-104)
// through the header-only class; linked against libcg_b200.so by tests/test_gpu_system.py (the member reduces on the GPU).
#include <coulombgalore/ewald_truncated.hpp>

#include <cstdint>
#include <vector>

using namespace CoulombGalore;

extern "C" int probe_ewaldt_surface(double cutoff, double alpha, double eps_sur, int64_t n, const double *x, const double *y, const double *z,
                                    const double *q, const double *mx, const double *my, const double *mz, double volume, double *energy) {
    try {
        const EwaldTruncated pot(cutoff, alpha, eps_sur);
        std::vector<vec3> pos, dip;
        std::vector<double> ch;
        for (int64_t i = 0; i < n; i++) {
            pos.push_back(vec3(x[i], y[i], z[i]));
            dip.push_back(vec3(mx[i], my[i], mz[i]));
            ch.push_back(q[i]);
        }
        *energy = pot.surface_energy(pos, ch, dip, volume);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
