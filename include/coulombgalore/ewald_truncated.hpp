// coulombgalore/ewald_truncated.hpp -- EwaldTruncated: same class name, constructor and members as the reference
// (ewald_truncated.hpp:37-120); constants and S(q) come from detail/descriptor.hpp and detail/srf.hpp, the code the sm_100a
// kernels are instantiated from.  surface_energy (reference :91-104) is an O(N) reduction and runs on the device: it needs
// libcg_b200 at link time, the per-pair members do not.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../cg_b200.h"
#include "core.hpp"

namespace CoulombGalore {

class EwaldTruncated : public detail::FunctorScheme<EwaldTruncated, detail::ErfcFamilySrf<detail::kEwaldT>> {
    // The constructor argument as given: the reference's repair `if (surface_dielectric_constant < 1) ... = infinity`
    // (ewald_truncated.hpp:56-58) assigns the PARAMETER, so its member -- which surface_energy reads -- keeps the argument.
    double surface_dielectric_constant_;

  public:
    inline EwaldTruncated(double cutoff, double alpha, double surface_dielectric_constant = infinity, double debye_length = infinity)
        : surface_dielectric_constant_(surface_dielectric_constant) {
        cg_scheme_desc d;
        detail::require_ok(detail::make_ewaldt(cutoff, alpha, surface_dielectric_constant, debye_length, d), "EwaldTruncated: invalid parameters");
        install(d, "EwaldT real-space", "XYZ");
    }

    /**
     * Surface term (reference ewald_truncated.hpp:91-104): 2 pi / (2 eps_sur + 1) / V * |sum z r + sum mu|^2, with the two sums
     * accumulated separately and squared as P.P + 2 P.D + D.D like the reference does.  The sums are reduced on the GPU
     * (cg_dipole_moments); `device` < 0 = the current device.
     */
    inline double surface_energy(const std::vector<vec3> &positions, const std::vector<double> &charges, const std::vector<vec3> &dipoles,
                                 double volume, int device = -1) const {
        if (positions.size() != charges.size() || positions.size() != dipoles.size())
            throw std::invalid_argument("EwaldTruncated::surface_energy: positions, charges and dipoles differ in length");
        const size_t n = positions.size();
        std::vector<double> soa(7 * n);
        for (size_t i = 0; i < n; i++) {
            for (int d = 0; d < 3; d++) {
                soa[d * n + i] = positions[i][d];
                soa[(4 + d) * n + i] = dipoles[i][d];
            }
            soa[3 * n + i] = charges[i];
        }
        cg_handle *h = nullptr;
        if (cg_create(&descriptor(), device, &h) != CG_OK) throw std::runtime_error(std::string("cg_create: ") + cg_last_error(nullptr));
        double m[6] = {0, 0, 0, 0, 0, 0};
        const double *s = soa.data();
        int rc = cg_set_system(h, (int64_t)n, s, s + n, s + 2 * n, s + 3 * n, s + 4 * n, s + 5 * n, s + 6 * n, nullptr, 0);
        if (rc == CG_OK) rc = cg_dipole_moments(h, m);
        const std::string err = rc == CG_OK ? std::string() : std::string(cg_last_error(h));
        cg_destroy(h);
        if (rc != CG_OK) throw std::runtime_error("EwaldTruncated::surface_energy: " + err);
        const double pi = 4.0 * std::atan(1.0);
        const double sq = (m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) + 2.0 * (m[0] * m[3] + m[1] * m[4] + m[2] * m[5]) +
                          (m[3] * m[3] + m[4] * m[4] + m[5] * m[5]);
        return 2.0 * pi / (2.0 * surface_dielectric_constant_ + 1.0) / volume * sq;
    }
};

} // namespace CoulombGalore
