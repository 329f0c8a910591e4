// coulombgalore/reciprocal.hpp -- ReciprocalEwaldState / ReciprocalEwaldGaussian with the reference's names and methods
// (reference ewald.hpp:44-94, 214-385), evaluated on the GPU through the C ABI (cg_recip_*, csrc/recip.cu).
// The state class holds the same public fields; the Gaussian class owns a device handle with the k-vectors, A_k and Q_mp.
// Per-particle methods of the reference (one position per call) exist as such and, for throughput, as batched overloads.
// Link against libcg_b200.so; there is no CPU fallback.
#pragma once
#include <functional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "cg_b200.h"
#include "ewald.hpp"

namespace CoulombGalore {

class ReciprocalEwaldState {
  public:
    int reciprocal_cutoff = 0;                //!< Inverse space cutoff
    double cutoff = 0.0;                      //!< Real-space cutoff
    double surface_dielectric_constant = 0.0; //!< Surface dielectric constant
    double kappa = 0.0;                       //!< Inverse Debye screening length
    double alpha = 0.0;
    vec3 box_length = {0.0, 0.0, 0.0};        //!< Box dimensions
    enum Policies { PBC, PBCEigen, IPBC, IPBCEigen, INVALID };
    Policies policy = PBC;
    inline double getVolume() const { return box_length[0] * box_length[1] * box_length[2]; }
};

class ReciprocalEwaldGaussian : public ReciprocalEwaldState {
    cg_recip *h_ = nullptr;
    void check(int rc, const char *where) const {
        if (rc != CG_OK) throw std::runtime_error(std::string(where) + ": " + cg_recip_last_error(h_));
    }
    struct Soa {
        std::vector<double> x, y, z, q, mx, my, mz;
    };
    template <class Positions, class Charges, class Dipoles> static Soa soa(const Positions &positions, const Charges &charges, const Dipoles &dipoles) {
        Soa s;
        for (const auto &p : positions) {
            s.x.push_back(p[0]);
            s.y.push_back(p[1]);
            s.z.push_back(p[2]);
        }
        for (const auto &c : charges) s.q.push_back(c);
        for (const auto &d : dipoles) {
            s.mx.push_back(d[0]);
            s.my.push_back(d[1]);
            s.mz.push_back(d[2]);
        }
        return s;
    }
    ReciprocalEwaldGaussian(const ReciprocalEwaldGaussian &) = delete;
    ReciprocalEwaldGaussian &operator=(const ReciprocalEwaldGaussian &) = delete;

  public:
    const Ewald real_space; //!< Real-space energy functions

    explicit ReciprocalEwaldGaussian(const ReciprocalEwaldState &state, int device = -1)
        : ReciprocalEwaldState(state),
          real_space(state.cutoff, state.alpha, state.surface_dielectric_constant, 1.0 / state.kappa) { // ewald.hpp:122-123
        generateKVectors(state.box_length, device);
    }
    ~ReciprocalEwaldGaussian() { cg_recip_destroy(h_); }

    inline void generateKVectors(const vec3 &box, int device = -1) { // :257-286; Q_mp is zeroed
        cg_recip_destroy(h_);
        h_ = nullptr;
        box_length = box;
        const double b[3] = {box[0], box[1], box[2]};
        if (cg_recip_create(cutoff, alpha, surface_dielectric_constant, kappa, reciprocal_cutoff, b, device, &h_) != CG_OK)
            throw std::runtime_error(std::string("cg_recip_create: ") + cg_recip_last_error(nullptr));
    }
    inline int numKVectors() const { return cg_recip_num_kvectors(h_); }
    inline void setZeroComplex() {
        const std::vector<double> zero(2 * (size_t)numKVectors(), 0.0);
        check(cg_recip_set_q(h_, zero.data()), "cg_recip_set_q");
    }
    cg_recip *handle() { return h_; }

    // updateComplex(positions, charges, dipoles[, std::minus<>()])  (:295-300)
    template <class Positions, class Charges, class Dipoles, class BinaryOp = std::plus<>>
    void updateComplex(Positions &positions, Charges &charges, Dipoles &dipoles, BinaryOp = BinaryOp()) {
        const Soa s = soa(positions, charges, dipoles);
        const int sign = std::is_same<BinaryOp, std::minus<>>::value ? -1 : 1;
        check(cg_recip_update(h_, (int64_t)s.x.size(), s.x.data(), s.y.data(), s.z.data(), s.q.data(), s.mx.data(), s.my.data(), s.mz.data(), sign),
              "cg_recip_update");
    }
    inline double reciprocal_energy() { // :305-317
        double e = 0.0;
        check(cg_recip_energy(h_, &e), "cg_recip_energy");
        return e;
    }
    inline vec3 reciprocal_force(const vec3 &position, const double charge, const vec3 &dipole_moment) { // :325-336
        const double r[3] = {position[0], position[1], position[2]}, m[3] = {dipole_moment[0], dipole_moment[1], dipole_moment[2]};
        double f[3];
        check(cg_recip_force_field(h_, 1, &r[0], &r[1], &r[2], &charge, &m[0], &m[1], &m[2], f, nullptr), "cg_recip_force_field");
        return vec3(f[0], f[1], f[2]);
    }
    inline vec3 reciprocal_field(const vec3 &position) { // :342-351
        const double r[3] = {position[0], position[1], position[2]};
        double f[3];
        check(cg_recip_force_field(h_, 1, &r[0], &r[1], &r[2], nullptr, nullptr, nullptr, nullptr, nullptr, f), "cg_recip_force_field");
        return vec3(f[0], f[1], f[2]);
    }
    // batched: force[3n] and / or field[3n] (AoS) for all particles in one launch
    template <class Positions, class Charges, class Dipoles>
    void reciprocal_force_field(const Positions &positions, const Charges &charges, const Dipoles &dipoles, double *force, double *field) {
        const Soa s = soa(positions, charges, dipoles);
        check(cg_recip_force_field(h_, (int64_t)s.x.size(), s.x.data(), s.y.data(), s.z.data(), s.q.data(), s.mx.data(), s.my.data(), s.mz.data(), force,
                                   field),
              "cg_recip_force_field");
    }
    template <class Positions, class Charges, class Dipoles>
    vec3 surface_field(const Positions &positions, const Charges &charges, const Dipoles &dipoles) const { // :355-362
        const Soa s = soa(positions, charges, dipoles);
        double f[3];
        check(cg_recip_surface(h_, (int64_t)s.x.size(), s.x.data(), s.y.data(), s.z.data(), s.q.data(), s.mx.data(), s.my.data(), s.mz.data(), nullptr, f),
              "cg_recip_surface");
        return vec3(f[0], f[1], f[2]);
    }
    template <class Positions, class Charges, class Dipoles>
    double surface_energy(const Positions &positions, const Charges &charges, const Dipoles &dipoles) const { // :367-374
        const Soa s = soa(positions, charges, dipoles);
        double e = 0.0;
        check(cg_recip_surface(h_, (int64_t)s.x.size(), s.x.data(), s.y.data(), s.z.data(), s.q.data(), s.mx.data(), s.my.data(), s.mz.data(), &e, nullptr),
              "cg_recip_surface");
        return e;
    }
    template <class Positions, class Charges, class Dipoles>
    vec3 surface_force(Positions &positions, Charges &charges, Dipoles &dipoles, const double charge) const { // :383-386
        const vec3 f = surface_field(positions, charges, dipoles);
        return vec3(charge * f[0], charge * f[1], charge * f[2]);
    }
};

} // namespace CoulombGalore
