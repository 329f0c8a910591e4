// coulombgalore/batched.hpp -- the batched host entry point the reference lacks: positions, charges and dipoles in
// SoA buffers, evaluated for ALL pairs within the cut-off by the sm_100a kernels behind the C ABI (include/cg_b200.h).
//
//     CoulombGalore::Wolf pot(12.0, 0.25);                          // exactly as with the reference
//     CoulombGalore::BatchedEvaluator gpu(pot);                     // cg_create(pot.descriptor())
//     gpu.set_system(n, x, y, z, q, nullptr, nullptr, nullptr, box, true);
//     gpu.eval(CG_MODE_ION, CG_FORCE, force, nullptr, nullptr, nullptr, &pairs);   // sum_j ion_ion_force(z_j, z_i, r_i - r_j)
//
// Link with -lcg_b200 (coulombgalore_b200/libcg_b200.so).  Errors throw std::runtime_error carrying cg_last_error().
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

#include "../cg_b200.h"
#include "core.hpp"

namespace CoulombGalore {

class BatchedEvaluator {
    cg_handle *h_ = nullptr;
    void check(int rc, const char *where) const {
        if (rc != CG_OK) throw std::runtime_error(std::string(where) + ": " + cg_last_error(h_));
    }

  public:
    explicit BatchedEvaluator(const SchemeBase &scheme, int device = -1) {
        const int rc = cg_create(&scheme.descriptor(), device, &h_);
        if (rc != CG_OK) throw std::runtime_error(std::string("cg_create: ") + cg_last_error(nullptr));
    }
    ~BatchedEvaluator() { cg_destroy(h_); }
    BatchedEvaluator(const BatchedEvaluator &) = delete;
    BatchedEvaluator &operator=(const BatchedEvaluator &) = delete;

    cg_handle *handle() const { return h_; }
    void set_stream(void *cuda_stream) { check(cg_set_stream(h_, cuda_stream), "cg_set_stream"); }
    // quad[n][9]: row-major 3x3 tensors for CG_MODE_MULTIPOLE_QUAD; call after set_system
    void set_quadrupoles(const double *quad) { check(cg_set_quadrupoles(h_, quad), "cg_set_quadrupoles"); }
    // evaluate on `source`'s system (same cut-off): one upload + sort per configuration for several schemes
    void share_system(BatchedEvaluator &source) { check(cg_share_system(h_, source.h_), "cg_share_system"); }
    void set_shard(int rank, int nranks) { check(cg_set_shard(h_, rank, nranks), "cg_set_shard"); }
    /** host SoA buffers; q / mu* may be null (all zero); box is read when pbc */
    void set_system(int64_t n, const double *x, const double *y, const double *z, const double *q, const double *mux, const double *muy,
                    const double *muz, const double box[3], bool pbc) {
        check(cg_set_system(h_, n, x, y, z, q, mux, muy, muz, box, pbc ? 1 : 0), "cg_set_system");
    }
    /** device SoA buffers (zero copy) */
    void set_system_device(int64_t n, const double *x, const double *y, const double *z, const double *q, const double *mux,
                           const double *muy, const double *muz, const double box[3], bool pbc) {
        check(cg_set_system_device(h_, n, x, y, z, q, mux, muy, muz, box, pbc ? 1 : 0), "cg_set_system_device");
    }
    /** host outputs: force[3n], field[3n], potential[n] (any may be null), total energy, ordered in-cutoff pair count */
    void eval(int mode, int what, double *force, double *field, double *potential, double *energy, int64_t *pairs) {
        check(cg_eval(h_, mode, what, force, field, potential, energy, pairs), "cg_eval");
    }
    // the same in two halves (several evaluators overlap their copies and kernels; buffers stay valid until eval_end)
    void eval_begin(int mode, int what, double *force, double *field, double *potential) {
        check(cg_eval_begin(h_, mode, what, force, field, potential), "cg_eval_begin");
    }
    void eval_end(double *energy, int64_t *pairs) { check(cg_eval_end(h_, energy, pairs), "cg_eval_end"); }
    // O(N) terms: out[0] = sum self_energy, out[1] = neutralization_energy, out[2] = sum z; self_field[3n] may be null
    void self_terms(double out[3], double *self_field) { check(cg_self_terms(h_, out, self_field), "cg_self_terms"); }
    void dipole_torque(const double *field, double *torque) { check(cg_dipole_torque(h_, field, torque), "cg_dipole_torque"); }
    /** device outputs; scalars[0] = U, scalars[1] = pair count; asynchronous on the handle's stream.  From the second
     *  evaluation of a system shape on, sizes come from the previous evaluation (no host round trip); a configuration that
     *  does not fit them yields scalars = {NaN, -1} and partial outputs -- check scalars[1] < 0 and call again, or use
     *  eval_device_checked */
    void eval_device(int mode, int what, double *force, double *field, double *potential, double *scalars) {
        check(cg_eval_device(h_, mode, what, force, field, potential, scalars), "cg_eval_device");
    }
    /** the per-particle results this evaluator's shard computed in its last (cell-list) evaluation, compacted for the way back to
     *  the host: slot g * 32 + l <-> particle l of i-group g, ids_out[slot] = ids[row] (ids == nullptr: the row itself), -1 for
     *  unused slots; returns the rows dst / ids_out must have (cap_rows = 0 only asks).  Device pointers, one launch. */
    int64_t collect_evaluated(const double *ids, int narrays, int width, const double *const *src, double *const *dst, double *ids_out,
                              int64_t cap_rows) {
        int64_t rows = 0;
        check(cg_collect_evaluated(h_, ids, narrays, width, src, dst, ids_out, cap_rows, &rows), "cg_collect_evaluated");
        return rows;
    }
    /** eval_device + a read of the two scalars (waits for the stream) + one exactly-sized re-evaluation if the first was invalid */
    void eval_device_checked(int mode, int what, double *force, double *field, double *potential, double *scalars, double host_scalars[2]) {
        for (int attempt = 0; attempt < 2; attempt++) {
            eval_device(mode, what, force, field, potential, scalars);
            check(cg_read_scalars(h_, scalars, host_scalars), "cg_read_scalars");
            if (!(host_scalars[1] < 0.0)) return;
        }
        throw std::runtime_error("cg_eval_device: evaluation invalid twice in a row");
    }
};

} // namespace CoulombGalore
