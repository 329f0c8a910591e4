// oracle/port_main.cpp -- TEST INFRASTRUCTURE ONLY.
// Builds oracle/libcg_port.so: the independent restatement (port_schemes.hpp) behind the cgo_api.h C ABI.
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

CGO_DEFINE_C_ABI("port")
