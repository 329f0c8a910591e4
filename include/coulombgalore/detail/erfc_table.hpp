// coulombgalore/detail/erfc_table.hpp -- host-side construction of the device erfc/exp table of specfun.hpp.
#pragma once
#include "specfun.hpp"
#include <cmath>
#include <vector>

namespace CoulombGalore {
namespace detail {

// interleaved {erfc(x_k), exp(-x_k^2)}, x_k = k/128: x_k and x_k^2 are exact in binary, so both entries carry only
// glibc's own (< 1 ulp) error
inline std::vector<double> make_erfc_table() {
    std::vector<double> t(2 * (size_t)kErfcTableEntries);
    for (int k = 0; k < kErfcTableEntries; k++) {
        const double x = (double)k * kErfcTableStep;
        t[2 * (size_t)k] = std::erfc(x);
        t[2 * (size_t)k + 1] = std::exp(-(x * x));
    }
    return t;
}

} // namespace detail
} // namespace CoulombGalore
