// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <boost/range/algorithm.hpp>.
// Only range::random_shuffle(container, gen) is used (src/phenotype.cpp:318,321).
// The permutation it produces is an arbitrary stand-in for Boost's, so the shuffled
// container is logged ('P' record) for replay.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>
#include "../../shim_hooks.h"

namespace boost {
namespace range {
template <class Container, class Gen>
Container& random_shuffle(Container& c, Gen& gen) {
    const long n = (long)c.size();
    for (long i = n - 1; i > 0; --i) {
        long j = (long)gen(i + 1);
        std::swap(c[i], c[j]);
    }
    std::vector<int32_t> rec(n + 1);
    rec[0] = (int32_t)n;
    for (long i = 0; i < n; i++) rec[i + 1] = (int32_t)c[i];
    const char tag = 'P';
    gmrm_shim_log(&tag, 1);
    gmrm_shim_log(rec.data(), rec.size() * sizeof(int32_t));
    return c;
}
}  // namespace range
}  // namespace boost
