// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for boost::math::gamma_p, used only by the
// out-of-scope --predict mode (src/bayes.cpp:205) with a = 0.5, where P(1/2, x) = erf(sqrt(x)).
#pragma once
#include <cmath>
namespace boost {
namespace math {
inline double gamma_p(double a, double x) {
    if (a == 0.5) return std::erf(std::sqrt(x));
    // series for general a (not reached by the reference)
    double sum = 1.0 / a, term = sum;
    for (int n = 1; n < 500; n++) { term *= x / (a + n); sum += term; if (term < 1e-17 * sum) break; }
    return sum * std::exp(-x + a * std::log(x) - std::lgamma(a));
}
}  // namespace math
}  // namespace boost
