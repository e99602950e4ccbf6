// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <boost/random.hpp>.
//
// Boost.Random is neither vendored by the reference nor installed here (SURVEY.md 8c;
// version unpinned upstream).  These classes give the reference's Distributions
// (src/distributions.hpp:5-61) something to compile against.  The engine is
// std::mt19937 (word-for-word the same generator as boost::mt19937); the distribution
// transforms are libstdc++'s, which are NOT Boost's -- so every variate is LOGGED
// (shim_hooks.h) and the GPU path is compared in replay mode, never by seed.
#pragma once
#include <cmath>
#include <random>
#include "shim_hooks.h"

namespace boost {
namespace random {

typedef std::mt19937 mt19937;

namespace shim_detail {
inline void log_rec(char tag, const double* v, int n) {
    char buf[1 + 3 * sizeof(double)];
    buf[0] = tag;
    __builtin_memcpy(buf + 1, v, n * sizeof(double));
    gmrm_shim_log(buf, 1 + n * sizeof(double));
}
}  // namespace shim_detail

template <class T = double>
class gamma_distribution {
    T a_, b_;
public:
    typedef T result_type;
    gamma_distribution(T a = 1, T b = 1) : a_(a), b_(b) {}
    template <class E> T operator()(E& e) {
        std::gamma_distribution<T> g(a_, T(1));
        T unit = g(e);
        double rec[3] = {double(a_), double(b_), double(unit)};
        shim_detail::log_rec('G', rec, 3);
        return unit * b_;
    }
};

template <class T = double>
class beta_distribution {
    T a_, b_;
public:
    typedef T result_type;
    beta_distribution(T a = 1, T b = 1) : a_(a), b_(b) {}
    template <class E> T operator()(E& e) {
        std::gamma_distribution<T> ga(a_, T(1)), gb(b_, T(1));
        T x = ga(e), y = gb(e);
        T v = x / (x + y);
        double rec[3] = {double(a_), double(b_), double(v)};
        shim_detail::log_rec('B', rec, 3);
        return v;
    }
};

template <class T = double>
class normal_distribution {
    T mean_, sd_;
public:
    typedef T result_type;
    normal_distribution(T mean = 0, T sd = 1) : mean_(mean), sd_(sd) {}
    template <class E> T operator()(E& e) {
        std::normal_distribution<T> n(T(0), T(1));
        T z = n(e);
        double rec[3] = {double(mean_), double(sd_), double(z)};
        shim_detail::log_rec('N', rec, 3);
        return mean_ + sd_ * z;
    }
};

template <class T = double>
class uniform_real_distribution {
    T lo_, hi_;
public:
    typedef T result_type;
    uniform_real_distribution(T lo = 0, T hi = 1) : lo_(lo), hi_(hi) {}
    template <class E> T operator()(E& e) {
        std::uniform_real_distribution<T> u(lo_, hi_);
        T v = u(e);
        double rec[1] = {double(v)};
        shim_detail::log_rec('U', rec, 1);
        return v;
    }
};

template <class T = int>
class uniform_int {
    T lo_, hi_;
public:
    typedef T result_type;
    uniform_int(T lo = 0, T hi = 9) : lo_(lo), hi_(hi) {}
    template <class E> T operator()(E& e) {
        std::uniform_int_distribution<T> u(lo_, hi_);
        return u(e);
    }
    // [0, n): what boost::variate_generator<..., uniform_int<>>::operator()(n) yields,
    // used by range::random_shuffle (src/phenotype.cpp:317-321)
    template <class E> T operator()(E& e, T n) {
        std::uniform_int_distribution<T> u(0, n - 1);
        return u(e);
    }
};

template <class Engine, class Dist>
class variate_generator {
    Engine eng_;   // Engine is a reference type (boost::mt19937&) at every call site
    Dist dist_;
public:
    typedef typename Dist::result_type result_type;
    variate_generator(Engine e, Dist d) : eng_(e), dist_(d) {}
    result_type operator()() { return dist_(eng_); }
    template <class T> result_type operator()(T n) { return dist_(eng_, n); }
};

}  // namespace random

using random::mt19937;
using random::normal_distribution;
using random::uniform_int;
using random::variate_generator;
}  // namespace boost
