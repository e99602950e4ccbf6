// Counter-based random numbers for the production (non-replay) Gibbs path.
//
// The reference draws from two boost::mt19937 engines per trait and rank
// (src/distributions.hpp:5-61, seeding src/bayes.cpp:796-803).  A sequential engine cannot
// feed thousands of markers in flight, and Boost's distribution transforms are not available
// (SURVEY.md 8c), so the production path uses Philox4x32-10 keyed by what is being drawn:
//
//     key     = (seed, stream tag)
//     counter = (iteration, global marker or group id, trait, draw index)
//
// so a draw does not depend on which GPU, virtual rank or thread asks for it.  Replay mode
// bypasses this file entirely and feeds the reference's logged variates instead.
//
// Everything here is __host__ __device__ so the same bits are produced by the CUDA kernels
// and by host code (the CPU oracle includes this header to follow the production streams).
#pragma once
#include <cmath>
#include <cstdint>

#ifndef GMRM_UNROLL
#if defined(__CUDA_ARCH__)
#define GMRM_UNROLL _Pragma("unroll")
#else
#define GMRM_UNROLL            // host pass: the pragma is unknown to g++
#endif
#endif

#if defined(__CUDACC__)
#define GMRM_HD __host__ __device__ __forceinline__
#else
#define GMRM_HD inline
#endif

namespace gmrm {

enum RngStream : uint32_t {
    STREAM_SAMPLER_U = 0x53414d55u,  // per (it, marker, trait): the component-selection uniform (bayes.cpp:435)
    STREAM_SAMPLER_N = 0x53414d4eu,  // per (it, marker, trait): the beta normal (bayes.cpp:456)
    STREAM_MU        = 0x4d555f5fu,  // per (it, trait): intercept normal (bayes.cpp:357)
    STREAM_SIGMAG0   = 0x53473030u,  // per (trait, group): initial sigmaG ~ Beta(1,1) (bayes.cpp:327)
    STREAM_SIGMAG    = 0x53474d41u,  // per (it, group, trait): sigmaG gamma (bayes.cpp:613)
    STREAM_PI        = 0x50495f5fu,  // per (it, group*K+k, trait): Dirichlet gammas (phenotype.cpp:227-237)
    STREAM_SIGMAE    = 0x5347455fu,  // per (it, trait): sigmaE gamma (bayes.cpp:635)
    STREAM_PERM      = 0x5045524du,  // per (it, virtual rank): marker permutation (phenotype.cpp:314-323)
};

struct U4 { uint32_t x, y, z, w; };

GMRM_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al., SC'11), constants as published.
GMRM_HD U4 philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
GMRM_UNROLL
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return U4{c0, c1, c2, c3};
}

// 52 random bits -> (0,1), never 0 or 1: (k + 1/2) * 2^-52 is exact in binary64.
GMRM_HD double u01(uint32_t a, uint32_t b) {
    const double k = (double)(a >> 6) * 67108864.0 + (double)(b >> 6);
    return (k + 0.5) * (1.0 / 4503599627370496.0);
}

GMRM_HD double box_muller(const U4& r) {
    const double u1 = u01(r.x, r.y), u2 = u01(r.z, r.w);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

GMRM_HD double draw_uniform(uint32_t seed, uint32_t stream, uint32_t it, uint32_t id, uint32_t trait, uint32_t n = 0) {
    U4 r = philox4x32(seed, stream, it, id, trait, n);
    return u01(r.x, r.y);
}

GMRM_HD double draw_normal(uint32_t seed, uint32_t stream, uint32_t it, uint32_t id, uint32_t trait, uint32_t n = 0) {
    return box_muller(philox4x32(seed, stream, it, id, trait, n));
}

// Gamma(shape, 1), Marsaglia & Tsang (2000).  Draw index n advances by one Philox block per
// attempt (block n: normal from all four words; block n+1: the acceptance uniform, and for
// shape < 1 the boosting uniform in .z/.w).
GMRM_HD double draw_gamma(double shape, uint32_t seed, uint32_t stream, uint32_t it, uint32_t id, uint32_t trait) {
    const bool boost = shape < 1.0;
    const double a = boost ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double out = d;
    for (uint32_t n = 0; n < 2000u; n += 2) {
        const double z = box_muller(philox4x32(seed, stream, it, id, trait, n));
        const U4 r = philox4x32(seed, stream, it, id, trait, n + 1);
        double v = 1.0 + c * z;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = u01(r.x, r.y);
        if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) {
            out = d * v;
            if (boost) out *= pow(u01(r.z, r.w), 1.0 / shape);
            break;
        }
    }
    return out;
}

// Pseudo-random permutation of [0, n) evaluated point-wise: 4-round Feistel network on the
// smallest even-bit domain >= n, cycle-walked back into range.  Replaces the in-place
// random_shuffle of midx (phenotype.cpp:314-323): no storage, no sort, any thread can ask for
// position s of rank r's order in iteration it.
GMRM_HD uint32_t perm_at(uint32_t s, uint32_t n, uint32_t seed, uint32_t it, uint32_t rank) {
    if (n <= 1) return 0;
    uint32_t bits = 2;
    while (bits < 32 && (1ull << bits) < (uint64_t)n) bits += 2;
    const uint32_t half = bits / 2, mask = (1u << half) - 1u;
    uint32_t x = s;
    do {
        uint32_t L = x >> half, R = x & mask;
GMRM_UNROLL
        for (uint32_t round = 0; round < 4; round++) {
            const uint32_t F = philox4x32(seed, STREAM_PERM, it, rank, R, round).x & mask;
            const uint32_t t = L ^ F;
            L = R;
            R = t;
        }
        x = (L << half) | R;
    } while (x >= n);
    return x;
}

}  // namespace gmrm
