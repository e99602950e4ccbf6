// Per-marker mixture sampler and per-iteration global draws.  Host + device; the CUDA kernels
// call these from one thread per (marker, trait) / per trait.
//
// Follows the sampling block of Bayes::process, src/bayes.cpp:389-492, and the
// epilogue src/bayes.cpp:594-650, INCLUDING the reference's quirks:
// N-1 in denom but nonas-1 in num/logl, beta_sqn*m0 in the sigmaG scale, N (not nonas) for sigmaE.
#pragma once
#include <cmath>
#include <cstdint>
#include "gmrm_rng.h"

namespace gmrm {

constexpr int kMaxK = 16;  // mixture components per group supported by the register arrays below

struct MarkerDraw {
    double beta_new;
    double dbeta;   // beta_old - beta_new (bayes.cpp:479)
    int comp;       // chosen component, -1 when the marker was skipped (sigmaG == 0)
    int need_z;     // 1 when a normal was consumed
};

// The per-marker mixture draw with the marker-independent pieces precomputed per (trait, group) by group_consts_kernel
// (kernels.cu): gc = [denom | log pi | -0.5 log(..) | sd], slot 0 of the first block = 1/(2 sigmaE).
// The arithmetic per marker is the reference's, term for term (bayes.cpp:421-477).
template <class ZDraw>
GMRM_HD MarkerDraw sample_marker_pre(double dot_raw, double beta, double sigmag, const double* gc, int K, int nonas,
                                     double u, ZDraw zdraw) {
    MarkerDraw r;
    r.need_z = 0;
    if (sigmag == 0.0) {                                     // bayes.cpp:396-400
        r.beta_new = 0.0; r.dbeta = 0.0; r.comp = -1;
        return r;
    }
    const double inv2sige = gc[0];
    const double *denom = gc, *logpi = gc + K, *hterm = gc + 2 * K, *sd = gc + 3 * K;   // denom[k] == reference denom[k-1]
    double muk[kMaxK], logl[kMaxK];
    const double num = dot_raw + beta * (double)(nonas - 1);                              // 421
    logl[0] = logpi[0];
    for (int i = 1; i < K; i++) {
        muk[i] = num / denom[i];                                                          // 425-426
        logl[i] = logpi[i] + (hterm[i] + muk[i] * num * inv2sige);                        // 428-433
    }
    bool zero_acum = false;                                                               // 437-445
    double tmp1 = 0.0;
    for (int i = 0; i < K; i++) {
        if (fabs(logl[i] - logl[0]) > 700.0) zero_acum = true;
        tmp1 += exp(logl[i] - logl[0]);
    }
    double acum = zero_acum ? 0.0 : 1.0 / tmp1;
    r.beta_new = 0.0; r.comp = K - 1;
    for (int i = 0; i < K; i++) {                                                         // 450-477
        if (u <= acum || i == K - 1) {
            if (i > 0) { r.beta_new = muk[i] + sd[i] * zdraw(); r.need_z = 1; }           // 456
            r.comp = i;
            break;
        } else {
            bool zero_inc = false;
            for (int j = i + 1; j < K; j++)
                if (fabs(logl[j] - logl[i + 1]) > 700.0) zero_inc = true;
            if (!zero_inc) {
                double esum = 0.0;
                for (int k = 0; k < K; k++) esum += exp(logl[k] - logl[i + 1]);
                acum += 1.0 / esum;
            }
        }
    }
    r.dbeta = beta - r.beta_new;                                                          // 479
    return r;
}

// inv_scaled_chisq_rng(a, b) = 1 / rgamma(a/2, 1/(a*b/2)) (distributions.hpp:24-30), given the
// unit-scale gamma variate of shape a/2.
GMRM_HD double inv_scaled_chisq_from_unit(double a, double b, double unit) {
    const double scale = 1.0 / (0.5 * a * b);
    return 1.0 / (unit * scale);
}

}  // namespace gmrm
