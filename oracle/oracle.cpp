// TEST INFRASTRUCTURE ONLY -- see oracle.h.  CPU restatement of the reference's Gibbs marker
// loop (medical-genomics-group/gmrm), plain scalar C++, one thread.  Every function cites the
// reference lines it follows.  Nothing in gmrm_b200/ links, includes or calls this file.
//
// Parity pin: replay mode consumes the variate logs written by the reference itself
// (oracle/_ref/gmrm_ref) and CHECKS, record by record, that the (mean, sd) of every normal and
// the (shape, scale) of every gamma the reference asked for equal what this restatement computes
// at the same point -- so the arithmetic is pinned per marker, not only through final outputs.
#include "oracle.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <regex>
#include <string>
#include <vector>

#include "../gmrm_b200/csrc/gmrm_rng.h"  // production Philox streams (rng_mode 1 follows them)

namespace {

std::string g_err;
const double kNaN = std::numeric_limits<double>::quiet_NaN();

// ---- decode tables: src/lut/mk_lut.cpp:25-32 (a), 54-61 (b); src/lut/mk_lut_na.cpp:25-29
double LUT_A[1024], LUT_B[1024], LUT_NA[64];
bool lut_ready = false;
void build_luts() {
    if (lut_ready) return;
    for (int byte = 0; byte < 256; byte++) {
        for (int k = 0; k < 4; k++) {
            const int code = (byte >> (2 * k)) & 3;
            LUT_A[byte * 4 + k] = code == 0 ? 2.0 : code == 2 ? 1.0 : 0.0;  // 00->2, 01->0 (missing), 10->1, 11->0
            LUT_B[byte * 4 + k] = code == 1 ? 0.0 : 1.0;                     // non-missing indicator
        }
    }
    for (int nib = 0; nib < 16; nib++)
        for (int k = 0; k < 4; k++) LUT_NA[nib * 4 + k] = (nib >> k) & 1 ? 1.0 : 0.0;
    lut_ready = true;
}

// ---- replay log reader (record format: oracle/ref_shim/shim_hooks.h)
struct LogReader {
    std::vector<unsigned char> buf;
    size_t pos = 0;
    bool load(const std::string& path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) return false;
        buf.assign(std::istreambuf_iterator<char>(f), {});
        pos = 0;
        return true;
    }
    bool expect(char tag, double* v, int n) {
        if (pos + 1 + n * sizeof(double) > buf.size() || buf[pos] != (unsigned char)tag) return false;
        memcpy(v, &buf[pos + 1], n * sizeof(double));
        pos += 1 + n * sizeof(double);
        return true;
    }
    bool expect_perm(std::vector<int>& out) {
        if (pos + 5 > buf.size() || buf[pos] != 'P') return false;
        int32_t n;
        memcpy(&n, &buf[pos + 1], 4);
        if (pos + 5 + (size_t)n * 4 > buf.size()) return false;
        out.resize(n);
        memcpy(out.data(), &buf[pos + 5], (size_t)n * 4);
        pos += 5 + (size_t)n * 4;
        return true;
    }
};

struct RankTrait {            // what class Phenotype holds per rank (src/phenotype.hpp:12-57)
    std::vector<double> betas;
    std::vector<int> comp, midx;
    std::vector<int> cass;    // [G*K]
    double mu = 0.0;
};

struct Gibbs {
    const OracleCfg& c;
    OracleOut& o;
    const uint8_t* bed;
    const double* eps0;
    const uint8_t* mask4;
    const int* nonas;
    const int* group_index;
    const double* cva;
    int im4, mbytes, Mm = 0;
    std::vector<int> S, M;                       // per rank block (bayes.cpp:903-925)
    std::vector<std::vector<RankTrait>> rt;      // [R][T]
    std::vector<std::vector<double>> eps;        // [nrep*T][im4*4]
    std::vector<std::vector<double>> dl;         // [nrep*T][im4*4] pending residual deltas (sync_rate > 1)
    std::vector<std::vector<double>> mave, msig; // [T][Mt]
    std::vector<double> cvai;                    // [G*K]
    std::vector<int> mtotgrp;                    // [G]
    std::vector<double> sigmag, pi_est;          // [T*G], [T*G*K]   (rank 0's, broadcast: bayes.cpp:626,649)
    std::vector<double> sigmae;                  // [T]
    std::vector<int> m0;                         // [T*G]
    std::vector<LogReader> logs;                 // [R]
    const double V0E = 0.0001, S02E = 0.0001, V0G = 0.0001, S02G = 0.0001;  // bayes.hpp:14-17

    Gibbs(const OracleCfg& cfg, OracleOut& out) : c(cfg), o(out) {}

    int rep_of(int r) const { return (int)((long)r * c.nrep / c.R); }
    double* eps_of(int rep, int t) { return eps[(size_t)rep * c.T + t].data(); }

    void check_pair(double got_a, double got_b, double want_a, double want_b) {
        auto rel = [](double g, double w) { return std::fabs(g - w) / std::fmax(std::fabs(w), 1e-300); };
        double e = std::fmax(want_a == got_a ? 0.0 : rel(got_a, want_a), want_b == got_b ? 0.0 : rel(got_b, want_b));
        if (e > o.max_log_relerr) o.max_log_relerr = e;
        o.n_log_checked++;
    }

    // ---- draws.  Replay: consume the rank's log, checking the parameters the reference used.
    bool draw_beta_init(int r, int t, int g, double& v) {          // bayes.cpp:327
        if (c.rng_mode == 0) {
            double rec[3];
            if (!logs[r].expect('B', rec, 3)) return fail("expected beta record", r);
            v = rec[2];
        } else {
            v = gmrm::draw_uniform(c.seed, gmrm::STREAM_SIGMAG0, 0, g, t);   // Beta(1,1) == U(0,1)
        }
        return true;
    }
    bool draw_norm(int r, double mean, double sd, double& z, uint32_t stream, uint32_t it, uint32_t id, uint32_t t) {
        if (c.rng_mode == 0) {
            double rec[3];
            if (!logs[r].expect('N', rec, 3)) return fail("expected normal record", r);
            check_pair(mean, sd, rec[0], rec[1]);
            z = rec[2];
        } else {
            z = gmrm::draw_normal(c.seed, stream, it, id, t);
        }
        return true;
    }
    bool draw_unif(int r, double& u, uint32_t it, uint32_t marker, uint32_t t) {
        if (c.rng_mode == 0) {
            double rec[1];
            if (!logs[r].expect('U', rec, 1)) return fail("expected uniform record", r);
            u = rec[0];
        } else {
            u = gmrm::draw_uniform(c.seed, gmrm::STREAM_SAMPLER_U, it, marker, t);
        }
        return true;
    }
    bool draw_gamma_unit(int r, double shape, double scale, double& unit, uint32_t stream, uint32_t it, uint32_t id, uint32_t t) {
        if (c.rng_mode == 0) {
            double rec[3];
            if (!logs[r].expect('G', rec, 3)) return fail("expected gamma record", r);
            check_pair(shape, scale, rec[0], rec[1]);
            unit = rec[2];
        } else {
            unit = gmrm::draw_gamma(shape, c.seed, stream, it, id, t);
        }
        return true;
    }
    bool fail(const char* what, int r) {
        char b[256];
        snprintf(b, sizeof b, "replay log of rank %d: %s at byte %zu", r, what, logs[r].pos);
        g_err = b;
        return false;
    }

    // Phenotype::offset_epsilon, phenotype.cpp:395-411
    void offset_epsilon(double* e, const uint8_t* m4, double off) {
        for (int i = 0; i < im4; i++)
            for (int j = 0; j < 4; j++) e[i * 4 + j] += off * LUT_NA[m4[i] * 4 + j];
    }

    int run();
};

int Gibbs::run() {
    build_luts();
    const int N = c.N, Mt = c.Mt, T = c.T, G = c.G, K = c.K, R = c.R;
    im4 = (N + 3) / 4;
    mbytes = im4;  // bayes.cpp:776
    S.resize(R); M.resize(R);
    for (int r = 0; r < R; r++) oracle_block_of_markers(Mt, R, r, &S[r], &M[r], &Mm);

    cvai.assign(G * K, 0.0);
    for (int g = 0; g < G; g++)
        for (int k = 1; k < K; k++) cvai[g * K + k] = 1.0 / cva[g * K + k];  // options.cpp:282

    mtotgrp.assign(G, 0);
    for (int i = 0; i < Mt; i++) mtotgrp[group_index[i]]++;  // bayes.cpp:807-809

    // pi_prior, bayes.hpp:34-47
    std::vector<double> pi_prior(G * K, 0.5);
    for (int g = 0; g < G; g++) {
        double sum_cva = 0.0;
        for (int j = 0; j < K - 1; j++) sum_cva += cva[g * K + j + 1];
        for (int j = 1; j < K; j++) pi_prior[g * K + j] = pi_prior[g * K] * cva[g * K + j] / sum_cva;
    }

    mave.assign(T, {}); msig.assign(T, {});
    for (int t = 0; t < T; t++) {
        mave[t].resize(Mt); msig[t].resize(Mt);
        oracle_marker_stats(bed, N, Mt, mask4 + (size_t)t * im4, nonas[t], mave[t].data(), msig[t].data());
    }

    if (c.rng_mode == 0) {
        logs.resize(R);
        for (int r = 0; r < R; r++)
            if (!logs[r].load(std::string(c.replay_dir) + "/rank" + std::to_string(r) + ".bin")) {
                g_err = "cannot read replay log for rank " + std::to_string(r);
                return -2;
            }
    }

    // ---- Bayes::process prologue, bayes.cpp:322-335
    rt.assign(R, std::vector<RankTrait>(T));
    sigmag.assign(T * G, 0.0); pi_est.assign(T * G * K, 0.0); sigmae.assign(T, 0.0); m0.assign(T * G, 0);
    for (int r = 0; r < R; r++)
        for (int t = 0; t < T; t++) {
            RankTrait& p = rt[r][t];
            p.betas.assign(M[r], 0.0); p.comp.assign(M[r], 0); p.cass.assign(G * K, 0);
            p.midx.resize(M[r]);
            for (int i = 0; i < M[r]; i++) p.midx[i] = i;       // set_midx, phenotype.cpp:308-312
            for (int g = 0; g < G; g++) {
                double v;
                if (!draw_beta_init(r, t, g, v)) return -3;
                if (mtotgrp[g] == 0) v = 0.0;
                if (r == 0) sigmag[t * G + g] = v;               // Bcast of rank 0, bayes.cpp:332
            }
        }
    for (int t = 0; t < T; t++)
        for (int i = 0; i < G * K; i++) pi_est[t * G * K + i] = pi_prior[i];  // bayes.cpp:333
    if (o.sigmag_init) memcpy(o.sigmag_init, sigmag.data(), sizeof(double) * T * G);

    eps.assign((size_t)c.nrep * T, {});
    for (int rep = 0; rep < c.nrep; rep++)
        for (int t = 0; t < T; t++)
            eps[(size_t)rep * T + t].assign(eps0 + (size_t)t * im4 * 4, eps0 + (size_t)(t + 1) * im4 * 4);

    std::vector<double> denom(K), muk(K), logl(K);
    std::vector<double> dbetas((size_t)R * T * 3);
    std::vector<char> share(R);
    std::vector<int> mloc_of(R);

    for (int it = 1; it <= c.iterations; it++) {
        const size_t ih = (size_t)(it - 1);

        // ---- per-iteration prologue, bayes.cpp:347-368
        for (int r = 0; r < R; r++) {
            const int rep = rep_of(r);
            const bool owner = (r == 0) || rep_of(r - 1) != rep;   // first rank of a replica drives it
            for (int t = 0; t < T; t++) {
                RankTrait& p = rt[r][t];
                double* e = eps_of(rep, t);
                const uint8_t* m4 = mask4 + (size_t)t * im4;
                if (owner) offset_epsilon(e, m4, p.mu);              // bayes.cpp:351
                if (it == 1 && r == 0) {                             // update_epsilon_sigma, phenotype.cpp:448-457
                    double s = 0.0;
                    for (int i = 0; i < im4; i++)
                        for (int j = 0; j < 4; j++) s += e[i * 4 + j] * e[i * 4 + j] * LUT_NA[m4[i] * 4 + j];
                    sigmae[t] = s / double(nonas[t]) * 0.5;
                }
                // sample_norm_rng(): norm_rng(epssum/nonas, sigmae/nonas) with epssum == 0, phenotype.cpp:279-282
                const double mean = 0.0 / double(nonas[t]), sd = std::sqrt(sigmae[t] / double(nonas[t]));
                double z;
                if (!draw_norm(r, mean, sd, z, gmrm::STREAM_MU, it, 0, t)) return -4;
                const double new_mu = mean + sd * z;
                // production streams have one global mu; in replay each rank keeps its own (Appendix A)
                p.mu = new_mu;
                if (owner) offset_epsilon(e, m4, -p.mu);             // bayes.cpp:359
                if (r == 0) {
                    if (o.mu) o.mu[ih * T + t] = p.mu;
                    if (o.mu_draw) o.mu_draw[ih * T + t] = p.mu;
                }
                if (c.shuffle) {                                     // bayes.cpp:363-364
                    if (c.rng_mode == 0) {
                        std::vector<int> perm;
                        if (!logs[r].expect_perm(perm) || (int)perm.size() != M[r]) { fail("expected permutation", r); return -5; }
                        p.midx = perm;
                    } else {
                        // production: order is a pure function of (seed, it, rank); no dependence on last order
                        for (int s = 0; s < M[r]; s++) p.midx[s] = (int)gmrm::perm_at(s, M[r], c.seed, it, r);
                    }
                }
                std::fill(p.cass.begin(), p.cass.end(), 0);          // reset_cass; m0 reset below
            }
            if (o.perm)
                for (int s = 0; s < Mm; s++)
                    o.perm[(ih * R + r) * Mm + s] = s < M[r] ? rt[r][0].midx[s] : -1;
        }
        std::fill(m0.begin(), m0.end(), 0);                          // reset_m0

        // ---- marker loop, bayes.cpp:375-555
        for (int mrki = 0; mrki < Mm; mrki++) {
            std::fill(dbetas.begin(), dbetas.end(), 0.0);
            for (int r = 0; r < R; r++) {
                share[r] = 0;
                if (mrki >= M[r]) continue;
                const int mloc = rt[r][0].midx[mrki];               // all traits follow trait 0's order, bayes.cpp:384
                mloc_of[r] = mloc;
                const int mglo = S[r] + mloc;
                const int mgrp = group_index[mglo];
                for (int t = 0; t < T; t++) {
                    RankTrait& p = rt[r][t];
                    const size_t vi = ((ih * Mm + mrki) * R + r) * T + t;
                    if (o.u) o.u[vi] = kNaN;
                    if (o.z) o.z[vi] = kNaN;
                    const double sigg = sigmag[t * G + mgrp];
                    if (sigg == 0.0) { p.betas[mloc] = 0.0; continue; }          // bayes.cpp:396-400
                    const double beta = p.betas[mloc];
                    const double sige = sigmae[t];
                    const double sige_g = sige / sigg;                            // 403
                    const double sigg_e = 1.0 / sige_g;                           // 404
                    const double inv2sige = 1.0 / (2.0 * sige);                   // 405
                    for (int i = 1; i <= K - 1; i++)
                        denom[i - 1] = (double)(N - 1) + sige_g * cvai[mgrp * K + i];   // 413-416 (N, not nonas)
                    double num = oracle_dot(bed + (size_t)mglo * mbytes, mbytes, eps_of(rep_of(r), t),
                                            mave[t][mglo], msig[t][mglo]);       // 418
                    num += beta * double(nonas[t] - 1);                           // 421
                    if (it == 1 && o.num_first) o.num_first[((size_t)mrki * R + r) * T + t] = num;
                    for (int i = 1; i <= K - 1; i++) muk[i] = num / denom[i - 1]; // 425-426
                    for (int i = 0; i < K; i++) {                                 // 428-433
                        logl[i] = std::log(pi_est[(t * G + mgrp) * K + i]);
                        if (i > 0)
                            logl[i] += -0.5 * std::log(sigg_e * double(nonas[t] - 1) * cva[mgrp * K + i] + 1.0)
                                       + muk[i] * num * inv2sige;
                    }
                    double prob;
                    if (!draw_unif(r, prob, it, mglo, t)) return -6;              // 435
                    if (o.u) o.u[vi] = prob;
                    bool zero_acum = false;                                       // 437-445
                    double tmp1 = 0.0;
                    for (int i = 0; i < K; i++) {
                        if (std::fabs(logl[i] - logl[0]) > 700.0) zero_acum = true;
                        tmp1 += std::exp(logl[i] - logl[0]);
                    }
                    double acum = zero_acum ? 0.0 : 1.0 / tmp1;
                    double dbeta = beta;                                          // 448
                    for (int i = 0; i < K; i++) {                                 // 450-477
                        if (prob <= acum || i == K - 1) {
                            if (i == 0) {
                                p.betas[mloc] = 0.0;
                            } else {
                                const double mean = muk[i], sd = std::sqrt(sige / denom[i - 1]);  // distributions.hpp:48-53
                                double z;
                                if (!draw_norm(r, mean, sd, z, gmrm::STREAM_SAMPLER_N, it, mglo, t)) return -7;
                                if (o.z) o.z[vi] = z;
                                p.betas[mloc] = mean + sd * z;
                            }
                            p.cass[mgrp * K + i] += 1;
                            p.comp[mloc] = i;
                            break;
                        } else {
                            bool zero_inc = false;
                            for (int j = i + 1; j < K; j++)
                                if (std::fabs(logl[j] - logl[i + 1]) > 700.0) zero_inc = true;
                            if (!zero_inc) {
                                double esum = 0.0;
                                for (int k = 0; k < K; k++) esum += std::exp(logl[k] - logl[i + 1]);
                                acum += 1.0 / esum;
                            }
                        }
                    }
                    dbeta -= p.betas[mloc];                                        // 479
                    if (std::fabs(dbeta) > 0.0) {                                  // 483-487
                        share[r] = 1;
                        dbetas[((size_t)r * T + t) * 3 + 0] = dbeta;
                        dbetas[((size_t)r * T + t) * 3 + 1] = mave[t][mglo];
                        dbetas[((size_t)r * T + t) * 3 + 2] = msig[t][mglo];
                    }
                }
            }
            // exchange + Bayes::update_epsilon, bayes.cpp:495-553, 681-706: every replica applies every
            // published update, in rank order, per trait only where that trait's dbeta != 0
            if (c.sync_rate == 1) {
                for (int rep = 0; rep < c.nrep; rep++)
                    for (int r = 0; r < R; r++) {
                        if (!share[r]) continue;
                        const uint8_t* col = bed + (size_t)(S[r] + mloc_of[r]) * mbytes;
                        for (int t = 0; t < T; t++) {
                            const double* d3 = &dbetas[((size_t)r * T + t) * 3];
                            if (d3[0] != 0.0) oracle_update_eps(eps_of(rep, t), mask4 + (size_t)t * im4, im4, col, d3);
                        }
                    }
            } else {
                // sync_rate > 1 (superset of the reference, hydra lineage): a replica (GPU) applies its own ranks'
                // updates at once and the others' only at the exchange, every sync_rate marker-steps
                if (dl.empty()) dl.assign((size_t)c.nrep * T, std::vector<double>((size_t)im4 * 4, 0.0));
                for (int r = 0; r < R; r++) {
                    if (!share[r]) continue;
                    const int rep = rep_of(r);
                    const uint8_t* col = bed + (size_t)(S[r] + mloc_of[r]) * mbytes;
                    for (int t = 0; t < T; t++) {
                        const double* d3 = &dbetas[((size_t)r * T + t) * 3];
                        if (d3[0] == 0.0) continue;
                        oracle_update_eps(eps_of(rep, t), mask4 + (size_t)t * im4, im4, col, d3);
                        oracle_update_eps(dl[(size_t)rep * T + t].data(), mask4 + (size_t)t * im4, im4, col, d3);
                    }
                }
                if ((mrki + 1) % c.sync_rate == 0 || mrki == Mm - 1) {
                    for (int rep = 0; rep < c.nrep; rep++)
                        for (int rep2 = 0; rep2 < c.nrep; rep2++) {
                            if (rep2 == rep) continue;
                            for (int t = 0; t < T; t++) {
                                double* e = eps_of(rep, t);
                                const std::vector<double>& d = dl[(size_t)rep2 * T + t];
                                for (int i = 0; i < im4 * 4; i++) e[i] += d[i];
                            }
                        }
                    for (auto& d : dl) std::fill(d.begin(), d.end(), 0.0);
                }
            }
        }

        // ---- per-iteration epilogue, bayes.cpp:562-651
        for (int t = 0; t < T; t++) {
            std::vector<double> beta_sqn(G, 0.0);
            std::vector<int> cass(G * K, 0);
            for (int r = 0; r < R; r++) {                            // local sums then Allreduce in rank order, 566-590
                std::vector<double> loc(G, 0.0);
                for (int i = 0; i < M[r]; i++) loc[group_index[S[r] + i]] += rt[r][t].betas[i] * rt[r][t].betas[i];
                for (int g = 0; g < G; g++) beta_sqn[g] = r == 0 ? loc[g] : beta_sqn[g] + loc[g];
                for (int i = 0; i < G * K; i++) cass[i] += rt[r][t].cass[i];
            }
            std::vector<double> new_sigmag(sigmag.begin() + t * G, sigmag.begin() + (t + 1) * G);
            std::vector<double> new_pi(pi_est.begin() + t * G * K, pi_est.begin() + (t + 1) * G * K);
            double new_sigmae = sigmae[t];
            for (int r = 0; r < R; r++) {                            // every rank draws; rank 0's values are broadcast
                std::vector<double> sg(sigmag.begin() + t * G, sigmag.begin() + (t + 1) * G);
                std::vector<double> pe(pi_est.begin() + t * G * K, pi_est.begin() + (t + 1) * G * K);
                for (int g = 0; g < G; g++) {
                    if (r == 0) {
                        if (o.sigg_unit) o.sigg_unit[(ih * T + t) * G + g] = kNaN;
                        if (o.pi_unit) for (int k = 0; k < K; k++) o.pi_unit[((ih * T + t) * G + g) * K + k] = kNaN;
                    }
                    if (mtotgrp[g] == 0) continue;                   // 597-598
                    const int m0g = mtotgrp[g] - cass[g * K];        // 605
                    if (r == 0) m0[t * G + g] = m0g;
                    int cass_sum = 0;
                    for (int k = 0; k < K; k++) cass_sum += cass[g * K + k];
                    if (m0g == 0 || cass_sum == 0) { sg[g] = 0.0; continue; }   // 608-611
                    // inv_scaled_chisq_rng(a, b) = 1 / rgamma(a/2, 1/(a*b/2)), distributions.hpp:24-30; args 613
                    const double a = V0G + (double)m0g;
                    const double b = (beta_sqn[g] * (double)m0g + V0G * S02G) / (V0G + (double)m0g);
                    const double shape = 0.5 * a, scale = 1.0 / (0.5 * a * b);
                    double unit;
                    if (!draw_gamma_unit(r, shape, scale, unit, gmrm::STREAM_SIGMAG, it, g, t)) return -8;
                    sg[g] = 1.0 / (unit * scale);
                    if (r == 0 && o.sigg_unit) o.sigg_unit[(ih * T + t) * G + g] = unit;
                    double sum = 0.0;                                // update_pi_est_dirichlet, phenotype.cpp:227-237
                    for (int k = 0; k < K; k++) {
                        const double sh = (double)cass[g * K + k] + 1.0;
                        double u1;
                        if (!draw_gamma_unit(r, sh, 1.0, u1, gmrm::STREAM_PI, it, g * K + k, t)) return -9;
                        pe[g * K + k] = u1 * 1.0;
                        sum += pe[g * K + k];
                        if (r == 0 && o.pi_unit) o.pi_unit[((ih * T + t) * G + g) * K + k] = u1;
                    }
                    for (int k = 0; k < K; k++) pe[g * K + k] /= sum;
                }
                // epsilon_sumsqr over the first N slots, no NA mask, phenotype.cpp:251-261; sigmaE, bayes.cpp:635
                const double* e = eps_of(rep_of(r), t);
                double e_sqn = 0.0;
                for (int i = 0; i < N; i++) e_sqn += e[i] * e[i];
                const double a = V0E + (double)N, b = (e_sqn + V0E * S02E) / (V0E + (double)N);
                const double shape = 0.5 * a, scale = 1.0 / (0.5 * a * b);
                double unit;
                if (!draw_gamma_unit(r, shape, scale, unit, gmrm::STREAM_SIGMAE, it, 0, t)) return -10;
                if (r == 0) {
                    new_sigmag = sg; new_pi = pe; new_sigmae = 1.0 / (unit * scale);
                    if (o.sige_unit) o.sige_unit[ih * T + t] = unit;
                }
                if (c.rng_mode == 1) break;                          // production: one global stream
            }
            std::copy(new_sigmag.begin(), new_sigmag.end(), sigmag.begin() + t * G);       // Bcast, 626
            std::copy(new_pi.begin(), new_pi.end(), pi_est.begin() + t * G * K);           // Bcast, 648-650
            sigmae[t] = new_sigmae;                                                        // Bcast, 638-639

            // history == what .bet / .cpn / .csv record (bayes.cpp:659-669)
            for (int r = 0; r < R; r++)
                for (int i = 0; i < M[r]; i++) {
                    if (o.betas) o.betas[(ih * T + t) * Mt + S[r] + i] = rt[r][t].betas[i];
                    if (o.comp) o.comp[(ih * T + t) * Mt + S[r] + i] = rt[r][t].comp[i];
                }
            if (o.sigmag) memcpy(&o.sigmag[(ih * T + t) * G], &sigmag[t * G], sizeof(double) * G);
            if (o.pi) memcpy(&o.pi[(ih * T + t) * G * K], &pi_est[t * G * K], sizeof(double) * G * K);
            if (o.sigmae) o.sigmae[ih * T + t] = sigmae[t];
            if (o.m0) memcpy(&o.m0[(ih * T + t) * G], &m0[t * G], sizeof(int) * G);
        }
    }
    if (o.eps_final)
        for (int t = 0; t < T; t++) memcpy(o.eps_final + (size_t)t * im4 * 4, eps_of(0, t), sizeof(double) * im4 * 4);
    if (c.rng_mode == 0)
        for (int r = 0; r < R; r++)
            if (logs[r].pos != logs[r].buf.size()) { fail("log not fully consumed", r); return -11; }
    return 0;
}

}  // namespace

extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

void oracle_decode_tables(double* a, double* b, double* na) {
    build_luts();
    memcpy(a, LUT_A, sizeof LUT_A); memcpy(b, LUT_B, sizeof LUT_B); memcpy(na, LUT_NA, sizeof LUT_NA);
}

int oracle_read_phen(const char* path, int N, double* eps, uint8_t* mask4, int* nonas_out, int* nas_out) {
    std::ifstream infile(path);
    if (!infile.is_open()) return -1;
    const int im4 = (N + 3) / 4;
    std::string line;
    std::regex re("\\s+");
    std::vector<double> data;
    std::vector<uint8_t> m4;
    double sum = 0.0;
    int line_n = 0, nonas = 0, nas = 0;
    while (getline(infile, line)) {                                   // phenotype.cpp:599-628
        const int k = line_n % 4;
        if (k == 0) m4.push_back(0b00001111);
        std::sregex_token_iterator first{line.begin(), line.end(), re, -1}, last;
        std::vector<std::string> tokens{first, last};
        if (tokens.size() < 3) return -1;
        if (tokens[2] == "NA") {
            nas++;
            data.push_back(std::numeric_limits<double>::max());
            m4[line_n / 4] &= ~(1 << k);
        } else {
            nonas++;
            data.push_back(atof(tokens[2].c_str()));
            sum += atof(tokens[2].c_str());
        }
        line_n++;
    }
    if (nas + nonas != N) return -1;                                  // assert, phenotype.cpp:631
    if (line_n % 4 != 0)
        for (int i = line_n % 4; i < 4; i++) m4[line_n / 4] &= ~(1 << i);   // 633-638
    const double avg = sum / double(nonas);                           // 648
    for (int i = 0; i < im4 * 4; i++) eps[i] = 0.0;                   // reference leaves pad slots uninitialised
    double sqn = 0.0;
    for (size_t i = 0; i < data.size(); i++) {                        // 652-662
        if (data[i] == std::numeric_limits<double>::max()) eps[i] = 0.0;
        else { eps[i] = data[i] - avg; sqn += eps[i] * eps[i]; }
    }
    sqn = std::sqrt(double(nonas - 1) / sqn);                         // 663
    for (size_t i = 0; i < data.size(); i++) eps[i] *= sqn;           // 666-667
    memcpy(mask4, m4.data(), im4);
    *nonas_out = nonas; *nas_out = nas;
    return 0;
}

void oracle_marker_stats(const uint8_t* bed, int N, int M, const uint8_t* mask4, int nonas, double* mave, double* msig) {
    build_luts();
    const int im4 = (N + 3) / 4;
    for (int i = 0; i < M; i++) {                                     // phenotype.cpp:528-550
        const uint8_t* bedm = bed + (size_t)i * im4;
        double suma = 0.0, sumb = 0.0;
        for (int j = 0; j < im4; j++)
            for (int k = 0; k < 4; k++) {
                suma += LUT_A[bedm[j] * 4 + k] * LUT_NA[mask4[j] * 4 + k];
                sumb += LUT_B[bedm[j] * 4 + k] * LUT_NA[mask4[j] * 4 + k];
            }
        mave[i] = suma / sumb;
        double sumsqr = 0.0;
        for (int j = 0; j < im4; j++)
            for (int k = 0; k < 4; k++) {
                const double val = (LUT_A[bedm[j] * 4 + k] - mave[i]) * LUT_B[bedm[j] * 4 + k] * LUT_NA[mask4[j] * 4 + k];
                sumsqr += val * val;
            }
        msig[i] = 1.0 / std::sqrt(sumsqr / (double(nonas) - 1.0));
    }
}

double oracle_dot(const uint8_t* bed, int mbytes, const double* phen, double mu, double sigma_inv) {
    build_luts();
    double dpa = 0.0, dpb = 0.0;                                      // bayes.cpp:749-766
    for (int i = 0; i < mbytes; i++)
        for (int j = 0; j < 4; j++) {
            dpa += LUT_A[bed[i] * 4 + j] * phen[i * 4 + j];
            dpb += LUT_B[bed[i] * 4 + j] * phen[i * 4 + j];
        }
    return sigma_inv * (dpa - mu * dpb);
}

void oracle_update_eps(double* epsilon, const uint8_t* mask4, int im4, const uint8_t* bed, const double* dbeta) {
    build_luts();
    const double bs_ = dbeta[0] * dbeta[2];                           // phenotype.cpp:328
    const double mdb = -dbeta[1];                                     // 329
    for (int i = 0; i < im4; i++) {                                   // 378-389
        const int bedi = bed[i] * 4, masi = mask4[i] * 4;
        for (int j = 0; j < 4; j++) {
            const double a = LUT_A[bedi + j], b = LUT_B[bedi + j], m = LUT_NA[masi + j];
            epsilon[i * 4 + j] += (mdb * b + a) * bs_ * m;
        }
    }
}

void oracle_block_of_markers(int Mt, int nranks, int rank, int* S, int* M, int* Mm) {
    const int modu = Mt % nranks, size = Mt / nranks;                 // bayes.cpp:905-921
    *Mm = Mt % nranks != 0 ? size + 1 : size;
    int cum = 0;
    for (int i = 0; i < nranks; i++) {
        const int len = i < modu ? size + 1 : size;
        if (i == rank) { *M = len; *S = cum; }
        cum += len;
    }
}

int oracle_gibbs(const OracleCfg* cfg, const uint8_t* bed, const double* eps0, const uint8_t* mask4,
                 const int* nonas, const int* group_index, const double* cva, OracleOut* out) {
    g_err.clear();
    if (cfg->sync_rate < 1) { g_err = "oracle: sync_rate must be >= 1"; return -20; }
    Gibbs g(*cfg, *out);
    g.bed = bed; g.mask4 = mask4; g.nonas = nonas; g.group_index = group_index; g.cva = cva;
    g.eps0 = eps0;
    out->n_log_checked = 0; out->max_log_relerr = 0.0;
    return g.run();
}

// Bayes::predict, src/bayes.cpp:14-284, for one trait and R ranks emulated in one thread.
//   bed [Mt][ceil(N/4)], y = the trait's residuals as read (Phenotype::get_centered_and_scaled_y returns epsilon_,
//   phenotype.hpp:229-234), mave/msig [Mt] from compute_markers_statistics, beta_hist [niter][Mt] the .bet history,
//   keep [Mt] != 0 where the marker's id is present in the reference .bim (bayes.cpp:95-100, 176-180).
// Outputs: g [4*im4] total genetic values (the Allreduce of line 136), and per marker (NaN where skipped)
//   beta, tdist, se, pval [Mt] (199-208), sigma [R] the residual variance every rank uses (149-152).
// Quirk kept: a rank subtracts from y only the OTHER ranks' genetic values (146-147), so results depend on R.
int oracle_predict(int N, int Mt, int R, const uint8_t* bed, const uint8_t* mask4, int nonas, const double* y,
                   const double* mave, const double* msig, const double* beta_hist, int niter, const uint8_t* keep,
                   double* g_out, double* beta_out, double* tdist_out, double* se_out, double* pval_out, double* sigma_out) {
    g_err.clear();
    if (N < 1 || Mt < 1 || R < 1 || niter < 1) { g_err = "oracle_predict: bad sizes"; return -1; }
    build_luts();
    const int mbytes = (N + 3) / 4, im4 = mbytes;
    std::vector<double> beta_sum(Mt, 0.0);                                  // 59-78: mean of the recorded betas
    for (int i = 0; i < niter; i++)
        for (int j = 0; j < Mt; j++) beta_sum[j] += beta_hist[(size_t)i * Mt + j];
    for (int j = 0; j < Mt; j++) beta_sum[j] /= double(niter);

    std::vector<std::vector<double>> g_k(R, std::vector<double>((size_t)im4 * 4, 0.0));
    std::vector<double> g((size_t)im4 * 4, 0.0);
    for (int r = 0; r < R; r++) {                                           // 87-124
        int S, M, Mm;
        oracle_block_of_markers(Mt, R, r, &S, &M, &Mm);
        for (int mrki = 0; mrki < M; mrki++) {
            const int mglo = S + mrki;
            if (keep && !keep[mglo]) continue;
            const uint8_t* bedm = bed + (size_t)mglo * mbytes;
            for (int j = 0; j < im4; j++)
                for (int k = 0; k < 4; k++) {
                    const double val = (LUT_A[bedm[j] * 4 + k] - mave[mglo]) * LUT_B[bedm[j] * 4 + k] * LUT_NA[mask4[j] * 4 + k] * msig[mglo];
                    g_k[r][j * 4 + k] += val * beta_sum[mglo];
                }
        }
        for (size_t i = 0; i < g.size(); i++) g[i] += g_k[r][i];            // 136
    }
    if (g_out) for (size_t i = 0; i < g.size(); i++) g_out[i] = g[i];
    for (int j = 0; j < Mt; j++) {
        if (beta_out) beta_out[j] = kNaN;
        if (tdist_out) tdist_out[j] = kNaN;
        if (se_out) se_out[j] = kNaN;
        if (pval_out) pval_out[j] = kNaN;
    }
    std::vector<double> y_k((size_t)im4 * 4);
    for (int r = 0; r < R; r++) {
        int S, M, Mm;
        oracle_block_of_markers(Mt, R, r, &S, &M, &Mm);
        for (size_t i = 0; i < y_k.size(); i++) y_k[i] = ((int)i < N ? y[i] : 0.0) - (g[i] - g_k[r][i]);   // 141-147
        double sigma = 0.0;                                                 // 149-152
        for (int i = 0; i < N; i++) sigma += y_k[i] * y_k[i];
        sigma /= nonas;
        if (sigma_out) sigma_out[r] = sigma;
        for (int mrki = 0; mrki < M; mrki++) {                              // 170-214
            const int mglo = S + mrki;
            if (keep && !keep[mglo]) continue;
            const uint8_t* bedm = bed + (size_t)mglo * mbytes;
            double xtx = 0.0, xty = 0.0;
            for (int j = 0; j < im4; j++)
                for (int k = 0; k < 4; k++) {
                    const double val = LUT_A[bedm[j] * 4 + k] * LUT_B[bedm[j] * 4 + k] * LUT_NA[mask4[j] * 4 + k];
                    xtx += val * val;
                    xty += val * y_k[j * 4 + k];
                }
            const double beta = xty / xtx;
            const double tdist = xty / std::sqrt(sigma * xtx);
            const double se = beta / tdist;
            const double pval = 1.0 - std::erf(std::sqrt(tdist * tdist * 0.5));   // boost::math::gamma_p(0.5, x) = erf(sqrt(x))
            if (beta_out) beta_out[mglo] = beta;
            if (tdist_out) tdist_out[mglo] = tdist;
            if (se_out) se_out[mglo] = se;
            if (pval_out) pval_out[mglo] = pval;
        }
    }
    return 0;
}

}  // extern "C"
