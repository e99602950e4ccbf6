/* TEST INFRASTRUCTURE ONLY.  C ABI of the CPU oracle (oracle/oracle.cpp): a plain, single-
 * threaded restatement of the reference's Gibbs marker loop, each function citing the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library -- and only as the checker.
 * The product (gmrm_b200/) never links or calls it.
 *
 * Pinning: the restatement is checked (tests/test_oracle_vs_reference.py) against outputs of
 * the reference itself -- oracle/_ref/gmrm_ref, the unmodified reference sources built against
 * oracle/ref_shim -- by replaying the reference's own logged random variates, and against the
 * fixtures that script committed under tests/golden/.  The reference ships no golden vectors
 * of its own. */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int N, Mt, T, G, K;
    int R;            /* ranks: MPI ranks in the reference == virtual ranks on the GPU            */
    int nrep;         /* epsilon replicas (R in replay mode: one per rank, as the reference has)   */
    int iterations;
    int shuffle;      /* --shuffle-markers                                                         */
    int rng_mode;     /* 0 = replay the reference's variate logs, 1 = production Philox streams    */
    int sync_rate;    /* marker-steps between residual exchanges; 1 == the reference               */
    uint32_t seed;
    const char* replay_dir; /* rng_mode 0: directory holding rank<r>.bin                           */
} OracleCfg;

typedef struct {
    /* history, one slab per iteration (all may be NULL to skip) */
    double*  betas;      /* [it][T][Mt]  global marker order, as .bet (xfiles.hpp:24-37)           */
    int32_t* comp;       /* [it][T][Mt]  as .cpn                                                   */
    double*  sigmag;     /* [it][T][G]                                                             */
    double*  sigmae;     /* [it][T]                                                                */
    double*  pi;         /* [it][T][G*K]                                                           */
    double*  mu;         /* [it][T]      rank 0's intercept                                        */
    int32_t* m0;         /* [it][T][G]                                                             */
    double*  eps_final;  /* [T][4*ceil(N/4)]  replica 0 after the last iteration                   */
    /* dense variates, in the layout gmrm_b200's replay interface takes (include/gmrm_b200.h)      */
    int32_t* perm;       /* [it][R][Mm]   local marker index of rank r at step s, -1 past M_r      */
    double*  u;          /* [it][Mm][R][T] selection uniforms, NaN where none was drawn            */
    double*  z;          /* [it][Mm][R][T] standard normals behind beta draws, NaN where none      */
    double*  mu_draw;    /* [it][T]       rank 0's new mu                                           */
    double*  sigg_unit;  /* [it][T][G]    unit-scale gamma behind sigmaG, NaN where skipped        */
    double*  pi_unit;    /* [it][T][G*K]  unit-scale gammas behind pi, NaN where skipped           */
    double*  sige_unit;  /* [it][T]                                                                */
    double*  sigmag_init;/* [T][G]        initial sigmaG (after zeroing empty groups)              */
    /* diagnostics */
    double*  num_first;  /* [Mm][R][T]    'num' of every marker-step of iteration 1 (dot + beta*(nonas-1)) */
    int64_t  n_log_checked; /* replay: number of logged (mean, sd)/(shape, scale) pairs compared   */
    double   max_log_relerr; /* replay: worst relative mismatch of those pairs                     */
} OracleOut;

/* decode tables regenerated as src/lut/mk_lut.cpp:25-32,54-61 and src/lut/mk_lut_na.cpp:25-29 do */
void oracle_decode_tables(double* lut_a /*1024*/, double* lut_b /*1024*/, double* lut_na /*64*/);

/* Phenotype::read_file, src/phenotype.cpp:587-673.  eps has 4*ceil(N/4) slots (pad = 0),
 * mask4 has ceil(N/4) bytes.  Returns 0, or -1 if the file cannot be read / has != N rows. */
int oracle_read_phen(const char* path, int N, double* eps, uint8_t* mask4, int* nonas, int* nas);

/* PhenMgr::compute_markers_statistics scalar path, src/phenotype.cpp:528-550 */
void oracle_marker_stats(const uint8_t* bed, int N, int M, const uint8_t* mask4, int nonas,
                         double* mave, double* msig);

/* Bayes::dot_product scalar path, src/bayes.cpp:749-766 */
double oracle_dot(const uint8_t* bedcol, int mbytes, const double* eps, double mave, double msig);

/* Phenotype::update_epsilon scalar path, src/phenotype.cpp:326-329,375-390.  dbeta3 = {dbeta, mave, msig} */
void oracle_update_eps(double* eps, const uint8_t* mask4, int im4, const uint8_t* bedcol, const double* dbeta3);

/* Bayes::set_block_of_markers, src/bayes.cpp:903-925 */
void oracle_block_of_markers(int Mt, int nranks, int rank, int* S, int* M, int* Mm);

/* Bayes::process, src/bayes.cpp:318-677, for R ranks emulated in one thread.
 * bed: [Mt][ceil(N/4)] all markers; eps0: [T][4*ceil(N/4)]; mask4: [T][ceil(N/4)];
 * group_index: [Mt]; cva: [G][K].  Returns 0, or a negative code (replay log mismatch). */
int oracle_gibbs(const OracleCfg* cfg, const uint8_t* bed, const double* eps0, const uint8_t* mask4,
                 const int* nonas, const int* group_index, const double* cva, OracleOut* out);

/* Bayes::predict, src/bayes.cpp:14-284, one trait, R ranks emulated (the results depend on R: a rank removes only the
 * other ranks' genetic values from y, lines 146-147).  beta_hist [niter][Mt] is the .bet history; keep [Mt] (or NULL)
 * marks markers whose id is in the reference .bim.  g [4*ceil(N/4)]; beta/tdist/se/pval [Mt], NaN where skipped;
 * sigma [R]. */
int oracle_predict(int N, int Mt, int R, const uint8_t* bed, const uint8_t* mask4, int nonas, const double* y,
                   const double* mave, const double* msig, const double* beta_hist, int niter, const uint8_t* keep,
                   double* g, double* beta, double* tdist, double* se, double* pval, double* sigma);

const char* oracle_last_error(void);

#ifdef __cplusplus
}
#endif
