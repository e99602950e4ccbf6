/* gmrm_b200 -- C ABI of the B200-native Gibbs marker loop of gmrm.
 *
 * The reference (medical-genomics-group/gmrm) has no plugin / FFI layer: its contract is the
 * process interface (command line, input and output files).  This header
 * is the seam a maintainer of the reference would bind instead of the bodies of
 *   Bayes::load_genotype            src/bayes.cpp:867-900      -> gmrm_upload_bed
 *   Phenotype::read_file's results  src/phenotype.cpp:587-673  -> gmrm_set_phenotype
 *   read_group_index_file / .grm    src/bayes.cpp:830-853, src/options.cpp:222-286 -> gmrm_set_groups
 *   compute_markers_statistics      src/phenotype.cpp:466-556  -> gmrm_compute_marker_stats
 *   Bayes::dot_product              src/bayes.cpp:709-770      -> gmrm_dot_products (test hook)
 *   the iteration body of Bayes::process  src/bayes.cpp:340-656 -> gmrm_init_chain + gmrm_run_iteration
 *   outputs read at bayes.cpp:659-669     -> gmrm_get_betas / gmrm_get_components / gmrm_get_state
 *   the sums of Bayes::predict            src/bayes.cpp:87-214  -> gmrm_predict
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * GMRM_E* code, gmrm_last_error() gives the text; the caller owns all host buffers; one host
 * thread drives one engine; one engine == one GPU == one shard of markers.  There is NO CPU
 * fallback: without a usable CUDA device gmrm_create fails with GMRM_ENODEVICE.
 */
#ifndef GMRM_B200_H
#define GMRM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMRM_OK 0
#define GMRM_EINVAL (-1)    /* bad argument / wrong call order                      */
#define GMRM_ENODEVICE (-2) /* no CUDA device, or device is not sm_100               */
#define GMRM_ECUDA (-3)     /* a CUDA runtime call or kernel failed                  */
#define GMRM_ENOMEM (-4)
#define GMRM_ENCCL (-5)
#define GMRM_EREPLAY (-6)   /* replay mode asked for a variate the log does not hold */

typedef struct gmrm_engine gmrm_engine;

typedef struct {
    int32_t device;        /* CUDA device ordinal                                                       */
    int32_t N;             /* individuals (Dimensions::Nt, src/dimensions.cpp:14-24)                      */
    int32_t Mt;            /* total markers over all shards, after --trunc-markers                        */
    int32_t T;             /* traits == --phen-files entries                                             */
    int32_t G, K;          /* groups, mixture components (.grm rows / columns)                           */
    int32_t world_size;    /* GPUs (processes) sharing the chain; 1 for a single GPU                     */
    int32_t world_rank;    /* this engine's index, 0 .. world_size-1                                     */
    int32_t vranks;        /* TOTAL virtual ranks R: a run is semantically the reference under
                              `mpirun -n R` (block partition bayes.cpp:903-925, one marker per rank per
                              step, updates applied in rank order bayes.cpp:681-706).  Must be a multiple
                              of world_size; engine g owns virtual ranks [g*R/ws, (g+1)*R/ws).          */
    int32_t sync_rate;     /* marker-steps between cross-GPU residual exchanges; 1 == the reference       */
    int32_t shuffle;       /* --shuffle-markers (default 1)                                              */
    uint32_t seed;         /* --seed                                                                    */
    int32_t nsm;           /* 0 = one tile per SM of the device; >0 overrides (tests)                     */
    int32_t flags;         /* GMRM_FLAG_*                                                               */
} gmrm_config;

#define GMRM_FLAG_DEFAULT 0

/* Replay of the reference's own random variates for ONE iteration (north-star replay mode).
 * Layouts are those oracle/oracle.h emits.  R = vranks (all of them, not only this shard's),
 * Mm = ceil(Mt / R).  NaN marks "the reference drew nothing here"; needing such a value is
 * GMRM_EREPLAY.  Any pointer may be NULL to take that family from the production Philox streams. */
typedef struct {
    const int32_t* perm;      /* [R][Mm]   local marker index of rank r at step s (midx, phenotype.cpp:314-323), -1 past M_r */
    const double* u;          /* [Mm][R][T] component-selection uniforms (bayes.cpp:435)               */
    const double* z;          /* [Mm][R][T] standard normals behind the beta draws (bayes.cpp:456)     */
    const double* mu_draw;    /* [T]       new intercept of rank 0 (bayes.cpp:357)                      */
    const double* sigg_unit;  /* [T][G]    unit-scale gamma variates behind sigmaG (bayes.cpp:613)      */
    const double* pi_unit;    /* [T][G*K]  unit-scale gamma variates behind pi (phenotype.cpp:227-237)  */
    const double* sige_unit;  /* [T]       unit-scale gamma variate behind sigmaE (bayes.cpp:635)       */
} gmrm_replay;

/* Global (replicated) chain parameters after an iteration: what .csv records (xfiles.cpp:17-42). */
typedef struct {
    double* sigmag;   /* [T][G]   */
    double* sigmae;   /* [T]      */
    double* pi;       /* [T][G*K] */
    double* mu;       /* [T]      */
    int32_t* m0;      /* [T][G]   */
    int32_t* cass;    /* [T][G*K] component counts of the last iteration, summed over shards */
} gmrm_state;

typedef struct {
    double marker_loop_ms;   /* device time of the marker loop of the last iteration (CUDA events)   */
    double iteration_ms;     /* device time of the whole last iteration                               */
    double dot_kernel_ms;    /* summed device time of the dot kernel launches (only if timing detail on) */
    double sample_kernel_ms; /* same for the sampler kernel                                                   */
    double update_kernel_ms; /* same for the residual-update kernel                                           */
    double exchange_ms;      /* same for the cross-GPU exchange (all-reduce + merge), 0 on one GPU             */
    double allreduce_ms;     /* the all-reduce part of exchange_ms (includes waiting for the slowest shard)    */
    int64_t launches;        /* kernels launched by the last iteration                                */
    int64_t steps;           /* marker-steps of the last iteration                                    */
    int64_t published;       /* marker updates with dbeta != 0 in the last iteration (this shard)     */
} gmrm_timing;

const char* gmrm_last_error(void);
const char* gmrm_version(void);

int gmrm_create(const gmrm_config* cfg, gmrm_engine** out);
void gmrm_destroy(gmrm_engine* e);

/* This engine's shard [marker_begin, marker_begin + marker_count) of the Mt markers, the HBM
 * layout chosen (bytes per column, individuals per lane, tiles), and HBM bytes it will hold. */
int gmrm_shard_info(const gmrm_engine* e, int32_t* marker_begin, int32_t* marker_count,
                    int64_t* column_stride_bytes, int32_t* individuals_per_lane, int32_t* tiles);

/* --- genotypes ------------------------------------------------------------------------------
 * gmrm_upload_bed: PLINK SNP-major bytes exactly as they sit in the .bed after its 3 magic bytes
 * (mbytes = ceil(N/4) per marker; src/bayes.cpp:776,882).  `marker_begin` is a GLOBAL marker
 * index; the range must lie inside this engine's shard.  May be called repeatedly with chunks.
 * The bytes are copied to the device and transcoded there into the tile-planar layout. */
int gmrm_upload_bed(gmrm_engine* e, const uint8_t* bed, int32_t marker_begin, int32_t marker_count);
/* Pinned host memory: a gmrm_upload_bed source allocated here is copied by DMA, chunk k+1 overlapping the transcode
 * of chunk k (.bed ingestion at 114 GB).  Any other host pointer works too, more slowly. */
void* gmrm_host_alloc(size_t bytes);
void gmrm_host_free(void* p);
/* Synthetic genotypes generated ON the device (MAF ~ U(maf_lo, maf_hi) per marker,
 * Binomial(2, p) dosages, optional missing), through the same transcode as gmrm_upload_bed. */
int gmrm_generate_bed(gmrm_engine* e, uint32_t seed, double maf_lo, double maf_hi, double missing_rate);
/* Inverse transcode back to PLINK bytes (bit-exact round trip; parity tests, CPU-baseline sample). */
int gmrm_download_bed(gmrm_engine* e, uint8_t* bed_out, int32_t marker_begin, int32_t marker_count);
/* Call once after the last upload/generate: builds the per-marker missing-genotype lists. */
int gmrm_finalize_bed(gmrm_engine* e);

/* --- phenotypes, groups ---------------------------------------------------------------------
 * eps0: N doubles, the centred-scaled phenotype with 0 at NA (Phenotype::read_file's epsilon_);
 * mask4: ceil(N/4) bytes, bit k of byte i set when individual 4i+k is observed. */
int gmrm_set_phenotype(gmrm_engine* e, int32_t trait, const double* eps0, const uint8_t* mask4, int32_t nonas);
/* group_index: Mt ints (all markers); cva: G*K mixture variances, cva[g*K] == 0. */
int gmrm_set_groups(gmrm_engine* e, const int32_t* group_index, const double* cva);

/* --- one-off marker statistics (phenotype.cpp:466-556) --------------------------------------- */
int gmrm_compute_marker_stats(gmrm_engine* e);
int gmrm_get_marker_stats(gmrm_engine* e, int32_t trait, double* mave, double* msig); /* shard-local, marker_count each */

/* --- test hook: Bayes::dot_product for n shard-local markers against the current residuals.
 * out[i*T + t] = msig*(sum a*eps - mave*sum b*eps) of marker local_ids[i] and trait t. */
int gmrm_dot_products(gmrm_engine* e, const int32_t* local_ids, int32_t n, double* out);
/* --- test hooks: decode the HBM layout back into the reference's table values, on the GPU.
 * a/b: N doubles each, dotp_lut_a / dotp_lut_b of every individual of one marker (src/dotp_lut.hpp);
 * na: N doubles, na_lut of every individual of one trait (src/na_lut.hpp). */
int gmrm_decode_marker(gmrm_engine* e, int32_t local_id, double* a, double* b);
int gmrm_decode_namask(gmrm_engine* e, int32_t trait, double* na);
/* --- test hook: apply one published update (Phenotype::update_epsilon) to trait t. */
int gmrm_apply_update(gmrm_engine* e, int32_t trait, int32_t local_id, double dbeta);

/* --- the chain --------------------------------------------------------------------------------
 * gmrm_init_chain: bayes.cpp:322-335 (+ sigmaE start, phenotype.cpp:432-459).  sigmag_init [T][G]
 * (replay) or NULL (Beta(1,1) from the production stream). */
int gmrm_init_chain(gmrm_engine* e, const double* sigmag_init);
/* One Gibbs iteration `it` (1-based), bayes.cpp:340-656, asynchronous on the engine's stream
 * except for the final small read-back of the global parameters. */
int gmrm_run_iteration(gmrm_engine* e, int32_t it, const gmrm_replay* replay);
/* The same in two halves: _async enqueues the iteration and returns at once (the host is free to write the previous
 * iteration's files, bayes.cpp:659-669, while the GPU works); gmrm_wait_iteration waits for it and reports its errors and
 * timings.  gmrm_run_iteration == _async + wait.  At most one iteration may be enqueued. */
int gmrm_run_iteration_async(gmrm_engine* e, int32_t it, const gmrm_replay* replay);
int gmrm_wait_iteration(gmrm_engine* e);

int gmrm_get_state(gmrm_engine* e, gmrm_state* out);
int gmrm_get_betas(gmrm_engine* e, int32_t trait, double* betas);          /* shard-local, marker_count */
int gmrm_get_components(gmrm_engine* e, int32_t trait, int32_t* comp);     /* shard-local, marker_count */
int gmrm_get_epsilon(gmrm_engine* e, int32_t trait, double* eps);          /* N doubles */
/* Asynchronous variant of get_betas / get_components (output staging for thin-rate-1 runs):
 * stage = device snapshot + device-to-host copy on a second stream into pinned memory; fetch = wait + hand out.
 * A gmrm_run_iteration between the two overlaps the copy with the next iteration. */
int gmrm_stage_outputs(gmrm_engine* e);
int gmrm_fetch_outputs(gmrm_engine* e, int32_t trait, double* betas, int32_t* comp);
/* the global parameters snapshotted by the same gmrm_stage_outputs call (gmrm_get_state without a blocking copy) */
int gmrm_fetch_state(gmrm_engine* e, gmrm_state* out);
int gmrm_get_timing(gmrm_engine* e, gmrm_timing* out);
int gmrm_set_timing_detail(gmrm_engine* e, int32_t level); /* 0: totals; 1: + step kernel (2 events/step); 2: every phase (6 events/step), the residual update of every step as its own launch */

/* --- association pass: Bayes::predict, src/bayes.cpp:14-284, for one trait on this GPU's shard.
 * The reference's MPI ranks are the engine's virtual ranks (cfg.vranks): block r of markers is tested against
 * y_k = y - (g - g_r), i.e. only the OTHER blocks' genetic values are removed from the phenotype (bayes.cpp:141-147),
 * so vranks = 1 reproduces `mpirun -n 1`.  Needs gmrm_compute_marker_stats; does not touch the chain's state.
 *   y          N doubles: the centred-scaled phenotype (what gmrm_set_phenotype took; NA entries are ignored)
 *   beta_mean  marker_count doubles: mean of the .bet history for this shard's markers (bayes.cpp:59-78)
 *   keep       marker_count bytes or NULL: 0 = the marker's id is absent from the reference .bim -> skipped (95-100)
 *   g          N doubles or NULL: genetic values summed over all blocks and GPUs (bayes.cpp:136)
 *   beta, tdist, se, pval   marker_count doubles each or NULL (bayes.cpp:199-208); NaN for skipped markers */
int gmrm_predict(gmrm_engine* e, int32_t trait, const double* y, const double* beta_mean, const uint8_t* keep, double* g,
                 double* beta, double* tdist, double* se, double* pval);

/* The genetic values alone (bayes.cpp:87-136; scale(X) %*% beta of example/data_sim.R:33): g[i] = na_i * sum over ALL markers
 * of every shard of ((a - mave) msig) beta.  beta: marker_count doubles (this shard's markers); g: N doubles, identical on
 * every GPU (one all-reduce).  Needs gmrm_compute_marker_stats. */
int gmrm_genetic_values(gmrm_engine* e, int32_t trait, const double* beta, double* g);

/* --- test hook without a device: launch plan of the step kernel and the rows each CTA owns in each pass */
int gmrm_debug_step_plan(int32_t N, int32_t nsm, int32_t V, int32_t T, int32_t* traits_per_launch, int32_t* rows_per_pass,
                         int32_t* npass, int32_t* smem_bytes, int32_t* nrows, int32_t* ranges);

/* --- test hook without a device: where the step kernel's lanes find their genotype bytes.  A CTA's share of a column in a pass
 * of `rows` (1..5) rows is one chunk of rows * 64 bytes; the byte at the returned offset is byte k (0..3) of the word that lane
 * `lane16` (0..15, of the 16 lanes serving a marker) looks up in table slot `slot` (0..rows-1).  Negative: bad argument. */
int gmrm_debug_chunk_offset(int32_t rows, int32_t slot, int32_t lane16, int32_t k);

/* --- multi-GPU (one process per GPU): rank 0 makes the id, the launcher broadcasts it. */
int gmrm_comm_unique_id(uint8_t id[128]);
int gmrm_comm_init(gmrm_engine* e, const uint8_t id[128]);
/* Exchange at sync_rate 1 (replaces Allgather(bool) + Allgatherv(dbetas) + Allgatherv(bed column), bayes.cpp:500-545):
 * the sampler kernel of every GPU pushes its compacted list of published (dbeta*msig, mave, column) items into the
 * peers' buffers over NVLink and raises a flag there; every GPU then applies every update itself, reading the other
 * shards' columns from peer memory.  After gmrm_finalize_bed each engine exports its buffers (genotypes, missing
 * lists, list buffer, flags) and imports every peer's -- as CUDA IPC handles across processes (5 x 64 bytes), or as
 * pointers inside one process.  GMRM_EXCHANGE=nccl falls back to an NCCL all-gather of the lists. */
int gmrm_comm_export_buffers(gmrm_engine* e, uint8_t handles[384]);
int gmrm_comm_import_buffers(gmrm_engine* e, int32_t rank, const uint8_t handles[384]);
int gmrm_comm_local_buffers(gmrm_engine* e, void* ptrs[6]);
int gmrm_comm_set_peer_buffers(gmrm_engine* e, int32_t rank, int32_t peer_device, void* const ptrs[6]);

#ifdef __cplusplus
}
#endif
#endif
