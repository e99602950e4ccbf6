// sm_100a kernels of the Gibbs marker loop.  See layout.h for the HBM layout and the table geometry,
// DESIGN.md for the roofline of each kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "gmrm_rng.h"
#include "layout.h"
#include "sampler.h"

namespace gmrm {

// Step-kernel shape: 16 warps of 128 registers, batches of 8 markers.  (Measured alternatives, tools/build_variant.sh: 12 and 14
// warps with more registers run the stream at the same speed; the hybrid plan below needs <= 12 warps for its registers.)
#ifndef GMRM_STEP_WARPS
#define GMRM_STEP_WARPS 16
#endif
#ifndef GMRM_STEP_BATCH
#define GMRM_STEP_BATCH 8
#endif
// Direct rows per pass the step kernel is built for (0: look-ups only -- the product build; 1: the hybrid plan, see
// StepParams::bed2 -- parity-tested on the GPU but 14 % SLOWER than look-ups alone at N = 458,000 (profiles/README.md), kept as
// a build variant: tools/build_variant.sh h12 -DGMRM_STEP_WARPS=12 -DGMRM_STEP_DIRECT=1)
#ifndef GMRM_STEP_DIRECT
#define GMRM_STEP_DIRECT 0
#endif
// Per-marker partial sums of a step live in shared memory, kPartSmemDoubles of them (16 KB next to the full set of table
// slots); a step of more (marker, trait) pairs is streamed in chunks of that size per pass, each chunk's sums being added to
// the CTA's slots of StepParams::partial (L2-resident) before the next chunk.
constexpr int kPartSmemDoubles = 2048;
constexpr int kBatch = GMRM_STEP_BATCH;   // markers per warp batch (one marker per half-warp: kBatch/2 pairs)
constexpr int kMaxGpus = 8;
constexpr int kPubCap = 128;           // published updates staged per round of the update phase
constexpr unsigned long long kXdSentinel = 0xFFF8DEADFFF8DEADull;   // "nothing here yet" in the exchange buffers (a NaN bit pattern)

struct PubEntry {   // one published update: Phenotype::update_epsilon's dbeta[3], phenotype.cpp:326-329
    double lam;     // dbeta * msig   (0 == nothing to apply)
    double mave;
};
// The step's published updates of one GPU and trait: a SEGMENTED list in virtual-rank order.  Every CTA of the sampler
// kernel samples kSegCap consecutive virtual ranks and writes their published items, compacted and in order, into its own
// segment -- a 16-byte header followed by up to kSegCap items -- with no dependency on any other CTA (and, across GPUs,
// straight into the peers' copies of the list).  The header's first word is  count | sequence << 8  (the sequence number of
// the sampled step, written last, after a system-scope fence): the consumer (step kernel) waits until every header carries
// the sequence number it expects -- that is the whole hand-shake, no flags, no last CTA -- then turns the counts into a
// prefix and walks the items of all segments of all GPUs in global virtual-rank order.
struct PubItem { double lam, mave; int32_t col, v; };     // col: column local to the publishing GPU
constexpr int kSegCap = 16;                                // virtual ranks per sampler CTA == items a segment can hold
constexpr int kSegDoubles = 2 + 3 * kSegCap;               // header + items, in doubles
__host__ __device__ constexpr int publist_segments(int V) { return (V + kSegCap - 1) / kSegCap; }
__host__ __device__ constexpr size_t publist_doubles(int V) { return (size_t)publist_segments(V) * kSegDoubles; }
__host__ __device__ constexpr unsigned long long seg_header(int count, unsigned long long seq) { return (unsigned long long)count | (seq << 8); }

// One marker-step on one GPU: (a) apply the updates published by the previous step to this CTA's rows of
// the residuals, (b) build the look-up tables of those rows, (c) stream the step's V columns through them.
struct StepParams {
    const uint8_t* bed;
    int64_t col_stride;
    int32_t nrows;
    const int32_t* cols;     // [V] shard-local column of each virtual rank this step, -1 = none
    int32_t V;               // 0: update only
    double* eps;             // [Ttot][npad]
    int64_t npad;
    int32_t Ttot, t0;        // this launch handles traits t0 .. t0+T-1 of Ttot
    int32_t rows_per_pass;   // table slots / T
    int32_t npass;           // step_npass(L, rows_per_pass)
    double* partial;         // [V][Ttot][nsm]   per-CTA partial sums of sum a*eps
    double* spart;           // [Ttot][nsm]      per-CTA sums of eps
    // pending updates (previous step), applied in virtual-rank order: pG lists of pV entries, list g published by
    // GPU g about ITS markers -- columns, genotypes and missing lists of list g are read from GPU g's buffers
    // (peer memory over NVLink when g is not this GPU)
    int32_t pG, pV;                      // lists, virtual ranks behind a list (0 lists: nothing pending)
    const double* plist;                 // [pG][Ttot][publist_doubles(pV)]: publist_segments(pV) segments each
    unsigned long long wait_seq;         // sequence number every segment header of the pending lists must carry
    const uint8_t* pbed[kMaxGpus];
    const uint32_t* pmiss_off[kMaxGpus];
    const uint32_t* pmiss_idx[kMaxGpus];
    // Row-sharded application of the pending lists (rs_world > 1: list exchange over several GPUs).  Every GPU holds the same
    // residuals; instead of every GPU applying every update to every row, GPU g applies ALL lists to 1/rs_world of each CTA's
    // rows (local row lr with lr % rs_world == rs_rank) and stores the updated rows into every GPU's residual array over
    // NVLink; same-index CTAs of the GPUs then tell each other through flags that their rows have landed.
    int32_t rs_world, rs_rank;
    unsigned long long row_seq;          // sequence number of this launch among the row-sharded ones (same on every GPU)
    double* peps[kMaxGpus];              // residual arrays of all GPUs (peer memory; [rs_rank] is this GPU's)
    unsigned long long* rflag_peer[kMaxGpus];   // GPU g's row-flag array [rs_world][nsm]: this GPU writes entry [rs_rank][cta]
    const unsigned long long* rflag_mine;       // this GPU's row-flag array: entry [g][cta] = sequence number GPU g's CTA `cta` has reached
    // Fused two-hop exchange of residual INCREMENTS (xd_world > 1; the default at sync_rate 1 on several GPUs).  Every GPU applies
    // only its OWN published list (local columns), as increments d_g of its CTAs' rows.  CTA c of every GPU owns the same rows;
    // they are cut into xd_world sub-slices.  Hop 1: CTA c of GPU g stores sub-slice o of d_g into GPU o's receive buffer over
    // NVLink.  GPU o's CTA c adds the xd_world increments of its sub-slice in GPU order to the old residuals (identical on every
    // GPU) and -- hop 2 -- stores the new residuals into every GPU's landing buffer, from where every CTA moves its rows into its
    // residual array: all replicas stay bit-identical, 2 x 7/8 of the residual array crosses NVLink per GPU and step.
    // There are NO flags and NO fences: both buffers are pre-filled with a sentinel (a NaN no increment or residual can equal),
    // the receiver polls the payload words themselves, 8 bytes at a time, and puts the sentinel back once it has them
    // (tools/p2p_micro.cu: 2 us per hop instead of 8-15 us with a system-scope fence per CTA).  Two copies of each buffer take
    // turns (parity of row_seq): a slot is rewritten two launches after it was consumed, with kernel boundaries in between.
    int32_t xd_world, xd_rank;
    double* xrecv[kMaxGpus];                     // GPU g's receive buffer [2][xd_world (source)][Ttot][npad]
    double* xland[kMaxGpus];                     // GPU g's landing buffer [2][Ttot][npad]
    const uint8_t* mask4;    // [Ttot][col_stride] NA nibble of every quad (bit k: individual 4q+k observed)
    double* delta;           // [Ttot][npad] or nullptr: increments applied since the last exchange (multi-GPU)
    const double* merge_tot; // [Ttot][npad] or nullptr: all-reduced deltas of the last exchange, still to be merged: every CTA first
                             // adds (merge_tot - delta) to its rows of eps and clears delta (the merge kernel, fused)
    // Hybrid plan (ndir = 1, one trait): the LAST row of every pass of two or more rows is a DIRECT row -- no table; its
    // genotypes are read from a second copy in plain 2-bit dosage fields, bed2 [column][drows = npass * nsm][64 bytes], row
    // slot pass * nsm + cta, and decoded on the fp64 / integer pipes while the look-ups keep the shared-memory pipe busy.
    const uint8_t* bed2;
    int32_t drows, ndir;
    uint32_t zero;           // 0 (see stream_rows)
    int32_t* err;
    int32_t pf;              // 1: L2 prefetch ahead of the streaming loads
    int32_t pdl;             // host side: launch with the programmatic-serialization attribute (the prologue overlaps the previous kernel)
    unsigned long long* prof;   // debug (GMRM_STEP_PROF): 8 cycle counters, see stream_rows / producers
};

struct SampleParams {
    int32_t V, T, G, K, N, nsm;
    int32_t it;
    uint32_t seed;
    int32_t r0;              // global index of this shard's first virtual rank
    int32_t R;               // total virtual ranks
    int32_t step;
    int32_t marker_begin;    // global index of the shard's first marker
    int32_t Mloc;
    const int32_t* cols;     // [V]
    const double* partial;   // [V][T][nsm]
    const double* spart;     // [T][nsm]
    const uint32_t* miss_off;// [Mloc+1]
    const uint32_t* miss_idx;
    const double* eps;
    int64_t npad;
    const double* mave;      // [T][Mloc]
    const double* msig;
    double* betas;           // [T][Mloc]
    int32_t* comp;
    const int32_t* group;    // [Mloc] group of each shard-local marker
    const double* sigmag;    // [T][G]
    const double* gc;        // [T][G][4K] per-iteration sampler constants (group_consts_kernel)
    const int32_t* nonas;    // [T]
    int32_t* cass;           // [T][G*K]
    PubEntry* pub;           // unused (kept for the layout of older callers)
    double* plist;           // [T][publist_doubles(V)] this GPU's list: CTA c writes segment c of every trait
    // peer-memory exchange (world > 1): every CTA also stores its segments into every peer's buffer over NVLink
    int32_t world, rank;
    double* peer_list[kMaxGpus];                 // where GPU g wants THIS GPU's list (nullptr: no exchange)
    // L2 prefetch of a published marker's column (it was streamed a step ago and has mostly left L2): the next step kernel's
    // update phase then finds its bytes in L2 instead of waiting for HBM.  nullptr: off
    const uint8_t* pf_bed;
    int64_t pf_col_stride;
    unsigned long long seq;                      // sequence number of this step, written into the segment headers
    const double* rep_u;     // replay: [Mm][R][T] or nullptr
    const double* rep_z;
    int32_t* err;            // device error flag
    int64_t* npublished;
    int32_t pdl;             // host side: launch with the programmatic-serialization attribute
};

// launchers (kernels.cu)
void launch_transcode(const uint8_t* plink, int nmark, const Layout& L, uint8_t* dst, uint32_t* miss_counts, cudaStream_t s);
void launch_fill_missing(const uint8_t* plink, int nmark, const Layout& L, const uint32_t* off, uint32_t* idx, cudaStream_t s);
// hybrid plan: the second (2-bit) copy of the direct rows of `nmark` staged PLINK columns, dst [nmark][npass * nsm][64]
void launch_direct_plane(const uint8_t* plink, int nmark, const Layout& L, int npass, uint8_t* dst, cudaStream_t s);
void launch_untranscode(const uint8_t* bed, int nmark, const Layout& L, const uint32_t* miss_off, const uint32_t* miss_idx,
                        uint8_t* plink_out, cudaStream_t s);
void launch_decode_column(const uint8_t* col, const Layout& L, const uint32_t* miss_idx, uint32_t nmiss, double* a, double* b, cudaStream_t s);
void launch_decode_namask(const uint8_t* mask4, const Layout& L, double* na, cudaStream_t s);
void launch_generate_plink(uint8_t* dst, int nmark, int first_global_marker, const Layout& L, uint32_t seed,
                           double maf_lo, double maf_hi, double missing_rate, cudaStream_t s);
// na_off [T+1] / na_idx: per trait the individuals (< N) without a phenotype, ascending
void launch_stats(const uint8_t* bed, int nmark, const Layout& L, const uint8_t* mask4, const uint32_t* miss_off, const uint32_t* miss_idx,
                  const int32_t* nonas, const uint32_t* na_off, const uint32_t* na_idx, int T, double* mave, double* msig, cudaStream_t s,
                  double* xtx = nullptr);
void launch_eps_offset(double* eps, const uint8_t* mask4, const Layout& L, int T, const double* mu_old, const double* mu_new, cudaStream_t s);
void launch_eps_merge(double* eps, double* loc, const double* tot, const Layout& L, int T, cudaStream_t s);
void launch_eps_sumsq(const double* eps, int64_t npad, int64_t n, int T, double* out, cudaStream_t s);
void launch_fill_u64(unsigned long long* dst, size_t n, unsigned long long value, cudaStream_t s);
// dynamic shared memory the step kernel needs, or -1 if (V, T, rows_per_pass, rows per CTA) do not fit
int step_npass(const Layout& L, int rows_per_pass);
void host_pass_rows(int nrows, int npass, int nsm, int cta, int* start, int* count);   // the kernel's row ownership [npass], for tests
int host_chunk_offset(int nr, int slot, int lane16, int k);                                // the kernel's lane -> byte map inside a pass chunk, for tests
int step_smem_bytes(const Layout& L, int V, int T, int rows_per_pass);
// traits per launch and rows per pass for a step of V markers (0 rows = does not fit)
void step_plan(const Layout& L, int V, int Ttot, int* traits_per_launch, int* rows_per_pass);
int launch_step(const Layout& L, int T, const StepParams& p, cudaStream_t s);
void launch_sample(const SampleParams& p, cudaStream_t s);
void launch_finish_dots(const SampleParams& p, double* out, cudaStream_t s);
void launch_steptab(int32_t* tab, int Mm, int Vl, int r0, int R, int Mt, int marker_begin, int shuffle,
                    uint32_t seed, int it, const int32_t* rep_perm, cudaStream_t s);
int beta_sq_scratch_doubles(int T, int G);
void launch_beta_sq(const double* betas, const int32_t* group, int Mloc, int T, int G, double* out, double* scratch, cudaStream_t s);

struct GlobalDrawParams {
    int32_t T, G, K, N, it;
    uint32_t seed;
    const int32_t* mtotgrp;   // [G]
    const double* bsq;        // [T][G]   (summed over shards)
    const int32_t* cass;      // [T][G*K] (summed over shards)
    const double* esq;        // [T]
    double* sigmag; double* sigmae; double* pi; int32_t* m0;
    const double* rep_sigg_unit; const double* rep_pi_unit; const double* rep_sige_unit;   // or nullptr
    int32_t* err;
};
void launch_global_draw(const GlobalDrawParams& p, cudaStream_t s);
void launch_group_consts(int T, int G, int K, int N, const double* sigmag, const double* sigmae, const double* pi, const double* cva,
                         const double* cvai, const int32_t* nonas, double* gc, cudaStream_t s);

struct MuDrawParams {
    int32_t T, it; uint32_t seed;
    const double* sigmae; const int32_t* nonas;
    double* mu; double* mu_old;
    const double* rep_mu;   // [T] or nullptr
};
void launch_mu_draw(const MuDrawParams& p, cudaStream_t s);
void launch_init_sigmae(const double* esq, const int32_t* nonas, int T, double* sigmae, cudaStream_t s);

// association pass, Bayes::predict (predict.cu)
int gvalue_chunks(int nmark);   // partial buffers launch_gvalues needs for a block of nmark markers: [chunks][npad] doubles
// g[i] = na_i * sum_{m in [m_begin, m_end)} ((a_im - mave_m) msig_m) beta_m  (columns local to the shard; keep[m] == 0: skipped);
// add (or nullptr) accumulates g as well
void launch_gvalues(const uint8_t* bed, const Layout& L, const uint32_t* miss_off, const uint32_t* miss_idx, int m_begin, int m_end,
                    const double* mave, const double* msig, const double* beta, const uint8_t* keep, const uint8_t* mask4,
                    double* part, double* g, double* add, cudaStream_t s);
void launch_predict_residual(const double* y, const double* g, const double* g_k, const Layout& L, double* y_k, cudaStream_t s);
void launch_predict_finish(const int32_t* cols, int V, int nsm, const double* partial, const double* xtx, const double* sumsq,
                           int32_t nonas, const uint8_t* keep, double* beta, double* tdist, double* se, double* pval, cudaStream_t s);
void launch_iota(int32_t* x, int n, cudaStream_t s);

}  // namespace gmrm
