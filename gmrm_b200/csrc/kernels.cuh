// sm_100a kernels of the Gibbs marker loop.  See layout.h for the HBM layout and the decode
// algebra, DESIGN.md for the roofline of each kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "gmrm_rng.h"
#include "layout.h"
#include "sampler.h"

namespace gmrm {

// ------------------------------------------------------------------------------------------
// dot kernel geometry
// shared-memory ring of the dot kernel: 64 tiles (<= 64 KB; 48 when 3 consumer warps share a sub-partition so
// that a stage always serves the same warp), cut into stages of BATCH markers
__host__ __device__ constexpr int dot_ring_tiles(int wps) { return wps == 3 ? 48 : 64; }
constexpr int kDotMaxThreads = 17 * 32; // 4 consumer warps per sub-partition + 1 producer warp, the widest variant
constexpr int kDotMaxBatch = 8;
constexpr int kUpdSplit = 4;                              // warps per sub-partition in the update kernel
constexpr int kUpdCap = 2048;                             // virtual ranks staged per round of the update kernel
constexpr int kUpdPF = 4;                                 // columns each update warp prefetches ahead
constexpr int kUpdThreads = kLanesPerTile * kUpdSplit;

struct PubEntry {   // one published update: Phenotype::update_epsilon's dbeta[3], phenotype.cpp:326-329
    double lam;     // dbeta * msig   (0 == nothing to apply)
    double mave;
};

struct DotParams {
    const uint8_t* bed;
    int64_t col_stride;
    const int32_t* cols;     // [V] shard-local column of each virtual rank this step, -1 = none
    int32_t V;
    const double* eps;       // [T][npad]
    int64_t npad;
    double* partial;         // [V][Ttot][nsl]
    int32_t nsl;             // nsm * 4
    int32_t Ttot, t0;        // this launch handles traits t0 .. t0+T-1 of Ttot
    const double* zeros;     // >= kDotMaxThreads * kDotMaxBatch zeros (opaque to ptxas, see set_lo)
    int32_t debug;           // 0 normal; 1 feed only (no arithmetic); 2 compute only (no TMA, no barriers) -- profiling aids
    int32_t variant;         // (consumer warps / sub-partition, markers / batch): 0 = (2,8), 1 = (4,4), 2 = (2,4), 3 = (3,4)
};

struct SampleParams {
    int32_t V, T, G, K, N, nsl, nsm;
    int32_t it;
    uint32_t seed;
    int32_t r0;              // global index of this shard's first virtual rank
    int32_t R;               // total virtual ranks
    int32_t step;
    int32_t marker_begin;    // global index of the shard's first marker
    int32_t Mloc;
    const int32_t* cols;     // [V]
    const double* partial;   // [V][T][nsl]
    const double* spart;     // [T][nsm]  per-tile sum of eps
    const uint32_t* miss_off;// [Mloc+1]
    const uint32_t* miss_idx;
    const double* eps;
    int64_t npad;
    const double* mave;      // [T][Mloc]
    const double* msig;
    double* betas;           // [T][Mloc]
    int32_t* comp;
    const int32_t* group;    // [Mloc] group of each shard-local marker
    const double* cva;       // [G*K]
    const double* cvai;
    const double* sigmag;    // [T][G]
    const double* sigmae;    // [T]
    const double* pi;        // [T][G*K]
    const double* gc;        // [T][G][4K] per-iteration sampler constants (group_consts_kernel)
    const int32_t* nonas;    // [T]
    int32_t* cass;           // [T][G*K]
    PubEntry* pub;           // [V][T]
    const double* rep_u;     // replay: [Mm][R][T] or nullptr
    const double* rep_z;
    int32_t* err;            // device error flag
    int64_t* npublished;
};

struct UpdateParams {
    const uint8_t* bed;
    int64_t col_stride;
    const int32_t* cols;     // [V]
    int32_t V, T;
    const PubEntry* pub;     // [V][T]
    const uint32_t* miss_off;// to pick the exact path for markers with missing genotypes
    const uint8_t* namask2;  // [T][col_stride] tile layout, field 01 = observed
    const uint8_t* na01;     // [T][npad] 1 = observed, per individual
    double* eps;             // [T][npad]
    int64_t npad;
    double* spart;           // [T][nsm]
    double* delta;           // [T][npad] or nullptr: accumulates the applied increments (multi-GPU exchange)
    int32_t exact;           // 1: always use the reference-order arithmetic
};

// launchers (kernels.cu)
void launch_transcode(const uint8_t* src, int nmark, const Layout& L, uint8_t* dst, cudaStream_t s);
void launch_decode_column(const uint8_t* col, const Layout& L, double* a, double* b, cudaStream_t s);
void launch_untranscode(const uint8_t* tiles, int nmark, const Layout& L, uint8_t* dst, cudaStream_t s);
void launch_generate_plink(uint8_t* dst, int nmark, int first_global_marker, const Layout& L, uint32_t seed,
                           double maf_lo, double maf_hi, double missing_rate, cudaStream_t s);
void launch_count_missing(const uint8_t* bed, int nmark, const Layout& L, uint32_t* counts, cudaStream_t s);
void launch_fill_missing(const uint8_t* bed, int nmark, const Layout& L, const uint32_t* off, uint32_t* idx, cudaStream_t s);
void launch_stats(const uint8_t* bed, int nmark, const Layout& L, const uint8_t* namask2, const int32_t* nonas, int T,
                  double* mave, double* msig, cudaStream_t s);
void launch_eps_offset(double* eps, const uint8_t* na01, const Layout& L, int T, const double* mu_old,
                       const double* mu_new, double* spart, cudaStream_t s);
void launch_eps_merge(double* eps, double* loc, const double* tot, const Layout& L, int T, double* spart, cudaStream_t s);
void launch_eps_sumsq(const double* eps, int64_t npad, int64_t n, int T, double* out, cudaStream_t s);
int launch_dot(const Layout& L, int T, const DotParams& p, cudaStream_t s);
int launch_dot_table(const Layout& L, const DotParams& p, cudaStream_t s);
int dot_table_passes(const Layout& L);
void launch_sample(const SampleParams& p, cudaStream_t s);
int launch_update(const Layout& L, const UpdateParams& p, cudaStream_t s);
void launch_finish_dots(const SampleParams& p, double* out, cudaStream_t s);
void launch_steptab(int32_t* tab, int Mm, int Vl, int r0, int R, int Mt, int marker_begin, int shuffle,
                    uint32_t seed, int it, const int32_t* rep_perm, cudaStream_t s);
void launch_beta_sq(const double* betas, const int32_t* group, int Mloc, int T, int G, double* out, cudaStream_t s);

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

}  // namespace gmrm
