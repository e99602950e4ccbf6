// sm_100a kernels of the association pass, Bayes::predict (src/bayes.cpp:14-284): genetic values g = X beta of a
// block of markers on the base-3 quad layout (layout.h), and the per-marker test statistics.  The marker sums
// sum_i a_im y_i themselves are taken by the step kernel of the Gibbs path (kernels.cu) with y as the residual.
//
// Bound: HBM (each column is read once per pass: N/4 bytes per marker); the accumulation is 16 shared-memory
// look-ups + 16 fp64 adds per 32-bit word of a column.
#include "kernels.cuh"

namespace gmrm {

namespace {

// [kernels-begin]  (tests/emu/ compiles the text between these markers for the host, CTA by CTA with one thread per
// CUDA thread, to check the kernels' indexing without a GPU; nothing in the product depends on that)
constexpr int kGvThreads = 128;     // one thread per 32-bit word (16 individuals) of a column
constexpr int kGvStage = 32;        // markers whose coefficients are staged in shared memory per round
constexpr int kGvLoads = 8;         // column words in flight per thread

// value a marker adds to an individual of dosage d: ((a - mave) * msig) * beta_mean, bayes.cpp:118-122 with b = na = 1
__device__ __forceinline__ double gv_coef(int d, double mave, double msig, double beta) { return (((double)d - mave) * msig) * beta; }

// Partial genetic values: part[c][i] = sum over the markers of chunk c of coef_m[dosage_im], missing genotypes
// counted as dosage 0 (gvalue_missing_kernel takes them out again).  grid = (words / 128, chunks).
__global__ void __launch_bounds__(kGvThreads) gvalue_partial_kernel(const uint8_t* __restrict__ bed, int64_t col_stride, int nwords,
                                                                    int m_begin, int m_end, int chunk_len,
                                                                    const double* __restrict__ mave, const double* __restrict__ msig,
                                                                    const double* __restrict__ beta, const uint8_t* __restrict__ keep,
                                                                    double* __restrict__ part, int64_t npad) {
    __shared__ uint8_t fld[kTabEntries];            // base-3 byte -> four 2-bit dosages
    __shared__ double coef[kGvStage][4];            // [marker of the round][dosage]
    const int tid = threadIdx.x;
    for (int e = tid; e < kTabEntries; e += kGvThreads) fld[e] = (uint8_t)tri_to_fields((uint32_t)e);
    const int w = blockIdx.x * kGvThreads + tid;
    const bool live = w < nwords;
    const int m0 = m_begin + (int)blockIdx.y * chunk_len;
    const int m1 = min(m_end, m0 + chunk_len);
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0;
    for (int mb = m0; mb < m1; mb += kGvStage) {
        __syncthreads();                            // previous round's coefficients are no longer read (and fld is written)
        const int nm = min(kGvStage, m1 - mb);
        if (tid < nm) {
            const int m = mb + tid;
            const bool k = keep == nullptr || keep[m] != 0;
            const double av = mave[m], sg = msig[m], b = beta[m];
#pragma unroll
            for (int d = 0; d < 3; d++) coef[tid][d] = k ? gv_coef(d, av, sg, b) : 0.0;
            coef[tid][3] = 0.0;
        }
        __syncthreads();
        for (int i0 = 0; i0 < nm; i0 += kGvLoads) {
            uint32_t wd[kGvLoads];
#pragma unroll
            for (int j = 0; j < kGvLoads; j++)
                wd[j] = (live && i0 + j < nm) ? __ldg(reinterpret_cast<const uint32_t*>(bed + (int64_t)(mb + i0 + j) * col_stride) + w) : 0u;
#pragma unroll
            for (int j = 0; j < kGvLoads; j++) {
                if (i0 + j < nm) {
                    const uint32_t x = wd[j];
                    const uint32_t f = (uint32_t)fld[x & 0xffu] | ((uint32_t)fld[(x >> 8) & 0xffu] << 8) |
                                       ((uint32_t)fld[(x >> 16) & 0xffu] << 16) | ((uint32_t)fld[x >> 24] << 24);
                    const double* c = coef[i0 + j];
#pragma unroll
                    for (int k = 0; k < 16; k++) acc[k] += c[(f >> (2 * k)) & 3u];
                }
            }
        }
    }
    if (live) {
        double* dst = part + (int64_t)blockIdx.y * npad + (int64_t)w * 16;
#pragma unroll
        for (int k = 0; k < 16; k++) dst[k] = acc[k];
    }
}

// Missing genotypes were counted as dosage 0 above: take coef_m[0] out again.  One CTA per chunk walks its markers
// in order; the individuals of one marker's list are distinct, so plain read-modify-writes are race-free and the
// result does not depend on scheduling.
__global__ void __launch_bounds__(256) gvalue_missing_kernel(const uint32_t* __restrict__ miss_off, const uint32_t* __restrict__ miss_idx,
                                                             int m_begin, int m_end, int chunk_len,
                                                             const double* __restrict__ mave, const double* __restrict__ msig,
                                                             const double* __restrict__ beta, const uint8_t* __restrict__ keep,
                                                             double* __restrict__ part, int64_t npad) {
    const int m0 = m_begin + (int)blockIdx.x * chunk_len;
    const int m1 = min(m_end, m0 + chunk_len);
    double* dst = part + (int64_t)blockIdx.x * npad;
    for (int m = m0; m < m1; m++) {
        const uint32_t k0 = miss_off[m], k1 = miss_off[m + 1];
        if (k1 == k0 || (keep != nullptr && keep[m] == 0)) continue;      // uniform over the CTA
        const double c0 = gv_coef(0, mave[m], msig[m], beta[m]);
        for (uint32_t i = k0 + threadIdx.x; i < k1; i += blockDim.x) dst[miss_idx[i]] -= c0;
        __syncthreads();
    }
}

// g[i] = na_i * sum_c part[c][i] in chunk order (the NA factor of bayes.cpp:118 is per individual); 0 past N.
// add != nullptr: g[i] is added to add[i] as well (running total over the blocks).
__global__ void __launch_bounds__(256) gvalue_reduce_kernel(const double* __restrict__ part, int nchunk, int64_t npad, int32_t N,
                                                            const uint8_t* __restrict__ mask4, double* __restrict__ g,
                                                            double* __restrict__ add) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    double s = 0.0;
    for (int c = 0; c < nchunk; c++) s += part[(int64_t)c * npad + i];
    const bool obs = i < N && ((mask4[i >> 2] >> (i & 3)) & 1u);
    const double v = obs ? s : 0.0;
    g[i] = v;
    if (add) add[i] += v;
}

// y_k = y - (g - g_k), bayes.cpp:141-147: only the OTHER blocks' genetic values leave the phenotype
__global__ void __launch_bounds__(256) predict_residual_kernel(const double* __restrict__ y, const double* __restrict__ g,
                                                               const double* __restrict__ g_k, int64_t npad, int32_t N,
                                                               double* __restrict__ y_k) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    y_k[i] = i < N ? y[i] - (g[i] - g_k[i]) : 0.0;
}

// Test statistics of bayes.cpp:199-208 for the V markers `cols` of a block: xty = sum of the step kernel's per-CTA
// partial sums (fixed order), xtx = the marker's count-based sum of squares, sigma = sumsq / nonas.  Warp per marker.
__global__ void __launch_bounds__(128) predict_finish_kernel(const int32_t* __restrict__ cols, int V, int nsm,
                                                             const double* __restrict__ partial, const double* __restrict__ xtx,
                                                             const double* __restrict__ sumsq, int32_t nonas,
                                                             const uint8_t* __restrict__ keep, double* __restrict__ beta,
                                                             double* __restrict__ tdist, double* __restrict__ se,
                                                             double* __restrict__ pval) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= V) return;
    const int col = cols[v];
    const double* part = partial + (int64_t)v * nsm;
    double s = 0.0;
    for (int i = lane; i < nsm; i += 32) s += part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);   // butterfly: same value and order on every lane
    if (lane != 0) return;
    if (keep != nullptr && keep[col] == 0) {                                    // id absent from the reference .bim: skipped
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        beta[col] = nan; tdist[col] = nan; se[col] = nan; pval[col] = nan;
        return;
    }
    const double sigma = sumsq[0] / (double)nonas;                              // bayes.cpp:149-152
    const double q = xtx[col], xty = s;
    const double b = xty / q;
    const double t = xty / sqrt(sigma * q);
    beta[col] = b;
    tdist[col] = t;
    se[col] = b / t;
    pval[col] = 1.0 - erf(sqrt(t * t * 0.5));                                   // 1 - gamma_p(1/2, t^2/2)
}

__global__ void iota_kernel(int32_t* __restrict__ x, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = i;
}

// enough (word-block, chunk) CTAs to fill the GPU for short blocks too, few enough partial buffers to stay small
inline int gvalue_chunks_of(int nmark) {
    if (nmark <= 0) return 1;
    const int c = (nmark + kGvStage - 1) / kGvStage;
    return c < 64 ? c : 64;
}
// launch geometry of the three genetic-value kernels for a block of nmark markers
struct GvPlan { int nchunk, chunk_len, nwords, word_blocks; };
inline GvPlan gvalue_plan(const Layout& L, int nmark) {
    GvPlan p;
    p.nchunk = gvalue_chunks_of(nmark);
    p.chunk_len = nmark > 0 ? (nmark + p.nchunk - 1) / p.nchunk : 1;
    p.nwords = (int)(L.col_stride / 4);
    p.word_blocks = (p.nwords + kGvThreads - 1) / kGvThreads;
    return p;
}
// [kernels-end]

}  // namespace

int gvalue_chunks(int nmark) { return gvalue_chunks_of(nmark); }

void launch_gvalues(const uint8_t* bed, const Layout& L, const uint32_t* miss_off, const uint32_t* miss_idx, int m_begin, int m_end,
                    const double* mave, const double* msig, const double* beta, const uint8_t* keep, const uint8_t* mask4,
                    double* part, double* g, double* add, cudaStream_t s) {
    const int nmark = m_end - m_begin;
    const GvPlan pl = gvalue_plan(L, nmark);
    if (nmark > 0) {
        dim3 grid((unsigned)pl.word_blocks, (unsigned)pl.nchunk);
        gvalue_partial_kernel<<<grid, kGvThreads, 0, s>>>(bed, L.col_stride, pl.nwords, m_begin, m_end, pl.chunk_len, mave, msig, beta, keep, part, L.npad);
        gvalue_missing_kernel<<<pl.nchunk, 256, 0, s>>>(miss_off, miss_idx, m_begin, m_end, pl.chunk_len, mave, msig, beta, keep, part, L.npad);
    }
    gvalue_reduce_kernel<<<(unsigned)((L.npad + 255) / 256), 256, 0, s>>>(part, nmark > 0 ? pl.nchunk : 0, L.npad, L.N, mask4, g, add);
}

void launch_predict_residual(const double* y, const double* g, const double* g_k, const Layout& L, double* y_k, cudaStream_t s) {
    predict_residual_kernel<<<(unsigned)((L.npad + 255) / 256), 256, 0, s>>>(y, g, g_k, L.npad, L.N, y_k);
}

void launch_predict_finish(const int32_t* cols, int V, int nsm, const double* partial, const double* xtx, const double* sumsq,
                           int32_t nonas, const uint8_t* keep, double* beta, double* tdist, double* se, double* pval, cudaStream_t s) {
    if (V <= 0) return;
    predict_finish_kernel<<<(unsigned)((V + 3) / 4), 128, 0, s>>>(cols, V, nsm, partial, xtx, sumsq, nonas, keep, beta, tdist, se, pval);
}

void launch_iota(int32_t* x, int n, cudaStream_t s) {
    if (n > 0) iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n);
}

}  // namespace gmrm
