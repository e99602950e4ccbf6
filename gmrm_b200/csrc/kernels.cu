// sm_100a kernels of the Gibbs marker loop (see kernels.cuh, layout.h).
#include <algorithm>
#include <atomic>
#include <type_traits>

#include "kernels.cuh"

#include <cstdio>

namespace gmrm {

// =====================================================================================
// small device helpers
// =====================================================================================
// [helpers-begin]  (tests/test_step_kernel_emulated.py compiles [helpers-*] and [step-*] for the host, see tests/emu/)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ double warp_sum_fixed(double v) {   // fixed xor tree: reproducible
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, fixed order; result valid in thread 0.  `red` holds >= blockDim/32 doubles.
__device__ __forceinline__ double block_sum_fixed(double v, double* red) {
    v = warp_sum_fixed(v);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < nw; i++) t += red[i];
    return t;
}
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start while
// its predecessor in the stream is still draining; pdl_wait() returns once the predecessor has completed and its writes
// are visible, pdl_trigger() lets the successor's CTAs be scheduled from here on.  Both are no-ops in a plain launch.
// 32-byte accesses of the multi-GPU exchange: full-sector stores (peer memory over NVLink), polling loads that bypass L1
__device__ __forceinline__ void st_v4_f64(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_poll_v4(const double* p, unsigned long long (&w)[4]) {
    asm volatile("ld.volatile.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// [helpers-end]

// =====================================================================================
// .bed ingestion: PLINK bytes <-> base-3 quad bytes + missing lists (bit-exact, invertible)
// =====================================================================================
// [ingest-begin]  (tests/test_chain_emulated.py compiles the [ingest|eps|epilogue] regions for the host too, see tests/emu/)
__global__ void transcode_kernel(const uint8_t* __restrict__ plink, int nmark, Layout L, uint8_t* __restrict__ dst,
                                 uint32_t* __restrict__ miss_counts) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (o >= L.col_stride || j >= nmark) return;
    uint8_t v = 0;
    if (o < L.mbytes) {
        uint32_t mm;
        v = plink_to_tri(plink[(int64_t)j * L.mbytes + o], &mm);
        if (mm) atomicAdd(&miss_counts[j], (uint32_t)__popc(mm));
    }
    dst[(int64_t)j * L.col_stride + o] = v;
}

// Missing-genotype lists (CSR over markers), ascending individual index, from the staged PLINK bytes.
// One warp per marker walks the column in byte order.  Pad individuals of the last byte are listed too
// (the round trip keeps arbitrary pad bits); their residual slots are 0 and their NA bits clear.
__global__ void fill_missing_kernel(const uint8_t* __restrict__ plink, int nmark, Layout L, const uint32_t* __restrict__ off,
                                    uint32_t* __restrict__ out) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= nmark) return;
    if (off[j + 1] == off[j]) return;
    const uint8_t* col = plink + (int64_t)j * L.mbytes;
    uint32_t base = off[j];
    for (int64_t idx0 = 0; idx0 < L.mbytes; idx0 += 32) {
        const int64_t idx = idx0 + lane;
        const uint32_t x = idx < L.mbytes ? col[idx] : 0xffu;
        const uint32_t m = x & (~x >> 1) & 0x55u;            // code 01: low bit set, high bit clear
        const uint32_t n = __popc(m);
        uint32_t incl = n;                                   // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t w = base + incl - n;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (m & (1u << (2 * k))) out[w++] = (uint32_t)(idx * 4 + k);
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__global__ void untranscode_kernel(const uint8_t* __restrict__ bed, int nmark, Layout L, uint8_t* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (idx >= L.mbytes || j >= nmark) return;
    out[(int64_t)j * L.mbytes + idx] = fields_to_plink(tri_to_fields(bed[(int64_t)j * L.col_stride + idx]), 0u);
}
// second half of the inverse: listed individuals get code 01 back (dosage 0 was written as 11: clear its high bit)
__global__ void unmiss_kernel(int nmark, Layout L, const uint32_t* __restrict__ off, const uint32_t* __restrict__ midx,
                              uint8_t* __restrict__ out) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= nmark) return;
    for (uint32_t i = off[j] + lane; i < off[j + 1]; i += 32) {
        const uint32_t ind = midx[i];
        const uintptr_t a = (uintptr_t)(out + (int64_t)j * L.mbytes + (ind >> 2));
        const uint32_t bit = (uint32_t)((a & 3) * 8) + 2 * (ind & 3) + 1;
        atomicAnd(reinterpret_cast<uint32_t*>(a & ~(uintptr_t)3), ~(1u << bit));
    }
}

// test hooks: the reference's table values of one column / one NA mask, individual by individual
__global__ void decode_column_kernel(const uint8_t* __restrict__ col, Layout L, double* __restrict__ a, double* __restrict__ b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.N) return;
    const uint32_t d = (tri_to_fields(col[i >> 2]) >> (2 * (i & 3))) & 3u;
    if (a) a[i] = (double)d;               // dotp_lut_a of a non-missing code
    if (b) b[i] = 1.0;                     // dotp_lut_b
}
__global__ void decode_missing_kernel(const uint32_t* __restrict__ midx, uint32_t nmiss, int N, double* __restrict__ a, double* __restrict__ b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmiss) return;
    const uint32_t ind = midx[i];
    if (ind >= (uint32_t)N) return;
    if (a) a[ind] = 0.0;                   // missing: a = b = 0 (lut/mk_lut.cpp:25-32)
    if (b) b[ind] = 0.0;
}
__global__ void decode_namask_kernel(const uint8_t* __restrict__ mask4, Layout L, double* __restrict__ na) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.N) return;
    na[i] = ((mask4[i >> 2] >> (i & 3)) & 1u) ? 1.0 : 0.0;   // na_lut (lut/mk_lut_na.cpp:25-29)
}
// [ingest-end]

// Synthetic PLINK bytes: per-marker MAF ~ U(lo, hi), dosage ~ Binomial(2, p).
__global__ void generate_plink_kernel(uint8_t* __restrict__ dst, int nmark, int first_marker, Layout L, uint32_t seed,
                                      double maf_lo, double maf_hi, double missing_rate) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (idx >= L.mbytes || j >= nmark) return;
    const uint32_t mg = (uint32_t)(first_marker + j);
    const U4 pm = philox4x32(seed, 0x47454e30u, mg, 0, 0, 0);
    const double p = maf_lo + (maf_hi - maf_lo) * u01(pm.x, pm.y);
    const uint32_t thr = (uint32_t)(p * 65536.0);
    const uint32_t mthr = (uint32_t)(missing_rate * 4294967296.0 > 4294967295.0 ? 4294967295.0 : missing_rate * 4294967296.0);
    const U4 r = philox4x32(seed, 0x47454e31u, mg, (uint32_t)idx, 0, 0);
    U4 rm = {~0u, ~0u, ~0u, ~0u};
    if (missing_rate > 0.0) rm = philox4x32(seed, 0x47454e32u, mg, (uint32_t)idx, 0, 0);
    const uint32_t rw[4] = {r.x, r.y, r.z, r.w}, mw[4] = {rm.x, rm.y, rm.z, rm.w};
    uint8_t byte = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t i = idx * 4 + k;
        uint32_t code = 0;   // PLINK pads with 00
        if (i < L.N) {
            const int d = ((rw[k] & 0xffffu) < thr) + ((rw[k] >> 16) < thr);
            code = d == 2 ? 0u : d == 1 ? 2u : 3u;          // 00 = dosage 2, 10 = 1, 11 = 0
            if (missing_rate > 0.0 && mw[k] < mthr) code = 1u;   // 01 = missing
        }
        byte |= (uint8_t)(code << (2 * k));
    }
    dst[(int64_t)j * L.mbytes + idx] = byte;
}

// =====================================================================================
// marker statistics, PhenMgr::compute_markers_statistics (phenotype.cpp:466-556), from integer
// counts of each dosage under the trait's NA mask: the sums of the
// reference's loops are sums of small integers, hence these counts exactly.
// =====================================================================================
// [stats-begin]  (tests/test_stats_kernel_emulated.py compiles the text up to [stats-end] for the host, see tests/emu/)
// One pass over a column, whatever the number of traits: the dosage-1 / dosage-2 counts over ALL individuals come from a
// per-byte count table (conflict-free: one copy per lane); per trait the few individuals WITHOUT a phenotype are then
// taken out again from a list (na_idx), and the observed individuals with a missing genotype from the marker's list.
// Persistent CTAs (the table is built once); HBM-bound: ceil(N/4) bytes per marker.
constexpr int kStatsThreads = 256;
#if !defined(__CUDA_ARCH__) && !defined(__CUDACC__)
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) { return emu_prmt(a, b, sel); }   // host emulation
#endif
__device__ __forceinline__ int stats_block_sum(int v, int* red) {     // every thread gets the total
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int i = 0; i < kStatsThreads / 32; i++) t += red[i];
    return t;
}
__global__ void __launch_bounds__(kStatsThreads) stats_kernel(const uint8_t* __restrict__ bed, int nmark, Layout L,
                                                              const uint8_t* __restrict__ mask4, const uint32_t* __restrict__ miss_off,
                                                              const uint32_t* __restrict__ miss_idx, const int32_t* __restrict__ nonas,
                                                              const uint32_t* __restrict__ na_off, const uint32_t* __restrict__ na_idx,
                                                              int T, double* __restrict__ mave, double* __restrict__ msig,
                                                              double* __restrict__ xtx) {
    __shared__ uint32_t cnt[kTabEntries * 64];   // [e][lane]: individuals of dosage 1 (bits 0-15) and 2 (bits 16-31) in byte e; stride 256 B
    __shared__ uint8_t fld[kTabEntries];                        // base-3 byte -> 2-bit dosage fields
    __shared__ int red[kStatsThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < kTabEntries * 32; i += kStatsThreads) {
        const int e = i >> 5;
        const uint32_t f = tri_to_fields((uint32_t)e), lo = f & 0x55u, hi = (f >> 1) & 0x55u;
        cnt[e * 64 + (i & 31)] = (uint32_t)__popc(lo & ~hi) | ((uint32_t)__popc(hi & ~lo) << 16);
    }
    for (int e = tid; e < kTabEntries; e += kStatsThreads) fld[e] = (uint8_t)tri_to_fields((uint32_t)e);
    __syncthreads();
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const char* cntb = reinterpret_cast<const char*>(cnt);
    const int nvec = (int)(L.col_stride / 16);
    for (int j = blockIdx.x; j < nmark; j += gridDim.x) {
        const uint8_t* colb = bed + (int64_t)j * L.col_stride;
        const uint4* col = reinterpret_cast<const uint4*>(colb);
        uint32_t acc = 0;                                       // packed counts of this thread's bytes (<= 4 per byte: no overflow below 16k bytes)
        for (int i0 = tid; i0 < nvec; i0 += 4 * kStatsThreads) {
            uint4 w[4];
#pragma unroll
            for (int u = 0; u < 4; u++) w[u] = i0 + u * kStatsThreads < nvec ? __ldg(col + i0 + u * kStatsThreads) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t x[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
                for (int q = 0; q < 4; q++) {                   // byte k -> offset e * 256 + lane * 4: one byte permute each
                    acc += *reinterpret_cast<const uint32_t*>(cntb + __byte_perm(x[q], lane4, 0x5504));
                    acc += *reinterpret_cast<const uint32_t*>(cntb + __byte_perm(x[q], lane4, 0x5514));
                    acc += *reinterpret_cast<const uint32_t*>(cntb + __byte_perm(x[q], lane4, 0x5524));
                    acc += *reinterpret_cast<const uint32_t*>(cntb + __byte_perm(x[q], lane4, 0x5534));
                }
            }
        }
        const int all1 = stats_block_sum((int)(acc & 0xffffu), red), all2 = stats_block_sum((int)(acc >> 16), red);
        const uint32_t m0 = miss_off[j], m1 = miss_off[j + 1];
        // pad slots of the last PLINK byte keep whatever bits the file had (00 = dosage 2 as PLINK writes them): not individuals
        int pad1 = 0, pad2 = 0;
        if (tid == 0)
            for (int ind = L.N; ind < 4 * L.mbytes; ind++) {
                const uint32_t d = ((uint32_t)fld[colb[ind >> 2]] >> (2 * (ind & 3))) & 3u;
                pad1 += d == 1u; pad2 += d == 2u;
            }
        for (int t = 0; t < T; t++) {
            const uint8_t* nmb = mask4 + (int64_t)t * L.col_stride;
            // individuals without a phenotype: their dosages leave the counts (na_lut == 0, phenotype.cpp:500-530)
            int d1 = pad1, d2 = pad2, mo = 0;
            for (uint32_t i = na_off[t] + tid; i < na_off[t + 1]; i += kStatsThreads) {
                const uint32_t ind = na_idx[i];
                const uint32_t d = ((uint32_t)fld[colb[ind >> 2]] >> (2 * (ind & 3))) & 3u;
                d1 += d == 1u; d2 += d == 2u;
            }
            // missing genotypes were stored as dosage 0: the observed ones are no dosage-0 individuals
            for (uint32_t i = m0 + tid; i < m1; i += kStatsThreads) {
                const uint32_t ind = miss_idx[i];
                mo += (nmb[ind >> 2] >> (ind & 3)) & 1;
            }
            const int n1 = all1 - stats_block_sum(d1, red), n2 = all2 - stats_block_sum(d2, red);
            const int n0 = nonas[t] - n1 - n2 - stats_block_sum(mo, red);
            if (tid == 0) {
                const double c0 = n0, c1 = n1, c2 = n2;
                const double suma = c1 + 2.0 * c2, sumb = c0 + c1 + c2;       // phenotype.cpp:536-537
                const double av = suma / sumb;                                // 540
                const double sumsqr = c0 * (0.0 - av) * (0.0 - av) + c1 * (1.0 - av) * (1.0 - av) + c2 * (2.0 - av) * (2.0 - av);  // 544-545
                mave[(int64_t)t * nmark + j] = av;
                msig[(int64_t)t * nmark + j] = 1.0 / sqrt(sumsqr / ((double)nonas[t] - 1.0));   // 548
                if (xtx) xtx[(int64_t)t * nmark + j] = c1 + 4.0 * c2;        // sum (a b na)^2 of Bayes::predict, bayes.cpp:190-195
            }
        }
    }
}
// [stats-end]

// =====================================================================================
// residual helpers
// =====================================================================================
// [eps-begin]
// Phenotype::offset_epsilon twice (bayes.cpp:351,359): eps += mu_old*na; eps -= mu_new*na
__global__ void __launch_bounds__(256) eps_offset_kernel(double* __restrict__ eps, const uint8_t* __restrict__ mask4, Layout L,
                                                         const double* __restrict__ mu_old, const double* __restrict__ mu_new) {
    const int t = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.npad) return;
    if (!((mask4[(int64_t)t * L.col_stride + (i >> 2)] >> (i & 3)) & 1u)) return;
    double v = eps[(int64_t)t * L.npad + i];
    v += mu_old[t];            // phenotype.cpp:408 (na == 1)
    v += -mu_new[t];
    eps[(int64_t)t * L.npad + i] = v;
}

// Multi-GPU exchange (replaces the per-marker Allgatherv + local recompute of bayes.cpp:500-547): after the
// all-reduce of the shards' residual deltas, add what the OTHER shards changed and clear the local delta.
__global__ void __launch_bounds__(256) eps_merge_kernel(double* __restrict__ eps, double* __restrict__ loc,
                                                        const double* __restrict__ tot, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    eps[i] = eps[i] + (tot[i] - loc[i]);
    loc[i] = 0.0;
}

// sum_{i<n} eps_i^2 per trait (Phenotype::epsilon_sumsqr, phenotype.cpp:251-261)
__global__ void __launch_bounds__(1024) eps_sumsq_kernel(const double* __restrict__ eps, int64_t npad, int64_t n, double* __restrict__ out) {
    __shared__ double red[32];
    const double* e = eps + (int64_t)blockIdx.x * npad;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += e[i] * e[i];
    const double tot = block_sum_fixed(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = tot;
}
// [eps-end]

// =====================================================================================
// K1: one marker-step of one GPU -- apply the previous step's published updates, build the look-up
// tables of this CTA's rows, stream the step's V columns through them (layout.h).
//
// One CTA per SM, 16 warps.  CTA c owns a short contiguous range of rows of every column and of the residuals in
// each pass (all_pass_rows below), at most rows_per_pass rows (as many (row, trait) table slots as shared memory
// holds).  Per CTA:
//   update: every thread owns up to 2 quads of the CTA's rows per round; for each published marker of the
//           previous step (rank order) it adds  v[dosage]  to its 4 residuals -- pair tables: one PRMT + LDS.64 + DADD
//           per individual serves TWO published markers; NA / missing individuals are routed to a zero entry.  On
//           several GPUs (sync rate 1) the sums start from zero and are exchanged as increments (exchange_increments)
//   build : half-warp <-> (slot, byte k, third d3 of the 81 entries); lane l owns one quad of the slot (chunk_offset):
//           reads its 4 residuals, writes 27 entries  sum_k d_k eps_k  (conflict-free 8-byte stores)
//   stream: warp w takes batches w, w+16, ... of 8 markers; in a batch, half-warp h works on marker 2i+h of
//           pair i = 0..3; lane l loads its words of the pass's rows with one 16-byte load (+ one 4-byte load for a fifth
//           row) straight from HBM/L2 into one of two register buffers, one batch ahead of its use (and prefetches into
//           L2 further ahead), and per byte does  PRMT -> LDS.64 -> DADD  into the pair's accumulator; a 16-lane
//           transposed butterfly leaves one total per marker, added to the marker's partial sum in shared memory (each
//           marker is always served by the same lane: plain read-modify-write).
//   (A shared-memory ring fed by a producer warp -- cp.async.bulk or cp.async -- was measured and dropped: with
//   227 KB of tables + partials only 32-40 KB are left for it, and the producer hand-off costs more than the
//   register double-buffer; DESIGN.md section 5 lists this and the other measured alternatives.)
// At the end the CTA writes partial[v][t][cta] and its sum of residuals spart[t][cta]; the sampler kernel
// adds the nsm partials of a marker in a fixed order.
// =====================================================================================
// [step-begin]
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
struct PubStage { double v[4]; };   // increment by dosage: (a - mave*b) * dbeta*msig for a = 0,1,2 (b = 1); v[3] = 0
struct PubInfo { int32_t col; uint32_t nmiss_g; };   // nmiss_g = missing genotypes of the column << 4 | publishing GPU

template <int IMM>
__device__ __forceinline__ double lds_f64_imm(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ uint32_t lds_u32_imm(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}
template <int K>
__device__ __forceinline__ uint32_t tab_addr(uint32_t word, uint32_t low) {   // byte K of word -> e*256 + low
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(word), "r"(low), "n"(0x5504 | (K << 4)));
    return d;
}
template <int K>
__device__ __forceinline__ uint32_t byte_into(uint32_t word, uint32_t base) { // (base & ~0xff) | byte K of word
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(word), "r"(base), "n"(0x7650 | K));
    return d;
}
__host__ __device__ constexpr int tab_imm(int slot, int k) {
    return (int)kTabBase + slot * kSlotBytes + (k >> 1) * kRegionBytes + (k & 1) * 128;
}
// acc[t] += table(slot SLOT + t, byte K)[e] for the T traits of a row: one LDS.64 + one DADD each
template <int SLOT, int K, int T, int TT = 0>
__device__ __forceinline__ void lookup_traits(double (&acc)[T], uint32_t a) {
    if constexpr (TT < T) {
        acc[TT] += lds_f64_imm<tab_imm(SLOT + TT, K)>(a);
        lookup_traits<SLOT, K, T, TT + 1>(acc, a);
    }
}

// Row ownership.  The rows of a column are cut into npass contiguous ranges (one per pass); in pass q the range is
// split evenly over the CTAs, so that during a pass the grid reads ONE contiguous run of every column (neighbouring
// CTAs read neighbouring 256..320-byte chunks: DRAM pages are used whole).  A CTA therefore owns npass short row
// ranges; `pr` lists them, local rows are numbered range after range.
constexpr int kMaxPasses = 64;
struct PassRows { int start[kMaxPasses], count[kMaxPasses], base[kMaxPasses]; int npass, total; };
// Pass q covers rows [nrows q/npass, nrows (q+1)/npass): n = every * nsm + extra rows, the first `extra` positions of the
// split get one more.  Position 0 of a pass is the CTA after the last one served an extra row by the passes before it
// (round robin), so a CTA's total over the step stays within one row of the average.
__host__ __device__ __forceinline__ void all_pass_rows(int nrows, int npass, int nsm, int cta, int* start, int* count) {
    int lo = 0, first = 0;
    for (int q = 0; q < npass; q++) {
        const int hi = (int)((int64_t)nrows * (q + 1) / npass), n = hi - lo, every = n / nsm, extra = n % nsm;
        const int c = (cta - first + nsm) % nsm;
        start[q] = lo + c * every + (c < extra ? c : extra);
        count[q] = every + (c < extra ? 1 : 0);
        first = (first + extra) % nsm;
        lo = hi;
    }
}
__device__ __forceinline__ int global_row(const PassRows& pr, int local_row) {
    int q = 0;
    while (q + 1 < pr.npass && local_row >= pr.base[q + 1]) q++;
    return pr.start[q] + (local_row - pr.base[q]);
}

// bytes of the table area; an update-only launch (V == 0) sizes it for staging kPubCap published columns
__host__ __device__ inline int step_area_bytes(int V, int T, int rows_per_pass, int npass) {
    if (V > 0) return rows_per_pass * T * kSlotBytes;
    // update-only launch: room for the staged column bytes, the pair tables, the prefix of the segment counts and, in the
    // row-sharded form, the entry groups' increments -- every launch asks for the full 227 KB anyway, so take a fixed 176 KB
    (void)npass; (void)rows_per_pass;
    return 180224;
}

// Exclusive prefix of the segment counts of all pending lists of trait tt (GPU-major = global virtual-rank order) into
// segpre[0 .. nseg], segpre[nseg] = number of published items.  Whole CTA; ends with a barrier.
template <int NT>
__device__ __forceinline__ void seg_prefix(const StepParams& p, int tt, int S, int nseg, int* segpre, int* wcnt) {
    const int tid = threadIdx.x;
    constexpr int kPer = 8;                             // segments per thread: nseg <= 8 * NT
    const int per = (nseg + NT - 1) / NT;
    int loc[kPer], sum = 0;
#pragma unroll
    for (int j = 0; j < kPer; j++) {
        const int i = tid * per + j;
        loc[j] = 0;
        if (j < per && i < nseg) {                      // wait for the segment: its header carries the expected sequence number
            const int g = i / S, sg = i - g * S;
            const volatile unsigned long long* h = reinterpret_cast<const volatile unsigned long long*>(
                p.plist + ((size_t)g * p.Ttot + tt) * publist_doubles(p.pV) + (size_t)sg * kSegDoubles);
            unsigned long long w = *h;
            for (uint32_t spins = 0; (w >> 8) != p.wait_seq; ++spins) {
                __nanosleep(100);
                // a lost peer must surface as an error, not as a hung GPU; the bound (some 40 s of polling) has to cover honest skew
                // between the processes (first-launch module loading, a rank writing output files)
                if (spins > (1u << 26)) __trap();
                w = *h;
            }
            loc[j] = (int)(w & 0xffu);
        }
        sum += loc[j];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += y;
    }
    if (p.pG > 1) __threadfence();                      // lists from other GPUs: the items behind the headers (written before them) are read after this point
    __syncthreads();                                    // wcnt / segpre of the previous trait are consumed
    if ((tid & 31) == 31) wcnt[tid >> 5] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int w = 0; w < (tid >> 5); w++) base += wcnt[w];
#pragma unroll
    for (int j = 0; j < kPer; j++) {
        const int i = tid * per + j;
        if (j < per && i < nseg) { segpre[i] = base; base += loc[j]; }
    }
    if (tid == NT - 1) segpre[nseg] = base;             // the last thread's running total is the grand total
    __syncthreads();
}

// item x (global virtual-rank order) of the pending lists of trait tt: its list (publishing GPU) and its address
__device__ __forceinline__ const double* seg_item(const StepParams& p, int tt, int S, int nseg, const int* segpre, int x, int* gpu) {
    int lo = 0, hi = nseg;                              // the segment holding item x: segpre[lo] <= x < segpre[lo + 1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (segpre[mid] <= x) lo = mid; else hi = mid;
    }
    const int g = lo / S, sg = lo - g * S;
    *gpu = g;
    return p.plist + ((size_t)g * p.Ttot + tt) * publist_doubles(p.pV) + (size_t)sg * kSegDoubles + 2 + 3 * (size_t)(x - segpre[lo]);
}

// ---- (a) pending updates: Phenotype::update_epsilon (phenotype.cpp:326-329,375-390) for every published marker
// of the previous step, in virtual-rank order, restricted to this CTA's rows.  All NT threads of the CTA work.
// With the fused increment exchange (p.xd_world > 1) the sums start from zero and go, instead of into the residuals, into the
// receive buffers of the GPUs that own the sub-slices of this CTA's rows (exchange_increments below finishes the job); the
// return value is the mask of the launch's traits whose list held anything.
template <int T, int NT>
__device__ int apply_pending(const StepParams& p, const PassRows& pr, PubStage* stage, PubInfo* info, uint32_t* lut,
                             uint32_t* bitmap, int* wcnt, uint8_t* bytes, int bytes_cap, bool profme) {
    long long tk = profme ? clock64() : 0;
    int nk = 40;
#define GMRM_ATICK() if (profme && nk < 60) { const long long t_ = clock64(); atomicAdd(&p.prof[nk++], (unsigned long long)(t_ - tk)); tk = t_; }
    const int tid = threadIdx.x;
    const int nr = pr.total, nq = nr * kRowBytes;       // quads of my rows
    constexpr int kQ = NT >= 512 ? 2 : 3;               // quads a thread owns per round: 13 rows (832 quads) in ONE round for 12 warps too
    const bool xd = p.xd_world > 1;
    const int xsub = (nq + max(p.xd_world, 1) - 1) / max(p.xd_world, 1);   // quads per sub-slice (fused increment exchange)
    int tmask = 0;
    // base-3 byte -> byte offsets (8 * dosage) of its four individuals into a PubStage
    for (int e = tid; e < kTabEntries; e += NT) {
        const uint32_t f = tri_to_fields(e);
        lut[e] = ((f & 3u) << 3) | (((f >> 2) & 3u) << 11) | (((f >> 4) & 3u) << 19) | (((f >> 6) & 3u) << 27);
    }
    const uint32_t lut_u32 = smem_u32(lut), stage_u32 = smem_u32(stage);
    // published markers handled per round: their bytes of this CTA's rows are staged in shared memory (the table
    // area, free at this point) with all loads in flight at once -- the columns were streamed a step ago and are
    // mostly out of L2, so fetching them entry by entry would expose one HBM round trip per 8 entries
    // pair tables at the head of the staging area (256-aligned: it starts at kTabBase): for entries 2i, 2i+1 of a round the 16
    // sums v_a[d_a] + v_b[d_b] (digit 3 = contributes nothing), 128 B -- one entry per 8-byte bank pair, so ANY mix of
    // indices within a half-warp is conflict-free -- and one look-up then serves TWO published markers of an individual
    double* ptab = reinterpret_cast<double*>(bytes);
    const uint32_t ptab_u32 = smem_u32(ptab);
    bytes += kPubCap * 64; bytes_cap -= kPubCap * 64;
    // exclusive prefix of the segment counts of all lists (GPU-major = global virtual-rank order), at the end of the staging area
    const int S = publist_segments(p.pV), nseg = p.pG * S;
    bytes_cap -= ((nseg + 1) * 4 + 15) & ~15;
    int* segpre = reinterpret_cast<int*>(bytes + bytes_cap);
    const int cap = max(1, min(kPubCap, bytes_cap / max(nq, 1)));
    for (int t = 0; t < T; t++) {
        const int tt = p.t0 + t;
        double* eps_t = p.eps + (int64_t)tt * p.npad;
        const uint8_t* mask_t = p.mask4 + (int64_t)tt * p.col_stride;
        seg_prefix<NT>(p, tt, S, nseg, segpre, wcnt);
        const int total = segpre[nseg];
        if (total > 0) tmask |= 1 << t;
        for (int q0 = 0; q0 < nq; q0 += kQ * NT) {
            double e[kQ][4], e0[kQ][4];
            uint32_t nmask[kQ];                          // 0x78 in byte k: individual k is not observed -> zero entry (15 of a pair table, 3 of a single one)
            int gq[kQ];                                 // global quad (byte of the column) of local quad q
            bool have[kQ];
#pragma unroll
            for (int qq = 0; qq < kQ; qq++) {
                const int q = q0 + qq * NT + tid;
                have[qq] = q < nq;
                gq[qq] = have[qq] ? global_row(pr, q >> 6) * kRowBytes + (q & 63) : 0;
                const uint32_t na = have[qq] ? mask_t[gq[qq]] : 0u;
                nmask[qq] = ((na & 1u) ? 0u : 0x78u) | ((na & 2u) ? 0u : 0x7800u) | ((na & 4u) ? 0u : 0x780000u) | ((na & 8u) ? 0u : 0x78000000u);
#pragma unroll
                for (int k = 0; k < 4; k++) { e[qq][k] = (have[qq] && !xd) ? eps_t[4 * (int64_t)gq[qq] + k] : 0.0; e0[qq][k] = e[qq][k]; }
            }
            GMRM_ATICK()   // [40] lut + eps/mask loads
            bool touched = false;
            GMRM_ATICK()   // [41] list headers
            {
                for (int r0 = 0; r0 < total; r0 += cap) {
                    const int n = min(cap, total - r0);
                    __syncthreads();
                    if (tid < n) {
                        int g;
                        const double* ip = seg_item(p, tt, S, nseg, segpre, r0 + tid, &g);
                        PubItem it;                          // .cg loads: peers rewrite this buffer between launches
                        it.lam = __ldcg(ip); it.mave = __ldcg(ip + 1);
                        const int2 cv = __ldcg(reinterpret_cast<const int2*>(ip + 2));
                        it.col = cv.x; it.v = cv.y;
                        PubStage& s = stage[tid];
                        const double mdb = -it.mave;                   // reference arithmetic: (mdb*b + a) * bs_, phenotype.cpp:328-329,388
                        s.v[0] = (mdb * 1.0 + 0.0) * it.lam;
                        s.v[1] = (mdb * 1.0 + 1.0) * it.lam;
                        s.v[2] = (mdb * 1.0 + 2.0) * it.lam;
                        s.v[3] = 0.0;
                        info[tid].col = it.col;
                        info[tid].nmiss_g = (uint32_t)g;               // the missing count is filled in with the column bytes below
                    } else if (tid < ((n + 7) & ~7)) {                 // pad the last group of 8 with no-op entries
                        stage[tid].v[0] = stage[tid].v[1] = stage[tid].v[2] = stage[tid].v[3] = 0.0;
                        info[tid].col = 0; info[tid].nmiss_g = 0u;
                    }
                    __syncthreads();
                    touched = true;
                    if (q0 == 0 || nq > kQ * NT) {            // (re)stage: once per round unless the rows need several quad chunks
                        uint32_t mo0 = 0, mo1 = 0;           // missing-list bounds of entry `tid`, in flight with the bytes
                        if (tid < n) { const uint32_t* mo = p.pmiss_off[info[tid].nmiss_g & 15u]; mo0 = mo[info[tid].col]; mo1 = mo[info[tid].col + 1]; }
                        const int npiece = n * nr * 4;       // 16-byte pieces: entry x local row x 4
                        for (int i0 = tid; i0 < npiece; i0 += 8 * NT) {
                            uint4 v[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int i = i0 + j * NT;
                                if (i < npiece) {
                                    const int en = i / (nr * 4), rem = i - en * (nr * 4), lr = rem >> 2, part = rem & 3;
                                    const uint8_t* src = p.pbed[info[en].nmiss_g & 15u] + (int64_t)info[en].col * p.col_stride + (int64_t)global_row(pr, lr) * kRowBytes + part * 16;
                                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(src));
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int i = i0 + j * NT;
                                if (i < npiece) *reinterpret_cast<uint4*>(bytes + (size_t)i * 16) = v[j];   // == entry*nq + lr*64 + part*16
                            }
                        }
                        if (tid < n) info[tid].nmiss_g = ((mo1 - mo0) << 4) | (info[tid].nmiss_g & 15u);
                        for (int i = tid; i < ((n + 7) & ~7) * 8; i += NT) {     // 16 sums per pair of entries (padded entries are all-zero)
                            const int pp = i >> 4, d1 = i & 3, d2 = (i >> 2) & 3;
                            ptab[i] = stage[2 * pp].v[d1] + stage[2 * pp + 1].v[d2];
                        }
                        __syncthreads();
                    }
                    GMRM_ATICK()   // [42] stage fill + column bytes
                    for (int g0 = 0; g0 < n; g0 += 8) {
                        uint32_t by[8][kQ];
#pragma unroll
                        for (int j = 0; j < 8; j++)
#pragma unroll
                            for (int qq = 0; qq < kQ; qq++) {
                                by[j][qq] = 0;
                                if (g0 + j < n && have[qq]) by[j][qq] = bytes[(size_t)(g0 + j) * nq + q0 + qq * NT + tid];
                            }
                        const uint32_t gbase = stage_u32 + (uint32_t)g0 * 32u;   // 256-aligned: stage is, g0 is a multiple of 8
                        // fast path (no entry of the group has missing genotypes): one branch-free block for the 8
                        // entries, two per look-up through the pair tables -- padded entries are all-zero, unobserved /
                        // absent quads hit the zero entry -- so that later look-ups are in flight while earlier ones are added
                        uint32_t anymiss = 0;
#pragma unroll
                        for (int j = 0; j < 8; j++) anymiss |= info[g0 + j].nmiss_g >> 4;
                        if (!anymiss) {
#define GMRM_APPLY_PAIR(J)                                                                                          \
    _Pragma("unroll") for (int qq = 0; qq < kQ; qq++) {                                                            \
        uint32_t o1, o2;                                                                                          \
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(o1) : "r"(lut_u32 + by[2 * J][qq] * 4u));                   \
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(o2) : "r"(lut_u32 + by[2 * J + 1][qq] * 4u));               \
        const uint32_t off = (o2 * 4u + o1) | nmask[qq];   /* byte k: 8 * (d_a + 4 d_b) of individual k, <= 80 */  \
        e[qq][0] += lds_f64_imm<J * 128>(byte_into<0>(off, pbase));   /* * na (phenotype.cpp:388) */              \
        e[qq][1] += lds_f64_imm<J * 128>(byte_into<1>(off, pbase));                                               \
        e[qq][2] += lds_f64_imm<J * 128>(byte_into<2>(off, pbase));                                               \
        e[qq][3] += lds_f64_imm<J * 128>(byte_into<3>(off, pbase));                                               \
    }
                            const uint32_t pbase = ptab_u32 + (uint32_t)g0 * 64u;    // 512-aligned: 4 pair tables per group of 8
                            GMRM_APPLY_PAIR(0) GMRM_APPLY_PAIR(1) GMRM_APPLY_PAIR(2) GMRM_APPLY_PAIR(3)
#undef GMRM_APPLY_PAIR
                            continue;
                        }
#define GMRM_APPLY(J)                                                                                              \
    if (g0 + J < n) {                                                                                             \
        const uint32_t nmiss = info[g0 + J].nmiss_g >> 4, pgpu = info[g0 + J].nmiss_g & 15u;                      \
        const bool hasmiss = nmiss != 0; /* CTA-uniform */                                                        \
        if (hasmiss) {                                                                                            \
            for (int i = tid; i < nr * 8; i += NT) bitmap[i] = 0u;                                                \
            __syncthreads();                                                                                      \
            const uint32_t mo = p.pmiss_off[pgpu][info[g0 + J].col];                                              \
            for (uint32_t i = tid; i < nmiss; i += NT) {                                                          \
                const int ind = (int)p.pmiss_idx[pgpu][mo + i], grow = ind >> 8;                                  \
                for (int qp = 0; qp < pr.npass; qp++)                                                             \
                    if (grow >= pr.start[qp] && grow < pr.start[qp] + pr.count[qp]) {                             \
                        const int loc = (pr.base[qp] + grow - pr.start[qp]) * kRowInd + (ind & 255);              \
                        atomicOr(&bitmap[loc >> 5], 1u << (loc & 31));                                            \
                    }                                                                                             \
            }                                                                                                     \
            __syncthreads();                                                                                      \
        }                                                                                                         \
        _Pragma("unroll") for (int qq = 0; qq < kQ; qq++) {                                                        \
            if (have[qq]) {                                                                                       \
                uint32_t off;                                                                                     \
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(off) : "r"(lut_u32 + by[J][qq] * 4u));              \
                off |= nmask[qq] & 0x18181818u;                                                                   \
                if (hasmiss) {   /* missing here: a = b = 0, no change */                                         \
                    const int q = q0 + qq * NT + tid;                                                             \
                    const uint32_t sk = (bitmap[q >> 3] >> ((q & 7) * 4)) & 0xfu;                                 \
                    off |= ((sk & 1u) ? 0x18u : 0u) | ((sk & 2u) ? 0x1800u : 0u) | ((sk & 4u) ? 0x180000u : 0u) | ((sk & 8u) ? 0x18000000u : 0u); \
                }                                                                                                 \
                e[qq][0] += lds_f64_imm<J * 32>(byte_into<0>(off, gbase));   /* * na (phenotype.cpp:388) */       \
                e[qq][1] += lds_f64_imm<J * 32>(byte_into<1>(off, gbase));                                        \
                e[qq][2] += lds_f64_imm<J * 32>(byte_into<2>(off, gbase));                                        \
                e[qq][3] += lds_f64_imm<J * 32>(byte_into<3>(off, gbase));                                        \
            }                                                                                                     \
        }                                                                                                         \
        if (hasmiss) __syncthreads();                                                                             \
    }
                        GMRM_APPLY(0) GMRM_APPLY(1) GMRM_APPLY(2) GMRM_APPLY(3) GMRM_APPLY(4) GMRM_APPLY(5) GMRM_APPLY(6) GMRM_APPLY(7)
#undef GMRM_APPLY
                    }
                }
            }
            GMRM_ATICK()   // [43] apply
            if (xd) {                                   // hop 1: the increments of quad q (zeros if nothing was published) go to the GPU that owns q's sub-slice
#pragma unroll
                for (int qq = 0; qq < kQ; qq++) {
                    if (!have[qq]) continue;
                    double* dst = p.xrecv[(q0 + qq * NT + tid) / xsub] + (((size_t)(p.row_seq & 1) * p.xd_world + p.xd_rank) * p.Ttot + tt) * p.npad + 4 * (int64_t)gq[qq];
                    st_v4_f64(dst, e[qq][0], e[qq][1], e[qq][2], e[qq][3]);
                }
            } else if (touched) {
#pragma unroll
                for (int qq = 0; qq < kQ; qq++) {
                    if (!have[qq]) continue;
#pragma unroll
                    for (int k = 0; k < 4; k++) eps_t[4 * (int64_t)gq[qq] + k] = e[qq][k];
                    if (p.delta) {
                        double* d = p.delta + (int64_t)tt * p.npad + 4 * (int64_t)gq[qq];
#pragma unroll
                        for (int k = 0; k < 4; k++) d[k] += e[qq][k] - e0[qq][k];   // what this shard changed since the last exchange
                    }
                }
            }
        }
    }
    return tmask;
}

// ---- fused increment exchange, second half (see StepParams::xd_world): reduction of this GPU's sub-slice of the CTA's rows in
// GPU order, hop 2 (new residuals into every GPU's landing buffer), then the CTA's rows from its own landing buffer into the
// residual array.  Data-driven: every wait is a poll of the payload itself.  Whole CTA; ends with a barrier.  `red` is scratch
// (the staging area: xd_world * quads of a sub-slice * 32 bytes).
__device__ __forceinline__ bool xd_ready(const unsigned long long (&w)[4]) {
    return w[0] != kXdSentinel && w[1] != kXdSentinel && w[2] != kXdSentinel && w[3] != kXdSentinel;
}
__device__ __forceinline__ void xd_wait(const double* src, unsigned long long (&w)[4]) {
    ld_poll_v4(src, w);
    for (uint32_t spins = 0; !xd_ready(w); ++spins) {
        __nanosleep(40);
        if (spins > (1u << 26)) __trap();                // a lost peer must surface as an error, not as a hung GPU
        ld_poll_v4(src, w);
    }
}
template <int T, int NT>
__device__ void exchange_increments(const StepParams& p, const PassRows& pr, double* red, int red_bytes) {
    const int tid = threadIdx.x;
    const int G = p.xd_world, me = p.xd_rank, par = (int)(p.row_seq & 1);
    const int nq = pr.total * kRowBytes, xsub = (nq + G - 1) / G;
    const int q_lo = min(nq, me * xsub), q_hi = min(nq, q_lo + xsub), nmine = q_hi - q_lo;
    const double sent = __longlong_as_double((long long)kXdSentinel);
    const int cap = max(1, red_bytes / (G * 32));        // quads of the sub-slice per round of the scratch area
    __syncthreads();                                     // the staging area is free
    for (int t = 0; t < T; t++)
    for (int c0 = 0; c0 < nmine; c0 += cap) {
        const int64_t tb = (int64_t)(p.t0 + t) * p.npad;
        const int nc = min(cap, nmine - c0), qc = q_lo + c0;
        // my sub-slice: one thread per (source GPU, quad) fetches, a barrier, one thread per quad adds in GPU order and sends
        for (int i = tid; i < nc * G; i += NT) {
            const int g = i / nc, q = qc + (i - g * nc);
            const int64_t o = tb + 4 * ((int64_t)global_row(pr, q >> 6) * kRowBytes + (q & 63));
            double* src = p.xrecv[me] + ((size_t)(par * G + g) * p.Ttot) * p.npad + o;
            unsigned long long w[4];
            xd_wait(src, w);
            st_v4_f64(src, sent, sent, sent, sent);      // consumed: the slot is rewritten two launches from now
            double* r = red + (size_t)i * 4;
            r[0] = __longlong_as_double((long long)w[0]); r[1] = __longlong_as_double((long long)w[1]);
            r[2] = __longlong_as_double((long long)w[2]); r[3] = __longlong_as_double((long long)w[3]);
        }
        __syncthreads();
        for (int i = tid; i < nc; i += NT) {
            const int q = qc + i;
            const int64_t o = tb + 4 * ((int64_t)global_row(pr, q >> 6) * kRowBytes + (q & 63));
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int g = 0; g < G; g++) {                // fixed order: every replica receives the same bits
                const double* r = red + ((size_t)g * nc + i) * 4;
                s0 += r[0]; s1 += r[1]; s2 += r[2]; s3 += r[3];
            }
            const double2 ea = __ldcg(reinterpret_cast<const double2*>(p.eps + o)), eb = __ldcg(reinterpret_cast<const double2*>(p.eps + o + 2));
            const double n0 = ea.x + s0, n1 = ea.y + s1, n2 = eb.x + s2, n3 = eb.y + s3;
            for (int g = 0; g < G; g++) st_v4_f64(p.xland[g] + ((size_t)par * p.Ttot) * p.npad + o, n0, n1, n2, n3);
        }
        __syncthreads();                                 // `red` is reused by the next round
    }
    // every quad of my rows, from whichever GPU reduced it
    for (int t = 0; t < T; t++) {
        const int64_t tb = (int64_t)(p.t0 + t) * p.npad;
        for (int q0 = 0; q0 < nq; q0 += 2 * NT) {
            int64_t o[2]; unsigned long long w[2][4]; bool have[2];
#pragma unroll
            for (int qq = 0; qq < 2; qq++) {
                const int q = q0 + qq * NT + tid;
                have[qq] = q < nq;
                o[qq] = have[qq] ? tb + 4 * ((int64_t)global_row(pr, q >> 6) * kRowBytes + (q & 63)) : 0;
                if (have[qq]) ld_poll_v4(p.xland[me] + ((size_t)par * p.Ttot) * p.npad + o[qq], w[qq]);
            }
#pragma unroll
            for (int qq = 0; qq < 2; qq++) {
                if (!have[qq]) continue;
                double* src = p.xland[me] + ((size_t)par * p.Ttot) * p.npad + o[qq];
                if (!xd_ready(w[qq])) xd_wait(src, w[qq]);
                st_v4_f64(src, sent, sent, sent, sent);
                st_v4_f64(p.eps + o[qq], __longlong_as_double((long long)w[qq][0]), __longlong_as_double((long long)w[qq][1]),
                          __longlong_as_double((long long)w[qq][2]), __longlong_as_double((long long)w[qq][3]));
            }
        }
    }
    __syncthreads();
}

// ---- (a') the same updates with several GPUs in the list exchange, ROW-SHARDED: every GPU would otherwise apply every GPU's
// updates to every row -- update work and NVLink column traffic that grow with the number of GPUs instead of shrinking.  Here
// GPU `rs_rank` applies ALL lists (global virtual-rank order) to the local rows rs_rank, rs_rank + rs_world, ... of the CTA
// only, stores the updated rows into the residual array of EVERY GPU over NVLink, and the same-index CTAs of the GPUs tell
// each other through flags that their rows have landed (every row is computed by exactly one GPU: all GPUs hold identical
// residuals by construction).  The few quads are spread over the CTA by splitting the staged entries of a round into
// entry groups whose partial increments are summed in a fixed order; missing genotypes (stored as dosage 0) are taken out
// again afterwards, entry by entry.
template <int T, int NT>
__device__ void apply_pending_sharded(const StepParams& p, const PassRows& pr, PubStage* stage, PubInfo* info, uint32_t* lut,
                                      int* wcnt, uint8_t* bytes, int bytes_cap) {
    const int tid = threadIdx.x, cta = blockIdx.x, nsm = gridDim.x;
    const int G = p.rs_world, me = p.rs_rank;
    const int nr = pr.total;
    const int nsub = nr > me ? (nr - me + G - 1) / G : 0;       // local rows me, me + G, ... of this CTA
    const int nq = nsub * kRowBytes;
    for (int e = tid; e < kTabEntries; e += NT) {
        const uint32_t f = tri_to_fields(e);
        lut[e] = ((f & 3u) << 3) | (((f >> 2) & 3u) << 11) | (((f >> 4) & 3u) << 19) | (((f >> 6) & 3u) << 27);
    }
    const uint32_t lut_u32 = smem_u32(lut);
    int EG = 1;                                                  // entry groups: 64 quads -> 8, 128 -> 4, up to 256 -> 2, more -> 1
    while (EG < 8 && nq * (EG * 2) <= NT) EG *= 2;
    const int W = NT / EG;                                       // threads per entry group == quads per chunk
    double* ptab = reinterpret_cast<double*>(bytes);             // pair tables (see apply_pending)
    const uint32_t ptab_u32 = smem_u32(ptab);
    bytes += kPubCap * 64; bytes_cap -= kPubCap * 64;
    double* redbuf = reinterpret_cast<double*>(bytes);           // [EG - 1][W][4] increments of the entry groups 1 ..
    bytes += (EG - 1) * W * 32; bytes_cap -= (EG - 1) * W * 32;
    double* corr = reinterpret_cast<double*>(bytes);             // [W][4] what the missing genotypes take out again
    bytes += W * 32; bytes_cap -= W * 32;
    const int S = publist_segments(p.pV), nseg = p.pG * S;
    bytes_cap -= ((nseg + 1) * 4 + 15) & ~15;
    int* segpre = reinterpret_cast<int*>(bytes + bytes_cap);
    const int cap = max(8, min(kPubCap, bytes_cap / W) & ~7);
    const int tq = tid % W, eg = tid / W;
    for (int t = 0; t < T && nq > 0; t++) {
        const int tt = p.t0 + t;
        const uint8_t* mask_t = p.mask4 + (int64_t)tt * p.col_stride;
        seg_prefix<NT>(p, tt, S, nseg, segpre, wcnt);
        const int total = segpre[nseg];
        if (total == 0) continue;                                // CTA-uniform
        for (int q0 = 0; q0 < nq; q0 += W) {
            const int q = q0 + tq;
            const bool have = q < nq;
            const int gq = have ? global_row(pr, me + (q >> 6) * G) * kRowBytes + (q & 63) : 0;   // global quad = byte of the column
            const uint32_t na = have ? mask_t[gq] : 0u;
            const uint32_t nmask = ((na & 1u) ? 0u : 0x78u) | ((na & 2u) ? 0u : 0x7800u) | ((na & 4u) ? 0u : 0x780000u) | ((na & 8u) ? 0u : 0x78000000u);
            const int rows_here = min(W >> 6, nsub - (q0 >> 6));
            double d[4] = {0.0, 0.0, 0.0, 0.0};
            if (eg == 0) { corr[tq * 4] = 0.0; corr[tq * 4 + 1] = 0.0; corr[tq * 4 + 2] = 0.0; corr[tq * 4 + 3] = 0.0; }
            for (int r0 = 0; r0 < total; r0 += cap) {
                const int n = min(cap, total - r0);
                __syncthreads();
                uint32_t mo0 = 0, mo1 = 0;
                if (tid < n) {
                    int g;
                    const double* ip = seg_item(p, tt, S, nseg, segpre, r0 + tid, &g);
                    PubItem it;                                  // .cg loads: peers rewrite this buffer between launches
                    it.lam = __ldcg(ip); it.mave = __ldcg(ip + 1);
                    const int2 cv = __ldcg(reinterpret_cast<const int2*>(ip + 2));
                    it.col = cv.x; it.v = cv.y;
                    PubStage& sg = stage[tid];
                    const double mdb = -it.mave;                 // reference arithmetic: (mdb*b + a) * bs_, phenotype.cpp:328-329,388
                    sg.v[0] = (mdb * 1.0 + 0.0) * it.lam;
                    sg.v[1] = (mdb * 1.0 + 1.0) * it.lam;
                    sg.v[2] = (mdb * 1.0 + 2.0) * it.lam;
                    sg.v[3] = 0.0;
                    info[tid].col = it.col;
                    info[tid].nmiss_g = (uint32_t)g;             // the missing count is filled in below
                    const uint32_t* mo = p.pmiss_off[g];         // bounds of the missing list: in flight with the column bytes
                    mo0 = mo[it.col]; mo1 = mo[it.col + 1];
                } else if (tid < ((n + 7) & ~7)) {               // pad the last group of 8 with no-op entries
                    stage[tid].v[0] = stage[tid].v[1] = stage[tid].v[2] = stage[tid].v[3] = 0.0;
                    info[tid].col = 0; info[tid].nmiss_g = 0u;
                }
                __syncthreads();
                {   // the entries' bytes of this chunk's rows: all loads in flight at once (remote shards' columns over NVLink)
                    const int npiece = n * rows_here * 4;        // 16-byte pieces: entry x chunk row x 4
                    for (int i0 = tid; i0 < npiece; i0 += 4 * NT) {
                        uint4 v[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int i = i0 + j * NT;
                            if (i < npiece) {
                                const int en = i / (rows_here * 4), rem = i - en * (rows_here * 4), k = rem >> 2, part = rem & 3;
                                const uint8_t* src = p.pbed[info[en].nmiss_g & 15u] + (int64_t)info[en].col * p.col_stride +
                                                     (int64_t)global_row(pr, me + ((q0 >> 6) + k) * G) * kRowBytes + part * 16;
                                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(src));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int i = i0 + j * NT;
                            if (i < npiece) {
                                const int en = i / (rows_here * 4), rem = i - en * (rows_here * 4);
                                *reinterpret_cast<uint4*>(bytes + (size_t)en * W + (size_t)rem * 16) = v[j];
                            }
                        }
                    }
                    for (int i = tid; i < ((n + 7) & ~7) * 8; i += NT) {     // 16 sums per pair of entries (padded entries are all-zero)
                        const int pp = i >> 4, d1 = i & 3, d2 = (i >> 2) & 3;
                        ptab[i] = stage[2 * pp].v[d1] + stage[2 * pp + 1].v[d2];
                    }
                    if (tid < n) info[tid].nmiss_g |= (mo1 - mo0) << 4;
                }
                __syncthreads();
                // this entry group's share of the round, in groups of 8 entries (4 pair tables each)
                const int ng8 = (n + 7) >> 3, per = (ng8 + EG - 1) / EG;
                for (int g8 = eg * per; g8 < min(ng8, (eg + 1) * per); g8++) {
                    const int g0 = g8 * 8;
                    uint32_t by[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) by[j] = (have && g0 + j < n) ? bytes[(size_t)(g0 + j) * W + tq] : 0u;
                    const uint32_t pbase = ptab_u32 + (uint32_t)g0 * 64u;
#define GMRM_SHARD_PAIR(J)                                                                                         \
    {                                                                                                             \
        uint32_t o1, o2;                                                                                          \
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(o1) : "r"(lut_u32 + by[2 * J] * 4u));                       \
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(o2) : "r"(lut_u32 + by[2 * J + 1] * 4u));                   \
        const uint32_t off = (o2 * 4u + o1) | nmask;                                                              \
        d[0] += lds_f64_imm<J * 128>(byte_into<0>(off, pbase));                                                   \
        d[1] += lds_f64_imm<J * 128>(byte_into<1>(off, pbase));                                                   \
        d[2] += lds_f64_imm<J * 128>(byte_into<2>(off, pbase));                                                   \
        d[3] += lds_f64_imm<J * 128>(byte_into<3>(off, pbase));                                                   \
    }
                    GMRM_SHARD_PAIR(0) GMRM_SHARD_PAIR(1) GMRM_SHARD_PAIR(2) GMRM_SHARD_PAIR(3)
#undef GMRM_SHARD_PAIR
                }
                // missing genotypes were stored as dosage 0 and have just received v[0]: take it out again, entry by entry
                // (the individuals of one entry are distinct; a barrier orders the entries)
                for (int en = 0; en < n; en++) {
                    const uint32_t nmiss = info[en].nmiss_g >> 4;
                    if (nmiss == 0) continue;                    // CTA-uniform
                    const uint32_t pgpu = info[en].nmiss_g & 15u, mo = p.pmiss_off[pgpu][info[en].col];
                    const double v0 = stage[en].v[0];
                    for (uint32_t i = tid; i < nmiss; i += NT) {
                        const int ind = (int)p.pmiss_idx[pgpu][mo + i], grow = ind >> 8;
                        for (int qp = 0; qp < pr.npass; qp++)
                            if (grow >= pr.start[qp] && grow < pr.start[qp] + pr.count[qp]) {
                                const int lr = pr.base[qp] + grow - pr.start[qp];
                                if (lr % G == me) {
                                    const int k = (lr - me) / G - (q0 >> 6);
                                    if (k >= 0 && k < rows_here) corr[(k * kRowBytes + ((ind & 255) >> 2)) * 4 + (ind & 3)] -= v0;
                                }
                            }
                    }
                    __syncthreads();
                }
            }
            // increments of the entry groups, in order, then the missing-genotype corrections; the rows go to every GPU
            if (eg > 0) { double* r = redbuf + ((size_t)(eg - 1) * W + tq) * 4; r[0] = d[0]; r[1] = d[1]; r[2] = d[2]; r[3] = d[3]; }
            __syncthreads();
            if (eg == 0 && have) {
                for (int g = 1; g < EG; g++) {
                    const double* r = redbuf + ((size_t)(g - 1) * W + tq) * 4;
                    d[0] += r[0]; d[1] += r[1]; d[2] += r[2]; d[3] += r[3];
                }
                const int64_t o = (int64_t)tt * p.npad + 4 * (int64_t)gq;
                const double2 ea = *reinterpret_cast<const double2*>(p.eps + o), eb = *reinterpret_cast<const double2*>(p.eps + o + 2);
                double2 na_, nb_;                                // unobserved individuals took the zero entry: no correction either
                na_.x = ea.x + (d[0] + ((na & 1u) ? corr[tq * 4] : 0.0));
                na_.y = ea.y + (d[1] + ((na & 2u) ? corr[tq * 4 + 1] : 0.0));
                nb_.x = eb.x + (d[2] + ((na & 4u) ? corr[tq * 4 + 2] : 0.0));
                nb_.y = eb.y + (d[3] + ((na & 8u) ? corr[tq * 4 + 3] : 0.0));
                for (int g = 0; g < G; g++) {
                    *reinterpret_cast<double2*>(p.peps[g] + o) = na_;
                    *reinterpret_cast<double2*>(p.peps[g] + o + 2) = nb_;
                }
                __threadfence_system();                          // by every warp that stored rows (one MEMBAR per warp), before the flags below
            }
            __syncthreads();                                     // corr / redbuf are reused by the next chunk
        }
    }
    // this CTA's rows are on their way to every GPU: fence, tell the same-index CTAs, wait for theirs
    __syncthreads();
    if (tid < G) {
        __threadfence_system();          // by the thread that writes the flag: the CTA's row stores (ordered before it by the barrier) first
        *reinterpret_cast<volatile unsigned long long*>(p.rflag_peer[tid] + (size_t)me * nsm + cta) = p.row_seq;
        const volatile unsigned long long* f = p.rflag_mine + (size_t)tid * nsm + cta;
        for (uint32_t spins = 0; *f < p.row_seq; ++spins) {
            __nanosleep(100);
            if (spins > (1u << 26)) __trap();                    // a lost peer must surface as an error, not as a hung GPU
        }
    }
    __threadfence();
    __syncthreads();
}

// Byte offset, inside the nr * 64 contiguous bytes a CTA owns of every column in a pass, of byte k of the word that lane l16
// (of the 16 lanes serving a marker) holds for table slot s.  Slots are grouped for vector loads: four slots are 16 contiguous
// bytes per lane, two are 8, one is 4 -- nr = 5: [16 lanes x 16 B][16 lanes x 4 B], nr = 3: [16 x 8 B][16 x 4 B].
__host__ __device__ __forceinline__ int chunk_offset(int nr, int s, int l16, int k) {
    int base = 0;
    if (nr >= 4) {
        if (s < 4) return 16 * l16 + 4 * s + k;
        base = 256; nr -= 4; s -= 4;
    }
    if (nr >= 2) {
        if (s < 2) return base + 8 * l16 + 4 * s + k;
        base += 128; s -= 2;
    }
    return base + 4 * l16 + k;
}

// ---- (b) tables of rows [row0, row0 + nrp) for T traits, by the NC consumer threads; es[t] accumulates this
// thread's share of sum eps
#ifndef GMRM_BUILD_PIPE
#define GMRM_BUILD_PIPE 1
#endif
template <int T, int NC>
__device__ __forceinline__ void build_tables(const StepParams& p, int row0, int nrp, double (&es)[T]) {
    const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    const int nunits = nrp * T * 4 * 3;
    constexpr int kStride = NC / 16;
    // a unit's four residuals: the loads of this half-warp's NEXT unit are issued before the current one's 27 entries are
    // computed and stored (a pass of 5 rows is two units per half-warp: without this the second L2 round trip is exposed)
    auto unit_ptr = [&](int u) -> const double* {
        const int line = u / 3, k = line & 3, slot = line >> 2, rr = slot / T, t = slot - rr * T;
        return p.eps + (int64_t)(p.t0 + t) * p.npad + ((int64_t)row0 * kRowBytes + chunk_offset(nrp, rr, l16, k)) * 4;
    };
    double2 n01 = make_double2(0.0, 0.0), n23 = make_double2(0.0, 0.0);
    if (hw < nunits) {   // .cg: in the multi-GPU exchanges these rows were just stored by other GPUs (they land in this GPU's L2)
        const double* e = unit_ptr(hw);
        n01 = __ldcg(reinterpret_cast<const double2*>(e)); n23 = __ldcg(reinterpret_cast<const double2*>(e + 2));
    }
    for (int u = hw; u < nunits; u += kStride) {
        const int d3 = u % 3, line = u / 3, k = line & 3, slot = line >> 2, rr = slot / T, t = slot - rr * T;
#if GMRM_BUILD_PIPE
        const double2 e01 = n01, e23 = n23;
        if (u + kStride < nunits) {
            const double* e = unit_ptr(u + kStride);
            n01 = __ldcg(reinterpret_cast<const double2*>(e)); n23 = __ldcg(reinterpret_cast<const double2*>(e + 2));
        }
#else       // measured alternative: every unit loads its own residuals when its turn comes
        const double* e_ = unit_ptr(u);
        const double2 e01 = __ldcg(reinterpret_cast<const double2*>(e_)), e23 = __ldcg(reinterpret_cast<const double2*>(e_ + 2));
#endif
        if (d3 == 0) {
            const double s4 = (e01.x + e01.y) + (e23.x + e23.y);
#pragma unroll
            for (int tt = 0; tt < T; tt++)
                if (tt == t) es[tt] += s4;
        }
        const double x0[3] = {0.0, e01.x, e01.x + e01.x}, x1[3] = {0.0, e01.y, e01.y + e01.y};
        const double x2[3] = {0.0, e23.x, e23.x + e23.x};
        const double x3 = d3 == 0 ? 0.0 : (d3 == 1 ? e23.y : e23.y + e23.y);
        const uint32_t base = (uint32_t)tab_imm(slot, k) + (uint32_t)(27 * d3) * 256u + (uint32_t)l16 * 8u;
#pragma unroll
        for (int d2 = 0; d2 < 3; d2++) {
            const double hi = x2[d2] + x3;
#pragma unroll
            for (int d1 = 0; d1 < 3; d1++)
#pragma unroll
                for (int d0 = 0; d0 < 3; d0++) {
                    const double val = (x0[d0] + x1[d1]) + hi;
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + (uint32_t)(d0 + 3 * d1 + 9 * d2) * 256u), "d"(val) : "memory");
                }
        }
    }
}

constexpr int kStepWarps = GMRM_STEP_WARPS, kStepThreads = kStepWarps * 32;
constexpr int kDirectMax = GMRM_STEP_DIRECT;   // direct rows per pass the kernel is built for (registers: 32 per row and thread)
constexpr int kPairs = kBatch / 2;

// L2 prefetch of this warp's first two batches of a pass (rows [row0, row0+nr) of their 16 columns each): issued
// before the pass's tables are built (and at kernel start for pass 0), it takes the HBM round trip of the first
// loads of a pass off the critical path.
__device__ __forceinline__ void prefetch_pass_head(const StepParams& p, int row0, int nr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = (p.V + kBatch - 1) / kBatch;
    if (warp >= kStepWarps) return;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int b = warp + k * kStepWarps, v = b * kBatch + (lane & (kBatch - 1)), g = lane / kBatch;
        if (b >= nb || v >= p.V) continue;
        const uint8_t* a = p.bed + (int64_t)max(p.cols[v], 0) * p.col_stride + (int64_t)row0 * kRowBytes;
        const int bytes = nr * kRowBytes;
        const int o1 = (32 / kBatch == 4 && g == 3) ? bytes - 1 : g * 128, o2 = 32 / kBatch == 4 ? 1 << 20 : (g ? bytes - 1 : 256);
        if (o1 < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o1));
        if (o2 < bytes && nr > 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o2));
    }
}

// ---- (c) stream the V columns through the tables of NR rows
// Markers [v0, v0 + Vc) of the step (a chunk whose Vc * T partial sums fit the shared-memory array `part`).
// A CTA's share of a column in a pass is one contiguous chunk of NR * 64 bytes.  Lane l (of the 16 that serve a marker) takes
// its words for the NR table slots with as few, as wide loads as the chunk allows -- four slots: 16 contiguous bytes (one
// LDG.128), two: 8 bytes, one: 4 bytes (chunk_offset below; build_tables uses the same map) -- so that a warp's load
// instruction covers whole 128-byte lines (the 4-byte form touched two half-used lines per instruction and cost twice the
// L1 wavefronts per byte).  Two register buffers take turns (A: in use, B: being loaded for the warp's next batch).
// The genotype stream is read once per step and never again before it has left L2 anyway: its lines carry the L2
// evict-first policy, so that what IS reused -- residuals, masks, the per-marker scalars the sampler reads at random, the
// published columns the sampler prefetched for the update phase -- stays resident next to 238 MB of streamed bytes per step.
#ifndef GMRM_STREAM_EVICT_FIRST
#define GMRM_STREAM_EVICT_FIRST 1
#endif
__device__ __forceinline__ uint64_t stream_policy() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint8_t* p, uint64_t pol) {
    uint32_t v;
#if GMRM_STREAM_EVICT_FIRST
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
#else
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
#endif
    return v;
}
template <int NR>
__device__ __forceinline__ void load_words(const uint8_t* pa, const uint8_t* pb, uint32_t (&w)[NR], uint64_t pol) {
    if constexpr (NR >= 4) {
#if GMRM_STREAM_EVICT_FIRST
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(pa), "l"(pol));
#else
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(pa));
#endif
        if constexpr (NR == 5) w[4] = ldg_stream_u32(pb, pol);
    } else if constexpr (NR >= 2) {
#if GMRM_STREAM_EVICT_FIRST
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(w[0]), "=r"(w[1]) : "l"(pa), "l"(pol));
#else
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(pa));
#endif
        if constexpr (NR == 3) w[2] = ldg_stream_u32(pb, pol);
    } else {
        w[0] = ldg_stream_u32(pa, pol);
    }
}

// DIRECT rows (ND > 0, one trait): rows of the pass that get no table.  Their genotypes come from a second copy kept in plain
// 2-bit dosage fields (StepParams::bed2), their residuals sit in registers -- lane l holds the 16 individuals of word l,
// pre-scaled -- and every genotype costs one byte permute plus one fp64 multiply-add on pipes the look-ups leave idle:
//   m_f = word & (0x03030303 << 2f)   isolates the fields f, f+4, f+8, f+12 in the four bytes (value d << 2f),
//   PRMT moves byte b of m_f into the low word of a double whose high word is 0: the denormal  d * 2^(2f) * 2^-1074,
//   DFMA with the weight  eps * 2^(kDirScale - 2f)  adds  d * eps * 2^(kDirScale - 1074)  exactly as  acc += d * eps  would
// (power-of-two scalings only), and the batch's direct sums are scaled back by 2^(1074 - kDirScale) when they join the
// look-up sums.  Missing genotypes are dosage 0 in both copies, residuals of unobserved individuals are 0.
constexpr int kDirScale = 960;
// D keeps its (zero) high word; its low word becomes byte B of m: the permute writes straight into the register pair
template <int B>
__device__ __forceinline__ void dir_operand(double& D, uint32_t m) {
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tprmt.b32 lo, %1, 0, %2;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+d"(D) : "r"(m), "n"(0x4440 | B));
}
template <int F>
__device__ __forceinline__ void dir_fields(uint32_t w, const double (&wt)[16], double (&D)[4], double& acc) {   // fields F, F+4, F+8, F+12 of the word
    const uint32_t m = w & (0x03030303u << (2 * F));
    dir_operand<0>(D[0], m); acc = fma(D[0], wt[F], acc);
    dir_operand<1>(D[1], m); acc = fma(D[1], wt[4 + F], acc);
    dir_operand<2>(D[2], m); acc = fma(D[2], wt[8 + F], acc);
    dir_operand<3>(D[3], m); acc = fma(D[3], wt[12 + F], acc);
}

// `es` non-null (GMRM_BUILD_FUSED=1, a measured alternative): the pass's tables are built HERE, by all threads, after the loads
// of every warp's first batch have been issued, so that their latency runs under the build -- followed by the barrier that
// publishes the tables.  A/B on one box at the UKB size: 61.1 ms per iteration against 60.4 with the build in front
// (profiles/r2_small_ab_summary.txt): off.
#ifndef GMRM_BUILD_FUSED
#define GMRM_BUILD_FUSED 0
#endif
// 1 (a measured alternative): the CTA's last pass stores every marker's sum straight into its global slot instead of the
// 2,048-store burst behind the final barrier.  A/B on one box: the step kernel got 1.5 us SLOWER (scattered stores inside the
// stream loop), profiles/r2_small_ab_summary.txt: off.
#ifndef GMRM_PARTIAL_DIRECT
#define GMRM_PARTIAL_DIRECT 0
#endif
template <int NR, int T, int ND = 0>
__device__ __forceinline__ void stream_rows(const StepParams& p, int row0, double* part, int v0, int Vc, int dslot = 0, double (*es)[T] = nullptr,
                                            bool last = false) {
    static_assert(ND == 0 || T == 1, "direct rows: one trait");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = lane >> 4, l16 = lane & 15;
    const uint32_t low = (uint32_t)l16 * 8u;
    const int nb = (Vc + kBatch - 1) / kBatch;
    int b = warp;
    if (b >= nb) {                                           // a warp without batches still builds and meets the barrier
        if (es != nullptr) { build_tables<T, kStepThreads>(p, row0, NR, *es); __syncthreads(); }
        return;
    }
    // Addresses: (column * nrows + row0) is a 32-bit ROW index (columns are whole rows: < 2^32 rows for any shard that fits a
    // GPU), so a lane's address is ONE multiply-add  lane_a + row * 64  on top of its 64-bit lane base.  The base is passed
    // through an empty asm statement: the compiler otherwise re-derives it from the kernel parameters in front of every load
    // (a dozen integer instructions per marker pair) to save two registers.
    const uint8_t* lane_a = p.bed + chunk_offset(NR, 0, l16, 0);
    asm volatile("" : "+l"(lane_a));
    const int d_b = chunk_offset(NR, NR - 1, l16, 0) - chunk_offset(NR, 0, l16, 0);    // the odd slot (NR 3, 5) behind the wide group
    const uint32_t nrows = (uint32_t)p.nrows, row0u = (uint32_t)row0;
    const int32_t* ccols = p.cols + v0;
    // direct rows: word l16 of row d of the second copy, weights of its 16 individuals
    // a zero neither compiler stage can fold (a launch parameter): the operands' high words live in registers instead of
    // being re-created by a move in front of every multiply-add
    [[maybe_unused]] const double dzero = __hiloint2double((int)p.zero, (int)p.zero);
    [[maybe_unused]] const uint8_t* lane_d = nullptr;
    [[maybe_unused]] double wt[ND > 0 ? ND : 1][16];
    if constexpr (ND > 0) {
        lane_d = p.bed2 + (int64_t)dslot * kRowBytes + l16 * 4;
        asm volatile("" : "+l"(lane_d));
#pragma unroll
        for (int d = 0; d < ND; d++) {
            const double* e = p.eps + (int64_t)p.t0 * p.npad + ((int64_t)(row0 + NR + d) * kRowInd + 16 * l16);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const double2 v = __ldcg(reinterpret_cast<const double2*>(e + j));
                wt[d][j] = scalbn(v.x, kDirScale - 2 * (j & 3));
                wt[d][j + 1] = scalbn(v.y, kDirScale - 2 * ((j + 1) & 3));
            }
        }
    }

    // raw column index (may be -1: no marker); the clamp is applied where the value is USED -- clamping here would make
    // the warp wait for this load at once instead of a batch later (measured: 17 % of the stream's stall samples)
    auto loadcols = [&](int bb) -> int {
        const int v = bb * kBatch + (l16 & (kBatch - 1));
        return (bb < nb && v < Vc) ? ccols[v] : 0;
    };
    const uint64_t pol = stream_policy();
    uint32_t WA[kPairs][NR + ND], WB[kPairs][NR + ND];
    auto load_pair = [&](int col, int i, uint32_t (&dst)[kPairs][NR + ND]) {      // col: the pair's column for this half-warp (raw)
        const uint32_t r = (uint32_t)max(col, 0) * nrows + row0u;
        const uint8_t* pa = lane_a + (uint64_t)r * kRowBytes;
        load_words<NR>(pa, pa + d_b, reinterpret_cast<uint32_t(&)[NR]>(dst[i]), pol);
        if constexpr (ND > 0) {
            const uint8_t* pd = lane_d + (uint64_t)((uint32_t)max(col, 0) * (uint32_t)p.drows) * kRowBytes;
#pragma unroll
            for (int d = 0; d < ND; d++) dst[i][NR + d] = ldg_stream_u32(pd + d * kRowBytes, pol);
        }
    };
    {
        const int c0 = loadcols(b);
#pragma unroll
        for (int i = 0; i < kPairs; i++) load_pair(__shfl_sync(0xffffffffu, c0, 2 * i + h), i, WA);
    }
    int cn = loadcols(b + kStepWarps);
    if (es != nullptr) { build_tables<T, kStepThreads>(p, row0, NR, *es); __syncthreads(); }
    // L2 prefetch kPfAhead batches ahead of the register loads (which run one batch ahead): the 16 warps then keep
    // ~192 KB per SM in flight towards HBM instead of 64 KB
    constexpr int kPfAhead = 4;
    int pcol = p.pf ? loadcols(b + kPfAhead * kStepWarps) : 0;
    // lane -> (marker of the batch, 128-byte line of its chunk): one prefetch per lane covers the chunk's (at most four) lines
    const int d_pf = min((lane / kBatch) * 128 * (kBatch / 8), NR * kRowBytes - 1) - chunk_offset(NR, 0, l16, 0);
    [[maybe_unused]] const bool hi8 = l16 & 8, hi4 = l16 & 4, hi2 = l16 & 2;
    // pair whose total this lane ends up with
    const int own = kPairs == 8 ? ((l16 >> 3) & 1) * 4 + ((l16 >> 2) & 1) * 2 + ((l16 >> 1) & 1)
                  : kPairs == 4 ? ((l16 >> 3) & 1) * 2 + ((l16 >> 2) & 1) : ((l16 >> 1) & 1);

    // one batch: look-ups on W while the warp's next batch (if there is one: `more`, a compile-time flag -- predicated loads
    // into Wn cost a register copy each) is loaded into Wn
    auto batch = [&](int bb, uint32_t (&W)[kPairs][NR + ND], uint32_t (&Wn)[kPairs][NR + ND], auto more_c) {
        constexpr bool more = decltype(more_c)::value;
        const int cnow = cn;
        if constexpr (more) cn = loadcols(bb + 2 * kStepWarps);
        if (p.pf && bb + kPfAhead * kStepWarps < nb) {
            const uint32_t r = (uint32_t)max(pcol, 0) * nrows + row0u;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(lane_a + (uint64_t)r * kRowBytes + d_pf));
            pcol = loadcols(bb + (kPfAhead + 1) * kStepWarps);
        }
        double acc[kPairs][T];
        [[maybe_unused]] double accd[kPairs];
        [[maybe_unused]] double Dop[4] = {dzero, dzero, dzero, dzero};   // multiplier operands: high words stay 0
#pragma unroll
        for (int i = 0; i < kPairs; i++) {
            accd[i] = 0.0;
#pragma unroll
            for (int t = 0; t < T; t++) acc[i][t] = 0.0;
        }

        // the loads of the next batch are spread over the first look-up groups (a burst of loads at the top of a batch
        // filled the load/store queue in front of the other warps' look-ups)
#define GMRM_LOOKUP(RR, K)                                                    \
    _Pragma("unroll") for (int i = 0; i < kPairs; i++)                             \
        lookup_traits<RR * T, K, T>(acc[i], tab_addr<K>(W[i][RR], low));      \
    {                                                                         \
        constexpr int G = RR * 4 + K;                                         \
        if constexpr (G < kPairs) {                                           \
            const int c_ = __shfl_sync(0xffffffffu, cnow, 2 * G + h);         \
            if constexpr (more) load_pair(c_, G, Wn);                         \
        }                                                                     \
    }
        // direct work is dealt out between the look-up groups: every (NR / ND)-th group is followed by one of the ND * 4 field
        // groups of the direct rows (4 per row: fields F, F+4, F+8, F+12 of every pair's word)
#define GMRM_DIRECT(G_)                                                                           \
    if constexpr (ND > 0) {                                                                       \
        constexpr int kEvery = NR / ND > 0 ? NR / ND : 1, D_ = (G_) / kEvery;                     \
        if constexpr ((G_) % kEvery == 0 && D_ < ND * 4) {                                        \
            _Pragma("unroll") for (int i = 0; i < kPairs; i++)                                    \
                dir_fields<D_ & 3>(W[i][NR + (D_ >> 2)], wt[D_ >> 2], Dop, accd[i]);              \
        }                                                                                         \
    }
#define GMRM_ROW(RR)                                                          \
    if constexpr (RR < NR) { GMRM_LOOKUP(RR, 0) GMRM_DIRECT(RR * 4 + 0) GMRM_LOOKUP(RR, 1) GMRM_DIRECT(RR * 4 + 1) \
                             GMRM_LOOKUP(RR, 2) GMRM_DIRECT(RR * 4 + 2) GMRM_LOOKUP(RR, 3) GMRM_DIRECT(RR * 4 + 3) }
        GMRM_ROW(0) GMRM_ROW(1) GMRM_ROW(2) GMRM_ROW(3) GMRM_ROW(4)
#undef GMRM_DIRECT
        if constexpr (ND > 0) {
#pragma unroll
            for (int i = 0; i < kPairs; i++) acc[i][0] = fma(accd[i], 0x1p114, acc[i][0]);       // 2^(1074 - kDirScale)
        }
#undef GMRM_ROW
#undef GMRM_LOOKUP
        if constexpr (NR * 4 < kPairs) {                  // fewer look-up groups than pairs (NR == 1, batches of 16)
#pragma unroll
            for (int i = NR * 4; i < kPairs; i++) {
                const int c_ = __shfl_sync(0xffffffffu, cnow, 2 * i + h);
                if constexpr (more) load_pair(c_, i, Wn);
            }
        }

        // 16-lane transposed butterfly: the pair accumulators -> the total of pair `own` (fixed order: reproducible)
#pragma unroll
        for (int t = 0; t < T; t++) {
            double b1;
            if constexpr (kPairs == 8) {
                double b4[4], b2[2];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double keep = hi8 ? acc[4 + j][t] : acc[j][t], send = hi8 ? acc[j][t] : acc[4 + j][t];
                    b4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const double keep = hi4 ? b4[2 + j] : b4[j], send = hi4 ? b4[j] : b4[2 + j];
                    b2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                const double keep = hi2 ? b2[1] : b2[0], send = hi2 ? b2[0] : b2[1];
                b1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            } else if constexpr (kPairs == 4) {
                double b2[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const double keep = hi8 ? acc[2 + j][t] : acc[j][t], send = hi8 ? acc[j][t] : acc[2 + j][t];
                    b2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                const double keep = hi4 ? b2[1] : b2[0], send = hi4 ? b2[0] : b2[1];
                b1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                b1 += __shfl_xor_sync(0xffffffffu, b1, 2);
            } else {
                const double keep = hi8 ? acc[1][t] : acc[0][t], send = hi8 ? acc[0][t] : acc[1][t];
                b1 = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                b1 += __shfl_xor_sync(0xffffffffu, b1, 4);
                b1 += __shfl_xor_sync(0xffffffffu, b1, 2);
            }
            b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
            const int v = bb * kBatch + 2 * own + h;
            if ((l16 & (kPairs == 8 ? 1 : kPairs == 4 ? 3 : 7)) == 0 && v < Vc) {
                // the CTA's last pass (GMRM_PARTIAL_DIRECT): the marker's sum goes straight to its global slot -- the stores
                // spread over the pass instead of a 2,048-store burst behind the final barrier
                if (GMRM_PARTIAL_DIRECT && last) p.partial[((int64_t)(v0 + v) * p.Ttot + p.t0 + t) * gridDim.x + blockIdx.x] = part[v * T + t] + b1;
                else part[v * T + t] += b1;
            }
        }
    };
    for (;;) {
        if (b + kStepWarps >= nb) { batch(b, WA, WB, std::false_type{}); break; }
        batch(b, WA, WB, std::true_type{});
        b += kStepWarps;
        if (b + kStepWarps >= nb) { batch(b, WB, WA, std::false_type{}); break; }
        batch(b, WB, WA, std::true_type{});
        b += kStepWarps;
    }
}

template <int T>
__global__ void __launch_bounds__(kStepThreads, 1) step_kernel(const StepParams p) {
    constexpr int NT = kStepThreads;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t b0 = smem_u32(smem_raw);
    if (b0 > kTabBase) {                                  // the LDS immediates assume tables at absolute address kTabBase
        if (threadIdx.x == 0) atomicExch(p.err, 10);
        return;
    }
    const int nsm = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    const int npass = p.npass;
    // shared memory carve-up (absolute addresses: kTabBase is 1024; `stage` stays 256-aligned)
    uint8_t* tabs = smem_raw + (kTabBase - b0);
    const int area = step_area_bytes(p.V, T, p.rows_per_pass, npass);                        // tables / update staging
    PubStage* stage = reinterpret_cast<PubStage*>(tabs + (size_t)area);                      // 256-aligned
    double* part = reinterpret_cast<double*>(stage + kPubCap);
    // a step of more markers than `part` holds sums for is streamed in chunks of kPartSmemDoubles / T markers per pass; each
    // chunk's sums are then added to the CTA's slots of the global partial array (L2-resident) before the next chunk
    const int Vc_max = kPartSmemDoubles / T, npart = min(p.V, Vc_max) * T;
    const bool chunked = p.V > Vc_max;
    PubInfo* info = reinterpret_cast<PubInfo*>(part + (size_t)npart);
    uint32_t* lut = reinterpret_cast<uint32_t*>(info + kPubCap);
    PassRows& pr = *reinterpret_cast<PassRows*>(lut + 82);
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(&pr + 1);
    double* red = reinterpret_cast<double*>(bitmap + ((npass * p.rows_per_pass * 8 + 1) & ~1));
    int* wcnt = reinterpret_cast<int*>(red + 32);                  // [4 windows][warps]

    if (tid == 0) {
        int tot = 0;
        all_pass_rows(p.nrows, npass, nsm, cta, pr.start, pr.count);
        for (int q = 0; q < npass; q++) {
            pr.base[q] = tot;
            tot += pr.count[q];
        }
        pr.npass = npass; pr.total = tot;
    }
    __syncthreads();
    const int nr = pr.total;
    const bool profme = p.prof && tid == 0 && (cta == 0 || cta == nsm / 2);
    long long tk = profme ? clock64() : 0;
    int nk = 8;
#define GMRM_TICK() if (profme) { const long long t_ = clock64(); atomicAdd(&p.prof[nk++], (unsigned long long)(t_ - tk)); tk = t_; }
    if (p.V > 0 && p.pf && npass > 0 && pr.count[0] > 0) prefetch_pass_head(p, pr.start[0], pr.count[0]);
    for (int i = tid; i < npart; i += NT) part[i] = 0.0;
    // everything above touched only this CTA's shared memory, the launch parameters and read-only inputs (the step table,
    // the genotypes): it may run while the sampler kernel of the previous step is still finishing
    pdl_wait();
    pdl_trigger();                                        // the sampler's CTAs may be scheduled as soon as SMs free up; they wait likewise
    if (p.merge_tot != nullptr && nr > 0) {               // residual deltas of the last exchange (all-reduced): what the OTHER shards changed
        for (int t = 0; t < T; t++) {
            const int64_t tb = (int64_t)(p.t0 + t) * p.npad;
            for (int q = tid; q < nr * kRowBytes; q += NT) {
                const int64_t o = tb + 4 * ((int64_t)global_row(pr, q >> 6) * kRowBytes + (q & 63));
                double2* e2 = reinterpret_cast<double2*>(p.eps + o);
                double2* l2 = reinterpret_cast<double2*>(p.delta + o);
                const double2* t2 = reinterpret_cast<const double2*>(p.merge_tot + o);
                const double2 ea = e2[0], eb = e2[1], la = l2[0], lb = l2[1], ta = t2[0], tb2 = t2[1];
                e2[0] = make_double2(ea.x + (ta.x - la.x), ea.y + (ta.y - la.y));
                e2[1] = make_double2(eb.x + (tb2.x - lb.x), eb.y + (tb2.y - lb.y));
                l2[0] = make_double2(0.0, 0.0); l2[1] = make_double2(0.0, 0.0);
            }
        }
        __syncthreads();
    }
    if (p.rs_world > 1 && p.pG * p.pV > 0) apply_pending_sharded<T, NT>(p, pr, stage, info, lut, wcnt, tabs, area);
    else
    if (p.pG * p.pV > 0 && nr > 0) {
        apply_pending<T, NT>(p, pr, stage, info, lut, bitmap, wcnt, tabs, area, p.prof && tid == 0 && (cta == 0 || cta == nsm / 2));
        if (p.xd_world > 1) exchange_increments<T, NT>(p, pr, reinterpret_cast<double*>(tabs), area);   // every GPU's CTA `cta` owns the same rows: all of them get here
    }
    __syncthreads();                                      // residual writes of (a) are visible to the whole CTA
    GMRM_TICK()                                           // [8] prologue + update phase
    if (p.V == 0) return;

    double es[T];
#pragma unroll
    for (int t = 0; t < T; t++) es[t] = 0.0;
    bool flushed = false;                                 // chunked steps: the global slots hold the sums of earlier passes
    int last_pass = -1;                                   // the CTA's last pass with rows: its stream writes the partial sums out itself
#if GMRM_PARTIAL_DIRECT
    if (!chunked)
        for (int q = 0; q < npass; q++)
            if (pr.count[q] > 0) last_pass = q;
#endif
    for (int pass = 0; pass < npass; pass++) {
        // rows of the pass beyond the table slots are DIRECT rows (hybrid plan, one trait): no table, genotypes from bed2
        // hybrid plan (one trait): the LAST row of a pass of two or more rows is a DIRECT row: no table, genotypes from bed2
        const int r_lo = pr.start[pass], nall = pr.count[pass], ndp = (kDirectMax >= 1 && T == 1 && p.ndir > 0 && nall >= 2) ? 1 : 0, nrp = nall - ndp;
        const int dslot = pass * nsm + cta;
        const bool fused = GMRM_BUILD_FUSED && ndp == 0 && nrp > 0;   // the stream builds the tables itself (see stream_rows)
        if (!fused) build_tables<T, NT>(p, r_lo, nrp, es);
        if (ndp > 0) {                                    // the direct rows' share of sum eps (their tables would have added it)
            for (int i = tid; i < ndp * kRowBytes; i += NT) {
                const double* e = p.eps + (int64_t)p.t0 * p.npad + ((int64_t)(r_lo + nrp) * kRowBytes + i) * 4;
                const double2 a = __ldcg(reinterpret_cast<const double2*>(e)), b = __ldcg(reinterpret_cast<const double2*>(e + 2));
                es[0] += (a.x + a.y) + (b.x + b.y);
            }
        }
        if (!fused) __syncthreads();
        GMRM_TICK()                                       // [9 + 3*pass] sync + build
        for (int v0 = 0; v0 < p.V; v0 += Vc_max) {
            const int Vc = min(Vc_max, p.V - v0);
            double (*esb)[T] = (fused && v0 == 0) ? &es : nullptr;
            const bool lastp = pass == last_pass && ndp == 0;
            if constexpr (T == 1 && kDirectMax >= 1) {
                if (ndp == 1) {
                    switch (nrp) {
                    case 1: stream_rows<1, 1, 1>(p, r_lo, part, v0, Vc, dslot); break;
                    case 2: stream_rows<2, 1, 1>(p, r_lo, part, v0, Vc, dslot); break;
                    case 3: stream_rows<3, 1, 1>(p, r_lo, part, v0, Vc, dslot); break;
                    case 4: stream_rows<4, 1, 1>(p, r_lo, part, v0, Vc, dslot); break;
                    default: break;
                    }
                }
            }
            if (ndp == 0)
            switch (nrp) {
            case 1: stream_rows<1, T>(p, r_lo, part, v0, Vc, 0, esb, lastp); break;
            case 2: if constexpr (2 * T <= kMaxSlots) stream_rows<2, T>(p, r_lo, part, v0, Vc, 0, esb, lastp); break;
            case 3: if constexpr (3 * T <= kMaxSlots) stream_rows<3, T>(p, r_lo, part, v0, Vc, 0, esb, lastp); break;
            case 4: if constexpr (4 * T <= kMaxSlots) stream_rows<4, T>(p, r_lo, part, v0, Vc, 0, esb, lastp); break;
            case 5: if constexpr (5 * T <= kMaxSlots) stream_rows<5, T>(p, r_lo, part, v0, Vc, 0, esb, lastp); break;
            default: break;
            }
            if (chunked && nall > 0) {                    // this chunk's sums of this pass -> the CTA's global slots
                __syncthreads();
                for (int i = tid; i < Vc * T; i += NT) {
                    const int v = v0 + i / T, t = i % T;
                    double* slot = p.partial + ((int64_t)v * p.Ttot + p.t0 + t) * nsm + cta;
                    *slot = flushed ? *slot + part[i] : part[i];
                    part[i] = 0.0;
                }
                __syncthreads();
            }
        }
        if (chunked && nall > 0) flushed = true;
        GMRM_TICK()                                       // [10 + 3*pass] this warp's streaming
        if (p.pf && pass + 1 < npass && pr.count[pass + 1] > 0) prefetch_pass_head(p, pr.start[pass + 1], pr.count[pass + 1]);
        if (pass + 1 < npass) __syncthreads();            // everyone is done with these tables
        GMRM_TICK()                                       // [11 + 3*pass] waiting for the slowest warp
    }
    __syncthreads();
    GMRM_TICK()
    // (hybrid plan: a last pass with a direct row keeps the shared-memory sums, written out here)
    const bool wrote_direct = last_pass >= 0 && !(kDirectMax >= 1 && T == 1 && p.ndir > 0 && pr.count[last_pass] >= 2);
    if (!chunked && !wrote_direct) {
        for (int i = tid; i < p.V * T; i += NT) {
            const int v = i / T, t = i - v * T;
            p.partial[((int64_t)v * p.Ttot + p.t0 + t) * nsm + cta] = part[i];
        }
    } else if (chunked && !flushed) {                                // a CTA that owns no rows at all (more CTAs than rows): its slots are zero
        for (int i = tid; i < p.V * T; i += NT) {
            const int v = i / T, t = i - v * T;
            p.partial[((int64_t)v * p.Ttot + p.t0 + t) * nsm + cta] = 0.0;
        }
    }
#pragma unroll
    for (int t = 0; t < T; t++) {
        const double tot = block_sum_fixed(es[t], red);
        if (tid == 0) p.spart[(int64_t)(p.t0 + t) * nsm + cta] = tot;
    }
    GMRM_TICK()                                           // epilogue: partial / spart writes
    if (profme) atomicAdd(&p.prof[7], 1ull);
#undef GMRM_TICK
}
// [step-end]

// [direct-begin]
// Second copy of the DIRECT rows (hybrid plan of the step kernel, StepParams::bed2): the row that closes the range of (pass,
// cta) in plain 2-bit dosage fields (missing = 0, as in the base-3 copy), [marker][npass * nsm slots][64 bytes].
__global__ void direct_plane_kernel(const uint8_t* __restrict__ plink, int nmark, Layout L, int npass, uint8_t* __restrict__ dst) {
    const int slot = blockIdx.x, q = slot / L.nsm, c = slot - q * L.nsm;
    __shared__ int srow;
    if (threadIdx.x == 0) {
        int start[kMaxPasses], count[kMaxPasses];
        all_pass_rows(L.nrows, npass, L.nsm, c, start, count);
        srow = count[q] >= 2 ? start[q] + count[q] - 1 : -1;
    }
    __syncthreads();
    const int row = srow;
    if (row < 0) return;                                       // a pass of fewer than two rows has no direct row
    const int64_t drows = (int64_t)npass * L.nsm;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < (int64_t)nmark * kRowBytes; i += (int64_t)gridDim.y * blockDim.x) {
        const int64_t j = i >> 6, o = (int64_t)row * kRowBytes + (i & 63);
        uint8_t v = 0;
        if (o < L.mbytes) {
            uint32_t mm;
            v = (uint8_t)tri_to_fields(plink_to_tri(plink[j * L.mbytes + o], &mm));
        }
        dst[(j * drows + slot) * kRowBytes + (i & 63)] = v;
    }
}
// [direct-end]

// =====================================================================================
// K2: one warp per virtual rank: finish the dot product, sample, publish.
// =====================================================================================
// [sample-begin]  (tests/test_marker_loop_emulated.py compiles the text up to [sample-end] for the host, see tests/emu/)
struct DotPieces { double dpa, dpb; };

// sum a*eps and sum b*eps of trait t for the marker of virtual rank r (all lanes get the result)
__device__ __forceinline__ DotPieces finish_dot(const SampleParams& p, int r, int col, int t, int lane) {
    const double* part = p.partial + ((int64_t)r * p.T + t) * p.nsm;
    double s = 0.0;
    for (int i0 = lane; i0 < p.nsm; i0 += 32 * 8) {          // 8 independent loads per round, fixed summation order
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = i0 + 32 * j < p.nsm ? part[i0 + 32 * j] : 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) s += x[j];
    }
    const double coded = warp_sum_fixed(s);                 // sum a*eps (missing genotypes are stored as dosage 0)
    double sa = 0.0;
    for (int i0 = lane; i0 < p.nsm; i0 += 32 * 8) {
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = i0 + 32 * j < p.nsm ? p.spart[(int64_t)t * p.nsm + i0 + 32 * j] : 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) sa += x[j];
    }
    const double sall = warp_sum_fixed(sa);                 // sum eps over all slots
    double sm = 0.0;
    const uint32_t m0 = p.miss_off[col], m1 = p.miss_off[col + 1];
    for (uint32_t i = m0 + lane; i < m1; i += 32) sm += p.eps[(int64_t)t * p.npad + p.miss_idx[i]];
    const double smiss = m1 > m0 ? warp_sum_fixed(sm) : 0.0;
    DotPieces d;
    d.dpa = coded;                   // a = 0 at missing (lut_a)
    d.dpb = sall - smiss;            // b = 0 at missing (lut_b)
    return d;
}
// Per-(trait, group) pieces of the sampler that do not depend on the marker (bayes.cpp:403-405,413-416,
// 429-431), evaluated once per iteration instead of once per marker:
//   gc[0..K)   denom[k-1] = (N-1) + sige_g * cvai[k]            (k >= 1; slot 0 holds inv2sige)
//   gc[K..2K)  log(pi[k])
//   gc[2K..3K) -0.5 * log(sigg_e * (nonas-1) * cva[k] + 1)        (k >= 1)
//   gc[3K..4K) sqrt(sigmae / denom[k-1])                          (k >= 1; sd of the beta draw, bayes.cpp:456)
__global__ void group_consts_kernel(int T, int G, int K, int N, const double* __restrict__ sigmag, const double* __restrict__ sigmae,
                                    const double* __restrict__ pi, const double* __restrict__ cva, const double* __restrict__ cvai,
                                    const int32_t* __restrict__ nonas, double* __restrict__ gc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * G) return;
    const int t = i / G, g = i % G;
    double* o = gc + (int64_t)i * 4 * K;
    const double sigg = sigmag[i], sige = sigmae[t];
    if (sigg == 0.0) { for (int k = 0; k < 4 * K; k++) o[k] = 0.0; return; }
    const double sige_g = sige / sigg;                       // 403
    const double sigg_e = 1.0 / sige_g;                      // 404
    o[0] = 1.0 / (2.0 * sige);                               // 405 inv2sige
    for (int k = 0; k < K; k++) {
        o[K + k] = log(pi[(int64_t)i * K + k]);              // 429
        if (k > 0) {
            const double denom = (double)(N - 1) + sige_g * cvai[g * K + k];                       // 414
            o[k] = denom;
            o[2 * K + k] = -0.5 * log(sigg_e * (double)(nonas[t] - 1) * cva[g * K + k] + 1.0);     // 431
            o[3 * K + k] = sqrt(sige / denom);
        }
    }
    o[2 * K] = 0.0; o[3 * K] = 0.0;
}

// Latency-ordered: the kernel is one dependent chain per warp, so independent loads are issued together --
// round 1: column index, the marker's partial sums and the residual sums (none depends on the column);
// round 2 (needs the column): missing-list bounds, group, mave, msig, beta;  round 3: sigmaG, sampler constants.
// The published item of lane t < T (trait t) is returned in (lam, mave); lam == 0: nothing to publish.
#ifndef GMRM_SAMPLE_EARLY
#define GMRM_SAMPLE_EARLY 1
#endif
__device__ __forceinline__ void sample_one(const SampleParams& p, int v, int lane, int col, double& out_lam, double& out_mave) {
    out_lam = 0.0; out_mave = 0.0;
#if !GMRM_SAMPLE_EARLY
    pdl_wait();
#endif
    // ---- before the grid dependency: everything about the marker that the step kernel still running does not write --
    // missing-list bounds, group, mave, msig, beta, the group's variance and sampler constants.  Their (often HBM) round
    // trips then run under the tail of the step kernel instead of behind the partial sums.
    const int cc = max(col, 0);
    const uint32_t m0 = p.miss_off[cc], m1 = p.miss_off[cc + 1];
    const int grp = p.group[cc];
    const int tl = lane < p.T ? lane : 0;
    const int64_t mi = (int64_t)tl * p.Mloc + cc;
    const double mave = p.mave[mi], msig = p.msig[mi], beta_old = p.betas[mi];
    const int64_t ri = ((int64_t)p.step * p.R + (p.r0 + v)) * p.T + tl;
    const double rep_u = p.rep_u ? p.rep_u[ri] : 0.0;
    const int nonas = p.nonas[tl];
    const double sigg = p.sigmag[tl * p.G + grp];
    const double* gc = p.gc + ((int64_t)tl * p.G + grp) * 4 * p.K;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(gc + (lane & 15)));
#if GMRM_SAMPLE_EARLY
    pdl_wait();                                              // the step kernel's partial sums are complete and visible
#endif
    // ---- sum a*eps partials and sum eps, trait by trait (fixed order)
    double my_coded = 0.0, my_sall = 0.0;
    for (int t = 0; t < p.T; t++) {
        const double* part = p.partial + ((int64_t)v * p.T + t) * p.nsm;
        double x[8], y[8], s = 0.0, sa = 0.0;
        for (int i0 = lane; i0 < p.nsm; i0 += 32 * 8) {      // 8 independent loads per round, fixed summation order
#pragma unroll
            for (int j = 0; j < 8; j++) {
                x[j] = i0 + 32 * j < p.nsm ? part[i0 + 32 * j] : 0.0;
                y[j] = i0 + 32 * j < p.nsm ? p.spart[(int64_t)t * p.nsm + i0 + 32 * j] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) s += x[j];
#pragma unroll
            for (int j = 0; j < 8; j++) sa += y[j];
        }
        const double coded = warp_sum_fixed(s), sall = warp_sum_fixed(sa);
        if (lane == t) { my_coded = coded; my_sall = sall; }
    }
    if (col < 0) return;
    double my_smiss = 0.0;
    if (m1 > m0)                                              // warp-uniform: sum of eps over the missing genotypes
        for (int t = 0; t < p.T; t++) {
            double sm = 0.0;
            for (uint32_t i = m0 + lane; i < m1; i += 32) sm += p.eps[(int64_t)t * p.npad + p.miss_idx[i]];
            const double smiss = warp_sum_fixed(sm);
            if (lane == t) my_smiss = smiss;
        }
    if (lane >= p.T) return;
    const int t = lane;
    const double my_dpa = my_coded;                           // a = 0 at missing (lut_a): stored as dosage 0
    const double my_dpb = my_sall - my_smiss;                 // b = 0 at missing (lut_b)
    const double dot_raw = msig * (my_dpa - mave * my_dpb);                // bayes.cpp:766
    const uint32_t mglo = (uint32_t)(p.marker_begin + col);
    double u;
    if (p.rep_u) {
        u = rep_u;
        if (sigg != 0.0 && !(u == u)) atomicExch(p.err, 1);               // reference drew nothing here
    } else {
        u = draw_uniform(p.seed, STREAM_SAMPLER_U, (uint32_t)p.it, mglo, (uint32_t)t);
    }
    // the normal is only materialised when a non-null component is chosen (bayes.cpp:456)
    auto zdraw = [&]() -> double {
        if (p.rep_z) {
            const double z = p.rep_z[ri];
            if (!(z == z)) atomicExch(p.err, 2);
            return z;
        }
        return draw_normal(p.seed, STREAM_SAMPLER_N, (uint32_t)p.it, mglo, (uint32_t)t);
    };
    const MarkerDraw d = sample_marker_pre(dot_raw, beta_old, sigg, gc, p.K, nonas, u, zdraw);
    p.betas[mi] = d.beta_new;
    if (d.comp >= 0) {
        p.comp[mi] = d.comp;                                               // bayes.cpp:462
        atomicAdd(&p.cass[(t * p.G + grp) * p.K + d.comp], 1);              // 460
    }
    out_lam = fabs(d.dbeta) > 0.0 ? d.dbeta * msig : 0.0;                  // 483-487, phenotype.cpp:328
    out_mave = mave;
}

// One warp per virtual rank, kSegCap virtual ranks per CTA.  The CTA compacts its published items (virtual-rank order) into
// ITS segment of this GPU's list -- and, across GPUs, into the same segment of every peer's copy over NVLink -- so no
// CTA waits for another; the last CTA to finish only raises the flags the peers' next step kernels wait on.
constexpr int kSampleMaxT = 32;
__global__ void __launch_bounds__(kSegCap * 32) sample_kernel(const SampleParams p) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int v = blockIdx.x * kSegCap + warp;
    __shared__ double s_lam[kSampleMaxT][kSegCap], s_mave[kSampleMaxT][kSegCap];
    __shared__ int s_col[kSegCap], s_npub;
    if (tid == 0) s_npub = 0;
    pdl_trigger();                                           // the next step kernel may start its prologue (shared-memory set-up, L2 prefetch)
    const int col = v < p.V ? p.cols[v] : -1;
    double lam = 0.0, mave = 0.0;
    if (v < p.V) sample_one(p, v, lane, col, lam, mave);     // waits for the step kernel (pdl_wait) after its marker-only loads
    else pdl_wait();
    if (p.pf_bed != nullptr && __ballot_sync(0xffffffffu, lam != 0.0) != 0u && lane == 0 && col >= 0)   // published: its column is needed again in a few microseconds
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.pf_bed + (int64_t)col * p.pf_col_stride), "r"((uint32_t)p.pf_col_stride) : "memory");
    if (lane < p.T) { s_lam[lane][warp] = lam; s_mave[lane][warp] = mave; }
    if (lane == 0) s_col[warp] = col;
    __syncthreads();
    const bool push = p.world > 1 && p.peer_list[0] != nullptr;
    const size_t S = publist_segments(p.V);
    for (int t = warp; t < p.T; t += kSegCap) {              // warp w compacts the traits w, w + 16, ...
        const bool live = lane < kSegCap && s_lam[t][lane < kSegCap ? lane : 0] != 0.0;
        const uint32_t bal = __ballot_sync(0xffffffffu, live);
        const int pos = __popc(bal & ((1u << lane) - 1u)), n = __popc(bal);
        const size_t off = ((size_t)t * S + blockIdx.x) * kSegDoubles;
        PubItem it;
        if (live) { it.lam = s_lam[t][lane]; it.mave = s_mave[t][lane]; it.col = s_col[lane]; it.v = blockIdx.x * kSegCap + lane; }
        for (int g = 0; g < (push ? p.world : 1); g++) {     // every GPU's copy of this list (own one included)
            double* seg = (push ? p.peer_list[g] : p.plist) + off;
            if (live) reinterpret_cast<PubItem*>(seg + 2)[pos] = it;
        }
        if (push) __threadfence_system();                    // one MEMBAR for the warp: the items first ...
        __syncwarp();
        if (lane == 0) {                                     // ... the header last: count | sequence
            for (int g = 0; g < (push ? p.world : 1); g++)
                *reinterpret_cast<volatile unsigned long long*>((push ? p.peer_list[g] : p.plist) + off) = seg_header(n, p.seq);
            if (n) atomicAdd(&s_npub, n);
        }
    }
    __syncthreads();
    if (tid == 0 && s_npub) atomicAdd(reinterpret_cast<unsigned long long*>(p.npublished), (unsigned long long)s_npub);
}
// [sample-end]

// test hook behind gmrm_dot_products: out[v*T+t] = Bayes::dot_product
__global__ void __launch_bounds__(128) finish_dots_kernel(const SampleParams p, double* __restrict__ out) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= p.V) return;
    const int col = p.cols[v];
    for (int t = 0; t < p.T; t++) {
        const DotPieces d = finish_dot(p, v, col, t, lane);
        const int64_t mi = (int64_t)t * p.Mloc + col;
        if (lane == 0) out[(int64_t)v * p.T + t] = p.msig[mi] * (d.dpa - p.mave[mi] * d.dpb);
    }
}
// =====================================================================================
// per-iteration prologue / epilogue
// =====================================================================================
// [epilogue-begin]
// marker of (step s, local virtual rank v): Bayes::set_block_of_markers (bayes.cpp:903-925) + midx
__global__ void steptab_kernel(int32_t* __restrict__ tab, int Mm, int Vl, int r0, int R, int Mt, int marker_begin,
                               int shuffle, uint32_t seed, int it, const int32_t* __restrict__ rep_perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Mm * Vl) return;
    const int s = (int)(i / Vl), v = (int)(i % Vl), r = r0 + v;
    const int size = Mt / R, modu = Mt % R;
    const int Mr = size + (r < modu ? 1 : 0);
    const int Sr = r * size + (r < modu ? r : modu);
    int out = -1;
    if (s < Mr) {
        int loc = s;
        if (shuffle) loc = rep_perm ? rep_perm[(int64_t)r * Mm + s] : (int)perm_at((uint32_t)s, (uint32_t)Mr, seed, (uint32_t)it, (uint32_t)r);
        out = Sr - marker_begin + loc;
    }
    tab[i] = out;
}

// sum of beta^2 per group over the shard (bayes.cpp:566-569); block = trait, fixed order
__global__ void __launch_bounds__(256) beta_sq_kernel(const double* __restrict__ betas, const int32_t* __restrict__ group,
                                                      int Mloc, int G, double* __restrict__ out) {
    extern __shared__ double acc[];                  // [G][256]
    const int t = blockIdx.x, tid = threadIdx.x;
    for (int g = 0; g < G; g++) acc[g * 256 + tid] = 0.0;
    const int chunk = (Mloc + 255) / 256;
    const int j0 = tid * chunk, j1 = min(Mloc, j0 + chunk);
    for (int j = j0; j < j1; j++) {
        const double b = betas[(int64_t)t * Mloc + j];
        acc[group[j] * 256 + tid] += b * b;
    }
    __syncthreads();
    for (int g = 0; g < G; g++) {
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) acc[g * 256 + tid] += acc[g * 256 + tid + o];
            __syncthreads();
        }
        if (tid == 0) out[t * G + g] = acc[g * 256];
    }
}

// The same sum for large shards in two levels (one CTA per trait took ~1 ms per iteration at 10^6 markers): CTA (t, b) sums
// slice b of the markers per group, a second launch adds the slices in order.  Fixed slicing: reproducible.
constexpr int kBsqSlices = 128;
__global__ void __launch_bounds__(256) beta_sq_part_kernel(const double* __restrict__ betas, const int32_t* __restrict__ group,
                                                           int Mloc, int G, double* __restrict__ part /* [T][kBsqSlices][G] */) {
    extern __shared__ double acc[];                  // [G][256]
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    for (int g = 0; g < G; g++) acc[g * 256 + tid] = 0.0;
    const int per = (Mloc + kBsqSlices - 1) / kBsqSlices, j0 = b * per, j1 = min(Mloc, j0 + per);
    for (int j = j0 + tid; j < j1; j += 256) {
        const double x = betas[(int64_t)t * Mloc + j];
        acc[group[j] * 256 + tid] += x * x;
    }
    __syncthreads();
    for (int g = 0; g < G; g++) {
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) acc[g * 256 + tid] += acc[g * 256 + tid + o];
            __syncthreads();
        }
        if (tid == 0) part[((int64_t)t * kBsqSlices + b) * G + g] = acc[g * 256];
    }
}
__global__ void beta_sq_final_kernel(const double* __restrict__ part, int T, int G, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * G) return;
    const int t = i / G, g = i - t * G;
    double s = 0.0;
    for (int b = 0; b < kBsqSlices; b++) s += part[((int64_t)t * kBsqSlices + b) * G + g];
    out[i] = s;
}

// sigmaE start value, Phenotype::update_epsilon_sigma (phenotype.cpp:448-457)
__global__ void init_sigmae_kernel(const double* esq, const int32_t* nonas, int T, double* sigmae) {
    const int t = threadIdx.x;
    if (t < T) sigmae[t] = esq[t] / (double)nonas[t] * 0.5;
}

// new intercept (bayes.cpp:357, phenotype.cpp:279-282: mean epssum/nonas with epssum == 0)
__global__ void mu_draw_kernel(const MuDrawParams p) {
    const int t = threadIdx.x;
    if (t >= p.T) return;
    p.mu_old[t] = p.mu[t];
    double m;
    if (p.rep_mu) m = p.rep_mu[t];
    else m = 0.0 / (double)p.nonas[t] + sqrt(p.sigmae[t] / (double)p.nonas[t]) * draw_normal(p.seed, STREAM_MU, (uint32_t)p.it, 0u, (uint32_t)t);
    p.mu[t] = m;
}

// group variances, mixture proportions, residual variance (bayes.cpp:594-650); thread = trait
__global__ void global_draw_kernel(const GlobalDrawParams p) {
    const int t = threadIdx.x;
    if (t >= p.T) return;
    const double V0E = 0.0001, S02E = 0.0001, V0G = 0.0001, S02G = 0.0001;   // bayes.hpp:14-17
    for (int g = 0; g < p.G; g++) {
        p.m0[t * p.G + g] = 0;                                                // reset_m0, bayes.cpp:366
        if (p.mtotgrp[g] == 0) continue;                                      // 597-598
        const int32_t* cs = p.cass + (t * p.G + g) * p.K;
        const int m0 = p.mtotgrp[g] - cs[0];                                  // 605
        p.m0[t * p.G + g] = m0;
        int csum = 0;
        for (int k = 0; k < p.K; k++) csum += cs[k];
        if (m0 == 0 || csum == 0) { p.sigmag[t * p.G + g] = 0.0; continue; }  // 608-611
        const double a = V0G + (double)m0;
        const double b = (p.bsq[t * p.G + g] * (double)m0 + V0G * S02G) / (V0G + (double)m0);   // 613
        double unit;
        if (p.rep_sigg_unit) { unit = p.rep_sigg_unit[t * p.G + g]; if (!(unit == unit)) atomicExch(p.err, 3); }
        else unit = draw_gamma(0.5 * a, p.seed, STREAM_SIGMAG, (uint32_t)p.it, (uint32_t)g, (uint32_t)t);
        p.sigmag[t * p.G + g] = inv_scaled_chisq_from_unit(a, b, unit);
        double* pi = p.pi + ((int64_t)t * p.G + g) * p.K;                     // phenotype.cpp:227-237
        double sum = 0.0;
        for (int k = 0; k < p.K; k++) {
            double val;
            if (p.rep_pi_unit) { val = p.rep_pi_unit[(t * p.G + g) * p.K + k]; if (!(val == val)) atomicExch(p.err, 4); }
            else val = draw_gamma((double)cs[k] + 1.0, p.seed, STREAM_PI, (uint32_t)p.it, (uint32_t)(g * p.K + k), (uint32_t)t);
            pi[k] = val;
            sum += val;
        }
        for (int k = 0; k < p.K; k++) pi[k] /= sum;
    }
    const double a = V0E + (double)p.N, b = (p.esq[t] + V0E * S02E) / (V0E + (double)p.N);   // 635
    double unit;
    if (p.rep_sige_unit) unit = p.rep_sige_unit[t];
    else unit = draw_gamma(0.5 * a, p.seed, STREAM_SIGMAE, (uint32_t)p.it, 0u, (uint32_t)t);
    p.sigmae[t] = inv_scaled_chisq_from_unit(a, b, unit);
}
// [epilogue-end]

// =====================================================================================
// launchers
// =====================================================================================
void launch_transcode(const uint8_t* plink, int nmark, const Layout& L, uint8_t* dst, uint32_t* miss_counts, cudaStream_t s) {
    if (nmark <= 0) return;
    dim3 grid((unsigned)((L.col_stride + 255) / 256), (unsigned)nmark);
    transcode_kernel<<<grid, 256, 0, s>>>(plink, nmark, L, dst, miss_counts);
}
void launch_direct_plane(const uint8_t* plink, int nmark, const Layout& L, int npass, uint8_t* dst, cudaStream_t s) {
    if (nmark <= 0 || npass <= 0) return;
    dim3 grid((unsigned)(npass * L.nsm), (unsigned)std::min(64, (nmark * kRowBytes + 255) / 256));
    direct_plane_kernel<<<grid, 256, 0, s>>>(plink, nmark, L, npass, dst);
}
void launch_fill_missing(const uint8_t* plink, int nmark, const Layout& L, const uint32_t* off, uint32_t* idx, cudaStream_t s) {
    if (nmark <= 0) return;
    fill_missing_kernel<<<(nmark + 3) / 4, 128, 0, s>>>(plink, nmark, L, off, idx);
}
void launch_untranscode(const uint8_t* bed, int nmark, const Layout& L, const uint32_t* miss_off, const uint32_t* miss_idx,
                        uint8_t* plink_out, cudaStream_t s) {
    if (nmark <= 0) return;
    dim3 grid((unsigned)((L.mbytes + 255) / 256), (unsigned)nmark);
    untranscode_kernel<<<grid, 256, 0, s>>>(bed, nmark, L, plink_out);
    unmiss_kernel<<<(nmark + 3) / 4, 128, 0, s>>>(nmark, L, miss_off, miss_idx, plink_out);
}
void launch_decode_column(const uint8_t* col, const Layout& L, const uint32_t* miss_idx, uint32_t nmiss, double* a, double* b, cudaStream_t s) {
    decode_column_kernel<<<(unsigned)((L.N + 255) / 256), 256, 0, s>>>(col, L, a, b);
    if (nmiss) decode_missing_kernel<<<(nmiss + 255) / 256, 256, 0, s>>>(miss_idx, nmiss, L.N, a, b);
}
void launch_decode_namask(const uint8_t* mask4, const Layout& L, double* na, cudaStream_t s) {
    decode_namask_kernel<<<(unsigned)((L.N + 255) / 256), 256, 0, s>>>(mask4, L, na);
}
void launch_generate_plink(uint8_t* dst, int nmark, int first_global_marker, const Layout& L, uint32_t seed, double maf_lo,
                           double maf_hi, double missing_rate, cudaStream_t s) {
    if (nmark <= 0) return;
    dim3 grid((unsigned)((L.mbytes + 255) / 256), (unsigned)nmark);
    generate_plink_kernel<<<grid, 256, 0, s>>>(dst, nmark, first_global_marker, L, seed, maf_lo, maf_hi, missing_rate);
}
void launch_stats(const uint8_t* bed, int nmark, const Layout& L, const uint8_t* mask4, const uint32_t* miss_off, const uint32_t* miss_idx,
                  const int32_t* nonas, const uint32_t* na_off, const uint32_t* na_idx, int T, double* mave, double* msig, cudaStream_t s,
                  double* xtx) {
    if (nmark <= 0) return;
    const int ctas = nmark < 8 * L.nsm ? nmark : 8 * L.nsm;          // persistent: 8 CTAs of 256 threads per SM
    stats_kernel<<<ctas, kStatsThreads, 0, s>>>(bed, nmark, L, mask4, miss_off, miss_idx, nonas, na_off, na_idx, T, mave, msig, xtx);
}
void launch_eps_offset(double* eps, const uint8_t* mask4, const Layout& L, int T, const double* mu_old, const double* mu_new, cudaStream_t s) {
    dim3 grid((unsigned)((L.npad + 255) / 256), (unsigned)T);
    eps_offset_kernel<<<grid, 256, 0, s>>>(eps, mask4, L, mu_old, mu_new);
}
void launch_eps_merge(double* eps, double* loc, const double* tot, const Layout& L, int T, cudaStream_t s) {
    const int64_t n = (int64_t)T * L.npad;
    eps_merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(eps, loc, tot, n);
}
__global__ void fill_u64_kernel(unsigned long long* dst, size_t n, unsigned long long value) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = value;
}
void launch_fill_u64(unsigned long long* dst, size_t n, unsigned long long value, cudaStream_t s) {
    if (n) fill_u64_kernel<<<1184, 256, 0, s>>>(dst, n, value);
}
void launch_eps_sumsq(const double* eps, int64_t npad, int64_t n, int T, double* out, cudaStream_t s) {
    eps_sumsq_kernel<<<T, 1024, 0, s>>>(eps, npad, n, out);
}

void host_pass_rows(int nrows, int npass, int nsm, int cta, int* start, int* count) { all_pass_rows(nrows, npass, nsm, cta, start, count); }
int host_chunk_offset(int nr, int slot, int lane16, int k) { return chunk_offset(nr, slot, lane16, k); }

constexpr int kMaxDynSmem = 232448;   // 227 KB: the most one CTA can opt into on sm_100

// passes of a step: the fewest such that a CTA never gets more than rows_per_pass rows in one
int step_npass(const Layout& L, int rows_per_pass) {
    const int64_t cap = (int64_t)rows_per_pass * L.nsm;
    int np = (int)((L.nrows + cap - 1) / cap);
    if (np < 1) np = 1;
    // even split of ceil(nrows/np) rows over nsm CTAs must fit rows_per_pass
    while (((L.nrows + np - 1) / np + L.nsm - 1) / L.nsm > rows_per_pass) np++;
    return np;
}
int step_smem_bytes(const Layout& L, int V, int T, int rows_per_pass) {
    if (step_npass(L, rows_per_pass) > kMaxPasses) return -1;
    const int nrmax = step_npass(L, rows_per_pass) * rows_per_pass;   // bound on the rows one CTA owns
    const int64_t bytes = (int64_t)kTabBase + (int64_t)step_area_bytes(V, T, rows_per_pass, step_npass(L, rows_per_pass)) + (int64_t)kPubCap * 32 +
                          (int64_t)(V < kPartSmemDoubles / T ? V : kPartSmemDoubles / T) * T * 8 + (int64_t)kPubCap * 8 + 82 * 4 + (int64_t)sizeof(PassRows) + (int64_t)((nrmax * 8 + 1) & ~1) * 4 + 32 * 8 + 4 * 32 * 4;
    return bytes <= kMaxDynSmem ? (int)bytes : -1;
}

void step_plan(const Layout& L, int V, int Ttot, int* traits_per_launch, int* rows_per_pass) {
    *traits_per_launch = 0; *rows_per_pass = 0;
    for (int tc = Ttot < 4 ? Ttot : 4; tc >= 1; tc--) {   // step_kernel<T> is instantiated for T = 1..4
        int rpp = kMaxSlots / tc;
        while (rpp >= 1 && step_smem_bytes(L, V, tc, rpp) < 0) rpp--;
        if (rpp < 1) continue;
        const int nchunks = (Ttot + tc - 1) / tc;
        const int bal = (Ttot + nchunks - 1) / nchunks;        // balanced chunks: 6 traits -> 3 + 3, not 5 + 1
        int rb = kMaxSlots / bal;
        while (rb >= 1 && step_smem_bytes(L, V, bal, rb) < 0) rb--;
        if (rb < 1) { *traits_per_launch = tc; *rows_per_pass = rpp; }
        else { *traits_per_launch = bal; *rows_per_pass = rb; }
        const int nrmax = L.max_rows_per_cta();
        if (*rows_per_pass > nrmax && nrmax >= 1) *rows_per_pass = nrmax;
        return;
    }
}

// cudaFuncSetAttribute acts on the current device only, and the CLI drives several GPUs from one process (one host
// thread each): the opt-in to more than 48 KB of dynamic shared memory is tracked per device.
constexpr int kMaxDevices = 64;
template <typename Kernel>
static bool ensure_dyn_smem(Kernel kernel, std::atomic<int>* have, int bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
    if (have[dev].load(std::memory_order_acquire) >= bytes) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
    have[dev].store(bytes, std::memory_order_release);
    return true;
}

template <int T>
static int step_launch_t(const Layout& L, const StepParams& p, cudaStream_t s) {
    const int smem = step_smem_bytes(L, p.V, T, p.rows_per_pass);
    if (smem < 0) return -3;
    static std::atomic<int> have[kMaxDevices];
    if (!ensure_dyn_smem(step_kernel<T>, have, kMaxDynSmem)) return -1;
    // every launch asks for the full 227 KB, update-only ones included: a different dynamic size would make the
    // driver re-partition L1/shared memory between consecutive launches of the marker loop
    (void)smem;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)L.nsm); cfg.blockDim = dim3(kStepThreads); cfg.dynamicSmemBytes = kMaxDynSmem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = p.pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, step_kernel<T>, p) == cudaSuccess ? 0 : -2;
}
// T = traits of this launch (1..4); p.rows_per_pass * T <= kMaxSlots
int launch_step(const Layout& L, int T, const StepParams& p, cudaStream_t s) {
    if (p.V > 0 && (p.rows_per_pass < 1 || p.rows_per_pass * T > kMaxSlots)) return -4;
    switch (T) {
    case 1: return step_launch_t<1>(L, p, s);
    case 2: return step_launch_t<2>(L, p, s);
    case 3: return step_launch_t<3>(L, p, s);
    case 4: return step_launch_t<4>(L, p, s);
    }
    return -5;
}

void launch_sample(const SampleParams& p, cudaStream_t s) {
    if (p.V <= 0) return;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)publist_segments(p.V)); cfg.blockDim = dim3(kSegCap * 32); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = p.pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, sample_kernel, p);
}
void launch_finish_dots(const SampleParams& p, double* out, cudaStream_t s) {
    if (p.V <= 0) return;
    finish_dots_kernel<<<(p.V + 3) / 4, 128, 0, s>>>(p, out);
}

void launch_steptab(int32_t* tab, int Mm, int Vl, int r0, int R, int Mt, int marker_begin, int shuffle, uint32_t seed,
                    int it, const int32_t* rep_perm, cudaStream_t s) {
    const int64_t n = (int64_t)Mm * Vl;
    if (n <= 0) return;
    steptab_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tab, Mm, Vl, r0, R, Mt, marker_begin, shuffle, seed, it, rep_perm);
}
int beta_sq_scratch_doubles(int T, int G) { return T * kBsqSlices * G; }
void launch_beta_sq(const double* betas, const int32_t* group, int Mloc, int T, int G, double* out, double* scratch, cudaStream_t s) {
    const int smem = G * 256 * (int)sizeof(double);
    if (scratch != nullptr && Mloc >= 4096) {                 // two levels for shards of any size worth it
        static std::atomic<int> have2[kMaxDevices];
        if (smem > 48 * 1024) ensure_dyn_smem(beta_sq_part_kernel, have2, smem);
        beta_sq_part_kernel<<<dim3((unsigned)T, kBsqSlices), 256, smem, s>>>(betas, group, Mloc, G, scratch);
        beta_sq_final_kernel<<<(T * G + 127) / 128, 128, 0, s>>>(scratch, T, G, out);
        return;
    }
    static std::atomic<int> have[kMaxDevices];
    if (smem > 48 * 1024) ensure_dyn_smem(beta_sq_kernel, have, smem);
    beta_sq_kernel<<<T, 256, smem, s>>>(betas, group, Mloc, G, out);
}
void launch_group_consts(int T, int G, int K, int N, const double* sigmag, const double* sigmae, const double* pi, const double* cva,
                         const double* cvai, const int32_t* nonas, double* gc, cudaStream_t s) {
    group_consts_kernel<<<(T * G + 127) / 128, 128, 0, s>>>(T, G, K, N, sigmag, sigmae, pi, cva, cvai, nonas, gc);
}
void launch_global_draw(const GlobalDrawParams& p, cudaStream_t s) { global_draw_kernel<<<1, 32, 0, s>>>(p); }
void launch_mu_draw(const MuDrawParams& p, cudaStream_t s) { mu_draw_kernel<<<1, 32, 0, s>>>(p); }
void launch_init_sigmae(const double* esq, const int32_t* nonas, int T, double* sigmae, cudaStream_t s) {
    init_sigmae_kernel<<<1, 32, 0, s>>>(esq, nonas, T, sigmae);
}

}  // namespace gmrm
