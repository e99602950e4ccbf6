// sm_100a kernels of the Gibbs marker loop (see kernels.cuh, layout.h).
#include "kernels.cuh"

#include <cstdio>

namespace gmrm {

// =====================================================================================
// small device helpers
// =====================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Replace the LOW word of a 64-bit register pair, keeping the high word where it is.  With the
// high word 0 the pair reads as the denormal double x * 2^-1074.  Writing it this way (and not as
// a fresh {x, 0} pack) is what lets ptxas keep one persistent zero register per multiplier, so
// that a genotype costs exactly one shift + one DFMA (checked with cuobjdump).
__device__ __forceinline__ void set_lo(double& D, uint32_t x) {
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmov.b64 %0, {%1, hi};\n\t}" : "+d"(D) : "r"(x));
}

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

__device__ __forceinline__ void tile_offset_to_slot_byte(int E4, int off, int& ls, int& b) {
    const int nw = E4 / 4, wbytes = nw * kLanesPerTile * 4;
    if (off < wbytes) {
        const int wi = off / (kLanesPerTile * 4), rem = off % (kLanesPerTile * 4);
        ls = rem / 4; b = wi * 4 + rem % 4;
        return;
    }
    off -= wbytes;
    int bb = 4 * nw;
    if (E4 & 2) {
        if (off < kLanesPerTile * 2) { ls = off / 2; b = bb + off % 2; return; }
        off -= kLanesPerTile * 2;
        bb += 2;
    }
    ls = off; b = bb;
}

// The groups (registers) of one lane-slot of one tile.
template <int E4>
struct SlotRegs {
    static constexpr int NW = E4 / 4, NH = (E4 % 4) / 2, NB = E4 % 2;
    uint32_t w[NW > 0 ? NW : 1];
    uint32_t h, b;
    __device__ __forceinline__ void load(const uint8_t* tile, int ls) {   // generic / shared / global pointer
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = *reinterpret_cast<const uint32_t*>(tile + i * kLanesPerTile * 4 + ls * 4);
        h = 0; b = 0;
        if (NH) h = *reinterpret_cast<const uint16_t*>(tile + NW * kLanesPerTile * 4 + ls * 2);
        if (NB) b = *(tile + NW * kLanesPerTile * 4 + NH * kLanesPerTile * 2 + ls);
    }
    // 2-bit field of individual k of the slot
    __device__ __forceinline__ uint32_t field(int k) const {
        if (k < 16 * NW) return (w[k / 16] >> (2 * (k % 16))) & 3u;
        k -= 16 * NW;
        if (NH) { if (k < 8) return (h >> (2 * k)) & 3u; k -= 8; }
        return (b >> (2 * k)) & 3u;
    }
};

// =====================================================================================
// .bed ingestion: transcode PLINK bytes <-> tile-planar dosage bytes (bit-exact, invertible)
// =====================================================================================
__global__ void transcode_kernel(const uint8_t* __restrict__ src, int nmark, Layout L, uint8_t* __restrict__ dst) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (o >= L.col_stride || j >= nmark) return;
    const int c = (int)(o / L.tile_bytes), off = (int)(o % L.tile_bytes);
    int ls, b;
    tile_offset_to_slot_byte(L.E4, off, ls, b);
    const int64_t idx = ((int64_t)c * kLanesPerTile + ls) * L.E4 + b;
    uint8_t v = 0;
    if (idx < L.mbytes) v = plink_to_dosage(src[(int64_t)j * L.mbytes + idx]);
    dst[(int64_t)j * L.col_stride + o] = v;
}

__global__ void untranscode_kernel(const uint8_t* __restrict__ tiles, int nmark, Layout L, uint8_t* __restrict__ dst) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (idx >= L.mbytes || j >= nmark) return;
    const int64_t s = idx / L.E4;
    const int b = (int)(idx % L.E4), c = (int)(s / kLanesPerTile), ls = (int)(s % kLanesPerTile);
    const uint8_t y = tiles[(int64_t)j * L.col_stride + (int64_t)c * L.tile_bytes + tile_byte_offset(L.E4, ls, b)];
    dst[(int64_t)j * L.mbytes + idx] = dosage_to_plink(y);
}

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

// Missing-genotype lists (CSR over shard-local markers), ascending individual index.
// One warp per marker walks the column in PLINK byte order.
__device__ __forceinline__ uint32_t missing_fields(const uint8_t* col, const Layout& L, int64_t idx) {
    const int64_t s = idx / L.E4;
    const int b = (int)(idx % L.E4), c = (int)(s / kLanesPerTile), ls = (int)(s % kLanesPerTile);
    const uint32_t y = col[(int64_t)c * L.tile_bytes + tile_byte_offset(L.E4, ls, b)];
    uint32_t m = y & (y >> 1) & 0x55u;            // field == 3
    const int64_t i0 = idx * 4;
    if (i0 + 3 >= L.N) {                          // drop pad individuals of the last byte
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (i0 + k >= L.N) m &= ~(1u << (2 * k));
    }
    return m;
}

__global__ void count_missing_kernel(const uint8_t* __restrict__ bed, int nmark, Layout L, uint32_t* __restrict__ counts) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= nmark) return;
    const uint8_t* col = bed + (int64_t)j * L.col_stride;
    uint32_t n = 0;
    for (int64_t idx = lane; idx < L.mbytes; idx += 32) n += __popc(missing_fields(col, L, idx));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) counts[j] = n;
}

__global__ void fill_missing_kernel(const uint8_t* __restrict__ bed, int nmark, Layout L, const uint32_t* __restrict__ off,
                                    uint32_t* __restrict__ out) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= nmark) return;
    const uint8_t* col = bed + (int64_t)j * L.col_stride;
    uint32_t base = off[j];
    for (int64_t idx0 = 0; idx0 < L.mbytes; idx0 += 32) {
        const int64_t idx = idx0 + lane;
        const uint32_t m = idx < L.mbytes ? missing_fields(col, L, idx) : 0u;
        const uint32_t n = __popc(m);
        uint32_t incl = n;                        // inclusive warp scan
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

// test hooks: the reference's table values of one column / one NA mask, individual by individual
__global__ void decode_column_kernel(const uint8_t* __restrict__ col, Layout L, double* __restrict__ a, double* __restrict__ b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.N) return;
    const int64_t s = i / L.E;
    const int k = (int)(i % L.E), c = (int)(s / kLanesPerTile), ls = (int)(s % kLanesPerTile);
    const uint32_t y = col[(int64_t)c * L.tile_bytes + tile_byte_offset(L.E4, ls, k / 4)];
    const uint32_t d = (y >> (2 * (k % 4))) & 3u;        // 0,1,2 = dosage, 3 = missing
    if (a) a[i] = d == 3 ? 0.0 : (double)d;              // dotp_lut_a
    if (b) b[i] = d == 3 ? 0.0 : 1.0;                    // dotp_lut_b (for a NA mask tile: na_lut)
}

// =====================================================================================
// marker statistics, PhenMgr::compute_markers_statistics (phenotype.cpp:466-556), from integer
// counts of each dosage under the trait's NA mask: the sums of the
// reference's loops are sums of small integers, hence these counts exactly.
// =====================================================================================
constexpr int kStatsMaxT = 32;
__global__ void __launch_bounds__(128) stats_kernel(const uint8_t* __restrict__ bed, int nmark, Layout L,
                                                    const uint8_t* __restrict__ namask2, const int32_t* __restrict__ nonas,
                                                    int T, double* __restrict__ mave, double* __restrict__ msig) {
    const int j = blockIdx.x;
    if (j >= nmark) return;
    const uint32_t* col = reinterpret_cast<const uint32_t*>(bed + (int64_t)j * L.col_stride);
    const int nwords = (int)(L.col_stride / 4);
    __shared__ int red[3][4];
    for (int t = 0; t < T; t++) {
        const uint32_t* nm = reinterpret_cast<const uint32_t*>(namask2 + (int64_t)t * L.col_stride);
        int n0 = 0, n1 = 0, n2 = 0;
        for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
            const uint32_t w = col[i], m = nm[i];
            const uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
            n1 += __popc(lo & ~hi & m);
            n2 += __popc(hi & ~lo & m);
            n0 += __popc(~(lo | hi) & 0x55555555u & m);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n0 += __shfl_xor_sync(0xffffffffu, n0, o);
            n1 += __shfl_xor_sync(0xffffffffu, n1, o);
            n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = n0; red[1][threadIdx.x >> 5] = n1; red[2][threadIdx.x >> 5] = n2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double c0 = red[0][0] + red[0][1] + red[0][2] + red[0][3];
            const double c1 = red[1][0] + red[1][1] + red[1][2] + red[1][3];
            const double c2 = red[2][0] + red[2][1] + red[2][2] + red[2][3];
            const double suma = c1 + 2.0 * c2, sumb = c0 + c1 + c2;       // phenotype.cpp:536-537
            const double av = suma / sumb;                                // 540
            const double sumsqr = c0 * (0.0 - av) * (0.0 - av) + c1 * (1.0 - av) * (1.0 - av) + c2 * (2.0 - av) * (2.0 - av);  // 544-545
            mave[(int64_t)t * nmark + j] = av;
            msig[(int64_t)t * nmark + j] = 1.0 / sqrt(sumsqr / ((double)nonas[t] - 1.0));   // 548
        }
    }
}

// =====================================================================================
// residual helpers
// =====================================================================================
// Phenotype::offset_epsilon twice (bayes.cpp:351,359): eps += mu_old*na; eps -= mu_new*na; and the
// per-tile sum of eps that the sampler turns into sum b*eps.
// A tile is the contiguous range of individuals [tile*128*E, (tile+1)*128*E): one block walks it with
// coalesced accesses; na01 is the per-individual 0/1 NA indicator.
__global__ void __launch_bounds__(256) eps_offset_kernel(double* __restrict__ eps, const uint8_t* __restrict__ na01, Layout L,
                                                         const double* __restrict__ mu_old, const double* __restrict__ mu_new,
                                                         double* __restrict__ spart) {
    const int t = blockIdx.y, per = kLanesPerTile * L.E;
    __shared__ double red[8];
    const int64_t base = (int64_t)t * L.npad + (int64_t)blockIdx.x * per;
    const double a = mu_old ? mu_old[t] : 0.0, b = mu_new ? -mu_new[t] : 0.0;
    double s = 0.0;
    for (int i = threadIdx.x; i < per; i += 256) {
        double v = eps[base + i];
        const double m = na01[base + i] ? 1.0 : 0.0;
        v += a * m;            // phenotype.cpp:408
        v += b * m;
        eps[base + i] = v;
        s += v;
    }
    const double tot = block_sum_fixed(s, red);
    if (threadIdx.x == 0) spart[(int64_t)t * L.nsm + blockIdx.x] = tot;
}

// Multi-GPU exchange (replaces the per-marker Allgatherv + local recompute of bayes.cpp:500-547): after the
// all-reduce of the shards' residual deltas, add what the OTHER shards changed, clear the local delta, refresh
// the per-tile sums.
__global__ void __launch_bounds__(256) eps_merge_kernel(double* __restrict__ eps, double* __restrict__ loc,
                                                        const double* __restrict__ tot, Layout L, double* __restrict__ spart) {
    const int t = blockIdx.y, per = kLanesPerTile * L.E;
    __shared__ double red[8];
    const int64_t base = (int64_t)t * L.npad + (int64_t)blockIdx.x * per;
    double s = 0.0;
    for (int i = threadIdx.x; i < per; i += 256) {
        const double v = eps[base + i] + (tot[base + i] - loc[base + i]);
        eps[base + i] = v;
        loc[base + i] = 0.0;
        s += v;
    }
    const double r = block_sum_fixed(s, red);
    if (threadIdx.x == 0) spart[(int64_t)t * L.nsm + blockIdx.x] = r;
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

// =====================================================================================
// K1: streamed decode-and-reduce of the step's V columns against the residuals.
//
// One CTA per tile (SM).  Warp 4*kWPS is the producer: one lane issues one cp.async.bulk per
// (marker, tile) into an 8-stage shared-memory ring, completion on mbarriers.  The other 4*kWPS
// warps are consumers: warp w serves sub-partition w&3 and takes every kWPS-th batch of 8 markers.
// A consumer lane owns E consecutive individuals: their weights w_k stay in registers for the whole
// launch, a marker costs E x (shift + DFMA) per lane (layout.h), then an 8-marker select-free
// transposed butterfly leaves one per-warp partial per marker, written to partial[r][t][tile*4+sp].
// =====================================================================================
template <int E4, int T, int WPS, int BATCH>
__global__ void __launch_bounds__((4 * WPS + 1) * 32, 1) dot_kernel(const DotParams p) {
    constexpr int TILE = kLanesPerTile * E4;
    constexpr int E = 4 * E4;
    constexpr int NW = E4 / 4, NH = (E4 % 4) / 2, NB = E4 % 2;
    constexpr int RING = dot_ring_tiles(WPS);
    constexpr int STAGES = RING / BATCH;
    constexpr int NTHREADS = (4 * WPS + 1) * 32;
    constexpr int LOGB = BATCH == 8 ? 3 : 2;
    static_assert(BATCH == 8 || BATCH == 4, "batch of 4 or 8 markers");
    static_assert(STAGES % WPS == 0, "a stage must always serve the same consumer group");
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + RING * TILE);
    uint64_t* empty = full + STAGES;
    int32_t* scols = reinterpret_cast<int32_t*>(empty + STAGES);     // the step's columns, padded with -1

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = (p.V + BATCH - 1) / BATCH;

    for (int i = threadIdx.x; i < nb * BATCH; i += NTHREADS) scols[i] = i < p.V ? p.cols[i] : -1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == 4 * WPS) {
        if (p.debug == 2) return;                     // debug: compute only, nothing is fed
        // ---------------- producer: lane j < BATCH issues the copy of marker j of every batch ----------------
        const uint8_t* tile0 = p.bed + (int64_t)blockIdx.x * TILE;
        for (int b = 0; b < nb; b++) {
            const int s = b % STAGES;
            if (b >= STAGES) mbar_wait(&empty[s], (uint32_t)((b / STAGES) - 1) & 1u);
            const int col = lane < BATCH ? scols[b * BATCH + lane] : -1;
            const uint32_t nvalid = __popc(__ballot_sync(0xffffffffu, col >= 0));
            if (lane == 0) mbar_expect_tx(&full[s], nvalid * TILE);
            __syncwarp();
            if (col >= 0) bulk_g2s(ring + (s * BATCH + lane) * TILE, tile0 + (int64_t)col * p.col_stride, TILE, &full[s]);
        }
        return;
    }

    // ---------------- consumers ----------------
    const int sp = warp & 3, q = warp >> 2;
    const int ls = sp * 32 + lane;
    const int64_t slot = (int64_t)blockIdx.x * kLanesPerTile + ls;

    // weights of this lane's E individuals, group by group (layout.h: w_k = 2^1000 (eps_k - eps_{k+1}/4))
    double wgt[E][T];
#pragma unroll
    for (int t = 0; t < T; t++) {
        const double* e = p.eps + (int64_t)(p.t0 + t) * p.npad + slot * E;
        double ev[E + 1];
#pragma unroll
        for (int k = 0; k < E; k++) ev[k] = e[k];
        ev[E] = 0.0;
#pragma unroll
        for (int k = 0; k < E; k++) {
            // last genotype of a group has no successor inside the group
            const bool last = (k < 16 * NW) ? ((k % 16) == 15) : (NH && k < 16 * NW + 8) ? (k == 16 * NW + 7) : (k == E - 1);
            wgt[k][t] = group_weight(ev[k], last ? 0.0 : ev[k + 1]);
        }
    }

    // accumulator slot j of this lane holds marker j ^ P (P = lane bits 4,3[,2] reversed): makes the
    // transposed butterfly below select-free
    int P = 0;
#pragma unroll
    for (int i = 0; i < LOGB; i++) P |= ((lane >> (4 - i)) & 1) << (LOGB - 1 - i);
    // shared-memory offsets (from the stage base) of this lane's word / half / byte of accumulator slot j
    uint32_t offw[BATCH], offh[BATCH], offb[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; j++) {
        offw[j] = (uint32_t)((j ^ P) * TILE + ls * 4);
        offh[j] = (uint32_t)((j ^ P) * TILE + NW * kLanesPerTile * 4 + ls * 2);
        offb[j] = (uint32_t)((j ^ P) * TILE + NW * kLanesPerTile * 4 + NH * kLanesPerTile * 2 + ls);
    }
    double D[BATCH];        // one persistent (lo = shifted word, hi = 0) multiplier pair per accumulator
#pragma unroll
    for (int j = 0; j < BATCH; j++) D[j] = p.zeros[j * NTHREADS + threadIdx.x];
    const uint32_t ring_u32 = smem_u32(ring);

    for (int b = q; b < nb; b += WPS) {
        const int s = b % STAGES;
        if (p.debug != 2) mbar_wait(&full[s], (uint32_t)(b / STAGES) & 1u);
        const uint32_t st = ring_u32 + (uint32_t)(s * BATCH * TILE);
        uint32_t gw[BATCH][NW > 0 ? NW : 1], gh[BATCH], gb[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; j++) {
#pragma unroll
            for (int wi = 0; wi < NW; wi++)
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(gw[j][wi]) : "r"(st + offw[j] + wi * kLanesPerTile * 4));
            gh[j] = 0; gb[j] = 0;
            if (NH) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(gh[j]) : "r"(st + offh[j]));
            if (NB) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(gb[j]) : "r"(st + offb[j]));
        }
        __syncwarp();
        if (lane == 0 && p.debug != 2) mbar_arrive(&empty[s]);     // registers hold the batch: the stage can be refilled
        if (p.debug == 1) continue;                   // debug: feed only, no arithmetic

        double acc[BATCH][T];
#pragma unroll
        for (int j = 0; j < BATCH; j++)
#pragma unroll
            for (int t = 0; t < T; t++) acc[j][t] = 0.0;

#pragma unroll
        for (int wi = 0; wi < NW; wi++)
#pragma unroll
            for (int k = 0; k < 16; k++)
#pragma unroll
                for (int j = 0; j < BATCH; j++) {
                    set_lo(D[j], gw[j][wi] << (30 - 2 * k));
#pragma unroll
                    for (int t = 0; t < T; t++) acc[j][t] = fma(D[j], wgt[wi * 16 + k][t], acc[j][t]);
                }
        if (NH) {
#pragma unroll
            for (int k = 0; k < 8; k++)
#pragma unroll
                for (int j = 0; j < BATCH; j++) {
                    set_lo(D[j], gh[j] << (30 - 2 * k));
#pragma unroll
                    for (int t = 0; t < T; t++) acc[j][t] = fma(D[j], wgt[16 * NW + k][t], acc[j][t]);
                }
        }
        if (NB) {
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int j = 0; j < BATCH; j++) {
                    set_lo(D[j], gb[j] << (30 - 2 * k));
#pragma unroll
                    for (int t = 0; t < T; t++) acc[j][t] = fma(D[j], wgt[16 * NW + 8 * NH + k][t], acc[j][t]);
                }
        }

        // transposed butterfly: BATCH markers x 32 lanes -> total of marker P in every lane sharing P
#pragma unroll
        for (int t = 0; t < T; t++) {
            double a[BATCH];
#pragma unroll
            for (int j = 0; j < BATCH; j++) a[j] = acc[j][t];
#pragma unroll
            for (int i = 0; i < LOGB; i++) {
                const int half = BATCH >> (i + 1);
#pragma unroll
                for (int j = 0; j < half; j++) a[j] += __shfl_xor_sync(0xffffffffu, a[j + half], 16 >> i);
            }
#pragma unroll
            for (int o = 16 >> LOGB; o > 0; o >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
            const int r = b * BATCH + P;
            if ((lane & ((32 >> LOGB) - 1)) == 0 && r < p.V)
                p.partial[((int64_t)r * p.Ttot + p.t0 + t) * p.nsl + blockIdx.x * 4 + sp] = a[0] * kDotUnscale;
        }
    }
}

// =====================================================================================
// K1 (single-trait fast path): table-lookup dot product.
//
// The shift+DFMA kernel above is bounded by instruction issue, not by the FP64 pipe: on sm_100a any
// integer instruction issued between two DFMAs costs about as much as the DFMA itself
// (tools/pipe_micro.cu: 48 DFMA/clk/SM alone, 30 with one shift each), so one trait cannot get past
// ~30 genotypes/clk/SM that way.  Here a group of 4 consecutive individuals (one byte of the column)
// is handled by ONE shared-memory lookup and ONE add:
//     table[q][byte] = sum_k field_k(byte) * eps[4q + k]          256 doubles per byte position
// The tile of a CTA (128*E4 bytes per column) is cut into chunks of <= 112 bytes; a pass builds the
// tables of one chunk (224 KB of shared memory) and streams that chunk of all V columns through
// them: lane l of a warp loads word l of the chunk (one coalesced <= 112-byte row per marker) and
// looks its 4 bytes up; 8 markers are reduced across lanes by a transposed butterfly.  Chunk c of
// pass p is tile (p*nsm + cta) / chunks_per_tile ..., i.e. in one pass the CTAs together read one
// contiguous nsm*CB-byte run of every column.  Per-marker partials are accumulated across passes
// in partial[r][t][cta] (each CTA owns its slot: plain read-modify-write).
// =====================================================================================
constexpr int kTabWarps = 16;
constexpr int kTabThreads = kTabWarps * 32;
constexpr int kTabMaxCW = 28;      // words per chunk: 28 * 4 bytes * 2 KB of table = 224 KB
constexpr int kTabBatch = 16;      // markers per warp batch (loads in flight per warp, width of the butterfly)

struct TabGeom { int cw, cwp, npass; };
__host__ __device__ inline TabGeom tab_geom(int tile_bytes) {
    TabGeom g;
    const int words = tile_bytes / 4;
    g.npass = (words + kTabMaxCW - 1) / kTabMaxCW;
    g.cw = (words + g.npass - 1) / g.npass;          // words per chunk (last chunk may be shorter)
    g.cwp = (g.cw + 3) & ~3;                         // padded so that the byte stride keeps lanes on distinct banks
    return g;
}

template <int E4>
__global__ void __launch_bounds__(kTabThreads, 1) dot_table_kernel(const DotParams p, Layout L) {
    constexpr int E = 4 * E4;
    constexpr int TILE = kLanesPerTile * E4;
    extern __shared__ __align__(16) uint8_t tsmem[];
    double* tab = reinterpret_cast<double*>(tsmem);          // [256][4][cwp]
    const TabGeom G = tab_geom(TILE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RS = 4 * G.cwp;                                // doubles between consecutive byte values
    constexpr int B = kTabBatch;
    const int nbat = (p.V + B - 1) / B;
    const double* eps = p.eps + (int64_t)p.t0 * p.npad;
    const uint32_t tab_u32 = smem_u32(tab);
    const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4, hi2 = lane & 2;
    // marker of the batch whose total this lane holds after the butterfly
    const int own = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);

    for (int pass = 0; pass < G.npass; pass++) {
        // unit u = (tile, chunk) handled by this CTA in this pass: neighbouring CTAs take neighbouring chunks
        // of the same tile, so that in one pass the grid reads one contiguous run of every column
        const int u = pass * gridDim.x + blockIdx.x;
        const int tile = u / G.npass, chunk = u % G.npass;
        const int w0 = chunk * G.cw;                         // first word of this chunk inside the tile
        const int cw = min(G.cw, TILE / 4 - w0);             // words in this chunk
        if (pass) __syncthreads();                           // everyone is done with the previous tables
        // ---- build: 4 threads per (byte position) quad, 64 byte values each
        for (int item = threadIdx.x; item < cw * 16; item += kTabThreads) {
            const int quad = item % (cw * 4), sub = item / (cw * 4);
            const int wl = quad >> 2, j = quad & 3;          // word (lane) and byte inside the word
            int ls, bb;
            tile_offset_to_slot_byte(E4, (w0 + wl) * 4 + j, ls, bb);
            const double* e = eps + ((int64_t)tile * kLanesPerTile + ls) * E + 4 * bb;
            const double e0 = e[0], e1 = e[1], e2 = e[2], e3 = e[3];
            double p01[16];
#pragma unroll
            for (int i = 0; i < 16; i++) p01[i] = (double)(i & 3) * e0 + (double)(i >> 2) * e1;
            double* dst = tab + j * G.cwp + wl;
#pragma unroll
            for (int hb = 0; hb < 4; hb++) {
                const int h = sub * 4 + hb;                  // high nibble of the byte value
                const double p23 = (double)(h & 3) * e2 + (double)(h >> 2) * e3;
#pragma unroll
                for (int i = 0; i < 16; i++) dst[(int64_t)(h * 16 + i) * RS] = p01[i] + p23;
            }
        }
        __syncthreads();

        // ---- stream the chunk of all V columns through the tables
        const uint8_t* chunk0 = p.bed + (int64_t)tile * TILE + (int64_t)w0 * 4 + lane * 4;
        const bool active = lane < cw;
        // inactive lanes (word 0 -> table[0] == 0) are parked on columns whose banks the active lanes of their
        // half-warp do not use
        const int colw = active ? lane : (lane & 15) % (cw < 16 ? cw : 16);
        uint32_t lb[4];                                      // shared-memory byte address of table[0][j][lane]
#pragma unroll
        for (int j = 0; j < 4; j++) lb[j] = tab_u32 + (uint32_t)((j * G.cwp + colw) * 8);
        const uint32_t bstride = (uint32_t)(RS * 8);

        uint32_t wn[B];
        auto fetch = [&](int bi) {
#pragma unroll
            for (int jj = 0; jj < B; jj++) {
                const int r = bi * B + jj;
                const int col = r < p.V ? p.cols[r] : -1;
                wn[jj] = (active && col >= 0) ? __ldg(reinterpret_cast<const uint32_t*>(chunk0 + (int64_t)col * p.col_stride)) : 0u;
            }
        };
        if (warp < nbat) fetch(warp);
        for (int bi = warp; bi < nbat; bi += kTabWarps) {
            uint32_t w[B];
#pragma unroll
            for (int jj = 0; jj < B; jj++) w[jj] = wn[jj];
            if (bi + kTabWarps < nbat) fetch(bi + kTabWarps);          // prefetch the next batch of this warp
            double a[B];
#pragma unroll
            for (int jj = 0; jj < B; jj++) {
                double v0, v1, v2, v3;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v0) : "r"(lb[0] + (w[jj] & 0xffu) * bstride));
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v1) : "r"(lb[1] + ((w[jj] >> 8) & 0xffu) * bstride));
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v2) : "r"(lb[2] + ((w[jj] >> 16) & 0xffu) * bstride));
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v3) : "r"(lb[3] + (w[jj] >> 24) * bstride));
                a[jj] = (v0 + v1) + (v2 + v3);               // inactive lanes hold word 0: table[0] == 0
            }
            // transposed butterfly over the 16 markers of the batch (fixed order: reproducible)
            double b8[8], b4[4], b2[2], b1;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const double send = hi16 ? a[i] : a[i + 8], keep = hi16 ? a[i + 8] : a[i];
                b8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double send = hi8 ? b8[i] : b8[i + 4], keep = hi8 ? b8[i + 4] : b8[i];
                b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double send = hi4 ? b4[i] : b4[i + 2], keep = hi4 ? b4[i + 2] : b4[i];
                b2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            {
                const double send = hi2 ? b2[0] : b2[1], keep = hi2 ? b2[1] : b2[0];
                b1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
            const int r = bi * B + own;
            if ((lane & 1) == 0 && r < p.V)                  // one slot per unit: plain stores, no read-modify-write
                p.partial[((int64_t)r * p.Ttot + p.t0) * p.nsl + u] = b1;
        }
    }
}

// =====================================================================================
// K2: one warp per virtual rank: finish the dot product, sample, publish.
// =====================================================================================
struct DotPieces { double dpa, dpb; };

// sum a*eps and sum b*eps of trait t for the marker of virtual rank r (all lanes get the result)
__device__ __forceinline__ DotPieces finish_dot(const SampleParams& p, int r, int col, int t, int lane) {
    const double* part = p.partial + ((int64_t)r * p.T + t) * p.nsl;
    double s = 0.0;
    for (int i0 = lane; i0 < p.nsl; i0 += 32 * 8) {          // 8 independent loads per round, fixed summation order
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = i0 + 32 * j < p.nsl ? part[i0 + 32 * j] : 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) s += x[j];
    }
    const double coded = warp_sum_fixed(s);                 // sum d*eps with missing coded 3
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
    d.dpa = coded - 3.0 * smiss;     // a = 0 at missing (lut_a)
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

__global__ void __launch_bounds__(128) sample_kernel(const SampleParams p) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= p.V) return;
    const int col = p.cols[v];
    if (col < 0) {
        if (lane < p.T) { p.pub[(int64_t)v * p.T + lane].lam = 0.0; p.pub[(int64_t)v * p.T + lane].mave = 0.0; }
        return;
    }
    double my_dpa = 0.0, my_dpb = 0.0;
    for (int t = 0; t < p.T; t++) {
        const DotPieces d = finish_dot(p, v, col, t, lane);
        if (lane == t) { my_dpa = d.dpa; my_dpb = d.dpb; }
    }
    if (lane >= p.T) return;
    const int t = lane;
    const int grp = p.group[col];
    const int64_t mi = (int64_t)t * p.Mloc + col;
    const double mave = p.mave[mi], msig = p.msig[mi];
    const double dot_raw = msig * (my_dpa - mave * my_dpb);                // bayes.cpp:766
    const uint32_t mglo = (uint32_t)(p.marker_begin + col);
    double u;
    const int64_t ri = ((int64_t)p.step * p.R + (p.r0 + v)) * p.T + t;
    const double sigg = p.sigmag[t * p.G + grp];
    if (p.rep_u) {
        u = p.rep_u[ri];
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
    const MarkerDraw d = sample_marker_pre(dot_raw, p.betas[mi], sigg, p.gc + ((int64_t)t * p.G + grp) * 4 * p.K, p.K, p.nonas[t], u, zdraw);
    p.betas[mi] = d.beta_new;
    if (d.comp >= 0) {
        p.comp[mi] = d.comp;                                               // bayes.cpp:462
        atomicAdd(&p.cass[(t * p.G + grp) * p.K + d.comp], 1);              // 460
    }
    PubEntry e;
    e.lam = fabs(d.dbeta) > 0.0 ? d.dbeta * msig : 0.0;                    // 483-487, phenotype.cpp:328
    e.mave = mave;
    p.pub[(int64_t)v * p.T + t] = e;
    if (e.lam != 0.0) atomicAdd(reinterpret_cast<unsigned long long*>(p.npublished), 1ull);
}

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
// K3: apply the step's published updates to the residuals.
// Phenotype::update_epsilon, phenotype.cpp:326-329,375-390:  eps += (a - mave*b) * (dbeta*msig) * na
//
// One CTA per (tile, trait), 4 x kUpdSplit warps.  Warp w serves the lane-slots of sub-partition
// w & 3 (same mapping as K1) and takes every kUpdSplit-th published marker, in rank order, prefetching
// kUpdPF columns ahead; its increments stay in registers.  The splits are then combined through shared
// memory in a fixed order (reproducible), masked by the NA mask and added to eps; the per-tile sum of
// eps that K2 needs is refreshed on the way.
// =====================================================================================
struct UpdEntry {           // one published marker, staged in shared memory
    double v[4];            // increment by genotype code: (a - mave*b) * dbeta*msig for dosage 0,1,2; 0 where missing
};

template <int E4>
__global__ void __launch_bounds__(kUpdThreads, 1) update_kernel(const UpdateParams p, Layout L) {
    constexpr int E = 4 * E4;
    extern __shared__ __align__(16) uint8_t usmem[];
    double* dlt = reinterpret_cast<double*>(usmem);                                  // [kUpdSplit][128][E + 1]
    UpdEntry* ent = reinterpret_cast<UpdEntry*>(usmem + sizeof(double) * kUpdSplit * (E + 1) * kLanesPerTile);
    int32_t* ecol = reinterpret_cast<int32_t*>(ent + kUpdCap);                        // column of each staged entry
    __shared__ int npub;
    __shared__ int wcnt[kUpdCap / 32];
    __shared__ double red[kUpdThreads / 32];
    const int t = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp = warp & 3, u = warp >> 2, ls = sp * 32 + lane;
    const uint8_t* tile0 = p.bed + (int64_t)blockIdx.x * L.tile_bytes;
    const uint32_t ent_u32 = smem_u32(ent);

    double d[E];
#pragma unroll
    for (int k = 0; k < E; k++) d[k] = 0.0;
    bool any = false;

    // rounds of kUpdCap virtual ranks: their published entries (rank order) fit the staging area
    for (int v_lo = 0; v_lo < p.V; v_lo += kUpdCap) {
        const int v_hi = min(p.V, v_lo + kUpdCap);
        __syncthreads();
        // ordered compaction of the round's published entries by the whole CTA: thread tid looks at virtual
        // ranks v_lo + tid + i*kUpdThreads (all loads in flight at once), warp counts are scanned through
        // shared memory, rank order is preserved
        constexpr int NI = kUpdCap / kUpdThreads;
        PubEntry pe[NI];
        uint32_t bal[NI];
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int v = v_lo + i * kUpdThreads + threadIdx.x;
            pe[i] = v < v_hi ? p.pub[(int64_t)v * p.T + t] : PubEntry{0.0, 0.0};
        }
#pragma unroll
        for (int i = 0; i < NI; i++) {
            bal[i] = __ballot_sync(0xffffffffu, pe[i].lam != 0.0);
            if (lane == 0) wcnt[i * (kUpdThreads / 32) + warp] = __popc(bal[i]);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int i = 0; i < NI * (kUpdThreads / 32); i++) { const int c = wcnt[i]; wcnt[i] = acc; acc += c; }
            npub = acc;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NI; i++) {
            if (pe[i].lam != 0.0) {
                const int idx = wcnt[i * (kUpdThreads / 32) + warp] + __popc(bal[i] & ((1u << lane) - 1u));
                // reference arithmetic: (mdb*b + a) * bs_  with mdb = -mave, bs_ = dbeta*msig (phenotype.cpp:328-329,388)
                const double mdb = -pe[i].mave;
                ent[idx].v[0] = (mdb * 1.0 + 0.0) * pe[i].lam;
                ent[idx].v[1] = (mdb * 1.0 + 1.0) * pe[i].lam;
                ent[idx].v[2] = (mdb * 1.0 + 2.0) * pe[i].lam;
                ent[idx].v[3] = 0.0;                 // missing: a = b = 0
                ecol[idx] = p.cols[v_lo + i * kUpdThreads + threadIdx.x];
            }
        }
        __syncthreads();
        const int n = npub;
        if (n == 0) continue;
        any = true;

        SlotRegs<E4> g[kUpdPF], gn[kUpdPF];
        const int mine = (n - u + kUpdSplit - 1) / kUpdSplit;      // entries i = u, u + S, u + 2S, ...
        auto fetch = [&](SlotRegs<E4>* dst, int c0) {
#pragma unroll
            for (int j = 0; j < kUpdPF; j++)
                if (c0 + j < mine) dst[j].load(tile0 + (int64_t)ecol[u + (c0 + j) * kUpdSplit] * p.col_stride, ls);
        };
        fetch(gn, 0);
        for (int c0 = 0; c0 < mine; c0 += kUpdPF) {
#pragma unroll
            for (int j = 0; j < kUpdPF; j++) g[j] = gn[j];
            fetch(gn, c0 + kUpdPF);                                  // prefetch the next chunk
#pragma unroll
            for (int j = 0; j < kUpdPF; j++) {
                if (c0 + j >= mine) break;
                const uint32_t eb = ent_u32 + (uint32_t)((u + (c0 + j) * kUpdSplit) * sizeof(UpdEntry));
#pragma unroll
                for (int k = 0; k < E; k++) {
                    double val;                      // v[code]: one shared-memory read (4 addresses: broadcast, no conflict)
                    asm("ld.shared.f64 %0, [%1];" : "=d"(val) : "r"(eb + g[j].field(k) * 8u));
                    d[k] += val;
                }
            }
        }
    }
    if (!__syncthreads_or(any)) return;              // nothing published for this trait: eps and its sums stand
    // increments of split u for slot ls, padded rows (E+1) keep both the writes and the reads below conflict-free
#pragma unroll
    for (int k = 0; k < E; k++) dlt[((int64_t)u * kLanesPerTile + ls) * (E + 1) + k] = d[k];
    __syncthreads();

    // combine the splits in fixed order and apply: the tile is a contiguous range of individuals -> coalesced
    constexpr int PER = kLanesPerTile * E;
    const int64_t base = (int64_t)t * p.npad + (int64_t)blockIdx.x * PER;
    double s = 0.0;
    for (int i = threadIdx.x; i < PER; i += kUpdThreads) {
        const int sl = i / E, k = i - sl * E;
        double inc = 0.0;
#pragma unroll
        for (int uu = 0; uu < kUpdSplit; uu++) inc += dlt[((int64_t)uu * kLanesPerTile + sl) * (E + 1) + k];
        double e = p.eps[base + i];
        if (p.na01[base + i]) {                      // * na  (phenotype.cpp:388)
            e += inc;
            if (p.delta) p.delta[base + i] += inc;   // multi-GPU: what this shard changed since the last exchange
        }
        p.eps[base + i] = e;
        s += e;
    }
    const double tot = block_sum_fixed(s, red);
    if (threadIdx.x == 0) p.spart[(int64_t)t * L.nsm + blockIdx.x] = tot;
}

// =====================================================================================
// per-iteration prologue / epilogue
// =====================================================================================
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

// =====================================================================================
// launchers
// =====================================================================================
void launch_transcode(const uint8_t* src, int nmark, const Layout& L, uint8_t* dst, cudaStream_t s) {
    if (nmark <= 0) return;
    dim3 grid((unsigned)((L.col_stride + 255) / 256), (unsigned)nmark);
    transcode_kernel<<<grid, 256, 0, s>>>(src, nmark, L, dst);
}
void launch_decode_column(const uint8_t* col, const Layout& L, double* a, double* b, cudaStream_t s) {
    decode_column_kernel<<<(unsigned)((L.N + 255) / 256), 256, 0, s>>>(col, L, a, b);
}
void launch_untranscode(const uint8_t* tiles, int nmark, const Layout& L, uint8_t* dst, cudaStream_t s) {
    if (nmark <= 0) return;
    dim3 grid((unsigned)((L.mbytes + 255) / 256), (unsigned)nmark);
    untranscode_kernel<<<grid, 256, 0, s>>>(tiles, nmark, L, dst);
}
void launch_generate_plink(uint8_t* dst, int nmark, int first_global_marker, const Layout& L, uint32_t seed, double maf_lo,
                           double maf_hi, double missing_rate, cudaStream_t s) {
    if (nmark <= 0) return;
    dim3 grid((unsigned)((L.mbytes + 255) / 256), (unsigned)nmark);
    generate_plink_kernel<<<grid, 256, 0, s>>>(dst, nmark, first_global_marker, L, seed, maf_lo, maf_hi, missing_rate);
}
void launch_count_missing(const uint8_t* bed, int nmark, const Layout& L, uint32_t* counts, cudaStream_t s) {
    if (nmark <= 0) return;
    count_missing_kernel<<<(nmark + 3) / 4, 128, 0, s>>>(bed, nmark, L, counts);
}
void launch_fill_missing(const uint8_t* bed, int nmark, const Layout& L, const uint32_t* off, uint32_t* idx, cudaStream_t s) {
    if (nmark <= 0) return;
    fill_missing_kernel<<<(nmark + 3) / 4, 128, 0, s>>>(bed, nmark, L, off, idx);
}
void launch_stats(const uint8_t* bed, int nmark, const Layout& L, const uint8_t* namask2, const int32_t* nonas, int T,
                  double* mave, double* msig, cudaStream_t s) {
    if (nmark <= 0) return;
    stats_kernel<<<nmark, 128, 0, s>>>(bed, nmark, L, namask2, nonas, T, mave, msig);
}

#define GMRM_DISPATCH_E4(E4v, CALL)                 \
    switch (E4v) {                                  \
    case 1: { constexpr int E4 = 1; CALL; } break;  \
    case 2: { constexpr int E4 = 2; CALL; } break;  \
    case 3: { constexpr int E4 = 3; CALL; } break;  \
    case 4: { constexpr int E4 = 4; CALL; } break;  \
    case 5: { constexpr int E4 = 5; CALL; } break;  \
    case 6: { constexpr int E4 = 6; CALL; } break;  \
    case 7: { constexpr int E4 = 7; CALL; } break;  \
    case 8: { constexpr int E4 = 8; CALL; } break;  \
    default: break;                                 \
    }

void launch_eps_offset(double* eps, const uint8_t* na01, const Layout& L, int T, const double* mu_old,
                       const double* mu_new, double* spart, cudaStream_t s) {
    dim3 grid((unsigned)L.nsm, (unsigned)T);
    eps_offset_kernel<<<grid, 256, 0, s>>>(eps, na01, L, mu_old, mu_new, spart);
}
void launch_eps_merge(double* eps, double* loc, const double* tot, const Layout& L, int T, double* spart, cudaStream_t s) {
    dim3 grid((unsigned)L.nsm, (unsigned)T);
    eps_merge_kernel<<<grid, 256, 0, s>>>(eps, loc, tot, L, spart);
}
void launch_eps_sumsq(const double* eps, int64_t npad, int64_t n, int T, double* out, cudaStream_t s) {
    eps_sumsq_kernel<<<T, 1024, 0, s>>>(eps, npad, n, out);
}

template <int E4, int T, int WPS, int BATCH>
static int dot_launch_v(const DotParams& p, int nsm, cudaStream_t s) {
    const int nbpad = (p.V + BATCH - 1) / BATCH * BATCH;
    const int smem = dot_ring_tiles(WPS) * kLanesPerTile * E4 + 2 * (dot_ring_tiles(WPS) / BATCH) * (int)sizeof(uint64_t) + nbpad * (int)sizeof(int32_t);
    static int attr = 0;
    if (smem > attr) {
        if (cudaFuncSetAttribute(dot_kernel<E4, T, WPS, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
        attr = smem;
    }
    dot_kernel<E4, T, WPS, BATCH><<<nsm, (4 * WPS + 1) * 32, smem, s>>>(p);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -2;
}
// variants (consumer warps per sub-partition, markers per batch): 0 = (2,8) default, 1 = (4,4), 2 = (2,4), 3 = (3,4).
// All are instantiated for single-trait runs (the headline path); multi-trait launches use the default.
template <int E4, int T>
static int dot_launch_t(const DotParams& p, int nsm, cudaStream_t s) {
    if (T == 1) {
        if (p.variant == 1) return dot_launch_v<E4, 1, 4, 4>(p, nsm, s);
        if (p.variant == 2) return dot_launch_v<E4, 1, 2, 4>(p, nsm, s);
        if (p.variant == 3) return dot_launch_v<E4, 1, 3, 4>(p, nsm, s);
        return dot_launch_v<E4, 1, 2, 8>(p, nsm, s);
    }
    return dot_launch_v<E4, T, 2, 4>(p, nsm, s);     // multi-trait: widest register budget (168 / thread)
}
template <int E4>
static int dot_launch_e(int T, const DotParams& p, int nsm, cudaStream_t s) {
    switch (T) {
    case 1: return dot_launch_t<E4, 1>(p, nsm, s);
    case 2: return dot_launch_t<E4, 2>(p, nsm, s);
    case 3: return dot_launch_t<E4, 3>(p, nsm, s);
    case 4: return dot_launch_t<E4, 4>(p, nsm, s);
    }
    return -1;
}
// T in 1..4 per launch (the engine chunks more traits)
int launch_dot(const Layout& L, int T, const DotParams& p, cudaStream_t s) {
    int rc = -1;
    GMRM_DISPATCH_E4(L.E4, (rc = dot_launch_e<E4>(T, p, L.nsm, s)));
    return rc;
}

template <int E4>
static int dot_table_launch_t(const DotParams& p, const Layout& L, cudaStream_t s) {
    const TabGeom G = tab_geom(L.tile_bytes);
    const int smem = 256 * 4 * G.cwp * (int)sizeof(double);
    static int attr = 0;
    if (smem > attr) {
        if (cudaFuncSetAttribute(dot_table_kernel<E4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
        attr = smem;
    }
    dot_table_kernel<E4><<<L.nsm, kTabThreads, smem, s>>>(p, L);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -2;
}
// one trait (p.t0) per launch; partial[r][t][unit] with p.nsl == nsm * passes
int dot_table_passes(const Layout& L) { return tab_geom(L.tile_bytes).npass; }
int launch_dot_table(const Layout& L, const DotParams& p, cudaStream_t s) {
    int rc = -1;
    GMRM_DISPATCH_E4(L.E4, (rc = dot_table_launch_t<E4>(p, L, s)));
    return rc;
}

void launch_sample(const SampleParams& p, cudaStream_t s) {
    if (p.V <= 0) return;
    sample_kernel<<<(p.V + 3) / 4, 128, 0, s>>>(p);
}
void launch_finish_dots(const SampleParams& p, double* out, cudaStream_t s) {
    if (p.V <= 0) return;
    finish_dots_kernel<<<(p.V + 3) / 4, 128, 0, s>>>(p, out);
}

template <int E4>
static int update_launch_t(const UpdateParams& p, const Layout& L, cudaStream_t s) {
    const int smem = (int)sizeof(double) * kUpdSplit * (4 * E4 + 1) * kLanesPerTile + kUpdCap * (int)(sizeof(UpdEntry) + sizeof(int32_t));
    static int attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        if (cudaFuncSetAttribute(update_kernel<E4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
        attr = smem;
    }
    dim3 grid((unsigned)L.nsm, (unsigned)p.T);
    update_kernel<E4><<<grid, kUpdThreads, smem, s>>>(p, L);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -2;
}
int launch_update(const Layout& L, const UpdateParams& p, cudaStream_t s) {
    int rc = -1;
    GMRM_DISPATCH_E4(L.E4, (rc = update_launch_t<E4>(p, L, s)));
    return rc;
}

void launch_steptab(int32_t* tab, int Mm, int Vl, int r0, int R, int Mt, int marker_begin, int shuffle, uint32_t seed,
                    int it, const int32_t* rep_perm, cudaStream_t s) {
    const int64_t n = (int64_t)Mm * Vl;
    if (n <= 0) return;
    steptab_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tab, Mm, Vl, r0, R, Mt, marker_begin, shuffle, seed, it, rep_perm);
}
void launch_beta_sq(const double* betas, const int32_t* group, int Mloc, int T, int G, double* out, cudaStream_t s) {
    const int smem = G * 256 * (int)sizeof(double);
    static int attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(beta_sq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr = smem;
    }
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
