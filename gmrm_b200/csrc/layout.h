// HBM layout of the genotype matrix ("base-3 quads") and the geometry of the table-lookup dot
// product.  Host + device.
//
// Reference layout (Bayes::load_genotype, src/bayes.cpp:867-900): marker-major, ceil(N/4) bytes
// per marker, individual 4q+k in bits 2k..2k+1 of byte q, PLINK codes 00 = dosage 2, 01 = missing,
// 10 = dosage 1, 11 = dosage 0 (decode tables src/lut/mk_lut.cpp:25-32).
//
// Device layout.  Same byte order as the .bed column (byte q <-> individuals 4q..4q+3, a "quad"),
// but the byte holds the quad's dosages in base 3:
//     e = d0 + 3 d1 + 9 d2 + 27 d3      (d = allele count 0,1,2;  e in 0..80)
// A missing genotype is stored as dosage 0 and listed in a per-marker CSR list of individuals
// (miss_off / miss_idx), which makes the transcode invertible bit for bit and gives
//     sum a*eps = sum_q table_q[e_q]          sum b*eps = sum eps - sum_{missing} eps.
// Columns are padded with zero bytes to a whole number of ROWS of 64 bytes (16 words, 256
// individuals); the residuals are padded with zeros likewise (npad = 256 * nrows).
//
// Why base 3: the dot product is taken by table look-up -- one shared-memory read and one fp64 add
// per QUAD instead of one multiply-add per genotype (on sm_100a the per-genotype
// fp64 path is limited to ~30 genotypes/clk/SM by issue, the look-up path to 64 by shared-memory
// bandwidth).  A quad's table has 81 entries x 8 B; with 2-bit codes it would need 256 (3.2x the
// shared memory, 3.2x the build work, 3x the passes).
//
// Table geometry in shared memory (one CTA per SM).  A CTA owns a contiguous range of rows.  The
// tables of one (row, trait) SLOT take kSlotBytes = 2 regions x 81 entries x 256 B:
//     address(slot, k, e, l) = kTabBase + slot*kSlotBytes + (k>>1)*kRegionBytes + e*256 + (k&1)*128 + l*8
// for byte k (0..3) of word l (0..15) of the row.  For a fixed k the 16 lanes of a half-warp hit 16
// distinct 8-byte bank pairs whatever their e: every look-up is conflict-free.  e*256 + l*8 is formed
// by ONE byte-permute (PRMT) of the genotype word with the lane constant l*8; the rest is the
// immediate offset of the LDS instruction.  That requires absolute shared addresses, hence kTabBase.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define GMRM_HD __host__ __device__ __forceinline__
#else
#define GMRM_HD inline
#endif
#ifndef GMRM_UNROLL
#if defined(__CUDA_ARCH__)
#define GMRM_UNROLL _Pragma("unroll")
#else
#define GMRM_UNROLL            // host pass: the pragma is unknown to g++
#endif
#endif

namespace gmrm {

constexpr int kRowBytes = 64;                        // bytes of one column per row
constexpr int kRowWords = 16;
constexpr int kRowInd = 256;                         // individuals per row
constexpr int kTabEntries = 81;
constexpr int kRegionBytes = kTabEntries * 256;      // 20,736
constexpr int kSlotBytes = 2 * kRegionBytes;         // 41,472
constexpr int kMaxSlots = 5;                         // (row, trait) slots per pass: 207,360 B of tables
constexpr uint32_t kTabBase = 1024;                  // absolute shared-memory address of slot 0

struct Layout {
    int32_t N = 0;        // individuals
    int32_t mbytes = 0;   // ceil(N/4): bytes per PLINK column
    int32_t nrows = 0;    // rows per column
    int32_t nsm = 0;      // CTAs sharing the rows (one per SM)
    int64_t col_stride = 0;   // nrows * 64
    int64_t npad = 0;         // nrows * 256 individuals incl. padding

    GMRM_HD int row_begin(int cta) const { return (int)((int64_t)cta * nrows / nsm); }
    GMRM_HD int max_rows_per_cta() const { return (nrows + nsm - 1) / nsm; }
};

inline Layout make_layout(int32_t N, int nsm) {
    Layout L;
    L.N = N;
    L.mbytes = (N + 3) / 4;
    L.nrows = (L.mbytes + kRowBytes - 1) / kRowBytes;
    L.nsm = nsm;
    L.col_stride = (int64_t)L.nrows * kRowBytes;
    L.npad = (int64_t)L.nrows * kRowInd;
    return L;
}

// PLINK byte (4 codes) -> base-3 quad byte + 4-bit mask of missing genotypes, and back.
GMRM_HD uint8_t plink_to_tri(uint8_t x, uint32_t* missmask) {
    uint32_t e = 0, mm = 0, w = 1;
GMRM_UNROLL
    for (int k = 0; k < 4; k++) {
        const uint32_t c = (x >> (2 * k)) & 3u;
        const uint32_t d = c == 0 ? 2u : (c == 2 ? 1u : 0u);   // 00 -> 2, 10 -> 1, 11 -> 0, 01 (missing) -> 0 + flag
        e += d * w;
        w *= 3;
        if (c == 1) mm |= 1u << k;
    }
    *missmask = mm;
    return (uint8_t)e;
}
// base-3 quad byte -> the four dosages as 2-bit fields (field k = dosage of individual 4q+k)
GMRM_HD uint32_t tri_to_fields(uint32_t e) {
    const uint32_t d3 = e / 27u, r3 = e - 27u * d3, d2 = r3 / 9u, r2 = r3 - 9u * d2, d1 = r2 / 3u, d0 = r2 - 3u * d1;
    return d0 | (d1 << 2) | (d2 << 4) | (d3 << 6);
}
// dosage fields + missing mask -> PLINK byte
GMRM_HD uint8_t fields_to_plink(uint32_t f, uint32_t missmask) {
    uint32_t x = 0;
GMRM_UNROLL
    for (int k = 0; k < 4; k++) {
        const uint32_t d = (f >> (2 * k)) & 3u;
        uint32_t c = d == 2 ? 0u : (d == 1 ? 2u : 3u);
        if ((missmask >> k) & 1u) c = 1u;
        x |= c << (2 * k);
    }
    return (uint8_t)x;
}

}  // namespace gmrm
