// HBM layout of the genotype matrix ("tile-planar 2-bit dosage") and the decode algebra the
// kernels use.  Host + device.
//
// Reference layout (Bayes::load_genotype, src/bayes.cpp:867-900): marker-major, ceil(N/4) bytes
// per marker, individual 4i+k in bits 2k..2k+1 of byte i, PLINK codes 00 = dosage 2, 01 = missing,
// 10 = dosage 1, 11 = dosage 0 (decode tables src/lut/mk_lut.cpp:25-32).
//
// Device layout.  The individuals are cut into `nsm` tiles (one per SM / CTA); a tile has 128
// lane-slots (4 SM sub-partitions x 32 lanes); a lane-slot owns E = 4*E4 CONSECUTIVE individuals
//     individual i  ->  slot s = i / E,  position k = i % E
// i.e. E4 consecutive bytes of the PLINK column.  A column is stored tile after tile
// (column stride = nsm * 128 * E4 bytes); inside a tile the E4 bytes of a slot are split into
// register-sized GROUPS -- E4/4 32-bit words, then a 16-bit half if E4 & 2, then a byte if E4 & 1 --
// and stored plane by plane (all 128 slots' word 0, then word 1, ..., then the halves, then the
// bytes), so that a warp's load of one group is one contiguous, conflict-free 128/64/32-byte run
// and one tile is one contiguous cp.async.bulk of 128*E4 bytes.
//
// Codes are re-coded so that the 2-bit field IS the dosage:  0,1,2 = allele count, 3 = missing.
// That makes a group register  g = sum_k d_k 4^k  and lets the dot product be taken WITHOUT
// extracting fields:  with  X_k = (g << (30-2k)) mod 2^32 = 2^30 * sum_{j<=k} d_j 4^(j-k)
//     sum_k X_k * w_k = 2^30 * sum_j d_j eps_j      when  w_k = eps_k - eps_{k+1}/4   (eps_n := 0)
// (the sum telescopes).  X_k is fed to the FP64 pipe as the denormal double (hi = 0, lo = X_k)
// = X_k * 2^-1074, and w_k is pre-scaled by 2^1000, so one marker costs one shift and one DFMA per
// genotype and the partial sums come out scaled by 2^-44 exactly.  See DESIGN.md "decode algebra".
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define GMRM_HD __host__ __device__ __forceinline__
#else
#define GMRM_HD inline
#endif

namespace gmrm {

constexpr int kLanesPerTile = 128;   // 4 sub-partitions x 32 lanes
constexpr int kMaxE4 = 8;            // <= 32 individuals per lane-slot (two 32-bit words)
constexpr double kWeightScale = 1.0715086071862673e+301;   // 2^1000
constexpr double kDotUnscale = 17592186044416.0;           // 2^44  (= 2^1074 / 2^30 / 2^1000)

struct Layout {
    int32_t N = 0;        // individuals
    int32_t mbytes = 0;   // ceil(N/4): bytes per PLINK column
    int32_t nsm = 0;      // tiles per column
    int32_t E4 = 0;       // bytes per lane-slot
    int32_t E = 0;        // individuals per lane-slot
    int32_t tile_bytes = 0;
    int64_t col_stride = 0;
    int64_t npad = 0;     // nsm * 128 * E individuals incl. padding

    GMRM_HD int nwords() const { return E4 / 4; }
    GMRM_HD int nhalf() const { return (E4 % 4) / 2; }
    GMRM_HD int nbyte() const { return E4 % 2; }
};

// Smallest E4 such that nsm tiles cover N individuals.  Returns 0 if N does not fit.
inline int choose_E4(int64_t N, int nsm) {
    const int64_t per4 = (int64_t)nsm * kLanesPerTile * 4;
    const int64_t e4 = (N + per4 - 1) / per4;
    return e4 < 1 ? 1 : (e4 > kMaxE4 ? 0 : (int)e4);
}

inline Layout make_layout(int32_t N, int nsm) {
    Layout L;
    L.N = N;
    L.mbytes = (N + 3) / 4;
    L.nsm = nsm;
    L.E4 = choose_E4(N, nsm);
    L.E = 4 * L.E4;
    L.tile_bytes = kLanesPerTile * L.E4;
    L.col_stride = (int64_t)nsm * L.tile_bytes;
    L.npad = (int64_t)nsm * kLanesPerTile * L.E;
    return L;
}

// Offset, inside a tile, of byte b (0..E4-1) of lane-slot ls (0..127).
GMRM_HD int tile_byte_offset(int E4, int ls, int b) {
    const int nw = E4 / 4;
    if (b < 4 * nw) return (b / 4) * (kLanesPerTile * 4) + ls * 4 + (b % 4);
    int off = nw * kLanesPerTile * 4;
    b -= 4 * nw;
    if (E4 & 2) {
        if (b < 2) return off + ls * 2 + b;
        off += kLanesPerTile * 2;
        b -= 2;
    }
    return off + ls + b;
}

// PLINK byte (4 codes) -> dosage byte (4 fields: 0,1,2 = allele count, 3 = missing), and back.
//   00->10, 01->11, 10->01, 11->00 :  out_hi = ~in_hi, out_lo = in_hi ^ in_lo
GMRM_HD uint8_t plink_to_dosage(uint8_t x) {
    return (uint8_t)(((~x) & 0xAA) | (((x >> 1) ^ x) & 0x55));
}
GMRM_HD uint8_t dosage_to_plink(uint8_t y) {
    return (uint8_t)(((~y) & 0xAA) | ((((~y) >> 1) ^ y) & 0x55));
}

// Weights of one group of n genotypes: w_k = 2^1000 * (eps_k - eps_{k+1}/4), eps_n := 0.
GMRM_HD double group_weight(double eps_k, double eps_k1) {
    return kWeightScale * (eps_k - 0.25 * eps_k1);
}

}  // namespace gmrm
