// Microbenchmark 4 (prepared for round 2, DESIGN.md section 5 "next levers" item 1): can the direct decode path
// (SHF + DFMA per genotype, fp64/ALU pipes) run next to the table look-up path (PRMT + LDS.64 + DADD per four
// genotypes, LSU pipe) in the same warps, and what does the mix deliver in genotypes per clock and SM?
//   k<L, D>: per iteration every warp handles L look-up words (4 quads each) and D direct words (16 genotypes each)
//   held in registers (an LCG makes new ones; bytes are kept below 81 for the look-ups).
//   ks<WL>: warp-specialised -- WL warps do look-ups only, the others the direct path only (the form a hybrid step
//   kernel would take: the direct warps keep their rows' residuals in registers, the look-up warps keep theirs in tables).
// The table geometry is the product's (layout.h): 81 entries x 256 B per region, 2 regions per slot, 5 slots.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hybrid_micro tools/hybrid_micro.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kRegion = 81 * 256, kSlot = 2 * kRegion, kSlots = 5;
constexpr uint32_t kBase = 1024;      // absolute shared address of slot 0 (as in the product)

__device__ __forceinline__ void set_lo(double& D, uint32_t x) {
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmov.b64 %0, {%1, hi};\n\t}" : "+d"(D) : "r"(x));
}
template <int K>
__device__ __forceinline__ uint32_t quad_addr(uint32_t word, uint32_t lanebase) {   // (e << 8) | lane * 8: one PRMT
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(word), "r"(lanebase), "n"(0x7604 | (K << 4)));
    return d;
}
template <int IMM>
__device__ __forceinline__ double lds_imm(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1 + %2];" : "=d"(v) : "r"(a), "n"(IMM));
    return v;
}
template <int SLOT>
__device__ __forceinline__ double lookup_word(uint32_t w, uint32_t lanebase, double acc) {
    acc += lds_imm<(int)kBase + SLOT * kSlot>(quad_addr<0>(w, lanebase));
    acc += lds_imm<(int)kBase + SLOT * kSlot + 128>(quad_addr<1>(w, lanebase));
    acc += lds_imm<(int)kBase + SLOT * kSlot + kRegion>(quad_addr<2>(w, lanebase));
    acc += lds_imm<(int)kBase + SLOT * kSlot + kRegion + 128>(quad_addr<3>(w, lanebase));
    return acc;
}
template <int S>
__device__ __forceinline__ uint32_t shf(uint32_t x) {
    uint32_t d;
    asm volatile("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "n"(S));
    return d;
}
// 16 genotypes of a 2-bit word: the shifted word goes into the low half of a denormal multiplier (round-1a trick)
template <int K = 0>
__device__ __forceinline__ void direct_word(uint32_t w, double& D, const double (&wt)[16], double& acc) {
    if constexpr (K < 16) {
        set_lo(D, shf<2 * K>(w));
        acc = fma(D, wt[K], acc);
        direct_word<K + 1>(w, D, wt, acc);
    }
}

__device__ __forceinline__ void fill_tables() {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(smem);
    double* tab = reinterpret_cast<double*>(smem + (kBase - b0));
    for (int i = threadIdx.x; i < kSlots * kSlot / 8; i += blockDim.x) tab[i] = 1e-3 * (i & 1023);
    __syncthreads();
}

// the stream of one warp: L look-up words and D direct words per iteration
template <int L, int D>
__device__ __forceinline__ double stream(const double* zero, int iters) {
    const uint32_t lanebase = (threadIdx.x & 15) * 8;
    uint32_t lw[L > 0 ? L : 1], dw[D > 0 ? D : 1];
    double lacc[L > 0 ? L : 1], dacc[D > 0 ? D : 1], Dm[D > 0 ? D : 1], wt[16];
#pragma unroll
    for (int i = 0; i < 16; i++) wt[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int i = 0; i < (L > 0 ? L : 1); i++) { lw[i] = 0x10203040u * (i + 1) + threadIdx.x; lacc[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < (D > 0 ? D : 1); i++) { dw[i] = 0x9e3779b9u * (i + 1) + threadIdx.x; dacc[i] = 0.0; Dm[i] = zero[threadIdx.x + 32 * i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < (L > D ? L : D); i++) {
            if (i < L) {
                const uint32_t w = lw[i] & 0x3f3f3f3fu;                  // bytes 0..63: valid table entries
                lacc[i] = (i % kSlots == 0) ? lookup_word<0>(w, lanebase, lacc[i]) : (i % kSlots == 1) ? lookup_word<1>(w, lanebase, lacc[i])
                        : (i % kSlots == 2) ? lookup_word<2>(w, lanebase, lacc[i]) : (i % kSlots == 3) ? lookup_word<3>(w, lanebase, lacc[i])
                        : lookup_word<4>(w, lanebase, lacc[i]);
                lw[i] = lw[i] * 1664525u + 1013904223u;
            }
            if (i < D) {
                direct_word(dw[i], Dm[i], wt, dacc[i]);
                dw[i] = dw[i] * 1664525u + 1013904223u;
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < (L > 0 ? L : 1); i++) s += lacc[i];
#pragma unroll
    for (int i = 0; i < (D > 0 ? D : 1); i++) s += dacc[i] + Dm[i];
    return s;
}

// every warp runs the same mix
template <int L, int D>
__global__ void __launch_bounds__(512, 1) k(double* out, const double* zero, int iters) {
    fill_tables();
    out[blockIdx.x * blockDim.x + threadIdx.x] = stream<L, D>(zero, iters);
}
// warp-specialised: warps below WL run look-ups only (8 words per iteration), the others the direct path only (4 words)
template <int WL>
__global__ void __launch_bounds__(512, 1) ks(double* out, const double* zero, int iters_lookup, int iters_direct) {
    fill_tables();
    const bool lookup = (int)(threadIdx.x >> 5) < WL;
    out[blockIdx.x * blockDim.x + threadIdx.x] = lookup ? stream<8, 0>(zero, iters_lookup) : stream<0, 4>(zero, iters_direct);
}

template <int L, int D>
void run(int warps, int nsm, double* out, double* zero) {
    const int iters = 2048, smem = 232448;
    cudaFuncSetAttribute(k<L, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<L, D><<<nsm, warps * 32, smem>>>(out, zero, 8);
    cudaEventRecord(a);
    k<L, D><<<nsm, warps * 32, smem>>>(out, zero, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double geno = (double)warps * 32 * iters * 16.0 * (L + D);
    const double per_clk = geno / (ms * 1e-3) / (clk * 1e3);
    printf("look-up words %d + direct words %d per iteration, %2d warps/SM: %7.3f ms  %6.2f genotypes/clk/SM (look-up share %5.1f, direct %5.1f)  (%s)\n",
           L, D, warps, ms, per_clk, per_clk * L / (L + D), per_clk * D / (L + D), cudaGetErrorString(cudaGetLastError()));
}
// iteration counts are set so that both roles finish together when the direct warps sustain `ratio` times the per-warp
// genotype rate of the look-up warps; the printed split shows what each role delivered in the common time
template <int WL>
void run_spec(int warps, int nsm, double* out, double* zero, double ratio) {
    const int it_l = 2048, it_d = (int)(2048 * ratio * 8 / 4), smem = 232448;
    cudaFuncSetAttribute(ks<WL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    ks<WL><<<nsm, warps * 32, smem>>>(out, zero, 8, 8);
    cudaEventRecord(a);
    ks<WL><<<nsm, warps * 32, smem>>>(out, zero, it_l, it_d);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double gl = (double)WL * 32 * it_l * 16.0 * 8, gd = (double)(warps - WL) * 32 * it_d * 16.0 * 4;
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("specialised: %2d look-up warps + %2d direct warps, direct/look-up per-warp rate %.2f: %7.3f ms  %6.2f genotypes/clk/SM (look-up %5.1f, direct %5.1f)  (%s)\n",
           WL, warps - WL, ratio, ms, (gl + gd) / cyc, gl / cyc, gd / cyc, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out, *zero;
    cudaMalloc(&out, 8 * nsm * 1024); cudaMalloc(&zero, 8 * 32 * 64);
    cudaMemset(zero, 0, 8 * 32 * 64);
    for (int warps : {8, 16}) {
        run<8, 0>(warps, nsm, out, zero); run<0, 4>(warps, nsm, out, zero);
        run<8, 1>(warps, nsm, out, zero); run<8, 2>(warps, nsm, out, zero); run<6, 2>(warps, nsm, out, zero);
        run<4, 2>(warps, nsm, out, zero); run<4, 4>(warps, nsm, out, zero);
    }
    for (double ratio : {0.5, 0.75, 1.0, 1.5}) {
        run_spec<12>(16, nsm, out, zero, ratio); run_spec<13>(16, nsm, out, zero, ratio); run_spec<14>(16, nsm, out, zero, ratio);
    }
    run_spec<16>(16, nsm, out, zero, 1.0);      // all look-up, through the specialised kernel (control)
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
