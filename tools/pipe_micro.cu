// Microbenchmark 3: which integer instruction co-issues with DFMA on sm_100a?  8 DFMA chains per warp;
// each DFMA is preceded by one integer op that produces the low word of its (denormal) multiplier.
//   OP 0: none   1: SHF (alu pipe)   2: IMAD (fma pipe)   3: LOP3 (alu)   4: IADD3 (alu)   5: PRMT (alu)
//   6: two ops (SHF + LOP3)          7: FFMA-pipe float op (FMUL) producing garbage lo
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void set_lo(double& D, uint32_t x) {
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmov.b64 %0, {%1, hi};\n\t}" : "+d"(D) : "r"(x));
}
template <int OP, int S>
__device__ __forceinline__ uint32_t iop(uint32_t x, uint32_t y) {
    uint32_t d;
    if (OP == 1) asm volatile("shf.l.clamp.b32 %0, %2, %1, %3;" : "=r"(d) : "r"(x), "r"(0u), "n"(S));
    else if (OP == 2) asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "n"(1u << S));
    else if (OP == 3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(x), "r"(y), "n"(0x01010101u * (S + 1)));
    else if (OP == 4) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "n"(S * 977 + 1));
    else if (OP == 5) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "n"(0x3210 ^ (S & 3)));
    else if (OP == 6) { uint32_t t; asm volatile("shf.l.clamp.b32 %0, %2, %1, %3;" : "=r"(t) : "r"(x), "r"(0u), "n"(S));
                        asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(t), "r"(y), "n"(0x0f0f0f0fu)); }
    else if (OP == 7) { float f; asm volatile("mul.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(x)), "f"(1.0f + S)); d = __float_as_uint(f); }
    else d = x;
    return d;
}
template <int OP, int K, int NM>
struct Body {
    static __device__ __forceinline__ void run(double (&acc)[NM], double (&D)[NM], const uint32_t (&g)[NM], const double (&w)[28]) {
#pragma unroll
        for (int m = 0; m < NM; m++) {
            if (OP != 0) set_lo(D[m], iop<OP, (K % 15) * 2 + 1>(g[m], g[(m + 1) % NM]));
            acc[m] = fma(D[m], w[K], acc[m]);
        }
        Body<OP, K + 1, NM>::run(acc, D, g, w);
    }
};
template <int OP, int NM> struct Body<OP, 28, NM> { static __device__ __forceinline__ void run(double (&)[NM], double (&)[NM], const uint32_t (&)[NM], const double (&)[28]) {} };

template <int OP, int NM>
__global__ void k(double* out, const double* zero, const uint32_t* words, int iters) {
    double w[28], acc[NM], D[NM];
    uint32_t g[NM];
#pragma unroll
    for (int i = 0; i < 28; i++) w[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int m = 0; m < NM; m++) { acc[m] = m; D[m] = zero[m * blockDim.x + threadIdx.x]; g[m] = words[m * blockDim.x + threadIdx.x]; }
    for (int it = 0; it < iters; it++) {
        Body<OP, 0, NM>::run(acc, D, g, w);
#pragma unroll
        for (int m = 0; m < NM; m++) g[m] = g[m] * 1664525u + 1013904223u;
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < NM; m++) s += acc[m] + D[m];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP, int NM>
void run(int warps, int nsm, double* out, double* zero, uint32_t* words) {
    const int iters = 512;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP, NM><<<nsm, warps * 32>>>(out, zero, words, 4);
    cudaEventRecord(a);
    k<OP, NM><<<nsm, warps * 32>>>(out, zero, words, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double fmas = (double)nsm * warps * 32 * iters * 28.0 * NM;
    printf("op %d chains %d warps/SM %2d : %7.3f ms  %6.2f DFMA/clk/SM  (%s)\n", OP, NM, warps, ms, fmas / (ms * 1e-3) / nsm / (clk * 1e3),
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out, *zero; uint32_t* words;
    cudaMalloc(&out, 8 * nsm * 1024); cudaMalloc(&zero, 8 * 8 * 1024); cudaMalloc(&words, 4 * 8 * 1024);
    cudaMemset(zero, 0, 8 * 8 * 1024); cudaMemset(words, 0x5a, 4 * 8 * 1024);
    for (int warps : {4, 8}) {
        run<0, 8>(warps, nsm, out, zero, words); run<1, 8>(warps, nsm, out, zero, words); run<2, 8>(warps, nsm, out, zero, words);
        run<3, 8>(warps, nsm, out, zero, words); run<4, 8>(warps, nsm, out, zero, words); run<5, 8>(warps, nsm, out, zero, words);
        run<6, 8>(warps, nsm, out, zero, words); run<7, 8>(warps, nsm, out, zero, words);
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
