// Microbenchmark 2: which ingredient of the decode loop costs throughput?  All variants run the same
// 8-chain DFMA stream; they differ in how the multiplier is produced.  tools/dfma_micro.cu measured the
// DFMA-only peak (~57 DFMA/clk/SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void set_lo(double& D, uint32_t x) {
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmov.b64 %0, {%1, hi};\n\t}" : "+d"(D) : "r"(x));
}

// MODE 0: constant multipliers, 28 distinct weights      (RF pressure only)
// MODE 1: + one shift per DFMA writing the low word in place (the decode loop)
// MODE 2: multiplier built by a fresh {x, 0} pack (shift + mov per DFMA)
// MODE 3: like 1, but only the alu pipe for shifts (shf via funnel) -- if ptxas honours it
template <int MODE, int NM>
__global__ void k(double* out, const double* zero, const uint32_t* words, int iters) {
    double w[28], acc[NM], D[NM];
    uint32_t g[NM];
#pragma unroll
    for (int i = 0; i < 28; i++) w[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int m = 0; m < NM; m++) { acc[m] = m; D[m] = zero[m * blockDim.x + threadIdx.x]; g[m] = words[m * blockDim.x + threadIdx.x]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 28; kk++) {
#pragma unroll
            for (int m = 0; m < NM; m++) {
                if (MODE == 0) {
                    acc[m] = fma(D[m], w[kk], acc[m]);
                } else if (MODE == 1 || MODE == 3) {
                    set_lo(D[m], g[m] << (kk % 16 * 2));
                    acc[m] = fma(D[m], w[kk], acc[m]);
                } else {
                    double d = __hiloint2double(0, (int)(g[m] << (kk % 16 * 2)));
                    acc[m] = fma(d, w[kk], acc[m]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < NM; m++) g[m] = g[m] * 1664525u + 1013904223u;   // new "genotype words" each round
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < NM; m++) s += acc[m] + D[m];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NM>
void run(int warps, int nsm, double* out, double* zero, uint32_t* words) {
    const int iters = 512;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, NM><<<nsm, warps * 32>>>(out, zero, words, 4);
    cudaEventRecord(a);
    k<MODE, NM><<<nsm, warps * 32>>>(out, zero, words, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double fmas = (double)nsm * warps * 32 * iters * 28.0 * NM;
    printf("mode %d chains %d warps/SM %2d : %7.3f ms  %6.2f DFMA/clk/SM\n", MODE, NM, warps, ms, fmas / (ms * 1e-3) / nsm / (clk * 1e3));
}

int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out, *zero; uint32_t* words;
    cudaMalloc(&out, 8 * nsm * 1024); cudaMalloc(&zero, 8 * 8 * 1024); cudaMalloc(&words, 4 * 8 * 1024);
    cudaMemset(zero, 0, 8 * 8 * 1024); cudaMemset(words, 0x5a, 4 * 8 * 1024);
    for (int warps : {4, 8, 12, 16}) {
        run<0, 8>(warps, nsm, out, zero, words);
        run<1, 8>(warps, nsm, out, zero, words);
        run<2, 8>(warps, nsm, out, zero, words);
        run<1, 4>(warps, nsm, out, zero, words);
        run<1, 2>(warps, nsm, out, zero, words);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
