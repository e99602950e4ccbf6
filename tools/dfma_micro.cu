// Microbenchmark behind DESIGN.md "decode algebra": FP64 FMA throughput on sm_100a when the
// multiplier is a DENORMAL (hi word 0, lo word = shifted genotype word) versus a normal number, and
// the dependent-issue latency of DFMA (how many independent chains a warp needs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dfma_micro tools/dfma_micro.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CHAINS, bool DENORM>
__global__ void k(double* out, const double* zero, int iters, uint32_t seedbits) {
    double acc[CHAINS], m[CHAINS];
    const double w = 1.0000001 + threadIdx.x * 1e-9;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
        acc[c] = c;
        uint32_t lo = seedbits * (c + 1) + threadIdx.x;
        m[c] = DENORM ? __hiloint2double((int)__double2loint(zero[threadIdx.x]), (int)lo) : (double)(lo | 1u) * 1e-9;
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) acc[c] = fma(m[c], w, acc[c]);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS, bool DENORM>
void run(int warps_per_sm, int nsm, double* out, double* zero) {
    const int iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<CHAINS, DENORM><<<nsm, warps_per_sm * 32>>>(out, zero, 16, 12345u);
    cudaEventRecord(a);
    k<CHAINS, DENORM><<<nsm, warps_per_sm * 32>>>(out, zero, iters, 12345u);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double fmas = (double)nsm * warps_per_sm * 32 * iters * 8 * CHAINS;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s chains=%d warps/SM=%2d : %7.3f ms  %8.1f GFMA/s  %6.2f DFMA/clk/SM (at %d MHz)  %.2f clk per dependent DFMA per warp-chain\n",
           DENORM ? "denormal" : "normal  ", CHAINS, warps_per_sm, ms, fmas / ms / 1e6, fmas / (ms * 1e-3) / nsm / (clk * 1e3), clk / 1000,
           (ms * 1e-3) * (clk * 1e3) / (iters * 8.0));
}

int main() {
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out, *zero;
    cudaMalloc(&out, sizeof(double) * nsm * 1024);
    cudaMalloc(&zero, sizeof(double) * 1024);
    cudaMemset(zero, 0, sizeof(double) * 1024);
    run<1, false>(4, nsm, out, zero);   // 1 warp per sub-partition, 1 chain: pure latency
    run<1, true>(4, nsm, out, zero);
    run<2, true>(4, nsm, out, zero);
    run<4, true>(4, nsm, out, zero);
    run<8, true>(4, nsm, out, zero);
    run<8, false>(4, nsm, out, zero);
    run<4, true>(8, nsm, out, zero);
    run<4, true>(12, nsm, out, zero);
    run<8, true>(8, nsm, out, zero);
    run<8, false>(8, nsm, out, zero);
    run<8, true>(16, nsm, out, zero);
    run<8, false>(16, nsm, out, zero);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
