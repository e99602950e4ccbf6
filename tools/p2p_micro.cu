// Microbenchmark 5 (round 2): what one hop of a peer-memory exchange costs between two B200s of one box -- every CTA of GPU A
// stores a block into GPU B's memory, makes it visible, raises a flag there, and waits for B's flag (and the other way round);
// 148 CTAs x 512 threads on each GPU as in the step kernel.  Modes:
//   0  flag only (volatile store, volatile poll): the bare NVLink round trip
//   1  data + __threadfence_system() by EVERY thread + barrier + fence by the flag writer   (round-2 scheme)
//   2  data + barrier + ONE __threadfence_system() by the flag writer
//   3  data + barrier + st.release.sys of the flag by the flag writer, ld.acquire.sys poll
//   4  data carries its own sequence number (16-byte stores {payload, seq}; the reader polls the DATA): no flag, no fence
// Every mode checks the received payload (errors are counted), so a scheme that is fast but wrong shows.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_micro tools/p2p_micro.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int kThreads = 512;

struct Side {
    unsigned long long* flag_mine;   // [ctas]
    unsigned long long* flag_peer;
    ulonglong2* buf_mine;            // [ctas][threads][words]
    ulonglong2* buf_peer;
    unsigned long long* errors;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 ld_volatile_v2(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) pingpong(Side s, int iters, int words, unsigned long long base) {
    const int tid = threadIdx.x, cta = blockIdx.x;
    unsigned long long nerr = 0;
    for (int it = 1; it <= iters; it++) {
        const unsigned long long seq = base + it;
        if (MODE != 0)
            for (int w = 0; w < words; w++) {
                ulonglong2 v; v.x = seq * 1000003ull + (unsigned long long)(tid * 16 + w); v.y = seq;
                s.buf_peer[(((size_t)(it & 1) * gridDim.x + cta) * kThreads + tid) * words + w] = v;   // two buffers take turns: the peer may still read the last one
            }
        if (MODE == 1) __threadfence_system();
        if (MODE != 4) {
            __syncthreads();
            if (tid == 0) {
                if (MODE == 1 || MODE == 2) __threadfence_system();
                if (MODE == 3) st_release_sys(s.flag_peer + cta, seq);
                else *reinterpret_cast<volatile unsigned long long*>(s.flag_peer + cta) = seq;
                if (MODE == 3) { while (ld_acquire_sys(s.flag_mine + cta) < seq) __nanosleep(20); }
                else { while (*reinterpret_cast<volatile unsigned long long*>(s.flag_mine + cta) < seq) __nanosleep(20); }
                if (MODE == 1 || MODE == 2) __threadfence();
            }
            __syncthreads();
        }
        if (MODE != 0)
            for (int w = 0; w < words; w++) {
                const ulonglong2* src = s.buf_mine + (((size_t)(it & 1) * gridDim.x + cta) * kThreads + tid) * words + w;
                ulonglong2 v;
                if (MODE == 4) { do { v = ld_volatile_v2(src); } while (v.y < seq); }
                else v = __ldcg(src);
                if (v.x != seq * 1000003ull + (unsigned long long)(tid * 16 + w) || v.y != seq) nerr++;
            }
        if (MODE == 4) __syncthreads();
    }
    if (nerr) atomicAdd(s.errors, nerr);
}

template <int MODE>
int run(int ctas, int words, int iters, Side* side, cudaStream_t* st, unsigned long long& base) {
    cudaEvent_t a[2], b[2];
    for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaEventCreate(&a[d])); CK(cudaEventCreate(&b[d])); }
    for (int rep = 0; rep < 2; rep++) {                      // first repetition warms up
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaEventRecord(a[d], st[d])); pingpong<MODE><<<ctas, kThreads, 0, st[d]>>>(side[d], iters, words, base); CK(cudaEventRecord(b[d], st[d])); }
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
        base += iters;
    }
    float ms[2];
    unsigned long long err[2];
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d)); CK(cudaEventElapsedTime(&ms[d], a[d], b[d]));
        CK(cudaMemcpy(&err[d], side[d].errors, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemset(side[d].errors, 0, 8));
    }
    printf("mode %d  ctas %3d  %5d B per thread (%7.1f KB per CTA): %7.2f us per hop (GPU0) %7.2f (GPU1)  payload errors %llu %llu\n", MODE, ctas,
           words * 16, words * 16.0 * kThreads / 1024, ms[0] * 1e3 / iters, ms[1] * 1e3 / iters, err[0], err[1]);
    return 0;
}

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
    const int ctas = 148, maxwords = 8;
    Side side[2];
    cudaStream_t st[2];
    unsigned long long* flags[2]; ulonglong2* bufs[2];
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d));
        const cudaError_t pe = cudaDeviceEnablePeerAccess(1 - d, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { printf("no peer access %d -> %d\n", d, 1 - d); return 1; }
        CK(cudaStreamCreate(&st[d]));
        CK(cudaMalloc(&flags[d], ctas * 8)); CK(cudaMemset(flags[d], 0, ctas * 8));
        CK(cudaMalloc(&bufs[d], (size_t)2 * ctas * kThreads * maxwords * 16)); CK(cudaMemset(bufs[d], 0, (size_t)2 * ctas * kThreads * maxwords * 16));
        CK(cudaMalloc(&side[d].errors, 8)); CK(cudaMemset(side[d].errors, 0, 8));
    }
    for (int d = 0; d < 2; d++) { side[d].flag_mine = flags[d]; side[d].flag_peer = flags[1 - d]; side[d].buf_mine = bufs[d]; side[d].buf_peer = bufs[1 - d]; }
    for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
    unsigned long long base = 0;
    const int iters = 300;
    if (run<0>(ctas, 1, iters, side, st, base)) return 1;
    if (run<0>(1, 1, iters, side, st, base)) return 1;
    for (int words : {1, 4, 8}) {
        if (run<1>(ctas, words, iters, side, st, base)) return 1;
        if (run<2>(ctas, words, iters, side, st, base)) return 1;
        if (run<3>(ctas, words, iters, side, st, base)) return 1;
        if (run<4>(ctas, words, iters, side, st, base)) return 1;
    }
    if (run<1>(1, 4, iters, side, st, base)) return 1;
    if (run<3>(1, 4, iters, side, st, base)) return 1;
    if (run<4>(1, 4, iters, side, st, base)) return 1;
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
