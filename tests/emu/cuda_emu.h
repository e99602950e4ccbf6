// TEST INFRASTRUCTURE ONLY.  A minimal host stand-in for the CUDA execution model, enough to run the kernels of
// gmrm_b200/csrc/predict.cu on the CPU: CTAs one after the other, one std::thread per CUDA thread, __syncthreads as a
// barrier, warp shuffles through a per-warp exchange buffer.  It checks indexing and logic, nothing about performance.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; EmuDim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline thread_local EmuDim3 threadIdx, blockIdx;
inline EmuDim3 blockDim, gridDim;
inline std::unique_ptr<std::barrier<>> emu_block_barrier;
inline std::vector<std::unique_ptr<std::barrier<>>> emu_warp_barrier;
inline std::vector<double> emu_warp_buf;     // [warp][32]
inline std::vector<uint64_t> emu_warp_buf64;  // [warp][32] raw 64-bit payloads (indexed shuffles)

// shared memory of the CTA being run: "absolute shared addresses" are offsets into this array
alignas(256) inline uint8_t emu_smem_storage[256 * 1024];

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct int2 { int x, y; };
struct uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct alignas(16) double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }

inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline double __hiloint2double(int hi, int lo) { const uint64_t r = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &r, 8); return d; }
inline double __shfl_xor_sync(unsigned, double v, int o) {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    emu_warp_buf[warp * 32 + lane] = v;
    emu_warp_barrier[warp]->arrive_and_wait();
    const double r = emu_warp_buf[warp * 32 + (lane ^ (unsigned)o)];
    emu_warp_barrier[warp]->arrive_and_wait();
    return r;
}
inline int __shfl_xor_sync(unsigned m, int v, int o) { return (int)__shfl_xor_sync(m, (double)v, o); }   // exact for |v| < 2^53
inline int __popc(unsigned x) { return __builtin_popcount(x); }
// indexed shuffle: every lane of the warp publishes its value, then reads lane `src`'s
template <typename T> inline T emu_shfl_idx(T v, int src) {
    static_assert(sizeof(T) <= 8, "payload");
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    emu_warp_buf64[warp * 32 + lane] = raw;
    emu_warp_barrier[warp]->arrive_and_wait();
    raw = emu_warp_buf64[warp * 32 + ((unsigned)src & 31u)];
    emu_warp_barrier[warp]->arrive_and_wait();
    T r; memcpy(&r, &raw, sizeof(T));
    return r;
}
inline int __shfl_sync(unsigned, int v, int src) { return emu_shfl_idx(v, src); }
inline double __shfl_sync(unsigned, double v, int src) { return emu_shfl_idx(v, src); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const unsigned lane = threadIdx.x & 31;
    const T r = emu_shfl_idx(v, lane >= delta ? (int)(lane - delta) : (int)lane);
    return lane >= delta ? r : v;
}
inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return emu_shfl_idx(v, src); }
inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    emu_warp_buf64[warp * 32 + lane] = pred ? 1u : 0u;
    emu_warp_barrier[warp]->arrive_and_wait();
    unsigned m = 0;
    for (unsigned l = 0; l < 32; l++) m |= (unsigned)emu_warp_buf64[warp * 32 + l] << l;
    emu_warp_barrier[warp]->arrive_and_wait();
    return m;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
template <typename T> inline T __ldcg(const T* p) { return *p; }
inline long long clock64() { return 0; }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline void __trap() { abort(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)((const uint8_t*)p - emu_smem_storage); }
template <typename T> inline T emu_lds(uint32_t addr) { T v; memcpy(&v, emu_smem_storage + addr, sizeof(T)); return v; }
template <typename T> inline void emu_sts(uint32_t addr, T v) { memcpy(emu_smem_storage + addr, &v, sizeof(T)); }
// prmt.b32 in its default mode: result byte i = byte (selector nibble i & 7) of {b, a}; bit 3 of a nibble replicates the sign
inline uint32_t emu_prmt(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t n = (sel >> (4 * i)) & 0xfu;
        uint32_t byte = (uint32_t)(src >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        d |= byte << (8 * i);
    }
    return d;
}
using std::max;
using std::min;

// kernel<<<grid, block>>>(args...)  ->  emu_launch(grid, block, [&] { kernel(args...); })
// A thread that returns early drops out of the barriers it has not reached (as exited CUDA threads do).
inline void emu_launch(EmuDim3 grid, EmuDim3 block, const std::function<void()>& body) {
    gridDim = grid; blockDim = block;
    const unsigned nthreads = block.x, nwarps = (nthreads + 31) / 32;
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
            emu_block_barrier = std::make_unique<std::barrier<>>(nthreads);
            emu_warp_barrier.clear();
            for (unsigned w = 0; w < nwarps; w++) emu_warp_barrier.push_back(std::make_unique<std::barrier<>>(std::min(32u, nthreads - 32 * w)));
            emu_warp_buf.assign((size_t)nwarps * 32, 0.0);
            emu_warp_buf64.assign((size_t)nwarps * 32, 0);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nthreads; t++)
                th.emplace_back([&, t] {
                    threadIdx = EmuDim3(t); blockIdx = EmuDim3(bx, by);
                    body();
                    emu_block_barrier->arrive_and_drop();
                    emu_warp_barrier[t >> 5]->arrive_and_drop();
                });
            for (auto& x : th) x.join();
        }
}

// individuals (< N) without a phenotype, per trait: what gmrm_set_phenotype collects for stats_kernel (engine.cu)
inline void emu_na_lists(const uint8_t* mask4, int64_t col_stride, int T, int N, std::vector<uint32_t>& off, std::vector<uint32_t>& idx) {
    off.assign((size_t)T + 1, 0u);
    idx.clear();
    for (int t = 0; t < T; t++) {
        for (int i = 0; i < N; i++)
            if (!((mask4[(int64_t)t * col_stride + (i >> 2)] >> (i & 3)) & 1)) idx.push_back((uint32_t)i);
        off[(size_t)t + 1] = (uint32_t)idx.size();
    }
    if (idx.empty()) idx.push_back(0u);
}
