// TEST INFRASTRUCTURE ONLY.  A minimal host stand-in for the CUDA execution model, enough to run the kernels of
// gmrm_b200/csrc/predict.cu on the CPU: CTAs one after the other, one std::thread per CUDA thread, __syncthreads as a
// barrier, warp shuffles through a per-warp exchange buffer.  It checks indexing and logic, nothing about performance.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
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

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)

inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
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
