// cuda_runtime.h -- TEST INFRASTRUCTURE: a host stand-in for the CUDA programming model, just large enough to compile
// the library's NON-tensor-core kernels (no PTX) with g++ and run them on the CPU.  tests/test_kernel_emulation.py puts
// this directory first on the include path, so `#include <cuda_runtime.h>` in csrc/mtm_internal.cuh resolves here.
//
// Two launchers:
//   emu_launch       every thread of every block one after the other (kernels whose threads do not communicate);
//   emu_launch_coop  one OS thread per CUDA thread of a block, blocks one after the other: __syncthreads(), warp
//                    shuffles / votes (full-warp, all lanes converged), shared memory (`__shared__` becomes `static`)
//                    and atomics behave as on the device.
// The emulation checks kernel LOGIC (index arithmetic, reductions, rounding) bit for bit; it says nothing about speed.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() = default;
    dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static thread_local dim3 threadIdx;
static dim3 blockIdx, blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

// ---- scalar intrinsics -------------------------------------------------------------------------------------------
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }            // approximate on the device: only compared kernel-to-kernel
static inline int __float2int_rn(float x) { return (int)nearbyintf(x); }     // default rounding mode: half to even
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c)
{
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 255u) * ((b >> (8 * k)) & 255u);
    return c;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift)
{
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (shift & 31));
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
using std::max;
using std::min;
static inline long long max(long long a, int b) { return a > b ? a : (long long)b; }
#define CUDART_INF_F (__builtin_inff())

// ---- atomics (shared or global memory: plain host memory here) ----------------------------------------------------
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline int atomicMax(int* p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- cooperative block emulation ------------------------------------------------------------------------------------
struct EmuWarp {
    std::unique_ptr<std::barrier<>> bar;
    uint64_t slot[32];
    int lanes = 0;
};
struct EmuBlock {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<EmuWarp> warps;
    std::atomic<int> vote{0};
};
static EmuBlock* emu_block = nullptr;
static uint8_t* emu_dyn_smem = nullptr;          // dynamic shared memory of the running block (`extern __shared__` is rewritten to it)
static thread_local int emu_lane = 0, emu_warp = 0;
static thread_local bool emu_coop = false;

static inline void __syncthreads() { if (emu_coop) emu_block->bar->arrive_and_wait(); }
static inline int __syncthreads_or(int pred)
{
    if (!emu_coop) return pred != 0;
    if (pred) emu_block->vote.store(1);
    emu_block->bar->arrive_and_wait();
    const int r = emu_block->vote.load();
    emu_block->bar->arrive_and_wait();
    if (threadIdx.x == 0 && threadIdx.y == 0) emu_block->vote.store(0);
    emu_block->bar->arrive_and_wait();
    return r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { if (emu_coop) emu_block->warps[emu_warp].bar->arrive_and_wait(); }

template <typename T> static inline T emu_exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    EmuWarp& w = emu_block->warps[emu_warp];
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.slot[emu_lane] = bits;
    w.bar->arrive_and_wait();
    T out = v;
    if (src_lane >= 0 && src_lane < w.lanes) memcpy(&out, &w.slot[src_lane], sizeof(T));
    w.bar->arrive_and_wait();
    return out;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_exchange(v, emu_lane + d < 32 ? emu_lane + d : -1); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_exchange(v, emu_lane - d); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_exchange(v, lane & 31); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
    EmuWarp& w = emu_block->warps[emu_warp];
    w.slot[emu_lane] = pred ? 1u : 0u;
    w.bar->arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < w.lanes; ++l) m |= (unsigned)(w.slot[l] != 0) << l;
    w.bar->arrive_and_wait();
    return m;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }

// every thread of every block, serially (valid for kernels whose threads do not communicate)
template <typename F> static void emu_launch(dim3 grid, dim3 block, F body)
{
    gridDim = grid; blockDim = block;
    emu_coop = false;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx) {
                            blockIdx = dim3(bx, by, bz);
                            threadIdx = dim3(tx, ty, tz);
                            body();
                        }
}

// one OS thread per CUDA thread (created once per launch, walking the blocks together); a thread that returns early leaves the
// block barrier like an exited CUDA thread
template <typename F> static void emu_launch_coop(dim3 grid, dim3 block, F body)
{
    gridDim = grid; blockDim = block;
    const int nthreads = (int)(block.x * block.y * block.z);
    const long long nblocks = (long long)grid.x * grid.y * grid.z;
    std::barrier<> between(nthreads);                      // separates the blocks of the launch
    std::unique_ptr<EmuBlock> blk;
    auto next_block = [&](long long b) {
        blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long long)grid.x * grid.y)));
        blk = std::make_unique<EmuBlock>();
        blk->bar = std::make_unique<std::barrier<>>(nthreads);
        blk->warps.resize((nthreads + 31) / 32);
        for (size_t w = 0; w < blk->warps.size(); ++w) {
            blk->warps[w].lanes = std::min(32, nthreads - 32 * (int)w);
            blk->warps[w].bar = std::make_unique<std::barrier<>>(blk->warps[w].lanes);
        }
        emu_block = blk.get();
    };
    next_block(0);
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&, t] {
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            emu_lane = t & 31; emu_warp = t >> 5; emu_coop = true;
            for (long long b = 0; b < nblocks; ++b) {
                body();
                emu_block->bar->arrive_and_drop();
                between.arrive_and_wait();                 // every thread has left block b
                if (b + 1 < nblocks) {
                    if (t == 0) next_block(b + 1);
                    between.arrive_and_wait();
                }
            }
        });
    for (auto& th : pool) th.join();
    emu_block = nullptr;
}
