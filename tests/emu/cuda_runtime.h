// cuda_runtime.h -- TEST INFRASTRUCTURE: a host stand-in for the CUDA programming model, just large enough to compile
// the library's NON-tensor-core kernels (no PTX) with g++ and run them on the CPU.  tests/test_kernel_emulation.py puts
// this directory first on the include path, so `#include <cuda_runtime.h>` in csrc/mtm_internal.cuh resolves here.
//
// Two launchers:
//   emu_launch       every thread of every block one after the other (kernels whose threads do not communicate);
//   emu_launch_coop  one FIBER (user-level context, ucontext) per CUDA thread of a block, blocks one after the other, all on
//                    the calling OS thread: __syncthreads(), warp shuffles / votes (full-warp, all lanes converged), shared
//                    memory (`__shared__` becomes `static`) and atomics behave as on the device.  A fiber that waits (block /
//                    warp barrier, mbarrier) is parked until its condition holds; when no fiber can run, the kernel has
//                    deadlocked and the process aborts with a message.  Scheduling is round-robin and deterministic.
// The emulation checks kernel LOGIC (index arithmetic, reductions, rounding) bit for bit; it says nothing about speed.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <ucontext.h>
#include <vector>

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() = default;
    dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static dim3 threadIdx, blockIdx, blockDim, gridDim;      // threadIdx follows the running fiber

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __grid_constant__
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
// __byte_perm(x, y, s): byte i of the result is byte (s >> 4i) & 7 of the 8-byte value {y, x}; selector bit 3 replicates its sign bit
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    const unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned out = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 15u;
        unsigned b = (unsigned)(v >> (8 * (sel & 7u))) & 255u;
        if (sel & 8u) b = (b & 128u) ? 255u : 0u;
        out |= b << (8 * i);
    }
    return out;
}
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long expected, unsigned long long desired)
{
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;                                                   // the value found, as CUDA's atomicCAS returns it
}
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

// ---- cooperative block emulation: fibers ------------------------------------------------------------------------------
struct EmuFiber {
    ucontext_t ctx;
    dim3 tid;
    int lane = 0, warp = 0;
    bool done = false;
    const volatile uint32_t* wait_ptr = nullptr;      // parked until *wait_ptr != wait_val
    uint32_t wait_val = 0;
};
struct EmuGate { volatile uint32_t gen = 0; int arrived = 0, expected = 0; };   // a re-usable barrier
struct EmuWarp { EmuGate gate; uint64_t slot[32]; int lanes = 0; };
struct EmuBlock {
    EmuGate gate;
    std::vector<EmuWarp> warps;
    int vote[2] = {0, 0};
    uint32_t vote_round = 0;
};
static constexpr size_t EMU_STACK = 256 * 1024;
static std::vector<std::unique_ptr<char[]>> emu_stacks;      // kept across launches
static std::vector<EmuFiber> emu_fibers;
static ucontext_t emu_sched_ctx;
static int emu_cur = -1;
static EmuBlock* emu_block = nullptr;
static uint8_t* emu_dyn_smem = nullptr;          // dynamic shared memory of the running block (`extern __shared__` is rewritten to it)
static int emu_lane = 0, emu_warp = 0;
static bool emu_coop = false;
static void (*emu_body_call)(void*) = nullptr;
static void* emu_body_ptr = nullptr;

// EMU_SCHED_SEED=<n> (environment): instead of round-robin, the scheduler visits the runnable fibers in a pseudo-random order and the
// modelled hardware operations (tests/emu/tcgen05_model.h) become preemption points -- other legal interleavings of the same kernel,
// to look for races in its synchronisation.  Deterministic for a given seed.
static uint64_t emu_rng_state = 0;
static bool emu_random_sched = false;
static inline uint32_t emu_rand()
{
    emu_rng_state = emu_rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(emu_rng_state >> 33);
}
static inline void emu_sched_init()
{
    const char* v = getenv("EMU_SCHED_SEED");
    emu_random_sched = v != nullptr;
    if (v) emu_rng_state = 0x9E3779B97F4A7C15ull ^ (uint64_t)atoll(v);
}
static const volatile uint32_t emu_always_ready = 1;
// a point where the running fiber may be preempted (random scheduling only)
static inline void emu_preempt_point()
{
    if (!emu_random_sched || emu_cur < 0 || (emu_rand() & 3u) != 0) return;
    EmuFiber& f = emu_fibers[emu_cur];
    f.wait_ptr = &emu_always_ready; f.wait_val = 0;         // runnable again at once
    swapcontext(&f.ctx, &emu_sched_ctx);
}

// parks the running fiber until *ptr != val
static inline void emu_wait_change(const volatile uint32_t* ptr, uint32_t val)
{
    if (*ptr != val) return;
    EmuFiber& f = emu_fibers[emu_cur];
    f.wait_ptr = ptr; f.wait_val = val;
    swapcontext(&f.ctx, &emu_sched_ctx);
}
static inline void emu_gate_release(EmuGate& g) { g.arrived = 0; g.gen = g.gen + 1; }
static inline void emu_gate_arrive_and_wait(EmuGate& g)
{
    const uint32_t gen = g.gen;
    if (++g.arrived == g.expected) emu_gate_release(g);
    else emu_wait_change(&g.gen, gen);
}
// an exiting thread no longer takes part in the barrier (like an exited CUDA thread)
static inline void emu_gate_drop(EmuGate& g)
{
    if (--g.expected > 0 && g.arrived == g.expected) emu_gate_release(g);
}

static inline void __syncthreads() { if (emu_coop) emu_gate_arrive_and_wait(emu_block->gate); }
static inline int __syncthreads_or(int pred)
{
    if (!emu_coop) return pred != 0;
    EmuBlock& b = *emu_block;
    const uint32_t round = b.vote_round;                    // changes only when the barrier of this vote releases
    if (pred) b.vote[round & 1] = 1;
    if (b.gate.arrived + 1 == b.gate.expected) { b.vote[(round + 1) & 1] = 0; b.vote_round = round + 1; }   // last arrival: next vote's slot
    emu_gate_arrive_and_wait(b.gate);
    return b.vote[round & 1];
}
static inline void __syncwarp(unsigned = 0xffffffffu) { if (emu_coop) emu_gate_arrive_and_wait(emu_block->warps[emu_warp].gate); }

template <typename T> static inline T emu_exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    EmuWarp& w = emu_block->warps[emu_warp];
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.slot[emu_lane] = bits;
    emu_gate_arrive_and_wait(w.gate);
    T out = v;
    if (src_lane >= 0 && src_lane < w.lanes) memcpy(&out, &w.slot[src_lane], sizeof(T));
    emu_gate_arrive_and_wait(w.gate);
    return out;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_exchange(v, emu_lane + d < 32 ? emu_lane + d : -1); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_exchange(v, emu_lane - d); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int d) { return emu_exchange(v, emu_lane ^ d); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_exchange(v, lane & 31); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
    EmuWarp& w = emu_block->warps[emu_warp];
    w.slot[emu_lane] = pred ? 1u : 0u;
    emu_gate_arrive_and_wait(w.gate);
    unsigned m = 0;
    for (int l = 0; l < w.lanes; ++l) m |= (unsigned)(w.slot[l] != 0) << l;
    emu_gate_arrive_and_wait(w.gate);
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

static void emu_fiber_main()
{
    emu_body_call(emu_body_ptr);
    EmuFiber& f = emu_fibers[emu_cur];
    f.done = true;
    emu_gate_drop(emu_block->gate);
    emu_gate_drop(emu_block->warps[f.warp].gate);
    swapcontext(&f.ctx, &emu_sched_ctx);                   // never resumed
}

// one fiber per CUDA thread, the blocks of the grid one after the other
template <typename F> static void emu_launch_coop(dim3 grid, dim3 block, F body)
{
    gridDim = grid; blockDim = block;
    const int nthreads = (int)(block.x * block.y * block.z);
    while ((int)emu_stacks.size() < nthreads) emu_stacks.emplace_back(new char[EMU_STACK]);
    emu_sched_init();
    emu_body_call = [](void* p) { (*static_cast<F*>(p))(); };
    emu_body_ptr = &body;
    emu_coop = true;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = dim3(bx, by, bz);
                EmuBlock blk;
                blk.gate.expected = nthreads;
                blk.warps.resize((nthreads + 31) / 32);
                for (size_t w = 0; w < blk.warps.size(); ++w)
                    blk.warps[w].lanes = blk.warps[w].gate.expected = std::min(32, nthreads - 32 * (int)w);
                emu_block = &blk;
                emu_fibers.assign(nthreads, EmuFiber());
                for (int t = 0; t < nthreads; ++t) {
                    EmuFiber& f = emu_fibers[t];
                    f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    f.lane = t & 31; f.warp = t >> 5;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = emu_stacks[t].get();
                    f.ctx.uc_stack.ss_size = EMU_STACK;
                    f.ctx.uc_link = nullptr;
                    makecontext(&f.ctx, emu_fiber_main, 0);
                }
                int live = nthreads;
                std::vector<int> visit(nthreads);
                for (int t = 0; t < nthreads; ++t) visit[t] = t;
                while (live > 0) {
                    bool ran = false;
                    if (emu_random_sched)
                        for (int t = nthreads - 1; t > 0; --t) std::swap(visit[t], visit[emu_rand() % (uint32_t)(t + 1)]);
                    for (int v = 0; v < nthreads; ++v) {
                        const int t = visit[v];
                        EmuFiber& f = emu_fibers[t];
                        if (f.done) continue;
                        if (f.wait_ptr) {
                            if (*f.wait_ptr == f.wait_val) continue;          // still parked
                            f.wait_ptr = nullptr;
                        }
                        emu_cur = t; threadIdx = f.tid; emu_lane = f.lane; emu_warp = f.warp;
                        swapcontext(&emu_sched_ctx, &f.ctx);
                        ran = true;
                        if (f.done) --live;
                    }
                    if (!ran) {
                        fprintf(stderr, "emulated kernel deadlocked in block (%u, %u, %u): %d threads wait for something that cannot happen "
                                "(first of them: thread %d)\n", bx, by, bz, live,
                                (int)(std::find_if(emu_fibers.begin(), emu_fibers.end(), [](const EmuFiber& f) { return !f.done; }) - emu_fibers.begin()));
                        abort();
                    }
                }
                emu_block = nullptr;
            }
    emu_coop = false;
    emu_cur = -1;
}

// ---- a stand-in for the CUDA runtime API (whole-library host build, tests/emu_library.py) -----------------------------------
// Everything is synchronous and "device memory" is host memory; one emulated device that reports compute capability 10.0.
// EMU_SM_COUNT keeps persistent / capped grids small so that the emulation stays fast.
#ifndef EMU_SM_COUNT
#define EMU_SM_COUNT 4
#endif
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaErrorNotReady = 600, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = EMU_SM_COUNT; char name[64] = "emulated sm_100a"; };
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = reinterpret_cast<cudaStream_t>(new int(0)); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete reinterpret_cast<int*>(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(new int(0)); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete reinterpret_cast<int*>(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }   // launches are synchronous here
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type = cudaMemoryTypeUnregistered; int device = 0; void* devicePointer = nullptr; void* hostPointer = nullptr; };
// every host pointer reads as pageable: the CPU build then exercises the chunked staging pipeline of host_staging.cu
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { *a = cudaPointerAttributes(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(aligned_alloc(256, (n + 255) / 256 * 256 + 256)); return cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
template <typename T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaHostGetDevicePointer(T** d, void* h, unsigned) { *d = static_cast<T*>(h); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr)
{
    for (size_t y = 0; y < h; ++y) memcpy(static_cast<char*>(d) + y * dp, static_cast<const char*>(s) + y * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }

// kernel<<<grid, block, smem, stream>>>(args) of the library's host code is rewritten to this (tests/emu_library.py)
static long long emu_launch_count = 0;
template <typename F> static void emu_launch_dyn(dim3 grid, dim3 block, size_t smem, F body)
{
    std::unique_ptr<uint8_t[]> store(new uint8_t[smem + 2048]);
    uint8_t* base = store.get() + ((1024 - (reinterpret_cast<uintptr_t>(store.get()) & 1023)) & 1023);
    memset(base, 0xCD, smem);
    emu_dyn_smem = base;
    ++emu_launch_count;
    emu_launch_coop(grid, block, body);
    emu_dyn_smem = nullptr;
}
