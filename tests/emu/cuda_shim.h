// cuda_shim.h -- TEST INFRASTRUCTURE: just enough of the CUDA programming model to compile the SIMPLE kernels of
// csrc/ (no shared memory, no warp intrinsics, no PTX) for the host and run their threads one after the other.
// tests/test_kernel_emulation.py pastes kernel source text from csrc/*.cu behind this header; the point is to check the
// kernels' index arithmetic and rounding rules on the CPU, bit for bit, where no GPU is available.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

struct dim3 { unsigned x = 1, y = 1, z = 1; };
static dim3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t a, uint32_t b) { uint2 r{a, b}; return r; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }          // only compared between two emulated kernels
static inline int __float2int_rn(float x) { return (int)nearbyintf(x); }   // default rounding mode: half to even
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
template <typename T> static inline T __ldg(const T* p) { return *p; }

// every thread of every block, serially (valid for kernels whose threads do not communicate)
template <typename F> static void emu_launch(dim3 grid, dim3 block, F body)
{
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx)
                for (unsigned ty = 0; ty < block.y; ++ty)
                    for (unsigned tx = 0; tx < block.x; ++tx) {
                        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz; threadIdx.x = tx; threadIdx.y = ty;
                        body();
                    }
}
