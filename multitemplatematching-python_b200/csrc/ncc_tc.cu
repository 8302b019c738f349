// ncc_tc.cu -- K3: sliding-window correlation on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// The numerator of cv2.matchTemplate (MTM/__init__.py:92)
//     CC[t](y, x) = sum_dy sum_dx I[y+dy][x+dx] * T_t[dy][dx]
// is computed EXACTLY as an integer GEMM per template row dy:
//     D[m][n] += A_dy[m][k] * B_dy[n][k]          tcgen05.mma kind::i8, u8 x u8 -> 32-bit in TMEM
//   n (N, up to 256)  : output row y            B_dy[n][k] = I[y0 + n + dy][x0 + k]
//   k (K = 32*nk)     : image column in the tile
//   m (M = 128)       : (template, x-offset)    A_dy[m][k] = T_t[dy][k - j]   (banded / Toeplitz)
// so the image tile is loaded ONCE into shared memory and every dy just moves the B
// descriptor's start address by one 16-byte row (no-swizzle K-major layout
// [k-block][row][16 B], SBO = 128 B, LBO = rows*16 B).  The Toeplitz A slabs are
// expanded once per template set (toeplitz_prep_kernel) and streamed with
// cp.async.bulk + mbarrier through a shared-memory ring.
//   mode A: m = (t, r):  8 templates x 16 x-offsets     slab [k-block][t][r][16 B]
//   mode B: m = (q', r): 1 template x 128 x-offsets, x = x0 + 16*(7-q') + r; slab
//           [block b][r][16 B] with LBO = 256 B so that consecutive K blocks alias the next
//           row groups (the band is shift-invariant) -- 8x less slab data.
// Accumulators (128 lanes x N columns, 32-bit) live in TMEM; the epilogue reads them with
// tcgen05.ld and applies OpenCV's normalisation in exact-integer + fp32 form:
//     score = (A*CC - S*sumT) * rsqrt(A*Q - S^2) * rsqrt(A*sumT2 - sumT^2),  clamped to [-1, 1],
// with S and rsqrt(A*Q - S^2) precomputed per window size (window_moments_kernel).
// Kernels: ncc_tc_persist_kernel<PROF, EW, MODE> (default: persistent, warp-specialised, two image tiles
// and two accumulators in flight) and ncc_tc_kernel<MODE> (one tile per CTA, for tiles too large to
// double-buffer).  MODE selects the epilogue: 0 the default above, 1 OpenCV's float64 rules for the
// other five methods, 2 weighted accumulation of a byte-plane product (16-bit images).
// Roofline: tensor pipe (kind::i8, 8192 MAC/clk/SM); useful fraction w / (32*nk).
#include "mtm_internal.cuh"
#include "ncc_epilogue.cuh"
#if defined(__CUDACC__)
#include <cuda.h>          // CUtensorMap + the encoder's prototype; cuTensorMapEncodeTiled itself is resolved at run time (no libcuda link)
#endif
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

constexpr int TC_THREADS = 256;
constexpr int TC_STAGES = 4;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    // the suspend-time hint (ns) lets the hardware park the thread instead of spinning: a polling lane
    // must not starve a producer warp that shares its scheduler (highest warp id wins the arbiter)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(200000u) : "memory");
    return ok != 0;
}
// one non-blocking poll (no suspend-time hint): true when the phase with this parity has completed
__device__ __forceinline__ bool mbar_try_wait_once(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    // try_wait suspends for a bounded time per call; a barrier that never completes (a bug, not a
    // load condition) traps instead of hanging the GPU.
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 16)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA: one box (16 bytes x box rows) of the image tensor map -> shared memory, completion on an mbarrier (bytes of the whole
// box, zero-filled parts included).  `tmap` points at a __grid_constant__ kernel parameter.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int x, int y, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// Tensor map of the padded u8 image (tile layout of mtm_set_image): dim0 = bytes of a row (pitch), dim1 = rows; box = 16 bytes x
// `box_rows` rows, no swizzle, out-of-bounds elements read as zero (rows below the image, bytes beyond the pitch).
static bool encode_tile_map(CUtensorMap* out, const uint8_t* base, int64_t pitch, int H, int box_rows)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeFn>(fn);
    }();
    if (!encode || box_rows < 1 || box_rows > 256 || (pitch & 15) || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
    const cuuint32_t elem_strides[2] = {1u, 1u};
    return encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, elem_strides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Exact A*CC - S*sumT (operands below 2^32, the difference inside +-2^63) rounded to float.  Written in PTX so that the
// operands stay 32-bit for ptxas: two IMAD.WIDE.U32 (the second with the negated 64-bit addend) and one I2F.S64; the C++
// form compiled to 64 x 32-bit multiplies with zero high words (six integer instructions per pixel).
__device__ __forceinline__ float n1_to_float(uint32_t area, uint32_t cc, uint32_t s, uint32_t sumT)
{
    float f;
    asm("{\n\t.reg .u64 t, u;\n\tmul.wide.u32 t, %3, %4;\n\tmul.wide.u32 u, %1, %2;\n\tsub.s64 u, u, t;\n\tcvt.rn.f32.s64 %0, u;\n\t}"
        : "=f"(f) : "r"(area), "r"(cc), "r"(s), "r"(sumT));
    return f;
}

// K-major, no-swizzle shared-memory matrix descriptor (sm_100 "version 1").
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}


// Normalise 16 consecutive accumulator columns (output rows y_first .. y_first+15) of one lane.
// Loads are issued for all 16 rows up front (addresses clamped, stores predicated) so that one
// L2 round trip covers the whole batch.
// Pixels scoring above the threshold are also appended to the candidate list (a few thousand at most):
// the 3x3 local-maximum test then only visits those instead of streaming every score map again.
struct CandSink {
    DevHit* list; int32_t* count; int cap; float thr; int tmpl, w, h;
};

__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// Single channel: the window moments are interleaved {S, bits of rsqrt(A*Q - S^2)} so that one 8-byte load serves
// a pixel.  `ahead` > 0: also pull the batch `ahead` rows further down into L1 (the next batch of this warp).
// (r_first: row inside the band; rows_avail: rows of the band that exist for this template; band: rows of the ring segment)
__device__ __forceinline__ void prefetch_moments16(const uint2* __restrict__ SR, int r_first, int rows_avail, int band, int x)
{
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int r = min(r_first + k, rows_avail - 1);
        prefetch_l1(SR + mom_index(x, r, band));
    }
}

// Hits-only epilogues (STORE == false: MODE 3) write no score map: pixels above the threshold go to the candidate list as
// before, and for the N_object == 1 search every lane keeps the best score it has seen with its row-major map index (the
// first occurrence: a lane walks its column downwards, ties keep the earlier row).
struct BestTrack { float r; uint32_t idx; };

template <bool STORE>
__device__ __forceinline__ void epilogue16(const uint32_t (&v)[16], int y_first, int mh, int mw, int x, uint32_t area,
                                           uint32_t sumT, float ct, bool is_const, const uint2* __restrict__ SR,
                                           float* __restrict__ out, const CandSink& sink, bool prefetch_next,
                                           BestTrack& bt, bool track, int y_base, int band)
{
    // mh: first row that is NOT computed for this template in this launch (end of its map or of the band)
    uint2 m[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int y = min(y_first + k, mh - 1);
        m[k] = __ldg(SR + mom_index(x, y - y_base, band));
    }
    if (prefetch_next) prefetch_moments16(SR, y_first + 16 - y_base, mh - y_base, band, x);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int y = y_first + k;
        // (32-bit operands: A <= 66051 and sumT <= 255 * 66051 on the tensor path)
        float r = n1_to_float(area, v[k], m[k].x, sumT) * __uint_as_float(m[k].y) * ct;
        r = fminf(1.0f, fmaxf(-1.0f, r));
        if (is_const) r = 1.0f;
        if (y < mh) {
            if (STORE) out[(int64_t)y * mw + x] = r;
            if (!STORE && track && r > bt.r) { bt.r = r; bt.idx = (uint32_t)(y * mw + x); }
            if (sink.list && r > sink.thr) {
                const int slot = atomicAdd(sink.count, 1);
                if (slot < sink.cap) {
                    DevHit c;
                    c.tmpl = sink.tmpl; c.x = x; c.y = y; c.w = sink.w; c.h = sink.h; c.score = r; c.seq = 0; c.key = 0.f;
                    sink.list[slot] = c;
                }
            }
        }
    }
}

// Fast form of epilogue16 for the common case (all 16 rows inside the map, template not constant): 32-bit operands,
// one running offset, no per-pixel bounds tests.  A <= 66051 and sumT <= 255*66051 fit 32 bits on the tensor path.
template <bool STORE>
__device__ __forceinline__ void epilogue16_fast(const uint32_t (&v)[16], int y_first, int mw, int x, uint32_t area, uint32_t sumT,
                                                float ct, const uint2* __restrict__ SR, float* __restrict__ out,
                                                const CandSink& sink, float thr, bool prefetch_next, BestTrack& bt, bool track,
                                                int y_base, int band)
{
    const uint2* sr = SR + mom_index(x, y_first - y_base, band);       // consecutive rows: 16 entries = 128 bytes apart (immediate offsets)
    float* o = out + (int64_t)y_first * mw + x;
    const uint32_t idx0 = (uint32_t)(y_first * mw + x);
    uint2 m[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m[k] = __ldg(sr + 16 * k);
    if (prefetch_next) {
#pragma unroll
        for (int k = 0; k < 16; ++k) prefetch_l1(sr + 16 * (16 + k));
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float r = n1_to_float(area, v[k], m[k].x, sumT) * __uint_as_float(m[k].y) * ct;
        r = fminf(1.0f, fmaxf(-1.0f, r));
        if (STORE) o[(uint32_t)(k * mw)] = r;
        if (!STORE && track) { const bool better = r > bt.r; bt.r = better ? r : bt.r; bt.idx = better ? idx0 + (uint32_t)(k * mw) : bt.idx; }
        if (r > thr) {
            const int slot = atomicAdd(sink.count, 1);
            if (slot < sink.cap) {
                DevHit c;
                c.tmpl = sink.tmpl; c.x = x; c.y = y_first + k; c.w = sink.w; c.h = sink.h; c.score = r; c.seq = 0; c.key = 0.f;
                sink.list[slot] = c;
            }
        }
    }
}

// Any method (cv2.TM_* 0..5): OpenCV's float64 epilogue (ncc_epilogue.cuh) on the exact tensor-core numerator and
// the exact window sums from the summed-area tables; four pixels' table loads are in flight at a time.
template <int C>
__device__ __forceinline__ void epilogue16_generic(const uint32_t (&v)[16], int y_first, int mh, int mw, int x, int method,
                                                   const TmplMeta& tm, const SatView& sat, float* __restrict__ out)
{
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const int y = y_first + k;
        if (y < mh) {
            uint32_t S[C];
#pragma unroll
            for (int c = 0; c < C; ++c) S[c] = sat_window_s(sat.s + c * sat.plane, sat.pitch, y, x, tm.h, tm.w);
            const unsigned long long Q = sat_window_q(sat.q, sat.pitch, y, x, tm.h, tm.w);
            out[(int64_t)y * mw + x] = ncc_epilogue<C>(method, (double)v[k], S, Q, tm);
        }
    }
}

// 16-bit path: one byte-plane product (exact u32) joins the exact numerator map, CC = 65536 hh + 256 (hl + lh) + ll.
__device__ __forceinline__ void epilogue16_accum(const uint32_t (&v)[16], int y_first, int mh, int mw, int x, double weight,
                                                 bool first, double* __restrict__ acc)
{
    double old[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int y = min(y_first + k, mh - 1);
        old[k] = first ? 0.0 : acc[(int64_t)y * mw + x];
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int y = y_first + k;
        if (y < mh) acc[(int64_t)y * mw + x] = fma(weight, (double)v[k], old[k]);
    }
}

// Multi-channel (interleaved RGB / RGBA) form: per-channel window sums, the squared sums share one table.
//   N1 = A*CC - sum_c S_c*sumT_c ;  rsD already holds rsqrt(A*Q - sum_c S_c^2).
template <int C, bool STORE>
__device__ __forceinline__ void epilogue16_mc(const uint32_t (&v)[16], int y_first, int mh, int mw, int x, long long area,
                                              const long long (&sumT)[MTM_MAX_CH], float ct, bool is_const,
                                              const uint32_t* __restrict__ S, int64_t plane, const float* __restrict__ rsD,
                                              float* __restrict__ out, const CandSink& sink, BestTrack& bt, bool track,
                                              int y_base, int band)
{
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float rs[8];
        uint32_t sw[8][C];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int y = min(y_first + 8 * half + k, mh - 1);
            const int64_t o = mom_index(x, y - y_base, band);
            rs[k] = __ldg(rsD + o);
#pragma unroll
            for (int c = 0; c < C; ++c) sw[k][c] = __ldg(S + c * plane + o);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int y = y_first + 8 * half + k;
            long long n1 = area * (long long)v[8 * half + k];
#pragma unroll
            for (int c = 0; c < C; ++c) n1 -= (long long)sw[k][c] * sumT[c];
            float r = (float)n1 * rs[k] * ct;
            r = fminf(1.0f, fmaxf(-1.0f, r));
            if (is_const) r = 1.0f;
            if (y < mh) {
                if (STORE) out[(int64_t)y * mw + x] = r;
                if (!STORE && track && r > bt.r) { bt.r = r; bt.idx = (uint32_t)(y * mw + x); }
                if (sink.list && r > sink.thr) {
                    const int slot = atomicAdd(sink.count, 1);
                    if (slot < sink.cap) {
                        DevHit cnd;
                        cnd.tmpl = sink.tmpl; cnd.x = x; cnd.y = y; cnd.w = sink.w; cnd.h = sink.h; cnd.score = r; cnd.seq = 0; cnd.key = 0.f;
                        sink.list[slot] = cnd;
                    }
                }
            }
        }
    }
}

// Stage the image tile rows [y0, y0+R) x bytes [x0, x0 + 16*kb) into smem as [k-block][row][16 B].
// Eight independent 16-byte loads per thread are in flight before the first store.
template <int NT>
__device__ __forceinline__ void stage_image_tile(uint8_t* __restrict__ tile, const uint8_t* __restrict__ img, int64_t pitch,
                                                 int H, int x0, int y0, int R, int kb, int tid)
{
    const int pieces = kb * R;
    for (int base = 0; base < pieces; base += 8 * NT) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * NT + tid;
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            if (idx < pieces) {
                const int c = idx / R, r = idx - c * R;
                const int gy = y0 + r;
                const int64_t gb = (int64_t)x0 + 16 * c;
                if (gy < H && gb + 16 <= pitch) v[u] = __ldg(reinterpret_cast<const uint4*>(img + (int64_t)gy * pitch + gb));
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * NT + tid;
            if (idx < pieces) *reinterpret_cast<uint4*>(tile + (size_t)idx * 16) = v[u];
        }
    }
}

// The same tile through the TMA unit: per 16-byte k-block, `chunks` boxes of `rc` rows (a box holds at most 256 rows; the last
// box is moved up so that it ends with the tile -- overlapping rows are written twice with the same bytes).  R and rc are
// multiples of 8: every destination is 128-byte aligned.  Issued by ONE thread after mbar_expect_tx(bar, tma_tile_bytes(..)).
__device__ __forceinline__ uint32_t tma_tile_bytes(int kb, int rc, int chunks) { return (uint32_t)(kb * chunks * rc * 16); }
__device__ __forceinline__ void tma_stage_tile(uint8_t* tile, const CUtensorMap* tmap, int xb, int y0, int R, int kb, int rc, int chunks,
                                               uint64_t* bar)
{
    for (int c = 0; c < kb; ++c)
        for (int j = 0; j < chunks; ++j) {
            const int r0 = min(j * rc, R - rc);
            tma_load_2d(tile + ((size_t)c * R + r0) * 16, tmap, xb + 16 * c, y0 + r0, bar);
        }
}

struct TcParams {
    const uint8_t* img; int64_t pitch; int H, W;
    const uint8_t* slabs;             // this group's Toeplitz slabs: h slabs of slab_bytes
    int slab_bytes;
    int a_kblk;                       // bytes between consecutive K blocks of a slab (2048 mode A, 256 mode B)
    int nk;                           // K chunks (32 bytes each) per dy
    int ds;                           // slabs (dy values) per ring stage
    int mode;                         // 0 = A (8 templates x 16 x), 1 = B (1 template x 128 x)
    int N;                            // output rows per tile == MMA N == TMEM columns used
    int R;                            // image tile rows = N + h - 1 rounded up to 8 (tc_tile_rows)
    int tma, tma_rc, tma_chunks;      // image tiles through the TMA unit (tensor map = second kernel parameter): rows per box, boxes per k-block
    int h, w, mh, mw;
    int y_base, rows, band_rows;      // this launch covers output rows [y_base, y_base + rows); band_rows: rows of the moment ring's segments
    const TmplMeta* meta; const int32_t* order; int count;
    const uint32_t* S; const float* rsD;   // window moments of this (h, w): [mh][mw]
    float* maps;
    DevHit* cand; int32_t* cand_count; int cand_cap; float cand_thr;   // optional candidate list (nullptr: off)
    unsigned long long* best;         // MODE 3, N_object == 1: per template arg-max key (ordered score << 32 | ~map index), nullptr: off
    int C; int64_t mom_plane;         // channels (1, 3, 4) and the element stride between the per-channel S planes
    int stages, tiles_x, tiles_total; // persistent kernel: slab ring depth, tile grid width, number of tiles
    int ctrl_first;                   // experiment MTM_B200_CTRL_FIRST=1: control warps first, one more (idle) warp per CTA
    long long* prof; int dbg;         // debug only (MTM_B200_PROF / MTM_B200_PDBG): per-CTA role clocks, phase knock-outs
    int method;                       // cv2 method id; != TM_CCOEFF_NORMED takes the float64 epilogue on the summed-area tables
    SatView sat;
    double* acc; double acc_weight; int acc_first;   // MODE 2 (16-bit byte planes): acc = (first ? 0 : acc) + weight * CC
};

// Epilogue of one tile for one of 8 epilogue warps: warp%4 selects the TMEM lane quarter, warp/4 the column half.
template <int MODE>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, uint32_t tmem_d, int x0, int y0, int warp, int lane, int parts, BestTrack& bt)
{
    constexpr bool STORE = MODE != 3;                          // MODE 3 = MODE 0 without the score map (hits only)
    const bool track = MODE == 3 && p.best != nullptr;
    const int m = 32 * (warp & 3) + lane;
    int x, tsel;
    if (p.mode == 0) { tsel = m >> 4; x = x0 + (m & 15); }
    else { tsel = 0; x = x0 + 16 * (7 - (m >> 4)) + (m & 15); }
    // per-lane template geometry: a mode-A group may mix template sizes
    const TmplMeta* tm = (tsel < p.count) ? &p.meta[p.order[tsel]] : nullptr;
    const int t_mh = tm ? min(tm->mh, p.y_base + p.rows) : 0, t_mw = tm ? tm->mw : 0;     // rows end with the map or with the band
    const bool live = tm && (x < t_mw);
    const long long area = tm ? (long long)tm->h * tm->w : 0;
    const long long sumT = tm ? tm->isum[0] : 0;
    long long sumT_c[MTM_MAX_CH];
#pragma unroll
    for (int c = 0; c < MTM_MAX_CH; ++c) sumT_c[c] = tm ? tm->isum[c] : 0;
    const float ct = tm ? tm->inv_sqrt_d2 : 0.f;
    const bool is_const = tm ? (tm->is_const != 0) : false;
    float* out = tm ? p.maps + tm->map_off : nullptr;
    const uint32_t* Sm = tm ? p.S + tm->mom_off : nullptr;
    const float* Rm = tm ? p.rsD + tm->mom_off : nullptr;
    const uint2* SRm = tm ? reinterpret_cast<const uint2*>(p.S) + tm->mom_off : nullptr;   // C == 1: interleaved {S, rsD}
    CandSink sink{is_const ? nullptr : p.cand, p.cand_count, p.cand_cap, p.cand_thr, tm ? p.order[tsel] : 0, tm ? tm->w : 0, tm ? tm->h : 0};
    const float thr_eff = p.cand ? p.cand_thr : 3.0e38f;       // scores never exceed 1: no list, no candidates
    // N is a multiple of 16: the `parts` warps of a lane quarter split the 16-column batches
    const int batches = p.N >> 4, part = warp >> 2;
    const int c_begin = 16 * ((batches * part) / parts), c_end = 16 * ((batches * (part + 1)) / parts);
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
        if (!live || y0 + c0 >= t_mh) continue;
        if (MODE == 2) {                                       // 16-bit path: weighted byte-plane product into the exact map
            epilogue16_accum(v, y0 + c0, t_mh, t_mw, x, p.acc_weight, p.acc_first != 0, p.acc + tm->map_off);
            continue;
        }
        if (MODE == 1) {                                       // compile-time: the default kernel carries none of the float64 code
            if (p.C == 1) epilogue16_generic<1>(v, y0 + c0, t_mh, t_mw, x, p.method, *tm, p.sat, out);
            else if (p.C == 3) epilogue16_generic<3>(v, y0 + c0, t_mh, t_mw, x, p.method, *tm, p.sat, out);
            else epilogue16_generic<4>(v, y0 + c0, t_mh, t_mw, x, p.method, *tm, p.sat, out);
            continue;
        }
        if (p.C == 1) {
            // the prefetched rows of the next batch must exist: +32
            if (!is_const && y0 + c0 + 32 <= t_mh) epilogue16_fast<STORE>(v, y0 + c0, t_mw, x, (uint32_t)area, (uint32_t)sumT, ct, SRm, out, sink, thr_eff, c0 + 16 < c_end, bt, track, p.y_base, p.band_rows);
            else epilogue16<STORE>(v, y0 + c0, t_mh, t_mw, x, (uint32_t)area, (uint32_t)sumT, ct, is_const, SRm, out, sink, c0 + 16 < c_end, bt, track, p.y_base, p.band_rows);
        }
        else if (p.C == 3) epilogue16_mc<3, STORE>(v, y0 + c0, t_mh, t_mw, x, area, sumT_c, ct, is_const, Sm, p.mom_plane, Rm, out, sink, bt, track, p.y_base, p.band_rows);
        else epilogue16_mc<4, STORE>(v, y0 + c0, t_mh, t_mw, x, area, sumT_c, ct, is_const, Sm, p.mom_plane, Rm, out, sink, bt, track, p.y_base, p.band_rows);
    }
}

// ---- single channel, default method (MODE 0 / 3): the epilogue the configs spend their time in -----------------------------
// One batch = 16 accumulator columns (output rows) of a lane.  The ring holds a lane's consecutive rows 128 bytes apart, so the 16
// moment loads of a batch are immediate offsets from one pointer; they are issued one batch AHEAD (two register sets) and the
// tcgen05.ld of the current batch is queued behind them.  Per output: exact 64-bit N1, one conversion, two multiplies; the
// above-threshold test of the 16 outputs is one accumulated predicate -- the (rare) append runs in a separate loop, and MODE 3
// clamps only there.
__device__ __forceinline__ void load_moments16(uint2 (&m)[16], const uint2* __restrict__ sr)
{
#pragma unroll
    for (int k = 0; k < 16; ++k) m[k] = __ldg(sr + 16 * k);
}

template <bool STORE>
__device__ __forceinline__ void epilogue16_c1(const uint32_t (&v)[16], const uint2 (&m)[16], int y_first, int mw, int x, uint32_t area,
                                              uint32_t sumT, float ct, float* __restrict__ out, const CandSink& sink, float thr_any,
                                              BestTrack& bt, bool track)
{
    float r[16];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float rk = n1_to_float(area, v[k], m[k].x, sumT) * __uint_as_float(m[k].y) * ct;
        if (STORE) rk = fminf(1.0f, fmaxf(-1.0f, rk));
        r[k] = rk;
        any = any || (rk > thr_any);
    }
    if (STORE) {
        float* o = out + (int64_t)y_first * mw + x;
#pragma unroll
        for (int k = 0; k < 16; ++k) o[(uint32_t)(k * mw)] = r[k];
    } else if (track) {
        const uint32_t idx0 = (uint32_t)(y_first * mw + x);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float rc = fminf(1.0f, r[k]);                // ties are decided on the clamped score (first occurrence)
            const bool better = rc > bt.r;
            bt.r = better ? rc : bt.r;
            bt.idx = better ? idx0 + (uint32_t)(k * mw) : bt.idx;
        }
    }
    if (any) {                                                     // rare; unrolled so that r[] stays in registers
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float rc = fminf(1.0f, fmaxf(-1.0f, r[k]));
            if (rc > sink.thr) {
                const int slot = atomicAdd(sink.count, 1);
                if (slot < sink.cap) {
                    DevHit c;
                    c.tmpl = sink.tmpl; c.x = x; c.y = y_first + k; c.w = sink.w; c.h = sink.h; c.score = rc; c.seq = 0; c.key = 0.f;
                    sink.list[slot] = c;
                }
            }
        }
    }
}

// Per-lane constants of the single-channel default-method epilogue.  The template a TMEM lane serves, its map geometry and the
// warp's column range do not change from tile to tile of a launch: fetched once per kernel (two dependent global loads and
// ~300 instructions that every epilogue warp used to repeat for every tile).
struct LaneC1 {
    const uint2* SRm; float* out;
    int t_mh, t_mw, xo;                                          // xo: the lane's x offset inside a tile (xo & 15 == lane & 15)
    uint32_t area, sumT; float ct, thr_any;
    int c_begin, c_end;                                          // accumulator columns of this warp
    int tmpl, w, h;
    bool has_tmpl, is_const;
};

__device__ __forceinline__ LaneC1 lane_c1_setup(const TcParams& p, int warp, int lane, int parts)
{
    LaneC1 L;
    const int m_row = 32 * (warp & 3) + lane;
    int tsel;
    if (p.mode == 0) { tsel = m_row >> 4; L.xo = m_row & 15; }
    else { tsel = 0; L.xo = 16 * (7 - (m_row >> 4)) + (m_row & 15); }
    const TmplMeta* tm = (tsel < p.count) ? &p.meta[p.order[tsel]] : nullptr;
    L.has_tmpl = tm != nullptr;
    L.t_mh = tm ? min(tm->mh, p.y_base + p.rows) : 0;           // rows end with the map or with the band
    L.t_mw = tm ? tm->mw : 0;
    L.area = tm ? (uint32_t)(tm->h * tm->w) : 0u;
    L.sumT = tm ? (uint32_t)tm->isum[0] : 0u;                    // <= 255 * 66051 on the tensor path
    L.ct = tm ? tm->inv_sqrt_d2 : 0.f;
    L.is_const = tm ? (tm->is_const != 0) : false;
    L.out = tm ? p.maps + tm->map_off : nullptr;
    L.SRm = tm ? reinterpret_cast<const uint2*>(p.S) + tm->mom_off : nullptr;
    L.tmpl = tm ? p.order[tsel] : 0; L.w = tm ? tm->w : 0; L.h = tm ? tm->h : 0;
    // the accumulated test runs on the UNclamped score in MODE 3: thresholds outside (-1, 1) are decided by the clamped re-test
    L.thr_any = 3.0e38f;                                          // no list: never
    if (p.cand) L.thr_any = p.cand_thr >= 1.0f ? 3.0e38f : (p.cand_thr < -1.0f ? -3.0e38f : p.cand_thr);
    // N is a multiple of 16: the `parts` warps of a lane quarter split the 16-column batches
    const int batches = p.N >> 4, part = warp >> 2;
    L.c_begin = 16 * ((batches * part) / parts); L.c_end = 16 * ((batches * (part + 1)) / parts);
    return L;
}

// The 16 lanes that serve one template share every 128-byte line of the tile-major moment ring (16 x-offsets = one line, rows
// 128 bytes apart): ONE prefetch instruction per warp pulls a whole batch into L1 -- lane k of a half fetches row k.
// `lead`: ring entry of the half's first lane (x & 15 == 0) at tile row 0, nullptr when that lane has no pixel in this tile.
__device__ __forceinline__ void prefetch_batch(const uint2* lead, int c0, int rows_ok, int lane)
{
    const int k = lane & 15;
    if (lead && k < rows_ok) prefetch_l1(lead + 16 * (c0 + k));
}

// PIPE: two register sets of moments (the loads run a whole batch ahead; needs ~170 registers: the 8-epilogue-warp kernel).
// !PIPE: one set, loaded just before the tcgen05.ld of its batch, the next batch prefetched into L1 (128-register kernels).
template <int MODE, bool PIPE>
__device__ __forceinline__ void epilogue_tile_c1(const TcParams& p, const LaneC1& L, uint32_t tmem_d, int x0, int y0, int warp, int lane, BestTrack& bt)
{
    constexpr bool STORE = MODE != 3;
    const bool track = MODE == 3 && p.best != nullptr;
    const int x = x0 + L.xo;
    const bool live = L.has_tmpl && (x < L.t_mw);
    const int t_mh = L.t_mh, t_mw = L.t_mw;
    const uint32_t area = L.area, sumT = L.sumT;
    const float ct = L.ct;
    const bool is_const = L.is_const;
    float* out = L.out;
    const uint2* SRm = L.SRm;
    // (a constant template's TM_CCOEFF_NORMED map is 1 everywhere: peak_local_max of a constant map is empty, so it lists nothing)
    CandSink sink{is_const ? nullptr : p.cand, p.cand_count, p.cand_cap, p.cand_thr, L.tmpl, L.w, L.h};
    const float thr_any = L.thr_any;
    const int c_begin = L.c_begin, c_end = L.c_end;
    const uint2* sr0 = live ? SRm + mom_index(x, y0 - p.y_base, p.band_rows) : nullptr;      // row c of the tile: sr0 + 16 * c
    auto fast_at = [&](int c0) { return live && !is_const && y0 + c0 + 16 <= t_mh; };
    if (!PIPE) {
        // the half's first lane: x - (lane & 15); its pixel exists whenever any pixel of the half does
        const uint2* lead = (L.has_tmpl && !is_const && x - (lane & 15) < t_mw) ? SRm + mom_index(x - (lane & 15), y0 - p.y_base, p.band_rows) : nullptr;
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
            const bool fast = fast_at(c0);
            uint2 m[16];
            if (fast) load_moments16(m, sr0 + 16 * c0);
            if (c0 + 16 < c_end) prefetch_batch(lead, c0 + 16, t_mh - (y0 + c0 + 16), lane);
            uint32_t v[16];
            tmem_ld16(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
            if (live && y0 + c0 < t_mh) {
                if (fast) epilogue16_c1<STORE>(v, m, y0 + c0, t_mw, x, area, sumT, ct, out, sink, thr_any, bt, track);
                else epilogue16<STORE>(v, y0 + c0, t_mh, t_mw, x, area, sumT, ct, is_const, SRm, out, sink, false, bt, track,
                                       p.y_base, p.band_rows);
            }
        }
        return;
    }
    uint2 m[16], mn[16];
    bool fast = c_begin < c_end && fast_at(c_begin);
    if (fast) load_moments16(m, sr0 + 16 * c_begin);
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
        const bool fast_next = c0 + 16 < c_end && fast_at(c0 + 16);
        if (fast_next) load_moments16(mn, sr0 + 16 * (c0 + 16));
        uint32_t v[16];
        tmem_ld16(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0, v);
        if (live && y0 + c0 < t_mh) {
            if (fast) epilogue16_c1<STORE>(v, m, y0 + c0, t_mw, x, area, sumT, ct, out, sink, thr_any, bt, track);
            else epilogue16<STORE>(v, y0 + c0, t_mh, t_mw, x, area, sumT, ct, is_const, SRm, out, sink, false, bt, track,
                                   p.y_base, p.band_rows);           // bottom rows of a map / constant template: bounds-tested form
        }
        fast = fast_next;
        if (fast_next) {
#pragma unroll
            for (int k = 0; k < 16; ++k) m[k] = mn[k];
        }
    }
}

// Pulls the window moments of a warp's first batch of the tile into L1 while the accumulator is still being computed.
__device__ __forceinline__ void epilogue_prefetch_first(const TcParams& p, const LaneC1& L, int x0, int y0, int lane)
{
    const int x_lead = x0 + L.xo - (lane & 15);
    if (!L.has_tmpl || L.is_const || x_lead >= L.t_mw) return;
    prefetch_batch(L.SRm + mom_index(x_lead, y0 - p.y_base, p.band_rows), L.c_begin, L.t_mh - (y0 + L.c_begin), lane);
}

// MODE 3, N_object == 1: the lanes of a warp that serve the same template (mode A: 16, mode B: 32) reduce their best pixel and
// one of them raises the template's arg-max key (larger score wins, then the smaller map index = cv2.minMaxLoc's first occurrence).
__device__ __forceinline__ void flush_best(const TcParams& p, const BestTrack& bt, int warp, int lane)
{
    const int m = 32 * (warp & 3) + lane;
    const int tsel = p.mode == 0 ? (m >> 4) : 0;
    unsigned long long key = bt.r > -2.0f ? (((unsigned long long)ordered_f32(bt.r) << 32) | (unsigned long long)(0xFFFFFFFFu - bt.idx)) : 0ull;
    const int span = p.mode == 0 ? 8 : 16;
    for (int d = span; d; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, d);
        key = o > key ? o : key;
    }
    if (tsel < p.count && key && (lane & (p.mode == 0 ? 15 : 31)) == 0) atomicMax(&p.best[p.order[tsel]], key);
}

// One CTA = one output tile.  Warp 0: slab producer, warp 1: MMA issuer, then all 8 warps: epilogue.
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 2)
ncc_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kb_img = 2 * p.nk;                               // 16-byte K blocks of the image tile
    const uint32_t tile_bytes = (uint32_t)kb_img * p.R * 16;
    const uint32_t stage_bytes = (uint32_t)p.ds * p.slab_bytes;
    uint8_t* tile = smem;
    uint8_t* ring = smem + ((tile_bytes + 127) & ~127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + TC_STAGES * stage_bytes);
    uint64_t* full = bars;                                     // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;                        // [TC_STAGES]
    uint64_t* accum = bars + 2 * TC_STAGES;                    // [1]
    uint64_t* tile_bar = accum + 1;                            // [1] image tile landed (TMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2);

    const int xw = p.mode == 0 ? 16 : 128;
    const int x0 = blockIdx.x * xw, y0 = p.y_base + blockIdx.y * p.N;
    uint32_t tmem_cols = 32;                                    // tcgen05.alloc: power of two >= 32
    while (tmem_cols < (uint32_t)p.N) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum, 1);
        mbar_init(tile_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);

    // ---- stage the image tile: rows [y0, y0+R) x bytes [x0, x0 + 32*nk), layout [k-block][row][16 B]
    if (p.tma) {                                               // TMA unit: one thread issues the boxes, everybody waits for the bytes
        __syncthreads();                                       // tile_bar initialised
        if (tid == 0) {
            mbar_expect_tx(tile_bar, tma_tile_bytes(kb_img, p.tma_rc, p.tma_chunks));
            tma_stage_tile(tile, &tmap, x0 * p.C, y0, p.R, kb_img, p.tma_rc, p.tma_chunks, tile_bar);
        }
        mbar_wait(tile_bar, 0);
    } else {
        stage_image_tile<TC_THREADS>(tile, p.img, p.pitch, p.H, x0 * p.C, y0, p.R, kb_img, tid);
        fence_async_smem();                                    // generic-proxy writes -> visible to the MMA (async proxy)
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    const int n_iters = (p.h + p.ds - 1) / p.ds;
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % TC_STAGES;
                if (it >= TC_STAGES) mbar_wait(&empty[s], ((it / TC_STAGES) - 1) & 1);
                const int rows = min(p.ds, p.h - it * p.ds);
                const uint32_t bytes = (uint32_t)rows * p.slab_bytes;
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(ring + (size_t)s * stage_bytes, p.slabs + (size_t)it * stage_bytes, bytes, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);   // S32 accum, u8 x u8, K-major
            const uint32_t tile_addr = smem_u32(tile), ring_addr = smem_u32(ring);
            const uint32_t lbo_b = (uint32_t)p.R * 16;
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % TC_STAGES;
                mbar_wait(&full[s], (it / TC_STAGES) & 1);
                tc_fence_after();
                const int rows = min(p.ds, p.h - it * p.ds);
                for (int d = 0; d < rows; ++d) {
                    const int dy = it * p.ds + d;
                    const uint32_t a_base = ring_addr + s * stage_bytes + d * p.slab_bytes;
                    const uint32_t b_base = tile_addr + dy * 16;
                    for (int i = 0; i < p.nk; ++i) {
                        const uint64_t ad = umma_desc(a_base + 2 * i * p.a_kblk, (uint32_t)p.a_kblk, 128);
                        const uint64_t bd = umma_desc(b_base + 2 * i * lbo_b, lbo_b, 128);
                        umma_i8(tmem_d, ad, bd, idesc, (dy | i) != 0);
                    }
                }
                umma_commit(&empty[s]);                        // frees the ring slot when these MMAs retire
            }
            umma_commit(accum);                                // all MMAs of the tile retired -> epilogue may read TMEM
        }
        __syncwarp();
    }

    // ---- epilogue: all 8 warps.  warp%4 selects the TMEM lane quarter, warp/4 the column half.
    // Only the MMA warp polls the accumulator barrier; everybody else blocks in bar.sync (no issue
    // slots stolen from the producer / MMA lanes while the main loop runs).
    if (warp == 1) { if (lane == 0) mbar_wait(accum, 0); __syncwarp(); }
    __syncthreads();
    tc_fence_after();
    BestTrack bt{-3.0e38f, 0u};
    if ((MODE == 0 || MODE == 3) && p.C == 1) epilogue_tile_c1<(MODE == 3 ? 3 : 0), false>(p, lane_c1_setup(p, warp, lane, 2), tmem_d, x0, y0, warp, lane, bt);
    else epilogue_tile<MODE>(p, tmem_d, x0, y0, warp, lane, 2, bt);
    if (MODE == 3 && p.best) flush_best(p, bt, warp, lane);
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_d, tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// Persistent, warp-specialised form of the same tile computation: one CTA per SM walks the tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...  Two image-tile buffers in shared memory and two
// accumulators in TMEM turn the three phases of a tile (stage the image rows / MMA / epilogue) into
// a pipeline, so the tensor pipe works on tile i+1 while tile i is normalised and stored and the
// rows of tile i+2 are fetched.
//   warps 0 .. EW-1   epilogue (TMEM -> registers -> score map), warp%4 = TMEM lane quarter; EW = 8 or 12
//   warp  EW          Toeplitz slab producer (cp.async.bulk into the ring, continuous over tiles)
//   warp  EW+1        MMA issuer (one elected lane)
//   warps EW+2, EW+3  image tile stagers (global -> registers -> shared memory, [k-block][row][16 B])
// EW = 12 serves small templates, whose tiles spend longer in the epilogue than in the MMAs.
constexpr int TCP_MAX_STAGES = 8;
constexpr double TCP_EPI_CLK_PER_ROW = 70.0;    // measured (role clocks, MMAs knocked out): epilogue clocks per output row of a tile, 8 warps, pipelined loads
constexpr double TCP_EPI_CLK_PER_ROW_12 = 55.0; // the same with 12 epilogue warps (128 registers, prefetch instead of the second register set)
constexpr int TCP_STAGERS = 32;                 // register staging (fallback when no tensor map can be made): one stager warp
constexpr size_t TCP_SMEM_SOFT = 188 * 1024;     // preferred ceiling of the persistent kernel's shared memory (see launch_ncc_tc)

// MMAs of `rows` consecutive template rows: NK K-chunks each.  Only the low descriptor words move
// (one 16-byte unit per image row, one slab per template row); they live in uniform registers.
template <int NK>
__device__ __forceinline__ void issue_rows(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                           uint32_t a_kstep, uint32_t b_kstep, uint32_t slab_step, int rows, uint32_t idesc,
                                           uint32_t accumulate, uint32_t b_row_step)
{
    uint32_t al[NK], bl[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) { al[k] = a_lo + k * a_kstep; bl[k] = b_lo + k * b_kstep; }
    for (int d = 0; d < rows; ++d) {
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            umma_i8(tmem_d, ((uint64_t)a_hi << 32) | al[k], ((uint64_t)b_hi << 32) | bl[k], idesc, accumulate);
            accumulate = 1;
            al[k] += slab_step; bl[k] += b_row_step;
        }
    }
}

// b_row_step: 1 (one 16-byte image row per template row); 0 only in the profiling knock-out that keeps B in place
__device__ __forceinline__ void issue_rows_any(int nk, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi,
                                               uint32_t a_kstep, uint32_t b_kstep, uint32_t slab_step, int rows, uint32_t idesc,
                                               uint32_t accumulate, uint32_t b_row_step)
{
    switch (nk) {
        case 1: issue_rows<1>(tmem_d, a_lo, b_lo, a_hi, b_hi, a_kstep, b_kstep, slab_step, rows, idesc, accumulate, b_row_step); break;
        case 2: issue_rows<2>(tmem_d, a_lo, b_lo, a_hi, b_hi, a_kstep, b_kstep, slab_step, rows, idesc, accumulate, b_row_step); break;
        case 3: issue_rows<3>(tmem_d, a_lo, b_lo, a_hi, b_hi, a_kstep, b_kstep, slab_step, rows, idesc, accumulate, b_row_step); break;
        case 4: issue_rows<4>(tmem_d, a_lo, b_lo, a_hi, b_hi, a_kstep, b_kstep, slab_step, rows, idesc, accumulate, b_row_step); break;
        case 5: issue_rows<5>(tmem_d, a_lo, b_lo, a_hi, b_hi, a_kstep, b_kstep, slab_step, rows, idesc, accumulate, b_row_step); break;
        case 6: issue_rows<6>(tmem_d, a_lo, b_lo, a_hi, b_hi, a_kstep, b_kstep, slab_step, rows, idesc, accumulate, b_row_step); break;
        default:
            for (int d = 0; d < rows; ++d) {
                for (int k = 0; k < nk; ++k) {
                    umma_i8(tmem_d, ((uint64_t)a_hi << 32) | (a_lo + k * a_kstep), ((uint64_t)b_hi << 32) | (b_lo + k * b_kstep), idesc, accumulate);
                    accumulate = 1;
                }
                a_lo += slab_step; b_lo += b_row_step;
            }
    }
}

// The MMA issuer of the persistent kernel.  NK = K chunks per template row (0: run-time p.nk), chosen ONCE per kernel so that
// no jump table sits in the per-stage path.  What the issuing thread does between two groups of MMAs decides whether the tensor
// pipe runs dry: the pipe holds about three queued MMAs (measured with i8_peak_kernel: a passing mbarrier wait + fence + commit
// after every 6 MMAs costs nothing at N = 208 but 30 % at N = 144, 43 % at N = 128), so every barrier is POLLED ONE STEP AHEAD --
// the try_wait of the next ring stage (next tile) is issued before the MMAs of the current one and its answer is read after
// them; only a poll that failed falls back to the blocking wait.
template <bool PROF, int NK>
__device__ __forceinline__ void mma_issuer_role(const TcParams& p, uint64_t* full, uint64_t* empty, uint64_t* tile_full, uint64_t* tile_empty,
                                                uint64_t* acc_full, uint64_t* acc_empty, uint32_t tile_addr0, uint32_t tile_bytes,
                                                uint32_t ring_addr, uint32_t stage_bytes, uint32_t tmem_base, uint32_t acc_stride,
                                                int my_tiles, int lane)
{
    const uint32_t idesc = (2u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);   // S32 accum, u8 x u8, K-major
    const uint32_t lbo_b = (uint32_t)p.R * 16;
    const int ds = p.ds, h = p.h, nk = NK ? NK : p.nk, stages = p.stages;
    const uint64_t a_desc0 = umma_desc(ring_addr, (uint32_t)p.a_kblk, 128);
    const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), a_lo0 = (uint32_t)a_desc0;
    const uint32_t a_kstep = (uint32_t)(2 * p.a_kblk) >> 4, b_kstep = (2 * lbo_b) >> 4;     // descriptor address unit: 16 B
    const uint32_t slab_step = (uint32_t)p.slab_bytes >> 4, stage_step = stage_bytes >> 4;
    // knock-outs (PROF): 2 = no MMAs, 32 = B stays on the tile's first rows (no 16-byte row shift), 64 = A always from ring slot 0
    const bool no_mma = PROF && (p.dbg & 2), b_fixed = PROF && (p.dbg & 32), a_fixed = PROF && (p.dbg & 64);
    int s = 0;
    uint32_t ph = 0;
    long long w_tile = 0, w_acc = 0, w_full = 0;
    const long long m_begin = PROF ? clock64() : 0;
    bool slab_ready = my_tiles > 0 && mbar_try_wait_once(&full[0], 0);
    bool tile_ready = my_tiles > 0 && mbar_try_wait_once(&tile_full[0], 0);
    bool acc_ready = true;                                      // fresh barrier
    for (int i = 0; i < my_tiles; ++i) {
        const int b = i & 1, u = i >> 1;
        const long long c0 = PROF ? clock64() : 0;
        if (!tile_ready) mbar_wait(&tile_full[b], u & 1);
        const long long c1 = PROF ? clock64() : 0;
        if (!acc_ready) mbar_wait(&acc_empty[b], (u & 1) ^ 1);  // fresh barrier: parity 1 passes
        if (PROF) { w_tile += c1 - c0; w_acc += clock64() - c1; }
        tc_fence_after();
        const uint64_t b_desc0 = umma_desc(tile_addr0 + (uint32_t)b * tile_bytes, lbo_b, 128);
        const uint32_t b_hi = (uint32_t)(b_desc0 >> 32), b_lo0 = (uint32_t)b_desc0;
        const uint32_t tmem_d = tmem_base + (uint32_t)b * acc_stride;
        for (int dy0 = 0; dy0 < h; dy0 += ds) {
            const long long c2 = PROF ? clock64() : 0;
            if (!slab_ready) mbar_wait(&full[s], ph);
            if (PROF) w_full += clock64() - c2;
            tc_fence_after();
            // poll the barriers of the NEXT step now, read the answers after this stage's MMAs have been queued
            int sn = s + 1;
            uint32_t phn = ph;
            if (sn == stages) { sn = 0; phn ^= 1u; }
            const bool last_stage = dy0 + ds >= h;
            const bool more = !last_stage || i + 1 < my_tiles;
            const bool slab_next = more && mbar_try_wait_once(&full[sn], phn);
            bool tile_next = false, acc_next = false;
            if (last_stage && i + 1 < my_tiles) {
                const int bn = (i + 1) & 1, un = (i + 1) >> 1;
                tile_next = mbar_try_wait_once(&tile_full[bn], un & 1);
                acc_next = mbar_try_wait_once(&acc_empty[bn], (un & 1) ^ 1);
            }
            if (elect_one()) {
                if (!no_mma) {
                    const uint32_t a_lo = a_lo0 + (a_fixed ? 0u : (uint32_t)s * stage_step), b_lo = b_lo0 + (b_fixed ? 0u : (uint32_t)dy0);
                    if (NK) issue_rows<NK ? NK : 1>(tmem_d, a_lo, b_lo, a_hi, b_hi, a_kstep, b_kstep, a_fixed ? 0u : slab_step, min(ds, h - dy0), idesc,
                                                    dy0 != 0, b_fixed ? 0u : 1u);
                    else issue_rows_any(nk, tmem_d, a_lo, b_lo, a_hi, b_hi, a_kstep, b_kstep, a_fixed ? 0u : slab_step, min(ds, h - dy0), idesc,
                                        dy0 != 0, b_fixed ? 0u : 1u);
                }
                umma_commit(&empty[s]);                        // frees the ring slot when these MMAs retire
            }
            s = sn; ph = phn;
            slab_ready = slab_next;
            if (last_stage) { tile_ready = tile_next; acc_ready = acc_next; }
        }
        if (elect_one()) {
            umma_commit(&acc_full[b]);                         // epilogue may read this accumulator
            umma_commit(&tile_empty[b]);                       // the loader may overwrite this image tile
        }
    }
    if (PROF && lane == 0) {
        long long* q = p.prof + 16 * blockIdx.x;
        q[0] = w_tile; q[1] = w_acc; q[2] = w_full; q[3] = clock64() - m_begin;
    }
    __syncwarp();
}

template <bool PROF, int EW, int MODE, bool PIPE>
__device__ __forceinline__ void ncc_tc_persist_body(const TcParams& p, const CUtensorMap& tmap)
{
    constexpr int TCP_THREADS = 32 * (EW + 3);
    constexpr int TCP_EPI_THREADS = 32 * EW;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // Warp roles.  Default: warps 0 .. EW-1 epilogue, EW producer, EW+1 MMA issuer, EW+2 loader.  p.ctrl_first (experiment): the three
    // control warps come FIRST (0 producer, 1 MMA issuer, 2 loader, 3 idle) and the epilogue warps are 4 .. EW+3 -- the TMEM lane
    // quarter of an epilogue warp stays warp % 4 either way.
    const int w_prod = p.ctrl_first ? 0 : EW, w_mma = p.ctrl_first ? 1 : EW + 1, w_load = p.ctrl_first ? 2 : EW + 2;
    const int ewarp = p.ctrl_first ? warp - 4 : warp;           // index among the epilogue warps (valid for those only)
    const bool is_epi = p.ctrl_first ? warp >= 4 : warp < EW;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* full = bars;                                      // [stages]  slab bytes landed
    uint64_t* empty = bars + TCP_MAX_STAGES;                    // [stages]  tcgen05.commit
    uint64_t* tile_full = bars + 2 * TCP_MAX_STAGES;            // [2] image tile staged (64 stager arrivals)
    uint64_t* tile_empty = tile_full + 2;                       // [2] tcgen05.commit: MMAs reading the tile retired
    uint64_t* acc_full = tile_full + 4;                         // [2] tcgen05.commit: accumulator complete
    uint64_t* acc_empty = tile_full + 6;                        // [2] 256 epilogue arrivals: accumulator read out
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_full + 8);

    const int kb_img = 2 * p.nk;
    const uint32_t tile_bytes = ((uint32_t)kb_img * p.R * 16 + 127) & ~127u;
    const uint32_t stage_bytes = (uint32_t)p.ds * p.slab_bytes;
    uint8_t* tiles = smem + 256;
    uint8_t* ring = tiles + 2 * tile_bytes;
    const int xw = p.mode == 0 ? 16 : 128;
    const int my_tiles = (p.tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    uint32_t tmem_cols = 32;                                    // two accumulators; tcgen05.alloc wants a power of two
    while (tmem_cols < 2u * (uint32_t)p.N) tmem_cols <<= 1;
    const uint32_t acc_stride = tmem_cols >> 1;

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tile_full[b], p.tma ? 1 : TCP_STAGERS); mbar_init(&tile_empty[b], 1);
            mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], TCP_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == w_prod) tmem_alloc(tmem_slot, tmem_cols);
    // register staging (p.tma == 0): the first image tile is staged by the whole CTA (nothing else to do yet); the stagers
    // take over from the second.  With the TMA unit the loader warp issues every tile, the first one included.
    if (!p.tma && tid < TCP_THREADS) stage_image_tile<TCP_THREADS>(tiles, p.img, p.pitch, p.H, ((int)blockIdx.x % p.tiles_x) * xw * p.C,
                                                                   p.y_base + ((int)blockIdx.x / p.tiles_x) * p.N, p.R, kb_img, tid);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == w_prod) {
        // ===== slab producer: the slab sequence of a tile repeats for every tile; the ring position runs on.
        // The whole warp walks the loop (warp-uniform control flow, uniform registers); one elected lane issues.
        int s = 0;
        uint32_t ph = 1;                                        // parity of the previous use of slot s (fresh barrier: passes)
        long long w_empty = 0;
        const int ds = p.ds, h = p.h, stages = p.stages, slab_bytes = p.slab_bytes;
        const uint8_t* slabs = p.slabs;
        for (int i = 0; i < my_tiles; ++i) {
            const uint8_t* src = slabs;
            for (int dy = 0; dy < h; dy += ds) {
                const long long c0 = PROF ? clock64() : 0;
                mbar_wait(&empty[s], ph);
                if (PROF) w_empty += clock64() - c0;
                const uint32_t bytes = (uint32_t)min(ds, h - dy) * (uint32_t)slab_bytes;
                if (elect_one()) {
                    if (PROF && (p.dbg & 8)) mbar_arrive(&full[s]);          // knock-out: no slab bytes move (what the L2 -> SM stream costs)
                    else {
                        mbar_expect_tx(&full[s], bytes);
                        bulk_g2s(ring + (size_t)s * stage_bytes, (PROF && (p.dbg & 16)) ? slabs : src, bytes, &full[s]);   // 16: always the same 24 KB
                    }
                }
                src += stage_bytes;
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
        if (PROF && lane == 0) p.prof[16 * blockIdx.x + 8] = w_empty;
        __syncwarp();
    } else if (warp == w_mma) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues (mma_issuer_role below)
        const uint32_t tile_addr0 = smem_u32(tiles), ring_addr = smem_u32(ring);
        switch (p.nk) {
            case 1: mma_issuer_role<PROF, 1>(p, full, empty, tile_full, tile_empty, acc_full, acc_empty, tile_addr0, tile_bytes, ring_addr, stage_bytes, tmem_base, acc_stride, my_tiles, lane); break;
            case 2: mma_issuer_role<PROF, 2>(p, full, empty, tile_full, tile_empty, acc_full, acc_empty, tile_addr0, tile_bytes, ring_addr, stage_bytes, tmem_base, acc_stride, my_tiles, lane); break;
            case 3: mma_issuer_role<PROF, 3>(p, full, empty, tile_full, tile_empty, acc_full, acc_empty, tile_addr0, tile_bytes, ring_addr, stage_bytes, tmem_base, acc_stride, my_tiles, lane); break;
            case 4: mma_issuer_role<PROF, 4>(p, full, empty, tile_full, tile_empty, acc_full, acc_empty, tile_addr0, tile_bytes, ring_addr, stage_bytes, tmem_base, acc_stride, my_tiles, lane); break;
            case 5: mma_issuer_role<PROF, 5>(p, full, empty, tile_full, tile_empty, acc_full, acc_empty, tile_addr0, tile_bytes, ring_addr, stage_bytes, tmem_base, acc_stride, my_tiles, lane); break;
            case 6: mma_issuer_role<PROF, 6>(p, full, empty, tile_full, tile_empty, acc_full, acc_empty, tile_addr0, tile_bytes, ring_addr, stage_bytes, tmem_base, acc_stride, my_tiles, lane); break;
            default: mma_issuer_role<PROF, 0>(p, full, empty, tile_full, tile_empty, acc_full, acc_empty, tile_addr0, tile_bytes, ring_addr, stage_bytes, tmem_base, acc_stride, my_tiles, lane); break;
        }
    } else if (!is_epi && p.tma) {
        // ===== image tile loader: one warp, one elected lane per tile feeds the TMA unit (launched with EW + 3 warps)
        if (warp == w_load) {
            const uint32_t bytes = tma_tile_bytes(kb_img, p.tma_rc, p.tma_chunks);
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1, u = i >> 1;
                const int ti = (int)blockIdx.x + i * (int)gridDim.x;
                const int x0 = (ti % p.tiles_x) * xw, y0 = p.y_base + (ti / p.tiles_x) * p.N;
                mbar_wait(&tile_empty[b], (u & 1) ^ 1);        // fresh barrier: parity 1 passes
                if (elect_one()) {
                    mbar_expect_tx(&tile_full[b], bytes);
                    tma_stage_tile(tiles + (size_t)b * tile_bytes, &tmap, x0 * p.C, y0, p.R, kb_img, p.tma_rc, p.tma_chunks, &tile_full[b]);
                }
                __syncwarp();
            }
        }
    } else if (!is_epi && warp == w_load) {
        // ===== image tile stager (register staging) =====
        const int t = tid - 32 * w_load;
        long long w_te = 0, w_work = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1, u = i >> 1;
            const int ti = (int)blockIdx.x + i * (int)gridDim.x;
            const int x0 = (ti % p.tiles_x) * xw, y0 = p.y_base + (ti / p.tiles_x) * p.N;
            const long long c0 = PROF ? clock64() : 0;
            mbar_wait(&tile_empty[b], (u & 1) ^ 1);
            const long long c1 = PROF ? clock64() : 0;
            if (i > 0 && (!PROF || !(p.dbg & 4)))
                stage_image_tile<TCP_STAGERS>(tiles + (size_t)b * tile_bytes, p.img, p.pitch, p.H, x0 * p.C, y0, p.R, kb_img, t);
            fence_async_smem();                                // generic-proxy writes -> visible to the MMA (async proxy)
            mbar_arrive(&tile_full[b]);
            if (PROF) { w_te += c1 - c0; w_work += clock64() - c1; }
        }
        if (PROF && t == 0) { p.prof[16 * blockIdx.x + 6] = w_te; p.prof[16 * blockIdx.x + 7] = w_work; }
    } else if (is_epi) {
        // ===== epilogue warps =====
        long long w_af = 0, w_epi = 0;
        BestTrack bt{-3.0e38f, 0u};
        const bool c1 = (MODE == 0 || MODE == 3) && p.C == 1;
        LaneC1 L{};
        if (c1) L = lane_c1_setup(p, ewarp, lane, EW / 4);
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1, u = i >> 1;
            const int ti = (int)blockIdx.x + i * (int)gridDim.x;
            const int x0 = (ti % p.tiles_x) * xw, y0 = p.y_base + (ti / p.tiles_x) * p.N;
            if (c1) epilogue_prefetch_first(p, L, x0, y0, lane);
            const long long c0 = PROF ? clock64() : 0;
            mbar_wait(&acc_full[b], u & 1);
            const long long c1 = PROF ? clock64() : 0;
            tc_fence_after();
            if (!PROF || !(p.dbg & 1)) {
                if (c1) epilogue_tile_c1<(MODE == 3 ? 3 : 0), PIPE>(p, L, tmem_base + (uint32_t)b * acc_stride, x0, y0, ewarp, lane, bt);
                else epilogue_tile<MODE>(p, tmem_base + (uint32_t)b * acc_stride, x0, y0, ewarp, lane, EW / 4, bt);
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[b]);
            if (PROF) { w_af += c1 - c0; w_epi += clock64() - c1; }
        }
        if (MODE == 3 && p.best) flush_best(p, bt, ewarp, lane);
        if (PROF && ewarp == 0 && lane == 0) { p.prof[16 * blockIdx.x + 4] = w_af; p.prof[16 * blockIdx.x + 5] = w_epi; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == w_prod) tmem_dealloc(tmem_base, tmem_cols);
}

// EW = 8: 352 threads, 168 registers (the pipelined epilogue holds two batches of moments); EW = 12: 480 threads, 128.
template <bool PROF, int EW, int MODE>
__global__ void __launch_bounds__(32 * (EW + 4), 1)
ncc_tc_persist_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmap)
{
    ncc_tc_persist_body<PROF, EW, MODE, EW == 8>(p, tmap);
}

// The 8-epilogue-warp kernel held to 128 registers (one set of moments, no load pipelining): 11 warps x 4096 registers leave
// exactly the 20 480 registers a box-moment CTA needs, so that the moment kernel of ANOTHER stream can run beside it (throughput
// mode: several contexts per GPU).  MTM_B200_LEAN=1.
template <int MODE>
__global__ void __maxnreg__(128)                    // (cannot be combined with __launch_bounds__; launched with 352 threads)
ncc_tc_persist_lean_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmap)
{
    ncc_tc_persist_body<false, 8, MODE, false>(p, tmap);
}

// Expands the templates of one group into Toeplitz slabs (see the header comment).
__global__ void toeplitz_prep_kernel(const uint8_t* __restrict__ tmpl, const TmplMeta* __restrict__ meta,
                                     const int32_t* __restrict__ order, int count, int mode, int h, int w,
                                     int nk, int slab_bytes, int C, uint8_t* __restrict__ slabs,
                                     const TmplPix8* __restrict__ pix8)      // optional: byte-plane arena geometry (16-bit templates)
{
    const int pieces_per_slab = slab_bytes / 16;
    const int64_t total = (int64_t)h * pieces_per_slab;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int dy = (int)(idx / pieces_per_slab);
        const int pc = (int)(idx - (int64_t)dy * pieces_per_slab);
        int t, r, ubase;                                   // ubase: template column of byte 0 of this piece
        if (mode == 0) { const int c = pc >> 7; t = (pc >> 4) & 7; r = pc & 15; ubase = 16 * c - C * r; }   // x-step = C bytes
        else { const int b = pc >> 4; t = 0; r = pc & 15; ubase = 16 * (b - 7) - r; }
        uint8_t bytes[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) bytes[k] = 0;
        if (t < count) {
            const TmplMeta& tm = meta[order[t]];                 // members smaller than the group are zero padded
            const int64_t t_off = pix8 ? pix8[order[t]].off : tm.pix_off;
            const int t_wp = pix8 ? pix8[order[t]].wp : tm.wp;
            const uint8_t* row = tmpl + t_off + (int64_t)dy * t_wp;
            if (dy < tm.h) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int col = ubase + k;
                    if (col >= 0 && col < tm.w * C) bytes[k] = row[col];
                }
            }
        }
        *reinterpret_cast<uint4*>(slabs + (int64_t)dy * slab_bytes + (int64_t)pc * 16) = *reinterpret_cast<const uint4*>(bytes);
    }
}

// S = window sum, rsD = rsqrt(A*Q - S^2) (0 for an exactly flat window), per window position, for
// every distinct template size in one launch (blockIdx.y = size).
// STREAM (experiment knob): write the maps with the evict-first hint (st.global.cs) so that a many-size sweep (C5: 4 GB of
// moments) cannot push the summed-area tables it keeps re-reading out of L2.
template <bool STREAM>
__global__ void window_moments_kernel(SatView sat, const uint32_t* __restrict__ sat_q32, const SizeDesc* __restrict__ sizes,
                                      uint32_t* __restrict__ S, float* __restrict__ rsD, int C, int64_t mom_plane)
{
    const SizeDesc sd = sizes[blockIdx.y];
    const int64_t n = (int64_t)sd.mh * sd.mw;
    const unsigned long long area = (unsigned long long)sd.h * sd.w;
    for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(pos / sd.mw), x = (int)(pos - (int64_t)y * sd.mw);
        const int64_t idx = mom_index(x, y, sd.band);          // whole maps: one band that starts at row 0
        // window sums of squares < 2^32 on the tensor path (h*w*C <= 66051): the 32-bit wrap-around table is exact
        const unsigned long long q = sat_window_s(sat_q32, sat.pitch, y, x, sd.h, sd.w);
        unsigned long long d1 = area * q;
        uint32_t s0 = 0;
        for (int c = 0; c < C; ++c) {
            const uint32_t s = sat_window_s(sat.s + c * sat.plane, sat.pitch, y, x, sd.h, sd.w);
            d1 -= (unsigned long long)s * s;
            if (C > 1) { if (STREAM) __stcs(S + c * mom_plane + sd.off + idx, s); else S[c * mom_plane + sd.off + idx] = s; }
            s0 = s;
        }
        const float rs = d1 ? rsqrtf((float)d1) : 0.0f;
        if (C > 1) { if (STREAM) __stcs(rsD + sd.off + idx, rs); else rsD[sd.off + idx] = rs; }
        else if (STREAM) __stcs(reinterpret_cast<uint2*>(S) + sd.off + idx, make_uint2(s0, __float_as_uint(rs)));
        else reinterpret_cast<uint2*>(S)[sd.off + idx] = make_uint2(s0, __float_as_uint(rs));   // one 8-byte load per pixel in the epilogue
    }
}

// Row-walking form of window_moments_kernel (experiment knob MTM_B200_MOM_ROWS=1; same arithmetic per position, hence
// bit-identical maps): blockIdx.z = size, blockIdx.y walks the rows, threads walk x -- no 64-bit division per position and
// the eight corner addresses are a row pointer plus x.  The grid-stride kernel above executes ~110 instructions per
// position and runs at ~0.75 positions per clock per SM (C5: 506 M positions in 2.4 ms), which is an issue-rate bound,
// not a memory one (evict-first stores changed nothing, profiles/README.md).  NOT YET MEASURED ON THE GPU.
template <int C>
__global__ void __launch_bounds__(256)
window_moments_rows_kernel(SatView sat, const uint32_t* __restrict__ sat_q32, const SizeDesc* __restrict__ sizes,
                           uint32_t* __restrict__ S, float* __restrict__ rsD, int64_t mom_plane)
{
    const SizeDesc sd = sizes[blockIdx.z];
    const uint32_t area = (uint32_t)sd.h * (uint32_t)sd.w;                 // <= 66051 on the tensor path
    const int x_first = blockIdx.x * blockDim.x + threadIdx.x, x_step = gridDim.x * blockDim.x;
    const int64_t down = (int64_t)sd.h * sat.pitch;                         // from the window's top SAT row to its bottom one
    for (int y = blockIdx.y; y < sd.mh; y += gridDim.y) {
        const uint32_t* qa = sat_q32 + (int64_t)y * sat.pitch;
        const uint32_t* sa = sat.s + (int64_t)y * sat.pitch;
        uint32_t* out_s = S + sd.off;
        float* out_r = rsD + sd.off;
        uint2* out_sr = reinterpret_cast<uint2*>(S) + sd.off;
        for (int x = x_first; x < sd.mw; x += x_step) {
            const int64_t idx = mom_index(x, y, sd.band);
            const uint32_t* q0 = qa + x;
            const uint32_t q = q0[down + sd.w] - q0[sd.w] - q0[down] + q0[0];      // modulo 2^32, exact (window sums < 2^32)
            unsigned long long d1 = (unsigned long long)area * q;
            uint32_t s0 = 0;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const uint32_t* p0 = sa + c * sat.plane + x;
                const uint32_t s = p0[down + sd.w] - p0[sd.w] - p0[down] + p0[0];
                d1 -= (unsigned long long)s * s;
                if (C > 1) out_s[c * mom_plane + idx] = s;
                s0 = s;
            }
            const float rs = d1 ? rsqrtf((float)d1) : 0.0f;
            if (C > 1) out_r[idx] = rs;
            else out_sr[idx] = make_uint2(s0, __float_as_uint(rs));
        }
    }
}

// Measurement helper (mtm_measure_i8_peak): the rate of the pipe the numerator kernels use.  One CTA per SM issues `iters`
// back-to-back tcgen05.mma kind::i8 M128 x N x K32 from operands resident in shared memory (zeros; two accumulators in
// turn), no loads, no epilogue: time / (grid * iters * 128 * N * 32) is the dense u8 MAC rate bench.py quotes rooflines against.
__global__ void __launch_bounds__(128, 1)
i8_peak_kernel(int iters, int n, int variant)
{
    // variant (MTM_B200_PEAK_VARIANT, experiments): bit 0 = the B descriptor moves one 16-byte row per MMA (64 positions, the
    // numerator kernel's row shift), bit 1 = A walks over 8 slabs of 4 KB, bit 2 = a tcgen05.commit after every 6 MMAs, bit 3 = pseudo-random operand bytes instead of zeros,
    // bit 4 = tcgen05.fence::after_thread_sync and bit 5 = a (passing) mbarrier wait after every 6 MMAs
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar2 = bar + 1;
    uint64_t* bar3 = bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32);
    uint8_t* A = smem + 1024;                                   // 8 x [2 k-blocks][128 rows][16 B]
    uint8_t* B = A + 8 * 4096;                                  // [2 k-blocks][n + 64 rows][16 B]
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rows_b = n + 64;
    for (int i = tid; i < (8 * 4096 + 32 * rows_b) / 16; i += blockDim.x) {
        uint4 fill = make_uint4(0u, 0u, 0u, 0u);
        if (variant & 8) {                                       // pseudo-random operand bytes (data-dependent power)
            uint32_t x = (uint32_t)i * 2654435761u + blockIdx.x * 40503u + 12345u;
            x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
            fill = make_uint4(x, x * 3266489917u, x * 668265263u + 1u, x * 374761393u + 7u);
        }
        reinterpret_cast<uint4*>(A)[i] = fill;
    }
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_init(bar3, 1); }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t a_desc = umma_desc(smem_u32(A), 128 * 16, 128), b_desc = umma_desc(smem_u32(B), (uint32_t)rows_b * 16, 128);
        if (elect_one()) {
            if ((variant & ~8) == 0) {
                for (int i = 0; i < iters; ++i) umma_i8(tmem_base + (uint32_t)(i & 1) * 256u, a_desc, b_desc, idesc, i > 1);
            } else {
                const uint32_t bs = variant & 1 ? 1u : 0u, as = variant & 2 ? 256u : 0u;     // descriptor address units of 16 bytes
                for (int i = 0; i < iters; ++i) {
                    umma_i8(tmem_base + (uint32_t)((i >> 6) & 1) * 256u, a_desc + as * (uint32_t)(i & 7), b_desc + bs * (uint32_t)(i & 63), idesc, (i & 63) != 0);
                    if (i % 6 == 5) {                                    // what the numerator kernel does once per ring stage
                        if (variant & 4) umma_commit(bar2);
                        if (variant & 16) tc_fence_after();
                        if (variant & 32) mbar_wait(bar3, 1);               // a barrier that never completed a phase: parity 1 passes at once
                    }
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int launch_i8_peak(mtm_ctx* ctx, int n, int iters)
{
    const size_t smem_bytes = 1024 + 8 * 4096 + (size_t)32 * (n + 64);
    static const int variant = getenv("MTM_B200_PEAK_VARIANT") ? atoi(getenv("MTM_B200_PEAK_VARIANT")) : 0;
    i8_peak_kernel<<<ctx->sm_count, 128, smem_bytes, ctx->stream>>>(iters, n, variant);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

// ------------------------------------------------------------------------------ host side
// Experiment knobs (environment, read once).  None of them is needed in production; they exist so that the
// measurements under profiles/ can be repeated: tile height, ring stage size, epilogue warps, shared-memory ceiling,
// the one-tile-per-CTA kernel, per-role clocks and phase knock-outs of the persistent kernel.
struct TcEnv {
    int force_n = 0, ds = 0, ew = 0, pdbg = 0, mom_cs = -1, stage_bytes = 65536;
    size_t smem_soft = 0;
    bool persist_off = false, persist_force = false, prof = false, mom_rows = false, tma_off = false, lean = false, plan_dbg = false, ctrl_first = false;
    TcEnv()
    {
        auto num = [](const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; };
        force_n = num("MTM_B200_FORCE_N");
        ds = num("MTM_B200_DS");
        if (num("MTM_B200_STAGE_KB") > 0) stage_bytes = num("MTM_B200_STAGE_KB") * 1024;
        ew = getenv("MTM_B200_EW") ? (num("MTM_B200_EW") == 12 ? 12 : 8) : 0;
        pdbg = num("MTM_B200_PDBG");
        smem_soft = getenv("MTM_B200_SMEM_SOFT") ? (size_t)num("MTM_B200_SMEM_SOFT") * 1024 : 0;
        persist_off = getenv("MTM_B200_PERSIST") && num("MTM_B200_PERSIST") == 0;
        persist_force = num("MTM_B200_PERSIST") == 2;
        prof = getenv("MTM_B200_PROF") != nullptr;
        mom_cs = getenv("MTM_B200_MOM_CS") ? num("MTM_B200_MOM_CS") : -1;
        mom_rows = num("MTM_B200_MOM_ROWS") != 0;
        tma_off = getenv("MTM_B200_TMA") && num("MTM_B200_TMA") == 0;
        lean = num("MTM_B200_LEAN") != 0;
        plan_dbg = num("MTM_B200_PLAN_DBG") != 0;
        ctrl_first = num("MTM_B200_CTRL_FIRST") != 0;
    }
};
static const TcEnv& tc_env()
{
    static const TcEnv e;
    return e;
}

// Rows of an image tile that yields n output rows of an h-row template, rounded up to 8 so that every k-block of the
// [k-block][row][16 B] layout starts on a 128-byte boundary (TMA destinations).
static inline int tc_tile_rows(int n, int h) { return (n + h - 1 + 7) & ~7; }

bool tc_path_supported(const mtm_ctx* ctx, int method, int h, int w)
{
    const int C = ctx->img.C;
    if ((C != 1 && C != 3 && C != 4) || method < 0 || method > 5) return false;
    if ((long long)h * w < 16 || (double)h * w * C * 65025.0 >= 4294967296.0) return false;   // 32-bit exact range; tiny windows -> fp64 path
    return true;
}

// Plans a group of `count` same-size templates.  Returns false when the tile does not fit shared memory.
bool tc_plan_group(int mode, int h, int w, int C, TcGroup& g)
{
    g.mode = mode; g.h = h; g.w = w;
    if (mode == 1 && C != 1) return false;                 // the aliased 128-offset layout needs a 1-byte x-step
    const int nx = mode == 0 ? 16 : 128;
    g.nk = (w * C + (nx - 1) * C + 31) / 32;               // band: template row bytes + the largest x-shift
    g.a_kblk = mode == 0 ? 2048 : 256;
    g.slab_bytes = mode == 0 ? 2 * g.nk * 2048 : (7 + 2 * g.nk) * 256;
    g.ds = 1;
    while ((g.ds + 1) * g.slab_bytes <= 16384 && g.ds < h) ++g.ds;
    const size_t ring = (size_t)TC_STAGES * g.ds * g.slab_bytes;
    const int force_n = tc_env().force_n;
    const int candidates[4] = {force_n ? force_n : 256, force_n ? force_n : 128, force_n ? force_n : 64, force_n ? force_n : 32};
    g.N = 0;
    for (int pass = 0; pass < 2 && !g.N; ++pass) {
        const size_t budget = pass == 0 ? 110 * 1024 : 224 * 1024;       // first try 2 CTAs per SM
        for (int n : candidates) {
            const size_t tile = ((size_t)2 * g.nk * tc_tile_rows(n, h) * 16 + 127) & ~(size_t)127;
            if (tile + ring + 256 <= budget && tc_tile_rows(n, h) * 16 < (1 << 18)) { g.N = n; break; }
        }
    }
    if (!g.N) return false;
    g.R = tc_tile_rows(g.N, h);
    g.smem = (((size_t)2 * g.nk * g.R * 16 + 127) & ~(size_t)127) + ring + 256;
    g.eff = (double)w * C / (32.0 * g.nk);
    return true;
}

int launch_toeplitz_prep(mtm_ctx* ctx, const TcGroup& g)
{
    const int64_t pieces = (int64_t)g.h * (g.slab_bytes / 16);
    const int blocks = (int)std::min<int64_t>((pieces + 255) / 256, 4096);
    if (ctx->tmpl_u16) {                                   // high- and low-byte planes of 16-bit templates
        for (int plane = 0; plane < 2; ++plane) {
            toeplitz_prep_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_tmpl8 + plane * ctx->tmpl8_plane, ctx->d_meta, ctx->d_order + g.first,
                                                                 g.count, g.mode, g.h, g.w, g.nk, g.slab_bytes, 1,
                                                                 ctx->d_slabs + plane * ctx->slab_plane + g.arena_off, ctx->d_pix8);
            MTM_LAUNCH_CHECK(ctx);
        }
        return MTM_OK;
    }
    toeplitz_prep_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_tmpl, ctx->d_meta, ctx->d_order + g.first, g.count, g.mode,
                                                         g.h, g.w, g.nk, g.slab_bytes, ctx->tmpl_C, ctx->d_slabs + g.arena_off, nullptr);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_window_moments(mtm_ctx* ctx, int size_first, int size_count)
{
    const ImageDev& im = ctx->img;
    SatView sv{im.sat_s, im.sat_q, im.sat_pitch, (int64_t)(im.H + 1) * im.sat_pitch};
    int64_t n = 0;
    const SizeDesc* h_first = ctx->h_sizes.data() + size_first;
    const SizeDesc* d_first = ctx->d_sizes + size_first;
    for (int q = 0; q < size_count; ++q) n = std::max<int64_t>(n, (int64_t)h_first[q].mh * h_first[q].mw);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16));
    if (tc_env().mom_rows) {
        int mh = 1, mw = 1;
        for (int q = 0; q < size_count; ++q) { mh = std::max(mh, h_first[q].mh); mw = std::max(mw, h_first[q].mw); }
        const dim3 rgrid((unsigned)std::min((mw + 255) / 256, 8), (unsigned)std::min(mh, ctx->sm_count), (unsigned)size_count);
        switch (im.C) {
            case 1: window_moments_rows_kernel<1><<<rgrid, 256, 0, ctx->stream>>>(sv, im.sat_q32, d_first, ctx->d_wS, ctx->d_wR, ctx->moments_total); break;
            case 3: window_moments_rows_kernel<3><<<rgrid, 256, 0, ctx->stream>>>(sv, im.sat_q32, d_first, ctx->d_wS, ctx->d_wR, ctx->moments_total); break;
            default: window_moments_rows_kernel<4><<<rgrid, 256, 0, ctx->stream>>>(sv, im.sat_q32, d_first, ctx->d_wS, ctx->d_wR, ctx->moments_total); break;
        }
        MTM_LAUNCH_CHECK(ctx);
        return MTM_OK;
    }
    const dim3 grid(blocks, (unsigned)size_count);
    // MTM_B200_MOM_CS=1 (experiment): evict-first stores.  Measured neutral on C5 (7.35 against 7.42 ms per step,
    // profiles/README.md): the 64-size sweep is not limited by the moment maps evicting the tables, so the default stays off.
    const bool stream_stores = tc_env().mom_cs > 0;
    if (stream_stores)
        window_moments_kernel<true><<<grid, 256, 0, ctx->stream>>>(sv, im.sat_q32, d_first, ctx->d_wS, ctx->d_wR, im.C, ctx->moments_total);
    else
        window_moments_kernel<false><<<grid, 256, 0, ctx->stream>>>(sv, im.sat_q32, d_first, ctx->d_wS, ctx->d_wR, im.C, ctx->moments_total);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

struct AccumArgs { int img_plane, tmpl_plane; double weight; bool first; };

static int launch_ncc_tc_impl(mtm_ctx* ctx, const TcGroup& g, int method, const AccumArgs* accum, int y_base, int rows)
{
    const ImageDev& im = ctx->img;
    TcParams p{};
    p.method = method;
    p.sat = SatView{im.sat_s, im.sat_q, im.sat_pitch, (int64_t)(im.H + 1) * im.sat_pitch};
    p.img = im.pix; p.pitch = im.pitch; p.H = im.H; p.W = im.W;
    // epilogue flavour = kernel instantiation: 0 default method, 1 float64 rules of the other methods, 2 byte-plane accumulation,
    // 3 default method without the score map (hits only: candidate list and / or per-template arg-max)
    const int kmode = accum ? 2 : (method != MTM_TM_CCOEFF_NORMED ? 1 : (ctx->hits_only ? 3 : 0));
    p.slabs = ctx->d_slabs + g.arena_off;
    if (accum) {
        if (accum->img_plane) p.img = im.pix_lo;
        p.slabs += accum->tmpl_plane * ctx->slab_plane;
        p.acc = ctx->d_acc; p.acc_weight = accum->weight; p.acc_first = accum->first ? 1 : 0;
    }
    p.slab_bytes = g.slab_bytes; p.a_kblk = g.a_kblk; p.nk = g.nk; p.ds = g.ds;
    p.mode = g.mode; p.h = g.h; p.w = g.w;
    p.mh = im.H - g.h_min + 1; p.mw = im.W - g.w_min + 1;      // tile grid covers the largest member map ...
    p.y_base = y_base; p.rows = std::min(rows, p.mh - y_base); p.band_rows = g.band_rows;      // ... rows [y_base, y_base + rows) of it in this launch
    p.meta = ctx->d_meta; p.order = ctx->d_order + g.first; p.count = g.count;
    p.S = ctx->d_wS; p.rsD = ctx->d_wR; p.maps = ctx->d_maps;
    p.C = im.C; p.mom_plane = ctx->moments_total;
    if (kmode == 3 && ctx->best_on) p.best = ctx->d_best;
    if (ctx->cand_on && (kmode == 0 || kmode == 3)) { p.cand = ctx->d_cand; p.cand_count = ctx->d_cand_count; p.cand_cap = MTM_CAND_CAP; p.cand_thr = ctx->cand_thr; }
    // Image tiles through the TMA unit (default; MTM_B200_TMA=0 or a failing encoder: register staging by the stager warps).
    // Called once p.R is final: a box holds at most 256 rows.
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    auto plan_tma = [&]() {
        p.tma_chunks = (p.R + 255) / 256;
        p.tma_rc = ((p.R + p.tma_chunks - 1) / p.tma_chunks + 7) & ~7;
        p.tma = (!tc_env().tma_off && encode_tile_map(&tmap, p.img, p.pitch, p.H, p.tma_rc)) ? 1 : 0;
        if (p.tma) ctx->ctr.tma_launches++;
    };

    // ---- plan of the one-tile-per-CTA kernel (ncc_tc_kernel).  It is the only choice when two tiles do not fit shared memory
    // (256 x 256 templates), and it is COMPARED with the persistent plan otherwise: both models count clocks per SM.  Measured on
    // 2048 x 2048 images (profiles/README.md, run r2ah): 8 templates of 200 x 200 px 1.71 ms persistent (112-row tiles, three
    // starved ring stages) against 1.15 ms one tile per CTA (240-row tiles); 240 x 240: 2.81 against 1.79 ms; 160 x 160: equal;
    // every BASELINE config stays persistent by a wide margin of the models.
    // Tile height AND rows per ring stage together (second session of round 2).  The planned (g.N, g.ds) keeps a 4 x 16-KB ring,
    // which for 256 x 256 templates (C3: mode B, 7.9-KB slabs that serve 12 MMAs each) capped the tile at 128 rows: 961 tiles, 61 of
    // them one pixel wide or high, 7 waves.  A mode-B launch needs ~13 KB of slabs in flight, so one row per stage frees 32 KB and
    // the tile grows to 208 rows: 589 tiles = 3.98 waves.  Cost (clocks per SM): waves x (MMAs of a tile at 1.1 (n/2 + 8) clocks,
    // stage bubbles and ring starvation as in the persistent model, + tile load and epilogue, which two co-resident CTAs hide).
    struct OnePlan { int N = 0, ds = 1; double cost = 1e300; } one;
    auto smem_one = [&](int n, int ds) { return (((size_t)2 * g.nk * tc_tile_rows(n, g.h) * 16 + 127) & ~(size_t)127) + (size_t)TC_STAGES * ds * g.slab_bytes + 256; };
    {
        const int xw_1 = g.mode == 0 ? 16 : 128;
        const int gx_1 = (p.mw + xw_1 - 1) / xw_1;
        for (int ds = g.ds; ds >= 1; --ds) {
            if (tc_env().ds && ds != std::min(tc_env().ds, g.ds)) continue;
            for (int n = 256; n >= 32; n -= 16) {
                if (tc_env().force_n && n != tc_env().force_n) continue;
                const size_t smem = smem_one(n, ds);
                if (smem > 227 * 1024 || tc_tile_rows(n, g.h) * 16 >= (1 << 18)) continue;
                const int per_sm = (2 * smem <= 226 * 1024) ? 2 : 1;              // TMEM (<= 256 columns) also allows 2
                const long long tiles = (long long)gx_1 * ((p.rows + n - 1) / n);
                const long long waves = (tiles + (long long)per_sm * ctx->sm_count - 1) / ((long long)per_sm * ctx->sm_count);
                const double t_mma = 1.1 * (0.5 * n + 8.0);
                const double need = 2500.0 * ((double)g.slab_bytes / g.nk) / t_mma;
                const double starve = std::max(1.0, need / ((double)(TC_STAGES - 1) * ds * g.slab_bytes));
                const double mma_tile = ((double)g.h * g.nk * t_mma + std::ceil((double)g.h / ds) * std::max(0.0, 390.0 - 3.0 * t_mma)) * starve;
                const double other = 40.0 * (n + g.h) + 70.0 * n;                  // tile load + epilogue (8 warps)
                const double cost = (double)waves * (per_sm * mma_tile + (per_sm == 1 ? other : 0.3 * other));
                if (cost < one.cost - 1e-9) { one.cost = cost; one.N = n; one.ds = ds; }
            }
        }
    }

    // ---- persistent pipeline (two image tiles + slab ring in shared memory, two accumulators in TMEM)
    if (!tc_env().persist_off) {
        const int xw_p = g.mode == 0 ? 16 : 128;
        const int gx_p = (p.mw + xw_p - 1) / xw_p;
        const int force_n = tc_env().force_n, force_ew = tc_env().ew, force_ds = tc_env().ds;
        // Cost model (clocks per CTA), fitted to measurements (profiles/README.md, round 2):
        //  * one MMA (M128 x n x K32, u8) from shared memory: n/2 + 8 clocks with resident zero operands (i8_peak_kernel),
        //    ~10 % more on real data;
        //  * between two ring stages the issuing thread spends ~390 clocks (barrier poll, fence, election, descriptors, commit)
        //    while the pipe holds about three queued MMAs: every stage costs max(0, 390 - 3 t_mma) clocks of idle pipe, so small
        //    tiles want MANY template rows per stage (C4: 24-KB stages 0.373 ms, 64-KB stages 0.323 ms);
        //  * the slab stream needs rate x latency bytes in flight (4 KB per MMA, ~2500 clocks from L2): a ring with fewer bytes
        //    beyond the stage being consumed runs at in_flight / need of the rate (C2: 2 x 64 KB slower than 3 x 48 KB);
        //  * the epilogue of a tile (~TCP_EPI_CLK_PER_ROW clocks per row with 8 warps) overlaps the MMAs of the next one; the
        //    first image tile and the last epilogue are exposed.
        int bestN = 0, best_stages = 0, best_ew = 8, best_ds = 1;
        double best_cost = 1e300;
        const int ds_cap = std::max(1, std::min(g.h, tc_env().stage_bytes / g.slab_bytes));
        for (int n = 256; n >= 32; n -= 16) {
            if (force_n && n != force_n) continue;
            const size_t tile_b = ((size_t)2 * g.nk * tc_tile_rows(n, g.h) * 16 + 127) & ~(size_t)127;
            if (256 + 2 * tile_b + 2 * (size_t)g.slab_bytes > 227 * 1024) continue;
            const size_t ring_space = 227 * 1024 - 256 - 2 * tile_b;
            const long long tiles = (long long)gx_p * ((p.rows + n - 1) / n);
            const long long per_cta = (tiles + ctx->sm_count - 1) / ctx->sm_count;
            const double t_mma = 1.1 * (0.5 * n + 8.0);
            const double bubble = std::max(0.0, 390.0 - 3.0 * t_mma);
            const double need = 2500.0 * 4096.0 / t_mma;                       // bytes in flight that hide the L2 latency
            for (int ds = 1; ds <= ds_cap; ++ds) {
                if (force_ds && ds != std::min(force_ds, g.h)) continue;
                const size_t stage_b = (size_t)ds * g.slab_bytes;
                const int stages = (int)std::min<size_t>(TCP_MAX_STAGES, ring_space / stage_b);
                if (stages < 2) break;
                const double in_flight = (double)(stages - 1) * (double)stage_b;
                const double starve = std::max(1.0, need / in_flight);
                const double mma_tile = ((double)g.h * g.nk * t_mma + std::ceil((double)g.h / ds) * bubble) * starve;
                for (int ew = 8; ew <= 12; ew += 4) {
                    if (force_ew && ew != force_ew) continue;
                    // the ew / 4 warps of a lane quarter split the n / 16 column batches: the slowest one sets the pace
                    const int parts = ew / 4, n_eff = 16 * parts * ((n / 16 + parts - 1) / parts);
                    const double epi_tile = n_eff * (ew == 8 ? TCP_EPI_CLK_PER_ROW : TCP_EPI_CLK_PER_ROW_12);
                    // (A/B on C2 / C4 / C5, round 2: 12 epilogue warps win everywhere, 10 % at C4 -- also with several streams per GPU)
                    // (the two phases of a tile overlap imperfectly: the shorter one still costs a share -- measured: the stage size
                    // of an epilogue-bound launch changes its time)
                    const double tile_clk = std::max(mma_tile, epi_tile) + 0.3 * std::min(mma_tile, epi_tile);
                    const double cost = (double)per_cta * tile_clk + epi_tile + 40.0 * (n + g.h);
                    if (cost < best_cost - 1e-9) { best_cost = cost; bestN = n; best_stages = stages; best_ew = ew; best_ds = ds; }
                }
            }
        }
        const int ds_p = best_ds;
        const size_t stage_b = (size_t)ds_p * g.slab_bytes;
        if (tc_env().plan_dbg && bestN)
            fprintf(stderr, "[mtm plan] h=%d w=%d nk=%d count=%d rows=%d: N=%d ds=%d (%zu KB) stages=%d ew=%d model %.0f clk\n", g.h, g.w, g.nk, g.count, p.rows,
                    bestN, ds_p, stage_b / 1024, best_stages, best_ew, best_cost);
        // MTM_B200_PERSIST=2 keeps the persistent kernel whenever it fits (A/B runs)
        const bool one_wins = one.N && one.cost < 0.9 * best_cost && !tc_env().persist_force;
        if (bestN && !one_wins) {
            const size_t tile_b = ((size_t)2 * g.nk * tc_tile_rows(bestN, g.h) * 16 + 127) & ~(size_t)127;
            p.N = bestN; p.R = tc_tile_rows(bestN, g.h); p.stages = best_stages; p.ds = ds_p;
            p.tiles_x = gx_p; p.tiles_total = gx_p * ((p.rows + bestN - 1) / bestN);
            plan_tma();
            const size_t smem_bytes = 256 + 2 * tile_b + (size_t)best_stages * stage_b;
            if (!ctx->tcp_attr_set) {
                const int big = 227 * 1024;
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<false, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<true, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<false, 12, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<true, 12, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<false, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<false, 12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<false, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<false, 12, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<false, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_kernel<false, 12, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_lean_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_persist_lean_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                ctx->tcp_attr_set = true;
            }
            const int grid_p = std::min(p.tiles_total, ctx->sm_count);
            const bool prof = tc_env().prof && kmode == 0;            // debug: per-CTA role clocks to stderr
            const int pdbg = tc_env().pdbg;
            long long* d_prof = nullptr;
            if (prof) {
                MTM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&d_prof), (size_t)grid_p * 16 * sizeof(long long)));
                MTM_CUDA(ctx, cudaMemsetAsync(d_prof, 0, (size_t)grid_p * 16 * sizeof(long long), ctx->stream));
                p.prof = d_prof;
            }
            p.dbg = pdbg;
            p.ctrl_first = tc_env().ctrl_first ? 1 : 0;
            const int ew = best_ew;
            if (kmode == 2) {
                if (ew == 12) ncc_tc_persist_kernel<false, 12, 2><<<grid_p, 32 * (15 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
                else ncc_tc_persist_kernel<false, 8, 2><<<grid_p, 32 * (11 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
            } else if (kmode == 1) {
                if (ew == 12) ncc_tc_persist_kernel<false, 12, 1><<<grid_p, 32 * (15 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
                else ncc_tc_persist_kernel<false, 8, 1><<<grid_p, 32 * (11 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
            } else if (kmode == 3) {
                if (ew == 12) ncc_tc_persist_kernel<false, 12, 3><<<grid_p, 32 * (15 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
                else if (tc_env().lean) ncc_tc_persist_lean_kernel<3><<<grid_p, 32 * (11 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
                else ncc_tc_persist_kernel<false, 8, 3><<<grid_p, 32 * (11 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
            } else if (ew == 12) {
                if (prof) ncc_tc_persist_kernel<true, 12, 0><<<grid_p, 32 * (15 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
                else ncc_tc_persist_kernel<false, 12, 0><<<grid_p, 32 * (15 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
            } else {
                if (prof) ncc_tc_persist_kernel<true, 8, 0><<<grid_p, 32 * (11 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
                else if (tc_env().lean) ncc_tc_persist_lean_kernel<0><<<grid_p, 32 * (11 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
                else ncc_tc_persist_kernel<false, 8, 0><<<grid_p, 32 * (11 + p.ctrl_first), smem_bytes, ctx->stream>>>(p, tmap);
            }
            MTM_LAUNCH_CHECK(ctx);
            if (prof) {
                std::vector<long long> hp((size_t)grid_p * 16);
                MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                MTM_CUDA(ctx, cudaMemcpy(hp.data(), d_prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                cudaFree(d_prof);
                double a[16] = {0};
                for (int c = 0; c < grid_p; ++c) for (int k = 0; k < 16; ++k) a[k] += (double)hp[16 * c + k] / grid_p;
                fprintf(stderr, "[mtm prof] persist: %d CTAs x %.2f tiles N=%d stages=%d ew=%d smem=%zu | mma: wait tile %.0f acc %.0f slabs %.0f total %.0f | "
                        "epilogue: wait %.0f work %.0f | stager: wait %.0f work %.0f | producer wait %.0f\n",
                        grid_p, (double)p.tiles_total / grid_p, p.N, p.stages, ew, smem_bytes, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
            }
            return MTM_OK;
        }
    }

    // ---- one tile per CTA (planned above)
    if (!one.N) return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "tensor path: no tile of a %d x %d template fits shared memory", g.h, g.w);
    p.N = one.N; p.R = tc_tile_rows(one.N, g.h); p.ds = one.ds;
    plan_tma();
    const size_t smem_bytes = smem_one(one.N, one.ds);
    if (tc_env().plan_dbg)
        fprintf(stderr, "[mtm plan] one tile per CTA: h=%d w=%d nk=%d mode=%d: N=%d ds=%d smem=%zu model %.0f clk\n", g.h, g.w, g.nk, g.mode, one.N, one.ds, smem_bytes, one.cost);
    if (!ctx->tc_attr_set) {
        MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MTM_CUDA(ctx, cudaFuncSetAttribute(ncc_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ctx->tc_attr_set = true;
    }
    const int xw = g.mode == 0 ? 16 : 128;
    dim3 grid((p.mw + xw - 1) / xw, (p.rows + p.N - 1) / p.N);
    if (kmode == 2) ncc_tc_kernel<2><<<grid, TC_THREADS, smem_bytes, ctx->stream>>>(p, tmap);
    else if (kmode == 1) ncc_tc_kernel<1><<<grid, TC_THREADS, smem_bytes, ctx->stream>>>(p, tmap);
    else if (kmode == 3) ncc_tc_kernel<3><<<grid, TC_THREADS, smem_bytes, ctx->stream>>>(p, tmap);
    else ncc_tc_kernel<0><<<grid, TC_THREADS, smem_bytes, ctx->stream>>>(p, tmap);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_ncc_tc(mtm_ctx* ctx, const TcGroup& g, int method, int y_base, int rows) { return launch_ncc_tc_impl(ctx, g, method, nullptr, y_base, rows); }

int launch_ncc_tc_accum(mtm_ctx* ctx, const TcGroup& g, int img_plane, int tmpl_plane, double weight, bool first)
{
    const AccumArgs a{img_plane, tmpl_plane, weight, first};
    return launch_ncc_tc_impl(ctx, g, MTM_TM_CCORR, &a, 0, ctx->img.H - g.h_min + 1);
}
