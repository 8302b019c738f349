// box_moments.cu -- window moments of every distinct template size, straight from the image (no summed-area tables).
//
// Same outputs, bit for bit, as window_moments_kernel (ncc_tc.cu): per window position the per-channel sums S_c and
// rsD = rsqrt(A*Q - sum_c S_c^2) (0 for an exactly flat window) -- the denominator statistics that OpenCV's
// common_matchTemplate takes from integral(img, sum, sqsum) (third-party; reached from MTM/__init__.py:92).  The
// summed-area route costs three launches per image (row prefixes, column prefixes, moment sweep: 27 + 15 us at
// BASELINE configs[1]) and, per window position and size, eight table corners through L2 (~24 B with the 8 B written;
// configs[4]: 64 sizes, 2.4 ms).  Box sums need ~2 image bytes per position instead:
//
//   vertical running sums   V_c(y, x') = sum_{dy < h} I_c[y+dy][x'],  VQ(y, x') = sum_c sum_{dy < h} I_c[y+dy][x']^2
//                           kept in registers, one add (row y+h-1) and one subtract (row y-1) per output row;
//   horizontal window sums  block-wide exclusive prefix P of V along x (warp shuffles + one shared-memory hop),
//                           S_c(y, x) = P_c[x+w] - P_c[x]  -- modulo 2^32, exact because window sums stay below 2^32
//                           on the tensor path (h*w*C <= 66051).
//
// A CTA owns one size (blockIdx.z), a strip of 1024 image columns (4 per thread, aligned 32-bit loads) and a band of
// output rows; it first accumulates the h-1 rows above its band (adds only, eight row loads in flight); the two rows that enter
// and leave the window are fetched one output row ahead.  HBM-bound by design: the image is read from L2 (each row
// (band overlap) times), the moment maps are written once: 8 B (C = 1) or 4(C+1) B per window position.
//
// Experiment knob MTM_B200_MOM_BOX=1 (mtm_api.cu: the summed-area tables are then built on demand only).  Checked on the
// CPU against window_moments_kernel (tests/test_kernel_emulation.py); NOT YET MEASURED ON THE GPU.
#include "mtm_internal.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int BM_THREADS = 256;
constexpr int BM_PX = 4;                               // image columns per thread
constexpr int BM_COLS = BM_THREADS * BM_PX;            // image columns per strip
constexpr int BM_ROW = BM_COLS + 4;                    // prefix row in shared memory (16-byte aligned rows, entry BM_COLS used)

struct BoxParams {
    const uint8_t* img; int64_t pitch;                 // zero-padded u8 rows, interleaved channels
    const SizeDesc* sizes;                             // blockIdx.z -> (h, w, mh, mw, ring segment, rows per band) of the launch's sizes
    uint32_t* S; float* rsD;                           // the moment ring (tile-major segments, mtm_internal.cuh)
    int64_t mom_plane;
    int y_begin, rows;                                 // output rows [y_begin, y_begin + rows) of every size (clipped to its map)
};

template <int C>
__global__ void __launch_bounds__(BM_THREADS)
box_moments_kernel(const BoxParams p)
{
    __shared__ __align__(16) uint32_t P[2][C + 1][BM_ROW];
    __shared__ uint32_t wtot[2][C + 1][BM_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const SizeDesc sd = p.sizes[blockIdx.z];
    const int strip_out = (BM_COLS - (sd.w - 1)) & ~3; // window positions per strip: a multiple of 4 (aligned loads)
    const int y_end = min(sd.mh, p.y_begin + p.rows);  // rows of this launch that exist for this size
    const int band = (max(y_end - p.y_begin, 0) + (int)gridDim.y - 1) / (int)gridDim.y;
    const int x0 = blockIdx.x * strip_out;             // first image column of the strip == its first window position
    const int y0 = p.y_begin + blockIdx.y * band, y1 = min(y_end, y0 + band);
    if (x0 >= sd.mw || y0 >= y_end) return;            // the grid is sized for the largest map of the launch
    const int64_t col_byte = ((int64_t)x0 + BM_PX * tid) * C;      // a multiple of 4: strip_out and BM_PX are
    uint32_t V[C + 1][BM_PX];
#pragma unroll
    for (int q = 0; q <= C; ++q)
#pragma unroll
        for (int j = 0; j < BM_PX; ++j) V[q][j] = 0u;

    // the C 32-bit words holding this thread's 4 pixels of image row r
    auto load_row = [&](int r, uint32_t (&wd)[C]) {
        const uint8_t* row = p.img + (int64_t)r * p.pitch + col_byte;
#pragma unroll
        for (int k = 0; k < C; ++k)
            wd[k] = (col_byte + 4 * k + 4 <= p.pitch) ? __ldg(reinterpret_cast<const uint32_t*>(row) + k) : 0u;   // beyond the row: no pixels
    };
    // adds or subtracts those words to / from the running column sums
    auto accumulate = [&](const uint32_t (&wd)[C], bool sub) {
#pragma unroll
        for (int j = 0; j < BM_PX; ++j) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int b = j * C + c;
                const uint32_t v = (wd[b >> 2] >> (8 * (b & 3))) & 255u;
                if (sub) { V[c][j] -= v; V[C][j] -= v * v; } else { V[c][j] += v; V[C][j] += v * v; }
            }
        }
    };

    // the h-1 rows above the band: eight independent loads in flight (one L2 round trip per eight rows, not per row)
    constexpr int WARM = 8;
    int r = y0;
    const int r_end = y0 + sd.h - 1;
    for (; r + WARM <= r_end; r += WARM) {
        uint32_t wd[WARM][C];
#pragma unroll
        for (int u = 0; u < WARM; ++u) load_row(r + u, wd[u]);
#pragma unroll
        for (int u = 0; u < WARM; ++u) accumulate(wd[u], false);
    }
    for (; r < r_end; ++r) {
        uint32_t wd[C];
        load_row(r, wd);
        accumulate(wd, false);
    }
    const uint32_t area = (uint32_t)sd.h * (uint32_t)sd.w;
    // the rows entering and leaving the window are fetched one output row ahead, behind the scan and the stores of the current one
    uint32_t w_in[C], w_out[C];
    load_row(y0 + sd.h - 1, w_in);
    load_row(y0, w_out);
    for (int y = y0; y < y1; ++y) {
        const int buf = (y - y0) & 1;
        accumulate(w_in, false);
        uint32_t n_in[C], n_out[C];
        const bool more = y + 1 < y1;
        if (more) { load_row(y + sd.h, n_in); load_row(y + 1, n_out); }
        // block-wide exclusive prefix of every quantity along x
        uint32_t incl[C + 1];
#pragma unroll
        for (int q = 0; q <= C; ++q) {
            uint32_t s = V[q][0] + V[q][1] + V[q][2] + V[q][3];
            const uint32_t own = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += n;
            }
            if (lane == 31) wtot[buf][q][wid] = s;
            incl[q] = s - own;                         // exclusive inside the warp
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q <= C; ++q) {
            uint32_t e = incl[q];
            for (int k = 0; k < wid; ++k) e += wtot[buf][q][k];
            uint4 v;
            v.x = e; v.y = v.x + V[q][0]; v.z = v.y + V[q][1]; v.w = v.z + V[q][2];
            *reinterpret_cast<uint4*>(&P[buf][q][BM_PX * tid]) = v;
            if (tid == BM_THREADS - 1) P[buf][q][BM_COLS] = v.w + V[q][3];
        }
        __syncthreads();
        // window positions tid, tid + 256, ... of the strip: conflict-free shared-memory reads, coalesced stores
#pragma unroll
        for (int j = 0; j < BM_PX; ++j) {
            const int xl = tid + j * BM_THREADS;
            const int x = x0 + xl;
            if (xl >= strip_out || x >= sd.mw) continue;
            const int64_t out_idx = sd.off + mom_index(x, y - p.y_begin, sd.band);
            const uint32_t qs = P[buf][C][xl + sd.w] - P[buf][C][xl];
            unsigned long long d1 = (unsigned long long)area * qs;
            uint32_t s0 = 0;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const uint32_t s = P[buf][c][xl + sd.w] - P[buf][c][xl];
                d1 -= (unsigned long long)s * s;
                if (C > 1) p.S[c * p.mom_plane + out_idx] = s;
                s0 = s;
            }
            const float rs = d1 ? rsqrtf((float)d1) : 0.0f;
            if (C > 1) p.rsD[out_idx] = rs;
            else reinterpret_cast<uint2*>(p.S)[out_idx] = make_uint2(s0, __float_as_uint(rs));
        }
        accumulate(w_out, true);
        if (more) {
#pragma unroll
            for (int k = 0; k < C; ++k) { w_in[k] = n_in[k]; w_out[k] = n_out[k]; }
        }
    }
}

// ---- single channel, throughput form ------------------------------------------------------------------------------------
// Same quantities, same bits; about a third of the instructions per window position (the generic kernel above spends ~84,
// which made the 64-size sweep of BASELINE configs[4] issue-bound at 1.3 TB/s of moment stores):
//   * 8 columns per thread (one aligned 8-byte load per image row), strips of 2048 columns: the block-wide scan is paid
//     once per 8 window positions;
//   * vertical sums of I as two 16-bit lanes per register (a column sum is at most 255 * h < 2^16 for h <= 257): the row
//     entering / leaving the window is spread into the lanes with two byte permutes per word;
//   * ONE barrier per output row: every warp publishes its own exclusive prefix and its total (double-buffered), the
//     readers add the warp bases themselves;
//   * {S prefix, Q prefix} interleaved in shared memory: one 8-byte read per window edge.
constexpr int B1_THREADS = 256;
constexpr int B1_PX = 8;
constexpr int B1_COLS = B1_THREADS * B1_PX;           // 2048 image columns per strip
constexpr int B1_WARPS = B1_THREADS / 32;

// LEAN: every window of the launch is at most 256 px wide (one instantiation per output loop: the row loop has to stay small --
// a kernel that carried both loops, or the eight positions unrolled without their tests, ran 8-17 % slower: instruction fetch).
template <bool LEAN, bool ROLLED = false, int OCC = 3>
__global__ void __launch_bounds__(B1_THREADS, OCC)
box_moments_c1_kernel(const BoxParams p)
{
    __shared__ __align__(16) uint2 P[2][B1_COLS + 8];  // exclusive prefix INSIDE the owning warp's 256 columns: {S, Q mod 2^32}
    __shared__ uint2 wtot[2][B1_WARPS];                // the warps' totals
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const SizeDesc sd = p.sizes[blockIdx.z];
    const int strip_out = (B1_COLS - (sd.w - 1)) & ~15;             // a multiple of 16: strips start on a tile column of the ring
    const int y_end = min(sd.mh, p.y_begin + p.rows);
    const int band = (max(y_end - p.y_begin, 0) + (int)gridDim.y - 1) / (int)gridDim.y;
    const int x0 = blockIdx.x * strip_out;
    const int y0 = p.y_begin + blockIdx.y * band, y1 = min(y_end, y0 + band);
    if (x0 >= sd.mw || y0 >= y_end) return;
    const int64_t col_byte = (int64_t)x0 + B1_PX * tid;            // a multiple of 8
    const bool in_row = col_byte + 8 <= p.pitch;                    // beyond the padded row: no pixels
    const uint8_t* col = p.img + col_byte;
    uint32_t VS[4] = {0u, 0u, 0u, 0u};                              // columns (0,1) (2,3) (4,5) (6,7) as 16-bit lanes
    uint32_t VQ[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) VQ[j] = 0u;
    if (tid == 0) { P[0][B1_COLS] = make_uint2(0u, 0u); P[1][B1_COLS] = make_uint2(0u, 0u); }

    auto load_row = [&](int r) -> uint2 {
        return in_row ? __ldg(reinterpret_cast<const uint2*>(col + (int64_t)r * p.pitch)) : make_uint2(0u, 0u);
    };
    auto add_row = [&](uint2 wd) {
        VS[0] += __byte_perm(wd.x, 0u, 0x4140); VS[1] += __byte_perm(wd.x, 0u, 0x4342);
        VS[2] += __byte_perm(wd.y, 0u, 0x4140); VS[3] += __byte_perm(wd.y, 0u, 0x4342);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t a = __byte_perm(wd.x, 0u, 0x4440 + j), b = __byte_perm(wd.y, 0u, 0x4440 + j);
            VQ[j] += a * a; VQ[4 + j] += b * b;
        }
    };
    auto sub_row = [&](uint2 wd) {                                  // lane-wise: a column sum contains the row being removed
        VS[0] -= __byte_perm(wd.x, 0u, 0x4140); VS[1] -= __byte_perm(wd.x, 0u, 0x4342);
        VS[2] -= __byte_perm(wd.y, 0u, 0x4140); VS[3] -= __byte_perm(wd.y, 0u, 0x4342);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t a = __byte_perm(wd.x, 0u, 0x4440 + j), b = __byte_perm(wd.y, 0u, 0x4440 + j);
            VQ[j] -= a * a; VQ[4 + j] -= b * b;
        }
    };

    // the h-1 rows above the band: adds only, eight loads in flight
    constexpr int WARM = 8;
    int r = y0;
    const int r_end = y0 + sd.h - 1;
    for (; r + WARM <= r_end; r += WARM) {
        uint2 wd[WARM];
#pragma unroll
        for (int u = 0; u < WARM; ++u) wd[u] = load_row(r + u);
#pragma unroll
        for (int u = 0; u < WARM; ++u) add_row(wd[u]);
    }
    for (; r < r_end; ++r) add_row(load_row(r));

    const uint32_t area = (uint32_t)sd.h * (uint32_t)sd.w;
    uint2* out_base = reinterpret_cast<uint2*>(p.S) + sd.off;
    const int seg_off = (tid + sd.w) >> 8;
    const bool wide = sd.w > B1_THREADS;
    const uint32_t seg_mask = seg_off ? 0xFFFFFFFFu : 0u;
    const int lim = min(strip_out, sd.mw - x0);                        // window positions of this strip
    const int jn = lim > tid ? (lim - tid + B1_THREADS - 1) / B1_THREADS : 0;     // this thread's positions: tid + 256 j, j < jn
    const int64_t dst_step = (int64_t)sd.band * (B1_THREADS / 16) * 16;
    uint2* out_row = out_base + ((int64_t)((x0 >> 4) + (tid >> 4)) * sd.band + (y0 - p.y_begin)) * 16 + (tid & 15);   // row y0 of position j = 0
    uint2 w_in = load_row(y0 + sd.h - 1), w_out = load_row(y0);
    for (int y = y0; y < y1; ++y) {
        const int buf = (y - y0) & 1;
        add_row(w_in);
        const bool more = y + 1 < y1;
        uint2 n_in = make_uint2(0u, 0u), n_out = make_uint2(0u, 0u);
        if (more) { n_in = load_row(y + sd.h); n_out = load_row(y + 1); }      // one output row ahead
        // exclusive prefix of this thread's 8 columns, then of the warp
        uint32_t es[8], eq[8];
        uint32_t ts = 0u, tq = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            es[j] = ts; eq[j] = tq;
            ts += (VS[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
            tq += VQ[j];
        }
        uint32_t is = ts, iq = tq;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ns = __shfl_up_sync(0xffffffffu, is, d), nq = __shfl_up_sync(0xffffffffu, iq, d);
            if (lane >= d) { is += ns; iq += nq; }
        }
        if (lane == 31) wtot[buf][wid] = make_uint2(is, iq);
        const uint32_t bs = is - ts, bq = iq - tq;                  // exclusive inside the warp
        uint4* dst = reinterpret_cast<uint4*>(&P[buf][B1_PX * tid]);
#pragma unroll
        for (int j = 0; j < 8; j += 2) dst[j >> 1] = make_uint4(bs + es[j], bq + eq[j], bs + es[j + 1], bq + eq[j + 1]);
        __syncthreads();                                            // the only barrier of the row (buffers alternate)
        if (LEAN) {
            // Windows no wider than a segment (w <= 256: everything but very wide templates).  Position xl = tid + 256 j lies in
            // segment j and its right edge in segment j + seg_off, seg_off = (tid + w) >> 8 in {0, 1}: the window sum is the
            // difference of the two segment-local prefixes plus, when the edge crossed into the next segment, segment j's total
            // (no running bases).  A thread's positions are a prefix j < jn of the eight; the ring entry of row y is 16 entries
            // (128 bytes) after row y - 1, the one of position j + 1 dst_step entries after position j.
            const uint2* pl = &P[buf][tid];
            const uint2* pr = pl + sd.w;
            uint2* o = out_row;
            auto emit = [&](int j) {
                const uint2 t = wtot[buf][j];
                const uint2 lo = pl[j * B1_THREADS], hi = pr[j * B1_THREADS];
                const uint32_t s = hi.x - lo.x + (t.x & seg_mask);
                const uint32_t qs = hi.y - lo.y + (t.y & seg_mask);
                const unsigned long long d1 = (unsigned long long)area * qs - (unsigned long long)s * s;
                const float rs = d1 ? mtm_rsqrt_normal((float)d1) : 0.0f;      // d1 >= 1: never subnormal
                o[j * dst_step] = make_uint2(s, __float_as_uint(rs));
            };
            if (ROLLED) {                                                  // default: the smallest row loop (see launch_box_moments)
#pragma unroll 1
                for (int j = 0; j < jn; ++j) emit(j);
            } else {
#pragma unroll
                for (int j = 0; j < B1_PX; ++j)
                    if (j < jn) emit(j);
            }
            out_row += 16;
        } else {
        uint2 wb[B1_WARPS + 1];                                     // base of every warp's segment; [B1_WARPS] = the strip total
        wb[0] = make_uint2(0u, 0u);
#pragma unroll
        for (int k = 0; k < B1_WARPS; ++k) { const uint2 t = wtot[buf][k]; wb[k + 1] = make_uint2(wb[k].x + t.x, wb[k].y + t.y); }
        // Window position xl = tid + 256 j lies in segment j; its right edge xl + w in segment j + seg_off, seg_off = (tid + w) >> 8:
        // the same for every j of a thread.  Ring entry of (x0 + xl, y): ((x >> 4) * band + r) * 16 + (x & 15) with x0 a multiple of
        // 16 -> a per-thread base plus j * 16 * band * 16 entries.
        const uint2* pl = &P[buf][tid];
        const uint2* pr = pl + sd.w;
        uint2* outp = out_base + ((int64_t)((x0 >> 4) + (tid >> 4)) * sd.band + (y - p.y_begin)) * 16 + (tid & 15);
#pragma unroll
        for (int j = 0; j < B1_PX; ++j) {
            const int xl = tid + j * B1_THREADS;
            if (xl >= strip_out || x0 + xl >= sd.mw) continue;
            const uint2 lo = pl[j * B1_THREADS], hi = pr[j * B1_THREADS];
            uint2 hb = wb[j];                                        // windows wider than a segment (w > 256): general selection
#pragma unroll
            for (int k = 1; k <= 4; ++k)
                if (j + k <= B1_WARPS && seg_off == k) hb = wb[j + k];
            const uint32_t s = (hi.x + hb.x) - (lo.x + wb[j].x);
            const uint32_t qs = (hi.y + hb.y) - (lo.y + wb[j].y);
            const unsigned long long d1 = (unsigned long long)area * qs - (unsigned long long)s * s;
            const float rs = d1 ? mtm_rsqrt_normal((float)d1) : 0.0f;      // d1 >= 1: never subnormal
            outp[j * dst_step] = make_uint2(s, __float_as_uint(rs));
        }
        }
        sub_row(w_out);
        w_in = n_in; w_out = n_out;
    }
}

}  // namespace

bool box_moments_enabled()
{
    // default ON since round 2 (profiles/README.md: C2 0.101 -> 0.082, C4 0.46 -> 0.41, C5 7.27 -> 6.00 ms per step against the
    // summed-area route); MTM_B200_MOM_BOX=0 restores the tables + window_moments_kernel for A/B runs.
    static const bool on = getenv("MTM_B200_MOM_BOX") == nullptr || atoi(getenv("MTM_B200_MOM_BOX")) != 0;
    return on;
}

// Can the resident (image, template set) take the box-sum route?  Plain uint8, every window narrower than half a strip.
bool box_moments_applicable(const mtm_ctx* ctx)
{
    if (ctx->img_dtype != MTM_U8 || ctx->masked || ctx->h_sizes.empty()) return false;
    const int C = ctx->img.C;
    if (C != 1 && C != 3 && C != 4) return false;
    for (const SizeDesc& sd : ctx->h_sizes) {
        if ((double)sd.h * sd.w * C * 65025.0 >= 4294967296.0) return false;     // window sums of squares must stay below 2^32
        if (sd.w > BM_COLS / 2) return false;
    }
    return true;
}

// ctx->d_sizes holds ctx->h_sizes (ensure_geometry).  Sizes [size_first, size_first + size_count), output rows [y_begin, y_begin + rows).
int launch_box_moments(mtm_ctx* ctx, int size_first, int size_count, int y_begin, int rows)
{
    const ImageDev& im = ctx->img;
    BoxParams p{};
    p.img = im.pix; p.pitch = im.pitch; p.sizes = ctx->d_sizes + size_first;
    p.S = ctx->d_wS; p.rsD = ctx->d_wR; p.mom_plane = ctx->moments_total;
    p.y_begin = y_begin; p.rows = rows;
    const SizeDesc* sizes = ctx->h_sizes.data() + size_first;
    int strips = 1, mh = 1;
    for (int q = 0; q < size_count; ++q) {
        const SizeDesc& sd = sizes[q];
        const int strip_out = (BM_COLS - (sd.w - 1)) & ~3;
        strips = std::max(strips, (sd.mw + strip_out - 1) / strip_out);
        mh = std::max(mh, std::min(sd.mh, y_begin + rows) - y_begin);
    }
    const int n_sizes = size_count;
    static const bool generic_only = getenv("MTM_B200_BOX_GENERIC") != nullptr;      // A/B runs
    bool lanes16 = true;                               // the throughput form keeps column sums in 16-bit lanes: 255 * h < 2^16
    for (int q = 0; q < size_count; ++q) lanes16 = lanes16 && sizes[q].h <= 257 && sizes[q].w <= B1_COLS / 2;
    if (im.C == 1 && !generic_only && lanes16) {
        // throughput form: strips of 2048 columns, about three resident CTAs per SM in total
        int strips1 = 1;
        for (int q = 0; q < size_count; ++q) {
            const int strip_out = (B1_COLS - (sizes[q].w - 1)) & ~15;
            strips1 = std::max(strips1, (sizes[q].mw + strip_out - 1) / strip_out);
        }
        // bands: enough CTAs for three per SM, but a band keeps at least 8 output rows (it first re-adds the h-1 rows above it)
        // CTAs per SM (registers per thread): A/B at C5 (ncu): 3 (72) 1.60 ms, 4 (64) 1.44 ms (default); MTM_B200_BOX_OCC = 3 / 5 for the others
        static const int occ = getenv("MTM_B200_BOX_OCC") ? std::min(6, std::max(3, atoi(getenv("MTM_B200_BOX_OCC")))) : 4;
        const int want = occ * ctx->sm_count;
        const int bands1 = std::max(1, std::min(std::max(1, mh / 8), (want + strips1 * n_sizes - 1) / (strips1 * n_sizes)));
        const dim3 grid1((unsigned)strips1, (unsigned)bands1, (unsigned)n_sizes);
        bool lean = getenv("MTM_B200_BOX_OUT") == nullptr || atoi(getenv("MTM_B200_BOX_OUT")) != 0;       // A/B knob: 0 = general output loop
        for (int q = 0; q < size_count; ++q) lean = lean && sizes[q].w <= B1_THREADS;
        // A/B on C5 (64 sizes, ncu): general loop 1.90 ms, lean unrolled 1.71 ms, lean rolled 1.60 ms (default) -- the row loop is
        // bound by instruction fetch, the smallest body wins.  MTM_B200_BOX_OUT = 0 / 1 select the other two.
        static const bool rolled = getenv("MTM_B200_BOX_OUT") == nullptr || atoi(getenv("MTM_B200_BOX_OUT")) >= 2;
        if (lean && rolled && occ == 4) box_moments_c1_kernel<true, true, 4><<<grid1, B1_THREADS, 0, ctx->stream>>>(p);
        else if (lean && rolled && occ == 5) box_moments_c1_kernel<true, true, 5><<<grid1, B1_THREADS, 0, ctx->stream>>>(p);
        else if (lean && rolled && occ == 6) box_moments_c1_kernel<true, true, 6><<<grid1, B1_THREADS, 0, ctx->stream>>>(p);
        else if (lean && rolled) box_moments_c1_kernel<true, true><<<grid1, B1_THREADS, 0, ctx->stream>>>(p);
        else if (lean) box_moments_c1_kernel<true><<<grid1, B1_THREADS, 0, ctx->stream>>>(p);
        else box_moments_c1_kernel<false><<<grid1, B1_THREADS, 0, ctx->stream>>>(p);
        MTM_LAUNCH_CHECK(ctx);
        return MTM_OK;
    }
    // about two CTAs per SM over all sizes; a band re-reads the h-1 rows above it, so bands stay as tall as that allows
    const int bands = std::max(1, std::min(mh, (2 * ctx->sm_count + strips * n_sizes - 1) / (strips * n_sizes)));
    const dim3 grid((unsigned)strips, (unsigned)bands, (unsigned)n_sizes);
    switch (im.C) {
        case 1: box_moments_kernel<1><<<grid, BM_THREADS, 0, ctx->stream>>>(p); break;
        case 3: box_moments_kernel<3><<<grid, BM_THREADS, 0, ctx->stream>>>(p); break;
        case 4: box_moments_kernel<4><<<grid, BM_THREADS, 0, ctx->stream>>>(p); break;
        default: return mtm_fail(ctx, MTM_ERR_INVALID, "unsupported channel count %d", im.C);
    }
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}
