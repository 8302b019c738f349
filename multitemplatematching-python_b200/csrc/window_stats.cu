// window_stats.cu -- K1: per-channel window statistics of the image and the
// template statistics, i.e. what OpenCV's integral(img, sum, sqsum, CV_64F) and
// meanStdDev(templ) feed into common_matchTemplate (third-party; call site
// MTM/__init__.py:92).  Everything is exact integer arithmetic:
//   sat_s[c] : (H+1)x(W+1) summed-area table of channel c, u32 with wrap-around
//              (any window sum < 2^32 is recovered exactly by the 4-corner
//              difference modulo 2^32);
//   sat_q    : (H+1)x(W+1) summed-area table of sum_c I_c^2, u64.
// HBM-bound: reads H*W*C bytes once, writes (4C+8) B per pixel.
#include "mtm_internal.cuh"

namespace {

// One block (256 threads) per image row: inclusive prefix along x of every channel and of the
// squared norm.  Each thread owns a contiguous run of pixels; runs are combined with a block
// scan.  Output (scratch): C planes of u32 [H][W] then one u32 plane [H][W] (row prefixes of
// squares stay < 2^32 for W*C < 66051).
template <int C>
__global__ void __launch_bounds__(256)
sat_rows_kernel(const uint8_t* __restrict__ img, int64_t pitch, int H, int W, int SP, uint32_t* __restrict__ scratch)
{
    __shared__ uint32_t wsum[C + 1][8];
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint8_t* row = img + (int64_t)y * pitch;
    const int64_t plane = (int64_t)H * SP;
    const int per = (W + 255) / 256;
    const int xa = min(W, tid * per), xb = min(W, xa + per);
    uint32_t tot[C + 1];
#pragma unroll
    for (int c = 0; c <= C; ++c) tot[c] = 0;
    for (int x = xa; x < xb; ++x) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t p = row[(int64_t)x * C + c];
            tot[c] += p;
            tot[C] += p * p;
        }
    }
    uint32_t excl[C + 1];
#pragma unroll
    for (int c = 0; c <= C; ++c) {
        uint32_t s = tot[c];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += n;
        }
        if (lane == 31) wsum[c][wid] = s;
        excl[c] = s - tot[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c <= C; ++c) {
        uint32_t base = 0;
        for (int k = 0; k < wid; ++k) base += wsum[c][k];
        excl[c] += base;
    }
    for (int x = xa; x < xb; ++x) {
        uint32_t sq = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t p = row[(int64_t)x * C + c];
            excl[c] += p;
            sq += p * p;
            scratch[c * plane + (int64_t)y * SP + x] = excl[c];
        }
        excl[C] += sq;
        scratch[C * plane + (int64_t)y * SP + x] = excl[C];
    }
}

// Single-channel fast path: each thread owns `per4` aligned 32-bit words (4 pixels each) of the row;
// 16-byte stores into the (4-aligned pitch) scratch rows.
__global__ void __launch_bounds__(256)
sat_rows_c1_kernel(const uint8_t* __restrict__ img, int64_t pitch, int H, int W, int SP, uint32_t* __restrict__ scratch)
{
    __shared__ uint32_t wsum[2][8];
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(img + (int64_t)y * pitch);
    const int64_t plane = (int64_t)H * SP;
    const int nwords = (W + 3) >> 2;
    const int per4 = (nwords + 255) / 256;
    const int wa = min(nwords, tid * per4), wb = min(nwords, wa + per4);
    uint32_t ts = 0, tq = 0;
    for (int k = wa; k < wb; ++k) {
        uint32_t v = __ldg(row + k);
        if (4 * k + 4 > W) v &= (1u << (8 * (W - 4 * k))) - 1u;        // last word: drop the padding bytes
        ts += __dp4a(v, 0x01010101u, 0u);
        tq += __dp4a(v, v, 0u);
    }
    uint32_t s = ts, q = tq;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ns = __shfl_up_sync(0xffffffffu, s, d), nq = __shfl_up_sync(0xffffffffu, q, d);
        if (lane >= d) { s += ns; q += nq; }
    }
    if (lane == 31) { wsum[0][wid] = s; wsum[1][wid] = q; }
    __syncthreads();
    uint32_t es = s - ts, eq = q - tq;
    for (int k = 0; k < wid; ++k) { es += wsum[0][k]; eq += wsum[1][k]; }
    uint4* ds = reinterpret_cast<uint4*>(scratch + (int64_t)y * SP);
    uint4* dq = reinterpret_cast<uint4*>(scratch + plane + (int64_t)y * SP);
    for (int k = wa; k < wb; ++k) {
        uint32_t v = __ldg(row + k);
        if (4 * k + 4 > W) v &= (1u << (8 * (W - 4 * k))) - 1u;
        const uint32_t p0 = v & 255u, p1 = (v >> 8) & 255u, p2 = (v >> 16) & 255u, p3 = v >> 24;
        uint4 a, b;
        a.x = es + p0; a.y = a.x + p1; a.z = a.y + p2; a.w = a.z + p3; es = a.w;
        b.x = eq + p0 * p0; b.y = b.x + p1 * p1; b.z = b.y + p2 * p2; b.w = b.z + p3 * p3; eq = b.w;
        ds[k] = a;                                                       // columns >= W of the padded row are never read
        dq[k] = b;
    }
}

// Column pass.  Block = 32 SAT columns x 32 row chunks: chunk total -> exclusive scan over the
// chunks -> running prefix (second sweep re-reads the chunk from L1/L2; loads are batched 8 deep).
// blockIdx.y picks the table: 0..C-1 -> sat_s[c] (u32), C -> sat_q (u64).
__global__ void __launch_bounds__(1024, 1)
sat_cols_kernel(const uint32_t* __restrict__ scratch, int H, int W, int C, int SP,
                uint32_t* __restrict__ sat_s, unsigned long long* __restrict__ sat_q, uint32_t* __restrict__ sat_q32,
                int64_t sat_pitch)
{
    __shared__ unsigned long long part[32][33];
    const int cx = threadIdx.x, ry = threadIdx.y;
    const int sx = blockIdx.x * 32 + cx;          // SAT column, 0..W
    const int table = blockIdx.y;
    const int64_t plane = (int64_t)H * SP;
    const int rc = (H + 31) / 32;
    const int y0 = ry * rc, y1 = min(H, y0 + rc);
    const bool live = (sx >= 1 && sx <= W);
    const uint32_t* src = scratch + table * plane + (live ? sx - 1 : 0);
    unsigned long long tot = 0;
    if (live) {
#pragma unroll 16
        for (int y = y0; y < y1; ++y) tot += __ldg(src + (int64_t)y * SP);
    }
    part[ry][cx] = tot;
    __syncthreads();
    unsigned long long run = 0;
    for (int k = 0; k < ry; ++k) run += part[k][cx];
    if (sx > W) return;
    const bool is_q = (table == C);
    const int64_t sat_plane = (int64_t)(H + 1) * sat_pitch;
    uint32_t* ds = sat_s + (is_q ? 0 : table * sat_plane);
    if (ry == 0) {                                 // SAT row 0 is all zeros
        if (is_q) { sat_q[sx] = 0ull; sat_q32[sx] = 0u; } else ds[sx] = 0u;
    }
#pragma unroll 16
    for (int y = y0; y < y1; ++y) {
        if (live) run += __ldg(src + (int64_t)y * SP);
        const int64_t o = (int64_t)(y + 1) * sat_pitch + sx;
        if (is_q) { sat_q[o] = run; sat_q32[o] = (uint32_t)run; } else ds[o] = (uint32_t)run;
    }
}

// One block per template: integer sums -> OpenCV's meanStdDev-derived constants.  The packed rows (pitch wp, a
// multiple of 4, zero padded beyond w*C bytes) are read as 32-bit words; byte b of a row belongs to channel b % C,
// which a 0/1 byte mask selects for dp4a (sum) and a 0/255 mask for the squares.
__global__ void __launch_bounds__(1024)
tmpl_stats_kernel(const uint8_t* __restrict__ tmpl, TmplMeta* __restrict__ meta, int C)
{
    TmplMeta& m = meta[blockIdx.x];
    const uint32_t* p = reinterpret_cast<const uint32_t*>(tmpl + m.pix_off);      // pix_off is a multiple of 16
    long long s[MTM_MAX_CH] = {0, 0, 0, 0}, q[MTM_MAX_CH] = {0, 0, 0, 0};
    const int wq = m.wp >> 2, nwords = m.h * wq;
#pragma unroll 4
    for (int k = threadIdx.x; k < nwords; k += blockDim.x) {
        const uint32_t v = __ldg(p + k);
        const int g = k % wq;                      // word index inside its row
        int ch = (4 * g) % C;                      // channel of the word's first byte
        uint32_t ones[MTM_MAX_CH] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
#pragma unroll
            for (int c = 0; c < MTM_MAX_CH; ++c)
                if (c == ch) ones[c] |= 1u << (8 * b);
            ch = (ch + 1 == C) ? 0 : ch + 1;
        }
#pragma unroll
        for (int c = 0; c < MTM_MAX_CH; ++c) {
            s[c] += __dp4a(v, ones[c], 0u);
            q[c] += __dp4a(v & (ones[c] * 255u), v, 0u);
        }
    }
    __shared__ long long red[2 * MTM_MAX_CH][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = 0; c < MTM_MAX_CH; ++c) {
        long long a = s[c], b = q[c];
        for (int d = 16; d; d >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, d);
            b += __shfl_down_sync(0xffffffffu, b, d);
        }
        if (lane == 0) { red[c][wid] = a; red[MTM_MAX_CH + c][wid] = b; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        const double inv_area = 1.0 / ((double)m.h * (double)m.w);
        double norm = 0.0, mean2 = 0.0;
        long long d2_all = 0;
        for (int c = 0; c < MTM_MAX_CH; ++c) {
            long long a = 0, b = 0;
            for (int k = 0; k < nw; ++k) { a += red[c][k]; b += red[MTM_MAX_CH + c][k]; }
            m.isum[c] = (c < C) ? a : 0;
            if (c < C) d2_all += (long long)m.h * m.w * b - a * a;          // exact; > 0 unless the template is constant
            const double mean = (double)a * inv_area;
            const double var = fmax((double)b * inv_area - mean * mean, 0.0);
            m.mean[c] = (c < C) ? mean : 0.0;
            if (c < C) { norm += var; mean2 += mean * mean; }
        }
        const double sum2 = norm + mean2;
        m.inv_sqrt_d2 = d2_all > 0 ? (float)(1.0 / sqrt((double)d2_all)) : 0.0f;
        m.inv_area = inv_area;
        m.is_const = norm < 2.220446049250313e-16 ? 1 : 0;      // DBL_EPSILON
        m.sum2 = sum2 / inv_area;
        m.norm_ccoeff = sqrt(norm) / sqrt(inv_area);
        m.norm_plain = sqrt(sum2) / sqrt(inv_area);
    }
}

}  // namespace

int launch_build_sat(mtm_ctx* ctx)
{
    ImageDev& im = ctx->img;
    const int H = im.H, W = im.W, C = im.C;
    const int SP = (W + 3) / 4 * 4;                    // scratch row pitch (elements), 16-byte aligned rows
    dim3 g1(H), b1(256);
    switch (C) {
        case 1: sat_rows_c1_kernel<<<g1, b1, 0, ctx->stream>>>(im.pix, im.pitch, H, W, SP, ctx->scratch); break;
        case 2: sat_rows_kernel<2><<<g1, b1, 0, ctx->stream>>>(im.pix, im.pitch, H, W, SP, ctx->scratch); break;
        case 3: sat_rows_kernel<3><<<g1, b1, 0, ctx->stream>>>(im.pix, im.pitch, H, W, SP, ctx->scratch); break;
        case 4: sat_rows_kernel<4><<<g1, b1, 0, ctx->stream>>>(im.pix, im.pitch, H, W, SP, ctx->scratch); break;
        default: return mtm_fail(ctx, MTM_ERR_INVALID, "unsupported channel count %d", C);
    }
    MTM_LAUNCH_CHECK(ctx);
    dim3 g2((W + 1 + 31) / 32, C + 1), b2(32, 32);
    sat_cols_kernel<<<g2, b2, 0, ctx->stream>>>(ctx->scratch, H, W, C, SP, im.sat_s, im.sat_q, im.sat_q32, im.sat_pitch);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_tmpl_stats(mtm_ctx* ctx)
{
    tmpl_stats_kernel<<<ctx->n_tmpl, 1024, 0, ctx->stream>>>(ctx->d_tmpl, ctx->d_meta, ctx->tmpl_C);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}
