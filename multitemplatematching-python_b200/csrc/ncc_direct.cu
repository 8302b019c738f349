// ncc_direct.cu -- K2 + K4: direct sliding-window correlation for uint8 inputs.
//
// Replaces the numerator of cv2.matchTemplate (MTM/__init__.py:92; OpenCV does it
// with an fp32 DFT / IPP) by an EXACT integer correlation: each thread owns a
// 4 (x) x 2 (y) x TT (templates) register tile of u32 accumulators and feeds
// byte-shifted image words and packed template words to dp4a (4 u8 MACs per
// lane-instruction).  Image rows and template rows are staged in shared memory
// per chunk of CH template rows.  The OpenCV normalisation (ncc_epilogue.cuh)
// is fused: the fp32 score is the only thing written to HBM.
//
// Roofline: compute-bound on the integer pipe (dp4a); algorithmic HBM traffic is
// the image tile reads (L2-resident) + 4 B per score pixel.
#include "mtm_internal.cuh"
#include "ncc_epilogue.cuh"

namespace {

constexpr int BX = 64, BY = 32, TXN = 16, NTHREADS = 256;

struct DirectParams {
    const uint8_t* img; int64_t pitch; int H, W;
    SatView sat;
    const uint8_t* tmpl; const TmplMeta* meta;
    const int32_t* order;                          // template indices of this group (already offset)
    float* maps;
    int count;                                     // templates in the group
    int h, w, wp, mh, mw;
    int method;
    int CH;                                        // template rows per chunk
    int TW;                                        // smem tile pitch in 32-bit words
    int wide;                                      // accumulate chunks into u64
};

template <int C, int TT>
__global__ void __launch_bounds__(NTHREADS, (TT >= 8) ? 1 : 2)
ncc_direct_u8_kernel(const DirectParams p)
{
    extern __shared__ uint32_t smem[];
    const int TW = p.TW, CH = p.CH, wq = p.wp >> 2;
    uint32_t* tile = smem;                                   // [(BY + CH)][TW]
    uint32_t* tws = smem + (BY + CH) * TW;                   // [TT][CH][wq]

    const int tid = threadIdx.x, tx = tid & (TXN - 1), ty = tid >> 4;
    const int bx0 = blockIdx.x * BX, by0 = blockIdx.y * BY, t0 = blockIdx.z * TT;
    constexpr int NW = (3 * C + 3) / 4 + 1;                  // image words needed per (row, g)

    uint32_t acc[TT][2][4];
    unsigned long long acc64[TT][2][4];
#pragma unroll
    for (int t = 0; t < TT; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[t][j][i] = 0u; acc64[t][j][i] = 0ull; }

    const int tile_rows = BY + CH - 1;
    for (int c0 = 0; c0 < p.h; c0 += CH) {
        __syncthreads();
        // ---- stage image rows [by0+c0, by0+c0+tile_rows) x TW words ----
        for (int idx = tid; idx < tile_rows * TW; idx += NTHREADS) {
            const int r = idx / TW, k = idx - r * TW;
            const int gy = by0 + c0 + r;
            const int64_t gb = (int64_t)bx0 * C + 4 * k;
            uint32_t v = 0u;
            if (gy < p.H && gb + 4 <= p.pitch)
                v = *reinterpret_cast<const uint32_t*>(p.img + (int64_t)gy * p.pitch + gb);
            tile[r * TW + k] = v;
        }
        // ---- stage template rows [c0, c0+CH) of the TT templates ----
        for (int idx = tid; idx < TT * CH * wq; idx += NTHREADS) {
            const int t = idx / (CH * wq), rem = idx - t * (CH * wq);
            const int r = rem / wq, g = rem - r * wq;
            uint32_t v = 0u;
            if (t0 + t < p.count && c0 + r < p.h)
                v = *reinterpret_cast<const uint32_t*>(p.tmpl + p.meta[p.order[t0 + t]].pix_off +
                                                       (int64_t)(c0 + r) * p.wp + 4 * g);
            tws[idx] = v;
        }
        __syncthreads();

        const uint32_t* trow0 = tile + (2 * ty) * TW + tx * C;
        for (int g = 0; g < wq; ++g) {
            uint32_t tprev[TT];
#pragma unroll
            for (int t = 0; t < TT; ++t) tprev[t] = 0u;
            for (int j = 0; j <= CH; ++j) {
                uint32_t Wd[NW];
                const uint32_t* src = trow0 + j * TW + g;
#pragma unroll
                for (int k = 0; k < NW; ++k) Wd[k] = src[k];
                uint32_t s[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int a = (i * C) / 4, sh = (i * C) % 4;
                    s[i] = (sh == 0) ? Wd[a] : __funnelshift_r(Wd[a], Wd[a + 1], 8 * sh);
                }
                uint32_t tcur[TT];
#pragma unroll
                for (int t = 0; t < TT; ++t) tcur[t] = (j < CH) ? tws[(t * CH + j) * wq + g] : 0u;
#pragma unroll
                for (int t = 0; t < TT; ++t) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[t][0][i] = __dp4a(s[i], tcur[t], acc[t][0][i]);
                        acc[t][1][i] = __dp4a(s[i], tprev[t], acc[t][1][i]);
                    }
                    tprev[t] = tcur[t];
                }
            }
        }
        if (p.wide) {
#pragma unroll
            for (int t = 0; t < TT; ++t)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) { acc64[t][j][i] += acc[t][j][i]; acc[t][j][i] = 0u; }
        }
    }

    // ---- fused normalisation + store ----
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int y = by0 + 2 * ty + j;
        if (y >= p.mh) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = bx0 + 4 * tx + i;
            if (x >= p.mw) continue;
            uint32_t S[C];
#pragma unroll
            for (int c = 0; c < C; ++c)
                S[c] = sat_window_s(p.sat.s + c * p.sat.plane, p.sat.pitch, y, x, p.h, p.w);
            const unsigned long long Q = sat_window_q(p.sat.q, p.sat.pitch, y, x, p.h, p.w);
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t0 + t >= p.count) break;
                const TmplMeta& tm = p.meta[p.order[t0 + t]];
                const double cc = p.wide ? (double)acc64[t][j][i] : (double)acc[t][j][i];
                p.maps[tm.map_off + (int64_t)y * p.mw + x] = ncc_epilogue<C>(p.method, cc, S, Q, tm);
            }
        }
    }
}

template <int C, int TT>
int launch_one(mtm_ctx* ctx, const DirectParams& p, dim3 grid, size_t smem)
{
    auto kern = ncc_direct_u8_kernel<C, TT>;
    MTM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NTHREADS, smem, ctx->stream>>>(p);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

template <int C>
int dispatch_tt(mtm_ctx* ctx, const DirectParams& p, int TT, dim3 grid, size_t smem)
{
    switch (TT) {
        case 1: return launch_one<C, 1>(ctx, p, grid, smem);
        case 2: return launch_one<C, 2>(ctx, p, grid, smem);
        case 4: return launch_one<C, 4>(ctx, p, grid, smem);
        default: return launch_one<C, 8>(ctx, p, grid, smem);
    }
}

}  // namespace

// Score maps of templates d_order[first .. first+count) (all of size h x w) with the direct kernel.
int launch_ncc_direct(mtm_ctx* ctx, int method, int first, int count)
{
    const ImageDev& im = ctx->img;
    const TmplMeta& m0 = ctx->h_meta[ctx->h_order[first]];
    DirectParams p{};
    p.img = im.pix; p.pitch = im.pitch; p.H = im.H; p.W = im.W;
    p.sat.s = im.sat_s; p.sat.q = im.sat_q; p.sat.pitch = im.sat_pitch;
    p.sat.plane = (int64_t)(im.H + 1) * im.sat_pitch;
    p.tmpl = ctx->d_tmpl; p.meta = ctx->d_meta; p.order = ctx->d_order + first; p.maps = ctx->d_maps;
    p.count = count; p.h = m0.h; p.w = m0.w; p.wp = m0.wp; p.mh = m0.mh; p.mw = m0.mw;
    p.method = method;
    const int C = im.C;

    int TT = count >= 8 ? 8 : count >= 4 ? 4 : count >= 2 ? 2 : 1;
    if (C > 1 && TT > 4) TT = 4;
    // smem tile pitch: covers BX*C + wp bytes (+ slack words), == 8 mod 16 words (bank spread)
    int TW = (BX * C + p.wp) / 4 + 4;
    TW = ((TW + 15) / 16) * 16 + 8;
    const size_t budget = 200 * 1024;
    int CH = p.h < 16 ? p.h : 16;
    auto need = [&](int tt, int ch) { return (size_t)(BY + ch) * TW * 4 + (size_t)tt * ch * p.wp; };
    while (need(TT, CH) > budget && CH > 1) CH = (CH + 1) / 2;
    while (need(TT, CH) > budget && TT > 1) TT /= 2;
    if (need(TT, CH) > budget)
        return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "template row of %d bytes does not fit the direct kernel", p.wp);
    p.CH = CH; p.TW = TW;
    // u32 accumulators are exact while 255*255*h*w*C < 2^32
    p.wide = ((double)p.h * p.w * C * 65025.0 >= 4294967296.0) ? 1 : 0;
    if (p.wide && (double)CH * p.wp * 65025.0 >= 4294967296.0)
        return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "template chunk overflows 32-bit accumulators");
    dim3 grid((p.mw + BX - 1) / BX, (p.mh + BY - 1) / BY, (count + TT - 1) / TT);
    const size_t smem = need(TT, CH);
    switch (C) {
        case 1: return dispatch_tt<1>(ctx, p, TT, grid, smem);
        case 3: return dispatch_tt<3>(ctx, p, TT, grid, smem);
        case 4: return dispatch_tt<4>(ctx, p, TT, grid, smem);
    }
    return mtm_fail(ctx, MTM_ERR_INVALID, "unsupported channel count %d", C);
}
