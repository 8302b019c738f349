// ncc_points.cu -- K2b: cv2.matchTemplate (MTM/__init__.py:92) for SMALL score maps of LARGE templates (uint8).
//
// A search box barely larger than the template (the re-localisation step of matchTemplatesPyramid, test.py's
// "template as large as the searchBox", MTM/__init__.py:140-144) has a handful of output pixels but tens of
// thousands of products per pixel.  The tiled kernels parallelise over output pixels: the tcgen05 kernel would run
// all template rows of its single tile as one serial chain of MMAs (0.2 ms for a 256 x 256 template), the dp4a
// kernel would leave most threads idle.  Here the parallelism is over the TEMPLATE: one CTA per output pixel,
// its 1024 threads stride over the template's 32-bit words (dp4a against the unaligned image word rebuilt with a
// funnel shift), block reduction, then the same float64 OpenCV epilogue as ncc_direct.cu on the exact integer
// sums -- maps are bit-identical to the dp4a kernel's.
// Bound: L2 bandwidth (every CTA reads the template and its window once: 2*h*w*C bytes per output pixel).
#include "mtm_internal.cuh"
#include "ncc_epilogue.cuh"
#include <cstdlib>

namespace {

constexpr int PT_THREADS = 1024;
constexpr int PT_MAX_PIXELS = 1024;        // largest score map this kernel is chosen for
constexpr int PT_MIN_BYTES = 16384;        // smallest template (h*w*C bytes) it is chosen for

struct PointsParams {
    const uint8_t* img; int64_t pitch;
    SatView sat;
    const uint8_t* tmpl; const TmplMeta* meta;
    const int32_t* order;                  // template indices of this launch
    float* maps;
    int method;
};

template <int C>
__global__ void __launch_bounds__(PT_THREADS)
ncc_points_kernel(const PointsParams p)
{
    __shared__ unsigned long long part[PT_THREADS / 32];
    const TmplMeta& tm = p.meta[p.order[blockIdx.y]];
    const int npos = tm.mh * tm.mw;
    const int wq = tm.wp >> 2, nwords = tm.h * wq;
    const uint8_t* tp = p.tmpl + tm.pix_off;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int pos = blockIdx.x; pos < npos; pos += gridDim.x) {
        const int y = pos / tm.mw, x = pos - y * tm.mw;
        unsigned long long total = 0ull;
#pragma unroll 4
        for (int k = tid; k < nwords; k += PT_THREADS) {
            const int r = k / wq, g = k - r * wq;
            const uint32_t tw = __ldg(reinterpret_cast<const uint32_t*>(tp + (int64_t)r * tm.wp + 4 * g));   // zero padded beyond w*C
            const int64_t a = (int64_t)(y + r) * p.pitch + (int64_t)x * C + 4 * g;
            const uint32_t* iw = reinterpret_cast<const uint32_t*>(p.img + (a & ~(int64_t)3));
            const uint32_t s = __funnelshift_r(__ldg(iw), __ldg(iw + 1), 8 * (int)(a & 3));                    // 4 image bytes from offset a
            total += __dp4a(s, tw, 0u);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) total += __shfl_down_sync(0xffffffffu, total, d);
        if (lane == 0) part[wid] = total;
        __syncthreads();
        if (tid == 0) {
            unsigned long long cc = 0ull;
#pragma unroll
            for (int k = 0; k < PT_THREADS / 32; ++k) cc += part[k];
            uint32_t S[C];
#pragma unroll
            for (int c = 0; c < C; ++c)
                S[c] = sat_window_s(p.sat.s + c * p.sat.plane, p.sat.pitch, y, x, tm.h, tm.w);
            const unsigned long long Q = sat_window_q(p.sat.q, p.sat.pitch, y, x, tm.h, tm.w);
            p.maps[tm.map_off + pos] = ncc_epilogue<C>(p.method, (double)cc, S, Q, tm);
        }
        __syncthreads();
    }
}

}  // namespace

// Should templates d_order[first .. first+count) take this kernel?  Every member must have a small map and a large
// template, the inputs must be plain uint8 and the caller must not have pinned another kernel.
bool points_path_preferred(const mtm_ctx* ctx, int first, int count)
{
    static const bool off = getenv("MTM_B200_NO_POINTS") != nullptr;      // experiments: keep the tiled kernels
    if (off || ctx->path != MTM_PATH_AUTO || ctx->img_dtype != MTM_U8 || ctx->masked) return false;
    const int C = ctx->img.C;
    if (C != 1 && C != 3 && C != 4) return false;
    for (int k = first; k < first + count; ++k) {
        const TmplMeta& m = ctx->h_meta[ctx->h_order[k]];
        if ((int64_t)m.mh * m.mw > PT_MAX_PIXELS || (int64_t)m.h * m.w * C < PT_MIN_BYTES) return false;
        if ((int64_t)m.h * m.w >= (int64_t)16000000) return false;         // window sums must stay below 2^32
    }
    return true;
}

int launch_ncc_points(mtm_ctx* ctx, int method, int first, int count)
{
    const ImageDev& im = ctx->img;
    PointsParams p{};
    p.img = im.pix; p.pitch = im.pitch;
    p.sat.s = im.sat_s; p.sat.q = im.sat_q; p.sat.pitch = im.sat_pitch;
    p.sat.plane = (int64_t)(im.H + 1) * im.sat_pitch;
    p.tmpl = ctx->d_tmpl; p.meta = ctx->d_meta; p.order = ctx->d_order + first; p.maps = ctx->d_maps;
    p.method = method;
    int npos = 1;
    for (int k = first; k < first + count; ++k) {
        const TmplMeta& m = ctx->h_meta[ctx->h_order[k]];
        npos = std::max(npos, m.mh * m.mw);
    }
    const dim3 grid((unsigned)npos, (unsigned)count);
    switch (im.C) {
        case 1: ncc_points_kernel<1><<<grid, PT_THREADS, 0, ctx->stream>>>(p); break;
        case 3: ncc_points_kernel<3><<<grid, PT_THREADS, 0, ctx->stream>>>(p); break;
        case 4: ncc_points_kernel<4><<<grid, PT_THREADS, 0, ctx->stream>>>(p); break;
        default: return mtm_fail(ctx, MTM_ERR_INVALID, "unsupported channel count %d", im.C);
    }
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}
