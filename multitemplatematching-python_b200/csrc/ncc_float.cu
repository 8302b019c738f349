// ncc_float.cu -- the float32 branch of MTM's dtype policy (MTM/__init__.py:71-74: unless image
// AND template are uint8, both are cast to float32 -- this is how 16-bit microscopy images reach
// cv2.matchTemplate at :92).  OpenCV computes this case with a float64 DFT and float64 integral
// images; here:
//   * window statistics: float64 summed-area tables (sat_rows_f32 / sat_cols_f64),
//   * numerator: direct FFMA correlation in fp32 against the MEAN-CENTRED template for the
//     TM_CCOEFF* methods (sum I*(T - mean_T) == CC - S*mean_T exactly, so the catastrophic
//     cancellation of the textbook form never happens and fp32 accumulation is accurate to ~1e-6),
//   * OpenCV's float64 epilogue (ncc_epilogue.cuh) fused into the kernel.
// Bound: fp32 FMA pipe.  A tensor-core (bf16x3 / tf32) variant is the next step.
#include "mtm_internal.cuh"
#include "ncc_epilogue.cuh"
#include <algorithm>

namespace {

// ------------------------------------------------------------------ float64 summed-area tables
template <int C>
__global__ void __launch_bounds__(256)
satf_rows_kernel(const float* __restrict__ img, int64_t pitch_e, int H, int W, double* __restrict__ scratch)
{
    __shared__ double wsum[C + 1][8];
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* row = img + (int64_t)y * pitch_e;
    const int64_t plane = (int64_t)H * W;
    const int per = (W + 255) / 256;
    const int xa = min(W, tid * per), xb = min(W, xa + per);
    double tot[C + 1];
#pragma unroll
    for (int c = 0; c <= C; ++c) tot[c] = 0.0;
    for (int x = xa; x < xb; ++x) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double p = (double)row[(int64_t)x * C + c];
            tot[c] += p;
            tot[C] += p * p;
        }
    }
    double excl[C + 1];
#pragma unroll
    for (int c = 0; c <= C; ++c) {
        double s = tot[c];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double n = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += n;
        }
        if (lane == 31) wsum[c][wid] = s;
        excl[c] = s - tot[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c <= C; ++c) {
        double base = 0.0;
        for (int k = 0; k < wid; ++k) base += wsum[c][k];
        excl[c] += base;
    }
    for (int x = xa; x < xb; ++x) {
        double sq = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double p = (double)row[(int64_t)x * C + c];
            excl[c] += p;
            sq += p * p;
            scratch[c * plane + (int64_t)y * W + x] = excl[c];
        }
        excl[C] += sq;
        scratch[C * plane + (int64_t)y * W + x] = excl[C];
    }
}

// tables 0..C-1 -> satf_s[c], table C -> satf_q; all (H+1) x sat_pitch doubles
__global__ void __launch_bounds__(1024, 1)
satf_cols_kernel(const double* __restrict__ scratch, int H, int W, int C, double* __restrict__ sat_s,
                 double* __restrict__ sat_q, int64_t sat_pitch)
{
    __shared__ double part[32][33];
    const int cx = threadIdx.x, ry = threadIdx.y;
    const int sx = blockIdx.x * 32 + cx;
    const int table = blockIdx.y;
    const int64_t plane = (int64_t)H * W;
    const int rc = (H + 31) / 32;
    const int y0 = ry * rc, y1 = min(H, y0 + rc);
    const bool live = (sx >= 1 && sx <= W);
    const double* src = scratch + table * plane + (live ? sx - 1 : 0);
    double tot = 0.0;
    if (live) {
#pragma unroll 8
        for (int y = y0; y < y1; ++y) tot += src[(int64_t)y * W];
    }
    part[ry][cx] = tot;
    __syncthreads();
    double run = 0.0;
    for (int k = 0; k < ry; ++k) run += part[k][cx];
    if (sx > W) return;
    const int64_t sat_plane = (int64_t)(H + 1) * sat_pitch;
    double* dst = (table == C) ? sat_q : sat_s + table * sat_plane;
    if (ry == 0) dst[sx] = 0.0;
#pragma unroll 8
    for (int y = y0; y < y1; ++y) {
        if (live) run += src[(int64_t)y * W];
        dst[(int64_t)(y + 1) * sat_pitch + sx] = run;
    }
}

// One block per template: OpenCV meanStdDev constants (float64) and the mean-centred copy.
__global__ void tmplf_stats_kernel(const float* __restrict__ tmpl, float* __restrict__ centred,
                                   TmplMeta* __restrict__ meta, int C)
{
    TmplMeta& m = meta[blockIdx.x];
    const float* p = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(tmpl) + m.pix_off);
    float* q = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(centred) + m.pix_off);
    const int rowe = m.wp >> 2;                                  // elements per packed row (w*C)
    double s[MTM_MAX_CH] = {0, 0, 0, 0}, sq[MTM_MAX_CH] = {0, 0, 0, 0};
    const int n = m.h * m.w;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int y = i / m.w, x = i - y * m.w;
        for (int c = 0; c < C; ++c) {
            const double v = (double)p[(int64_t)y * rowe + x * C + c];
            s[c] += v;
            sq[c] += v * v;
        }
    }
    __shared__ double red[2 * MTM_MAX_CH][32];
    __shared__ double mean_sh[MTM_MAX_CH];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = 0; c < MTM_MAX_CH; ++c) {
        double a = s[c], b = sq[c];
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
        for (int c = 0; c < MTM_MAX_CH; ++c) {
            double a = 0, b = 0;
            for (int k = 0; k < nw; ++k) { a += red[c][k]; b += red[MTM_MAX_CH + c][k]; }
            const double mean = a * inv_area;
            const double var = fmax(b * inv_area - mean * mean, 0.0);
            m.mean[c] = (c < C) ? mean : 0.0;
            mean_sh[c] = m.mean[c];
            if (c < C) { norm += var; mean2 += mean * mean; }
        }
        const double sum2 = norm + mean2;
        m.inv_area = inv_area;
        m.is_const = norm < 2.220446049250313e-16 ? 1 : 0;
        m.sum2 = sum2 / inv_area;
        m.norm_ccoeff = sqrt(norm) / sqrt(inv_area);
        m.norm_plain = sqrt(sum2) / sqrt(inv_area);
        for (int c = 0; c < MTM_MAX_CH; ++c) m.isum[c] = 0;
        m.inv_sqrt_d2 = 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int y = i / m.w, x = i - y * m.w;
        for (int c = 0; c < C; ++c) {
            const int64_t o = (int64_t)y * rowe + x * C + c;
            q[o] = (float)((double)p[o] - mean_sh[c]);
        }
    }
}

// ------------------------------------------------------------------ direct fp32 correlation
constexpr int FBX = 64, FBY = 32, FTHREADS = 256;

struct FloatParams {
    const float* img; int64_t pitch_e; int H, W;
    const double* sat_s; const double* sat_q; int64_t sat_pitch, sat_plane;
    const float* tmpl;                                 // raw or centred template arena
    const TmplMeta* meta; const int32_t* order; float* maps;
    int count, h, w, mh, mw, method, CH, TW, centred;
};

__device__ __forceinline__ double satf_window(const double* __restrict__ t, int64_t pitch, int y, int x, int h, int w)
{
    const double* a = t + (int64_t)y * pitch + x;
    const double* b = a + (int64_t)h * pitch;
    return b[w] - a[w] - b[0] + a[0];
}

// Thread tile: XO x-outputs (4 for single channel, 1 otherwise) x 2 rows x TT templates.
template <int C, int TT>
__global__ void __launch_bounds__(FTHREADS, 2)
ncc_direct_f32_kernel(const FloatParams p)
{
    constexpr int XO = (C == 1) ? 4 : 1;
    constexpr int BX = (C == 1) ? FBX : 16;
    extern __shared__ float smemf[];
    const int TW = p.TW, CH = p.CH, we = p.w * C;                  // template row elements
    float* tile = smemf;                                           // [(FBY + CH)][TW]
    float* tws = smemf + (FBY + CH) * TW;                          // [TT][CH][we4]
    const int we4 = (we + 3) & ~3;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int bx0 = blockIdx.x * BX, by0 = blockIdx.y * FBY, t0 = blockIdx.z * TT;
    float acc[TT][2][XO];
#pragma unroll
    for (int t = 0; t < TT; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < XO; ++i) acc[t][j][i] = 0.f;

    const int tile_rows = FBY + CH - 1;
    for (int c0 = 0; c0 < p.h; c0 += CH) {
        __syncthreads();
        for (int idx = tid; idx < tile_rows * TW; idx += FTHREADS) {
            const int r = idx / TW, k = idx - r * TW;
            const int gy = by0 + c0 + r;
            const int64_t ge = (int64_t)bx0 * C + k;
            tile[idx] = (gy < p.H && ge < (int64_t)p.W * C) ? p.img[(int64_t)gy * p.pitch_e + ge] : 0.f;
        }
        for (int idx = tid; idx < TT * CH * we4; idx += FTHREADS) {
            const int t = idx / (CH * we4), rem = idx - t * (CH * we4);
            const int r = rem / we4, e = rem - r * we4;
            float v = 0.f;
            if (t0 + t < p.count && c0 + r < p.h && e < we) {
                const TmplMeta& tm = p.meta[p.order[t0 + t]];
                v = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.tmpl) + tm.pix_off)[(int64_t)(c0 + r) * (tm.wp >> 2) + e];
            }
            tws[idx] = v;
        }
        __syncthreads();

        const float* trow0 = tile + (2 * ty) * TW + (XO * tx) * C;
        for (int j = 0; j <= CH; ++j) {                           // tile row 2*ty + j feeds out row 0 (T row j) and out row 1 (T row j-1)
            const float* src = trow0 + j * TW;
            const float* ta = tws + j * we4;                       // template row j of template 0
            const float* tb = tws + (j - 1) * we4;                 // template row j-1
            if constexpr (C == 1) {
                float v0 = src[0], v1 = src[1], v2 = src[2], v3;
                for (int e = 0; e < we4; e += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v3 = src[e + u + 3];
#pragma unroll
                        for (int t = 0; t < TT; ++t) {
                            const float a = (j < CH) ? ta[t * CH * we4 + e + u] : 0.f;
                            const float b = (j >= 1) ? tb[t * CH * we4 + e + u] : 0.f;
                            acc[t][0][0] = fmaf(v0, a, acc[t][0][0]); acc[t][1][0] = fmaf(v0, b, acc[t][1][0]);
                            acc[t][0][1] = fmaf(v1, a, acc[t][0][1]); acc[t][1][1] = fmaf(v1, b, acc[t][1][1]);
                            acc[t][0][2] = fmaf(v2, a, acc[t][0][2]); acc[t][1][2] = fmaf(v2, b, acc[t][1][2]);
                            acc[t][0][3] = fmaf(v3, a, acc[t][0][3]); acc[t][1][3] = fmaf(v3, b, acc[t][1][3]);
                        }
                        v0 = v1; v1 = v2; v2 = v3;
                    }
                }
            } else {
                for (int e = 0; e < we; ++e) {
                    const float v = src[e];
#pragma unroll
                    for (int t = 0; t < TT; ++t) {
                        const float a = (j < CH) ? ta[t * CH * we4 + e] : 0.f;
                        const float b = (j >= 1) ? tb[t * CH * we4 + e] : 0.f;
                        acc[t][0][0] = fmaf(v, a, acc[t][0][0]);
                        acc[t][1][0] = fmaf(v, b, acc[t][1][0]);
                    }
                }
            }
        }
    }

    // ---- OpenCV epilogue in float64 (window sums from the float64 SATs)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int y = by0 + 2 * ty + j;
        if (y >= p.mh) continue;
#pragma unroll
        for (int i = 0; i < XO; ++i) {
            const int x = bx0 + XO * tx + i;
            if (x >= p.mw) continue;
            double S[C];
            double Q = 0.0;
            if (p.method != MTM_TM_CCORR) {                       // plain correlation needs no window statistics
#pragma unroll
                for (int c = 0; c < C; ++c) S[c] = satf_window(p.sat_s + c * p.sat_plane, p.sat_pitch, y, x, p.h, p.w);
                Q = satf_window(p.sat_q, p.sat_pitch, y, x, p.h, p.w);
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) S[c] = 0.0;
            }
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t0 + t >= p.count) break;
                const TmplMeta& tm = p.meta[p.order[t0 + t]];
                double cc = (double)acc[t][j][i];
                if (p.centred) {                                  // acc == CC - sum_c S_c*mean_c already: undo for the shared epilogue
#pragma unroll
                    for (int c = 0; c < C; ++c) cc += S[c] * tm.mean[c];
                }
                p.maps[tm.map_off + (int64_t)y * p.mw + x] = ncc_epilogue_f64<C>(p.method, cc, S, Q, tm);
            }
        }
    }
}

template <int C, int TT>
int launch_f(mtm_ctx* ctx, const FloatParams& p, dim3 grid, size_t smem)
{
    auto kern = ncc_direct_f32_kernel<C, TT>;
    MTM_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, FTHREADS, smem, ctx->stream>>>(p);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

template <int C>
int dispatch_f(mtm_ctx* ctx, const FloatParams& p, int TT, dim3 grid, size_t smem)
{
    switch (TT) {
        case 1: return launch_f<C, 1>(ctx, p, grid, smem);
        case 2: return launch_f<C, 2>(ctx, p, grid, smem);
        default: return launch_f<C, 4>(ctx, p, grid, smem);
    }
}

// ------------------------------------------------------------------ masked matching (methods 0 / 3)
// OpenCV matchTemplateMask (third-party; reached from MTM/__init__.py:92 with mask=...):
//   TM_SQDIFF        R = sum I^2 M^2 - 2 sum I (T M^2) + sum (T M)^2
//   TM_CCORR_NORMED  R = sum I (T M^2) / sqrt( sum (T M)^2 * sum I^2 M^2 )
// uint8 masks are binary (non-zero -> 1), float32 masks are weights; everything is float32.

// raw template + mask (u8 or f32) -> T*M^2 into `tm2`, M^2 into `m2` (float layouts of the template arena);
// meta.sum2 := sum (T*M)^2 in float64.  One block per template.
__global__ void masked_prep_kernel(const uint8_t* __restrict__ raw_t, const uint8_t* __restrict__ raw_m, int is_f32,
                                   float* __restrict__ tm2, float* __restrict__ m2, TmplMeta* __restrict__ meta, int C)
{
    TmplMeta& m = meta[blockIdx.x];
    const int rowe = m.wp >> 2, n = m.h * m.w * C, we = m.w * C;
    const int64_t raw_off = m.pix_off / 4 * (is_f32 ? 4 : 1);     // raw arrays are packed with the element size
    float* q1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tm2) + m.pix_off);
    float* q2 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(m2) + m.pix_off);
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int y = i / we, e = i - y * we;
        float t, k;
        if (is_f32) {
            t = reinterpret_cast<const float*>(raw_t + raw_off)[(int64_t)y * we + e];
            k = reinterpret_cast<const float*>(raw_m + raw_off)[(int64_t)y * we + e];
        } else {
            t = (float)raw_t[raw_off + (int64_t)y * we + e];
            k = raw_m[raw_off + (int64_t)y * we + e] ? 1.0f : 0.0f;
        }
        const float tm = t * k;                                    // templ.mul(mask)
        q1[(int64_t)y * rowe + e] = t * (k * k);                   // templ.mul(mask.mul(mask))
        q2[(int64_t)y * rowe + e] = k * k;
        acc += (double)tm * (double)tm;
    }
    __shared__ double red[32];
    for (int d = 16; d; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += red[k];
        m.sum2 = tot;                                              // templ2_mask2_sum
        m.is_const = 0; m.inv_area = 1.0 / ((double)m.h * m.w);
    }
}

__global__ void u8_to_f32_sq_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, const float* __restrict__ srcf,
                                    float* __restrict__ dst, float* __restrict__ dst2, int64_t pitch_e, int H, int WE)
{
    const int64_t n = (int64_t)H * WE;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / WE), e = (int)(i - (int64_t)y * WE);
        const float v = src ? (float)src[(int64_t)y * src_pitch + e] : srcf[(int64_t)y * pitch_e + e];
        if (src) dst[(int64_t)y * pitch_e + e] = v;
        dst2[(int64_t)y * pitch_e + e] = v * v;
    }
}

// A = corr(I, T M^2), B = corr(I^2, M^2) (both fp32 maps) -> the masked score, in place into A.
__global__ void masked_combine_kernel(float* __restrict__ A, const float* __restrict__ B, const TmplMeta* __restrict__ meta,
                                      int method)
{
    const TmplMeta& tm = meta[blockIdx.y];
    const int64_t n = (int64_t)tm.mh * tm.mw;
    const double t2 = tm.sum2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = A[tm.map_off + i], b = B[tm.map_off + i];
        float r;
        if (method == MTM_TM_SQDIFF) r = (float)(-2.0 * (double)a + (double)b + t2);
        else r = a / sqrtf((float)(t2 * (double)b));
        A[tm.map_off + i] = r;
    }
}

}  // namespace

int launch_masked_prep(mtm_ctx* ctx, const uint8_t* d_raw_t, const uint8_t* d_raw_m, int is_f32)
{
    masked_prep_kernel<<<ctx->n_tmpl, 256, 0, ctx->stream>>>(d_raw_t, d_raw_m, is_f32, reinterpret_cast<float*>(ctx->d_tmpl),
                                                           reinterpret_cast<float*>(ctx->d_tmpl_centred), ctx->d_meta, ctx->tmpl_C);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

// float image (and its square) for the masked path; src is the u8 image or (srcf) the float image already resident
int launch_masked_image(mtm_ctx* ctx)
{
    ImageDev& im = ctx->img;
    const int WE = im.W * im.C;
    const int64_t n = (int64_t)im.H * WE;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    const bool from_u8 = (ctx->img_dtype == MTM_U8);
    u8_to_f32_sq_kernel<<<blocks, 256, 0, ctx->stream>>>(from_u8 ? im.pix : nullptr, im.pitch, im.pixf, im.pixf, im.pixf2,
                                                        im.pitch_e, im.H, WE);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_masked_combine(mtm_ctx* ctx, int method, const float* mapsB)
{
    int64_t max_px = 1;
    for (int t = 0; t < ctx->n_tmpl; ++t) max_px = std::max<int64_t>(max_px, (int64_t)ctx->h_meta[t].mh * ctx->h_meta[t].mw);
    const int blocks = (int)std::min<int64_t>((max_px + 255) / 256, (int64_t)ctx->sm_count * 8);
    masked_combine_kernel<<<dim3(blocks, ctx->n_tmpl), 256, 0, ctx->stream>>>(ctx->d_maps, mapsB, ctx->d_meta, method);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

// ------------------------------------------------------------------ 16-bit images (MTM_U16)
namespace {

// raw u16 rows -> float32 image (statistics, fallback kernels) + high / low byte planes in the u8 tile layout of the
// tensor-core kernel (zero padded pitch).
__global__ void u16_split_image_kernel(const uint16_t* __restrict__ src, int64_t src_stride_e, int H, int W,
                                       float* __restrict__ pixf, int64_t pitch_e, uint8_t* __restrict__ hi,
                                       uint8_t* __restrict__ lo, int64_t pitch)
{
    const int y = blockIdx.y;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
        const uint32_t v = src[(int64_t)y * src_stride_e + x];
        pixf[(int64_t)y * pitch_e + x] = (float)v;
        hi[(int64_t)y * pitch + x] = (uint8_t)(v >> 8);
        lo[(int64_t)y * pitch + x] = (uint8_t)(v & 255u);
    }
}

// float32 image whose pixels may all be integers in [0, 65535]: the same byte planes, and a flag when some pixel is not
__global__ void f32_split_image_kernel(const float* __restrict__ pixf, int64_t pitch_e, int H, int W, uint8_t* __restrict__ hi,
                                       uint8_t* __restrict__ lo, int64_t pitch, int32_t* __restrict__ not_integral)
{
    const int y = blockIdx.y;
    int bad = 0;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
        const float f = pixf[(int64_t)y * pitch_e + x];
        const uint32_t v = (f >= 0.0f && f <= 65535.0f) ? (uint32_t)f : 0u;
        if (!((float)v == f)) bad = 1;                       // negative, too large, fractional, NaN
        hi[(int64_t)y * pitch + x] = (uint8_t)(v >> 8);
        lo[(int64_t)y * pitch + x] = (uint8_t)(v & 255u);
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(not_integral, 1);
}

// exact numerator (double) + float64 window statistics -> OpenCV's epilogue, any method.  blockIdx.y = template.
__global__ void cc16_epilogue_kernel(const double* __restrict__ acc, float* __restrict__ maps, const TmplMeta* __restrict__ meta,
                                     int tmpl_first, int method, const double* __restrict__ sat_s, const double* __restrict__ sat_q,
                                     int64_t sat_pitch)
{
    const TmplMeta& tm = meta[tmpl_first + blockIdx.y];
    const int64_t n = (int64_t)tm.mh * tm.mw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(idx / tm.mw), x = (int)(idx - (int64_t)y * tm.mw);
        double S[1] = {0.0}, Q = 0.0;
        if (method != MTM_TM_CCORR) {
            S[0] = satf_window(sat_s, sat_pitch, y, x, tm.h, tm.w);
            Q = satf_window(sat_q, sat_pitch, y, x, tm.h, tm.w);
        }
        maps[tm.map_off + idx] = ncc_epilogue_f64<1>(method, acc[tm.map_off + idx], S, Q, tm);
    }
}

}  // namespace

int launch_u16_split_image(mtm_ctx* ctx, const uint16_t* src, int64_t src_stride_bytes)
{
    ImageDev& im = ctx->img;
    dim3 grid((unsigned)std::min(8, (im.W + 255) / 256), (unsigned)im.H);
    u16_split_image_kernel<<<grid, 256, 0, ctx->stream>>>(src, src_stride_bytes / 2, im.H, im.W, im.pixf, im.pitch_e, im.pix, im.pix_lo, im.pitch);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_f32_split_image(mtm_ctx* ctx, int* not_integral)
{
    ImageDev& im = ctx->img;
    int32_t* flag = ctx->d_cand_count + 8;                   // a spare word of the 64-byte counter block
    MTM_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int32_t), ctx->stream));
    dim3 grid((unsigned)std::min(8, (im.W + 255) / 256), (unsigned)im.H);
    f32_split_image_kernel<<<grid, 256, 0, ctx->stream>>>(im.pixf, im.pitch_e, im.H, im.W, im.pix, im.pix_lo, im.pitch, flag);
    MTM_LAUNCH_CHECK(ctx);
    int32_t h = 1;
    MTM_CUDA(ctx, cudaMemcpyAsync(&h, flag, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->ctr.d2h_bytes += (int64_t)sizeof h;
    *not_integral = h;
    return MTM_OK;
}

int launch_cc16_epilogue(mtm_ctx* ctx, int method, int tmpl)
{
    const ImageDev& im = ctx->img;
    int64_t n = 0;
    const int first = tmpl < 0 ? 0 : tmpl, count = tmpl < 0 ? ctx->n_tmpl : 1;
    for (int t = first; t < first + count; ++t) n = std::max<int64_t>(n, (int64_t)ctx->h_meta[t].mh * ctx->h_meta[t].mw);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8));
    cc16_epilogue_kernel<<<dim3(blocks, count), 256, 0, ctx->stream>>>(ctx->d_acc, ctx->d_maps, ctx->d_meta, first, method, im.satf_s,
                                                                       im.satf_q, im.sat_pitch);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_build_sat_f32(mtm_ctx* ctx)
{
    ImageDev& im = ctx->img;
    const int H = im.H, W = im.W, C = im.C;
    double* scratch = reinterpret_cast<double*>(ctx->scratch);
    switch (C) {
        case 1: satf_rows_kernel<1><<<H, 256, 0, ctx->stream>>>(im.pixf, im.pitch_e, H, W, scratch); break;
        case 3: satf_rows_kernel<3><<<H, 256, 0, ctx->stream>>>(im.pixf, im.pitch_e, H, W, scratch); break;
        case 4: satf_rows_kernel<4><<<H, 256, 0, ctx->stream>>>(im.pixf, im.pitch_e, H, W, scratch); break;
        default: return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "float32 images with %d channels", C);
    }
    MTM_LAUNCH_CHECK(ctx);
    dim3 g2((W + 1 + 31) / 32, C + 1), b2(32, 32);
    satf_cols_kernel<<<g2, b2, 0, ctx->stream>>>(scratch, H, W, C, im.satf_s, im.satf_q, im.sat_pitch);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_tmpl_stats_f32(mtm_ctx* ctx)
{
    tmplf_stats_kernel<<<ctx->n_tmpl, 256, 0, ctx->stream>>>(reinterpret_cast<const float*>(ctx->d_tmpl),
                                                            reinterpret_cast<float*>(ctx->d_tmpl_centred), ctx->d_meta, ctx->tmpl_C);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

// Score maps of templates d_order[first .. first+count) (same size) for float32 inputs.
// img_override / tmpl_override / maps_override: the masked path (ncc_masked) runs two plain correlations
// (image x T*M^2 and image^2 x M^2) through this kernel with method TM_CCORR.
int launch_ncc_direct_f32(mtm_ctx* ctx, int method, int first, int count, const float* img_override,
                          const uint8_t* tmpl_override, float* maps_override)
{
    const ImageDev& im = ctx->img;
    const TmplMeta& m0 = ctx->h_meta[ctx->h_order[first]];
    FloatParams p{};
    p.img = img_override ? img_override : im.pixf; p.pitch_e = im.pitch_e; p.H = im.H; p.W = im.W;
    p.sat_s = im.satf_s; p.sat_q = im.satf_q; p.sat_pitch = im.sat_pitch; p.sat_plane = (int64_t)(im.H + 1) * im.sat_pitch;
    p.centred = (!tmpl_override && (method == MTM_TM_CCOEFF || method == MTM_TM_CCOEFF_NORMED)) ? 1 : 0;
    p.tmpl = reinterpret_cast<const float*>(tmpl_override ? tmpl_override : (p.centred ? ctx->d_tmpl_centred : ctx->d_tmpl));
    p.meta = ctx->d_meta; p.order = ctx->d_order + first; p.maps = maps_override ? maps_override : ctx->d_maps;
    p.count = count; p.h = m0.h; p.w = m0.w; p.mh = m0.mh; p.mw = m0.mw; p.method = method;
    const int C = im.C;
    const int we4 = (m0.w * C + 3) & ~3;
    const int BX = (C == 1) ? FBX : 16;
    int TT = count >= 4 ? 4 : count >= 2 ? 2 : 1;
    int TW = BX * C + we4 + 8;
    TW = ((TW + 31) / 32) * 32 + 8;
    const size_t budget = 200 * 1024;
    int CH = m0.h < 8 ? m0.h : 8;
    auto need = [&](int tt, int ch) { return ((size_t)(FBY + ch) * TW + (size_t)tt * ch * we4) * sizeof(float); };
    while (need(TT, CH) > budget && CH > 1) CH = (CH + 1) / 2;
    while (need(TT, CH) > budget && TT > 1) TT /= 2;
    if (need(TT, CH) > budget) return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "float32 template row too wide for the direct kernel");
    p.CH = CH; p.TW = TW;
    dim3 grid((p.mw + BX - 1) / BX, (p.mh + FBY - 1) / FBY, (count + TT - 1) / TT);
    const size_t smem = need(TT, CH);
    switch (C) {
        case 1: return dispatch_f<1>(ctx, p, TT, grid, smem);
        case 3: return dispatch_f<3>(ctx, p, TT, grid, smem);
        case 4: return dispatch_f<4>(ctx, p, TT, grid, smem);
    }
    return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "float32 images with %d channels", C);
}
