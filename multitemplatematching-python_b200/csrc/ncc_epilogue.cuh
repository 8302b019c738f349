// ncc_epilogue.cuh -- K4: OpenCV's common_matchTemplate per-pixel normalisation
// (third-party modules/imgproc/src/templmatch.cpp; reached from
// MTM/__init__.py:92) restated for the device.  Inputs are the EXACT integer
// numerator and window sums; all arithmetic is float64 like OpenCV's, the result
// is cast to float32 last.  Used by both numerator kernels (direct and tcgen05).
#pragma once
#include "mtm_internal.cuh"

struct SatView {
    const uint32_t* s;                 // C planes
    const unsigned long long* q;
    int64_t pitch;                     // elements per SAT row
    int64_t plane;                     // elements per channel plane
};

__device__ __forceinline__ uint32_t sat_window_s(const uint32_t* __restrict__ t, int64_t pitch,
                                                 int y, int x, int h, int w)
{
    const uint32_t* a = t + (int64_t)y * pitch + x;
    const uint32_t* b = a + (int64_t)h * pitch;
    return b[w] - a[w] - b[0] + a[0];            // modulo 2^32, exact for window sums < 2^32
}

__device__ __forceinline__ unsigned long long sat_window_q(const unsigned long long* __restrict__ t,
                                                           int64_t pitch, int y, int x, int h, int w)
{
    const unsigned long long* a = t + (int64_t)y * pitch + x;
    const unsigned long long* b = a + (int64_t)h * pitch;
    return b[w] - a[w] - b[0] + a[0];
}

// cc: sum I*T over all channels.  S[c]: window sum per channel.  Q: window sum of squares.
template <int C>
__device__ __forceinline__ float ncc_epilogue_f64(int method, double cc, const double (&S)[C], double Q, const TmplMeta& tm)
{
    if (method == MTM_TM_CCORR) return (float)cc;
    double num = cc;
    const bool coeff = (method == MTM_TM_CCOEFF || method == MTM_TM_CCOEFF_NORMED);
    const bool sqdiff = (method == MTM_TM_SQDIFF || method == MTM_TM_SQDIFF_NORMED);
    const bool normed = (method == MTM_TM_SQDIFF_NORMED || method == MTM_TM_CCORR_NORMED ||
                         method == MTM_TM_CCOEFF_NORMED);
    if (method == MTM_TM_CCOEFF_NORMED && tm.is_const) return 1.0f;
    double wnd_mean2 = 0.0, wnd_sum2 = 0.0;
    if (coeff) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double t = S[c];
            wnd_mean2 += t * t;
            num -= t * tm.mean[c];
        }
        wnd_mean2 *= tm.inv_area;
    }
    if (normed || sqdiff) {
        wnd_sum2 = Q;
        if (sqdiff) num = fmax(wnd_sum2 - 2.0 * num + tm.sum2, 0.0);
    }
    if (normed) {
        const double diff2 = fmax(wnd_sum2 - wnd_mean2, 0.0);
        const double tnorm = coeff ? tm.norm_ccoeff : tm.norm_plain;
        double t;
        if (diff2 <= fmin(0.5, 10.0 * 1.1920928955078125e-07 * wnd_sum2)) t = 0.0;   // FLT_EPSILON
        else t = sqrt(diff2) * tnorm;
        const double a = fabs(num);
        if (a < t) num /= t;
        else if (a < t * 1.125) num = num > 0.0 ? 1.0 : -1.0;
        else num = (method != MTM_TM_SQDIFF_NORMED) ? 0.0 : 1.0;
    }
    return (float)num;
}

// Exact-integer window sums (uint8 images).
template <int C>
__device__ __forceinline__ float ncc_epilogue(int method, double cc, const uint32_t (&S)[C],
                                              unsigned long long Q, const TmplMeta& tm)
{
    double Sd[C];
#pragma unroll
    for (int c = 0; c < C; ++c) Sd[c] = (double)S[c];
    return ncc_epilogue_f64<C>(method, cc, Sd, (double)Q, tm);
}
