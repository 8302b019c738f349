/* oracle/ncc_exact.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Exact cross-correlation numerator  CC(y,x) = sum_{c,dy,dx} I[y+dy][x+dx][c] * T[dy][dx][c]
 * for uint8 inputs with int64 accumulation, i.e. the quantity cv2.matchTemplate
 * (call site /root/reference MTM/__init__.py:92) obtains approximately by fp32
 * DFT/IPP before OpenCV's common_matchTemplate epilogue (restated in
 * oracle/ncc_exact.py).  Layout: HxWxC / hxwxC interleaved, contiguous.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC (see oracle/build.py).
 */
#include <stdint.h>
#include <stddef.h>

int oracle_cc_u8(const uint8_t* img, int H, int W, int C,
                 const uint8_t* tpl, int h, int w,
                 int64_t* out /* (H-h+1) x (W-w+1) */)
{
    if (h > H || w > W || h <= 0 || w <= 0 || C <= 0) return -1;
    const int mh = H - h + 1, mw = W - w + 1;
    const size_t irow = (size_t)W * C, trow = (size_t)w * C;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < mh; ++y) {
        for (int x = 0; x < mw; ++x) {
            int64_t acc = 0;
            for (int dy = 0; dy < h; ++dy) {
                const uint8_t* ip = img + (size_t)(y + dy) * irow + (size_t)x * C;
                const uint8_t* tp = tpl + (size_t)dy * trow;
                uint32_t racc = 0;                 /* <= 255*255*w*C fits for w*C < 66051 */
                if (trow < 66000) {
                    for (size_t k = 0; k < trow; ++k) racc += (uint32_t)ip[k] * (uint32_t)tp[k];
                    acc += racc;
                } else {
                    for (size_t k = 0; k < trow; ++k) acc += (int64_t)ip[k] * (int64_t)tp[k];
                }
            }
            out[(size_t)y * mw + x] = acc;
        }
    }
    return 0;
}

/* float32 inputs, float64 accumulation (oracle for the non-uint8 dtype policy,
 * MTM/__init__.py:71-74). */
int oracle_cc_f32(const float* img, int H, int W, int C,
                  const float* tpl, int h, int w,
                  double* out)
{
    if (h > H || w > W || h <= 0 || w <= 0 || C <= 0) return -1;
    const int mh = H - h + 1, mw = W - w + 1;
    const size_t irow = (size_t)W * C, trow = (size_t)w * C;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < mh; ++y) {
        for (int x = 0; x < mw; ++x) {
            double acc = 0.0;
            for (int dy = 0; dy < h; ++dy) {
                const float* ip = img + (size_t)(y + dy) * irow + (size_t)x * C;
                const float* tp = tpl + (size_t)dy * trow;
                double racc = 0.0;
                for (size_t k = 0; k < trow; ++k) racc += (double)ip[k] * (double)tp[k];
                acc += racc;
            }
            out[(size_t)y * mw + x] = acc;
        }
    }
    return 0;
}
