// transform.cu -- SURVEY §8 f3: what the reference's tutorials do in user code before calling
// matchTemplates, done on the device so that only the base pixels cross PCIe:
//   * template augmentation by the 8 symmetries of the square (np.rot90 k = 1..3, np.fliplr, np.flipud,
//     transpose, anti-transpose; Tutorial2-Template_Augmentation.ipynb cell 15), and
//   * integer-factor area downscale = cv2.resize(src, (W/f, H/f), interpolation=cv2.INTER_AREA)
//     (Tutorial3-SpeedingUp.ipynb cells 17-21) of the image and of the templates.
// One kernel serves both: every output pixel is the f x f box mean of the source, read through the
// index map of the symmetry.  Pure data movement: HBM-bound, reads f*f*C elements and writes C
// elements per output pixel.
//
// Rounding of the box mean follows OpenCV's resizeAreaFast for integer pixels (third-party, pinned
// against the live cv2 in tests/test_oracle.py):  f == 2: (sum + 2) >> 2;  f >= 3:
// cvRound((float)sum * (1.f / (f*f))) (round half to even on the float32 product).  float32 pixels
// are summed in double and rounded once (OpenCV sums in float32; the two agree to ~1e-7 relative).
#include "mtm_internal.cuh"

namespace {

// (i, j) of the transformed array -> (si, sj) of the dh x dw source array.
__device__ __forceinline__ void xform_source_index(int op, int i, int j, int dh, int dw, int& si, int& sj)
{
    switch (op) {
    default:
    case MTM_XF_IDENTITY:      si = i;          sj = j;          break;
    case MTM_XF_ROT90:         si = j;          sj = dw - 1 - i; break;   // np.rot90(m, 1)
    case MTM_XF_ROT180:        si = dh - 1 - i; sj = dw - 1 - j; break;   // np.rot90(m, 2)
    case MTM_XF_ROT270:        si = dh - 1 - j; sj = i;          break;   // np.rot90(m, 3)
    case MTM_XF_FLIPLR:        si = i;          sj = dw - 1 - j; break;   // np.fliplr(m)
    case MTM_XF_FLIPUD:        si = dh - 1 - i; sj = j;          break;   // np.flipud(m)
    case MTM_XF_TRANSPOSE:     si = j;          sj = i;          break;   // m.swapaxes(0, 1)
    case MTM_XF_ANTITRANSPOSE: si = dh - 1 - j; sj = dw - 1 - i; break;   // np.rot90(m, 2).swapaxes(0, 1)
    }
}

template <typename T> struct BoxMean;
template <> struct BoxMean<uint8_t> {
    using Acc = uint32_t;
    static __device__ __forceinline__ uint8_t finish(uint32_t s, int f, float scale) {
        if (f == 1) return (uint8_t)s;
        if (f == 2) return (uint8_t)((s + 2u) >> 2);
        return (uint8_t)__float2int_rn(__fmul_rn((float)s, scale));
    }
};
template <> struct BoxMean<uint16_t> {
    using Acc = uint32_t;                     // f <= 16: sum <= 256 * 65535 < 2^24, exact in float32 as well
    static __device__ __forceinline__ uint16_t finish(uint32_t s, int f, float scale) {
        if (f == 1) return (uint16_t)s;
        if (f == 2) return (uint16_t)((s + 2u) >> 2);
        return (uint16_t)__float2int_rn(__fmul_rn((float)s, scale));
    }
};
template <> struct BoxMean<float> {
    using Acc = double;
    static __device__ __forceinline__ float finish(double s, int f, float) {
        return f == 1 ? (float)s : (float)(s / (double)(f * f));
    }
};

// grid = (pixel blocks, outputs); one thread per output pixel (all channels).
template <typename T>
__global__ void __launch_bounds__(256)
transform_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const XformDesc* __restrict__ descs,
                 int C, int f, float scale)
{
    const XformDesc d = descs[blockIdx.y];
    const T* s0 = reinterpret_cast<const T*>(src + d.src_off);
    const int64_t sp = d.src_pitch / (int64_t)sizeof(T);              // source row pitch in elements
    const int64_t npix = (int64_t)d.oh * d.ow;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / d.ow), j = (int)(p - (int64_t)i * d.ow);
        int si, sj;
        xform_source_index(d.op, i, j, d.dh, d.dw, si, sj);
        T* out = reinterpret_cast<T*>(dst + d.dst_off + (int64_t)i * d.dst_pitch) + (int64_t)j * C;
        const T* box = s0 + (int64_t)si * f * sp + (int64_t)sj * f * C;
        for (int c = 0; c < C; ++c) {
            typename BoxMean<T>::Acc acc = 0;
            for (int dy = 0; dy < f; ++dy) {
                const T* row = box + (int64_t)dy * sp + c;
                for (int dx = 0; dx < f; ++dx) acc += row[(int64_t)dx * C];
            }
            out[c] = BoxMean<T>::finish(acc, f, scale);
        }
    }
}

}  // namespace

// `d_descs` holds `n_out` descriptors; `max_pixels` = largest oh*ow among them.
int launch_transform(mtm_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const XformDesc* d_descs, int n_out,
                     int64_t max_pixels, int C, int dtype, int factor)
{
    if (n_out <= 0 || max_pixels <= 0) return MTM_OK;
    const float scale = 1.f / (float)(factor * factor);
    const int64_t want = (max_pixels + 255) / 256;
    dim3 grid((unsigned)std::min<int64_t>(want, (int64_t)ctx->sm_count * 32), (unsigned)n_out);
    if (dtype == MTM_U8) transform_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(d_src, d_dst, d_descs, C, factor, scale);
    else if (dtype == MTM_U16) transform_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(d_src, d_dst, d_descs, C, factor, scale);
    else transform_kernel<float><<<grid, 256, 0, ctx->stream>>>(d_src, d_dst, d_descs, C, factor, scale);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}
