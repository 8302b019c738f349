// mtm_api.cu -- the C ABI of libmtm_b200.so (include/mtm_b200.h): context, uploads,
// and the host-side sequencing of the kernels.  No arithmetic happens here.
#include "mtm_internal.cuh"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <utility>

static thread_local std::string g_create_err;

int mtm_fail(mtm_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

template <typename T>
int mtm_reserve(mtm_ctx* ctx, T*& ptr, size_t& cap, size_t need)
{
    if (need <= cap && ptr) return MTM_OK;
    if (ptr) { MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); MTM_CUDA(ctx, cudaFree(ptr)); ptr = nullptr; cap = 0; }
    const size_t want = need + need / 8 + 64;
    MTM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ptr), want * sizeof(T)));
    MTM_CUDA(ctx, cudaMemsetAsync(ptr, 0, want * sizeof(T), ctx->stream));
    cap = want;
    return MTM_OK;
}

template <typename T>
static int reserve_pinned(mtm_ctx* ctx, T*& ptr, size_t& cap, size_t need)
{
    if (need <= cap && ptr) return MTM_OK;
    if (ptr) { MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); MTM_CUDA(ctx, cudaFreeHost(ptr)); ptr = nullptr; cap = 0; }
    const size_t want = need + need / 8 + 64;
    MTM_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&ptr), want * sizeof(T)));
    cap = want;
    return MTM_OK;
}


// Folds completed MTM_OPT_TIME_NCC event brackets into the counters.  `all`: wait for every pending
// bracket; otherwise only wait when the ring is full.
static int harvest_ncc_time(mtm_ctx* ctx, bool all)
{
    while (ctx->ncc_tail != ctx->ncc_head) {
        const int k = ctx->ncc_tail % MTM_NCC_RING;
        const bool full = (ctx->ncc_head - ctx->ncc_tail) >= MTM_NCC_RING;
        if (all || full) MTM_CUDA(ctx, cudaEventSynchronize(ctx->ev_ncc[k][1]));
        float ms = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms, ctx->ev_ncc[k][0], ctx->ev_ncc[k][1]);
        if (e == cudaErrorNotReady) { (void)cudaGetLastError(); return MTM_OK; }
        MTM_CUDA(ctx, e);
        ctx->ctr.ncc_ms += ms;
        ctx->ctr.ncc_launches += ctx->ncc_launches_of[k];
        ctx->ncc_tail++;
    }
    return MTM_OK;
}

// MTM_OPT_TIME_NCC: an event bracket around a run of numerator launches (closed and reopened around anything else that is
// queued in between, e.g. the window moments of the next band).
static int ncc_bracket_open(mtm_ctx* ctx)
{
    if (!ctx->time_ncc) return MTM_OK;
    MTM_TRY(harvest_ncc_time(ctx, false));
    MTM_CUDA(ctx, cudaEventRecord(ctx->ev_ncc[ctx->ncc_head % MTM_NCC_RING][0], ctx->stream));
    ctx->ncc_mark = ctx->ctr.kernel_launches;
    return MTM_OK;
}
static int ncc_bracket_close(mtm_ctx* ctx)
{
    if (!ctx->time_ncc) return MTM_OK;
    const int k = ctx->ncc_head % MTM_NCC_RING;
    MTM_CUDA(ctx, cudaEventRecord(ctx->ev_ncc[k][1], ctx->stream));
    ctx->ncc_launches_of[k] = (int)(ctx->ctr.kernel_launches - ctx->ncc_mark);
    ctx->ncc_head++;
    return MTM_OK;
}

// Debug aid (MTM_B200_STAGES=1): CUDA events between the stages of a call, printed at its end.
struct StageMarks {
    std::vector<std::pair<const char*, cudaEvent_t>> ev;
    bool on = getenv("MTM_B200_STAGES") != nullptr;
    void mark(mtm_ctx* ctx, const char* name) {
        if (!on) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, ctx->stream); ev.emplace_back(name, e);
    }
    void report(mtm_ctx* ctx) {
        if (!on || ev.size() < 2) return;
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[mtm stages]");
        for (size_t i = 1; i < ev.size(); ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second); fprintf(stderr, " %s=%.1fus", ev[i].first, ms * 1e3f); }
        float tot = 0; cudaEventElapsedTime(&tot, ev.front().second, ev.back().second);
        fprintf(stderr, " total=%.1fus\n", tot * 1e3f);
        for (auto& p : ev) cudaEventDestroy(p.second);
        ev.clear();
    }
};
static StageMarks g_marks;

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int reserve_hits(mtm_ctx* ctx, int cap)
{
    cap = next_pow2(std::max(cap, 1024));
    if (cap <= ctx->hit_cap) return MTM_OK;
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_blockA) cudaFree(ctx->d_blockA);
    if (ctx->d_blockB) cudaFree(ctx->d_blockB);
    if (ctx->d_keep) cudaFree(ctx->d_keep);
    ctx->d_blockA = ctx->d_blockB = nullptr; ctx->d_keep = nullptr; ctx->hit_cap = 0;
    const size_t bytes = MTM_HIT_HEADER + (size_t)cap * sizeof(DevHit);
    MTM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_blockA), bytes));
    MTM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_blockB), bytes));
    MTM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_keep), (size_t)cap * sizeof(int32_t)));
    MTM_CUDA(ctx, cudaMemsetAsync(ctx->d_blockA, 0, bytes, ctx->stream));
    MTM_CUDA(ctx, cudaMemsetAsync(ctx->d_blockB, 0, bytes, ctx->stream));
    ctx->hit_cap = cap;
    MTM_TRY(reserve_pinned(ctx, ctx->h_stage, ctx->h_stage_cap, bytes));
    if (!ctx->h_mirror) {
        const size_t mb = MTM_HIT_HEADER + (size_t)MTM_MIRROR_HITS * sizeof(DevHit);
        MTM_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_mirror), mb, cudaHostAllocMapped));
        memset(ctx->h_mirror, 0, mb);
        MTM_CUDA(ctx, cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->d_mirror), ctx->h_mirror, 0));
    }
    return MTM_OK;
}

extern "C" {

int mtm_abi_version(void) { return MTM_ABI_VERSION; }

int mtm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

const char* mtm_last_error(const mtm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int mtm_create(int device, mtm_ctx** out)
{
    if (!out) return mtm_fail(nullptr, MTM_ERR_INVALID, "mtm_create: null output pointer");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return mtm_fail(nullptr, MTM_ERR_CUDA, "mtm_create: no CUDA device (%s); libmtm_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n_dev)
        return mtm_fail(nullptr, MTM_ERR_INVALID, "mtm_create: device %d out of range [0, %d)", device, n_dev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return mtm_fail(nullptr, MTM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return mtm_fail(nullptr, MTM_ERR_CUDA, "mtm_create: device %d is sm_%d%d; libmtm_b200 is built for sm_100a only",
                        device, prop.major, prop.minor);
    mtm_ctx* ctx = new mtm_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    auto bail = [&](const char* what, cudaError_t err) {
        int rc = mtm_fail(nullptr, MTM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
        delete ctx;
        return rc;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    ctx->stream = ctx->own_stream;
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int k = 0; k < MTM_NCC_RING; ++k)
        for (int j = 0; j < 2; ++j)
            if ((e = cudaEventCreate(&ctx->ev_ncc[k][j])) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_cand), (size_t)MTM_CAND_CAP * sizeof(DevHit))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_cand_count), 64)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_hkeys), (size_t)MTM_HASH_SLOTS * sizeof(unsigned long long))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_hvals), (size_t)MTM_HASH_SLOTS * sizeof(int32_t))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMemsetAsync(ctx->d_hkeys, 0xFF, (size_t)MTM_HASH_SLOTS * sizeof(unsigned long long), ctx->stream)) != cudaSuccess) return bail("cudaMemsetAsync", e);
    int rc = reserve_hits(ctx, 1 << 16);
    if (rc != MTM_OK) { g_create_err = ctx->err; mtm_destroy(ctx); return rc; }
    *out = ctx;
    return MTM_OK;
}

int mtm_destroy(mtm_ctx* ctx)
{
    if (!ctx) return MTM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->img.pix); cudaFree(ctx->img.sat_s); cudaFree(ctx->img.sat_q); cudaFree(ctx->img.sat_q32); cudaFree(ctx->scratch);
    cudaFree(ctx->d_meta); cudaFree(ctx->d_tmpl); cudaFree(ctx->d_maps); cudaFree(ctx->d_order);
    cudaFree(ctx->d_blockA); cudaFree(ctx->d_blockB); cudaFree(ctx->d_keep);
    cudaFree(ctx->d_nontrivial); cudaFree(ctx->d_best); cudaFree(ctx->d_cand); cudaFree(ctx->d_cand_count); cudaFree(ctx->d_hkeys); cudaFree(ctx->d_hvals);
    cudaFree(ctx->img.pixf2); cudaFree(ctx->d_raw_t); cudaFree(ctx->d_raw_m); cudaFree(ctx->d_maps2);
    cudaFree(ctx->img.pixf); cudaFree(ctx->img.satf_s); cudaFree(ctx->img.satf_q); cudaFree(ctx->d_tmpl_centred);
    cudaFree(ctx->d_raw16); cudaFree(ctx->img.pix_lo); cudaFree(ctx->d_tmpl8); cudaFree(ctx->d_pix8); cudaFree(ctx->d_acc);
    cudaFree(ctx->d_full); cudaFree(ctx->d_small); cudaFree(ctx->d_xform);
    cudaFree(ctx->d_slabs); cudaFree(ctx->d_wS); cudaFree(ctx->d_wR); cudaFree(ctx->d_sizes);
    for (int k = 0; k < MTM_MAX_INFLIGHT; ++k) { cudaFree(ctx->d_slot[k]); cudaFreeHost(ctx->h_slot[k]); if (ctx->ev_slot[k]) cudaEventDestroy(ctx->ev_slot[k]); }
    cudaFreeHost(ctx->h_tmpl_stage); cudaFreeHost(ctx->h_stage); cudaFreeHost(ctx->h_geom); cudaFreeHost(ctx->h_mirror);
    for (int k = 0; k < MTM_STAGE_BUFS; ++k) { cudaFreeHost(ctx->h_chunk[k]); if (ctx->ev_chunk[k]) cudaEventDestroy(ctx->ev_chunk[k]); }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int k = 0; k < MTM_NCC_RING; ++k)
        for (int j = 0; j < 2; ++j)
            if (ctx->ev_ncc[k][j]) cudaEventDestroy(ctx->ev_ncc[k][j]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return MTM_OK;
}

int mtm_set_stream(mtm_ctx* ctx, void* cuda_stream)
{
    MTM_ENTER(ctx);
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return MTM_OK;
}

int mtm_synchronize(mtm_ctx* ctx)
{
    MTM_ENTER(ctx);
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MTM_OK;
}

int mtm_set_option(mtm_ctx* ctx, int option, int64_t value)
{
    if (!ctx) return MTM_ERR_INVALID;
    if (option == MTM_OPT_PATH && value >= MTM_PATH_AUTO && value <= MTM_PATH_TENSOR) { ctx->path = (int)value; return MTM_OK; }
    if (option == MTM_OPT_TIME_NCC) { ctx->time_ncc = value ? 1 : 0; return MTM_OK; }
    return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_option: unknown option %d / value %lld", option, (long long)value);
}

int mtm_get_counters(mtm_ctx* ctx, mtm_counters* out)
{
    if (!ctx || !out) return MTM_ERR_INVALID;
    MTM_CUDA(ctx, cudaSetDevice(ctx->device));
    MTM_TRY(harvest_ncc_time(ctx, true));
    *out = ctx->ctr;
    return MTM_OK;
}

int mtm_reset_counters(mtm_ctx* ctx)
{
    if (!ctx) return MTM_ERR_INVALID;
    ctx->ctr = mtm_counters{};
    return MTM_OK;
}

int mtm_measure_i8_peak(mtm_ctx* ctx, int n_cols, int iters, double* tmacs_per_s)
{
    MTM_ENTER(ctx);
    if (!tmacs_per_s || n_cols < 16 || n_cols > 256 || n_cols % 16 || iters < 1)
        return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_measure_i8_peak: N in 16..256 (multiple of 16), iters >= 1");
    MTM_TRY(launch_i8_peak(ctx, n_cols, 64));                  // warm-up: module load, clocks
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        MTM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        MTM_TRY(launch_i8_peak(ctx, n_cols, iters));
        MTM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        MTM_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        MTM_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        best = std::min(best, ms);
    }
    *tmacs_per_s = (double)ctx->sm_count * iters * 128.0 * n_cols * 32.0 / (best * 1e-3) / 1e12;
    return MTM_OK;
}

int mtm_timer_begin(mtm_ctx* ctx)
{
    MTM_ENTER(ctx);
    MTM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return MTM_OK;
}

int mtm_timer_end(mtm_ctx* ctx, float* elapsed_ms)
{
    MTM_ENTER(ctx);
    if (!elapsed_ms) return MTM_ERR_INVALID;
    MTM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    MTM_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    MTM_CUDA(ctx, cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
    return MTM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------
// Summed-area tables of the resident uint8 image (window_stats.cu).
static int ensure_sat(mtm_ctx* ctx)
{
    if (ctx->sat_valid) return MTM_OK;
    MTM_TRY(launch_build_sat(ctx));
    ctx->sat_valid = true;
    return MTM_OK;
}

static int set_image_impl(mtm_ctx* ctx, const void* pixels, int H, int W, int C, int dtype,
                          int64_t row_stride, bool on_device)
{
    MTM_ENTER(ctx);
    if (!pixels || H <= 0 || W <= 0) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image: empty image (%d x %d)", H, W);
    if (C < 1 || C > MTM_MAX_CH) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image: %d channels (1..4 supported)", C);
    if (dtype != MTM_U8 && dtype != MTM_F32 && dtype != MTM_U16) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image: unknown dtype %d", dtype);
    if ((int64_t)W * C > 66000) return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "mtm_set_image: rows wider than 66000 elements");
    if (dtype == MTM_U16 && C != 1) return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "mtm_set_image: 16-bit images must be single channel (cast to float32 otherwise)");
    const int64_t esz = dtype == MTM_F32 ? 4 : (dtype == MTM_U16 ? 2 : 1);
    if (row_stride < (int64_t)W * C * esz) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image: row stride %lld < %lld", (long long)row_stride, (long long)W * C * esz);
    ImageDev& im = ctx->img;
    if (dtype == MTM_F32 || dtype == MTM_U16) {
        // MTM_U16 = the reference's uint16 -> float32 cast done on the device, plus the byte planes of the exact numerator
        const bool u16 = dtype == MTM_U16;
        if (C == 2) return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "mtm_set_image: 2-channel float32 images");
        im.pitch_e = ((int64_t)W * C + 3) / 4 * 4;
        MTM_TRY(mtm_reserve(ctx, im.pixf, ctx->imgf_cap, (size_t)(H * im.pitch_e + 64)));
        // float32 pixels that are all integers in [0, 65535] (uint16 data cast on the host, mixed uint8 / uint16 inputs after the
        // reference's float32 cast) take the same exact byte-plane numerator as MTM_U16: detected on the device (MTM_B200_F32_EXACT=0: off)
        static const bool f32_exact = !(getenv("MTM_B200_F32_EXACT") && atoi(getenv("MTM_B200_F32_EXACT")) == 0);
        const bool probe = !u16 && C == 1 && f32_exact;
        const bool same_shape_f = (im.H == H && im.W == W && im.C == C && ctx->img_dtype == MTM_F32 && ctx->img_src_dtype == dtype);
        ctx->img_src_dtype = dtype;
        im.H = H; im.W = W; im.C = C;
        im.sat_pitch = ((int64_t)W + 1 + 3) / 4 * 4;
        g_marks.mark(ctx, "begin");
        if (u16 || probe) {
            const int64_t pitch = (((int64_t)W + 64 + 64) + 127) / 128 * 128;      // u8 tile layout of the tensor-core kernel
            const size_t plane_bytes = (size_t)(H * pitch + 256);
            const bool fresh = ctx->img_cap < plane_bytes || ctx->pix_lo_cap < plane_bytes || im.pitch != pitch || !same_shape_f;
            MTM_TRY(mtm_reserve(ctx, im.pix, ctx->img_cap, plane_bytes));
            MTM_TRY(mtm_reserve(ctx, im.pix_lo, ctx->pix_lo_cap, plane_bytes));
            if (fresh) {                                           // the padding beyond W must read as zero
                MTM_CUDA(ctx, cudaMemsetAsync(im.pix, 0, ctx->img_cap, ctx->stream));
                MTM_CUDA(ctx, cudaMemsetAsync(im.pix_lo, 0, ctx->pix_lo_cap, ctx->stream));
            }
            im.pitch = pitch;
        }
        bool planes = u16;
        if (u16) {
            const uint16_t* src16 = static_cast<const uint16_t*>(pixels);
            int64_t src_stride = row_stride;
            if (!on_device) {
                MTM_TRY(mtm_reserve(ctx, ctx->d_raw16, ctx->raw16_cap, (size_t)H * W + 8));
                MTM_TRY(mtm_upload_rows(ctx, ctx->d_raw16, (size_t)W * 2, pixels, (size_t)row_stride, (size_t)W * 2, H));
                ctx->ctr.h2d_bytes += (int64_t)H * W * 2;
                src16 = ctx->d_raw16; src_stride = (int64_t)W * 2;
            }
            MTM_TRY(launch_u16_split_image(ctx, src16, src_stride));
        } else {
            if (on_device)
                MTM_CUDA(ctx, cudaMemcpy2DAsync(im.pixf, (size_t)im.pitch_e * 4, pixels, (size_t)row_stride, (size_t)W * C * 4, (size_t)H,
                                                cudaMemcpyDeviceToDevice, ctx->stream));
            else
                MTM_TRY(mtm_upload_rows(ctx, im.pixf, (size_t)im.pitch_e * 4, pixels, (size_t)row_stride, (size_t)W * C * 4, H));
            if (!on_device) ctx->ctr.h2d_bytes += (int64_t)H * W * C * 4;
            if (probe) {
                int not_integral = 1;
                MTM_TRY(launch_f32_split_image(ctx, &not_integral));           // byte planes + "some pixel is no uint16" (one 4-byte read-back)
                planes = !not_integral;
            }
        }
        ctx->img_u16 = planes;
        dtype = MTM_F32;
        MTM_TRY(mtm_reserve(ctx, im.satf_s, ctx->satf_s_cap, (size_t)C * (H + 1) * im.sat_pitch));
        MTM_TRY(mtm_reserve(ctx, im.satf_q, ctx->satf_q_cap, (size_t)(H + 1) * im.sat_pitch));
        MTM_TRY(mtm_reserve(ctx, ctx->scratch, ctx->scratch_cap, (size_t)2 * (C + 1) * H * W + 16));
        ctx->img_dtype = dtype;
        if (!same_shape_f) ctx->geometry_valid = false;     // map offsets depend on the image shape only
        ctx->moments_valid = false;
        ctx->masked_image_valid = false;
        MTM_TRY(launch_build_sat_f32(ctx));
        return MTM_OK;
    }
    const int64_t pitch = (((int64_t)W * C + 64 * C + 64) + 127) / 128 * 128;
    const bool reshape = (im.H != H || im.W != W || im.C != C);
    const bool same_shape = !reshape && ctx->img_dtype == dtype && ctx->img_src_dtype == MTM_U8;
    ctx->img_u16 = false;
    ctx->img_src_dtype = MTM_U8;
    MTM_TRY(mtm_reserve(ctx, im.pix, ctx->img_cap, (size_t)(H * pitch + 256)));
    if (reshape || !same_shape || im.pitch != pitch) MTM_CUDA(ctx, cudaMemsetAsync(im.pix, 0, ctx->img_cap, ctx->stream));
    im.pitch = pitch; im.H = H; im.W = W; im.C = C;
    im.sat_pitch = ((int64_t)W + 1 + 3) / 4 * 4;
    g_marks.mark(ctx, "begin");
    if (on_device)
        MTM_CUDA(ctx, cudaMemcpy2DAsync(im.pix, (size_t)pitch, pixels, (size_t)row_stride, (size_t)W * C, (size_t)H,
                                        cudaMemcpyDeviceToDevice, ctx->stream));
    else
        MTM_TRY(mtm_upload_rows(ctx, im.pix, (size_t)pitch, pixels, (size_t)row_stride, (size_t)W * C, H));      // pageable sources: staged through pinned chunks
    g_marks.mark(ctx, "copy_image");
    if (!on_device) ctx->ctr.h2d_bytes += (int64_t)H * W * C;
    MTM_TRY(mtm_reserve(ctx, im.sat_s, ctx->sat_s_cap, (size_t)C * (H + 1) * im.sat_pitch));
    MTM_TRY(mtm_reserve(ctx, im.sat_q, ctx->sat_q_cap, (size_t)(H + 1) * im.sat_pitch));
    MTM_TRY(mtm_reserve(ctx, im.sat_q32, ctx->sat_q32_cap, (size_t)(H + 1) * im.sat_pitch));
    MTM_TRY(mtm_reserve(ctx, ctx->scratch, ctx->scratch_cap, (size_t)(C + 1) * H * ((W + 3) / 4 * 4)));
    ctx->img_dtype = dtype;
    if (!same_shape) ctx->geometry_valid = false;           // map offsets depend on the image shape only: a stream of
                                                            // equal-sized images keeps them (and skips the sync in ensure_geometry)
    ctx->moments_valid = false;
    ctx->masked_image_valid = false;
    ctx->sat_valid = false;
    if (!box_moments_enabled()) MTM_TRY(ensure_sat(ctx));   // knob: compute_maps builds the tables when a kernel of the call reads them
    g_marks.mark(ctx, "sat");
    return MTM_OK;
}

int ensure_geometry(mtm_ctx* ctx)
{
    if (ctx->img.H == 0) return mtm_fail(ctx, MTM_ERR_INVALID, "no image set (call mtm_set_image first)");
    if (ctx->n_tmpl == 0) return mtm_fail(ctx, MTM_ERR_INVALID, "no templates set (call mtm_set_templates first)");
    if (ctx->geometry_valid) return MTM_OK;
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // h_geom (pinned) may still feed an earlier async copy
    if (ctx->tmpl_C != ctx->img.C || (!ctx->masked && ctx->tmpl_dtype != ctx->img_dtype))
        return mtm_fail(ctx, MTM_ERR_INVALID, "image and templates differ in channel count or dtype");
    const int n = ctx->n_tmpl;
    int64_t off = 0;
    for (int t = 0; t < n; ++t) {
        TmplMeta& m = ctx->h_meta[t];
        if (m.h > ctx->img.H || m.w > ctx->img.W)
            return mtm_fail(ctx, MTM_ERR_INVALID, "template %d (%d x %d) is larger than the image (%d x %d)", t, m.h, m.w,
                            ctx->img.H, ctx->img.W);
        m.mh = ctx->img.H - m.h + 1;
        m.mw = ctx->img.W - m.w + 1;
        m.map_off = off;
        off += ((int64_t)m.mh * m.mw + 31) / 32 * 32;
    }
    ctx->maps_total = off;
    ctx->moments_valid = false;
    // Window moments: the distinct sizes in (h, w)-sorted order.  They are produced per template group and per band of
    // output rows into a ring that every group and band reuses (mtm_internal.cuh, SizeDesc): offsets are per group.
    ctx->h_sizes.clear();
    std::vector<int> size_of((size_t)n, 0);
    for (int k = 0; k < n; ++k) {
        const TmplMeta& m = ctx->h_meta[ctx->h_order[k]];
        if (ctx->h_sizes.empty() || ctx->h_sizes.back().h != m.h || ctx->h_sizes.back().w != m.w)
            ctx->h_sizes.push_back(SizeDesc{m.h, m.w, m.mh, m.mw, 0, m.mh, 0});
        size_of[k] = (int)ctx->h_sizes.size() - 1;
    }
    ctx->moments_total = 0;
    ctx->box_ok = false;
    if (ctx->tc_ready && ctx->img_dtype == MTM_U8) {
        // box sums straight from the image (banded) unless some input rules them out: then the summed-area route, whole maps
        bool box = box_moments_enabled() && box_moments_applicable(ctx);
        for (const TcGroup& g : ctx->tc_groups) box = box && !points_path_preferred(ctx, g.first, g.count);
        ctx->box_ok = box;
        // Two placements of the moments (profiles/README.md, round 2):
        //  * resident (default): every size's whole map, produced by ONE launch per image before the groups' numerator launches;
        //  * ring (MTM_B200_RING_KB=<budget>, read at every geometry change): per template group and per band of rows into a
        //    buffer of that size which the next band overwrites, so that the moments never leave L2.  Measured on
        //    BASELINE configs[4]: the DRAM traffic of the moment stage disappears, but every band costs the fill and drain of one
        //    more persistent numerator launch (~25 us each), which outweighs it (sync 6.7 -> 8.7 ms at 96 MB, 10.2 ms at 48 MB).
        const char* ring_env = getenv("MTM_B200_RING_KB");
        ctx->moments_ring = ring_env != nullptr && box;
        for (size_t gi = 1; gi < ctx->tc_groups.size() && ctx->moments_ring; ++gi)       // a size shared by two groups: its segment
            if (size_of[ctx->tc_groups[gi].first] == size_of[ctx->tc_groups[gi].first - 1]) ctx->moments_ring = false;   // cannot sit in both rings
        const int64_t ring_bytes = ring_env ? (int64_t)std::max(1, atoi(ring_env)) << 10 : 0;
        const int64_t entry_bytes = ctx->img.C == 1 ? 8 : 4 * (ctx->img.C + 1);
        int mh_all = 1;
        for (const SizeDesc& sd : ctx->h_sizes) mh_all = std::max(mh_all, sd.mh);
        if (!ctx->moments_ring) {                               // one band; the segments of every size side by side
            int64_t moff_all = 0;
            for (SizeDesc& sd : ctx->h_sizes) {
                sd.off = moff_all; sd.band = mh_all;
                moff_all += mom_segment(sd.mw, mh_all);
            }
            ctx->moments_total = moff_all;
        }
        for (TcGroup& g : ctx->tc_groups) {                     // (two groups of equal-sized templates share their sizes)
            g.size_first = size_of[g.first];
            g.size_count = size_of[g.first + g.count - 1] - g.size_first + 1;
            g.band_rows = mh_all;
        }
        for (TcGroup& g : ctx->tc_groups) {
            if (!ctx->moments_ring) break;
            g.size_first = size_of[g.first];
            g.size_count = size_of[g.first + g.count - 1] - g.size_first + 1;
            int mh_max = 1;
            int64_t per_row = 0;
            for (int q = 0; q < g.size_count; ++q) {
                const SizeDesc& sd = ctx->h_sizes[g.size_first + q];
                mh_max = std::max(mh_max, sd.mh);
                per_row += (int64_t)((sd.mw + 15) >> 4) * 16;
            }
            int band = mh_max;
            const int64_t rows_budget = std::max<int64_t>(16, ring_bytes / entry_bytes / per_row);
            if (box && rows_budget < mh_max) {                  // several bands of (nearly) equal height, multiples of 16 rows
                const int bands = (int)((mh_max + rows_budget - 1) / rows_budget);
                band = (((mh_max + bands - 1) / bands) + 15) & ~15;
            }
            g.band_rows = band;
            int64_t moff = 0;
            for (int q = 0; q < g.size_count; ++q) {
                SizeDesc& sd = ctx->h_sizes[g.size_first + q];
                sd.off = moff; sd.band = band;
                moff += mom_segment(sd.mw, band);
            }
            ctx->moments_total = std::max(ctx->moments_total, moff);
        }
        MTM_TRY(mtm_reserve(ctx, ctx->d_wS, ctx->wS_cap, (size_t)ctx->moments_total * std::max(2, ctx->img.C)));   // C == 1: interleaved {S, rsD}
        MTM_TRY(mtm_reserve(ctx, ctx->d_wR, ctx->wR_cap, (size_t)ctx->moments_total));
        MTM_TRY(mtm_reserve(ctx, ctx->d_sizes, ctx->sizes_cap, ctx->h_sizes.size()));
        MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_sizes, ctx->h_sizes.data(), ctx->h_sizes.size() * sizeof(SizeDesc),
                                      cudaMemcpyHostToDevice, ctx->stream));      // pageable source: staged before returning
    }
    for (int k = 0; k < n; ++k) ctx->h_meta[ctx->h_order[k]].mom_off = ctx->h_sizes[size_of[k]].off;
    for (int t = 0; t < n; ++t) {
        const TmplMeta& m = ctx->h_meta[t];
        ctx->h_geom[t].map_off = m.map_off; ctx->h_geom[t].mh = m.mh; ctx->h_geom[t].mw = m.mw; ctx->h_geom[t].mom_off = m.mom_off;
    }
    MTM_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_meta, sizeof(TmplMeta), ctx->h_geom, sizeof(TmplGeom), sizeof(TmplGeom),
                                    (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->maps_resident = false;                             // the map arena is reserved by compute_maps when a search needs it
    ctx->geometry_valid = true;
    return MTM_OK;
}

// Groups the (h, w)-sorted templates into tcgen05 launches and expands their Toeplitz slabs
// (template-only work: done once per mtm_set_templates).  Leaves tc_ready false when some
// template cannot take the tensor path (multi-channel, too large for shared memory, ...).
static int plan_tensor_path(mtm_ctx* ctx)
{
    ctx->tc_groups.clear();
    ctx->tc_ready = false;
    ctx->moments_valid = false;
    const bool planes16 = ctx->tmpl_u16 && ctx->tmpl_C == 1;      // 16-bit templates: slabs of the high- and the low-byte plane
    if ((ctx->tmpl_C != 1 && ctx->tmpl_C != 3 && ctx->tmpl_C != 4) || (ctx->tmpl_dtype != MTM_U8 && !planes16)) return MTM_OK;
    const int n = ctx->n_tmpl;
    const int TC = ctx->tmpl_C;
    // Grouping = the cheapest partition of the (h, w)-sorted templates into launches (dynamic programme over prefixes; a
    // mode-A launch takes up to 8 CONSECUTIVE templates, zero padded to the largest member, a mode-B launch one template).
    // Cost of a launch in clocks per output row of a 16-column tile (what the measured role clocks of the persistent kernel
    // suggest, profiles/README.md): MMAs 0.6 * h * nk (~1.1 (N/2 + 8) clocks each at N ~ 200, h * nk of them per tile), the
    // epilogue ~55 -- whichever is longer plus a share of the other (they overlap imperfectly); a mode-B launch covers eight
    // such columns at once for its single template.  The epilogue term is what the first planner missed: a launch of four
    // templates costs as much as one of eight (C5, 64 sizes: 11 launches, three of them with 4-5 templates -> 8 full ones).
    // Every launch also pays ~20 us (launch, first tile, last epilogue: ~40 K clocks), expressed in the same unit through the
    // score-map area of the resident image (2 MP when none is set yet): on small images fewer launches win, on large ones
    // the term only breaks ties against single-template launches.
    const double area = ctx->img.H > 0 ? (double)ctx->img.H * ctx->img.W : 2.0e6;
    const double per_launch = 40000.0 * ctx->sm_count * 16.0 / std::max(area, 1.0);
    auto launch_cost = [per_launch](const TcGroup& g) {
        const double mma = 0.6 * g.h * g.nk, epi = 55.0;
        const double c = std::max(mma, epi) + 0.3 * std::min(mma, epi);
        return (g.mode == 0 ? c : c / 8.0) + per_launch;
    };
    int64_t arena = 0;
    auto emit = [&](TcGroup g, int first, int count, int h_min, int w_min) {
        g.first = first; g.count = count; g.h_min = h_min; g.w_min = w_min;
        g.arena_off = arena;
        arena += ((int64_t)g.h * g.slab_bytes + 127) / 128 * 128;
        ctx->tc_groups.push_back(g);
    };
    std::vector<double> best((size_t)n + 1, 1e300);
    std::vector<int> take((size_t)n + 1, 0), mode_of((size_t)n + 1, 0);
    best[0] = 0.0;
    for (int i = 1; i <= n; ++i) {
        int hg = 0, wg = 0;
        for (int k = 1; k <= 8 && k <= i; ++k) {                     // a mode-A launch of the templates [i - k, i)
            const TmplMeta& m = ctx->h_meta[ctx->h_order[i - k]];
            hg = std::max(hg, (int)m.h); wg = std::max(wg, (int)m.w);
            TcGroup g{};
            if (!tc_plan_group(0, hg, wg, TC, g)) break;
            const double c = best[i - k] + launch_cost(g);
            if (c < best[i] - 1e-12) { best[i] = c; take[i] = k; mode_of[i] = 0; }
        }
        const TmplMeta& m1 = ctx->h_meta[ctx->h_order[i - 1]];
        TcGroup b{};
        if (tc_plan_group(1, m1.h, m1.w, TC, b)) {                   // ... or template i - 1 alone in mode B
            const double c = best[i - 1] + launch_cost(b);
            if (c < best[i] - 1e-12) { best[i] = c; take[i] = 1; mode_of[i] = 1; }
        }
        if (best[i] >= 1e299) return MTM_OK;                         // a template no tile fits: no tensor path for this set
    }
    std::vector<std::pair<int, int>> cuts;                           // (first, count) back to front
    for (int i = n; i > 0; i -= take[i]) cuts.push_back({i - take[i], take[i] | (mode_of[i] << 8)});
    for (auto it = cuts.rbegin(); it != cuts.rend(); ++it) {
        const int first = it->first, count = it->second & 255, mode = it->second >> 8;
        int hg = 0, wg = 0, h_min = 1 << 30, w_min = 1 << 30;
        for (int k = first; k < first + count; ++k) {
            const TmplMeta& m = ctx->h_meta[ctx->h_order[k]];
            hg = std::max(hg, (int)m.h); wg = std::max(wg, (int)m.w);
            h_min = std::min(h_min, (int)m.h); w_min = std::min(w_min, (int)m.w);
        }
        TcGroup g{};
        if (!tc_plan_group(mode, hg, wg, TC, g)) return MTM_OK;
        emit(g, first, count, h_min, w_min);
    }
    ctx->slab_plane = (arena + 127) / 128 * 128;
    MTM_TRY(mtm_reserve(ctx, ctx->d_slabs, ctx->slabs_cap, (size_t)(planes16 ? 2 : 1) * ctx->slab_plane + 128));
    for (const TcGroup& g : ctx->tc_groups) MTM_TRY(launch_toeplitz_prep(ctx, g));
    ctx->tc_ready = true;
    return MTM_OK;
}

static bool use_tensor_path(const mtm_ctx* ctx, int method)
{
    if (ctx->path == MTM_PATH_DIRECT || !ctx->tc_ready) return false;
    if (ctx->img_dtype == MTM_F32 && !(ctx->img_u16 && ctx->tmpl_u16 && ctx->img.C == 1)) return false;   // float32 data: byte planes needed
    for (const TcGroup& g : ctx->tc_groups)
        if (!tc_path_supported(ctx, method, g.h, g.w)) return false;
    return true;
}

// Masked templates (methods 0 / 3): two plain fp32 correlations per template + the combine kernel.
static int compute_maps_masked(mtm_ctx* ctx, int method)
{
    if (method != MTM_TM_SQDIFF && method != MTM_TM_CCORR_NORMED)
        return mtm_fail(ctx, MTM_ERR_INVALID, "masks are only defined for TM_SQDIFF (0) and TM_CCORR_NORMED (3)");
    ImageDev& im = ctx->img;
    if (!ctx->masked_image_valid) {
        if (ctx->img_dtype == MTM_U8) {
            im.pitch_e = ((int64_t)im.W * im.C + 3) / 4 * 4;
            MTM_TRY(mtm_reserve(ctx, im.pixf, ctx->imgf_cap, (size_t)(im.H * im.pitch_e + 64)));
        }
        MTM_TRY(mtm_reserve(ctx, im.pixf2, ctx->pixf2_cap, (size_t)(im.H * im.pitch_e + 64)));
        MTM_TRY(launch_masked_image(ctx));
        ctx->masked_image_valid = true;
    }
    MTM_TRY(mtm_reserve(ctx, ctx->d_maps2, ctx->maps2_cap, (size_t)ctx->maps_total));
    const int n = ctx->n_tmpl;
    int i = 0;
    while (i < n) {
        const TmplMeta& a = ctx->h_meta[ctx->h_order[i]];
        int j = i + 1;
        while (j < n && ctx->h_meta[ctx->h_order[j]].h == a.h && ctx->h_meta[ctx->h_order[j]].w == a.w) ++j;
        MTM_TRY(launch_ncc_direct_f32(ctx, MTM_TM_CCORR, i, j - i, im.pixf, ctx->d_tmpl, ctx->d_maps));
        MTM_TRY(launch_ncc_direct_f32(ctx, MTM_TM_CCORR, i, j - i, im.pixf2, ctx->d_tmpl_centred, ctx->d_maps2));
        i = j;
    }
    MTM_TRY(launch_masked_combine(ctx, method, ctx->d_maps2));
    return MTM_OK;
}

// May the numerator kernel's epilogue list the above-threshold pixels for the peak search?  Only for the
// default method, the multi-object search and maps larger than the list (so that a constant map -- the one
// case peak_local_max treats specially -- can never hide behind a short list).
void request_candidates(mtm_ctx* ctx, int method, int64_t n_object, double thr)
{
    ctx->cand_on = false;
    ctx->want_n1 = n_object == 1;
    static const bool no_cand = getenv("MTM_B200_NO_CAND") != nullptr;     // experiments: always stream the maps for peaks
    if (method != MTM_TM_CCOEFF_NORMED || n_object == 1 || no_cand) return;
    for (int t = 0; t < ctx->n_tmpl; ++t) {
        const TmplMeta& m = ctx->h_meta[t];
        if (m.mh == 1 || m.mw == 1 || (int64_t)m.mh * m.mw <= MTM_CAND_CAP) return;
    }
    ctx->cand_on = true;
    ctx->cand_thr = (float)thr;
}

// Score maps of every template (tmpl < 0) or of one template, grouped by template size.
int compute_maps(mtm_ctx* ctx, int method, int tmpl, bool hits_ok)
{
    if (method < 0 || method > 5) return mtm_fail(ctx, MTM_ERR_INVALID, "unknown method %d", method);
    const int n = ctx->n_tmpl;
    int i = 0;
    ctx->hits_only = false; ctx->best_on = false; ctx->best_valid = false;
    if (ctx->masked) {
        ctx->cand_on = false; ctx->cand_valid = false;
        MTM_TRY(mtm_reserve(ctx, ctx->d_maps, ctx->maps_cap, (size_t)ctx->maps_total));
        ctx->maps_resident = true;
        return compute_maps_masked(ctx, method);
    }
    const bool tensor = use_tensor_path(ctx, method);
    if (!tensor && ctx->path == MTM_PATH_TENSOR)
        return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "tensor-core path requested but not available for these inputs/method");
    const bool tensor16 = tensor && ctx->img_dtype == MTM_F32;     // 16-bit byte-plane path
    ctx->cand_on = ctx->cand_on && tensor && !tensor16 && tmpl < 0;
    // Hits-only search: the caller reads no score map, every template runs the default-method tcgen05 kernel, and the peak
    // search has what it needs without one (candidate list, or the per-template arg-max for N_object == 1): the numerator
    // kernels then store no map at all (MODE 3) and the map arena is not even reserved.  MTM_B200_NO_HITS_ONLY=1: A/B runs.
    static const bool no_hits = getenv("MTM_B200_NO_HITS_ONLY") != nullptr;
    bool hits = hits_ok && !no_hits && tensor && !tensor16 && method == MTM_TM_CCOEFF_NORMED && tmpl < 0 && ctx->img_dtype == MTM_U8 &&
                (ctx->cand_on || ctx->want_n1);
    if (hits)
        for (const TcGroup& g : ctx->tc_groups) hits = hits && !points_path_preferred(ctx, g.first, g.count);
    if (!hits) MTM_TRY(mtm_reserve(ctx, ctx->d_maps, ctx->maps_cap, (size_t)ctx->maps_total));
    ctx->hits_only = hits;
    if (hits) ctx->ctr.hits_only_searches++;
    ctx->maps_resident = false;
    if (hits && ctx->want_n1 && !ctx->cand_on) {
        ctx->best_on = true;
        MTM_CUDA(ctx, cudaMemsetAsync(ctx->d_best, 0, (size_t)n * sizeof(unsigned long long), ctx->stream));
    }
    if (ctx->img_dtype == MTM_U8) {
        // Everything but the default method's tensor-core epilogue reads the summed-area tables; under MTM_B200_MOM_BOX the
        // window moments come from the image instead (box_moments.cu) and the tables are not built for such a call.
        // window moments: banded box sums (ctx->box_ok, ensure_geometry) or, when some input rules those out, the tables
        const bool box = ctx->box_ok && tensor && method == MTM_TM_CCOEFF_NORMED;
        if (!box) MTM_TRY(ensure_sat(ctx));
    }
    if (tensor16) { MTM_TRY(mtm_reserve(ctx, ctx->d_acc, ctx->acc_cap, (size_t)ctx->maps_total)); ctx->cand_on = false; }
    if (tensor && !tensor16 && method == MTM_TM_CCOEFF_NORMED && !ctx->moments_ring && ctx->img_dtype == MTM_U8) {
        // resident placement: the window moments of every distinct size in one launch
        bool any_tc = false;
        for (const TcGroup& g : ctx->tc_groups) any_tc = any_tc || !points_path_preferred(ctx, g.first, g.count);
        if (any_tc) {
            int mh_all = 1;
            for (const SizeDesc& sd : ctx->h_sizes) mh_all = std::max(mh_all, sd.mh);
            if (ctx->box_ok) MTM_TRY(launch_box_moments(ctx, 0, (int)ctx->h_sizes.size(), 0, mh_all));
            else MTM_TRY(launch_window_moments(ctx, 0, (int)ctx->h_sizes.size()));
        }
    }
    MTM_TRY(ncc_bracket_open(ctx));
    ctx->cand_valid = false;
    if (ctx->cand_on) MTM_CUDA(ctx, cudaMemsetAsync(ctx->d_cand_count, 0, sizeof(int32_t), ctx->stream));
    if (tensor) {
        for (const TcGroup& g : ctx->tc_groups) {
            if (tmpl >= 0) {
                bool has = false;
                for (int k = 0; k < g.count; ++k) has = has || ctx->h_order[g.first + k] == tmpl;
                if (!has) continue;
            }
            if (tensor16) {
                // CC = 65536 Ih*Th + 256 (Ih*Tl + Il*Th) + Il*Tl: four exact u8 x u8 correlations into the double map
                MTM_TRY(launch_ncc_tc_accum(ctx, g, 0, 0, 65536.0, true));
                MTM_TRY(launch_ncc_tc_accum(ctx, g, 0, 1, 256.0, false));
                MTM_TRY(launch_ncc_tc_accum(ctx, g, 1, 0, 256.0, false));
                MTM_TRY(launch_ncc_tc_accum(ctx, g, 1, 1, 1.0, false));
            } else if (points_path_preferred(ctx, g.first, g.count)) {
                MTM_TRY(launch_ncc_points(ctx, method, g.first, g.count));          // tiny maps of large templates
            } else if (method != MTM_TM_CCOEFF_NORMED) {
                MTM_TRY(launch_ncc_tc(ctx, g, method, 0, ctx->img.H - g.h_min + 1));       // float64 epilogue on the tables: no moments
            } else if (!ctx->moments_ring) {
                MTM_TRY(launch_ncc_tc(ctx, g, method, 0, ctx->img.H - g.h_min + 1));       // resident moments (produced above)
            } else {
                // the group's window moments band by band (ring in L2), each band followed by its numerator launch
                const int mh_max = ctx->img.H - g.h_min + 1;
                for (int y_base = 0; y_base < mh_max; y_base += g.band_rows) {
                    const int rows = std::min(g.band_rows, mh_max - y_base);
                    MTM_TRY(ncc_bracket_close(ctx));               // the moments are not numerator time
                    if (ctx->box_ok) MTM_TRY(launch_box_moments(ctx, g.size_first, g.size_count, y_base, rows));
                    else MTM_TRY(launch_window_moments(ctx, g.size_first, g.size_count));
                    MTM_TRY(ncc_bracket_open(ctx));
                    MTM_TRY(launch_ncc_tc(ctx, g, method, y_base, rows));
                }
            }
        }
        if (tensor16) MTM_TRY(launch_cc16_epilogue(ctx, method, tmpl));
        i = n;
        ctx->cand_valid = ctx->cand_on;
        ctx->best_valid = ctx->best_on;
    }
    ctx->cand_on = false;
    ctx->best_on = false;
    ctx->maps_resident = !hits;
    while (i < n) {
        const TmplMeta& a = ctx->h_meta[ctx->h_order[i]];
        int j = i + 1;
        while (j < n && ctx->h_meta[ctx->h_order[j]].h == a.h && ctx->h_meta[ctx->h_order[j]].w == a.w) ++j;
        const bool f32 = (ctx->img_dtype == MTM_F32);
        const bool points = !f32 && points_path_preferred(ctx, i, j - i);
        auto direct = [&](int first, int count) {
            return f32 ? launch_ncc_direct_f32(ctx, method, first, count)
                       : (points ? launch_ncc_points(ctx, method, first, count) : launch_ncc_direct(ctx, method, first, count));
        };
        if (tmpl < 0) {
            MTM_TRY(direct(i, j - i));
        } else {
            for (int k = i; k < j; ++k)
                if (ctx->h_order[k] == tmpl) MTM_TRY(direct(k, 1));
        }
        i = j;
    }
    MTM_TRY(ncc_bracket_close(ctx));
    return MTM_OK;
}

// Second half of every plain template upload: ctx->h_meta holds sizes / pitches / arena offsets of the n templates and
// the packed pixels are (queued to be) in ctx->d_tmpl.  Uploads the metadata and the (h, w) order, computes the
// template statistics and plans the tensor-core launches.
static int finish_templates(mtm_ctx* ctx, int n, int C, int dtype, size_t total)
{
    MTM_TRY(mtm_reserve(ctx, ctx->d_meta, ctx->meta_cap, (size_t)n));
    MTM_TRY(mtm_reserve(ctx, ctx->d_order, ctx->order_cap, (size_t)n));
    MTM_TRY(mtm_reserve(ctx, ctx->d_nontrivial, ctx->per_tmpl_cap, (size_t)n));
    MTM_TRY(mtm_reserve(ctx, ctx->d_best, ctx->best_cap, (size_t)n));
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_meta, ctx->h_meta.data(), (size_t)n * sizeof(TmplMeta), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h_order.resize(n);
    for (int t = 0; t < n; ++t) ctx->h_order[t] = t;
    std::stable_sort(ctx->h_order.begin(), ctx->h_order.end(), [&](int a, int b) {
        const TmplMeta &x = ctx->h_meta[a], &y = ctx->h_meta[b];
        return x.h != y.h ? x.h < y.h : x.w < y.w;
    });
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_order, ctx->h_order.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    // h_meta / h_order are pageable: the copies above are staged synchronously by the runtime.
    ctx->ctr.h2d_bytes += (int64_t)n * (sizeof(TmplMeta) + sizeof(int32_t));
    ctx->n_tmpl = n; ctx->tmpl_C = C; ctx->tmpl_dtype = dtype;
    ctx->geometry_valid = false;
    ctx->masked = false;
    if (dtype == MTM_F32) {
        MTM_TRY(mtm_reserve(ctx, ctx->d_tmpl_centred, ctx->tmplc_cap, total + 64));
        MTM_TRY(launch_tmpl_stats_f32(ctx));
    } else {
        MTM_TRY(launch_tmpl_stats(ctx));
    }
    MTM_TRY(plan_tensor_path(ctx));
    ctx->tmpl_hash_valid = true;
    return MTM_OK;
}

// Content hash of a template upload (pixels, sizes, channel count, dtype and, for transformed uploads, the
// transform list): an identical re-submission keeps everything derived from it resident.
struct TmplHasher {
    uint64_t h;
    explicit TmplHasher(uint64_t seed) : h(0x9E3779B97F4A7C15ull ^ seed) {}
    void mix(uint64_t v) { h ^= v; h *= 0x100000001B3ull; h ^= h >> 29; }
    void bytes(const void* p, size_t n) {
        const uint8_t* b = static_cast<const uint8_t*>(p);
        size_t k = 0;
        for (; k + 8 <= n; k += 8) { uint64_t v; memcpy(&v, b + k, 8); mix(v); }
        uint64_t tail = 0;
        if (k < n) { memcpy(&tail, b + k, n - k); mix(tail ^ ((uint64_t)(n - k) << 56)); }
    }
};

// A hash match alone does not prove that the resident template set is the submitted one (a 64-bit non-cryptographic hash can
// collide): the pixels are compared with the copy kept from the last upload.
static bool same_pixels(const mtm_ctx* ctx, int n, const void* const* pixels, const int32_t* h, const int32_t* w, size_t elem_bytes)
{
    size_t off = 0;
    for (int t = 0; t < n; ++t) {
        const size_t bytes = (size_t)h[t] * w[t] * elem_bytes;
        if (off + bytes > ctx->h_tmpl_copy.size() || memcmp(ctx->h_tmpl_copy.data() + off, pixels[t], bytes) != 0) return false;
        off += bytes;
    }
    return off == ctx->h_tmpl_copy.size();
}
static void keep_pixels(mtm_ctx* ctx, int n, const void* const* pixels, const int32_t* h, const int32_t* w, size_t elem_bytes)
{
    size_t total = 0;
    for (int t = 0; t < n; ++t) total += (size_t)h[t] * w[t] * elem_bytes;
    ctx->h_tmpl_copy.resize(total);
    size_t off = 0;
    for (int t = 0; t < n; ++t) {
        const size_t bytes = (size_t)h[t] * w[t] * elem_bytes;
        memcpy(ctx->h_tmpl_copy.data() + off, pixels[t], bytes);
        off += bytes;
    }
}

// The candidate list of a search overflowed (more than MTM_CAND_CAP pixels above the threshold): the peak search must stream
// the score maps.  A hits-only search has none: compute them (no list this time).
int candidates_overflowed(mtm_ctx* ctx, int method)
{
    ctx->cand_valid = false;
    if (!ctx->maps_resident) {
        ctx->cand_on = false;
        MTM_TRY(compute_maps(ctx, method, -1, false));
    }
    return MTM_OK;
}

// Downloads header + hits of a block into h_stage.  Returns the raw count in *n_raw.
int download_block(mtm_ctx* ctx, const uint8_t* d_block, int* n_raw, int* n_valid, int* declined)
{
    const int PRE = 256;
    const int pre = std::min(PRE, ctx->hit_cap);
    const size_t first = MTM_HIT_HEADER + (size_t)pre * sizeof(DevHit);
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage, d_block, first, cudaMemcpyDeviceToHost, ctx->stream));
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->ctr.d2h_bytes += (int64_t)first;
    const int32_t* hdr = reinterpret_cast<const int32_t*>(ctx->h_stage);
    const int n = hdr[0];
    *n_valid = n;
    *n_raw = std::max(hdr[0], hdr[1]);
    if (declined) { *declined = hdr[2]; if (hdr[2]) return MTM_OK; }
    if (n > pre && n <= ctx->hit_cap) {
        const size_t rest = (size_t)(n - pre) * sizeof(DevHit);
        MTM_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage + first, d_block + first, rest, cudaMemcpyDeviceToHost, ctx->stream));
        MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->ctr.d2h_bytes += (int64_t)rest;
    }
    return MTM_OK;
}

// Same contract as download_block for a result that finalize_small_kernel also stored in the mapped mirror:
// one stream synchronise, no copy engine round trip.  Results longer than the mirror fetch the rest from the block.
int download_mirror(mtm_ctx* ctx, const uint8_t* d_block, int* n_raw, int* n_valid, int* declined)
{
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int32_t* hdr = reinterpret_cast<const int32_t*>(ctx->h_mirror);
    const int n = hdr[0];
    *n_valid = n;
    *n_raw = std::max(hdr[0], hdr[1]);
    *declined = hdr[2];
    const int have = std::min(std::max(n, 0), MTM_MIRROR_HITS);
    memcpy(ctx->h_stage, ctx->h_mirror, MTM_HIT_HEADER + (size_t)have * sizeof(DevHit));
    ctx->ctr.d2h_bytes += (int64_t)(MTM_HIT_HEADER + (size_t)have * sizeof(DevHit));
    if (hdr[2]) return MTM_OK;
    if (n > have && n <= ctx->hit_cap) {
        const size_t first = MTM_HIT_HEADER + (size_t)have * sizeof(DevHit), rest = (size_t)(n - have) * sizeof(DevHit);
        MTM_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage + first, d_block + first, rest, cudaMemcpyDeviceToHost, ctx->stream));
        MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->ctr.d2h_bytes += (int64_t)rest;
    }
    return MTM_OK;
}

void copy_out(const mtm_ctx* ctx, mtm_hit* hits, int n)
{
    const DevHit* src = reinterpret_cast<const DevHit*>(ctx->h_stage + MTM_HIT_HEADER);
    for (int i = 0; i < n; ++i) {
        hits[i].tmpl = src[i].tmpl; hits[i].x = src[i].x; hits[i].y = src[i].y;
        hits[i].w = src[i].w; hits[i].h = src[i].h; hits[i].score = src[i].score;
    }
}

extern "C" {

int mtm_set_image(mtm_ctx* ctx, const void* pixels, int H, int W, int C, int dtype, int64_t row_stride_bytes)
{
    return set_image_impl(ctx, pixels, H, W, C, dtype, row_stride_bytes, false);
}

int mtm_set_image_device(mtm_ctx* ctx, const void* d_pixels, int H, int W, int C, int dtype, int64_t row_stride_bytes)
{
    return set_image_impl(ctx, d_pixels, H, W, C, dtype, row_stride_bytes, true);
}

int mtm_set_templates(mtm_ctx* ctx, int n, const void* const* pixels, const int32_t* h, const int32_t* w, int C, int dtype)
{
    MTM_ENTER(ctx);
    if (n <= 0 || !pixels || !h || !w) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates: empty template list");
    if (C < 1 || C > MTM_MAX_CH) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates: %d channels (1..4 supported)", C);
    if (dtype != MTM_U8 && dtype != MTM_F32 && dtype != MTM_U16) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates: unknown dtype %d", dtype);
    if (dtype == MTM_U16 && C != 1) return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "mtm_set_templates: 16-bit templates must be single channel (cast to float32 otherwise)");
    const bool u16 = dtype == MTM_U16;
    const int in_esz = dtype == MTM_F32 ? 4 : (u16 ? 2 : 1);     // element size of the caller's arrays
    const int esz = (dtype == MTM_F32 || u16) ? 4 : 1;           // element size of the device arena (16-bit -> float32)
    // Same template set as last time (content hash)?  Then everything derived from it -- packed pixels,
    // statistics, Toeplitz slabs, launch plan -- is still resident: nothing to upload or recompute.
    {
        TmplHasher hasher(((uint64_t)n << 32) ^ ((uint64_t)C << 8) ^ (uint64_t)dtype);
        for (int t = 0; t < n; ++t) {
            if (!pixels[t] || h[t] <= 0 || w[t] <= 0)
                return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates: template %d is empty", t);
            hasher.mix(((uint64_t)(uint32_t)h[t] << 32) | (uint32_t)w[t]);
            hasher.bytes(pixels[t], (size_t)h[t] * w[t] * C * in_esz);
        }
        const int dev_dtype = u16 ? MTM_F32 : dtype;
        if (ctx->n_tmpl == n && ctx->tmpl_hash == hasher.h && ctx->tmpl_C == C && ctx->tmpl_dtype == dev_dtype && ctx->tmpl_src_dtype == dtype &&
            ctx->tmpl_hash_valid && same_pixels(ctx, n, pixels, h, w, (size_t)C * in_esz)) return MTM_OK;
        ctx->tmpl_hash = hasher.h;
        ctx->tmpl_hash_valid = false;           // set again once the upload below has been queued
        keep_pixels(ctx, n, pixels, h, w, (size_t)C * in_esz);
    }
    // float32 templates whose pixels are all integers in [0, 65535] get byte planes too (see set_image_impl): with such an image
    // the numerator is then exact on the tensor cores, otherwise they are plain float32 templates
    static const bool f32_exact = !(getenv("MTM_B200_F32_EXACT") && atoi(getenv("MTM_B200_F32_EXACT")) == 0);
    bool f32_planes = dtype == MTM_F32 && C == 1 && f32_exact;
    for (int t = 0; t < n && f32_planes; ++t) {
        const float* f = static_cast<const float*>(pixels[t]);
        const size_t count = (size_t)h[t] * w[t];
        for (size_t k = 0; k < count; ++k) {
            const float v = f[k];
            if (!(v >= 0.0f && v <= 65535.0f && (float)(uint32_t)v == v)) { f32_planes = false; break; }
        }
    }
    const bool planes16 = u16 || f32_planes;
    // the previous upload may still be reading the pinned staging buffers
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_tmpl = 0; ctx->geometry_valid = false;          // a failed upload leaves "no templates set", not a half-updated list
    ctx->tmpl_src_dtype = dtype;
    ctx->h_meta.assign(n, TmplMeta{});
    size_t total = 0;
    for (int t = 0; t < n; ++t) {
        if (!pixels[t] || h[t] <= 0 || w[t] <= 0)
            return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates: template %d is empty", t);
        TmplMeta& m = ctx->h_meta[t];
        m.h = h[t]; m.w = w[t];
        m.wp = (w[t] * C * esz + 3) / 4 * 4;
        m.pix_off = (int64_t)total;
        total += ((size_t)m.wp * m.h + 15) / 16 * 16;
    }
    MTM_TRY(reserve_pinned(ctx, ctx->h_tmpl_stage, ctx->tmpl_stage_cap, total));
    MTM_TRY(reserve_pinned(ctx, ctx->h_geom, ctx->geom_cap, (size_t)n));
    memset(ctx->h_tmpl_stage, 0, total);
    for (int t = 0; t < n; ++t) {
        const TmplMeta& m = ctx->h_meta[t];
        const uint8_t* src = static_cast<const uint8_t*>(pixels[t]);
        uint8_t* dst = ctx->h_tmpl_stage + m.pix_off;
        const size_t row = (size_t)m.w * C * esz;
        if (u16) {                                                 // the reference's uint16 -> float32 cast
            for (int y = 0; y < m.h; ++y) {
                const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src) + (size_t)y * m.w;
                float* d32 = reinterpret_cast<float*>(dst + (size_t)y * m.wp);
                for (int x = 0; x < m.w; ++x) d32[x] = (float)s16[x];
            }
        } else {
            for (int y = 0; y < m.h; ++y) memcpy(dst + (size_t)y * m.wp, src + (size_t)y * row, row);
        }
    }
    if (planes16) {
        // byte planes for the exact tensor-core numerator: [high-byte arena][low-byte arena], u8 packing (pitch w rounded up to 4)
        std::vector<TmplPix8> pix8((size_t)n);
        size_t total8 = 0;
        for (int t = 0; t < n; ++t) {
            pix8[t].off = (int64_t)total8; pix8[t].wp = (w[t] + 3) / 4 * 4; pix8[t].pad = 0;
            total8 += ((size_t)pix8[t].wp * h[t] + 15) / 16 * 16;
        }
        std::vector<uint8_t> planes(2 * total8, 0);
        for (int t = 0; t < n; ++t) {
            const uint16_t* s16 = static_cast<const uint16_t*>(pixels[t]);
            const float* f32 = static_cast<const float*>(pixels[t]);
            for (int y = 0; y < h[t]; ++y)
                for (int x = 0; x < w[t]; ++x) {
                    const uint16_t v = u16 ? s16[(size_t)y * w[t] + x] : (uint16_t)f32[(size_t)y * w[t] + x];
                    planes[(size_t)pix8[t].off + (size_t)y * pix8[t].wp + x] = (uint8_t)(v >> 8);
                    planes[total8 + (size_t)pix8[t].off + (size_t)y * pix8[t].wp + x] = (uint8_t)(v & 255u);
                }
        }
        MTM_TRY(mtm_reserve(ctx, ctx->d_tmpl8, ctx->tmpl8_cap, 2 * total8 + 64));
        MTM_TRY(mtm_reserve(ctx, ctx->d_pix8, ctx->pix8_cap, (size_t)n));
        ctx->tmpl8_plane = (int64_t)total8;
        // pageable sources: the runtime stages these copies before returning
        MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmpl8, planes.data(), 2 * total8, cudaMemcpyHostToDevice, ctx->stream));
        MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_pix8, pix8.data(), (size_t)n * sizeof(TmplPix8), cudaMemcpyHostToDevice, ctx->stream));
        ctx->ctr.h2d_bytes += (int64_t)(2 * total8);
    }
    ctx->tmpl_u16 = planes16;
    if (u16) dtype = MTM_F32;
    MTM_TRY(mtm_reserve(ctx, ctx->d_tmpl, ctx->tmpl_cap, total + 64));
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmpl, ctx->h_tmpl_stage, total, cudaMemcpyHostToDevice, ctx->stream));
    ctx->ctr.h2d_bytes += (int64_t)total;
    return finish_templates(ctx, n, C, dtype, total);
}

int mtm_set_templates_masked(mtm_ctx* ctx, int n, const void* const* pixels, const void* const* masks,
                             const int32_t* h, const int32_t* w, int C, int dtype)
{
    MTM_ENTER(ctx);
    if (n <= 0 || !pixels || !masks || !h || !w) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_masked: empty template list");
    if (C < 1 || C > MTM_MAX_CH || C == 2) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_masked: %d channels (1, 3, 4 supported)", C);
    if (dtype != MTM_U8 && dtype != MTM_F32) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_masked: unknown dtype %d", dtype);
    const int esz = dtype == MTM_F32 ? 4 : 1;
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->tmpl_hash_valid = false;
    ctx->n_tmpl = 0; ctx->geometry_valid = false;          // a failed upload leaves "no templates set", not a half-updated list
    ctx->h_meta.assign(n, TmplMeta{});
    size_t total = 0;                                        // bytes of the float32 arenas (T*M^2 and M^2)
    for (int t = 0; t < n; ++t) {
        if (!pixels[t] || !masks[t] || h[t] <= 0 || w[t] <= 0)
            return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_masked: template %d is empty", t);
        TmplMeta& m = ctx->h_meta[t];
        m.h = h[t]; m.w = w[t];
        m.wp = w[t] * C * 4;
        m.pix_off = (int64_t)total;
        total += ((size_t)m.wp * m.h + 15) / 16 * 16;
    }
    const size_t raw_total = total / 4 * esz;
    MTM_TRY(reserve_pinned(ctx, ctx->h_tmpl_stage, ctx->tmpl_stage_cap, 2 * raw_total + 64));
    MTM_TRY(reserve_pinned(ctx, ctx->h_geom, ctx->geom_cap, (size_t)n));
    memset(ctx->h_tmpl_stage, 0, 2 * raw_total);
    for (int t = 0; t < n; ++t) {
        const TmplMeta& m = ctx->h_meta[t];
        const size_t off = (size_t)m.pix_off / 4 * esz, bytes = (size_t)m.h * m.w * C * esz;
        memcpy(ctx->h_tmpl_stage + off, pixels[t], bytes);
        memcpy(ctx->h_tmpl_stage + raw_total + off, masks[t], bytes);
    }
    MTM_TRY(mtm_reserve(ctx, ctx->d_raw_t, ctx->raw_t_cap, raw_total + 64));
    MTM_TRY(mtm_reserve(ctx, ctx->d_raw_m, ctx->raw_m_cap, raw_total + 64));
    MTM_TRY(mtm_reserve(ctx, ctx->d_tmpl, ctx->tmpl_cap, total + 64));
    MTM_TRY(mtm_reserve(ctx, ctx->d_tmpl_centred, ctx->tmplc_cap, total + 64));
    MTM_TRY(mtm_reserve(ctx, ctx->d_meta, ctx->meta_cap, (size_t)n));
    MTM_TRY(mtm_reserve(ctx, ctx->d_order, ctx->order_cap, (size_t)n));
    MTM_TRY(mtm_reserve(ctx, ctx->d_nontrivial, ctx->per_tmpl_cap, (size_t)n));
    MTM_TRY(mtm_reserve(ctx, ctx->d_best, ctx->best_cap, (size_t)n));
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_raw_t, ctx->h_tmpl_stage, raw_total, cudaMemcpyHostToDevice, ctx->stream));
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_raw_m, ctx->h_tmpl_stage + raw_total, raw_total, cudaMemcpyHostToDevice, ctx->stream));
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_meta, ctx->h_meta.data(), (size_t)n * sizeof(TmplMeta), cudaMemcpyHostToDevice, ctx->stream));
    ctx->h_order.resize(n);
    for (int t = 0; t < n; ++t) ctx->h_order[t] = t;
    std::stable_sort(ctx->h_order.begin(), ctx->h_order.end(), [&](int a, int b) {
        const TmplMeta &x = ctx->h_meta[a], &y = ctx->h_meta[b];
        return x.h != y.h ? x.h < y.h : x.w < y.w;
    });
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_order, ctx->h_order.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->ctr.h2d_bytes += (int64_t)2 * raw_total + (int64_t)n * (sizeof(TmplMeta) + sizeof(int32_t));
    ctx->n_tmpl = n; ctx->tmpl_C = C; ctx->tmpl_dtype = MTM_F32; ctx->tmpl_u16 = false;
    ctx->geometry_valid = false;
    ctx->masked = true;
    ctx->tc_groups.clear(); ctx->tc_ready = false;
    MTM_TRY(launch_masked_prep(ctx, ctx->d_raw_t, ctx->d_raw_m, dtype == MTM_F32));
    return MTM_OK;
}

int mtm_set_templates_transformed(mtm_ctx* ctx, int n, const void* const* pixels, const int32_t* h, const int32_t* w,
                                  int C, int dtype, int n_ops, const int32_t* ops, int downscale)
{
    MTM_ENTER(ctx);
    if (n <= 0 || !pixels || !h || !w) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_transformed: empty template list");
    if (n_ops <= 0 || !ops) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_transformed: empty transform list");
    if (C < 1 || C > MTM_MAX_CH) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_transformed: %d channels (1..4 supported)", C);
    if (dtype != MTM_U8 && dtype != MTM_F32)
        return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "mtm_set_templates_transformed: dtype %d (uint8 and float32 templates only)", dtype);
    if (downscale < 1 || downscale > MTM_MAX_DOWNSCALE)
        return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_transformed: downscale %d outside [1, %d]", downscale, MTM_MAX_DOWNSCALE);
    for (int k = 0; k < n_ops; ++k)
        if (ops[k] < MTM_XF_IDENTITY || ops[k] > MTM_XF_ANTITRANSPOSE)
            return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_transformed: unknown transform %d", ops[k]);
    const int esz = dtype == MTM_F32 ? 4 : 1;
    const int f = downscale;
    const int n_out = n * n_ops;
    TmplHasher hasher(((uint64_t)n_out << 32) ^ ((uint64_t)C << 8) ^ (uint64_t)dtype ^ 0x7466000000000000ull ^ ((uint64_t)f << 16));
    for (int k = 0; k < n_ops; ++k) hasher.mix((uint64_t)ops[k] + 1);
    size_t raw_total = 0;
    for (int t = 0; t < n; ++t) {
        if (!pixels[t] || h[t] <= 0 || w[t] <= 0)
            return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_transformed: template %d is empty", t);
        if (h[t] / f < 1 || w[t] / f < 1)
            return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_templates_transformed: template %d (%d x %d) vanishes at downscale %d", t, h[t], w[t], f);
        hasher.mix(((uint64_t)(uint32_t)h[t] << 32) | (uint32_t)w[t]);
        hasher.bytes(pixels[t], (size_t)h[t] * w[t] * C * esz);
        raw_total += ((size_t)h[t] * w[t] * C * esz + 15) / 16 * 16;
    }
    if (ctx->n_tmpl == n_out && ctx->tmpl_hash == hasher.h && ctx->tmpl_C == C && ctx->tmpl_dtype == dtype && !ctx->tmpl_u16 &&
        !ctx->masked && ctx->tmpl_hash_valid && same_pixels(ctx, n, pixels, h, w, (size_t)C * esz)) return MTM_OK;
    ctx->tmpl_hash = hasher.h;
    ctx->tmpl_hash_valid = false;
    keep_pixels(ctx, n, pixels, h, w, (size_t)C * esz);
    // the previous upload may still be reading the pinned staging buffer
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MTM_TRY(reserve_pinned(ctx, ctx->h_tmpl_stage, ctx->tmpl_stage_cap, raw_total + 64));
    MTM_TRY(reserve_pinned(ctx, ctx->h_geom, ctx->geom_cap, (size_t)n_out));
    ctx->n_tmpl = 0; ctx->geometry_valid = false;          // a failed upload leaves "no templates set", not a half-updated list
    ctx->h_meta.assign(n_out, TmplMeta{});
    std::vector<XformDesc> descs((size_t)n_out);
    size_t src_off = 0, total = 0;
    int64_t max_pixels = 0;
    for (int t = 0; t < n; ++t) {
        const size_t bytes = (size_t)h[t] * w[t] * C * esz;
        memcpy(ctx->h_tmpl_stage + src_off, pixels[t], bytes);
        const int dh = h[t] / f, dw = w[t] / f;
        for (int k = 0; k < n_ops; ++k) {
            const int op = ops[k];
            const bool swap = (op == MTM_XF_ROT90 || op == MTM_XF_ROT270 || op == MTM_XF_TRANSPOSE || op == MTM_XF_ANTITRANSPOSE);
            TmplMeta& m = ctx->h_meta[(size_t)t * n_ops + k];
            m.h = swap ? dw : dh; m.w = swap ? dh : dw;
            m.wp = (m.w * C * esz + 3) / 4 * 4;
            m.pix_off = (int64_t)total;
            total += ((size_t)m.wp * m.h + 15) / 16 * 16;
            XformDesc& d = descs[(size_t)t * n_ops + k];
            d.src_off = (int64_t)src_off; d.src_pitch = (int64_t)w[t] * C * esz;
            d.dst_off = m.pix_off; d.dst_pitch = m.wp;
            d.dh = dh; d.dw = dw; d.oh = m.h; d.ow = m.w; d.op = op; d.pad = 0;
            max_pixels = std::max(max_pixels, (int64_t)m.h * m.w);
        }
        src_off += (bytes + 15) / 16 * 16;
    }
    MTM_TRY(mtm_reserve(ctx, ctx->d_raw_t, ctx->raw_t_cap, raw_total + 64));
    MTM_TRY(mtm_reserve(ctx, ctx->d_tmpl, ctx->tmpl_cap, total + 64));
    MTM_TRY(mtm_reserve(ctx, ctx->d_xform, ctx->xform_cap, (size_t)n_out));
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_raw_t, ctx->h_tmpl_stage, raw_total, cudaMemcpyHostToDevice, ctx->stream));
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_xform, descs.data(), (size_t)n_out * sizeof(XformDesc), cudaMemcpyHostToDevice, ctx->stream));
    ctx->ctr.h2d_bytes += (int64_t)raw_total + (int64_t)n_out * (int64_t)sizeof(XformDesc);
    MTM_CUDA(ctx, cudaMemsetAsync(ctx->d_tmpl, 0, total, ctx->stream));        // row padding must read as zero
    MTM_TRY(launch_transform(ctx, ctx->d_raw_t, ctx->d_tmpl, ctx->d_xform, n_out, max_pixels, C, dtype, f));
    ctx->tmpl_u16 = false;
    return finish_templates(ctx, n_out, C, dtype, total);
}

int mtm_set_image_scaled(mtm_ctx* ctx, const void* pixels, int H, int W, int C, int dtype, int64_t row_stride_bytes,
                         int downscale)
{
    MTM_ENTER(ctx);
    if (!pixels || H <= 0 || W <= 0) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image_scaled: empty image (%d x %d)", H, W);
    if (C < 1 || C > MTM_MAX_CH) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image_scaled: %d channels (1..4 supported)", C);
    if (dtype != MTM_U8 && dtype != MTM_F32 && dtype != MTM_U16) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image_scaled: unknown dtype %d", dtype);
    if (downscale < 1 || downscale > MTM_MAX_DOWNSCALE)
        return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image_scaled: downscale %d outside [1, %d]", downscale, MTM_MAX_DOWNSCALE);
    const int f = downscale;
    if (H / f < 1 || W / f < 1) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image_scaled: image %d x %d vanishes at downscale %d", H, W, f);
    const int64_t esz = dtype == MTM_F32 ? 4 : (dtype == MTM_U16 ? 2 : 1);
    const int64_t row = (int64_t)W * C * esz;
    if (row_stride_bytes < row) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image_scaled: row stride %lld < %lld", (long long)row_stride_bytes, (long long)row);
    ctx->full_dtype = -1;                                        // invalid until the upload below is queued
    MTM_TRY(mtm_reserve(ctx, ctx->d_full, ctx->full_cap, (size_t)(row * H + 64)));
    MTM_TRY(mtm_upload_rows(ctx, ctx->d_full, (size_t)row, pixels, (size_t)row_stride_bytes, (size_t)row, H));
    ctx->ctr.h2d_bytes += row * H;
    ctx->full_H = H; ctx->full_W = W; ctx->full_C = C; ctx->full_dtype = dtype;
    if (f == 1) return set_image_impl(ctx, ctx->d_full, H, W, C, dtype, row, true);
    const int sh = H / f, sw = W / f;
    const int64_t srow = (int64_t)sw * C * esz;
    MTM_TRY(mtm_reserve(ctx, ctx->d_small, ctx->small_cap, (size_t)(srow * sh + 64)));
    MTM_TRY(mtm_reserve(ctx, ctx->d_xform, ctx->xform_cap, (size_t)1));
    XformDesc d{};
    d.src_off = 0; d.src_pitch = row; d.dst_off = 0; d.dst_pitch = srow;
    d.dh = sh; d.dw = sw; d.oh = sh; d.ow = sw; d.op = MTM_XF_IDENTITY;
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_xform, &d, sizeof d, cudaMemcpyHostToDevice, ctx->stream));   // pageable: staged before returning
    MTM_TRY(launch_transform(ctx, ctx->d_full, ctx->d_small, ctx->d_xform, 1, (int64_t)sh * sw, C, dtype, f));
    return set_image_impl(ctx, ctx->d_small, sh, sw, C, dtype, srow, true);
}

int mtm_set_image_roi(mtm_ctx* ctx, int x, int y, int w, int h)
{
    MTM_ENTER(ctx);
    if (ctx->full_dtype < 0) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image_roi: no full-resolution image resident (call mtm_set_image_scaled first)");
    if (x < 0 || y < 0 || w <= 0 || h <= 0 || (int64_t)x + w > ctx->full_W || (int64_t)y + h > ctx->full_H)
        return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_set_image_roi: region (%d, %d, %d, %d) outside the %d x %d image", x, y, w, h,
                        ctx->full_W, ctx->full_H);
    const int64_t esz = ctx->full_dtype == MTM_F32 ? 4 : (ctx->full_dtype == MTM_U16 ? 2 : 1);
    const int64_t row = (int64_t)ctx->full_W * ctx->full_C * esz;
    const uint8_t* origin = ctx->d_full + (int64_t)y * row + (int64_t)x * ctx->full_C * esz;
    return set_image_impl(ctx, origin, h, w, ctx->full_C, ctx->full_dtype, row, true);
}

int mtm_score_map(mtm_ctx* ctx, int tmpl, int method, float* out_host, int64_t out_elems)
{
    MTM_ENTER(ctx);
    MTM_TRY(ensure_geometry(ctx));
    if (tmpl < 0 || tmpl >= ctx->n_tmpl) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_score_map: template index %d out of range", tmpl);
    const TmplMeta& m = ctx->h_meta[tmpl];
    const int64_t need = (int64_t)m.mh * m.mw;
    if (!out_host || out_elems < need) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_score_map: output buffer holds %lld floats, need %lld",
                                                       (long long)out_elems, (long long)need);
    MTM_TRY(compute_maps(ctx, method, tmpl));
    MTM_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->d_maps + m.map_off, (size_t)need * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->ctr.d2h_bytes += need * (int64_t)sizeof(float);
    return MTM_OK;
}

int mtm_find_matches(mtm_ctx* ctx, int method, int64_t n_object, double score_threshold,
                     mtm_hit* hits, int capacity, int* n_hits)
{
    MTM_ENTER(ctx);
    if (!n_hits || (capacity > 0 && !hits)) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_find_matches: null output");
    MTM_TRY(ensure_geometry(ctx));
    request_candidates(ctx, method, n_object, score_threshold);
    MTM_TRY(compute_maps(ctx, method, -1, true));
    const int minimize = method_is_min(method) ? 1 : 0;
    for (int attempt = 0; attempt < 8; ++attempt) {
        MTM_TRY(launch_peaks(ctx, method, n_object, (float)score_threshold, score_threshold));
        int n_raw = 0, n = 0, declined = 0;
        if (n_object != 1) {
            MTM_TRY(launch_finalize_small(ctx, minimize, 1, 0, 0, 0.f, 0, -1, 0.f));
            MTM_TRY(download_block(ctx, ctx->d_blockA, &n_raw, &n, &declined));
            if (declined == 2) { MTM_TRY(candidates_overflowed(ctx, method)); continue; }      // candidate list overflowed: stream the maps
            if (declined) {                                  // more than 1024 raw hits: general path
                if (n_raw > ctx->hit_cap) { MTM_TRY(reserve_hits(ctx, n_raw)); continue; }
                MTM_TRY(launch_sort_hits(ctx, 0, minimize, 0, 1));
                MTM_TRY(download_block(ctx, ctx->d_blockA, &n_raw, &n));
            }
        } else {
            MTM_TRY(download_block(ctx, ctx->d_blockA, &n_raw, &n));
        }
        if (n_raw > ctx->hit_cap) { MTM_TRY(reserve_hits(ctx, n_raw)); continue; }
        *n_hits = n;
        if (n > capacity) return mtm_fail(ctx, MTM_ERR_CAPACITY, "mtm_find_matches: %d hits, caller capacity %d", n, capacity);
        copy_out(ctx, hits, n);
        return MTM_OK;
    }
    return mtm_fail(ctx, MTM_ERR_CUDA, "mtm_find_matches: hit buffer kept overflowing");
}

int mtm_match_templates(mtm_ctx* ctx, int method, int64_t n_object, double score_threshold,
                        double max_overlap, mtm_hit* hits, int capacity, int* n_hits)
{
    MTM_ENTER(ctx);
    if (!n_hits || (capacity > 0 && !hits)) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates: null output");
    if (method == MTM_TM_SQDIFF) return mtm_fail(ctx, MTM_ERR_INVALID, "The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.");
    MTM_TRY(ensure_geometry(ctx));
    g_marks.mark(ctx, "geometry");
    request_candidates(ctx, method, n_object, score_threshold);
    MTM_TRY(compute_maps(ctx, method, -1, true));
    g_marks.mark(ctx, "moments+ncc");
    const int minimize = method_is_min(method) ? 1 : 0;
    const int ascending = (method == MTM_TM_SQDIFF_NORMED) ? 1 : 0;
    const float thr_nms = ascending ? (float)(1.0 - score_threshold) : (float)score_threshold;
    for (int attempt = 0; attempt < 8; ++attempt) {
        MTM_TRY(launch_peaks(ctx, method, n_object, (float)score_threshold, score_threshold));
        g_marks.mark(ctx, "peaks");
        int n_raw = 0, n = 0, declined = 0;
        MTM_TRY(launch_finalize_small(ctx, minimize, n_object != 1, n_object == 1, 1, thr_nms, ascending, n_object,
                                      (float)max_overlap, nullptr, true));
        g_marks.mark(ctx, "finalize");
        MTM_TRY(download_mirror(ctx, ctx->d_blockB, &n_raw, &n, &declined));
        g_marks.mark(ctx, "download");
        g_marks.report(ctx);
        if (declined == 2) { MTM_TRY(candidates_overflowed(ctx, method)); continue; }          // candidate list overflowed: stream the maps
        if (declined) {                                      // more than 1024 raw hits: general path
            if (n_raw > ctx->hit_cap) { MTM_TRY(reserve_hits(ctx, n_raw)); continue; }
            if (n_object != 1) {
                MTM_TRY(launch_sort_hits(ctx, 0, minimize, 0, 1));
                MTM_TRY(launch_sort_hits(ctx, 1, minimize, ascending, 0));
            }
            MTM_TRY(launch_nms(ctx, thr_nms, ascending, n_object, (float)max_overlap));
            MTM_TRY(download_block(ctx, ctx->d_blockB, &n_raw, &n));
        }
        if (n_raw > ctx->hit_cap) { MTM_TRY(reserve_hits(ctx, n_raw)); continue; }
        *n_hits = n;
        if (n > capacity) return mtm_fail(ctx, MTM_ERR_CAPACITY, "mtm_match_templates: %d hits, caller capacity %d", n, capacity);
        copy_out(ctx, hits, n);
        return MTM_OK;
    }
    return mtm_fail(ctx, MTM_ERR_CUDA, "mtm_match_templates: hit buffer kept overflowing");
}

int mtm_match_templates_async(mtm_ctx* ctx, int method, int64_t n_object, double score_threshold,
                              double max_overlap, int slot)
{
    MTM_ENTER(ctx);
    if (slot < 0 || slot >= MTM_MAX_INFLIGHT) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_async: slot %d out of range", slot);
    if (ctx->slot_busy[slot]) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_async: slot %d not collected yet", slot);
    if (method == MTM_TM_SQDIFF) return mtm_fail(ctx, MTM_ERR_INVALID, "The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.");
    const size_t bytes = MTM_HIT_HEADER + (size_t)MTM_SLOT_HITS * sizeof(DevHit);
    if (!ctx->d_slot[slot]) {
        MTM_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_slot[slot]), bytes));
        MTM_CUDA(ctx, cudaMemsetAsync(ctx->d_slot[slot], 0, bytes, ctx->stream));
        MTM_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->h_slot[slot]), bytes));
        MTM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_slot[slot], cudaEventDisableTiming));
    }
    MTM_TRY(ensure_geometry(ctx));
    request_candidates(ctx, method, n_object, score_threshold);
    MTM_TRY(compute_maps(ctx, method, -1, true));
    const int minimize = method_is_min(method) ? 1 : 0;
    const int ascending = (method == MTM_TM_SQDIFF_NORMED) ? 1 : 0;
    const float thr_nms = ascending ? (float)(1.0 - score_threshold) : (float)score_threshold;
    MTM_TRY(launch_peaks(ctx, method, n_object, (float)score_threshold, score_threshold));
    MTM_TRY(launch_finalize_small(ctx, minimize, n_object != 1, n_object == 1, 1, thr_nms, ascending, n_object,
                                  (float)max_overlap, ctx->d_slot[slot]));
    // header + the first 256 hits ride along; _collect fetches the (rare) remainder
    const size_t first = MTM_HIT_HEADER + 256 * sizeof(DevHit);
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->h_slot[slot], ctx->d_slot[slot], first, cudaMemcpyDeviceToHost, ctx->stream));
    MTM_CUDA(ctx, cudaEventRecord(ctx->ev_slot[slot], ctx->stream));
    ctx->ctr.d2h_bytes += (int64_t)first;
    ctx->slot_busy[slot] = true;
    return MTM_OK;
}

int mtm_match_templates_collect(mtm_ctx* ctx, int slot, mtm_hit* hits, int capacity, int* n_hits)
{
    MTM_ENTER(ctx);
    if (slot < 0 || slot >= MTM_MAX_INFLIGHT || !ctx->slot_busy[slot])
        return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_collect: slot %d has no submission in flight", slot);
    if (!n_hits || (capacity > 0 && !hits)) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_collect: null output");
    MTM_CUDA(ctx, cudaEventSynchronize(ctx->ev_slot[slot]));
    ctx->slot_busy[slot] = false;
    const int32_t* hdr = reinterpret_cast<const int32_t*>(ctx->h_slot[slot]);
    const int n = hdr[0];
    *n_hits = n;
    if (hdr[2]) { *n_hits = hdr[1]; return mtm_fail(ctx, MTM_ERR_CAPACITY, "mtm_match_templates_collect: %d raw peaks exceed the fast path; use mtm_match_templates", hdr[1]); }
    if (n > capacity) return mtm_fail(ctx, MTM_ERR_CAPACITY, "mtm_match_templates_collect: %d hits, caller capacity %d", n, capacity);
    if (n > 256) {
        const size_t first = MTM_HIT_HEADER + 256 * sizeof(DevHit), rest = (size_t)(n - 256) * sizeof(DevHit);
        MTM_CUDA(ctx, cudaMemcpyAsync(ctx->h_slot[slot] + first, ctx->d_slot[slot] + first, rest, cudaMemcpyDeviceToHost, ctx->stream));
        MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->ctr.d2h_bytes += (int64_t)rest;
    }
    const DevHit* src = reinterpret_cast<const DevHit*>(ctx->h_slot[slot] + MTM_HIT_HEADER);
    for (int i = 0; i < n; ++i) {
        hits[i].tmpl = src[i].tmpl; hits[i].x = src[i].x; hits[i].y = src[i].y;
        hits[i].w = src[i].w; hits[i].h = src[i].h; hits[i].score = src[i].score;
    }
    return MTM_OK;
}

int mtm_nms(mtm_ctx* ctx, const mtm_hit* hits, int n, double score_threshold, int sort_ascending,
            int64_t n_object, double max_overlap, int32_t* keep, int* n_keep)
{
    MTM_ENTER(ctx);
    if (n < 0 || !n_keep || (n > 0 && (!hits || !keep))) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_nms: null argument");
    if (n == 0) { *n_keep = 0; return MTM_OK; }
    MTM_TRY(reserve_hits(ctx, n));
    MTM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int32_t* hdr = reinterpret_cast<int32_t*>(ctx->h_stage);
    memset(hdr, 0, MTM_HIT_HEADER);
    hdr[0] = n;
    DevHit* dst = reinterpret_cast<DevHit*>(ctx->h_stage + MTM_HIT_HEADER);
    for (int i = 0; i < n; ++i) {
        dst[i].tmpl = hits[i].tmpl; dst[i].x = hits[i].x; dst[i].y = hits[i].y; dst[i].w = hits[i].w; dst[i].h = hits[i].h;
        dst[i].score = hits[i].score; dst[i].seq = i; dst[i].key = 0.f;
    }
    const size_t bytes = MTM_HIT_HEADER + (size_t)n * sizeof(DevHit);
    MTM_CUDA(ctx, cudaMemcpyAsync(ctx->d_blockA, ctx->h_stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->ctr.h2d_bytes += (int64_t)bytes;
    const int ascending = sort_ascending ? 1 : 0;
    const float thr_nms = ascending ? (float)(1.0 - score_threshold) : (float)score_threshold;
    int n_raw = 0, nk = 0, declined = 0;
    MTM_TRY(launch_finalize_small(ctx, 0, 0, 1, 1, thr_nms, ascending, n_object, (float)max_overlap));
    MTM_TRY(download_block(ctx, ctx->d_blockB, &n_raw, &nk, &declined));
    if (declined) {
        MTM_TRY(launch_sort_hits(ctx, 1, 0, ascending, 0));
        MTM_TRY(launch_nms(ctx, thr_nms, ascending, n_object, (float)max_overlap));
        MTM_TRY(download_block(ctx, ctx->d_blockB, &n_raw, &nk));
    }
    const DevHit* src = reinterpret_cast<const DevHit*>(ctx->h_stage + MTM_HIT_HEADER);
    for (int i = 0; i < nk; ++i) keep[i] = src[i].seq;
    *n_keep = nk;
    return MTM_OK;
}

}  // extern "C"

template int mtm_reserve<uint8_t>(mtm_ctx*, uint8_t*&, size_t&, size_t);
