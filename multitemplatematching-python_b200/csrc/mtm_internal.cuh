// mtm_internal.cuh -- shared declarations of libmtm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/mtm_b200.h"

#define MTM_MAX_CH 4
#define MTM_NCC_RING 256     // event brackets of MTM_OPT_TIME_NCC in flight (a banded search opens one per band)
#define MTM_CAND_CAP 32768      // above-threshold pixels the tensor-core epilogue may list per call
#define MTM_HASH_SLOTS (1 << 17) // slots of the candidate hash table (load factor <= 1/4)
#define MTM_SLOT_HITS 1024      // hits a slot of the asynchronous API can return (== the fused fast path)
#define MTM_STAGE_BUFS 4        // pinned chunks of the pageable-upload pipeline (host_staging.cu)
#define MTM_HIT_HEADER 32      // bytes: int32 count[8]; count[0] = number of hits (may exceed capacity)

// Device-side record of one hit (32 bytes); the public mtm_hit is its 24-byte prefix.
struct DevHit {
    int32_t tmpl;
    int32_t x, y, w, h;
    float score;
    int32_t seq;     // tie-break rank: position in the canonical (template, peak) order / input order
    float key;       // NMS sort key: score, or 1-score when sortAscending
};

// Per-template constants (host fills geometry, tmpl_stats_kernel fills the statistics).
struct TmplGeom {         // image-dependent part, re-uploaded when the image size changes
    int64_t map_off;      // element offset of the template's score map in the map arena
    int32_t mh, mw;       // score map size
    int64_t mom_off;      // element offset of its window-moment maps (shared by templates of one size)
};
struct TmplMeta {
    int64_t map_off;      // -- TmplGeom prefix (24 bytes) --
    int32_t mh, mw;
    int64_t mom_off;
    int64_t pix_off;      // byte offset of the packed template in the template arena
    int32_t h, w;         // template size (pixels)
    int32_t wp;           // packed row pitch in bytes (w*C rounded up to 4, zero padded)
    int32_t is_const;     // templNorm < DBL_EPSILON (OpenCV: TM_CCOEFF_NORMED map := 1)
    double mean[MTM_MAX_CH];   // meanStdDev mean per channel
    double sum2;               // sum T^2 over all channels  (templSum2 after "/= invArea")
    double norm_ccoeff;        // sqrt(sum_c var_c) / sqrt(invArea)
    double norm_plain;         // sqrt(sum_c var_c + mean_c^2) / sqrt(invArea)
    double inv_area;
    long long isum[MTM_MAX_CH]; // integer sum of the template per channel: tensor-core epilogue
    float inv_sqrt_d2;         // 1 / sqrt(sum_c (A*sumT2_c - sumT_c^2))
    float pad_f;
};

// Where a template's packed u8 pixels live inside a byte-plane arena (MTM_U16 templates).
struct TmplPix8 {
    int64_t off;
    int32_t wp, pad;
};

// One output of transform_kernel (transform.cu): `op` of the f-times area-reduced source array.
struct XformDesc {
    int64_t src_off;      // byte offset of the source array in the source buffer
    int64_t src_pitch;    // bytes between source rows
    int64_t dst_off;      // byte offset of the output in the destination buffer
    int64_t dst_pitch;    // bytes between output rows
    int32_t dh, dw;       // size of the reduced source (source size / f, rounded down)
    int32_t oh, ow;       // output size: (dh, dw), swapped by the transposing symmetries
    int32_t op, pad;      // mtm_transform
};

// One distinct template size: where its window moments live.  Moments are produced per template group and per BAND of output
// rows, into a ring the next band overwrites (it stays in L2: the moment stage writes to DRAM only what the cache evicts).  A
// size's segment holds `band` rows in tile-major order, the order the tcgen05 epilogue reads them in:
//     entry(x, r) = ((x >> 4) * band + r) * 16 + (x & 15),   r = y - first row of the band
// i.e. the 16 x-offsets of a tile column are one 128-byte line and consecutive rows follow each other 128 bytes apart.
struct SizeDesc {
    int32_t h, w, mh, mw;
    int64_t off;          // element offset of the segment in the ring
    int32_t band, pad;    // rows per band of this size's group
};
__host__ __device__ inline int64_t mom_index(int x, int r, int band) { return ((int64_t)(x >> 4) * band + r) * 16 + (x & 15); }
__host__ __device__ inline int64_t mom_segment(int mw, int band) { return (int64_t)((mw + 15) >> 4) * 16 * band; }

// One launch of the tcgen05 kernel: `count` templates d_order[first .. first+count); (h, w) is the
// group's padded size (mode A may mix sizes: smaller templates are zero padded in the Toeplitz slabs).
struct TcGroup {
    int mode;                  // 0: 8 templates x 16 x-offsets, 1: 1 template x 128 x-offsets
    int first, count;
    int h, w, nk, a_kblk, slab_bytes, ds, N, R;
    size_t smem;
    int64_t arena_off;         // byte offset of the group's Toeplitz slabs in d_slabs
    int h_min, w_min;          // smallest member (largest score map): the tile grid covers its map
    int size_first, size_count;// its distinct sizes: ctx->h_sizes[size_first .. size_first + size_count)   (ensure_geometry)
    int band_rows;             // output rows per band of window moments (>= the largest member map: one band)
    double eff;
};

struct ImageDev {
    uint8_t* pix = nullptr;      // u8, interleaved channels, zero padded rows
    int64_t pitch = 0;           // bytes
    int H = 0, W = 0, C = 0;
    uint32_t* sat_s = nullptr;   // C tables of (H+1) x sat_pitch, wrap-around u32 (window sums < 2^32)
    unsigned long long* sat_q = nullptr;  // (H+1) x sat_pitch, sum over channels of I^2
    uint32_t* sat_q32 = nullptr;          // the same table modulo 2^32 (exact for window sums < 2^32: tensor path)
    int64_t sat_pitch = 0;       // elements
    // float32 images (MTM/__init__.py:71-74): pixels + float64 summed-area tables
    float* pixf = nullptr; int64_t pitch_e = 0;
    uint8_t* pix_lo = nullptr;           // MTM_U16 images: `pix` holds the high bytes, `pix_lo` the low bytes (same pitch)
    float* pixf2 = nullptr;              // image squared (masked matching)
    double* satf_s = nullptr; double* satf_q = nullptr;
};

struct mtm_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    mtm_counters ctr{};
    int path = MTM_PATH_AUTO;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // MTM_OPT_TIME_NCC: ring of event pairs bracketing the numerator kernels of recent compute_maps calls
    cudaEvent_t ev_ncc[MTM_NCC_RING][2] = {};
    int ncc_launches_of[MTM_NCC_RING] = {};
    int time_ncc = 0, ncc_head = 0, ncc_tail = 0;       // [tail, head) are recorded but not yet folded into ctr
    int64_t ncc_mark = 0;                               // kernel_launches when the open bracket began
    int sm_count = 0;

    // image
    ImageDev img;
    size_t img_cap = 0, sat_s_cap = 0, sat_q_cap = 0, sat_q32_cap = 0, scratch_cap = 0, imgf_cap = 0, satf_s_cap = 0, satf_q_cap = 0;
    uint32_t* scratch = nullptr;         // row-prefix scratch for the SAT build
    bool sat_valid = false;              // the u8 summed-area tables belong to the resident image (built on demand under MTM_B200_MOM_BOX)
    int img_dtype = -1;

    // full-resolution image of mtm_set_image_scaled (source of the reduced image and of mtm_set_image_roi)
    uint8_t* d_full = nullptr; size_t full_cap = 0;
    int full_H = 0, full_W = 0, full_C = 0, full_dtype = -1;
    uint8_t* d_small = nullptr; size_t small_cap = 0;          // its reduced copy (contiguous), handed to set_image
    XformDesc* d_xform = nullptr; size_t xform_cap = 0;        // descriptors of the transform launches

    // templates
    int n_tmpl = 0;
    std::vector<TmplMeta> h_meta;
    TmplMeta* d_meta = nullptr; size_t meta_cap = 0;
    uint8_t* d_tmpl = nullptr; size_t tmpl_cap = 0;
    uint8_t* d_tmpl_centred = nullptr; size_t tmplc_cap = 0;   // float32 templates minus their mean
    uint8_t* h_tmpl_stage = nullptr; size_t tmpl_stage_cap = 0;   // pinned
    int tmpl_C = 0, tmpl_dtype = -1;
    // MTM_U16 uploads: float32 everywhere (img_dtype / tmpl_dtype == MTM_F32) plus the byte planes for the exact tensor-core numerator
    bool img_u16 = false, tmpl_u16 = false;                    // byte planes are resident (MTM_U16 uploads, and float32 data found to be integers in [0, 65535])
    int img_src_dtype = -1, tmpl_src_dtype = -1;               // dtypes the resident image / templates were submitted with
    uint16_t* d_raw16 = nullptr; size_t raw16_cap = 0;         // staging of the raw 16-bit image
    size_t pix_lo_cap = 0;
    uint8_t* d_tmpl8 = nullptr; size_t tmpl8_cap = 0;          // packed u8 templates: high-byte arena, then low-byte arena
    int64_t tmpl8_plane = 0;                                   // bytes between the two arenas
    TmplPix8* d_pix8 = nullptr; size_t pix8_cap = 0;           // per-template offset / pitch inside an arena
    int64_t slab_plane = 0;                                    // bytes between the high- and low-byte Toeplitz slabs in d_slabs
    double* d_acc = nullptr; size_t acc_cap = 0;               // exact numerator maps (double) of the 16-bit path
    uint64_t tmpl_hash = 0; bool tmpl_hash_valid = false;   // content hash of the resident template set ...
    std::vector<uint8_t> h_tmpl_copy;                       // ... and the submitted pixels themselves (compared on a hash match)
    bool geometry_valid = false;         // map offsets computed for (image, templates)
    bool masked = false;                 // templates carry masks (methods 0 / 3): d_tmpl = T*M^2, d_tmpl_centred = M^2
    bool masked_image_valid = false;     // pixf / pixf2 hold the current image
    uint8_t* d_raw_t = nullptr; uint8_t* d_raw_m = nullptr; size_t raw_t_cap = 0, raw_m_cap = 0;
    float* d_maps2 = nullptr; size_t maps2_cap = 0; size_t pixf2_cap = 0;

    // score maps
    float* d_maps = nullptr; size_t maps_cap = 0;   // elements
    int64_t maps_total = 0, moments_total = 0;
    int maps_method = -1;                // method the resident maps were computed with (-1: stale)

    // tensor-core path (ncc_tc.cu)
    std::vector<TcGroup> tc_groups;                      // covers every template when tc_ready
    bool tc_ready = false;
    uint8_t* d_slabs = nullptr; size_t slabs_cap = 0;    // Toeplitz-expanded template rows
    std::vector<SizeDesc> h_sizes;                       // distinct (h, w) of the current templates
    SizeDesc* d_sizes = nullptr; size_t sizes_cap = 0;
    uint32_t* d_wS = nullptr; size_t wS_cap = 0;         // window sums S
    float* d_wR = nullptr; size_t wR_cap = 0;            // rsqrt(A*Q - S^2)
    bool moments_valid = false;                          // (unused since the moments are produced per group and band)
    bool moments_ring = false;                           // moments per group and band into a reused ring (MTM_B200_RING_KB) instead of resident maps
    bool box_ok = false;                                 // window moments by box sums straight from the image (banded); else summed-area tables
    bool tc_attr_set = false;
    bool tcp_attr_set = false;

    int32_t* d_order = nullptr; size_t order_cap = 0;   // template indices sorted by (h, w)
    std::vector<int32_t> h_order;
    TmplGeom* h_geom = nullptr; size_t geom_cap = 0;     // pinned

    // hits: two device blocks, each = HitHeader + DevHit[hit_cap]
    uint8_t* d_blockA = nullptr; uint8_t* d_blockB = nullptr;
    int hit_cap = 0;                     // device capacity per block (power of two)
    int32_t* d_nontrivial = nullptr;     // per template: some pixel is not a 3x3 max
    unsigned long long* d_best = nullptr;// per template arg-max key
    int32_t* d_keep = nullptr;           // kept indices (capacity hit_cap)
    size_t per_tmpl_cap = 0, best_cap = 0;
    uint8_t* h_stage = nullptr; size_t h_stage_cap = 0;   // pinned up/download buffer
    uint8_t* h_mirror = nullptr; uint8_t* d_mirror = nullptr;   // mapped pinned result mirror (header + MTM_MIRROR_HITS hits) and its device alias

    // candidate list written by the tcgen05 epilogue (pixels above the threshold), consumed by verify_candidates
    DevHit* d_cand = nullptr; int32_t* d_cand_count = nullptr;
    // hits-only searches (MODE 3 of the tcgen05 kernels): no score map is written; the peak search works on the candidate list
    // (3x3 maxima resolved inside the list, resolve_candidates_kernel) or on the per-template arg-max keys of the epilogue
    bool want_n1 = false;                // the pending search is the N_object == 1 one (request_candidates)
    bool hits_only = false;              // the current / last compute_maps call ran without score maps
    bool best_on = false;                // ... and its epilogues raise d_best (N_object == 1)
    bool best_valid = false;             // d_best holds the arg-max keys of the resident search
    bool maps_resident = false;          // d_maps holds the score maps of the resident (image, templates, method)
    unsigned long long* d_hkeys = nullptr; int32_t* d_hvals = nullptr;   // open-addressing table of resolve_candidates_kernel (kept empty between calls)
    bool cand_on = false;                // the current compute_maps call fills the list
    bool cand_valid = false;             // the list belongs to the resident score maps
    float cand_thr = 0.f;

    // pageable-upload pipeline (host_staging.cu): pinned chunks + the event of each chunk's last DMA
    uint8_t* h_chunk[MTM_STAGE_BUFS] = {};
    cudaEvent_t ev_chunk[MTM_STAGE_BUFS] = {};
    bool chunk_used[MTM_STAGE_BUFS] = {};
    int chunk_next = 0;

    // asynchronous submissions (mtm_match_templates_async / _collect): per-slot result blocks
    uint8_t* d_slot[MTM_MAX_INFLIGHT] = {};        // header + DevHit[MTM_SLOT_HITS]
    uint8_t* h_slot[MTM_MAX_INFLIGHT] = {};        // pinned mirrors
    cudaEvent_t ev_slot[MTM_MAX_INFLIGHT] = {};
    bool slot_busy[MTM_MAX_INFLIGHT] = {};

    int32_t* countA() const { return reinterpret_cast<int32_t*>(d_blockA); }
    int32_t* countB() const { return reinterpret_cast<int32_t*>(d_blockB); }
    DevHit* hitsA() const { return reinterpret_cast<DevHit*>(d_blockA + MTM_HIT_HEADER); }
    DevHit* hitsB() const { return reinterpret_cast<DevHit*>(d_blockB + MTM_HIT_HEADER); }
};

int mtm_fail(mtm_ctx* ctx, int code, const char* fmt, ...);
#define MTM_CUDA(ctx, call)                                                          \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess)                                                      \
            return mtm_fail(ctx, MTM_ERR_CUDA, "%s failed: %s (%s:%d)", #call,       \
                            cudaGetErrorString(e__), __FILE__, __LINE__);            \
    } while (0)
#define MTM_LAUNCH_CHECK(ctx)                                                        \
    do {                                                                             \
        (ctx)->ctr.kernel_launches++;                                                \
        MTM_CUDA(ctx, cudaGetLastError());                                           \
    } while (0)

#define MTM_TRY(expr) do { int rc__ = (expr); if (rc__ != MTM_OK) return rc__; } while (0)
#define MTM_ENTER(ctx)                                                        \
    if (!(ctx)) return MTM_ERR_INVALID;                                       \
    MTM_CUDA(ctx, cudaSetDevice((ctx)->device))

template <typename T>
int mtm_reserve(mtm_ctx* ctx, T*& ptr, size_t& cap, size_t need_elems);

// host rows -> device rows on the context's stream; pageable sources go through pinned chunks filled by a host thread pool
int mtm_upload_rows(mtm_ctx* ctx, void* dst, size_t dst_pitch, const void* src, size_t src_stride, size_t row_bytes, int H);

// ---- host-side stages of a search (mtm_api.cu), shared with the multi-GPU entry points (mtm_comm.cu)
int reserve_hits(mtm_ctx* ctx, int cap);
int ensure_geometry(mtm_ctx* ctx);
void request_candidates(mtm_ctx* ctx, int method, int64_t n_object, double thr);
int candidates_overflowed(mtm_ctx* ctx, int method);
int compute_maps(mtm_ctx* ctx, int method, int tmpl, bool hits_ok = false);   // hits_ok: the caller only needs the hit list (no score map read-back)
int download_block(mtm_ctx* ctx, const uint8_t* d_block, int* n_raw, int* n_valid, int* declined = nullptr);
int download_mirror(mtm_ctx* ctx, const uint8_t* d_block, int* n_raw, int* n_valid, int* declined);
void copy_out(const mtm_ctx* ctx, mtm_hit* hits, int n);

// ---- kernels (defined in the .cu files) -------------------------------------
int launch_build_sat(mtm_ctx* ctx);
int launch_tmpl_stats(mtm_ctx* ctx);
// templates d_order[first .. first+count) share (h, w)
int launch_ncc_direct(mtm_ctx* ctx, int method, int first, int count);
// small score maps of large uint8 templates (ncc_points.cu); same group convention
bool points_path_preferred(const mtm_ctx* ctx, int first, int count);
int launch_ncc_points(mtm_ctx* ctx, int method, int first, int count);
// float32 branch (ncc_float.cu)
int launch_build_sat_f32(mtm_ctx* ctx);
int launch_tmpl_stats_f32(mtm_ctx* ctx);
int launch_ncc_direct_f32(mtm_ctx* ctx, int method, int first, int count, const float* img_override = nullptr,
                          const uint8_t* tmpl_override = nullptr, float* maps_override = nullptr);
// masked matching, methods 0 / 3 (ncc_float.cu)
int launch_masked_prep(mtm_ctx* ctx, const uint8_t* d_raw_t, const uint8_t* d_raw_m, int is_f32);
int launch_masked_image(mtm_ctx* ctx);
int launch_masked_combine(mtm_ctx* ctx, int method, const float* mapsB);
// tensor-core path
bool tc_path_supported(const mtm_ctx* ctx, int method, int h, int w);
bool tc_plan_group(int mode, int h, int w, int C, TcGroup& g);
int launch_toeplitz_prep(mtm_ctx* ctx, const TcGroup& g);
// window moments of the sizes h_sizes[size_first .. +size_count) for output rows [y_begin, y_begin + rows) -> the ring
int launch_window_moments(mtm_ctx* ctx, int size_first, int size_count);          // summed-area route: whole maps only (one band)
// experiment knob MTM_B200_MOM_BOX (box_moments.cu): the same moment maps from running box sums, without summed-area tables
bool box_moments_enabled();
bool box_moments_applicable(const mtm_ctx* ctx);
int launch_box_moments(mtm_ctx* ctx, int size_first, int size_count, int y_begin, int rows);
// rows [y_base, y_base + rows) of the group's score maps (rows counted on its largest map)
int launch_ncc_tc(mtm_ctx* ctx, const TcGroup& g, int method, int y_base, int rows);
int launch_i8_peak(mtm_ctx* ctx, int n, int iters);      // measurement helper: back-to-back kind::i8 MMAs, no loads, no epilogue
// 16-bit path: one byte-plane product of the group accumulated into ctx->d_acc (img_plane / tmpl_plane: 0 = high, 1 = low bytes)
int launch_ncc_tc_accum(mtm_ctx* ctx, const TcGroup& g, int img_plane, int tmpl_plane, double weight, bool first);
int launch_u16_split_image(mtm_ctx* ctx, const uint16_t* src, int64_t src_stride_bytes);
// byte planes of the resident float32 image (im.pixf) when every pixel is an integer in [0, 65535]; *not_integral otherwise (synchronises)
int launch_f32_split_image(mtm_ctx* ctx, int* not_integral);
int launch_cc16_epilogue(mtm_ctx* ctx, int method, int tmpl);
// augmentation / area downscale (transform.cu)
int launch_transform(mtm_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const XformDesc* d_descs, int n_out,
                     int64_t max_pixels, int C, int dtype, int factor);
// raw (unsorted) peaks of every template -> block A
int launch_peaks(mtm_ctx* ctx, int method, int64_t n_object, float thr32, double thr64, bool allow_candidates = true);
// in-place sort of block A (mode 0: findMatches order, mode 1: NMSBoxes order)
int launch_sort_hits(mtm_ctx* ctx, int mode, int minimize, int ascending_key, int check_trivial, bool prepped = false);
// fast path (raw count <= 1024): sort(s) [+ NMS] in one launch; sets header[2] = 1 when it declines
constexpr int MTM_MIRROR_HITS = 256;
int launch_finalize_small(mtm_ctx* ctx, int minimize, int check_trivial, int presorted, int do_nms, float thr32,
                          int ascending, int64_t n_object, float max_overlap, uint8_t* out_block = nullptr, bool mirror = false,
                          bool prepped = false);
// block A (sorted mode 1) -> block B
int launch_nms(mtm_ctx* ctx, float thr32, int ascending, int64_t n_object, float max_overlap);

// ---- shared device helpers ----------------------------------------------------
__host__ __device__ inline bool method_is_min(int method) { return method == 0 || method == 1; }

// mode-0 sort keys are cached in the hit itself so that comparisons never touch global memory:
// seq = row-major index in the score map, key = 1 for 1-D (find_peaks) maps.
__device__ __forceinline__ void prep_mode0(DevHit& h, const TmplMeta* __restrict__ meta)
{
    const TmplMeta& tm = meta[h.tmpl];
    h.seq = h.y * tm.mw + h.x;
    h.key = (tm.mh == 1 || tm.mw == 1) ? 1.0f : 0.0f;
}

// 1 / sqrt(x) for NORMAL positive x (callers pass integers >= 1 converted to float): the bare MUFU.RSQ, without rsqrtf()'s
// subnormal scaling -- same bits as rsqrtf() on such inputs.
__device__ __forceinline__ float mtm_rsqrt_normal(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return rsqrtf(x);
#endif
}

__device__ __forceinline__ uint32_t ordered_f32(float f) {
    f += 0.0f;                                    // -0 -> +0
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
