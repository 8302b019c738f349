/* mtm_b200.h -- C ABI of libmtm_b200.so, the B200-native replacement for the
 * third-party kernels on the hot path of MTM 2.0.1 (MultiTemplateMatching-Python).
 *
 * Plain C, no torch / numpy types.  Every entry point names the reference call
 * site (file:line under the reference tree) it stands in for.  All functions
 * return MTM_OK (0) or a negative mtm_status; a human-readable message for the
 * last failure on a context is available through mtm_last_error().  There is no
 * CPU fallback: without a CUDA device mtm_create() fails with MTM_ERR_CUDA.
 *
 * Threading: a context owns one CUDA stream and its device workspaces; calls on
 * the same context must not overlap (the Python layer holds a lock).  Use one
 * context per GPU / per concurrent caller.
 */
#ifndef MTM_B200_H
#define MTM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTM_ABI_VERSION 2

typedef enum mtm_status {
    MTM_OK = 0,
    MTM_ERR_INVALID = -1,      /* bad argument (shape, dtype, null pointer, state)          */
    MTM_ERR_CUDA = -2,         /* CUDA runtime / driver failure (message has the details)   */
    MTM_ERR_CAPACITY = -3,     /* caller's hit buffer too small; *n_hits holds the need     */
    MTM_ERR_UNSUPPORTED = -4,  /* valid for the reference, not implemented on the GPU yet   */
    MTM_ERR_PEER = -5          /* another rank of the communicator failed inside a collective call */
} mtm_status;

/* MTM_U16: 16-bit unsigned pixels (image AND templates).  The reference casts them to float32 before
 * cv2.matchTemplate (MTM/__init__.py:71-74); here the same float32 semantics are kept for the window statistics and the
 * OpenCV epilogue, while the numerator is computed EXACTLY on the tensor cores from the high/low byte planes
 * (single-channel images; other shapes take the float32 kernels).
 * MTM_F32: what the reference's float32 cast produces.  Single-channel float32 images / templates whose pixels are all
 * integers in [0, 65535] (16-bit data cast by the caller; a uint16 image with float32 templates) are detected on upload and
 * take the same exact byte-plane route (MTM_B200_F32_EXACT=0 turns the detection off); everything else runs the fp32 kernels. */
typedef enum mtm_dtype { MTM_U8 = 0, MTM_F32 = 1, MTM_U16 = 2 } mtm_dtype;

/* cv2.TM_* codes, MTM/__init__.py:56 `method` */
typedef enum mtm_method {
    MTM_TM_SQDIFF = 0, MTM_TM_SQDIFF_NORMED = 1, MTM_TM_CCORR = 2,
    MTM_TM_CCORR_NORMED = 3, MTM_TM_CCOEFF = 4, MTM_TM_CCOEFF_NORMED = 5
} mtm_method;

/* Numerator kernel selection (mtm_set_option MTM_OPT_PATH). */
typedef enum mtm_path { MTM_PATH_AUTO = 0, MTM_PATH_DIRECT = 1, MTM_PATH_TENSOR = 2 } mtm_path;
/* The 8 symmetries of the square applied to a template on the device (mtm_set_templates_transformed): what the
 * tutorials of the reference do with numpy before calling matchTemplates (Tutorial2-Template_Augmentation.ipynb
 * cell 15: np.rot90(template, k), "We could also do some flipping with np.fliplr, flipud"). */
typedef enum mtm_transform {
    MTM_XF_IDENTITY = 0,
    MTM_XF_ROT90 = 1,          /* np.rot90(t, 1): counter-clockwise, (h, w) -> (w, h) */
    MTM_XF_ROT180 = 2,         /* np.rot90(t, 2) */
    MTM_XF_ROT270 = 3,         /* np.rot90(t, 3) */
    MTM_XF_FLIPLR = 4,         /* np.fliplr(t) */
    MTM_XF_FLIPUD = 5,         /* np.flipud(t) */
    MTM_XF_TRANSPOSE = 6,      /* t.swapaxes(0, 1) */
    MTM_XF_ANTITRANSPOSE = 7   /* np.rot90(t, 2).swapaxes(0, 1) */
} mtm_transform;
#define MTM_MAX_DOWNSCALE 16

typedef enum mtm_option {
    MTM_OPT_PATH = 0,          /* mtm_path */
    MTM_OPT_TIME_NCC = 1       /* 1: bracket the numerator kernels with CUDA events (roofline timing) */
} mtm_option;

/* Device/host mirror of the reference's Hit tuple (label, (x, y, w, h), score),
 * MTM/NMS.py:18; `tmpl` indexes the template list instead of carrying the label. */
typedef struct mtm_hit {
    int32_t tmpl;
    int32_t x, y, w, h;
    float score;
} mtm_hit;

typedef struct mtm_counters {
    int64_t kernel_launches;   /* kernels of this library launched since the last reset */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t ncc_launches;      /* numerator (K2/K3) kernel launches timed with MTM_OPT_TIME_NCC   */
    double ncc_ms;             /* their summed device time in ms (CUDA events on the ctx stream) */
    int64_t tma_launches;      /* numerator launches whose image tiles went through the TMA unit (cp.async.bulk.tensor) */
    int64_t hits_only_searches;/* searches that wrote no score map (candidate list / arg-max straight from the epilogue)  */
} mtm_counters;

typedef struct mtm_ctx mtm_ctx;

/* ---- lifetime ------------------------------------------------------------- */
int mtm_abi_version(void);
/* Number of usable (sm_100) CUDA devices visible to the process, 0 without a GPU / driver (never fails). */
int mtm_device_count(void);
int mtm_create(int device, mtm_ctx** out);
int mtm_destroy(mtm_ctx* ctx);
/* ctx may be NULL: returns the message of the last failed mtm_create(). */
const char* mtm_last_error(const mtm_ctx* ctx);
/* Run on an existing CUDA stream (cudaStream_t passed as void*), e.g. the
 * caller's timing stream; NULL restores the context's own stream. */
int mtm_set_stream(mtm_ctx* ctx, void* cuda_stream);
int mtm_synchronize(mtm_ctx* ctx);
int mtm_set_option(mtm_ctx* ctx, int option, int64_t value);
int mtm_get_counters(mtm_ctx* ctx, mtm_counters* out);
int mtm_reset_counters(mtm_ctx* ctx);
/* CUDA-event stopwatch on the context's stream (bench.py): begin, work, end -> ms. */
int mtm_timer_begin(mtm_ctx* ctx);
int mtm_timer_end(mtm_ctx* ctx, float* elapsed_ms);
/* Measurement helper for the roofline bench.py reports: the dense rate of the tensor pipe the numerator kernels use.
 * One CTA per SM issues `iters` back-to-back tcgen05.mma kind::i8 (u8 x u8 -> s32, M 128 x N n_cols x K 32) from
 * operands resident in shared memory -- no loads, no epilogue; best of 5 launches timed with CUDA events.
 * *tmacs_per_s = 1e-12 * MACs / s (x2 = TOP/s). */
int mtm_measure_i8_peak(mtm_ctx* ctx, int n_cols, int iters, double* tmacs_per_s);

/* ---- inputs ---------------------------------------------------------------
 * mtm_set_image: the `image` operand of cv2.matchTemplate (MTM/__init__.py:92)
 * after MTM's searchBox crop (MTM/__init__.py:140-144).  `pixels` is a host
 * pointer to H rows of W*C elements, rows `row_stride_bytes` apart (a cropped
 * numpy view needs no host copy).  Uploads and builds the per-channel window
 * statistics (what OpenCV's integral() provides to common_matchTemplate).
 * mtm_set_image_device: same with a DEVICE pointer (inputs already in HBM). */
int mtm_set_image(mtm_ctx* ctx, const void* pixels, int H, int W, int C, int dtype,
                  int64_t row_stride_bytes);
int mtm_set_image_device(mtm_ctx* ctx, const void* d_pixels, int H, int W, int C, int dtype,
                         int64_t row_stride_bytes);
/* mtm_set_templates: the `listTemplates` arrays (MTM/__init__.py:95,147-167),
 * each contiguous h[i] x w[i] x C, same dtype/C as the image. */
int mtm_set_templates(mtm_ctx* ctx, int n, const void* const* pixels,
                      const int32_t* h, const int32_t* w, int C, int dtype);

/* mtm_set_templates_masked: templates WITH masks, the `mask=` operand of cv2.matchTemplate
 * (MTM/__init__.py:76-92, only honoured for TM_SQDIFF and TM_CCORR_NORMED).  masks[i] has the shape and
 * dtype of pixels[i]; uint8 masks are binary (non-zero = use), float32 masks are weights (OpenCV). */
int mtm_set_templates_masked(mtm_ctx* ctx, int n, const void* const* pixels, const void* const* masks,
                             const int32_t* h, const int32_t* w, int C, int dtype);

/* ---- on-device augmentation and search-region / pyramid helpers ----------------------------
 * The reference leaves these to user code in its tutorials; here they run on the device so that only the base
 * pixels cross PCIe once.
 *
 * mtm_set_templates_transformed: the template list becomes, for every base template b (in order) and every
 * ops[k] (in order), transform ops[k] of the `downscale`-times reduced base: index b*n_ops + k.  The reduction is
 * cv2.resize(t[:h/f*f, :w/f*f], (w/f, h/f), interpolation=cv2.INTER_AREA) (Tutorial3-SpeedingUp.ipynb cells
 * 17-19; bit-identical for integer pixels), downscale == 1 keeps the pixels.  MTM_U8 and MTM_F32 templates. */
int mtm_set_templates_transformed(mtm_ctx* ctx, int n, const void* const* pixels, const int32_t* h, const int32_t* w,
                                  int C, int dtype, int n_ops, const int32_t* ops, int downscale);

/* mtm_set_image_scaled: uploads the full-resolution image (kept resident for mtm_set_image_roi) and searches its
 * `downscale`-times reduced copy: cv2.resize(image[:H/f*f, :W/f*f], (W/f, H/f), interpolation=cv2.INTER_AREA),
 * Tutorial3-SpeedingUp.ipynb cell 17.  downscale in [1, MTM_MAX_DOWNSCALE]. */
int mtm_set_image_scaled(mtm_ctx* ctx, const void* pixels, int H, int W, int C, int dtype, int64_t row_stride_bytes,
                         int downscale);

/* mtm_set_image_roi: the searchBox crop of MTM.findMatches (MTM/__init__.py:140-144) taken from the resident
 * full-resolution image of the last mtm_set_image_scaled: no host copy, no upload.  Hit coordinates of the
 * following searches are relative to (x, y), as with a host-side crop. */
int mtm_set_image_roi(mtm_ctx* ctx, int x, int y, int w, int h);

/* ---- the hot path ---------------------------------------------------------
 * mtm_score_map: cv2.matchTemplate(image, template, method) at
 * MTM/__init__.py:92 for template `tmpl`; writes (H-h+1)*(W-w+1) floats. */
int mtm_score_map(mtm_ctx* ctx, int tmpl, int method, float* out_host, int64_t out_elems);

/* mtm_find_matches: the per-template loop of MTM.findMatches
 * (MTM/__init__.py:172-175 -> _multi_compute :179-244): score map, then
 * cv2.minMaxLoc (n_object == 1, :225-230) or _findLocalMax_/_findLocalMin_
 * (:231-235 -> skimage peak_local_max :45, scipy find_peaks :34/:40, 1x1 map
 * :25-30).  n_object < 0 means inf.  Hits come back in template order, each
 * template's peaks in the reference's order.  Offsets are NOT added. */
int mtm_find_matches(mtm_ctx* ctx, int method, int64_t n_object, double score_threshold,
                     mtm_hit* hits, int capacity, int* n_hits);

/* mtm_nms: MTM.NMS (MTM/NMS.py:20-84) including cv2.dnn.NMSBoxes (:78).
 * keep[] receives indices into hits[] in output order. */
int mtm_nms(mtm_ctx* ctx, const mtm_hit* hits, int n, double score_threshold,
            int sort_ascending, int64_t n_object, double max_overlap,
            int32_t* keep, int* n_keep);

/* mtm_match_templates: MTM.matchTemplates after validation
 * (MTM/__init__.py:289-296): find + NMS fused on the device, one D2H copy. */
int mtm_match_templates(mtm_ctx* ctx, int method, int64_t n_object, double score_threshold,
                        double max_overlap, mtm_hit* hits, int capacity, int* n_hits);

/* Pipelined form of mtm_match_templates for streams of images (the caller's loop over images,
 * e.g. the 16-image batch of BASELINE.json configs[4]): _async enqueues the whole search + NMS of
 * the CURRENT image/templates and the copy of its result into pinned memory, and returns without
 * waiting; _collect waits for that slot and hands the hits over.  Up to MTM_MAX_INFLIGHT
 * submissions may be in flight, so the GPU runs back to back while the host prepares the next
 * image.  Limited to the fused fast path: a submission with more than 1024 raw peaks makes
 * _collect return MTM_ERR_CAPACITY (re-run that image with mtm_match_templates). */
#define MTM_MAX_INFLIGHT 8
int mtm_match_templates_async(mtm_ctx* ctx, int method, int64_t n_object, double score_threshold,
                              double max_overlap, int slot);
int mtm_match_templates_collect(mtm_ctx* ctx, int slot, mtm_hit* hits, int capacity, int* n_hits);

/* ---- multi-GPU: the two cuts of the reference's parallel axis (SURVEY 8 b/e) ------------------------------
 * The reference runs one task per template (MTM/__init__.py:172-175) and couples them again only in the NMS
 * (MTM/__init__.py:294-296).  Across GPUs the same cut needs ONE exchange: an all-gather of the ranks' hit
 * blocks (32-byte count header + 32-byte records) over NCCL / NVLink, issued by the library on the context's
 * stream between the peak kernels and the replicated NMS kernel -- no host round trip, no torch.
 *
 * A communicator (mtm_comm) is one rank's endpoint:
 *   mtm_comm_init_rank  one process per GPU (torchrun): rank 0 makes an id with mtm_comm_unique_id, the launcher
 *                       hands its MTM_COMM_ID_BYTES bytes to every rank (file, socket, environment ...);
 *   mtm_comm_create     one process driving n GPUs: out[0..n) receive n endpoints (ncclCommInitAll); each must
 *                       then be driven by its own host thread.  When every entry of `devices` names the SAME
 *                       GPU the group runs on an in-process loop-back (device-to-device copies ordered by CUDA
 *                       events instead of NCCL), which lets a single-GPU box exercise the whole exchange.
 * libnccl.so.2 is resolved at run time (dlopen; override with MTM_B200_NCCL_LIB): a single-GPU user of the
 * library needs no NCCL.  world == 1 never touches NCCL. */
typedef struct mtm_comm mtm_comm;
#define MTM_COMM_ID_BYTES 128
int mtm_comm_unique_id(void* id_out);
int mtm_comm_init_rank(int device, int world, int rank, const void* id, mtm_comm** out);
int mtm_comm_create(int n, const int* devices, mtm_comm** out);
int mtm_comm_destroy(mtm_comm* comm);
int mtm_comm_info(const mtm_comm* comm, int* world, int* rank, int* device);
/* comm may be NULL: message of the last failed mtm_comm_init_rank / mtm_comm_create / mtm_comm_unique_id. */
const char* mtm_comm_last_error(const mtm_comm* comm);
/* Element-wise MAX of `n` doubles over the ranks (device-timed durations: "max over ranks"), and a barrier. */
int mtm_comm_allreduce_max(mtm_comm* comm, double* values, int n);
int mtm_comm_barrier(mtm_comm* comm);

/* mtm_match_templates_sharded: MTM.matchTemplates (MTM/__init__.py:247-296) with the TEMPLATE LIST cut into
 * contiguous slices, one per rank (BASELINE configs[3]).  Every rank holds the whole image and its own slice of
 * `n_local` templates in `ctx` (mtm_set_templates with the slice; n_local == 0: this rank contributes nothing and
 * the templates resident in ctx are ignored); `tmpl_base` is the list index of the slice's first template.
 * Local stage: score maps -> peaks of the slice (the per-template tasks of :172-175).  Exchange: one all-gather of
 * the hit blocks; the rank-ordered concatenation is the list findMatches would have returned.  Then every rank
 * runs the identical global NMS (:294-296) on its GPU and returns the identical list; hits[].tmpl indexes the
 * WHOLE template list.  A rank whose local stage fails still takes part in the exchange and every rank returns
 * an error (the failing one its own, the others MTM_ERR_PEER): nobody is left waiting in the collective. */
int mtm_match_templates_sharded(mtm_ctx* ctx, mtm_comm* comm, int tmpl_base, int n_local, int method, int64_t n_object,
                                double score_threshold, double max_overlap, mtm_hit* hits, int capacity, int* n_hits);

/* mtm_gather_results: the IMAGE cut (BASELINE configs[4]: a batch of images, NMS local to an image).  Rank r has
 * submitted its n_local <= images_per_rank images with mtm_match_templates_async: submission i sits in slot
 * slots[i] of context ctxs[i].  One all-gather of images_per_rank blocks per rank (header + hits_per_image
 * records) hands every image's final hit list to every rank: out_hits[(r * images_per_rank + i) * hits_per_image
 * + k], out_counts[r * images_per_rank + i] = number of hits, -1 for an unused entry, -2 when that image did not
 * fit (more than hits_per_image hits, or outside the fused fast path): re-run it with mtm_match_templates.
 * The gathered slots are released (no mtm_match_templates_collect afterwards). */
int mtm_gather_results(mtm_comm* comm, int n_local, mtm_ctx* const* ctxs, const int* slots, int images_per_rank,
                       int hits_per_image, mtm_hit* out_hits, int32_t* out_counts);

#ifdef __cplusplus
}
#endif
#endif /* MTM_B200_H */
