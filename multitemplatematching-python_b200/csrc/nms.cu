// nms.cu -- K7: hit ordering and greedy IoU suppression on the device.
//
//  * sort_hits_kernel : bitonic sort of the hit buffer (single CTA, no host sync;
//      the count is read from device memory).
//      mode 0 = the order MTM.findMatches produces when templates are visited in
//               list order (MTM/__init__.py:172-175,238-244): template index, then
//               the peak finder's order (descending score with row-major ties for
//               peak_local_max; ascending index for the 1-D find_peaks maps);
//      mode 1 = cv2.dnn.NMSBoxes' std::stable_sort by descending score
//               (GetMaxScoreIndex), ties keep list order (seq).
//  * nms_kernel : MTM.NMS (MTM/NMS.py:20-84): the nHits<=1 and N_object==1
//      short-cuts, the 'score > threshold' filter, NMSFast_'s greedy loop with
//      rectOverlap = 1.f - (float)jaccardDistance(Rect_<int>) and the [:N_object] cut.
#include "mtm_internal.cuh"
#include <math_constants.h>

namespace {

struct SortCtx {
    const TmplMeta* meta;
    int mode;
    int minimize;
};

__device__ __forceinline__ bool hit_less(const DevHit& a, const DevHit& b, const SortCtx& s)
{
    if (s.mode == 0) {
        if (a.tmpl != b.tmpl) return a.tmpl < b.tmpl;
        if (a.tmpl == 0x7fffffff) return false;
        // prepared by prep_mode0(): seq = row-major index in the map, key = 1 for 1-D (find_peaks) maps
        if (a.key == 0.0f) {
            const float ka = s.minimize ? -a.score : a.score, kb = s.minimize ? -b.score : b.score;
            if (ka != kb) return ka > kb;
        }
        return a.seq < b.seq;
    }
    if (a.key != b.key) return a.key > b.key;
    if (s.mode == 2) {                                              // canonical list order without sorting into it first
        if (a.tmpl != b.tmpl) return a.tmpl < b.tmpl;
        // minimising methods: two different scores can share the float32 key 1 - score.  The stable sort of the reference keeps
        // the peak finder's order there (2-D maps: ascending score, then row-major) -- what mode 0 followed by mode 1 produces.
        if (s.minimize && a.score != b.score && a.tmpl != 0x7fffffff) {
            const TmplMeta& tm = s.meta[a.tmpl];
            if (tm.mh != 1 && tm.mw != 1) return a.score < b.score;
        }
    }
    return a.seq < b.seq;
}

__device__ __forceinline__ DevHit load_hit(const DevHit* p)
{
    DevHit h;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    *reinterpret_cast<uint4*>(&h) = a;
    *(reinterpret_cast<uint4*>(&h) + 1) = b;
    return h;
}
__device__ __forceinline__ void store_hit(DevHit* p, const DevHit& h)
{
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = *reinterpret_cast<const uint4*>(&h);
    q[1] = *(reinterpret_cast<const uint4*>(&h) + 1);
}

// count[0] = number of hits.  mode 0 drops the hits of "trivial" templates (constant map
// rule of peak_local_max) and, when assign_seq, numbers the survivors.
__global__ void __launch_bounds__(1024, 1)
sort_hits_kernel(DevHit* __restrict__ hits, int cap, int32_t* __restrict__ count, const TmplMeta* __restrict__ meta,
                 const int32_t* __restrict__ nontrivial, int mode, int minimize, int ascending_key, int check_trivial,
                 int prepped)
{
    __shared__ int dead;
    const int tid = threadIdx.x, nth = blockDim.x;
    int n = count[0];
    if (n > cap) n = cap;                       // overflow is reported by the host from count[0]
    if (tid == 0) dead = 0;
    __syncthreads();
    int npad = 1;
    while (npad < n) npad <<= 1;
    // pre-pass
    for (int i = tid; i < npad; i += nth) {
        if (i >= n) {
            DevHit s;
            s.tmpl = 0x7fffffff; s.x = s.y = s.w = s.h = 0; s.score = 0.f; s.seq = 0x7fffffff; s.key = -CUDART_INF_F;
            store_hit(hits + i, s);
        } else if (mode == 0) {
            if (check_trivial && !nontrivial[hits[i].tmpl]) {
                hits[i].tmpl = 0x7fffffff;
                hits[i].key = -CUDART_INF_F;
                hits[i].seq = 0x7fffffff;
                atomicAdd(&dead, 1);
            } else if (!prepped) {              // gathered hits (mtm_match_templates_sharded) carry their keys already
                DevHit h = load_hit(hits + i);
                prep_mode0(h, meta);
                store_hit(hits + i, h);
            }
        } else {
            hits[i].key = ascending_key ? 1.0f - hits[i].score : hits[i].score;
        }
    }
    __syncthreads();
    SortCtx sc{meta, mode, minimize};
#pragma unroll 1
    for (int k = 2; k <= npad; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 1
            for (int i = tid; i < npad; i += nth) {
                const int l = i ^ j;
                if (l > i) {
                    DevHit a = load_hit(hits + i), b = load_hit(hits + l);
                    const bool up = ((i & k) == 0);
                    const bool swap = up ? hit_less(b, a, sc) : hit_less(a, b, sc);
                    if (swap) { store_hit(hits + i, b); store_hit(hits + l, a); }
                }
            }
            __syncthreads();
        }
    }
    if (mode == 0) {
        const int live = n - dead;
        for (int i = tid; i < live; i += nth) hits[i].seq = i;
        if (tid == 0) { count[1] = live; if (count[0] <= cap) count[0] = live; }
    }
}

__device__ __forceinline__ float rect_overlap(const DevHit& a, const DevHit& b)
{
    const int Aa = a.w * a.h, Ab = b.w * b.h;
    if (Aa + Ab <= 0) return 1.0f;                       // jaccardDistance() == 0
    const int x1 = max(a.x, b.x), y1 = max(a.y, b.y);
    const int iw = min(a.x + a.w, b.x + b.w) - x1, ih = min(a.y + a.h, b.y + b.h) - y1;
    const double Aab = (iw > 0 && ih > 0) ? (double)(iw * ih) : 0.0;
    const double dist = 1.0 - Aab / ((double)(Aa + Ab) - Aab);
    return 1.0f - (float)dist;
}

// hits sorted with mode 1.  out[] receives the kept hits in output order, out_count[0] their number.
__global__ void __launch_bounds__(1024, 1)
nms_kernel(const DevHit* __restrict__ hits, int cap, const int32_t* __restrict__ count, DevHit* __restrict__ out,
           int32_t* __restrict__ out_count, int32_t* __restrict__ keep, float thr32, int ascending,
           long long n_object, float max_overlap)
{
    __shared__ int s_kept;
    __shared__ unsigned long long s_best;
    const int tid = threadIdx.x, nth = blockDim.x;
    int n = count[0];
    if (tid == 0) out_count[1] = n;                       // raw count: lets the host see an overflow
    if (n > cap) n = cap;
    if (n <= 1) {                                         // MTM/NMS.py:53-55: no thresholding
        if (tid == 0) { if (n == 1) { out[0] = hits[0]; keep[0] = 0; } out_count[0] = n; }
        return;
    }
    if (n_object == 1) {                                  // MTM/NMS.py:61-69: best raw score, first in list order
        if (tid == 0) s_best = 0ull;
        __syncthreads();
        unsigned long long k = 0ull;
        for (int i = tid; i < n; i += nth) {
            const float v = ascending ? -hits[i].score : hits[i].score;
            const unsigned long long key = ((unsigned long long)ordered_f32(v) << 32) |
                                           (unsigned long long)(0xFFFFFFFFu - (uint32_t)hits[i].seq);
            k = key > k ? key : k;
        }
        atomicMax(&s_best, k);
        __syncthreads();
        for (int i = tid; i < n; i += nth) {
            const float v = ascending ? -hits[i].score : hits[i].score;
            const unsigned long long key = ((unsigned long long)ordered_f32(v) << 32) |
                                           (unsigned long long)(0xFFFFFFFFu - (uint32_t)hits[i].seq);
            if (key == s_best) { out[0] = hits[i]; keep[0] = i; out_count[0] = 1; }
        }
        return;
    }
    if (tid == 0) s_kept = 0;
    __syncthreads();
    const long long limit = n_object < 0 ? (long long)n : n_object;
    for (int i = 0; i < n; ++i) {
        const DevHit cand = hits[i];
        if (!(cand.key > thr32)) break;                   // sorted by key: the rest fails too
        const int kept = s_kept;
        if (kept >= limit) break;
        int sup = 0;
        for (int k = tid; k < kept; k += nth) {
            if (!(rect_overlap(cand, hits[keep[k]]) <= max_overlap)) sup = 1;
        }
        sup = __syncthreads_or(sup);
        if (!sup && tid == 0) { keep[kept] = i; out[kept] = cand; s_kept = kept + 1; }
        __syncthreads();
    }
    if (tid == 0) out_count[0] = (s_kept < limit) ? s_kept : (int)limit;
}

// ---- fast path: the whole post-peak pipeline in ONE launch when the raw hit list is small ----
// (the common case: tens of hits).  Shared-memory bitonic sorts + the greedy scan; falls through
// (out_count[2] = 1) to the general multi-kernel path when there are more than FIN_CAP raw hits.
constexpr int FIN_CAP = 1024;

__device__ void smem_bitonic(DevHit* sh, int npad, const SortCtx& sc)
{
    const int tid = threadIdx.x, nth = blockDim.x;
#pragma unroll 1
    for (int k = 2; k <= npad; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 1
            for (int i = tid; i < npad; i += nth) {
                const int l = i ^ j;
                if (l > i) {
                    const DevHit a = sh[i], b = sh[l];
                    const bool up = ((i & k) == 0);
                    const bool swap = up ? hit_less(b, a, sc) : hit_less(a, b, sc);
                    if (swap) { sh[i] = b; sh[l] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// do_nms == 0: findMatches order -> written back to `hits` (block A), count[0]/[1] updated.
// do_nms == 1: ... then MTM.NMS -> `out` (block B), out_count[0] = kept, out_count[1] = raw count.
// PRESORTED / DO_NMS are compile-time: the launch-latency-bound single CTA then walks ~8 KB of code instead of 32 KB
// (instruction fetch of a cold kernel is a visible part of its ~15 us).
template <bool PRESORTED, bool DO_NMS>
__device__ __forceinline__ void
finalize_small_body(DevHit* __restrict__ hits, int cap, int32_t* __restrict__ count, const TmplMeta* __restrict__ meta,
                    const int32_t* __restrict__ nontrivial, int minimize, int check_trivial,
                    DevHit* __restrict__ out, int32_t* __restrict__ out_count, float thr32, int ascending,
                    long long n_object, float max_overlap, int prepped)
{
    constexpr bool presorted = PRESORTED, do_nms = DO_NMS;
    __shared__ DevHit sh[FIN_CAP];
    __shared__ int s_live, s_kept;
    __shared__ unsigned long long s_best;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int n_raw = count[0];
    int32_t* flag_hdr = do_nms ? out_count : count;
    // sharded calls: the merge kernel left the largest per-rank raw count in [4] and a failing rank's status in [5]
    if (do_nms && tid < 4) out_count[4 + tid] = count[4 + tid];            // [4] largest per-rank count, [5] status, [6] largest raw count
    if (count[3]) { if (tid == 0) { flag_hdr[2] = 2; if (do_nms) out_count[1] = 0; } return; }   // candidate list overflowed
    if (n_raw > FIN_CAP) { if (tid == 0) { flag_hdr[2] = 1; if (do_nms) out_count[1] = n_raw; } return; }
    if (tid == 0) { s_live = 0; s_kept = 0; s_best = 0ull; flag_hdr[2] = 0; }
    __syncthreads();
    int npad = 1;
    while (npad < n_raw) npad <<= 1;
    int dead_local = 0;
#pragma unroll 1
    for (int i = tid; i < npad; i += nth) {
        DevHit h;
        if (i < n_raw) {
            h = load_hit(hits + i);
            if (check_trivial && !nontrivial[h.tmpl]) { h.tmpl = 0x7fffffff; h.key = -CUDART_INF_F; h.seq = 0x7fffffff; dead_local++; }
            else if (!presorted && !prepped) prep_mode0(h, meta);
        } else {
            h.tmpl = 0x7fffffff; h.x = h.y = h.w = h.h = 0; h.score = 0.f; h.seq = 0x7fffffff; h.key = -CUDART_INF_F;
        }
        sh[i] = h;
    }
    if (dead_local) atomicAdd(&s_live, dead_local);
    __syncthreads();
    const int n = n_raw - s_live;                               // live hits (s_live counted the dead ones)
    __syncthreads();
    if (!presorted && !do_nms) {                                // findMatches order
        SortCtx sc0{meta, 0, minimize};
        smem_bitonic(sh, npad, sc0);
#pragma unroll 1
        for (int i = tid; i < n; i += nth) store_hit(hits + i, sh[i]);
#pragma unroll 1
        for (int i = tid; i < n; i += nth) hits[i].seq = i;
        if (tid == 0) { count[0] = n; count[1] = n; }
        return;
    }
    if (tid == 0) out_count[1] = n_raw;
    // ---- MTM.NMS (MTM/NMS.py:20-84) ----
    // NMSBoxes' stable sort by descending key of the canonical (template, peak order) list ==
    // ONE sort by (key desc, template asc, row-major index asc): mode 2.  Presorted input
    // (standalone mtm_nms, N_object == 1) keeps its own seq: mode 1.
    if (!presorted) {
#pragma unroll 1
        for (int i = tid; i < npad; i += nth)
            if (sh[i].tmpl != 0x7fffffff) sh[i].key = ascending ? 1.0f - sh[i].score : sh[i].score;   // seq = row-major index (prep_mode0)
        __syncthreads();
        SortCtx sc2{meta, 2, minimize};
        smem_bitonic(sh, npad, sc2);
    }
    if (n <= 1) {
        if (tid == 0) { if (n == 1) out[0] = sh[0]; out_count[0] = n; }
        return;
    }
    if (n_object == 1) {
        unsigned long long kbest = 0ull;
#pragma unroll 1
        for (int i = tid; i < n; i += nth) {
            const float v = ascending ? -sh[i].score : sh[i].score;
            const unsigned long long key = ((unsigned long long)ordered_f32(v) << 32) |
                                           (unsigned long long)(0xFFFFFFFFu - (uint32_t)sh[i].seq);
            kbest = key > kbest ? key : kbest;
        }
        atomicMax(&s_best, kbest);
        __syncthreads();
#pragma unroll 1
        for (int i = tid; i < n; i += nth) {
            const float v = ascending ? -sh[i].score : sh[i].score;
            const unsigned long long key = ((unsigned long long)ordered_f32(v) << 32) |
                                           (unsigned long long)(0xFFFFFFFFu - (uint32_t)sh[i].seq);
            if (key == s_best) { out[0] = sh[i]; out_count[0] = 1; }
        }
        return;
    }
    if (presorted) {
#pragma unroll 1
        for (int i = tid; i < npad; i += nth)
            if (i < n) sh[i].key = ascending ? 1.0f - sh[i].score : sh[i].score;
        __syncthreads();
        SortCtx sc1{meta, 1, minimize};
        smem_bitonic(sh, npad, sc1);
    }
    // greedy scan by warp 0 alone (warp votes instead of block barriers); kept_idx lists the survivors
    __shared__ unsigned short kept_idx[FIN_CAP];
    const long long limit = n_object < 0 ? (long long)n : n_object;
    if (tid < 32) {
        int kept = 0;
#pragma unroll 1
        for (int i = 0; i < n && kept < limit; ++i) {
            const DevHit cand = sh[i];
            if (!(cand.key > thr32)) break;
            int sup = 0;
#pragma unroll 1
            for (int k = tid; k < kept; k += 32) {
                const DevHit& o = sh[kept_idx[k]];
                // disjoint boxes have overlap 1.f - (float)1.0 == 0 <= max_overlap: skip the fp64 division
                const bool meet = (cand.x < o.x + o.w && o.x < cand.x + cand.w && cand.y < o.y + o.h && o.y < cand.y + cand.h) ||
                                  (cand.w * cand.h + o.w * o.h <= 0);          // degenerate rects: keep OpenCV's rule
                if (meet && !(rect_overlap(cand, o) <= max_overlap)) sup = 1;
            }
            if (!__any_sync(0xffffffffu, sup)) {
                if (tid == 0) kept_idx[kept] = (unsigned short)i;
                ++kept;
                __syncwarp();
            }
        }
        if (tid == 0) s_kept = kept;
    }
    __syncthreads();
    const int kept = s_kept;
#pragma unroll 1
    for (int k = tid; k < kept; k += nth) store_hit(out + k, sh[kept_idx[k]]);
    if (tid == 0) out_count[0] = kept;
}

// `mirror` (optional): mapped pinned host memory.  The header and the first MTM_MIRROR_HITS hits of the result are
// stored there as well, so the synchronous API reads them after a stream synchronise instead of a D2H copy + synchronise.
template <bool PRESORTED, bool DO_NMS>
__global__ void __launch_bounds__(256, 1)
finalize_small_kernel(DevHit* __restrict__ hits, int cap, int32_t* __restrict__ count, const TmplMeta* __restrict__ meta,
                      const int32_t* __restrict__ nontrivial, int minimize, int check_trivial,
                      DevHit* __restrict__ out, int32_t* __restrict__ out_count, float thr32, int ascending,
                      long long n_object, float max_overlap, uint8_t* __restrict__ mirror, int prepped)
{
    constexpr bool do_nms = DO_NMS;
    finalize_small_body<PRESORTED, DO_NMS>(hits, cap, count, meta, nontrivial, minimize, check_trivial, out, out_count, thr32,
                                           ascending, n_object, max_overlap, prepped);
    if (!mirror) return;
    __syncthreads();                                           // the block's own global writes are visible to all its threads
    const int32_t* hdr = do_nms ? out_count : count;
    const uint4* src = reinterpret_cast<const uint4*>(do_nms ? out : hits);
    const int n = min(max(hdr[0], 0), MTM_MIRROR_HITS);
    uint4* dst = reinterpret_cast<uint4*>(mirror);
    const int tid = threadIdx.x;
    if (tid < 2) dst[tid] = reinterpret_cast<const uint4*>(hdr)[tid];          // 32-byte header
    for (int i = tid; i < 2 * n; i += blockDim.x) dst[2 + i] = src[i];         // 32-byte hits (the host reads them after a stream synchronise: no fence needed)
}

}  // namespace

int launch_finalize_small(mtm_ctx* ctx, int minimize, int check_trivial, int presorted, int do_nms, float thr32,
                          int ascending, int64_t n_object, float max_overlap, uint8_t* out_block, bool mirror, bool prepped)
{
    uint8_t* ob = out_block ? out_block : ctx->d_blockB;
    static_assert(MTM_HIT_HEADER == 32 && sizeof(DevHit) == 32, "mirror layout");
    DevHit* oh = reinterpret_cast<DevHit*>(ob + MTM_HIT_HEADER);
    int32_t* oc = reinterpret_cast<int32_t*>(ob);
    uint8_t* mr = mirror ? ctx->d_mirror : nullptr;
#define MTM_FIN(P, N) finalize_small_kernel<P, N><<<1, 256, 0, ctx->stream>>>(ctx->hitsA(), ctx->hit_cap, ctx->countA(), ctx->d_meta, \
        ctx->d_nontrivial, minimize, check_trivial, oh, oc, thr32, ascending, (long long)n_object, max_overlap, mr, prepped ? 1 : 0)
    if (presorted) { if (do_nms) MTM_FIN(true, true); else MTM_FIN(true, false); }
    else { if (do_nms) MTM_FIN(false, true); else MTM_FIN(false, false); }
#undef MTM_FIN
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_sort_hits(mtm_ctx* ctx, int mode, int minimize, int ascending_key, int check_trivial, bool prepped)
{
    sort_hits_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->hitsA(), ctx->hit_cap, ctx->countA(), ctx->d_meta,
                                                  ctx->d_nontrivial, mode, minimize, ascending_key, check_trivial, prepped ? 1 : 0);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}

int launch_nms(mtm_ctx* ctx, float thr32, int ascending, int64_t n_object, float max_overlap)
{
    nms_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->hitsA(), ctx->hit_cap, ctx->countA(), ctx->hitsB(), ctx->countB(),
                                            ctx->d_keep, thr32, ascending, (long long)n_object, max_overlap);
    MTM_LAUNCH_CHECK(ctx);
    return MTM_OK;
}
