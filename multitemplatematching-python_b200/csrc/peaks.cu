// peaks.cu -- K5/K6: peak extraction from the fp32 score maps on the device.
//
//  * peaks2d_kernel  : skimage.feature.peak_local_max(map, threshold_abs=thr,
//                      exclude_border=False) as called at MTM/__init__.py:45
//                      (3x3 maximum, borders included, strict '> thr', constant
//                      map -> no peaks, whole plateaus kept).
//  * peaks1d_kernel  : the degenerate maps of MTM/__init__.py:25-41 -- 1x1 map
//                      ('>= thr') and 1xn / nx1 maps (scipy.signal.find_peaks
//                      with height=thr: strict neighbours, plateau midpoint,
//                      edges excluded, float64 comparison, ascending order).
//  * argbest_kernel  : cv2.minMaxLoc (MTM/__init__.py:226): global max (min for
//                      methods 0/1), first occurrence in row-major order.
// For methods 0/1 (MTM/__init__.py:51-53, 232-233) the map and the threshold are
// negated, exactly like _findLocalMin_.  HBM-bound: one 4-byte read per score
// pixel (neighbours come from L1/L2).
#include "mtm_internal.cuh"

namespace {

// Streaming pass: one aligned float4 per thread.  The "every pixel is a 3x3 maximum" rule of
// peak_local_max only fires for a constant map, i.e. iff no pixel differs from pixel 0; the 3x3
// test (8 extra loads, L1/L2 hits) is only run for the pixels above the threshold.
__global__ void peaks2d_kernel(const TmplMeta* __restrict__ meta, const float* __restrict__ maps,
                               float thr32, int minimize, DevHit* __restrict__ hits, int cap,
                               int32_t* __restrict__ count, int32_t* __restrict__ nontrivial, int skip_const)
{
    const TmplMeta& tm = meta[blockIdx.y];
    const int mh = tm.mh, mw = tm.mw;
    if (mh == 1 || mw == 1) return;                          // handled by peaks1d_kernel
    if (skip_const && tm.is_const) return;                   // TM_CCOEFF_NORMED map of a constant template: all 1 -> no peaks (nontrivial stays 0)
    const int64_t n = (int64_t)mh * mw;
    const float* m = maps + tm.map_off;                      // 128-byte aligned (map offsets are multiples of 32)
    const float sgn = minimize ? -1.0f : 1.0f;
    const float m0 = m[0];
    int differs = 0;
    // each thread takes UNR aligned float4 per sweep (UNR independent 16-byte loads in flight)
    constexpr int UNR = 4;
    const int64_t sweep = 4 * (int64_t)gridDim.x * blockDim.x;
    for (int64_t base0 = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); base0 < n; base0 += UNR * sweep) {
        float4 q[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t base = base0 + u * sweep;
            if (base + 4 <= n) q[u] = __ldg(reinterpret_cast<const float4*>(m + base));
            else {
                q[u].x = (base + 0 < n) ? m[base + 0] : m0; q[u].y = (base + 1 < n) ? m[base + 1] : m0;
                q[u].z = (base + 2 < n) ? m[base + 2] : m0; q[u].w = (base + 3 < n) ? m[base + 3] : m0;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t base = base0 + u * sweep;
            const float v4[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float raw = v4[k];
                if (raw != m0) differs = 1;
                const float v = sgn * raw;
                if (!(v > thr32) || base + k >= n) continue;
                const int64_t idx = base + k;
                const int y = (int)(idx / mw), x = (int)(idx - (int64_t)y * mw);
                bool is_max = true;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= mh) continue;
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int xx = x + dx;
                        if (xx < 0 || xx >= mw || (dx == 0 && dy == 0)) continue;
                        if (sgn * m[(int64_t)yy * mw + xx] > v) is_max = false;
                    }
                }
                if (is_max) {
                    const int slot = atomicAdd(count, 1);
                    if (slot < cap) {
                        DevHit h;
                        h.tmpl = blockIdx.y; h.x = x; h.y = y; h.w = tm.w; h.h = tm.h;
                        h.score = raw; h.seq = 0; h.key = 0.f;
                        hits[slot] = h;
                    }
                }
            }
        }
    }
    if (__syncthreads_or(differs) && threadIdx.x == 0 && !nontrivial[blockIdx.y]) atomicOr(&nontrivial[blockIdx.y], 1);
}

// One thread per degenerate template (these maps have at most max(H, W) entries).
__global__ void peaks1d_kernel(const TmplMeta* __restrict__ meta, int n_tmpl, const float* __restrict__ maps,
                               float thr32, double thr64, int minimize, DevHit* __restrict__ hits, int cap,
                               int32_t* __restrict__ count, int32_t* __restrict__ nontrivial)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tmpl) return;
    const TmplMeta& tm = meta[t];
    if (tm.mh != 1 && tm.mw != 1) return;
    nontrivial[t] = 1;                                        // the "constant map" rule is 2-D only
    const float* m = maps + tm.map_off;
    const double sgn = minimize ? -1.0 : 1.0;
    auto emit = [&](int i) {
        const int slot = atomicAdd(count, 1);
        if (slot < cap) {
            DevHit h;
            h.tmpl = t; h.w = tm.w; h.h = tm.h; h.score = m[i]; h.seq = 0; h.key = 0.f;
            if (tm.mh == 1) { h.x = i; h.y = 0; } else { h.x = 0; h.y = i; }
            hits[slot] = h;
        }
    };
    const int n = tm.mh * tm.mw;
    if (n == 1) {                                             // MTM/__init__.py:25-30, float32 '>='
        const float v = minimize ? -m[0] : m[0];
        if (v >= thr32) emit(0);
        return;
    }
    int i = 1;                                                // scipy _local_maxima_1d
    while (i < n - 1) {
        const double vi = sgn * (double)m[i];
        if (sgn * (double)m[i - 1] < vi) {
            int ahead = i + 1;
            while (ahead < n - 1 && sgn * (double)m[ahead] == vi) ++ahead;
            if (sgn * (double)m[ahead] < vi) {
                const int mid = (i + ahead - 1) / 2;
                if (sgn * (double)m[mid] >= thr64) emit(mid);
                i = ahead;
            }
        }
        ++i;
    }
}

// 3x3 test for the above-threshold pixels listed by the tensor-core epilogue (instead of streaming
// every score map again).  count[3] = 1 reports a list overflow: the host re-runs peaks2d_kernel.
__global__ void verify_candidates_kernel(const TmplMeta* __restrict__ meta, int n_tmpl, const float* __restrict__ maps,
                                         const DevHit* __restrict__ cand, const int32_t* __restrict__ cand_count, int cand_cap,
                                         DevHit* __restrict__ hits, int cap, int32_t* __restrict__ count,
                                         int32_t* __restrict__ nontrivial)
{
    const int n = *cand_count;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n_tmpl) nontrivial[gid] = 1;           // only used for maps with more pixels than the list holds
    if (n > cand_cap) { if (gid == 0) count[3] = 1; return; }
    for (int i = gid; i < n; i += gridDim.x * blockDim.x) {
        const DevHit c = cand[i];
        const TmplMeta& tm = meta[c.tmpl];
        const float* m = maps + tm.map_off;
        bool is_max = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = c.y + dy;
            if (yy < 0 || yy >= tm.mh) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = c.x + dx;
                if (xx < 0 || xx >= tm.mw || (dx == 0 && dy == 0)) continue;
                if (m[(int64_t)yy * tm.mw + xx] > c.score) is_max = false;
            }
        }
        if (is_max) {
            const int slot = atomicAdd(count, 1);
            if (slot < cap) hits[slot] = c;
        }
    }
}

// Hits-only searches have no score map to look the neighbours up in.  They are not needed: a neighbour that beats a candidate
// scores above the threshold itself, so it is in the list too.  ONE CTA: (1) every candidate enters an open-addressing table
// keyed by (template, map index); (2) every candidate probes its (up to) eight neighbours and survives unless one of them is
// listed with a larger score -- the same test as verify_candidates_kernel, '>' on the same float32 scores; (3) the used
// slots are emptied again, so the table needs no memset between calls.  count[3] = 1 reports a list overflow.
__device__ __forceinline__ uint32_t cand_hash(unsigned long long key)
{
    key ^= key >> 33; key *= 0xff51afd7ed558ccdull; key ^= key >> 29;
    return (uint32_t)key & (MTM_HASH_SLOTS - 1);
}

__global__ void __launch_bounds__(1024, 1)
resolve_candidates_kernel(const TmplMeta* __restrict__ meta, int n_tmpl, DevHit* __restrict__ cand, const int32_t* __restrict__ cand_count,
                          int cand_cap, unsigned long long* __restrict__ hkeys, int32_t* __restrict__ hvals,
                          DevHit* __restrict__ hits, int cap, int32_t* __restrict__ count, int32_t* __restrict__ nontrivial)
{
    constexpr unsigned long long EMPTY = ~0ull;
    const int n = *cand_count;
    const int tid = threadIdx.x;
    for (int t = tid; t < n_tmpl; t += blockDim.x) nontrivial[t] = 1;     // only used for maps with more pixels than the list holds
    if (n > cand_cap) { if (tid == 0) count[3] = 1; return; }
    for (int i = tid; i < n; i += blockDim.x) {
        const DevHit c = cand[i];
        const unsigned long long key = ((unsigned long long)(uint32_t)c.tmpl << 32) | (uint32_t)(c.y * meta[c.tmpl].mw + c.x);
        uint32_t h = cand_hash(key);
        while (atomicCAS(&hkeys[h], EMPTY, key) != EMPTY) h = (h + 1) & (MTM_HASH_SLOTS - 1);     // keys are unique: a pixel is listed once
        hvals[h] = i;
        cand[i].seq = (int32_t)h;
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const DevHit c = cand[i];
        const TmplMeta& tm = meta[c.tmpl];
        // the eight first probes are independent loads (one round trip); collisions continue one by one
        unsigned long long want[8], found[8];
        uint32_t slot_of[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dy = (q < 3) ? -1 : (q < 5 ? 0 : 1);
            const int dx = (q < 3) ? q - 1 : (q == 3 ? -1 : (q == 4 ? 1 : q - 6));
            const int yy = c.y + dy, xx = c.x + dx;
            const bool inside = yy >= 0 && yy < tm.mh && xx >= 0 && xx < tm.mw;
            want[q] = inside ? (((unsigned long long)(uint32_t)c.tmpl << 32) | (uint32_t)(yy * tm.mw + xx)) : EMPTY;
            slot_of[q] = cand_hash(want[q]);
            found[q] = inside ? hkeys[slot_of[q]] : EMPTY;
        }
        bool is_max = true;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            unsigned long long k = found[q];
            uint32_t h = slot_of[q];
            while (k != EMPTY && k != want[q]) { h = (h + 1) & (MTM_HASH_SLOTS - 1); k = hkeys[h]; }
            if (k != EMPTY && cand[hvals[h]].score > c.score) is_max = false;      // an empty slot: not listed, its score is at most the threshold
        }
        if (is_max) {
            const int slot = atomicAdd(count, 1);
            if (slot < cap) { DevHit o = c; o.seq = 0; hits[slot] = o; }
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) hkeys[cand[i].seq] = EMPTY;
}

// N_object == 1 of a hits-only search: the epilogues of the numerator kernels raised best[t] (flush_best, ncc_tc.cu).
__global__ void emit_best_keys_kernel(const TmplMeta* __restrict__ meta, int n_tmpl, const unsigned long long* __restrict__ best,
                                      DevHit* __restrict__ hits, int32_t* __restrict__ count, int cap)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) count[0] = n_tmpl;
    if (t >= n_tmpl || t >= cap) return;
    const TmplMeta& tm = meta[t];
    const unsigned long long key = best[t];
    const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
    const uint32_t o = (uint32_t)(key >> 32);
    const uint32_t bits = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;    // inverse of ordered_f32
    DevHit h;
    h.tmpl = t; h.y = (int)(idx / (uint32_t)tm.mw); h.x = (int)(idx - (uint32_t)h.y * (uint32_t)tm.mw);
    h.w = tm.w; h.h = tm.h; h.score = __uint_as_float(bits); h.seq = t; h.key = 0.f;
    hits[t] = h;
}

__global__ void argbest_kernel(const TmplMeta* __restrict__ meta, const float* __restrict__ maps,
                               int minimize, unsigned long long* __restrict__ best)
{
    const TmplMeta& tm = meta[blockIdx.y];
    const int64_t n = (int64_t)tm.mh * tm.mw;
    const float* m = maps + tm.map_off;
    unsigned long long k = 0ull;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const float v = minimize ? -m[idx] : m[idx];
        const unsigned long long key =
            ((unsigned long long)ordered_f32(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
        k = key > k ? key : k;
    }
    for (int d = 16; d; d >>= 1) {
        const unsigned long long o = __shfl_down_sync(0xffffffffu, k, d);
        k = o > k ? o : k;
    }
    if ((threadIdx.x & 31) == 0 && k) atomicMax(&best[blockIdx.y], k);
}

__global__ void emit_best_kernel(const TmplMeta* __restrict__ meta, int n_tmpl, const float* __restrict__ maps,
                                 const unsigned long long* __restrict__ best, DevHit* __restrict__ hits,
                                 int32_t* __restrict__ count, int cap)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) count[0] = n_tmpl;                           // more templates than the block holds: the host grows it and repeats
    if (t >= n_tmpl || t >= cap) return;
    const TmplMeta& tm = meta[t];
    const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(best[t] & 0xFFFFFFFFull);
    DevHit h;
    h.tmpl = t; h.y = (int)(idx / (uint32_t)tm.mw); h.x = (int)(idx - (uint32_t)h.y * (uint32_t)tm.mw);
    h.w = tm.w; h.h = tm.h; h.score = maps[tm.map_off + idx]; h.seq = t; h.key = 0.f;
    hits[t] = h;
}

}  // namespace

// Fills block A (hits + count[0]) with the raw (unsorted) peaks of every template.
int launch_peaks(mtm_ctx* ctx, int method, int64_t n_object, float thr32, double thr64, bool allow_candidates)
{
    const int nt = ctx->n_tmpl;
    const int minimize = method_is_min(method) ? 1 : 0;
    int64_t max_px = 0;
    for (int t = 0; t < nt; ++t) {
        const int64_t n = (int64_t)ctx->h_meta[t].mh * ctx->h_meta[t].mw;
        if (n >= (1ll << 31)) return mtm_fail(ctx, MTM_ERR_UNSUPPORTED, "score map larger than 2^31 pixels");
        max_px = n > max_px ? n : max_px;
    }
    MTM_CUDA(ctx, cudaMemsetAsync(ctx->countA(), 0, MTM_HIT_HEADER, ctx->stream));
    int bx = (int)((max_px + 4095) / 4096);
    const int bx_cap = ctx->sm_count * 8;
    if (bx > bx_cap) bx = bx_cap;
    if (bx < 1) bx = 1;
    if (n_object == 1 && ctx->best_valid) {                  // hits-only search: the numerator kernels' epilogues found the arg-max
        emit_best_keys_kernel<<<(nt + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_meta, nt, ctx->d_best, ctx->hitsA(), ctx->countA(), ctx->hit_cap);
        MTM_LAUNCH_CHECK(ctx);
        return MTM_OK;
    }
    const bool listed = allow_candidates && ctx->cand_valid && ctx->cand_thr == thr32 && !minimize && n_object != 1;
    if (!ctx->maps_resident && !listed)
        return mtm_fail(ctx, MTM_ERR_INVALID, "internal: the peak search needs score maps that this search did not write");
    if (n_object == 1) {
        MTM_CUDA(ctx, cudaMemsetAsync(ctx->d_best, 0, nt * sizeof(unsigned long long), ctx->stream));
        argbest_kernel<<<dim3(bx, nt), 256, 0, ctx->stream>>>(ctx->d_meta, ctx->d_maps, minimize, ctx->d_best);
        MTM_LAUNCH_CHECK(ctx);
        emit_best_kernel<<<(nt + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_meta, nt, ctx->d_maps, ctx->d_best,
                                                                   ctx->hitsA(), ctx->countA(), ctx->hit_cap);
        MTM_LAUNCH_CHECK(ctx);
        return MTM_OK;
    }
    if (listed && !ctx->maps_resident) {
        // hits-only search: 3x3 maxima resolved inside the list of above-threshold pixels
        resolve_candidates_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_meta, nt, ctx->d_cand, ctx->d_cand_count, MTM_CAND_CAP, ctx->d_hkeys,
                                                              ctx->d_hvals, ctx->hitsA(), ctx->hit_cap, ctx->countA(), ctx->d_nontrivial);
        MTM_LAUNCH_CHECK(ctx);
        return MTM_OK;
    }
    if (listed) {
        // the epilogue of the numerator kernel already listed every pixel above the threshold
        verify_candidates_kernel<<<64, 256, 0, ctx->stream>>>(ctx->d_meta, nt, ctx->d_maps, ctx->d_cand, ctx->d_cand_count,
                                                             MTM_CAND_CAP, ctx->hitsA(), ctx->hit_cap, ctx->countA(),
                                                             ctx->d_nontrivial);
        MTM_LAUNCH_CHECK(ctx);
        return MTM_OK;
    }
    MTM_CUDA(ctx, cudaMemsetAsync(ctx->d_nontrivial, 0, nt * sizeof(int32_t), ctx->stream));
    bool any2d = false, any1d = false;
    for (int t = 0; t < nt; ++t) {
        if (ctx->h_meta[t].mh == 1 || ctx->h_meta[t].mw == 1) any1d = true; else any2d = true;
    }
    const float t32 = minimize ? -thr32 : thr32;
    const double t64 = minimize ? -thr64 : thr64;
    if (any2d) {
        peaks2d_kernel<<<dim3(bx, nt), 256, 0, ctx->stream>>>(ctx->d_meta, ctx->d_maps, t32, minimize,
                                                              ctx->hitsA(), ctx->hit_cap, ctx->countA(),
                                                              ctx->d_nontrivial, method == MTM_TM_CCOEFF_NORMED && ctx->img_dtype == MTM_U8 ? 1 : 0);
        MTM_LAUNCH_CHECK(ctx);
    }
    if (any1d) {
        peaks1d_kernel<<<(nt + 63) / 64, 64, 0, ctx->stream>>>(ctx->d_meta, nt, ctx->d_maps, t32, t64, minimize,
                                                               ctx->hitsA(), ctx->hit_cap, ctx->countA(),
                                                               ctx->d_nontrivial);
        MTM_LAUNCH_CHECK(ctx);
    }
    return MTM_OK;
}
