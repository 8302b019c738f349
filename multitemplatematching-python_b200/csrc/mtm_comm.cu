// mtm_comm.cu -- the multi-GPU entry points of the C ABI (include/mtm_b200.h, "multi-GPU").
//
// The reference's parallel axis is "one task per template" (MTM/__init__.py:172-175), coupled again only by the
// NMS (MTM/__init__.py:294-296).  Across GPUs that needs ONE exchange per call: an all-gather of fixed-size hit
// blocks (32-byte header + 32-byte DevHit records).  It is issued here, on the context's stream, between the
// peak kernels and the replicated NMS kernel:
//     peaks (block A) -> shard_pack_kernel (send) -> ncclAllGather (recv) -> shard_merge_kernel (block A) -> finalize / NMS
// so that a sharded call has no host round trip besides the final read of the result mirror.
// Transports: NCCL (libnccl.so.2 resolved with dlopen: a single-GPU user needs no NCCL), an in-process loop-back
// for several endpoints on ONE device (device-to-device copies ordered by CUDA events + a host barrier; lets a
// single-GPU box run the whole exchange), and a plain copy for world == 1.
#include "mtm_internal.cuh"
#if defined(__CUDACC__) && __has_include(<nccl.h>)
#include <nccl.h>          // types and enums only; every entry point is looked up at run time
#else                      // the part of NCCL's (stable) C interface used here, for builds without its header
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclChar = 0, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclMax = 2 } ncclRedOp_t;
#endif
#include <dlfcn.h>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

thread_local std::string g_comm_err;

struct NcclApi {
    void* handle = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi& nccl_api()
{
    static const NcclApi api = [] {
        NcclApi a;
        const char* override_path = getenv("MTM_B200_NCCL_LIB");
        const char* names[3] = {override_path, "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            a.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (a.handle) break;
            a.err = dlerror();
        }
        if (!a.handle) return a;
        auto sym = [&](const char* nm) {
            void* p = dlsym(a.handle, nm);
            if (!p) { a.err = std::string("libnccl lacks ") + nm; }
            return p;
        };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        if (!a.GetUniqueId || !a.CommInitRank || !a.CommInitAll || !a.CommDestroy || !a.AllGather || !a.AllReduce || !a.GetErrorString) {
            dlclose(a.handle);
            a.handle = nullptr;
        }
        return a;
    }();
    return api;
}

// Host-side rendezvous of the endpoints of an in-process loop-back group (one host thread per endpoint).
struct LoopGroup {
    int n = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t gen = 0;
    std::vector<mtm_comm*> members;
    std::vector<double> red;
    int refs = 0;
    void barrier()
    {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t g = gen;
        if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

}  // namespace

struct mtm_comm {
    int world = 1, rank = 0, device = 0;
    ncclComm_t nccl = nullptr;
    LoopGroup* loop = nullptr;
    cudaStream_t stream = nullptr;                 // mtm_gather_results / mtm_comm_allreduce_max
    cudaEvent_t ev_sent = nullptr, ev_consumed = nullptr;
    uint8_t* d_send = nullptr; uint8_t* d_recv = nullptr;
    size_t send_cap = 0, recv_cap = 0;             // bytes
    uint8_t* h_recv = nullptr; size_t h_recv_cap = 0;   // pinned
    double* d_red = nullptr; double* h_red = nullptr;   // 64 doubles each
    int cap_g = 1024;                              // hits per rank block of the template cut (grows in lock step on every rank)
    std::string err;
};

namespace {

constexpr int COMM_RED_MAX = 64;

int comm_fail(mtm_comm* c, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_comm_err = buf;
    return code;
}

#define COMM_CUDA(c, call)                                                                             \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return comm_fail(c, MTM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define COMM_NCCL(c, call)                                                                             \
    do {                                                                                               \
        ncclResult_t r__ = (call);                                                                     \
        if (r__ != ncclSuccess)                                                                        \
            return comm_fail(c, MTM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

int comm_need_nccl(mtm_comm* c)
{
    const NcclApi& a = nccl_api();
    if (!a.handle)
        return comm_fail(c, MTM_ERR_UNSUPPORTED, "NCCL is not available (%s); set MTM_B200_NCCL_LIB to libnccl.so.2", a.err.c_str());
    return MTM_OK;
}

int comm_setup_local(mtm_comm* c)
{
    COMM_CUDA(c, cudaSetDevice(c->device));
    COMM_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    COMM_CUDA(c, cudaEventCreateWithFlags(&c->ev_sent, cudaEventDisableTiming));
    COMM_CUDA(c, cudaEventCreateWithFlags(&c->ev_consumed, cudaEventDisableTiming));
    COMM_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_red), COMM_RED_MAX * sizeof(double)));
    COMM_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_red), COMM_RED_MAX * sizeof(double)));
    return MTM_OK;
}

// Exchange buffers: `send_bytes` of this rank, `world * send_bytes` gathered.  Growing waits for the device: nobody may
// still read the old ones (peers only write into d_recv before this rank's consumer kernel, which has completed by then).
int comm_reserve(mtm_comm* c, size_t send_bytes, bool host_mirror)
{
    const size_t recv_bytes = send_bytes * (size_t)c->world;
    if (send_bytes > c->send_cap || recv_bytes > c->recv_cap) {
        COMM_CUDA(c, cudaDeviceSynchronize());
        if (c->d_send) cudaFree(c->d_send);
        if (c->d_recv) cudaFree(c->d_recv);
        c->d_send = c->d_recv = nullptr; c->send_cap = c->recv_cap = 0;
        COMM_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_send), send_bytes));
        COMM_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->d_recv), recv_bytes));
        c->send_cap = send_bytes; c->recv_cap = recv_bytes;
    }
    if (host_mirror && recv_bytes > c->h_recv_cap) {
        if (c->h_recv) cudaFreeHost(c->h_recv);
        c->h_recv = nullptr; c->h_recv_cap = 0;
        COMM_CUDA(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_recv), recv_bytes));
        c->h_recv_cap = recv_bytes;
    }
    return MTM_OK;
}

// recv[r * bytes .. ) := rank r's send[0 .. bytes) for every r, ordered on stream `s`.
int comm_all_gather(mtm_comm* c, cudaStream_t s, size_t bytes)
{
    if (c->world == 1) {
        COMM_CUDA(c, cudaMemcpyAsync(c->d_recv, c->d_send, bytes, cudaMemcpyDeviceToDevice, s));
        return MTM_OK;
    }
    if (c->nccl) {
        COMM_NCCL(c, nccl_api().AllGather(c->d_send, c->d_recv, bytes, ncclChar, c->nccl, s));
        return MTM_OK;
    }
    LoopGroup* g = c->loop;
    g->barrier();                                  // every endpoint is inside the call: buffers sized, last consumption recorded
    for (mtm_comm* p : g->members) {
        COMM_CUDA(c, cudaStreamWaitEvent(s, p->ev_consumed, 0));
        COMM_CUDA(c, cudaMemcpyAsync(p->d_recv + (size_t)c->rank * bytes, c->d_send, bytes, cudaMemcpyDeviceToDevice, s));
    }
    COMM_CUDA(c, cudaEventRecord(c->ev_sent, s));
    g->barrier();                                  // every endpoint has queued its copies
    for (mtm_comm* p : g->members) COMM_CUDA(c, cudaStreamWaitEvent(s, p->ev_sent, 0));
    return MTM_OK;
}

// The kernel reading d_recv has been queued on `s`: peers may overwrite it in the next exchange once it ran.
int comm_consumed(mtm_comm* c, cudaStream_t s)
{
    if (c->loop) COMM_CUDA(c, cudaEventRecord(c->ev_consumed, s));
    return MTM_OK;
}

// ---------------------------------------------------------------------------- template cut: pack / merge
// Block layout: int32 header[8] + DevHit[cap].  header: [0] hits of this rank (may exceed cap: the receivers then ask for a
// larger block), [1] raw peak count of the rank's own block A, [3] candidate-list overflow, [5] status of the rank's local stage.
__global__ void __launch_bounds__(1024, 1)
shard_pack_kernel(const DevHit* __restrict__ hits, const int32_t* __restrict__ count, int cap_in, const TmplMeta* __restrict__ meta,
                  const int32_t* __restrict__ nontrivial, int check_trivial, int presorted, int tmpl_base,
                  uint8_t* __restrict__ send, int cap_g, int status)
{
    __shared__ int s_n;
    int32_t* hdr = reinterpret_cast<int32_t*>(send);
    DevHit* out = reinterpret_cast<DevHit*>(send + MTM_HIT_HEADER);
    const int tid = threadIdx.x;
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int n_raw = hits ? count[0] : 0;
    const int n = min(n_raw, cap_in);
    for (int i = tid; i < n; i += blockDim.x) {
        DevHit h = hits[i];
        if (check_trivial && !nontrivial[h.tmpl]) continue;          // constant-map rule of peak_local_max: per template, hence local
        if (presorted) h.seq = tmpl_base + h.tmpl;                   // N_object == 1: one hit per template, list order = template order
        else prep_mode0(h, meta);
        h.tmpl += tmpl_base;
        const int slot = atomicAdd(&s_n, 1);
        if (slot < cap_g) out[slot] = h;
    }
    __syncthreads();
    if (tid < 8) {
        int v = 0;
        if (tid == 0) v = s_n;
        else if (tid == 1) v = n_raw;
        else if (tid == 3) v = hits ? count[3] : 0;
        else if (tid == 5) v = status;
        hdr[tid] = v;
    }
}

// Rank-ordered concatenation of the gathered blocks -> block A (header + hits).
__global__ void shard_merge_kernel(const uint8_t* __restrict__ recv, int world, int cap_g, size_t block_bytes,
                                   DevHit* __restrict__ out, int out_cap, int32_t* __restrict__ out_count)
{
    const int64_t total_slots = (int64_t)world * cap_g;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_slots; idx += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / cap_g), k = (int)(idx - (int64_t)r * cap_g);
        const int32_t* hdr = reinterpret_cast<const int32_t*>(recv + (size_t)r * block_bytes);
        if (k >= min(max(hdr[0], 0), cap_g)) continue;
        int off = 0;
        for (int q = 0; q < r; ++q) off += min(max(reinterpret_cast<const int32_t*>(recv + (size_t)q * block_bytes)[0], 0), cap_g);
        if (off + k < out_cap)
            out[off + k] = reinterpret_cast<const DevHit*>(recv + (size_t)r * block_bytes + MTM_HIT_HEADER)[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int total = 0, cand_over = 0, need = 0, status = 0, raw_max = 0;
        for (int r = 0; r < world; ++r) {
            const int32_t* hdr = reinterpret_cast<const int32_t*>(recv + (size_t)r * block_bytes);
            total += min(max(hdr[0], 0), cap_g);
            need = max(need, hdr[0]);
            raw_max = max(raw_max, hdr[1]);
            cand_over |= hdr[3];
            status = min(status, hdr[5]);
        }
        out_count[0] = total; out_count[1] = 0; out_count[2] = 0; out_count[3] = cand_over;
        out_count[4] = need; out_count[5] = status; out_count[6] = raw_max; out_count[7] = 0;
    }
}

// ---------------------------------------------------------------------------- image cut: pack the final lists of the slots
constexpr int GATHER_MAX_SRC = 32;
struct GatherSrc { const uint8_t* slot[GATHER_MAX_SRC]; };

// One CTA per image entry of this rank's send buffer: header + the first hits_per_image hits of its result slot.
__global__ void gather_pack_kernel(GatherSrc src, int first, uint8_t* __restrict__ send, int hits_per_image)
{
    const size_t block_img = MTM_HIT_HEADER + (size_t)hits_per_image * sizeof(DevHit);
    const uint8_t* s = src.slot[blockIdx.x];
    uint8_t* d = send + (size_t)(first + blockIdx.x) * block_img;
    int32_t* dh = reinterpret_cast<int32_t*>(d);
    if (!s) { if (threadIdx.x < 8) dh[threadIdx.x] = threadIdx.x == 0 ? -1 : 0; return; }      // unused entry
    const int32_t* sh = reinterpret_cast<const int32_t*>(s);
    int n = sh[0];
    if (sh[2] || n > hits_per_image || n < 0) n = -2;                 // outside the fused fast path / does not fit
    if (threadIdx.x < 8) dh[threadIdx.x] = threadIdx.x == 0 ? n : sh[threadIdx.x];
    const uint4* sv = reinterpret_cast<const uint4*>(s + MTM_HIT_HEADER);
    uint4* dv = reinterpret_cast<uint4*>(d + MTM_HIT_HEADER);
    for (int i = threadIdx.x; i < 2 * max(n, 0); i += blockDim.x) dv[i] = sv[i];
}

int next_pow2_i(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

extern "C" {

const char* mtm_comm_last_error(const mtm_comm* comm) { return comm ? comm->err.c_str() : g_comm_err.c_str(); }

int mtm_comm_unique_id(void* id_out)
{
    if (!id_out) return comm_fail(nullptr, MTM_ERR_INVALID, "mtm_comm_unique_id: null output");
    MTM_TRY(comm_need_nccl(nullptr));
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == MTM_COMM_ID_BYTES, "NCCL id size");
    COMM_NCCL(nullptr, nccl_api().GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return MTM_OK;
}

int mtm_comm_destroy(mtm_comm* c)
{
    if (!c) return MTM_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->nccl) nccl_api().CommDestroy(c->nccl);
    if (c->loop) {
        bool last;
        { std::lock_guard<std::mutex> lk(c->loop->m); last = (--c->loop->refs == 0); }
        if (last) delete c->loop;
    }
    cudaFree(c->d_send); cudaFree(c->d_recv); cudaFree(c->d_red);
    cudaFreeHost(c->h_recv); cudaFreeHost(c->h_red);
    if (c->ev_sent) cudaEventDestroy(c->ev_sent);
    if (c->ev_consumed) cudaEventDestroy(c->ev_consumed);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return MTM_OK;
}

int mtm_comm_init_rank(int device, int world, int rank, const void* id, mtm_comm** out)
{
    if (!out) return comm_fail(nullptr, MTM_ERR_INVALID, "mtm_comm_init_rank: null output pointer");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return comm_fail(nullptr, MTM_ERR_INVALID, "mtm_comm_init_rank: rank %d of %d", rank, world);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return comm_fail(nullptr, MTM_ERR_CUDA, "mtm_comm_init_rank: no CUDA device; libmtm_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev) return comm_fail(nullptr, MTM_ERR_INVALID, "mtm_comm_init_rank: device %d out of range [0, %d)", device, n_dev);
    mtm_comm* c = new mtm_comm();
    c->world = world; c->rank = rank; c->device = device;
    int rc = comm_setup_local(c);
    if (rc == MTM_OK && world > 1) {
        if (!id) rc = comm_fail(c, MTM_ERR_INVALID, "mtm_comm_init_rank: null id");
        if (rc == MTM_OK) rc = comm_need_nccl(c);
        if (rc == MTM_OK) {
            ncclUniqueId uid;
            memcpy(&uid, id, sizeof uid);
            ncclResult_t r = nccl_api().CommInitRank(&c->nccl, world, uid, rank);
            if (r != ncclSuccess) rc = comm_fail(c, MTM_ERR_CUDA, "ncclCommInitRank failed: %s", nccl_api().GetErrorString(r));
        }
    }
    if (rc != MTM_OK) { g_comm_err = c->err; mtm_comm_destroy(c); return rc; }
    *out = c;
    return MTM_OK;
}

int mtm_comm_create(int n, const int* devices, mtm_comm** out)
{
    if (n < 1 || !devices || !out) return comm_fail(nullptr, MTM_ERR_INVALID, "mtm_comm_create: empty device list");
    for (int i = 0; i < n; ++i) out[i] = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return comm_fail(nullptr, MTM_ERR_CUDA, "mtm_comm_create: no CUDA device; libmtm_b200 has no CPU fallback");
    bool same = true, distinct = true;
    for (int i = 0; i < n; ++i) {
        if (devices[i] < 0 || devices[i] >= n_dev) return comm_fail(nullptr, MTM_ERR_INVALID, "mtm_comm_create: device %d out of range [0, %d)", devices[i], n_dev);
        same = same && devices[i] == devices[0];
        for (int j = 0; j < i; ++j) distinct = distinct && devices[i] != devices[j];
    }
    if (n > 1 && !same && !distinct)
        return comm_fail(nullptr, MTM_ERR_INVALID, "mtm_comm_create: devices must be all distinct (NCCL) or all the same (loop-back)");
    std::vector<mtm_comm*> cs((size_t)n, nullptr);
    int rc = MTM_OK;
    for (int i = 0; i < n && rc == MTM_OK; ++i) {
        cs[i] = new mtm_comm();
        cs[i]->world = n; cs[i]->rank = i; cs[i]->device = devices[i];
        rc = comm_setup_local(cs[i]);
        if (rc != MTM_OK) g_comm_err = cs[i]->err;
    }
    if (rc == MTM_OK && n > 1) {
        if (same) {
            LoopGroup* g = new LoopGroup();
            g->n = n; g->members = cs; g->red.assign((size_t)n * COMM_RED_MAX, 0.0); g->refs = n;
            for (mtm_comm* c : cs) c->loop = g;
        } else {
            rc = comm_need_nccl(nullptr);
            if (rc == MTM_OK) {
                std::vector<ncclComm_t> nc((size_t)n, nullptr);
                ncclResult_t r = nccl_api().CommInitAll(nc.data(), n, devices);
                if (r != ncclSuccess) rc = comm_fail(nullptr, MTM_ERR_CUDA, "ncclCommInitAll failed: %s", nccl_api().GetErrorString(r));
                else for (int i = 0; i < n; ++i) cs[i]->nccl = nc[i];
            }
        }
    }
    if (rc != MTM_OK) {
        for (mtm_comm* c : cs) if (c) { c->loop = nullptr; mtm_comm_destroy(c); }
        return rc;
    }
    for (int i = 0; i < n; ++i) out[i] = cs[i];
    return MTM_OK;
}

int mtm_comm_info(const mtm_comm* c, int* world, int* rank, int* device)
{
    if (!c) return MTM_ERR_INVALID;
    if (world) *world = c->world;
    if (rank) *rank = c->rank;
    if (device) *device = c->device;
    return MTM_OK;
}

int mtm_comm_allreduce_max(mtm_comm* c, double* values, int n)
{
    if (!c || !values || n < 0 || n > COMM_RED_MAX) return comm_fail(c, MTM_ERR_INVALID, "mtm_comm_allreduce_max: bad arguments (n <= %d)", COMM_RED_MAX);
    if (c->world == 1 || n == 0) return MTM_OK;
    if (c->loop) {
        LoopGroup* g = c->loop;
        g->barrier();
        for (int i = 0; i < n; ++i) g->red[(size_t)c->rank * COMM_RED_MAX + i] = values[i];
        g->barrier();
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < c->world; ++r) values[i] = std::max(values[i], g->red[(size_t)r * COMM_RED_MAX + i]);
        g->barrier();
        return MTM_OK;
    }
    COMM_CUDA(c, cudaSetDevice(c->device));
    memcpy(c->h_red, values, (size_t)n * sizeof(double));
    COMM_CUDA(c, cudaMemcpyAsync(c->d_red, c->h_red, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    COMM_NCCL(c, nccl_api().AllReduce(c->d_red, c->d_red, (size_t)n, ncclDouble, ncclMax, c->nccl, c->stream));
    COMM_CUDA(c, cudaMemcpyAsync(c->h_red, c->d_red, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    COMM_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(values, c->h_red, (size_t)n * sizeof(double));
    return MTM_OK;
}

int mtm_comm_barrier(mtm_comm* c)
{
    double one = 1.0;
    return mtm_comm_allreduce_max(c, &one, 1);
}

int mtm_match_templates_sharded(mtm_ctx* ctx, mtm_comm* comm, int tmpl_base, int n_local, int method, int64_t n_object,
                                double score_threshold, double max_overlap, mtm_hit* hits, int capacity, int* n_hits)
{
    MTM_ENTER(ctx);
    // argument errors that are identical on every rank may return before the exchange
    if (!comm) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_sharded: null communicator");
    if (comm->device != ctx->device) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_sharded: context on device %d, communicator on %d", ctx->device, comm->device);
    if (!n_hits || (capacity > 0 && !hits)) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_sharded: null output");
    if (method == MTM_TM_SQDIFF) return mtm_fail(ctx, MTM_ERR_INVALID, "The method TM_SQDIFF is not supported. Use TM_SQDIFF_NORMED instead.");
    if (tmpl_base < 0 || n_local < 0) return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_sharded: negative slice");

    // ---- local stage: score maps of this rank's slice.  A failure is carried through the exchange (header[5]).
    auto local_stage = [&]() -> int {
        if (n_local == 0) return MTM_OK;
        if (n_local != ctx->n_tmpl)
            return mtm_fail(ctx, MTM_ERR_INVALID, "mtm_match_templates_sharded: slice of %d templates, %d resident in the context", n_local, ctx->n_tmpl);
        MTM_TRY(ensure_geometry(ctx));
        request_candidates(ctx, method, n_object, score_threshold);
        MTM_TRY(compute_maps(ctx, method, -1, true));
        return MTM_OK;
    };
    int status = local_stage();
    std::string local_err = ctx->err;
    const int minimize = method_is_min(method) ? 1 : 0;
    const int ascending = (method == MTM_TM_SQDIFF_NORMED) ? 1 : 0;
    const float thr_nms = ascending ? (float)(1.0 - score_threshold) : (float)score_threshold;
    const int presorted = n_object == 1 ? 1 : 0;
    for (int attempt = 0; attempt < 8; ++attempt) {
        const int cap_g = comm->cap_g;
        const size_t block = MTM_HIT_HEADER + (size_t)cap_g * sizeof(DevHit);
        if (comm_reserve(comm, block, false) != MTM_OK) return mtm_fail(ctx, MTM_ERR_CUDA, "%s", comm->err.c_str());
        if (status == MTM_OK && n_local > 0) {
            status = launch_peaks(ctx, method, n_object, (float)score_threshold, score_threshold);
            if (status != MTM_OK) local_err = ctx->err;
        }
        const bool have = status == MTM_OK && n_local > 0;
        shard_pack_kernel<<<1, 1024, 0, ctx->stream>>>(have ? ctx->hitsA() : nullptr, ctx->countA(), ctx->hit_cap, ctx->d_meta, ctx->d_nontrivial,
                                                      presorted ? 0 : 1, presorted, tmpl_base, comm->d_send, cap_g, status);
        MTM_LAUNCH_CHECK(ctx);
        if (comm_all_gather(comm, ctx->stream, block) != MTM_OK) return mtm_fail(ctx, MTM_ERR_CUDA, "%s", comm->err.c_str());
        const int64_t slots = (int64_t)comm->world * cap_g;
        shard_merge_kernel<<<(int)std::min<int64_t>((slots + 255) / 256, 64), 256, 0, ctx->stream>>>(comm->d_recv, comm->world, cap_g, block,
                                                                                                   ctx->hitsA(), ctx->hit_cap, ctx->countA());
        MTM_LAUNCH_CHECK(ctx);
        if (comm_consumed(comm, ctx->stream) != MTM_OK) return mtm_fail(ctx, MTM_ERR_CUDA, "%s", comm->err.c_str());
        // ---- replicated global NMS on the merged list (keys prepared by the owners: prepped)
        MTM_TRY(launch_finalize_small(ctx, minimize, 0, presorted, 1, thr_nms, ascending, n_object, (float)max_overlap, nullptr, true, true));
        int n_raw = 0, n = 0, declined = 0;
        MTM_TRY(download_mirror(ctx, ctx->d_blockB, &n_raw, &n, &declined));
        const int32_t* hdr = reinterpret_cast<const int32_t*>(ctx->h_stage);
        const int need = hdr[4], peer_status = hdr[5], raw_max = hdr[6];
        if (peer_status != 0 || status != MTM_OK) {
            if (status != MTM_OK) { ctx->err = local_err; return status; }
            return mtm_fail(ctx, MTM_ERR_PEER, "mtm_match_templates_sharded: another rank failed its local search (status %d)", peer_status);
        }
        // every rank reads the same header, so every rank takes the same branch below
        if (raw_max > ctx->hit_cap) { MTM_TRY(reserve_hits(ctx, raw_max)); continue; }
        if (need > cap_g) { comm->cap_g = next_pow2_i(need); continue; }
        if (declined == 2) { MTM_TRY(candidates_overflowed(ctx, method)); continue; }          // a candidate list overflowed: stream the maps
        if (declined) {                                                      // more than 1024 merged hits: general path
            if (n_raw > ctx->hit_cap) { MTM_TRY(reserve_hits(ctx, n_raw)); continue; }
            if (!presorted) {
                MTM_TRY(launch_sort_hits(ctx, 0, minimize, 0, 0, true));
                MTM_TRY(launch_sort_hits(ctx, 1, minimize, ascending, 0));
            }
            MTM_TRY(launch_nms(ctx, thr_nms, ascending, n_object, (float)max_overlap));
            MTM_TRY(download_block(ctx, ctx->d_blockB, &n_raw, &n));
        }
        if (n_raw > ctx->hit_cap) { MTM_TRY(reserve_hits(ctx, n_raw)); continue; }
        *n_hits = n;
        if (n > capacity) return mtm_fail(ctx, MTM_ERR_CAPACITY, "mtm_match_templates_sharded: %d hits, caller capacity %d", n, capacity);
        copy_out(ctx, hits, n);
        return MTM_OK;
    }
    return mtm_fail(ctx, MTM_ERR_CUDA, "mtm_match_templates_sharded: hit buffers kept overflowing");
}

int mtm_gather_results(mtm_comm* c, int n_local, mtm_ctx* const* ctxs, const int* slots, int images_per_rank,
                       int hits_per_image, mtm_hit* out_hits, int32_t* out_counts)
{
    if (!c) return MTM_ERR_INVALID;
    if (images_per_rank < 1 || n_local < 0 || n_local > images_per_rank || hits_per_image < 1 || hits_per_image > MTM_SLOT_HITS)
        return comm_fail(c, MTM_ERR_INVALID, "mtm_gather_results: %d local images of %d per rank, %d hits per image (1..%d)", n_local,
                         images_per_rank, hits_per_image, MTM_SLOT_HITS);
    if (!out_hits || !out_counts || (n_local > 0 && (!ctxs || !slots))) return comm_fail(c, MTM_ERR_INVALID, "mtm_gather_results: null argument");
    COMM_CUDA(c, cudaSetDevice(c->device));
    int status = MTM_OK;
    for (int i = 0; i < n_local; ++i) {
        if (!ctxs[i] || slots[i] < 0 || slots[i] >= MTM_MAX_INFLIGHT || !ctxs[i]->slot_busy[slots[i]] || ctxs[i]->device != c->device)
            status = comm_fail(c, MTM_ERR_INVALID, "mtm_gather_results: entry %d has no submission in flight on device %d", i, c->device);
    }
    const size_t block_img = MTM_HIT_HEADER + (size_t)hits_per_image * sizeof(DevHit);
    const size_t block = block_img * (size_t)images_per_rank;
    MTM_TRY(comm_reserve(c, block, true));
    for (int first = 0; first < images_per_rank; first += GATHER_MAX_SRC) {
        GatherSrc src{};
        const int cnt = std::min(GATHER_MAX_SRC, images_per_rank - first);
        for (int k = 0; k < cnt; ++k) {
            const int i = first + k;
            src.slot[k] = nullptr;
            if (status == MTM_OK && i < n_local) {
                COMM_CUDA(c, cudaStreamWaitEvent(c->stream, ctxs[i]->ev_slot[slots[i]], 0));
                src.slot[k] = ctxs[i]->d_slot[slots[i]];
            }
        }
        gather_pack_kernel<<<cnt, 128, 0, c->stream>>>(src, first, c->d_send, hits_per_image);
        COMM_CUDA(c, cudaGetLastError());
    }
    if (status != MTM_OK) {                          // a rank with bad arguments still takes part: its entries read "unused" + status
        const int32_t bad[8] = {-1, 0, 0, 0, 0, status, 0, 0};
        COMM_CUDA(c, cudaMemcpyAsync(c->d_send, bad, sizeof bad, cudaMemcpyHostToDevice, c->stream));
    }
    MTM_TRY(comm_all_gather(c, c->stream, block));
    COMM_CUDA(c, cudaMemcpyAsync(c->h_recv, c->d_recv, block * (size_t)c->world, cudaMemcpyDeviceToHost, c->stream));
    MTM_TRY(comm_consumed(c, c->stream));
    COMM_CUDA(c, cudaStreamSynchronize(c->stream));
    if (status == MTM_OK)
        for (int i = 0; i < n_local; ++i) ctxs[i]->slot_busy[slots[i]] = false;
    int peer_status = 0;
    for (int r = 0; r < c->world; ++r) {
        for (int i = 0; i < images_per_rank; ++i) {
            const uint8_t* b = c->h_recv + (size_t)r * block + (size_t)i * block_img;
            const int32_t* hdr = reinterpret_cast<const int32_t*>(b);
            if (i == 0 && hdr[0] == -1 && hdr[5] != 0) peer_status = hdr[5];
            const int n = hdr[0];
            out_counts[r * images_per_rank + i] = n;
            const DevHit* src = reinterpret_cast<const DevHit*>(b + MTM_HIT_HEADER);
            mtm_hit* dst = out_hits + ((size_t)r * images_per_rank + i) * hits_per_image;
            for (int k = 0; k < n; ++k) {
                dst[k].tmpl = src[k].tmpl; dst[k].x = src[k].x; dst[k].y = src[k].y;
                dst[k].w = src[k].w; dst[k].h = src[k].h; dst[k].score = src[k].score;
            }
        }
    }
    if (status != MTM_OK) return status;
    if (peer_status != 0) return comm_fail(c, MTM_ERR_PEER, "mtm_gather_results: another rank reported status %d", peer_status);
    return MTM_OK;
}

}  // extern "C"
