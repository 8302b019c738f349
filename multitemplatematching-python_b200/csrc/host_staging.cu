// host_staging.cu -- upload pipeline for PAGEABLE host images (what a drop-in caller of
// MTM.matchTemplates(listTemplates, image) passes is a plain numpy array).
//
// cudaMemcpy2DAsync from pageable memory is staged by the driver through its own bounce buffer on the calling thread
// (~10 GB/s, synchronous): 0.18 ms for the 2 MB image of BASELINE configs[1] on top of a 0.2 ms call.  Here the rows are
// copied into page-locked chunks by a small pool of host threads while the copy engine already moves the previous chunk:
//     for every chunk of rows:   wait until the chunk buffer's last DMA has finished (event)
//                                workers: memcpy their share of the rows into the pinned chunk        (host, parallel)
//                                cudaMemcpy2DAsync(pinned chunk -> device rows) on the context's stream (copy engine)
// Page-locked sources (cudaPointerGetAttributes says so) skip all of this and are copied in place.
// MTM_B200_COPY_THREADS=n sets the pool size (default 4, 0 = leave pageable copies to the driver).
#include "mtm_internal.cuh"
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

namespace {

// Workers claim task indices of the current job; the caller takes part.  Idle workers spin briefly before they sleep, so a
// stream of calls a few hundred microseconds apart keeps them hot without burning a core for an idle process.
class CopyPool {
public:
    explicit CopyPool(int n_workers)
    {
        for (int i = 0; i < n_workers; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++generation_; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }

    void parallel_for(int n_tasks, const std::function<void(int)>& fn)
    {
        std::lock_guard<std::mutex> job_lock(job_m_);          // one job at a time (contexts of several host threads share the pool)
        fn_ = &fn; n_tasks_ = n_tasks;
        next_.store(0, std::memory_order_relaxed);
        done_.store(0, std::memory_order_relaxed);
        { std::lock_guard<std::mutex> lk(m_); ++generation_; }
        gen_atomic_.fetch_add(1, std::memory_order_release);
        cv_.notify_all();
        work();
        while (done_.load(std::memory_order_acquire) < n_tasks) std::this_thread::yield();
        fn_ = nullptr;
    }

private:
    void work()
    {
        for (;;) {
            const int i = next_.fetch_add(1, std::memory_order_acq_rel);
            if (i >= n_tasks_) return;
            (*fn_)(i);
            done_.fetch_add(1, std::memory_order_acq_rel);
        }
    }
    void run()
    {
        uint64_t seen = 0;
        for (;;) {
            // spin for ~1 ms on the generation counter, then sleep on the condition variable
            bool woke = false;
            for (int spin = 0; spin < 20000; ++spin) {
                if (gen_atomic_.load(std::memory_order_acquire) != seen) { woke = true; break; }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            if (!woke) {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_ != seen_gen_of(seen) || stop_; });
            }
            if (stop_) return;
            seen = gen_atomic_.load(std::memory_order_acquire);
            if (fn_) work();
        }
    }
    uint64_t seen_gen_of(uint64_t seen) const { return seen; }     // generation_ and gen_atomic_ advance together

    std::vector<std::thread> workers_;
    std::mutex m_, job_m_;
    std::condition_variable cv_;
    uint64_t generation_ = 0;
    std::atomic<uint64_t> gen_atomic_{0};
    bool stop_ = false;
    const std::function<void(int)>* fn_ = nullptr;
    int n_tasks_ = 0;
    std::atomic<int> next_{0}, done_{0};
};

CopyPool* copy_pool()
{
    static CopyPool* pool = [] {
        int n = 4;
        if (const char* v = getenv("MTM_B200_COPY_THREADS")) n = atoi(v);
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0) n = std::min(n, std::max(hw - 1, 0));
        return n > 0 ? new CopyPool(n - 1) : nullptr;          // the caller is the n-th copier
    }();
    return pool;
}

constexpr size_t STAGE_CHUNK = 512 << 10;                      // bytes per pinned chunk: small enough that a 2 MB image already overlaps copy and DMA
constexpr int STAGE_BUFS = MTM_STAGE_BUFS;

}  // namespace

// H rows of row_bytes bytes, `src_stride` apart on the host, to `dst` (device, rows dst_pitch apart) on the context's stream.
int mtm_upload_rows(mtm_ctx* ctx, void* dst, size_t dst_pitch, const void* src, size_t src_stride, size_t row_bytes, int H)
{
    cudaPointerAttributes attr{};
    const cudaError_t pe = cudaPointerGetAttributes(&attr, src);
    if (pe != cudaSuccess) (void)cudaGetLastError();
    const bool pageable = pe != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
    CopyPool* pool = pageable ? copy_pool() : nullptr;
    if (!pool || (size_t)H * row_bytes < 65536) {               // page-locked (or tiny) source: the copy engine reads it in place
        MTM_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch, src, src_stride, row_bytes, (size_t)H, cudaMemcpyHostToDevice, ctx->stream));
        return MTM_OK;
    }
    if (!ctx->h_chunk[0]) {
        for (int k = 0; k < STAGE_BUFS; ++k) {
            MTM_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->h_chunk[k]), STAGE_CHUNK));
            MTM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk[k], cudaEventDisableTiming));
        }
    }
    const int rows_per_chunk = (int)std::max<size_t>(1, STAGE_CHUNK / row_bytes);
    if (row_bytes > STAGE_CHUNK) {                              // rows wider than a chunk (never for images <= 66000 elements of 4 bytes... keep it safe)
        MTM_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch, src, src_stride, row_bytes, (size_t)H, cudaMemcpyHostToDevice, ctx->stream));
        return MTM_OK;
    }
    const int parts = pool->size() + 1;
    const uint8_t* s = static_cast<const uint8_t*>(src);
    uint8_t* d = static_cast<uint8_t*>(dst);
    for (int r0 = 0, c = 0; r0 < H; r0 += rows_per_chunk, ++c) {
        const int k = ctx->chunk_next;
        ctx->chunk_next = (k + 1) % STAGE_BUFS;
        const int rows = std::min(rows_per_chunk, H - r0);
        if (ctx->chunk_used[k]) MTM_CUDA(ctx, cudaEventSynchronize(ctx->ev_chunk[k]));     // its last DMA has left the buffer
        uint8_t* stage = ctx->h_chunk[k];
        const int per = (rows + parts - 1) / parts;
        pool->parallel_for(parts, [&](int t) {
            const int a = std::min(rows, t * per), b = std::min(rows, a + per);
            if (src_stride == row_bytes) {
                if (b > a) memcpy(stage + (size_t)a * row_bytes, s + (size_t)(r0 + a) * src_stride, (size_t)(b - a) * row_bytes);
            } else {
                for (int r = a; r < b; ++r) memcpy(stage + (size_t)r * row_bytes, s + (size_t)(r0 + r) * src_stride, row_bytes);
            }
        });
        MTM_CUDA(ctx, cudaMemcpy2DAsync(d + (size_t)r0 * dst_pitch, dst_pitch, stage, row_bytes, row_bytes, (size_t)rows,
                                        cudaMemcpyHostToDevice, ctx->stream));
        MTM_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[k], ctx->stream));
        ctx->chunk_used[k] = true;
    }
    return MTM_OK;
}
