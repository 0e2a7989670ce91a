// staging.cu -- host <-> device copies of LARGE pageable buffers through a ring of pinned staging buffers.
//
// The h_* entry points of include/pmrd.h take ordinary (pageable) host arrays -- NumPy arrays, as the reference's
// Python API hands them over.  cudaMemcpyAsync from pageable memory is a synchronous, single-threaded staging copy
// inside the driver (about 3 GB/s measured for the 600 MB of a 5e7-read collapse call): the PCIe link idles.  Here
// a small pool of host threads copies chunk i+1 into one pinned buffer while the DMA engine moves chunk i out of
// another, so the call runs at the host-memcpy / PCIe rate.  Buffers that are already pinned are copied directly.
#include "common.cuh"
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include <stdlib.h>

#define STG_BUFS 4
#define STG_CHUNK ((size_t)16 << 20)
#define STG_MIN ((size_t)4 << 20)

struct CopyPool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    char* dst = nullptr;
    const char* src = nullptr;
    size_t bytes = 0;
    uint64_t epoch = 0;
    int pending = 0;
    bool quit = false;
    explicit CopyPool(int n) {
        for (int t = 0; t < n; ++t) workers.emplace_back([this, t, n] { run(t, n); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> l(mu);
            quit = true;
        }
        cv_go.notify_all();
        for (auto& w : workers) w.join();
    }
    void run(int t, int n) {
        uint64_t seen = 0;
        for (;;) {
            char* d;
            const char* s;
            size_t b;
            {
                std::unique_lock<std::mutex> l(mu);
                cv_go.wait(l, [&] { return quit || epoch != seen; });
                if (quit) return;
                seen = epoch;
                d = dst; s = src; b = bytes;
            }
            // slice t of n, 4 KB aligned
            const size_t per = ((b / (size_t)n) + 4095) & ~(size_t)4095;
            const size_t lo = per * (size_t)t, hi = lo + per < b ? lo + per : b;
            if (lo < b) memcpy(d + lo, s + lo, hi - lo);
            {
                std::lock_guard<std::mutex> l(mu);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    void copy(void* d, const void* s, size_t b) {
        std::unique_lock<std::mutex> l(mu);
        dst = (char*)d; src = (const char*)s; bytes = b;
        pending = (int)workers.size();
        ++epoch;
        cv_go.notify_all();
        cv_done.wait(l, [&] { return pending == 0; });
    }
};

struct Stager {
    void* pin[STG_BUFS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[STG_BUFS];
    bool ok = false;
    CopyPool* pool = nullptr;
    ~Stager() {
        delete pool;
        for (int b = 0; b < STG_BUFS; ++b) {
            if (pin[b]) { cudaFreeHost(pin[b]); cudaEventDestroy(ev[b]); }
        }
    }
};

static Stager* stager_get(pmrd_ctx* ctx) {
    if (ctx->stager) return (Stager*)ctx->stager;
    Stager* s = new (std::nothrow) Stager();
    if (!s) return nullptr;
    s->ok = true;
    for (int b = 0; b < STG_BUFS; ++b) {
        if (cudaHostAlloc(&s->pin[b], STG_CHUNK, cudaHostAllocDefault) != cudaSuccess) { s->pin[b] = nullptr; s->ok = false; break; }
        if (cudaEventCreateWithFlags(&s->ev[b], cudaEventDisableTiming) != cudaSuccess) { cudaFreeHost(s->pin[b]); s->pin[b] = nullptr; s->ok = false; break; }
    }
    int n = 0;
    if (const char* e = getenv("PMRD_COPY_THREADS")) n = atoi(e);
    if (n <= 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        n = (int)(hw / 2);
        if (n > 8) n = 8;
        if (n < 2) n = 2;
    }
    s->pool = new (std::nothrow) CopyPool(n);
    if (!s->pool) s->ok = false;
    cudaGetLastError();
    ctx->stager = s;
    return s;
}

void pmrd_stager_destroy(pmrd_ctx* ctx) {
    delete (Stager*)ctx->stager;
    ctx->stager = nullptr;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// host -> device, ordered on ctx->stream.  Returns when the host buffer has been read completely (the last chunks
// may still be in flight from the pinned ring) -- the same contract as cudaMemcpyAsync from pageable memory.
int32_t pmrd_stage_up(pmrd_ctx* ctx, void* d, const void* h, size_t bytes) {
    if (bytes == 0) return PMRD_OK;
    Stager* s = (bytes >= STG_MIN && !is_pinned(h)) ? stager_get(ctx) : nullptr;
    if (!s || !s->ok) {
        PMRD_CUDA(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return PMRD_OK;
    }
    int b = 0;
    for (size_t off = 0; off < bytes; off += STG_CHUNK, b = (b + 1) % STG_BUFS) {
        const size_t len = bytes - off < STG_CHUNK ? bytes - off : STG_CHUNK;
        PMRD_CUDA(ctx, cudaEventSynchronize(s->ev[b]));                  // the DMA that last used this buffer is done
        s->pool->copy(s->pin[b], (const char*)h + off, len);
        PMRD_CUDA(ctx, cudaMemcpyAsync((char*)d + off, s->pin[b], len, cudaMemcpyHostToDevice, ctx->stream));
        PMRD_CUDA(ctx, cudaEventRecord(s->ev[b], ctx->stream));
    }
    return PMRD_OK;
}

// device -> host, ordered on ctx->stream; SYNCHRONOUS for large pageable destinations (the host threads copy out)
int32_t pmrd_stage_down(pmrd_ctx* ctx, void* h, const void* d, size_t bytes) {
    if (bytes == 0 || !h) return PMRD_OK;
    Stager* s = (bytes >= STG_MIN && !is_pinned(h)) ? stager_get(ctx) : nullptr;
    if (!s || !s->ok) {
        PMRD_CUDA(ctx, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return PMRD_OK;
    }
    const size_t n_chunks = (bytes + STG_CHUNK - 1) / STG_CHUNK;
    // keep STG_BUFS - 1 DMAs in flight ahead of the chunk being copied out
    size_t issued = 0;
    auto issue = [&](size_t c) -> cudaError_t {
        const int b = (int)(c % STG_BUFS);
        const size_t off = c * STG_CHUNK, len = bytes - off < STG_CHUNK ? bytes - off : STG_CHUNK;
        cudaError_t e = cudaMemcpyAsync(s->pin[b], (const char*)d + off, len, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(s->ev[b], ctx->stream);
        return e;
    };
    for (int b = 0; b < STG_BUFS; ++b) PMRD_CUDA(ctx, cudaEventSynchronize(s->ev[b]));   // earlier uploads have left the ring
    for (; issued < n_chunks && issued < (size_t)(STG_BUFS - 1); ++issued) PMRD_CUDA(ctx, issue(issued));
    for (size_t c = 0; c < n_chunks; ++c) {
        const int b = (int)(c % STG_BUFS);
        const size_t off = c * STG_CHUNK, len = bytes - off < STG_CHUNK ? bytes - off : STG_CHUNK;
        if (issued < n_chunks) { PMRD_CUDA(ctx, issue(issued)); ++issued; }       // (its buffer was drained one iteration ago)
        PMRD_CUDA(ctx, cudaEventSynchronize(s->ev[b]));
        s->pool->copy((char*)h + off, s->pin[b], len);
    }
    return PMRD_OK;
}
