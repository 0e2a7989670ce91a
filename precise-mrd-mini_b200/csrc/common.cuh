// common.cuh -- context, error plumbing, launch accounting, Philox4x32-10, warp helpers.
// sm_100a only: no multi-arch dispatch, no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pmrd.h"

#define PMRD_SM_COUNT_DEFAULT 148

struct pmrd_ctx {
    int device;
    int sm_count;
    uint64_t seed;
    cudaStream_t own_stream;
    cudaStream_t stream;
    int64_t launches;
    int prof_on;   // per-kernel CUDA-event timing (pmrd_profile_enable)
    void* prof;    // opaque: api.cu
    void* stager;  // opaque: staging.cu (pinned ring + copy threads, created on the first large pageable copy)
    char err[512];
};

int32_t pmrd_stage_up(pmrd_ctx* ctx, void* d, const void* h, size_t bytes);
int32_t pmrd_stage_down(pmrd_ctx* ctx, void* h, const void* d, size_t bytes);
void pmrd_stager_destroy(pmrd_ctx* ctx);
void pmrd_prof_begin(pmrd_ctx* ctx, const char* kernel_name);
void pmrd_prof_end(pmrd_ctx* ctx);

extern char g_pmrd_create_err[512];

#define PMRD_CUDA(ctx, expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__,   \
                     #expr, cudaGetErrorString(_e));                                          \
            return PMRD_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define PMRD_ARG(ctx, cond, msg)                                                              \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: invalid argument: %s", __FILE__, \
                     __LINE__, msg);                                                          \
            return PMRD_ERR_ARG;                                                              \
        }                                                                                     \
    } while (0)

#define PMRD_TRY(expr)                 \
    do {                               \
        int32_t _s = (expr);           \
        if (_s != PMRD_OK) return _s;  \
    } while (0)

// launch + count + check
#define PMRD_LAUNCH_NAMED(ctx, name, kernel, grid, block, smem, ...)                          \
    do {                                                                                      \
        if ((ctx)->prof_on) pmrd_prof_begin(ctx, name);                                       \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                      \
        if ((ctx)->prof_on) pmrd_prof_end(ctx);                                               \
        (ctx)->launches++;                                                                    \
        PMRD_CUDA(ctx, cudaGetLastError());                                                   \
    } while (0)
#define PMRD_LAUNCH(ctx, kernel, grid, block, smem, ...) PMRD_LAUNCH_NAMED(ctx, #kernel, kernel, grid, block, smem, __VA_ARGS__)
// (arguments that are themselves macros holding a comma list are expanded first)
#define PMRD_LAUNCH_NAMED_X(...) PMRD_LAUNCH_NAMED(__VA_ARGS__)

// stream-ordered scratch buffer; freed on scope exit (cudaFreeAsync on the same stream)
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(cudaStream_t stream, size_t bytes) {
        release();
        s = stream;
        if (bytes == 0) bytes = 16;
        return cudaMallocAsync(&p, bytes, stream);
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

static inline int pmrd_blocks(int64_t n, int per_block) { return (int)((n + per_block - 1) / per_block); }

// ----------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter-based: (key = seed, counter = who/what).
// ----------------------------------------------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
        lo = a * b;
        hi = __umulhi(a, b);
#else
        uint64_t p = (uint64_t)a * b;
        lo = (uint32_t)p;
        hi = (uint32_t)(p >> 32);
#endif
    }
    __host__ __device__ inline uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t key0 = k0, key1 = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ key0;
            uint32_t n1 = lo1;
            uint32_t n2 = hi0 ^ c3 ^ key1;
            uint32_t n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            key0 += 0x9E3779B9u;
            key1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// Philox2x32-10 (same paper): 64-bit counter, 32-bit key, one multiply per round.  Used for the
// per-read draws, whose key is a per-family word drawn from the Philox4x32 family block -- half the
// multiplies of a 4x32 block for the 64 random bits a read needs.
__host__ __device__ static inline uint2 philox2x32(uint32_t key, uint32_t c0, uint32_t c1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi, lo;
        Philox::mulhilo(0xD256D193u, c0, hi, lo);
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += 0x9E3779B9u;
    }
    return make_uint2(c0, c1);
}

// stream ids: the 4th counter word's top byte names the stage so that streams never overlap
#define PMRD_STREAM_CELLS 1u     /* simulate_reads per-cell draws            */
#define PMRD_STREAM_GD_FAMILY 2u /* G-D collapse_umis per-family draws       */
#define PMRD_STREAM_READS 3u     /* read-level generator                     */
#define PMRD_STREAM_ERRMODEL 4u  /* error-model modifiers + bootstrap        */
#define PMRD_STREAM_LODBOOT 5u   /* LoD bootstrap resampling                 */
#define PMRD_STREAM_SHUFFLE 6u   /* tile shuffle                             */
#define PMRD_STREAM_CONTAM 7u    /* hop / collision injection                */

// uniform in (0,1): 53 random bits, never 0 or 1
__host__ __device__ static inline double u53(uint32_t hi, uint32_t lo) {
    uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;  // 53 bits
    return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}
// uniform in (0,1) from 32 bits
__host__ __device__ static inline float u24(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// ----------------------------------------------------------------------------------------
// warp helpers
// ----------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// streaming 128-bit load / store (read-once data: do not pollute L1)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u1(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
#endif
