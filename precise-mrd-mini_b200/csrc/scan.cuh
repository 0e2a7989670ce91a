// scan.cuh -- device-wide exclusive prefix sum (reduce / recurse / scan-with-offset).
// Deadlock-free by construction (no inter-block waiting), deterministic.
#pragma once
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* smem_warp /*[32]*/, T& block_total) {
    // v: per-thread value; returns exclusive prefix over the block
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) smem_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        T w = (lane < (uint32_t)nw) ? smem_warp[lane] : T(0);
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (uint32_t)o) wi += t;
        }
        smem_warp[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) smem_warp[32] = wi;
    }
    __syncthreads();
    block_total = smem_warp[32];
    T r = smem_warp[warp] + incl - v;
    __syncthreads();
    return r;
}

template <typename TIn, typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const TIn* __restrict__ in, T* __restrict__ sums, int64_t n) {
    __shared__ T sw[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    T acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i * SCAN_THREADS + threadIdx.x;
        if (idx < n) acc += (T)in[idx];
    }
    T tot;
    block_exclusive_scan<T>(acc, sw, tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

template <typename TIn, typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const TIn* __restrict__ in, T* __restrict__ out, const T* __restrict__ block_offsets, int64_t n, T* __restrict__ total_out) {
    __shared__ T sw[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i;
        v[i] = (idx < n) ? (T)in[idx] : T(0);
        acc += v[i];
    }
    T tot;
    T excl = block_exclusive_scan<T>(acc, sw, tot);
    T off = (block_offsets ? block_offsets[blockIdx.x] : T(0)) + excl;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i;
        if (idx < n) out[idx] = off;
        off += v[i];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total_out = off;
}

// out[i] = sum_{j<i} in[j]; in == out allowed when TIn == T.  total_out (device, optional)
// receives the grand total.
template <typename TIn, typename T>
static int32_t exclusive_scan(pmrd_ctx* ctx, const TIn* in, T* out, int64_t n, T* total_out) {
    if (n <= 0) {
        if (total_out) PMRD_CUDA(ctx, cudaMemsetAsync(total_out, 0, sizeof(T), ctx->stream));
        return PMRD_OK;
    }
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        PMRD_LAUNCH(ctx, (scan_apply_kernel<TIn, T>), 1, SCAN_THREADS, 0, in, out, (const T*)nullptr, n, total_out);
        return PMRD_OK;
    }
    DevBuf sums;
    PMRD_CUDA(ctx, sums.alloc(ctx->stream, (size_t)nb * sizeof(T)));
    PMRD_LAUNCH(ctx, (scan_reduce_kernel<TIn, T>), (int)nb, SCAN_THREADS, 0, in, sums.as<T>(), n);
    PMRD_TRY((exclusive_scan<T, T>(ctx, sums.as<T>(), sums.as<T>(), nb, (T*)nullptr)));
    PMRD_LAUNCH(ctx, (scan_apply_kernel<TIn, T>), (int)nb, SCAN_THREADS, 0, in, out, (const T*)sums.as<T>(), n, total_out);
    return PMRD_OK;
}
