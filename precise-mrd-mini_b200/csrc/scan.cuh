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

// Layout of the apply pass: a warp owns 512 consecutive items as 4 groups of 128; lane l holds the 4
// consecutive items [group * 128 + 4 l, +4) of each group, so a 4-byte input is one 16-byte load per
// group and the outputs of a warp form contiguous 512-byte (u32) / 1-KB (u64) runs.
template <typename TIn, typename T>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const TIn* __restrict__ in, T* __restrict__ out, const T* __restrict__ block_offsets, int64_t n, T* __restrict__ total_out) {
    __shared__ T s_warp[SCAN_THREADS / 32 + 1];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)warp * (SCAN_TILE / (SCAN_THREADS / 32));
    T v[4][4], s[4], incl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t idx = wbase + j * 128 + lane * 4;
        if (sizeof(TIn) == 4 && idx + 3 < n) {
            const uint4 q = *reinterpret_cast<const uint4*>(in + idx);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { TIn t; memcpy(&t, &w[k], sizeof(TIn) < 4 ? sizeof(TIn) : 4); v[j][k] = (T)t; }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[j][k] = (idx + k < n) ? (T)in[idx + k] : T(0);
        }
        s[j] = v[j][0] + v[j][1] + v[j][2] + v[j][3];
        incl[j] = s[j];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const T t = __shfl_up_sync(0xffffffffu, incl[j], o);
            if (lane >= (uint32_t)o) incl[j] += t;
        }
    }
    T carry[4], wtot = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        carry[j] = wtot;
        wtot += __shfl_sync(0xffffffffu, incl[j], 31);
    }
    if (lane == 0) s_warp[warp] = wtot;
    __syncthreads();
    if (threadIdx.x == 0) {
        T run = block_offsets ? block_offsets[blockIdx.x] : T(0);
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) { const T c = s_warp[w]; s_warp[w] = run; run += c; }
        s_warp[SCAN_THREADS / 32] = run;
    }
    __syncthreads();
    const T wexcl = s_warp[warp];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t idx = wbase + j * 128 + lane * 4;
        T off = wexcl + carry[j] + incl[j] - s[j];
        T o4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { o4[k] = off; off += v[j][k]; }
        if (idx + 3 < n && sizeof(T) == 8) {
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + idx);
            dst[0] = make_ulonglong2((unsigned long long)o4[0], (unsigned long long)o4[1]);
            dst[1] = make_ulonglong2((unsigned long long)o4[2], (unsigned long long)o4[3]);
        } else if (idx + 3 < n && sizeof(T) == 4) {
            *reinterpret_cast<uint4*>(out + idx) = make_uint4((uint32_t)o4[0], (uint32_t)o4[1], (uint32_t)o4[2], (uint32_t)o4[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (idx + k < n) out[idx + k] = o4[k];
        }
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = s_warp[SCAN_THREADS / 32];
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
