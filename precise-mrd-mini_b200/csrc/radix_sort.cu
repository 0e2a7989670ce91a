// radix_sort.cu -- kernel K4 (see radix_sort.cuh).  Replaces the reference's dict-of-lists
// grouping: src/precise_mrd/umi.py:97-144 (group_umis_by_site / exact-UMI groups),
// precise-mrd-mini/src/mrd/umi.py:65-67 (groupby), rust_ext/src/lib.rs:7-22 (FNV map).
#include "radix_sort.cuh"
#include <stdlib.h>
#ifndef RS_DIRECT_MIN_BLOCKS
#define RS_DIRECT_MIN_BLOCKS 3
#endif
#ifndef RS_WIDE_MIN_BLOCKS        /* blocks per SM the 256 x 16 scatter kernels are compiled for (register cap = 65536 / (256 x this)) */
#define RS_WIDE_MIN_BLOCKS 4
#endif
#ifndef RS_SWIDE_MIN_BLOCKS
#define RS_SWIDE_MIN_BLOCKS 4
#endif

// zero `words` (a multiple of 4 * THREADS) 32-bit shared-memory counters with 16-byte stores, fully unrolled
// (the scalar loop was 9 % of the direct downsweep's instructions)
template <int THREADS, int WORDS>
__device__ __forceinline__ void rs_zero_smem(uint32_t* p) {
    static_assert(WORDS % (4 * THREADS) == 0, "whole 16-byte stores per thread");
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < WORDS / (4 * THREADS); ++i) q[i * THREADS + threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
}

// ---------------------------------------------------------------- varying bits ----------
__global__ void __launch_bounds__(256) rs_varying_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                         unsigned long long* __restrict__ or_and /*[2]*/) {
    uint64_t o = 0, a = ~0ull;
    const int64_t n2 = n >> 1;
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v = ld_stream_u4(k4 + i);
        uint64_t x = ((uint64_t)v.y << 32) | v.x, y = ((uint64_t)v.w << 32) | v.z;
        o |= x | y;
        a &= x & y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        uint64_t x = keys[n - 1];
        o |= x;
        a &= x;
    }
    uint32_t olo = __reduce_or_sync(0xffffffffu, (uint32_t)o), ohi = __reduce_or_sync(0xffffffffu, (uint32_t)(o >> 32));
    uint32_t alo = __reduce_and_sync(0xffffffffu, (uint32_t)a), ahi = __reduce_and_sync(0xffffffffu, (uint32_t)(a >> 32));
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&or_and[0], ((unsigned long long)ohi << 32) | olo);
        atomicAnd(&or_and[1], ((unsigned long long)ahi << 32) | alo);
    }
}

int32_t radix_varying_bits(pmrd_ctx* ctx, const uint64_t* d_keys, int64_t n, uint64_t* h_varying, uint64_t* h_or) {
    *h_varying = 0;
    if (h_or) *h_or = 0;
    if (n <= 0) return PMRD_OK;
    if (n == 1 && !h_or) return PMRD_OK;
    DevBuf buf;
    PMRD_CUDA(ctx, buf.alloc(ctx->stream, 16));
    unsigned long long init[2] = {0ull, ~0ull};
    PMRD_CUDA(ctx, cudaMemcpyAsync(buf.p, init, 16, cudaMemcpyHostToDevice, ctx->stream));
    int grid = (int)((n / 2 + 255) / 256);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    if (grid < 1) grid = 1;
    PMRD_LAUNCH(ctx, rs_varying_kernel, grid, 256, 0, d_keys, n, buf.as<unsigned long long>());
    unsigned long long res[2];
    PMRD_CUDA(ctx, cudaMemcpyAsync(res, buf.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *h_varying = res[0] ^ res[1];
    if (h_or) *h_or = res[0];
    return PMRD_OK;
}

RadixPlan radix_make_plan(uint64_t varying, int max_digit_bits) {
    RadixPlan p;
    p.n_passes = 0;
    // greedy: an (up to) 8-bit window starting at the lowest varying bit not yet covered; at most 8 windows cover 64 bits.
    // max_digit_bits = 9: if 9-bit windows save a pass, the varying span is cut into equal windows of at most 9 bits.
    int lo = 0, hi = 63, nbits = 0;
    while (lo < 64 && !((varying >> lo) & 1ull)) ++lo;
    while (hi >= 0 && !((varying >> hi) & 1ull)) --hi;
    for (int b = 0; b < 64; ++b) nbits += (int)((varying >> b) & 1ull);
    if (max_digit_bits >= 9 && lo <= hi && nbits == hi - lo + 1) {            // one contiguous run of varying bits
        const int span = hi - lo + 1, p8 = (span + 7) / 8, p9 = (span + 8) / 9;
        if (p9 < p8) {
            int b = lo;
            for (int k = 0; k < p9; ++k) {
                const int w = (span - (b - lo) + (p9 - k) - 1) / (p9 - k);    // balanced: 17 bits -> 9 + 8
                p.shift[p.n_passes] = b;
                p.bits[p.n_passes] = w;
                ++p.n_passes;
                b += w;
            }
            return p;
        }
    }
    int b = 0;
    while (b < 64) {
        if (!((varying >> b) & 1ull)) { ++b; continue; }
        int w = 64 - b < 8 ? 64 - b : 8;
        // trim non-varying bits off the top of the window
        while (w > 1 && !((varying >> (b + w - 1)) & 1ull)) --w;
        p.shift[p.n_passes] = b;
        p.bits[p.n_passes] = w;
        ++p.n_passes;
        b += w;
    }
    return p;
}

// ---------------------------------------------------------------- upsweep ---------------
// hist[tile][BINS] (tile-major: one coalesced row per tile).  BINS = 256 (8-bit digits) or 512 (9-bit digits: a 17-bit
// group index is partitioned in two passes instead of three)
template <int BINS>
__global__ void __launch_bounds__(RS_THREADS)
rs_upsweep_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ hist, int shift, uint32_t mask) {
    extern __shared__ __align__(16) uint32_t s_hist_raw[];
    uint32_t (*s_hist)[BINS] = reinterpret_cast<uint32_t (*)[BINS]>(s_hist_raw);          // [RS_WARPS][BINS]
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    rs_zero_smem<RS_THREADS, RS_WARPS * BINS>(&s_hist[0][0]);
    __syncthreads();
    const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
    const int64_t rem = n - tile_base;
    const int tile_n = rem < RS_TILE ? (int)rem : RS_TILE;
    // each warp owns 512 consecutive keys; 128-bit loads = 2 keys per lane per step
    const int64_t wbase = tile_base + (int64_t)warp * (RS_TILE / RS_WARPS);
#pragma unroll
    for (int i = 0; i < RS_ITEMS / 2; ++i) {
        const int loc = (int)warp * (RS_TILE / RS_WARPS) + i * 64 + (int)lane * 2;
        uint32_t d0 = 0xffffffffu, d1 = 0xffffffffu;
        if (loc + 1 < tile_n) {
            uint4 v = ld_stream_u4(keys + wbase + i * 64 + lane * 2);
            uint64_t x = ((uint64_t)v.y << 32) | v.x, y = ((uint64_t)v.w << 32) | v.z;
            d0 = (uint32_t)(x >> shift) & mask;
            d1 = (uint32_t)(y >> shift) & mask;
        } else if (loc < tile_n) {
            d0 = (uint32_t)(keys[wbase + i * 64 + lane * 2] >> shift) & mask;
        }
        // one shared-memory atomic per key on a warp-private histogram: digits of a scrambled
        // input are spread over the bins, so same-address conflicts inside a warp are rare
        if (d0 != 0xffffffffu) atomicAdd(&s_hist[warp][d0], 1u);
        if (d1 != 0xffffffffu) atomicAdd(&s_hist[warp][d1], 1u);
    }
    __syncthreads();
    if (threadIdx.x < BINS) {
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) c += s_hist[w][threadIdx.x];
        hist[(int64_t)blockIdx.x * BINS + threadIdx.x] = c;
    }
}

// Whole tiles only (the cell-major layout), digit inside the low 32 key bits: every load of the tile is issued before
// the first counter is touched (8 x 16 bytes in flight per thread at 256 threads), no bounds tests.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
rs_upsweep_tiles_kernel(const uint64_t* __restrict__ keys, uint32_t* __restrict__ hist, int shift, uint32_t mask) {
    constexpr int WARPS = THREADS / 32, LOADS = RS_TILE / 2 / THREADS;
    __shared__ __align__(16) uint32_t s_hist[WARPS][RS_MAX_BINS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(keys + (int64_t)blockIdx.x * RS_TILE) + warp * (LOADS * 32) + lane;
    uint4 v[LOADS];
#pragma unroll
    for (int i = 0; i < LOADS; ++i) v[i] = ld_stream_u4(src + i * 32);
    rs_zero_smem<THREADS, WARPS * RS_MAX_BINS>(&s_hist[0][0]);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < LOADS; ++i) {
        atomicAdd(&s_hist[warp][(v[i].x >> shift) & mask], 1u);
        atomicAdd(&s_hist[warp][(v[i].z >> shift) & mask], 1u);
    }
    __syncthreads();
    if (threadIdx.x < RS_MAX_BINS) {
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) c += s_hist[w][threadIdx.x];
        hist[(int64_t)blockIdx.x * RS_MAX_BINS + threadIdx.x] = c;
    }
}

// ---------------------------------------------------------------- column scan -----------
// H[tile][d]  ->  exclusive over (d major, tile minor) order, computed column-wise.
// (the three column kernels run with one thread per bin: blockDim.x = 256 or 512)
__global__ void __launch_bounds__(512)
rs_colsum_kernel(const uint32_t* __restrict__ hist, int n_tiles, int chunk, uint32_t* __restrict__ colsum /*[n_chunks][bins]*/) {
    const int bins = blockDim.x;
    const int t0 = blockIdx.x * chunk;
    const int t1 = min(t0 + chunk, n_tiles);
    uint32_t acc = 0;
#pragma unroll 8
    for (int t = t0; t < t1; ++t) acc += hist[(int64_t)t * bins + threadIdx.x];
    colsum[(int64_t)blockIdx.x * bins + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(512) rs_colbase_kernel(uint32_t* __restrict__ colsum, int n_chunks) {
    __shared__ uint32_t sw[33];
    const int bins = blockDim.x;
    uint32_t run = 0;
#pragma unroll 8
    for (int c = 0; c < n_chunks; ++c) {
        uint32_t v = colsum[(int64_t)c * bins + threadIdx.x];
        colsum[(int64_t)c * bins + threadIdx.x] = run;
        run += v;
    }
    uint32_t tot;
    uint32_t base = block_exclusive_scan<uint32_t>(run, sw, tot);
    for (int c = 0; c < n_chunks; ++c) colsum[(int64_t)c * bins + threadIdx.x] += base;
}

__global__ void __launch_bounds__(512)
rs_colapply_kernel(uint32_t* __restrict__ hist, int n_tiles, int chunk, const uint32_t* __restrict__ colsum) {
    const int bins = blockDim.x;
    const int t0 = blockIdx.x * chunk;
    const int t1 = min(t0 + chunk, n_tiles);
    uint32_t run = colsum[(int64_t)blockIdx.x * bins + threadIdx.x];
    int t = t0;
    for (; t + 8 <= t1; t += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = hist[(int64_t)(t + j) * bins + threadIdx.x];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            hist[(int64_t)(t + j) * bins + threadIdx.x] = run;
            run += v[j];
        }
    }
    for (; t < t1; ++t) {
        uint32_t v = hist[(int64_t)t * bins + threadIdx.x];
        hist[(int64_t)t * bins + threadIdx.x] = run;
        run += v;
    }
}

// ---------------------------------------------------------------- downsweep -------------
template <int BINS>
struct RsSmemT {
    uint64_t keys[RS_TILE];
    uint32_t vals[RS_TILE];
    uint32_t warp_cnt[RS_WARPS][BINS];
    uint32_t bin_base[BINS];
    uint32_t glob_delta[BINS];
    uint32_t sw[33];
};
typedef RsSmemT<RS_MAX_BINS> RsSmem;

template <bool HAS_VALS, int BINS = RS_MAX_BINS>
__global__ void __launch_bounds__(RS_THREADS, BINS == 256 ? 3 : 2)
rs_downsweep_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                    uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                    const uint32_t* __restrict__ hist_scanned, int64_t n, int shift, uint32_t mask) {
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmemT<BINS>& S = *reinterpret_cast<RsSmemT<BINS>*>(rs_smem_raw);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    rs_zero_smem<RS_THREADS, RS_WARPS * BINS>(&S.warp_cnt[0][0]);
    const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
    const int64_t rem = n - tile_base;
    const int tile_n = rem < RS_TILE ? (int)rem : RS_TILE;
    const int wloc = (int)warp * (RS_TILE / RS_WARPS);

    // warp-striped load: item i of lane l = wloc + i*32 + l  (memory order = i-major, lane-minor)
    uint64_t key[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int loc = wloc + i * 32 + (int)lane;
        key[i] = (loc < tile_n) ? keys_in[tile_base + loc] : ~0ull;
    }
    __syncthreads();

    // stable in-warp ranking per digit (match-any): rank = #earlier items of this warp with my digit
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int loc = wloc + i * 32 + (int)lane;
        const bool valid = loc < tile_n;
        const uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & mask) : 0xffffffffu;
        // peers = lanes holding the same digit (lanes past the end of the input form their own group)
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        const uint32_t m = rs_peer_mask<BINS == 256 ? 8 : 9>(d, valid ? vmask : ~vmask);
        const uint32_t before = __popc(m & lt);
        // the lowest lane of each digit group bumps the warp's counter once for the whole group and
        // hands the old value to its peers by shuffle: one shared atomic + one shuffle per item,
        // no read-modify-write hazard and therefore no warp barriers in the loop
        const int leader = __ffs((int)m) - 1;
        uint32_t old = 0;
        if (valid && before == 0) old = atomicAdd(&S.warp_cnt[warp][d], (uint32_t)__popc(m));
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = (uint16_t)(old + before);
    }
    __syncthreads();

    // per-bin: exclusive prefix over warps, then block scan over bins
    {
        const uint32_t b = threadIdx.x;  // threads 0..BINS-1 own one bin each (BINS <= RS_THREADS)
        uint32_t run = 0;
        if (b < BINS) {
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                uint32_t c = S.warp_cnt[w][b];
                S.warp_cnt[w][b] = run;
                run += c;
            }
        }
        uint32_t tot;
        uint32_t base = block_exclusive_scan<uint32_t>(run, S.sw, tot);
        if (b < BINS) {
            S.bin_base[b] = base;
            S.glob_delta[b] = hist_scanned[(int64_t)blockIdx.x * BINS + b] - base;
        }
    }
    __syncthreads();

    // reorder into shared memory by digit (stable)
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int loc = wloc + i * 32 + (int)lane;
        if (loc < tile_n) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            const uint32_t pos = S.bin_base[d] + S.warp_cnt[warp][d] + rank[i];
            S.keys[pos] = key[i];
            if (HAS_VALS) S.vals[pos] = vals_in[tile_base + loc];
        }
    }
    __syncthreads();

    // coalesced per-digit run stores
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int j = i * RS_THREADS + (int)threadIdx.x;
        if (j < tile_n) {
            const uint64_t k = S.keys[j];
            const uint32_t d = (uint32_t)(k >> shift) & mask;
            const uint32_t g = (uint32_t)j + S.glob_delta[d];
            keys_out[g] = k;
            if (HAS_VALS) vals_out[g] = S.vals[j];
        }
    }
}

// ---------------------------------------------------------------- direct downsweep ------
// Keys-only variant for cell-major input: ranks are computed exactly as above, but every record is
// stored straight from registers to its final position.  A cell's output region is at most a
// few MB and its tiles are processed back to back, so the partially written 32-byte sectors meet
// in L2 before they are evicted (DRAM write traffic stays at 8 B / record -- checked with ncu),
// and the shared-memory reorder (a bank-conflicted 64-bit scatter + a gather + a barrier) is gone.
// (whole tiles only: the cell-major layout pads every cell to a multiple of RS_TILE)
template <int THREADS, int ITEMS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
rs_downsweep_direct_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                           const uint32_t* __restrict__ hist_scanned, int shift, uint32_t mask) {
    static_assert(THREADS * ITEMS == RS_TILE, "one 4096-record tile per block");
    constexpr int WARPS = THREADS / 32;
    __shared__ __align__(16) uint32_t s_cnt[WARPS][RS_MAX_BINS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    rs_zero_smem<THREADS, WARPS * RS_MAX_BINS>(&s_cnt[0][0]);
    const uint64_t* __restrict__ src = keys_in + (int64_t)blockIdx.x * RS_TILE + warp * (RS_TILE / WARPS) + lane;
    uint64_t key[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) key[i] = src[i * 32];
    __syncthreads();
    uint32_t rank2[ITEMS / 2];   // two 16-bit ranks per register
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        const uint32_t m = rs_peer_mask(d, 0xffffffffu);
        const uint32_t before = __popc(m & lt);
        const int leader = __ffs((int)m) - 1;
        uint32_t old = 0;
        if (before == 0) old = atomicAdd(&s_cnt[warp][d], (uint32_t)__popc(m));
        old = __shfl_sync(0xffffffffu, old, leader);
        const uint32_t r = old + before;
        if (i & 1) rank2[i >> 1] |= r << 16;
        else rank2[i >> 1] = r;
    }
    __syncthreads();
    // per bin: global start of this tile's run + exclusive prefix over the warps
    if (threadIdx.x < RS_MAX_BINS) {
        uint32_t run = hist_scanned[(int64_t)blockIdx.x * RS_MAX_BINS + threadIdx.x];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_cnt[w][threadIdx.x];
            s_cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        const uint32_t r = (i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu;
        keys_out[s_cnt[warp][d] + r] = key[i];
    }
}

// ---------------------------------------------------------------- staged downsweep ------
// Same ranking, but the tile is first put in digit order in shared memory and then written out
// position by position, so that a warp's store covers a few contiguous digit runs instead of 32
// scattered records.  Worth its shared-memory round trip only when the INPUT is scattered: on the
// first pass over generator output a direct warp store touches 30.5 distinct 32-byte sectors
// (the kernel is bound by that L1->L2 sector rate), here ~10; on later passes the first pass has
// already clustered each family's reads and the direct kernel is the faster one.
template <int WARPS>
struct RsStage {
    uint32_t cnt[WARPS][RS_MAX_BINS];
    uint64_t rec[RS_TILE];
    uint32_t delta[RS_MAX_BINS];
    uint32_t sw[33];
};
template <int THREADS, int ITEMS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
rs_downsweep_staged_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                           const uint32_t* __restrict__ hist_scanned, int shift, uint32_t mask) {
    static_assert(THREADS * ITEMS == RS_TILE && THREADS >= RS_MAX_BINS, "one 4096-record tile per block");
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char rs_stage_raw[];
    RsStage<WARPS>& S = *reinterpret_cast<RsStage<WARPS>*>(rs_stage_raw);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    rs_zero_smem<THREADS, WARPS * RS_MAX_BINS>(&S.cnt[0][0]);
    const uint64_t* __restrict__ src = keys_in + (int64_t)blockIdx.x * RS_TILE + warp * (RS_TILE / WARPS) + lane;
    uint64_t key[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) key[i] = src[i * 32];
    __syncthreads();
    uint32_t rank2[ITEMS / 2];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        const uint32_t m = rs_peer_mask(d, 0xffffffffu);
        const uint32_t before = __popc(m & lt);
        const int leader = __ffs((int)m) - 1;
        uint32_t old = 0;
        if (before == 0) old = atomicAdd(&S.cnt[warp][d], (uint32_t)__popc(m));
        old = __shfl_sync(0xffffffffu, old, leader);
        const uint32_t r = old + before;
        if (i & 1) rank2[i >> 1] |= r << 16;
        else rank2[i >> 1] = r;
    }
    __syncthreads();
    {
        // where each digit's run starts inside the tile: exclusive scan of the tile's own digit counts
        // (the sum of the warp counters -- no histogram of the tile has to be read for it)
        uint32_t tot = 0;
        if (threadIdx.x < RS_MAX_BINS) {
#pragma unroll
            for (int w = 0; w < WARPS; ++w) tot += S.cnt[w][threadIdx.x];
        }
        uint32_t all;
        uint32_t run = block_exclusive_scan<uint32_t>(tot, S.sw, all);
        if (threadIdx.x < RS_MAX_BINS) {
            S.delta[threadIdx.x] = hist_scanned[(int64_t)blockIdx.x * RS_MAX_BINS + threadIdx.x] - run;   // tile position -> global position
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = S.cnt[w][threadIdx.x];
                S.cnt[w][threadIdx.x] = run;
                run += c;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        const uint32_t r = (i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu;
        S.rec[S.cnt[warp][d] + r] = key[i];
    }
    __syncthreads();
    uint64_t* __restrict__ dst = keys_out;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = i * THREADS + threadIdx.x;
        const uint64_t k = S.rec[j];
        dst[j + S.delta[(uint32_t)(k >> shift) & mask]] = k;
    }
}

// ---------------------------------------------------------------- segmented scan --------
// Cell-major input (each cell = whole tiles [t0, t1)): every cell is an independent sort, so
// the scan restarts at each cell -- one block per cell, thread d owns digit d.
//   base(d, t) = t0 * RS_TILE + sum_{d' < d} sum_{t' in cell} H[t'][d'] + sum_{t0 <= t' < t} H[t'][d]
// a cell's tiles are cut into RS_SCAN_PARTS contiguous ranges, walked side by side (1: cells of a few tiles only)
template <int RS_SCAN_PARTS>
__global__ void __launch_bounds__(256 * RS_SCAN_PARTS)
rs_cellscan_kernel(uint32_t* __restrict__ hist, const int64_t* __restrict__ cell_tile_start, int n_cells) {
    __shared__ uint32_t sw[33];
    __shared__ uint32_t s_part[RS_SCAN_PARTS][256];
    const int c = blockIdx.x;
    if (c >= n_cells) return;
    const uint32_t d = threadIdx.x & 255u, part = threadIdx.x >> 8;
    const int64_t t0 = cell_tile_start[c], t1 = cell_tile_start[c + 1];
    const int64_t per = (t1 - t0 + RS_SCAN_PARTS - 1) / RS_SCAN_PARTS;
    const int64_t p0 = min(t0 + (int64_t)part * per, t1), p1 = min(p0 + per, t1);
    // the serial walk over a cell's (up to ~60) tiles is pure load latency: four ranges in flight cut it fourfold
    uint32_t acc = 0;
#pragma unroll 4
    for (int64_t t = p0; t < p1; ++t) acc += hist[t * 256 + d];
    s_part[part][d] = acc;
    __syncthreads();
    uint32_t tot_d = 0, before = 0;
#pragma unroll
    for (int p = 0; p < RS_SCAN_PARTS; ++p) {
        const uint32_t v = s_part[p][d];
        before += p < (int)part ? v : 0u;
        tot_d += v;
    }
    uint32_t tot;
    // exclusive scan over the digits of the cell's totals (the first 256 threads carry them, the others add zero)
    const uint32_t ex = block_exclusive_scan<uint32_t>(part == 0 ? tot_d : 0u, sw, tot);
    if (part == 0) s_part[0][d] = ex;
    __syncthreads();
    uint32_t run = (uint32_t)(t0 * RS_TILE) + s_part[0][d] + before;
    int64_t t = p0;
    for (; t + 4 <= p1; t += 4) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = hist[(t + j) * 256 + d];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hist[(t + j) * 256 + d] = run;
            run += v[j];
        }
    }
    for (; t < p1; ++t) {
        const uint32_t v = hist[t * 256 + d];
        hist[t * 256 + d] = run;
        run += v;
    }
}

static int32_t rs_set_attrs(pmrd_ctx* ctx) {
    // (function attributes belong to the device of the current context: set on every call, it costs microseconds)
    {
        PMRD_CUDA(ctx, cudaFuncSetAttribute(rs_downsweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        PMRD_CUDA(ctx, cudaFuncSetAttribute(rs_downsweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        PMRD_CUDA(ctx, cudaFuncSetAttribute((rs_downsweep_kernel<true, 512>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmemT<512>)));
        PMRD_CUDA(ctx, cudaFuncSetAttribute((rs_downsweep_kernel<false, 512>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmemT<512>)));
        PMRD_CUDA(ctx, cudaFuncSetAttribute(rs_upsweep_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RS_WARPS * 512 * sizeof(uint32_t))));
        PMRD_CUDA(ctx, cudaFuncSetAttribute(rs_downsweep_staged_kernel<RS_THREADS, RS_ITEMS, RS_DIRECT_MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsStage<RS_WARPS>)));
        PMRD_CUDA(ctx, cudaFuncSetAttribute(rs_downsweep_staged_kernel<256, 16, RS_SWIDE_MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsStage<8>)));
    }
    return PMRD_OK;
}

// Per-cell sort of the low `key_bits` key bits; n must be a whole number of tiles and
// d_cell_tile_start[n_cells] == n / RS_TILE.  The bits above key_bits must be constant
// inside a cell (they are: the cell index).
// d_hist0 (optional): a [n / RS_TILE][256] buffer that already holds the per-tile histogram of the FIRST
// digit (the producer of the records counted it as a by-product): the first counting pass is skipped and
// the buffer serves as the histogram scratch of the later passes.
int32_t radix_sort_pairs_segmented(pmrd_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                                   int64_t n, const int64_t* d_cell_tile_start, int64_t n_cells, int key_bits, int* in_b,
                                   uint32_t* d_hist0) {
    *in_b = 0;
    if (n <= 0 || n_cells <= 0) return PMRD_OK;
    PMRD_ARG(ctx, n % RS_TILE == 0 && n < (1ll << 32) - RS_TILE, "segmented sort: n must be a multiple of the tile size and < 2^32");
    PMRD_ARG(ctx, key_bits >= 1 && key_bits <= 32, "segmented sort: key_bits");
    PMRD_TRY(rs_set_attrs(ctx));
    const int n_tiles = (int)(n / RS_TILE);
    DevBuf hist_own;
    if (!d_hist0) PMRD_CUDA(ctx, hist_own.alloc(ctx->stream, (size_t)n_tiles * RS_MAX_BINS * sizeof(uint32_t)));
    uint32_t* const hist = d_hist0 ? d_hist0 : hist_own.as<uint32_t>();
    uint64_t* kin = keys_a; uint32_t* vin = vals_a; uint64_t* kout = keys_b; uint32_t* vout = vals_b;
    const int n_passes = (key_bits + 7) / 8;
    // development knobs, read on every call (no static state: contexts of different threads may sort at once)
    const char* e_knob = getenv("PMRD_RS_DIRECT");
    const int direct = e_knob ? atoi(e_knob) : 1;
    e_knob = getenv("PMRD_RS_STAGED");
    const int staged = e_knob ? atoi(e_knob) : 1;
    e_knob = getenv("PMRD_RS_SWIDE");
    const int swide = e_knob ? atoi(e_knob) : 1;      // 1: 256 threads x 16 records, 4 blocks / SM (8.1 vs 9.2 ms per step)
    e_knob = getenv("PMRD_RS_WIDE");
    const int wide = e_knob ? atoi(e_knob) : 2;       // 2: 256 threads x 16 records, 4 blocks / SM (measured best, by 3 %)
    for (int p = 0; p < n_passes; ++p) {
        const int shift = p * 8;
        const int w = key_bits - shift < 8 ? key_bits - shift : 8;
        const uint32_t mask = (1u << w) - 1u;
        // bit p of `staged`: pass p goes through the shared-memory staged scatter (default: the first pass only)
        const bool stage = !vals_a && direct && ((staged >> p) & 1);
        if (!(p == 0 && d_hist0)) {
            // whole tiles, digit in the low key word: the loads-first counting kernel (5.2 vs 6.6 ms per step)
            if (shift + 8 <= 32) PMRD_LAUNCH_NAMED(ctx, "rs_upsweep_kernel", (rs_upsweep_tiles_kernel<256>), n_tiles, 256, 0, kin, hist, shift, mask);
            else PMRD_LAUNCH_NAMED(ctx, "rs_upsweep_kernel", rs_upsweep_kernel<256>, n_tiles, RS_THREADS, RS_WARPS * 256 * sizeof(uint32_t), kin, n, hist, shift, mask);
        }
        if ((int64_t)n_tiles >= 8 * n_cells) PMRD_LAUNCH_NAMED(ctx, "rs_cellscan_kernel", rs_cellscan_kernel<4>, (int)n_cells, 1024, 0, hist, d_cell_tile_start, (int)n_cells);
        else PMRD_LAUNCH_NAMED(ctx, "rs_cellscan_kernel", rs_cellscan_kernel<1>, (int)n_cells, 256, 0, hist, d_cell_tile_start, (int)n_cells);
        if (vals_a) {
            PMRD_LAUNCH(ctx, rs_downsweep_kernel<true>, n_tiles, RS_THREADS, sizeof(RsSmem), kin, vin, kout, vout, hist, n, shift, mask);
        } else if (stage) {
            if (swide) PMRD_LAUNCH_NAMED(ctx, "rs_downsweep_staged_kernel", (rs_downsweep_staged_kernel<256, 16, RS_SWIDE_MIN_BLOCKS>), n_tiles, 256, sizeof(RsStage<8>), kin, kout, hist, shift, mask);
            else PMRD_LAUNCH_NAMED(ctx, "rs_downsweep_staged_kernel", (rs_downsweep_staged_kernel<RS_THREADS, RS_ITEMS, RS_DIRECT_MIN_BLOCKS>), n_tiles, RS_THREADS, sizeof(RsStage<RS_WARPS>), kin, kout, hist, shift, mask);
        } else if (direct) {   // packed 8-byte records, stored straight to their final place (L2 merges the sectors)
            if (wide == 2) PMRD_LAUNCH_NAMED(ctx, "rs_downsweep_direct_kernel", (rs_downsweep_direct_kernel<256, 16, RS_WIDE_MIN_BLOCKS>), n_tiles, 256, 0, kin, kout, hist, shift, mask);
            else PMRD_LAUNCH_NAMED(ctx, "rs_downsweep_direct_kernel", (rs_downsweep_direct_kernel<RS_THREADS, RS_ITEMS, RS_DIRECT_MIN_BLOCKS>), n_tiles, RS_THREADS, 0, kin, kout, hist, shift, mask);
        } else {   // packed 8-byte records: the "key" carries its payload in the bits above key_bits
            PMRD_LAUNCH(ctx, rs_downsweep_kernel<false>, n_tiles, RS_THREADS, sizeof(RsSmem), kin, vin, kout, vout, hist, n, shift, mask);
        }
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *in_b = n_passes & 1;
    return PMRD_OK;
}

// ---------------------------------------------------------------- driver ----------------
size_t radix_scratch_bytes(int64_t n) {
    const int64_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    int64_t chunk = 1;
    while (chunk * chunk < n_tiles) ++chunk;
    const int64_t n_chunks = chunk > 0 ? (n_tiles + chunk - 1) / chunk : 0;
    return (size_t)(n_tiles + n_chunks + 2) * 512 * sizeof(uint32_t);      // (sized for 9-bit digits)
}

int32_t radix_sort_pairs(pmrd_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                         int64_t n, uint64_t varying, int* in_b, uint32_t* d_scratch, int max_digit_bits) {
    *in_b = 0;
    if (n <= 1) return PMRD_OK;
    PMRD_ARG(ctx, n < (1ll << 32) - RS_TILE, "radix_sort_pairs: n must be < 2^32 (chunk the input)");
    RadixPlan plan = radix_make_plan(varying, max_digit_bits);
    if (plan.n_passes == 0) return PMRD_OK;

    const int n_tiles = (int)((n + RS_TILE - 1) / RS_TILE);
    int chunk = 1;
    while ((int64_t)chunk * chunk < n_tiles) ++chunk;
    const int n_chunks = (n_tiles + chunk - 1) / chunk;

    // caller-owned scratch (radix_scratch_bytes(n)): nothing is allocated here (plans: allocation-free steps)
    DevBuf hist_own, colsum_own;
    if (!d_scratch) {
        PMRD_CUDA(ctx, hist_own.alloc(ctx->stream, (size_t)n_tiles * 512 * sizeof(uint32_t)));
        PMRD_CUDA(ctx, colsum_own.alloc(ctx->stream, (size_t)n_chunks * 512 * sizeof(uint32_t)));
    }
    uint32_t* const hist = d_scratch ? d_scratch : hist_own.as<uint32_t>();
    uint32_t* const colsum = d_scratch ? d_scratch + (size_t)n_tiles * 512 : colsum_own.as<uint32_t>();

    PMRD_TRY(rs_set_attrs(ctx));

    uint64_t* kin = keys_a;
    uint32_t* vin = vals_a;
    uint64_t* kout = keys_b;
    uint32_t* vout = vals_b;
    for (int p = 0; p < plan.n_passes; ++p) {
        const int shift = plan.shift[p];
        const uint32_t mask = (1u << plan.bits[p]) - 1u;
        const bool wide = plan.bits[p] > 8;                      // a 9-bit digit: 512 bins
        const int bins = wide ? 512 : 256;
        if (wide) PMRD_LAUNCH_NAMED(ctx, "rs_upsweep_kernel", rs_upsweep_kernel<512>, n_tiles, RS_THREADS, RS_WARPS * 512 * sizeof(uint32_t), kin, n, hist, shift, mask);
        else PMRD_LAUNCH_NAMED(ctx, "rs_upsweep_kernel", rs_upsweep_kernel<256>, n_tiles, RS_THREADS, RS_WARPS * 256 * sizeof(uint32_t), kin, n, hist, shift, mask);
        PMRD_LAUNCH(ctx, rs_colsum_kernel, n_chunks, bins, 0, hist, n_tiles, chunk, colsum);
        PMRD_LAUNCH(ctx, rs_colbase_kernel, 1, bins, 0, colsum, n_chunks);
        PMRD_LAUNCH(ctx, rs_colapply_kernel, n_chunks, bins, 0, hist, n_tiles, chunk, colsum);
        if (vals_a && wide) {
            PMRD_LAUNCH_NAMED(ctx, "rs_downsweep_kernel<true>", (rs_downsweep_kernel<true, 512>), n_tiles, RS_THREADS, sizeof(RsSmemT<512>), kin, vin, kout, vout, hist, n, shift, mask);
        } else if (vals_a) {
            PMRD_LAUNCH(ctx, rs_downsweep_kernel<true>, n_tiles, RS_THREADS, sizeof(RsSmem), kin, vin, kout, vout,
                        hist, n, shift, mask);
        } else if (wide) {
            PMRD_LAUNCH_NAMED(ctx, "rs_downsweep_kernel<false>", (rs_downsweep_kernel<false, 512>), n_tiles, RS_THREADS, sizeof(RsSmemT<512>), kin, vin, kout, vout, hist, n, shift, mask);
        } else {
            PMRD_LAUNCH(ctx, rs_downsweep_kernel<false>, n_tiles, RS_THREADS, sizeof(RsSmem), kin, vin, kout, vout,
                        hist, n, shift, mask);
        }
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *in_b = (plan.n_passes & 1);
    return PMRD_OK;
}
