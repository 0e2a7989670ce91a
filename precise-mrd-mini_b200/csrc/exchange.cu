// exchange.cu -- K18: the one exchange step of the path (third bullet).
//
// The reference collapses ONE process's reads (UMIProcessor.process_reads, src/precise_mrd/umi.py:205-285; the Rust
// group_umis_fast, rust_ext/src/lib.rs:7-109).  When the reads of a sample are held by several GPUs in arrival order
// (a run striped over the ranks), the families of a locus are spread over all of them: every locus gets an OWNER rank,
// each rank cuts its reads into one contiguous run per owner, the runs change hands (all-to-all over NVLink, issued by
// the host through torch.distributed / NCCL), and the owner collapses what it received with the single-GPU kernels.
//
//   owner(group) = group mod world           (world a power of two: the low bits -- dense ids spread evenly)
//
// The cut is a STABLE partition (one radix pass over those bits), and all-to-all delivers the runs in source-rank order,
// so an owner sees the reads of a family in (rank, arrival) order: the order of the concatenated input.  Its table is
// therefore the single-GPU table of the concatenated reads, bit for bit (the weighted vote adds in read order).
// After the exchange a rank only holds groups congruent to its rank: owner_rekey_kernel divides the group field by world
// so that the ids are dense again for the group-major collapse (collapse.cu: one block per group id).
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

#define OWNER_MAX_LOG2 8

__global__ void __launch_bounds__(256)
owner_count_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t mask, unsigned long long* __restrict__ counts) {
    __shared__ uint32_t s_c[1 << OWNER_MAX_LOG2];
    s_c[threadIdx.x] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&s_c[(uint32_t)(keys[i] >> 32) & mask], 1u);
    __syncthreads();
    if (threadIdx.x <= mask && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
}

__global__ void __launch_bounds__(256) owner_rekey_kernel(uint64_t* __restrict__ keys, int64_t n, int log2_world) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        keys[i] = ((k >> 32) >> log2_world) << 32 | (k & 0xffffffffull);
    }
}

int32_t k18_owner_partition(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp, uint32_t* d_payload_tmp,
                            int64_t n, int32_t log2_world, int64_t* h_counts, int32_t* in_tmp) {
    PMRD_ARG(ctx, log2_world >= 0 && log2_world <= OWNER_MAX_LOG2, "world = 2^k ranks, k <= 8");
    PMRD_ARG(ctx, h_counts && in_tmp, "null output");
    const int world = 1 << log2_world;
    *in_tmp = 0;
    for (int r = 0; r < world; ++r) h_counts[r] = 0;
    if (n <= 0) return PMRD_OK;
    PMRD_ARG(ctx, d_keys && d_payload && d_keys_tmp && d_payload_tmp, "null buffer");
    const uint32_t mask = (uint32_t)world - 1u;
    DevBuf cnt;
    PMRD_CUDA(ctx, cnt.alloc(ctx->stream, (size_t)world * 8));
    PMRD_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, (size_t)world * 8, ctx->stream));
    int grid = pmrd_blocks(n, 256 * 8);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    PMRD_LAUNCH(ctx, owner_count_kernel, grid, 256, 0, (const uint64_t*)d_keys, n, mask, cnt.as<unsigned long long>());
    int in_b = 0;
    if (log2_world > 0) PMRD_TRY(radix_sort_pairs(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n, (uint64_t)mask << 32, &in_b));
    *in_tmp = in_b;
    unsigned long long h[1 << OWNER_MAX_LOG2];
    PMRD_CUDA(ctx, cudaMemcpyAsync(h, cnt.p, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < world; ++r) h_counts[r] = (int64_t)h[r];
    return PMRD_OK;
}

int32_t k18_owner_rekey(pmrd_ctx* ctx, uint64_t* d_keys, int64_t n, int32_t log2_world) {
    PMRD_ARG(ctx, log2_world >= 0 && log2_world <= OWNER_MAX_LOG2, "world = 2^k ranks, k <= 8");
    if (n <= 0 || log2_world == 0) return PMRD_OK;
    PMRD_ARG(ctx, d_keys != nullptr, "null buffer");
    int grid = pmrd_blocks(n, 256 * 4);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    PMRD_LAUNCH(ctx, owner_rekey_kernel, grid, 256, 0, d_keys, n, (int)log2_world);
    return PMRD_OK;
}

// ---------------------------------------------------------------- cut + exchange in ONE kernel, over peer memory ------
// The NCCL form above stages every read twice (partitioned copy, then the all-to-all's receive buffer).  On one NVSwitch
// node a rank can store straight into the owner's receive buffer: the scatter kernel below IS the all-to-all.  Each owner
// exports its receive arrays (cudaMalloc + cudaIpcGetMemHandle, pmrd_peer_alloc), every rank maps them
// (cudaIpcOpenMemHandle, pmrd_peer_open) and the cut writes record i of a tile to
//     peer[owner].keys[ base[owner] + tile_off[owner][tile] + (stable rank of i among the tile's reads of that owner) ]
// where base[owner] = reads that lower ranks send to that owner (one small all-gather of the count matrix on the host side).
// Same order at the owner as the NCCL form: source rank, then arrival.
#define OX_THREADS 256
#define OX_ITEMS 16
#define OX_TILE (OX_THREADS * OX_ITEMS)
#define OX_MAX_WORLD 8

// per tile, reads per owner: hist[owner * n_tiles + tile]
__global__ void __launch_bounds__(OX_THREADS)
owner_tile_count_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t mask, int64_t n_tiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_c[OX_MAX_WORLD];
    if (threadIdx.x < OX_MAX_WORLD) s_c[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * OX_TILE;
    // eight 8-bit counters in one word: a thread sees at most 16 reads
    unsigned long long acc = 0ull;
#pragma unroll
    for (int i = 0; i < OX_ITEMS; ++i) {
        const int64_t j = base + (int64_t)i * OX_THREADS + threadIdx.x;
        if (j < n) acc += 1ull << (8u * ((uint32_t)(keys[j] >> 32) & mask));
    }
    // widen to 16-bit fields (a warp sees at most 512) and add over the warp
    unsigned long long lo = (acc & 0xffull) | ((acc & 0xff00ull) << 8) | ((acc & 0xff0000ull) << 16) | ((acc & 0xff000000ull) << 24);
    unsigned long long hi = ((acc >> 32) & 0xffull) | (((acc >> 32) & 0xff00ull) << 8) | (((acc >> 32) & 0xff0000ull) << 16) | (((acc >> 32) & 0xff000000ull) << 24);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo += __shfl_xor_sync(0xffffffffu, lo, o);
        hi += __shfl_xor_sync(0xffffffffu, hi, o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const uint32_t a = (uint32_t)(lo >> (16 * d)) & 0xffffu, b = (uint32_t)(hi >> (16 * d)) & 0xffffu;
            if (a) atomicAdd(&s_c[d], a);
            if (b) atomicAdd(&s_c[4 + d], b);
        }
    }
    __syncthreads();
    if (threadIdx.x <= mask) hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = s_c[threadIdx.x];
}

// one block per owner: exclusive scan of its row over the tiles, in place; totals[owner] = reads bound for it
__global__ void __launch_bounds__(1024) owner_tile_scan_kernel(uint32_t* __restrict__ hist, int64_t n_tiles, unsigned long long* __restrict__ totals) {
    __shared__ uint32_t sw[33];
    uint32_t* row = hist + (int64_t)blockIdx.x * n_tiles;
    uint32_t run = 0;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += 1024) {
        const int64_t t = t0 + threadIdx.x;
        const uint32_t v = t < n_tiles ? row[t] : 0u;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan<uint32_t>(v, sw, tot);
        if (t < n_tiles) row[t] = run + ex;
        run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = run;
}

struct OwnerPeers {
    uint64_t* keys[OX_MAX_WORLD];
    uint32_t* payload[OX_MAX_WORLD];
    unsigned long long base[OX_MAX_WORLD];
};

struct OxSmem {
    uint64_t key[OX_TILE];
    uint32_t pl[OX_TILE];
    unsigned long long delta[OX_MAX_WORLD];                       // position in the tile -> position at the owner
    uint32_t run[OX_THREADS / 32][OX_MAX_WORLD];                  // first: reads of (warp, owner); then: where that run starts in the tile
};

// The tile is first put in owner order in shared memory and then streamed out position by position, so that a warp's
// store is 32 consecutive records of ONE owner (256 + 128 bytes): at 8 ranks a direct store from the ranking registers
// would leave a warp with ~4 lanes per owner, 32-byte NVLink writes.
__global__ void __launch_bounds__(OX_THREADS, 3)
owner_scatter_peer_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ payload, int64_t n, int log2_world, int64_t n_tiles,
                          const uint32_t* __restrict__ tile_off /*[world][n_tiles]*/, OwnerPeers P) {
    constexpr int WARPS = OX_THREADS / 32;
    extern __shared__ __align__(16) unsigned char ox_raw[];
    OxSmem& S = *reinterpret_cast<OxSmem*>(ox_raw);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lt = lanemask_lt();
    const uint32_t mask = (1u << log2_world) - 1u;
    if (threadIdx.x < WARPS * OX_MAX_WORLD) (&S.run[0][0])[threadIdx.x] = 0u;
    const int64_t tile_base = (int64_t)blockIdx.x * OX_TILE;
    const int64_t base = tile_base + (int64_t)warp * (OX_TILE / WARPS);
    const int tile_n = n - tile_base < OX_TILE ? (int)(n - tile_base) : OX_TILE;
    uint64_t key[OX_ITEMS];
    uint32_t pl[OX_ITEMS];
#pragma unroll
    for (int i = 0; i < OX_ITEMS; ++i) {
        const int64_t j = base + i * 32 + lane;
        key[i] = j < n ? keys[j] : ~0ull;
        pl[i] = j < n ? payload[j] : 0u;
    }
    __syncthreads();
    uint32_t rank2[OX_ITEMS / 2];                                   // two 16-bit ranks per register (a warp holds 512 reads)
#pragma unroll
    for (int i = 0; i < OX_ITEMS; ++i) {
        const int64_t j = base + i * 32 + lane;
        const bool valid = j < n;
        const uint32_t d = (uint32_t)(key[i] >> 32) & mask;
        // lanes with my owner (log2_world ballots); lanes past the end of the input form their own group
        uint32_t m = __ballot_sync(0xffffffffu, valid);
        m = valid ? m : ~m;
        for (int b = 0; b < log2_world; ++b) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
            m &= ((d >> b) & 1u) ? bal : ~bal;
        }
        const uint32_t before = __popc(m & lt);
        const int leader = __ffs((int)m) - 1;
        uint32_t old = 0u;
        if (valid && before == 0) {
            old = S.run[warp][d];                                   // (only this warp touches its row: no atomics)
            S.run[warp][d] = old + (uint32_t)__popc(m);
        }
        __syncwarp();
        old = __shfl_sync(0xffffffffu, old, leader);
        const uint32_t r = old + before;
        if (i & 1) rank2[i >> 1] |= r << 16;
        else rank2[i >> 1] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // owner runs of the tile, owner 0 first; inside a run the warps in order
        uint32_t at = 0u;
        for (uint32_t o = 0; o <= mask; ++o) {
            S.delta[o] = P.base[o] + tile_off[(int64_t)o * n_tiles + blockIdx.x] - at;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = S.run[w][o];
                S.run[w][o] = at;
                at += c;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < OX_ITEMS; ++i) {
        const int64_t j = base + i * 32 + lane;
        if (j < n) {
            const uint32_t d = (uint32_t)(key[i] >> 32) & mask;
            const uint32_t r = (i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu;
            const uint32_t at = S.run[warp][d] + r;
            S.key[at] = key[i];
            S.pl[at] = pl[i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < OX_ITEMS; ++i) {
        const int j = i * OX_THREADS + (int)threadIdx.x;
        if (j < tile_n) {
            const uint64_t k = S.key[j];
            const uint32_t d = (uint32_t)(k >> 32) & mask;
            const unsigned long long at = S.delta[d] + (unsigned long long)j;
            // stored with the owner's dense group id (group / world): no separate re-keying pass over the received reads
            P.keys[d][at] = (((k >> 32) >> log2_world) << 32) | (k & 0xffffffffull);
            P.payload[d][at] = S.pl[j];
        }
    }
}

int32_t k18_owner_tiles(pmrd_ctx* ctx, const uint64_t* d_keys, int64_t n, int32_t log2_world, uint32_t* d_tile_off, int64_t* h_counts) {
    PMRD_ARG(ctx, log2_world >= 0 && (1 << log2_world) <= OX_MAX_WORLD, "the peer-memory exchange serves up to 8 ranks");
    PMRD_ARG(ctx, h_counts != nullptr, "null output");
    const int world = 1 << log2_world;
    for (int r = 0; r < world; ++r) h_counts[r] = 0;
    if (n <= 0) return PMRD_OK;
    PMRD_ARG(ctx, d_keys && d_tile_off, "null buffer");
    const int64_t n_tiles = (n + OX_TILE - 1) / OX_TILE;
    DevBuf tot;
    PMRD_CUDA(ctx, tot.alloc(ctx->stream, (size_t)world * 8));
    PMRD_LAUNCH(ctx, owner_tile_count_kernel, (int)n_tiles, OX_THREADS, 0, d_keys, n, (uint32_t)world - 1u, n_tiles, d_tile_off);
    PMRD_LAUNCH(ctx, owner_tile_scan_kernel, world, 1024, 0, d_tile_off, n_tiles, tot.as<unsigned long long>());
    unsigned long long h[OX_MAX_WORLD];
    PMRD_CUDA(ctx, cudaMemcpyAsync(h, tot.p, (size_t)world * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < world; ++r) h_counts[r] = (int64_t)h[r];
    return PMRD_OK;
}

int32_t k18_owner_scatter_peer(pmrd_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_payload, int64_t n, int32_t log2_world,
                               const uint32_t* d_tile_off, uint64_t* const* peer_keys, uint32_t* const* peer_payload, const int64_t* h_base) {
    PMRD_ARG(ctx, log2_world >= 0 && (1 << log2_world) <= OX_MAX_WORLD, "the peer-memory exchange serves up to 8 ranks");
    if (n <= 0) return PMRD_OK;
    PMRD_ARG(ctx, d_keys && d_payload && d_tile_off && peer_keys && peer_payload && h_base, "null argument");
    const int world = 1 << log2_world;
    OwnerPeers P{};
    for (int r = 0; r < world; ++r) {
        PMRD_ARG(ctx, peer_keys[r] && peer_payload[r] && h_base[r] >= 0, "peer buffer");
        P.keys[r] = peer_keys[r];
        P.payload[r] = peer_payload[r];
        P.base[r] = (unsigned long long)h_base[r];
    }
    const int64_t n_tiles = (n + OX_TILE - 1) / OX_TILE;
    PMRD_CUDA(ctx, cudaFuncSetAttribute(owner_scatter_peer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OxSmem)));
    PMRD_LAUNCH(ctx, owner_scatter_peer_kernel, (int)n_tiles, OX_THREADS, sizeof(OxSmem), d_keys, d_payload, n, (int)log2_world, n_tiles, d_tile_off, P);
    return PMRD_OK;
}

// ---- receive buffers other ranks' kernels can store into (one process per GPU, same node)
int32_t k18_peer_alloc(pmrd_ctx* ctx, int64_t bytes, void** d_ptr, uint8_t* handle64) {
    PMRD_ARG(ctx, bytes > 0 && d_ptr && handle64, "peer_alloc arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle travels as 64 bytes");
    void* p = nullptr;
    PMRD_CUDA(ctx, cudaMalloc(&p, (size_t)bytes));                 // (not the stream-ordered pool: its memory cannot be exported)
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        PMRD_CUDA(ctx, e);
    }
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return PMRD_OK;
}

int32_t k18_peer_open(pmrd_ctx* ctx, const uint8_t* handle64, void** d_ptr) {
    PMRD_ARG(ctx, handle64 && d_ptr, "peer_open arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PMRD_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PMRD_OK;
}

int32_t k18_peer_close(pmrd_ctx* ctx, void* d_ptr) {
    if (d_ptr) PMRD_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return PMRD_OK;
}

int32_t k18_peer_free(pmrd_ctx* ctx, void* d_ptr) {
    if (d_ptr) PMRD_CUDA(ctx, cudaFree(d_ptr));
    return PMRD_OK;
}
