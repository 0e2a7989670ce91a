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
