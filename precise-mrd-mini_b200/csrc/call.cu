// call.cu -- K7: per-sample segmented reduction of the collapsed family table, then K8/K9.
// Replaces the O(samples x rows) boolean-mask scan of call_mrd
// (src/precise_mrd/call.py:161-229).  Counts are exact integers; the fp64 means use a
// fixed-shape reduction (per-thread strided partial sums, then a fixed shuffle tree), so
// they are deterministic run-to-run and within ~1e-16 relative of NumPy's pairwise mean.
#include "kernels.cuh"

#define CALL_THREADS 256

__global__ void __launch_bounds__(CALL_THREADS)
call_segment_kernel(const int64_t* __restrict__ seg_offsets, int64_t n_segments, const uint8_t* __restrict__ is_variant,
                    const double* __restrict__ quality, const double* __restrict__ agreement, double mean_error_rate,
                    int64_t* __restrict__ n_variants, int64_t* __restrict__ n_total, double* __restrict__ variant_fraction,
                    double* __restrict__ expected_variants, double* __restrict__ fold_change,
                    double* __restrict__ mean_quality, double* __restrict__ mean_consensus) {
    __shared__ double s_q[CALL_THREADS / 32], s_a[CALL_THREADS / 32];
    __shared__ unsigned long long s_k[CALL_THREADS / 32];
    for (int64_t s = blockIdx.x; s < n_segments; s += gridDim.x) {
        const int64_t lo = seg_offsets[s], hi = seg_offsets[s + 1];
        unsigned long long k = 0;
        double sq = 0.0, sa = 0.0;
        for (int64_t r = lo + threadIdx.x; r < hi; r += CALL_THREADS) {
            k += is_variant[r] ? 1ull : 0ull;
            if (quality) sq += quality[r];
            if (agreement) sa += agreement[r];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            k += __shfl_xor_sync(0xffffffffu, k, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
        }
        if ((threadIdx.x & 31) == 0) {
            s_k[threadIdx.x >> 5] = k;
            s_q[threadIdx.x >> 5] = sq;
            s_a[threadIdx.x >> 5] = sa;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long kk = 0;
            double q = 0.0, a = 0.0;
            for (int w = 0; w < CALL_THREADS / 32; ++w) {
                kk += s_k[w];
                q += s_q[w];
                a += s_a[w];
            }
            const int64_t nt = hi - lo;
            const double ntd = (double)nt, kd = (double)kk;
            n_variants[s] = (int64_t)kk;
            n_total[s] = nt;
            // call.py:179,190-193,203: single IEEE operations, no contraction
            const double expected = __dmul_rn(ntd, mean_error_rate);
            expected_variants[s] = expected;
            variant_fraction[s] = nt > 0 ? __ddiv_rn(kd, ntd) : 0.0;
            fold_change[s] = expected > 0.0 ? __ddiv_rn(kd, expected) : (kk > 0 ? INFINITY : 1.0);
            mean_quality[s] = quality ? __ddiv_rn(q, ntd) : NAN;
            mean_consensus[s] = agreement ? __ddiv_rn(a, ntd) : NAN;
        }
        __syncthreads();
    }
}

int32_t k7_call(pmrd_ctx* ctx, const int64_t* d_seg_offsets, int64_t n_segments, const uint8_t* d_is_variant,
                const double* d_quality, const double* d_agreement, double mean_error_rate, int32_t test_type,
                double alpha, const pmrd_call_table* o) {
    PMRD_ARG(ctx, test_type == PMRD_TEST_POISSON || test_type == PMRD_TEST_BINOMIAL, "Unknown test type");
    if (n_segments <= 0) return PMRD_OK;
    int grid = (int)(n_segments < (int64_t)ctx->sm_count * 64 ? n_segments : (int64_t)ctx->sm_count * 64);
    PMRD_LAUNCH(ctx, call_segment_kernel, grid, CALL_THREADS, 0, d_seg_offsets, n_segments, d_is_variant, d_quality,
                d_agreement, mean_error_rate, o->n_variants, o->n_total, o->variant_fraction, o->expected_variants,
                o->fold_change, o->mean_quality, o->mean_consensus);
    // p-values: poisson_test(k, n*rate) / binomial_test(k, n, rate)   (call.py:182-185)
    DevBuf rate;
    PMRD_CUDA(ctx, rate.alloc(ctx->stream, 8));
    PMRD_CUDA(ctx, cudaMemcpyAsync(rate.p, &mean_error_rate, 8, cudaMemcpyHostToDevice, ctx->stream));
    PMRD_TRY(k8_pvalues(ctx, o->n_variants, o->n_total, rate.as<double>(), 0, n_segments, test_type, PMRD_ALT_TWO_SIDED, o->p_value));
    PMRD_TRY(k9_bh(ctx, o->p_value, n_segments, alpha, o->significant, o->p_adjusted));
    return PMRD_OK;
}
