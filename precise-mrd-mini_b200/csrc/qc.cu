// qc.cu -- the two per-family reductions either side of the call stage that SURVEY.md §8f lists as "next":
//   * FeatureEngineer.extract_features (src/precise_mrd/advanced_stats.py:33-83): the feature matrix the
//     ML / Bayesian callers consume, straight from the collapsed-family columns;
//   * QCAnalyzer (src/precise_mrd/qc.py:54-215): family-size distribution and consensus efficiency of a
//     family table -- an exact size histogram plus the counts the guard-rails compare with their thresholds.
// Both are one pass over 12-28 bytes per family (HBM-bound); everything the report derives from them
// (moments, percentiles, rates) is exact integer arithmetic on the histogram, done by the host mirror.
#include "kernels.cuh"

// out[i][0..5] = family_size, quality_score, consensus_agreement, allele_fraction,
//                family_size / (quality_score + 1), consensus_agreement * allele_fraction
// (single IEEE operations, no contraction: bit-equal to the pandas expressions)
__global__ void __launch_bounds__(256)
gd_features_kernel(const int64_t* __restrict__ size, const double* __restrict__ q, const double* __restrict__ agr,
                   const double* __restrict__ af, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = (double)size[i], qi = q[i], a = agr[i], f = af[i];
    double* o = out + i * 6;
    o[0] = s;
    o[1] = qi;
    o[2] = a;
    o[3] = f;
    o[4] = __ddiv_rn(s, __dadd_rn(qi, 1.0));
    o[5] = __dmul_rn(a, f);
}

int32_t k15_gd_features(pmrd_ctx* ctx, const int64_t* d_size, const double* d_q, const double* d_agr, const double* d_af, int64_t n,
                        double* d_out) {
    if (n <= 0) return PMRD_OK;
    PMRD_LAUNCH(ctx, gd_features_kernel, pmrd_blocks(n, 256), 256, 0, d_size, d_q, d_agr, d_af, n, d_out);
    return PMRD_OK;
}

// hist[min(size, n_bins - 1)] += 1 (the last bin collects every larger family); counts[0] += families whose
// flags meet `flag_mask` (PMRD_FAM_HAS_CONSENSUS: call_consensus returned an allele), counts[1] = largest size
#define QC_THREADS 256
__global__ void __launch_bounds__(QC_THREADS)
qc_family_kernel(const uint32_t* __restrict__ size, const uint32_t* __restrict__ flags /* or NULL */, uint32_t flag_mask, int64_t n,
                 int n_bins, unsigned long long* __restrict__ hist, unsigned long long* __restrict__ counts) {
    extern __shared__ uint32_t s_h[];
    for (int b = threadIdx.x; b < n_bins; b += QC_THREADS) s_h[b] = 0u;
    __syncthreads();
    unsigned long long n_cons = 0;
    uint32_t vmax = 0;
    for (int64_t i = (int64_t)blockIdx.x * QC_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * QC_THREADS) {
        const uint32_t s = size[i];
        atomicAdd(&s_h[s < (uint32_t)(n_bins - 1) ? s : (uint32_t)(n_bins - 1)], 1u);
        vmax = s > vmax ? s : vmax;
        if (flags) n_cons += (flags[i] & flag_mask) != 0u;
    }
    n_cons = warp_sum(n_cons);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if ((threadIdx.x & 31) == 0) {
        if (n_cons) atomicAdd(&counts[0], n_cons);
        atomicMax(&counts[1], (unsigned long long)vmax);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += QC_THREADS)
        if (s_h[b]) atomicAdd(&hist[b], (unsigned long long)s_h[b]);
}

int32_t k16_qc_family(pmrd_ctx* ctx, const uint32_t* d_size, const uint32_t* d_flags, uint32_t flag_mask, int64_t n, int n_bins,
                      unsigned long long* d_hist, unsigned long long* d_counts) {
    PMRD_ARG(ctx, n_bins >= 2 && n_bins <= 8192, "2 <= n_bins <= 8192");
    PMRD_CUDA(ctx, cudaMemsetAsync(d_hist, 0, (size_t)n_bins * 8, ctx->stream));
    PMRD_CUDA(ctx, cudaMemsetAsync(d_counts, 0, 16, ctx->stream));
    if (n <= 0) return PMRD_OK;
    int grid = pmrd_blocks(n, QC_THREADS * 8);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    PMRD_LAUNCH(ctx, qc_family_kernel, grid, QC_THREADS, (size_t)n_bins * 4, d_size, d_flags, flag_mask, n, n_bins, d_hist, d_counts);
    return PMRD_OK;
}
