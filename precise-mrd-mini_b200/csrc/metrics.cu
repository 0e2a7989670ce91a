// metrics.cu -- K12: ROC-AUC / average precision of one score vector under many bootstrap
// resamples (src/precise_mrd/metrics.py:18-33 roc_auc_score / average_precision and
// :48-118 bootstrap_metric, which runs 2 x n_bootstrap sklearn calls per smoke run).
//
// The scores are sorted ONCE (descending, K4 radix sort on the order-preserving bit pattern)
// and cut into groups of equal score -- sklearn's "distinct thresholds".  A resample is then
// only a multiset of row indices: one launch histograms every resample's draws into per-group
// (negatives, positives) counters, and one block per resample walks its groups in threshold
// order accumulating
//   AUC = sum (fpr_i - fpr_{i-1}) * (tpr_i + tpr_{i-1}) / 2        (np.trapz over roc_curve)
//   AP  = sum (R_i - R_{i-1}) * P_i                                 (average_precision_score)
// with fpr = fps / fps[-1], tpr = R = tps / tps[-1], P = tps / (tps + fps), each formed by the
// same IEEE divisions as sklearn.  Groups a resample did not draw add no threshold (and no
// term), exactly as in a per-resample sklearn call.  The resample indices are either the
// caller's (metrics.py:76-79 draws them up-front with rng.choice -- hand them over and the
// result is the reference's resample for resample) or Philox draws keyed (seed, stream, b, t).
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "rng.cuh"
#include "scan.cuh"

#define PMRD_STREAM_METRICS 8u
#define BM_THREADS 256

__global__ void __launch_bounds__(256)
bm_pack_kernel(const double* __restrict__ score, int64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = score[i];
    if (v == 0.0) v = 0.0;                                   // -0.0 and +0.0 are one threshold
    const uint64_t b = (uint64_t)__double_as_longlong(v);
    const uint64_t asc = (b >> 63) ? ~b : (b | 0x8000000000000000ull);   // total order of the doubles
    keys[i] = ~asc;                                          // descending score == ascending key
    idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
bm_flag_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i > 0 && keys[i] != keys[i - 1]) ? 1u : 0u;
}

// group_of[row] = index of the row's threshold group (0 = highest score)
__global__ void __launch_bounds__(256)
bm_group_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ excl,
                int64_t n, uint32_t* __restrict__ group_of) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    group_of[idx[i]] = excl[i] + flag[i];
}

// one thread per (resample, draw): count the drawn row into its group's (neg, pos) pair
__global__ void __launch_bounds__(256)
bm_count_kernel(uint64_t seed, uint64_t stream_id, const int64_t* __restrict__ indices /* [n_res][n] or NULL */,
                const uint8_t* __restrict__ y, const uint32_t* __restrict__ group_of, int64_t n, int64_t b0, int64_t n_res,
                int64_t identity_b /* resample that is the sample itself, or -1 */, uint32_t n_groups,
                uint32_t* __restrict__ cnt /* [n_res][n_groups][2] */, int* __restrict__ bad) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;                                    // resample inside this batch
    if (t >= n || r >= n_res) return;
    const int64_t b = b0 + r;
    int64_t i;
    if (b == identity_b) {
        i = t;
    } else if (indices) {
        i = indices[b * n + t];
        if (i < 0 || i >= n) { *bad = 1; return; }
    } else {
        const Philox ph(seed);
        const uint4 w = ph((uint32_t)t, (uint32_t)b, (uint32_t)stream_id,
                           (PMRD_STREAM_METRICS << 24) | (uint32_t)((stream_id >> 32) & 0xffffffu));
        i = (int64_t)(u53(w.x, w.y) * (double)n);
        if (i >= n) i = n - 1;
    }
    atomicAdd(&cnt[((size_t)r * n_groups + group_of[i]) * 2 + (y[i] ? 1 : 0)], 1u);
}

// one block per resample: walk the groups in threshold order
__global__ void __launch_bounds__(BM_THREADS)
bm_reduce_kernel(const uint32_t* __restrict__ cnt, uint32_t n_groups, int64_t b0, double* __restrict__ auc, double* __restrict__ ap) {
    __shared__ unsigned long long sw[33];
    __shared__ double s_red[2][BM_THREADS / 32];
    const uint32_t* c = cnt + (size_t)blockIdx.x * n_groups * 2;
    const uint32_t per = (n_groups + BM_THREADS - 1) / BM_THREADS;
    const uint32_t lo = min(threadIdx.x * per, n_groups), hi = min(lo + per, n_groups);
    unsigned long long neg = 0, pos = 0;
    for (uint32_t g = lo; g < hi; ++g) { neg += c[2 * g]; pos += c[2 * g + 1]; }
    unsigned long long tot_neg, tot_pos;
    unsigned long long fp = block_exclusive_scan<unsigned long long>(neg, sw, tot_neg);
    unsigned long long tp = block_exclusive_scan<unsigned long long>(pos, sw, tot_pos);
    const double TP = (double)tot_pos, FP = (double)tot_neg;
    double a = 0.0, p = 0.0;
    if (tot_pos > 0) {
        const bool roc = tot_neg > 0;
        double tpr0 = __ddiv_rn((double)tp, TP), fpr0 = roc ? __ddiv_rn((double)fp, FP) : 0.0;
        for (uint32_t g = lo; g < hi; ++g) {
            const uint32_t cn = c[2 * g], cp = c[2 * g + 1];
            if (cn + cp == 0) continue;                              // not drawn: no threshold here
            fp += cn; tp += cp;
            const double tpr1 = __ddiv_rn((double)tp, TP);
            if (roc) {
                const double fpr1 = __ddiv_rn((double)fp, FP);
                a = __dadd_rn(a, __ddiv_rn(__dmul_rn(__dadd_rn(fpr1, -fpr0), __dadd_rn(tpr1, tpr0)), 2.0));
                fpr0 = fpr1;
            }
            p = __dadd_rn(p, __dmul_rn(__dadd_rn(tpr1, -tpr0), __ddiv_rn((double)tp, (double)(tp + fp))));
            tpr0 = tpr1;
        }
    }
    // fixed-shape reduction: lanes, then warps in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); p += __shfl_down_sync(0xffffffffu, p, o); }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = a; s_red[1][threadIdx.x >> 5] = p; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double A = 0.0, P = 0.0;
        for (int w = 0; w < BM_THREADS / 32; ++w) { A += s_red[0][w]; P += s_red[1][w]; }
        // metrics.py:18-33: one class only -> 0.5 / mean(y_true)
        auc[b0 + blockIdx.x] = (tot_pos > 0 && tot_neg > 0) ? A : 0.5;
        ap[b0 + blockIdx.x] = tot_pos > 0 ? P : 0.0;
    }
}

// d_auc / d_ap hold n_boot entries (+1 at index n_boot for the un-resampled sample when want_point)
int32_t k12_bootstrap_metrics(pmrd_ctx* ctx, const uint8_t* d_y, const double* d_score, int64_t n, const int64_t* d_indices,
                              int64_t n_boot, int want_point, uint64_t stream_id, double* d_auc, double* d_ap) {
    PMRD_ARG(ctx, n > 0 && n < (1ll << 31), "bootstrap metrics: 0 < n < 2^31");
    PMRD_ARG(ctx, n_boot >= 0, "bootstrap metrics: n_boot");
    const int64_t n_res_all = n_boot + (want_point ? 1 : 0);
    if (n_res_all == 0) return PMRD_OK;
    DevBuf ka, kb, va, vb, flag, excl, grp, tot, bad;
    PMRD_CUDA(ctx, ka.alloc(ctx->stream, (size_t)n * 8));
    PMRD_CUDA(ctx, kb.alloc(ctx->stream, (size_t)n * 8));
    PMRD_CUDA(ctx, va.alloc(ctx->stream, (size_t)n * 4));
    PMRD_CUDA(ctx, vb.alloc(ctx->stream, (size_t)n * 4));
    PMRD_CUDA(ctx, flag.alloc(ctx->stream, (size_t)n * 4));
    PMRD_CUDA(ctx, excl.alloc(ctx->stream, (size_t)n * 4));
    PMRD_CUDA(ctx, grp.alloc(ctx->stream, (size_t)n * 4));
    PMRD_CUDA(ctx, tot.alloc(ctx->stream, 4));
    PMRD_CUDA(ctx, bad.alloc(ctx->stream, 4));
    PMRD_CUDA(ctx, cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
    PMRD_LAUNCH(ctx, bm_pack_kernel, pmrd_blocks(n, 256), 256, 0, d_score, n, ka.as<uint64_t>(), va.as<uint32_t>());
    uint64_t varying = 0;
    PMRD_TRY(radix_varying_bits(ctx, ka.as<uint64_t>(), n, &varying));
    int in_b = 0;
    PMRD_TRY(radix_sort_pairs(ctx, ka.as<uint64_t>(), va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), n, varying, &in_b));
    const uint64_t* keys = in_b ? kb.as<uint64_t>() : ka.as<uint64_t>();
    const uint32_t* idx = in_b ? vb.as<uint32_t>() : va.as<uint32_t>();
    PMRD_LAUNCH(ctx, bm_flag_kernel, pmrd_blocks(n, 256), 256, 0, keys, n, flag.as<uint32_t>());
    PMRD_TRY((exclusive_scan<uint32_t, uint32_t>(ctx, flag.as<uint32_t>(), excl.as<uint32_t>(), n, tot.as<uint32_t>())));
    PMRD_LAUNCH(ctx, bm_group_kernel, pmrd_blocks(n, 256), 256, 0, idx, (const uint32_t*)flag.as<uint32_t>(),
                (const uint32_t*)excl.as<uint32_t>(), n, grp.as<uint32_t>());
    uint32_t h_tot = 0;
    PMRD_CUDA(ctx, cudaMemcpyAsync(&h_tot, tot.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t n_groups = h_tot + 1;
    // resamples in batches whose counters stay under 1 GiB (and under the grid.y limit)
    int64_t batch = (1ll << 30) / ((int64_t)n_groups * 8);
    batch = batch < 1 ? 1 : (batch > 32768 ? 32768 : batch);
    DevBuf cnt;
    PMRD_CUDA(ctx, cnt.alloc(ctx->stream, (size_t)(batch < n_res_all ? batch : n_res_all) * n_groups * 8));
    for (int64_t b0 = 0; b0 < n_res_all; b0 += batch) {
        const int64_t nr = n_res_all - b0 < batch ? n_res_all - b0 : batch;
        PMRD_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, (size_t)nr * n_groups * 8, ctx->stream));
        const dim3 grid((unsigned)pmrd_blocks(n, 256), (unsigned)nr);
        PMRD_LAUNCH(ctx, bm_count_kernel, grid, 256, 0, ctx->seed, stream_id, d_indices, d_y, (const uint32_t*)grp.as<uint32_t>(), n, b0,
                    nr, want_point ? n_boot : (int64_t)-1, n_groups, cnt.as<uint32_t>(), bad.as<int>());
        PMRD_LAUNCH(ctx, bm_reduce_kernel, (int)nr, BM_THREADS, 0, (const uint32_t*)cnt.as<uint32_t>(), n_groups, b0, d_auc, d_ap);
    }
    int h_bad = 0;
    PMRD_CUDA(ctx, cudaMemcpyAsync(&h_bad, bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    PMRD_ARG(ctx, h_bad == 0, "bootstrap metrics: resample index out of range");
    return PMRD_OK;
}
