// pvalues.cu -- K8 (tail p-values) and K9 (Benjamini-Hochberg) + their C-ABI entry points.
//   K8 replaces call.py:14-38 / stats.py:47-133 (scipy.stats.poisson / binomtest)
//   K9 replaces call.py:41-79 benjamini_hochberg_correction (models/statistical.py:108-140)
#include "common.cuh"
#include "radix_sort.cuh"
#include "stats.cuh"
#include "kernels.cuh"

// ---------------------------------------------------------------- K8 --------------------
__global__ void __launch_bounds__(128)
pvalues_kernel(const int64_t* __restrict__ k, const int64_t* __restrict__ n, const double* __restrict__ rate,
               int64_t rate_stride, const int32_t* __restrict__ rate_index /* or NULL: rate[i * rate_stride] */, int64_t m,
               int test_type, int alternative, double* __restrict__ p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t ki = k[i];
    const double r = rate_index ? rate[rate_index[i]] : rate[i * rate_stride];
    double out;
    if (test_type == PMRD_TEST_POISSON) {
        // call.py:179 expected = n_total * mean_error_rate: one IEEE multiply, no contraction
        const double expected = n ? __dmul_rn((double)n[i], r) : r;
        out = alternative == PMRD_ALT_TWO_SIDED ? pmrd_poisson_two_sided(ki, expected)
              : alternative == PMRD_ALT_GREATER ? pmrd_poisson_greater(ki, expected) : pmrd_poisson_less(ki, expected);
    } else {
        const int64_t ni = n[i];
        out = alternative == PMRD_ALT_TWO_SIDED ? pmrd_binomial_two_sided(ki, ni, r)
              : alternative == PMRD_ALT_GREATER ? pmrd_binomial_greater(ki, ni, r) : pmrd_binomial_less(ki, ni, r);
    }
    p[i] = out;
}

int32_t k8_pvalues(pmrd_ctx* ctx, const int64_t* d_k, const int64_t* d_n, const double* d_rate, int64_t rate_stride,
                   int64_t m, int32_t test_type, int32_t alternative, double* d_p, const int32_t* d_rate_index) {
    PMRD_ARG(ctx, test_type == PMRD_TEST_POISSON || test_type == PMRD_TEST_BINOMIAL, "Unknown test type");
    PMRD_ARG(ctx, alternative == PMRD_ALT_TWO_SIDED || alternative == PMRD_ALT_GREATER || alternative == PMRD_ALT_LESS,
             "unknown alternative");
    PMRD_ARG(ctx, test_type == PMRD_TEST_POISSON || d_n != nullptr, "binomial test needs n");
    if (m <= 0) return PMRD_OK;
    PMRD_LAUNCH(ctx, pvalues_kernel, pmrd_blocks(m, 128), 128, 0, d_k, d_n, d_rate, rate_stride, d_rate_index, m, test_type, alternative, d_p);
    return PMRD_OK;
}

// ---------------------------------------------------------------- K14 -------------------
// Fisher's exact test on m 2x2 tables (filters.py:50-100 strand bias): one thread per table
__global__ void __launch_bounds__(128)
fisher_kernel(const int64_t* __restrict__ table /* [m][4] = c00 c01 c10 c11 */, int64_t m, double* __restrict__ odds,
              double* __restrict__ p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double o;
    p[i] = pmrd_fisher_two_sided(table[4 * i], table[4 * i + 1], table[4 * i + 2], table[4 * i + 3], &o);
    odds[i] = o;
}

int32_t k14_fisher(pmrd_ctx* ctx, const int64_t* d_table, int64_t m, double* d_odds, double* d_p) {
    if (m <= 0) return PMRD_OK;
    PMRD_LAUNCH(ctx, fisher_kernel, pmrd_blocks(m, 128), 128, 0, d_table, m, d_odds, d_p);
    return PMRD_OK;
}

// ---------------------------------------------------------------- K9 --------------------
#define BH_THREADS 256
#define BH_ITEMS 8
#define BH_TILE (BH_THREADS * BH_ITEMS)

__global__ void __launch_bounds__(256) bh_pack_kernel(const double* __restrict__ p, int64_t m, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    // non-negative doubles order like their bit patterns; -0.0 is folded onto +0.0
    double v = p[i];
    if (v == 0.0) v = 0.0;
    keys[i] = (uint64_t)__double_as_longlong(v);
    idx[i] = (uint32_t)i;
}

// adj_i = p_(i) * m / (i+1)   (call.py:59-61: one multiply, one divide, in that order)
__device__ __forceinline__ double bh_raw_adjusted(uint64_t key, int64_t i, double m) {
    return __ddiv_rn(__dmul_rn(__longlong_as_double((long long)key), m), (double)(i + 1));
}

// per tile: minimum of the raw adjusted values; largest rank meeting p_(i) <= (i+1)/m*alpha
__global__ void __launch_bounds__(BH_THREADS)
bh_tile_kernel(const uint64_t* __restrict__ keys, int64_t m, double alpha, double* __restrict__ tile_min,
               unsigned long long* __restrict__ kstar /* max (i+1) */) {
    __shared__ double s_min[BH_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * BH_TILE;
    const double md = (double)m;
    double mn = INFINITY;
    unsigned long long best = 0;
#pragma unroll
    for (int j = 0; j < BH_ITEMS; ++j) {
        const int64_t i = base + j * BH_THREADS + threadIdx.x;
        if (i < m) {
            const uint64_t key = keys[i];
            mn = fmin(mn, bh_raw_adjusted(key, i, md));
            // call.py:55: sorted_p[i] <= (i + 1) / m * alpha
            const double thr = __dmul_rn(__ddiv_rn((double)(i + 1), md), alpha);
            if (__longlong_as_double((long long)key) <= thr) best = (unsigned long long)(i + 1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        unsigned long long b2 = __shfl_xor_sync(0xffffffffu, best, o);
        best = b2 > best ? b2 : best;
    }
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = mn;
        if (best) atomicMax(kstar, best);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = s_min[0];
        for (int w = 1; w < BH_THREADS / 32; ++w) v = fmin(v, s_min[w]);
        tile_min[blockIdx.x] = v;
    }
}

// suffix-min over tile minima: tile_suffix[t] = min over tiles > t   (single block)
__global__ void __launch_bounds__(256) bh_suffix_kernel(double* __restrict__ tile_min, int n_tiles) {
    // chunked backwards scan; n_tiles is m/2048 so even 1e8 tests is 5e4 entries
    __shared__ double s_part[256];
    const int per = (n_tiles + 255) / 256;
    const int lo = threadIdx.x * per, hi = min(lo + per, n_tiles);
    double mn = INFINITY;
    for (int t = hi - 1; t >= lo; --t) mn = fmin(mn, tile_min[t]);
    s_part[threadIdx.x] = mn;
    __syncthreads();
    double carry = INFINITY;  // min over all chunks after mine
    for (int c = threadIdx.x + 1; c < 256; ++c) carry = fmin(carry, s_part[c]);
    __syncthreads();
    for (int t = hi - 1; t >= lo; --t) {
        const double v = tile_min[t];
        tile_min[t] = carry;  // exclusive suffix min
        carry = fmin(carry, v);
    }
}

// apply: in-tile reverse running min (+ carry from later tiles), clip, scatter
__global__ void __launch_bounds__(BH_THREADS)
bh_apply_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, int64_t m,
                const double* __restrict__ tile_suffix, const unsigned long long* __restrict__ kstar,
                uint8_t* __restrict__ rejected, double* __restrict__ q) {
    __shared__ double s_w[BH_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * BH_TILE + (int64_t)threadIdx.x * BH_ITEMS;  // blocked arrangement
    const double md = (double)m;
    double v[BH_ITEMS];
    double mn = INFINITY;
#pragma unroll
    for (int j = BH_ITEMS - 1; j >= 0; --j) {
        const int64_t i = base + j;
        v[j] = (i < m) ? bh_raw_adjusted(keys[i], i, md) : INFINITY;
        mn = fmin(mn, v[j]);
    }
    // suffix-min across threads of the block: thread t needs min over threads > t
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double incl = mn;  // inclusive suffix within warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl = fmin(incl, t);
    }
    if (lane == 0) s_w[warp] = incl;
    __syncthreads();
    double after = tile_suffix[blockIdx.x];
    for (int w = warp + 1; w < BH_THREADS / 32; ++w) after = fmin(after, s_w[w]);
    double next = __shfl_down_sync(0xffffffffu, incl, 1);
    if (lane < 31) after = fmin(after, next);
    const unsigned long long ks = *kstar;
    double run = after;
#pragma unroll
    for (int j = BH_ITEMS - 1; j >= 0; --j) {
        const int64_t i = base + j;
        run = fmin(run, v[j]);
        if (i < m) {
            const uint32_t dst = idx[i];
            q[dst] = fmin(fmax(run, 0.0), 1.0);
            rejected[dst] = (unsigned long long)(i + 1) <= ks ? 1 : 0;
        }
    }
}

// Workspace of one BH run over up to m tests.  A plan owns one, so that its step allocates nothing, copies nothing
// between host and device and never synchronises (the pmrd_plan contract; it also keeps the step graph-capturable).
size_t bh_workspace_bytes(int64_t m) {
    if (m < 1) m = 1;
    const size_t a = ((size_t)m * 8 + 255) & ~(size_t)255, b = ((size_t)m * 4 + 255) & ~(size_t)255;
    const size_t tiles = (size_t)((m + BH_TILE - 1) / BH_TILE);
    return 2 * a + 2 * b + ((tiles * 8 + 255) & ~(size_t)255) + 256 + radix_scratch_bytes(m);
}

int32_t k9_bh(pmrd_ctx* ctx, const double* d_p, int64_t m, double alpha, uint8_t* d_rejected, double* d_q, void* d_workspace) {
    if (m <= 0) return PMRD_OK;
    PMRD_ARG(ctx, m < (1ll << 31), "BH: m must be < 2^31");
    const size_t a = ((size_t)m * 8 + 255) & ~(size_t)255, b = ((size_t)m * 4 + 255) & ~(size_t)255;
    const int n_tiles = (int)((m + BH_TILE - 1) / BH_TILE);
    DevBuf own;
    if (!d_workspace) PMRD_CUDA(ctx, own.alloc(ctx->stream, bh_workspace_bytes(m)));
    unsigned char* w = d_workspace ? (unsigned char*)d_workspace : own.as<unsigned char>();
    uint64_t* ka = (uint64_t*)w; w += a;
    uint64_t* kb = (uint64_t*)w; w += a;
    uint32_t* va = (uint32_t*)w; w += b;
    uint32_t* vb = (uint32_t*)w; w += b;
    double* tmin = (double*)w; w += ((size_t)n_tiles * 8 + 255) & ~(size_t)255;
    unsigned long long* kstar = (unsigned long long*)w; w += 256;
    uint32_t* rs_scratch = (uint32_t*)w;
    PMRD_LAUNCH(ctx, bh_pack_kernel, pmrd_blocks(m, 256), 256, 0, d_p, m, ka, va);
    // Digit windows: with a caller-owned workspace (plans) the mask is FIXED -- p lies in [0, 1], so its bit pattern
    // is at most 0x3FF0...0 and bits 0..61 may vary: no reduction, no device -> host copy, no sync.  Otherwise the
    // varying bits are measured first (one sync; fewer passes for the host-buffer entry points).
    uint64_t varying = (1ull << 62) - 1ull;
    if (!d_workspace) PMRD_TRY(radix_varying_bits(ctx, ka, m, &varying));
    int in_b = 0;
    PMRD_TRY(radix_sort_pairs(ctx, ka, va, kb, vb, m, varying, &in_b, rs_scratch));
    const uint64_t* keys = in_b ? kb : ka;
    const uint32_t* idx = in_b ? vb : va;
    PMRD_CUDA(ctx, cudaMemsetAsync(kstar, 0, 8, ctx->stream));
    PMRD_LAUNCH(ctx, bh_tile_kernel, n_tiles, BH_THREADS, 0, keys, m, alpha, tmin, kstar);
    PMRD_LAUNCH(ctx, bh_suffix_kernel, 1, 256, 0, tmin, n_tiles);
    PMRD_LAUNCH(ctx, bh_apply_kernel, n_tiles, BH_THREADS, 0, keys, idx, m, (const double*)tmin, (const unsigned long long*)kstar,
                d_rejected, d_q);
    return PMRD_OK;
}

// ---------------------------------------------------------------- K9 per call group --------
// Benjamini-Hochberg inside every call group (cells [gs[g], gs[g+1]), at most BHG_MAX of them): one block per group.
// Ranks by counting (ties by index: tied p-values get the same adjusted value and never straddle the rejection
// cut, Appendix B of SURVEY.md), then exactly the arithmetic of the global path -- p * m / rank in that order, reverse
// running minimum, clip -- so a group's q-values are bit-equal to pmrd_bh run on that group alone.
#define BHG_MAX 4096
#define BHG_THREADS 1024
__global__ void __launch_bounds__(BHG_THREADS)
bh_group_kernel(const double* __restrict__ p, const int64_t* __restrict__ gs, double alpha, uint8_t* __restrict__ rejected, double* __restrict__ q) {
    extern __shared__ __align__(16) unsigned char bhg_raw[];
    double* const s_in = reinterpret_cast<double*>(bhg_raw);           // [BHG_MAX] the group's p-values (every thread scans all of them)
    double* const s_p = s_in + BHG_MAX;                                // [BHG_MAX] sorted p, then raw adjusted values
    uint16_t* const s_src = reinterpret_cast<uint16_t*>(s_p + BHG_MAX); // [BHG_MAX] sorted position -> cell of the group
    __shared__ unsigned int s_kstar;
    const int64_t lo = gs[blockIdx.x];
    const int m = (int)(gs[blockIdx.x + 1] - lo);
    if (m <= 0) return;
    if (threadIdx.x == 0) s_kstar = 0u;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        double v = p[lo + i];
        if (v == 0.0) v = 0.0;                                   // -0.0 folded onto +0.0, as bh_pack_kernel does
        s_in[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double v = s_in[i];
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const double w = s_in[j];                            // (same address for the whole warp: a broadcast)
            rank += (w < v) || (w == v && j < i);
        }
        s_p[rank] = v;
        s_src[rank] = (uint16_t)i;
    }
    __syncthreads();
    const double md = (double)m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double v = s_p[i];
        const double thr = __dmul_rn(__ddiv_rn((double)(i + 1), md), alpha);      // call.py:55
        if (v <= thr) atomicMax(&s_kstar, (unsigned int)(i + 1));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_p[i] = __ddiv_rn(__dmul_rn(s_p[i], md), (double)(i + 1));   // call.py:59-61
    __syncthreads();
    // reverse running minimum (call.py:64-70): warp 0, each lane a contiguous slice, then the carries from the slices after it
    if (threadIdx.x < 32) {
        const int per = (m + 31) / 32;
        const int a = (int)threadIdx.x * per, b = min(a + per, m);
        double mn = INFINITY;
        for (int i = b - 1; i >= a; --i) mn = fmin(mn, s_p[i]);
        double carry = INFINITY;                                  // min over the slices after mine
#pragma unroll
        for (int l = 31; l >= 0; --l) {
            const double other = __shfl_sync(0xffffffffu, mn, l);
            if (l > (int)threadIdx.x) carry = fmin(carry, other);
        }
        double run = carry;
        for (int i = b - 1; i >= a; --i) {
            run = fmin(run, s_p[i]);
            s_p[i] = run;
        }
    }
    __syncthreads();
    const unsigned int ks = s_kstar;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int64_t dst = lo + s_src[i];
        q[dst] = fmin(fmax(s_p[i], 0.0), 1.0);
        rejected[dst] = (unsigned int)(i + 1) <= ks ? 1 : 0;
    }
}

int32_t k9_bh_groups(pmrd_ctx* ctx, const double* d_p, const int64_t* d_gs, int64_t n_groups, double alpha, uint8_t* d_rejected, double* d_q) {
    if (n_groups <= 0) return PMRD_OK;
    const size_t smem = (size_t)BHG_MAX * (8 + 8 + 2);
    PMRD_CUDA(ctx, cudaFuncSetAttribute(bh_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PMRD_LAUNCH(ctx, bh_group_kernel, (int)n_groups, BHG_THREADS, smem, d_p, d_gs, alpha, d_rejected, d_q);
    return PMRD_OK;
}
