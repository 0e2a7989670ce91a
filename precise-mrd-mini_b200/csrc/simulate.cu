// simulate.cu -- the G-D (sample-level) Monte-Carlo generators on Philox:
//   K1  sim_cells        simulate_reads        src/precise_mrd/simulate.py:133-180
//   K2  sim_gd_families  collapse_umis         src/precise_mrd/collapse.py:70-118 (v2),
//                                              "collapse 2.py":54-78 (v1)
//   K10 error model      fit_error_model       src/precise_mrd/error_model.py:139-186
// The reference threads ONE PCG64 stream through all of these; here every draw is keyed
// (seed; stream, cell, family, draw), so results do not depend on GPU count or call order.
// Parity class: distributional (KS / chi-square at alpha = 0.01, tests/test_gpu_simulate.py).
#include "kernels.cuh"
#include "rng.cuh"
#include "scan.cuh"

// ---------------------------------------------------------------- K1 --------------------
__global__ void __launch_bounds__(128)
sim_cells_kernel(uint64_t seed, const int64_t* __restrict__ cell_id, const double* __restrict__ af,
                 const int64_t* __restrict__ depth, const double* __restrict__ bg_mult /* or NULL */, int64_t n_cells,
                 int max_family_size, pmrd_cell_table out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const Philox ph(seed);
    const uint64_t cid = (uint64_t)cell_id[i];
    const uint32_t c2 = (uint32_t)cid, c3 = (PMRD_STREAM_CELLS << 24) | (uint32_t)((cid >> 32) & 0xffffffu);
    const uint4 b0 = ph(0, 0, c2, c3), b1 = ph(1, 0, c2, c3), b2 = ph(2, 0, c2, c3);
    // simulate.py:144  background_rates = rng.uniform(1e-5, 1e-3)
    double bg = 1e-5 + (1e-3 - 1e-5) * u53(b0.x, b0.y);
    // eval/stratified.py:186-201: the context multiplier scales the stored rate, and the false
    // positives are drawn at rate * multiplier again (the reference applies it twice)
    const double m = bg_mult ? bg_mult[i] : 1.0;
    bg *= m;
    const double fp_rate = fmin(bg * m, 1.0);
    // simulate.py:147-148  clip(poisson(5), 1, max_family_size)
    int fs = sample_poisson_small(u53(b0.z, b0.w), 5.0);
    fs = fs < 1 ? 1 : (fs > max_family_size ? max_family_size : fs);
    // simulate.py:152,155-158
    const int64_t d = depth[i];
    const int64_t n_true = sample_binomial(u53(b1.x, b1.y), d, af[i]);
    const int64_t n_fp = sample_binomial(u53(b1.z, b1.w), d - n_true, fp_rate);
    // simulate.py:161-162  clip(normal(25, 5), 10, 40)
    double q = 25.0 + 5.0 * sample_normal(u53(b2.x, b2.y), u53(b2.z, b2.w));
    q = fmin(fmax(q, 10.0), 40.0);
    out.background_rate[i] = bg;
    out.mean_family_size[i] = fs;
    out.n_true_variants[i] = n_true;
    out.n_false_positives[i] = n_fp;
    out.mean_quality[i] = q;
}

int32_t k1_simulate_cells(pmrd_ctx* ctx, const int64_t* d_cell_id, const double* d_af, const int64_t* d_depth,
                          const double* d_bg_mult, int64_t n_cells, int32_t max_family_size, const pmrd_cell_table* d_out) {
    if (n_cells <= 0) return PMRD_OK;
    PMRD_LAUNCH(ctx, sim_cells_kernel, pmrd_blocks(n_cells, 128), 128, 0, ctx->seed, d_cell_id, d_af, d_depth, d_bg_mult,
                n_cells, max_family_size, *d_out);
    return PMRD_OK;
}

// ---------------------------------------------------------------- K2 --------------------
#define GD_TILE 1024 /* families per tile; one 256-thread block, 4 families per thread */

struct GdCell {
    uint64_t cell_id;
    double af, bg, lambda;
};

struct GdFamily {
    int size;
    double quality, agreement;
    uint32_t flags;
    bool kept;
};

__device__ __forceinline__ int gd_family_size(const Philox& ph, uint32_t fam, uint32_t c2, uint32_t c3, double lambda,
                                              int max_fs, uint4& b0) {
    b0 = ph(0, fam, c2, c3);
    // collapse.py:80-82  min(max(1, int(poisson(lambda))), max_family_size)
    int fs = sample_poisson_small(u53(b0.x, b0.y), lambda);
    fs = fs < 1 ? 1 : fs;
    return fs > max_fs ? max_fs : fs;
}

__device__ __forceinline__ GdFamily gd_family(const Philox& ph, uint32_t fam, const GdCell& c, const pmrd_umi_params& P, int variant) {
    GdFamily f;
    const uint32_t c2 = (uint32_t)c.cell_id, c3 = (PMRD_STREAM_GD_FAMILY << 24) | (uint32_t)((c.cell_id >> 32) & 0xffffffu);
    uint4 b0;
    f.size = gd_family_size(ph, fam, c2, c3, c.lambda, P.max_family_size, b0);
    f.kept = f.size >= P.min_family_size;  // collapse.py:85-86: skipped BEFORE any other draw
    f.quality = f.agreement = 0.0;
    f.flags = 0;
    if (!f.kept) return f;
    const uint4 b1 = ph(1, fam, c2, c3), b2 = ph(2, fam, c2, c3);
    // collapse.py:90-92  clip(normal(25 + 100*AF, 3), 10, 40)      (v1: normal(25, 3))
    const double base_q = variant == 1 ? 25.0 : 25.0 + c.af * 100.0;
    double q = base_q + 3.0 * sample_normal(u53(b1.x, b1.y), u53(b1.z, b1.w));
    q = fmin(fmax(q, 10.0), 40.0);
    const bool pass_q = q >= (double)P.quality_threshold;
    // collapse.py:99-101  uniform(min(0.95, 0.7 + (q-10)/60), 1)    (v1: uniform(0.5, 1))
    const double cb = variant == 1 ? 0.5 : fmin(0.95, 0.7 + (q - 10.0) / 60.0);
    const double agree = cb + (1.0 - cb) * u53(b0.z, b0.w);
    const bool pass_c = agree >= P.consensus_threshold;
    // collapse.py:104-106
    const bool is_true = c.af > 0.01 && u53(b2.x, b2.y) < 0.8;
    const bool is_bg = u53(b2.z, b2.w) < c.bg;
    const bool is_var = pass_q && pass_c && (is_true || is_bg);
    f.quality = q;
    f.agreement = agree;
    f.flags = (pass_q ? 1u : 0u) | (pass_c ? 2u : 0u) | (is_var ? 4u : 0u);
    return f;
}

// tile t covers families [tile_fam0[t], min(+GD_TILE, n_families)) of cell tile_cell[t]
__global__ void __launch_bounds__(256)
gd_count_kernel(uint64_t seed, const int32_t* __restrict__ tile_cell, const int32_t* __restrict__ tile_fam0,
                const GdCell* __restrict__ cells, const int64_t* __restrict__ n_families, pmrd_umi_params P,
                uint32_t* __restrict__ tile_count) {
    __shared__ uint32_t sw[8];
    const Philox ph(seed);
    const int ci = tile_cell[blockIdx.x];
    const GdCell c = cells[ci];
    const int64_t nf = n_families[ci];
    const uint32_t c2 = (uint32_t)c.cell_id, c3 = (PMRD_STREAM_GD_FAMILY << 24) | (uint32_t)((c.cell_id >> 32) & 0xffffffu);
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < GD_TILE / 256; ++j) {
        const int64_t fam = (int64_t)tile_fam0[blockIdx.x] + j * 256 + threadIdx.x;
        if (fam < nf) {
            uint4 b0;
            cnt += gd_family_size(ph, (uint32_t)fam, c2, c3, c.lambda, P.max_family_size, b0) >= P.min_family_size;
        }
    }
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += sw[w];
        tile_count[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
gd_write_kernel(uint64_t seed, const int32_t* __restrict__ tile_cell, const int32_t* __restrict__ tile_fam0,
                const GdCell* __restrict__ cells, const int64_t* __restrict__ n_families, pmrd_umi_params P, int variant,
                const uint64_t* __restrict__ tile_base, pmrd_gd_family_table out) {
    __shared__ uint32_t sw[33];
    const Philox ph(seed);
    const int ci = tile_cell[blockIdx.x];
    const GdCell c = cells[ci];
    const int64_t nf = n_families[ci];
    // blocked: thread t owns families fam0 + 4t .. 4t+3 so that ranks follow family order
    GdFamily f[GD_TILE / 256];
    uint32_t cnt = 0;
    const int64_t fam0 = (int64_t)tile_fam0[blockIdx.x] + (int64_t)threadIdx.x * (GD_TILE / 256);
#pragma unroll
    for (int j = 0; j < GD_TILE / 256; ++j) {
        f[j].kept = false;
        if (fam0 + j < nf) f[j] = gd_family(ph, (uint32_t)(fam0 + j), c, P, variant);
        cnt += f[j].kept ? 1u : 0u;
    }
    uint32_t tot;
    uint64_t r = tile_base[blockIdx.x] + block_exclusive_scan<uint32_t>(cnt, sw, tot);
#pragma unroll
    for (int j = 0; j < GD_TILE / 256; ++j) {
        if (f[j].kept) {
            out.cell[r] = ci;
            out.family_id[r] = (int32_t)(fam0 + j);
            out.family_size[r] = f[j].size;
            out.quality_score[r] = f[j].quality;
            out.consensus_agreement[r] = f[j].agreement;
            out.flags[r] = (uint8_t)f[j].flags;
            ++r;
        }
    }
}

// d_tile_base (out): exclusive scan of kept rows per tile (+ total in [n_tiles])
int32_t k2_gd_families(pmrd_ctx* ctx, const int32_t* d_tile_cell, const int32_t* d_tile_fam0, int64_t n_tiles,
                       const void* d_cells, const int64_t* d_n_families, const pmrd_umi_params* P, int variant,
                       uint64_t* d_tile_base, const pmrd_gd_family_table* d_out /* null: count only */) {
    if (n_tiles <= 0) return PMRD_OK;
    if (!d_out) {
        DevBuf cnt;
        PMRD_CUDA(ctx, cnt.alloc(ctx->stream, (size_t)n_tiles * 4));
        PMRD_LAUNCH(ctx, gd_count_kernel, (int)n_tiles, 256, 0, ctx->seed, d_tile_cell, d_tile_fam0, (const GdCell*)d_cells,
                    d_n_families, *P, cnt.as<uint32_t>());
        PMRD_TRY((exclusive_scan<uint32_t, uint64_t>(ctx, cnt.as<uint32_t>(), d_tile_base, n_tiles, d_tile_base + n_tiles)));
        return PMRD_OK;
    }
    PMRD_LAUNCH(ctx, gd_write_kernel, (int)n_tiles, 256, 0, ctx->seed, d_tile_cell, d_tile_fam0, (const GdCell*)d_cells,
                d_n_families, *P, variant, (const uint64_t*)d_tile_base, *d_out);
    return PMRD_OK;
}

// ---------------------------------------------------------------- K10 -------------------
// One block per trinucleotide context; thread b draws bootstrap resample b.
// Resampling n_neg rows with replacement and taking mean(is_variant) is exactly
// Binomial(n_neg, n_var/n_neg) / n_neg (error_model.py:165-170).
__global__ void __launch_bounds__(256)
error_model_kernel(uint64_t seed, uint64_t stream_id, int64_t n_neg, int64_t n_var, int n_boot,
                   double* __restrict__ rate64, double* __restrict__ ci_lo64, double* __restrict__ ci_hi64) {
    extern __shared__ double s_boot[];  // n_boot
    const Philox ph(seed);
    const uint32_t ctxi = blockIdx.x;
    const uint32_t c2 = (uint32_t)stream_id, c3 = (PMRD_STREAM_ERRMODEL << 24) | (uint32_t)((stream_id >> 32) & 0xffffffu);
    const uint4 bm = ph(0, ctxi, c2, c3);
    double rate, modifier = 0.5 + 1.5 * u53(bm.x, bm.y);  // error_model.py:155 uniform(0.5, 2.0)
    const double vr = n_neg > 0 ? (double)n_var / (double)n_neg : 0.0;
    if (n_neg > 0) rate = vr * modifier;
    else rate = 1e-5 + (1e-3 - 1e-5) * u53(bm.z, bm.w);  // error_model.py:159
    if (n_neg > 10) {
        for (int b = threadIdx.x; b < n_boot; b += blockDim.x) {
            const uint4 r = ph(1 + (uint32_t)b, ctxi, c2, c3);
            const int64_t k = sample_binomial(u53(r.x, r.y), n_neg, vr);
            s_boot[b] = ((double)k / (double)n_neg) * modifier;
        }
        __syncthreads();
        // np.percentile(linear): rank-by-counting selection of the two bracketing order statistics
        if (threadIdx.x < 2) {
            const double qq = threadIdx.x == 0 ? 2.5 : 97.5;
            const double pos = qq / 100.0 * (double)(n_boot - 1);
            const int lo = (int)floor(pos), hi = min(lo + 1, n_boot - 1);
            const double frac = pos - (double)lo;
            double vlo = 0.0, vhi = 0.0;
            for (int i = 0; i < n_boot; ++i) {
                int rank = 0, ties = 0;
                const double v = s_boot[i];
                for (int j = 0; j < n_boot; ++j) {
                    rank += s_boot[j] < v;
                    ties += s_boot[j] == v;
                }
                if (rank <= lo && lo < rank + ties) vlo = v;
                if (rank <= hi && hi < rank + ties) vhi = v;
            }
            // numpy's lerp: a + (b - a) * t  (t < 0.5 branch is what 2.5 / 97.5 of 100 hit; same value either way to 1 ulp)
            const double val = vlo + (vhi - vlo) * frac;
            if (threadIdx.x == 0) ci_lo64[ctxi] = val;
            else ci_hi64[ctxi] = val;
        }
    } else if (threadIdx.x == 0) {
        ci_lo64[ctxi] = rate * 0.5;  // error_model.py:176-177
        ci_hi64[ctxi] = rate * 2.0;
    }
    if (threadIdx.x == 0) rate64[ctxi] = rate;
}

int32_t k10_error_model(pmrd_ctx* ctx, int64_t n_neg, int64_t n_var, int32_t n_boot, uint64_t stream_id, double* d_rate64,
                        double* d_lo64, double* d_hi64) {
    PMRD_ARG(ctx, n_boot >= 1 && n_boot <= 4096, "n_boot must be in [1, 4096]");
    PMRD_ARG(ctx, n_neg >= 0 && n_var >= 0 && n_var <= n_neg, "0 <= n_var <= n_neg");
    PMRD_LAUNCH(ctx, error_model_kernel, 64, 256, (size_t)n_boot * 8, ctx->seed, stream_id, n_neg, n_var, (int)n_boot, d_rate64,
                d_lo64, d_hi64);
    return PMRD_OK;
}
