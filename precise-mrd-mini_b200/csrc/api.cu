// api.cu -- the extern "C" boundary (include/pmrd.h): context management and the
// host-buffer wrappers that stage data through device memory around the K* kernels.
#include <math.h>
#include <new>

#include "kernels.cuh"
#include "radix_sort.cuh"

char g_pmrd_create_err[512] = "";

// ---------------------------------------------------------------- per-kernel timing -----
// CUDA events recorded around every launch on the launching stream while enabled; the
// report aggregates by kernel name.  bench.py uses it INSIDE the timed region to get the
// dominant kernel's average launch duration for the roofline line.
#include <map>
#include <string>
#include <vector>
struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
};
struct PmrdProf {
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t take() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};
void pmrd_prof_begin(pmrd_ctx* ctx, const char* kernel_name) {
    PmrdProf* p = (PmrdProf*)ctx->prof;
    if (!p) return;
    ProfRec r{kernel_name, p->take(), p->take()};
    cudaEventRecord(r.a, ctx->stream);
    p->recs.push_back(r);
}
void pmrd_prof_end(pmrd_ctx* ctx) {
    PmrdProf* p = (PmrdProf*)ctx->prof;
    if (!p || p->recs.empty()) return;
    cudaEventRecord(p->recs.back().b, ctx->stream);
}

int32_t collapse_upload_weights(pmrd_ctx* ctx, const double* h_w64);
int32_t fused_upload_tables(pmrd_ctx* ctx, const double* h_w64);
int32_t k6_hamming(pmrd_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, int64_t n, int32_t* d_d);
int32_t k2_gd_families(pmrd_ctx* ctx, const int32_t* d_tile_cell, const int32_t* d_tile_fam0, int64_t n_tiles,
                       const void* d_cells, const int64_t* d_n_families, const pmrd_umi_params* P, int variant,
                       uint64_t* d_tile_base, const pmrd_gd_family_table* d_out);
int32_t k10_error_model(pmrd_ctx* ctx, int64_t n_neg, int64_t n_var, int32_t n_boot, uint64_t stream_id, double* d_rate64,
                        double* d_lo64, double* d_hi64);
int32_t k11_fit_lod_batch(pmrd_ctx* ctx, const double* d_af, const double* d_hits, int32_t n, int32_t n_series, double target,
                          int32_t n_boot, const uint64_t* d_stream_ids, double* d_out3);
int32_t call_counts_dev(pmrd_ctx* ctx, const int64_t* d_tot, const int64_t* d_var, const uint8_t* d_neg, int64_t n_cells,
                        uint64_t stream_id, int32_t test_type, int32_t alternative, double alpha, double* d_mean_rate,
                        int64_t* d_negc, double* d_p, double* d_q, uint8_t* d_sig, void* d_bh_workspace);
int32_t k15_gd_features(pmrd_ctx* ctx, const int64_t* d_size, const double* d_q, const double* d_agr, const double* d_af, int64_t n,
                        double* d_out);
int32_t k16_qc_family(pmrd_ctx* ctx, const uint32_t* d_size, const uint32_t* d_flags, uint32_t flag_mask, int64_t n, int n_bins,
                      unsigned long long* d_hist, unsigned long long* d_counts);
struct pmrd_plan;
int32_t plan_create(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                    int64_t n_cells, const pmrd_read_sim_params* sp, const pmrd_umi_params* up, double neg_af_max,
                    int32_t test_type, int32_t alternative, double alpha, const pmrd_cell_knobs* knobs, pmrd_plan** out);
int32_t plan_group_rates(pmrd_ctx* ctx, pmrd_plan* P, double* h_rates, int64_t cap);
int32_t plan_run(pmrd_ctx* ctx, pmrd_plan* P);
int32_t plan_run_simulate(pmrd_ctx* ctx, pmrd_plan* P, int64_t which);
int32_t plan_run_collapse(pmrd_ctx* ctx, pmrd_plan* P, int64_t which);
int64_t plan_n_chunks(const pmrd_plan* P);
int64_t plan_chunk_reads(const pmrd_plan* P, int64_t which);
int32_t plan_info(const pmrd_plan* P, int64_t* o);
int32_t plan_run_call(pmrd_ctx* ctx, pmrd_plan* P);
int32_t plan_results(pmrd_ctx* ctx, pmrd_plan* P, const pmrd_pipeline_result* h, double* h_mean_rate, int64_t* h_totals);
void plan_destroy(pmrd_plan* P);
int64_t plan_n_reads(const pmrd_plan* P);
int64_t plan_n_families(const pmrd_plan* P);
#include "reads.cuh"

extern "C" {

int32_t pmrd_abi_version(void) { return PMRD_ABI_VERSION; }

int32_t pmrd_create(int32_t device, uint64_t seed, pmrd_ctx** out) {
    if (!out) return PMRD_ERR_ARG;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev <= 0) {
        snprintf(g_pmrd_create_err, sizeof(g_pmrd_create_err),
                 "pmrd_create: no CUDA device visible (%s); libpmrd_b200 has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return PMRD_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n_dev) {
        snprintf(g_pmrd_create_err, sizeof(g_pmrd_create_err), "pmrd_create: device %d out of range [0,%d)", device, n_dev);
        return PMRD_ERR_ARG;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        snprintf(g_pmrd_create_err, sizeof(g_pmrd_create_err), "pmrd_create: %s", cudaGetErrorString(e));
        return PMRD_ERR_CUDA;
    }
    if (prop.major != 10) {
        snprintf(g_pmrd_create_err, sizeof(g_pmrd_create_err),
                 "pmrd_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return PMRD_ERR_NO_DEVICE;
    }
    pmrd_ctx* ctx = new (std::nothrow) pmrd_ctx();
    if (!ctx) return PMRD_ERR_ARG;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : PMRD_SM_COUNT_DEFAULT;
    ctx->seed = seed;
    ctx->launches = 0;
    ctx->prof_on = 0;
    ctx->prof = nullptr;
    ctx->stager = nullptr;
    ctx->err[0] = 0;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        snprintf(g_pmrd_create_err, sizeof(g_pmrd_create_err), "pmrd_create: %s", cudaGetErrorString(e));
        delete ctx;
        return PMRD_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    // keep freed scratch in the stream-ordered pool instead of returning it to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    // 10 ** (q / 10.0) exactly as CPython evaluates it: libm pow on doubles (umi.py:161)
    double w[64];
    for (int q = 0; q < 64; ++q) w[q] = pow(10.0, (double)q / 10.0);
    int32_t s = collapse_upload_weights(ctx, w);
    if (s == PMRD_OK) s = reads_upload_tables(ctx);
    if (s == PMRD_OK) s = fused_upload_tables(ctx, w);
    if (s == PMRD_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) s = PMRD_ERR_CUDA;
    if (s != PMRD_OK) {
        snprintf(g_pmrd_create_err, sizeof(g_pmrd_create_err), "pmrd_create: %s", ctx->err);
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return s;
    }
    *out = ctx;
    return PMRD_OK;
}

void pmrd_destroy(pmrd_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pmrd_stager_destroy(ctx);
    cudaStreamDestroy(ctx->own_stream);
    if (ctx->prof) {
        PmrdProf* p = (PmrdProf*)ctx->prof;
        for (ProfRec& r : p->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        for (cudaEvent_t e : p->pool) cudaEventDestroy(e);
        delete p;
    }
    delete ctx;
}

const char* pmrd_last_error(const pmrd_ctx* ctx) { return ctx ? ctx->err : g_pmrd_create_err; }

int32_t pmrd_set_seed(pmrd_ctx* ctx, uint64_t seed) {
    if (!ctx) return PMRD_ERR_ARG;
    ctx->seed = seed;
    return PMRD_OK;
}

int32_t pmrd_set_stream(pmrd_ctx* ctx, void* cuda_stream) {
    if (!ctx) return PMRD_ERR_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PMRD_OK;
}

int32_t pmrd_sync(pmrd_ctx* ctx) {
    if (!ctx) return PMRD_ERR_ARG;
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

int64_t pmrd_launch_count(const pmrd_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t pmrd_peak_memory_bytes(pmrd_ctx* ctx) {
    if (!ctx) return 0;
    cudaMemPool_t pool;
    uint64_t high = 0;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) != cudaSuccess) return 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &high) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (int64_t)high;
}

int32_t pmrd_profile_enable(pmrd_ctx* ctx, int32_t on) {
    if (!ctx) return PMRD_ERR_ARG;
    if (on && !ctx->prof) ctx->prof = new (std::nothrow) PmrdProf();
    ctx->prof_on = (on && ctx->prof) ? 1 : 0;
    return PMRD_OK;
}

// Synchronises, then writes one line per kernel: "<name> <launches> <total_ms> <min_ms> <max_ms>\n" and clears
// the records.  Returns the number of bytes written (truncated to cap-1).
int64_t pmrd_profile_report(pmrd_ctx* ctx, char* buf, int64_t cap) {
    if (!ctx || !buf || cap <= 0) return 0;
    buf[0] = 0;
    PmrdProf* p = (PmrdProf*)ctx->prof;
    if (!p) return 0;
    cudaStreamSynchronize(ctx->stream);
    struct Agg { long long n = 0; double total = 0.0, mn = 1e300, mx = 0.0; };
    std::map<std::string, Agg> agg;
    std::vector<std::string> order;
    for (ProfRec& r : p->recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        std::string name(r.name);
        if (name.size() && name[0] == '(') name = name.substr(1, name.size() - 2);
        if (!agg.count(name)) order.push_back(name);
        Agg& a = agg[name];
        a.n += 1;
        a.total += ms;
        if (ms < a.mn) a.mn = ms;
        if (ms > a.mx) a.mx = ms;
        p->pool.push_back(r.a);
        p->pool.push_back(r.b);
    }
    p->recs.clear();
    cudaGetLastError();
    int64_t off = 0;
    for (const std::string& n : order) {
        char line[320];
        const Agg& a = agg[n];
        const int w = snprintf(line, sizeof(line), "%s %lld %.6f %.6f %.6f\n", n.c_str(), a.n, a.total, a.mn, a.mx);
        if (off + w >= cap) break;
        memcpy(buf + off, line, (size_t)w);
        off += w;
    }
    buf[off] = 0;
    return off;
}

}  // extern "C"

// ---------------------------------------------------------------- staging helpers -------
#define ENTER(ctx)                              \
    if (!(ctx)) return PMRD_ERR_ARG;            \
    (ctx)->err[0] = 0;                          \
    PMRD_CUDA(ctx, cudaSetDevice((ctx)->device))

static int32_t up(pmrd_ctx* ctx, DevBuf& b, const void* h, size_t bytes) {
    PMRD_CUDA(ctx, b.alloc(ctx->stream, bytes));
    return pmrd_stage_up(ctx, b.p, h, bytes);     // (large pageable buffers go through the pinned ring: staging.cu)
}
static int32_t down(pmrd_ctx* ctx, void* h, const void* d, size_t bytes) {
    return pmrd_stage_down(ctx, h, d, bytes);
}

extern "C" {

// ---------------------------------------------------------------- K8 --------------------
int32_t pmrd_pvalues_dev(pmrd_ctx* ctx, const int64_t* d_k, const int64_t* d_n, const double* d_rate, int64_t rate_stride,
                         int64_t m, int32_t test_type, int32_t alternative, double* d_p) {
    ENTER(ctx);
    return k8_pvalues(ctx, d_k, d_n, d_rate, rate_stride, m, test_type, alternative, d_p);
}

int32_t pmrd_pvalues(pmrd_ctx* ctx, const int64_t* h_k, const int64_t* h_n, const double* h_rate, int64_t rate_stride,
                     int64_t m, int32_t test_type, int32_t alternative, double* h_p) {
    ENTER(ctx);
    PMRD_ARG(ctx, m >= 0, "m");
    if (m == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_k && h_rate && h_p, "null buffer");
    PMRD_ARG(ctx, rate_stride == 0 || rate_stride == 1, "rate_stride must be 0 or 1");
    DevBuf k, n, r, p;
    PMRD_TRY(up(ctx, k, h_k, (size_t)m * 8));
    if (h_n) PMRD_TRY(up(ctx, n, h_n, (size_t)m * 8));
    PMRD_TRY(up(ctx, r, h_rate, (size_t)(rate_stride ? m : 1) * 8));
    PMRD_CUDA(ctx, p.alloc(ctx->stream, (size_t)m * 8));
    PMRD_TRY(k8_pvalues(ctx, k.as<int64_t>(), h_n ? n.as<int64_t>() : nullptr, r.as<double>(), rate_stride, m, test_type, alternative, p.as<double>()));
    PMRD_TRY(down(ctx, h_p, p.p, (size_t)m * 8));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- K9 --------------------
int32_t pmrd_bh_dev(pmrd_ctx* ctx, const double* d_p, int64_t m, double alpha, uint8_t* d_rejected, double* d_q) {
    ENTER(ctx);
    return k9_bh(ctx, d_p, m, alpha, d_rejected, d_q);
}

int32_t pmrd_bh(pmrd_ctx* ctx, const double* h_p, int64_t m, double alpha, uint8_t* h_rejected, double* h_q) {
    ENTER(ctx);
    PMRD_ARG(ctx, m >= 0, "m");
    if (m == 0) return PMRD_OK;  // call.py:46-47: empty in, empty out
    PMRD_ARG(ctx, h_p && h_rejected && h_q, "null buffer");
    DevBuf p, rej, q;
    PMRD_TRY(up(ctx, p, h_p, (size_t)m * 8));
    PMRD_CUDA(ctx, rej.alloc(ctx->stream, (size_t)m));
    PMRD_CUDA(ctx, q.alloc(ctx->stream, (size_t)m * 8));
    PMRD_TRY(k9_bh(ctx, p.as<double>(), m, alpha, rej.as<uint8_t>(), q.as<double>()));
    PMRD_TRY(down(ctx, h_rejected, rej.p, (size_t)m));
    PMRD_TRY(down(ctx, h_q, q.p, (size_t)m * 8));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- call ------------------
int32_t pmrd_call_dev(pmrd_ctx* ctx, const int64_t* d_seg_offsets, int64_t n_segments, const uint8_t* d_is_variant,
                      const double* d_quality, const double* d_agreement, double mean_error_rate, int32_t test_type,
                      double alpha, pmrd_call_table* d_out) {
    ENTER(ctx);
    PMRD_ARG(ctx, d_out != nullptr, "out");
    return k7_call(ctx, d_seg_offsets, n_segments, d_is_variant, d_quality, d_agreement, mean_error_rate, test_type, alpha, d_out);
}

int32_t pmrd_call(pmrd_ctx* ctx, const int64_t* h_seg_offsets, int64_t n_segments, const uint8_t* h_is_variant,
                  const double* h_quality, const double* h_agreement, double mean_error_rate, int32_t test_type,
                  double alpha, pmrd_call_table* h_out) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_segments >= 0, "n_segments");
    PMRD_ARG(ctx, test_type == PMRD_TEST_POISSON || test_type == PMRD_TEST_BINOMIAL, "Unknown test type");
    if (n_segments == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_seg_offsets && h_is_variant && h_out, "null buffer");
    const int64_t S = n_segments, R = h_seg_offsets[S];
    PMRD_ARG(ctx, h_seg_offsets[0] == 0 && R >= 0, "seg_offsets must start at 0");
    DevBuf so, iv, q, a, nv, nt, vf, ev, fc, pv, mq, mc, pa, sg;
    PMRD_TRY(up(ctx, so, h_seg_offsets, (size_t)(S + 1) * 8));
    PMRD_TRY(up(ctx, iv, h_is_variant, (size_t)R));
    if (h_quality) PMRD_TRY(up(ctx, q, h_quality, (size_t)R * 8));
    if (h_agreement) PMRD_TRY(up(ctx, a, h_agreement, (size_t)R * 8));
    PMRD_CUDA(ctx, nv.alloc(ctx->stream, S * 8)); PMRD_CUDA(ctx, nt.alloc(ctx->stream, S * 8));
    PMRD_CUDA(ctx, vf.alloc(ctx->stream, S * 8)); PMRD_CUDA(ctx, ev.alloc(ctx->stream, S * 8));
    PMRD_CUDA(ctx, fc.alloc(ctx->stream, S * 8)); PMRD_CUDA(ctx, pv.alloc(ctx->stream, S * 8));
    PMRD_CUDA(ctx, mq.alloc(ctx->stream, S * 8)); PMRD_CUDA(ctx, mc.alloc(ctx->stream, S * 8));
    PMRD_CUDA(ctx, pa.alloc(ctx->stream, S * 8)); PMRD_CUDA(ctx, sg.alloc(ctx->stream, S));
    pmrd_call_table d{nv.as<int64_t>(), nt.as<int64_t>(), vf.as<double>(), ev.as<double>(), fc.as<double>(),
                      pv.as<double>(), mq.as<double>(), mc.as<double>(), pa.as<double>(), sg.as<uint8_t>()};
    PMRD_TRY(k7_call(ctx, so.as<int64_t>(), S, iv.as<uint8_t>(), h_quality ? q.as<double>() : nullptr,
                     h_agreement ? a.as<double>() : nullptr, mean_error_rate, test_type, alpha, &d));
    PMRD_TRY(down(ctx, h_out->n_variants, d.n_variants, S * 8));
    PMRD_TRY(down(ctx, h_out->n_total, d.n_total, S * 8));
    PMRD_TRY(down(ctx, h_out->variant_fraction, d.variant_fraction, S * 8));
    PMRD_TRY(down(ctx, h_out->expected_variants, d.expected_variants, S * 8));
    PMRD_TRY(down(ctx, h_out->fold_change, d.fold_change, S * 8));
    PMRD_TRY(down(ctx, h_out->p_value, d.p_value, S * 8));
    PMRD_TRY(down(ctx, h_out->mean_quality, d.mean_quality, S * 8));
    PMRD_TRY(down(ctx, h_out->mean_consensus, d.mean_consensus, S * 8));
    PMRD_TRY(down(ctx, h_out->p_adjusted, d.p_adjusted, S * 8));
    PMRD_TRY(down(ctx, h_out->significant, d.significant, S));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- sort ------------------
int32_t pmrd_sort_pairs_dev(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, int64_t n, uint64_t* d_keys_tmp,
                            uint32_t* d_vals_tmp, int32_t* out_in_tmp) {
    ENTER(ctx);
    PMRD_ARG(ctx, out_in_tmp != nullptr, "out_in_tmp");
    PMRD_ARG(ctx, ((uintptr_t)d_keys & 15) == 0 && ((uintptr_t)d_keys_tmp & 15) == 0, "key buffers must be 16-byte aligned");
    uint64_t varying = 0;
    PMRD_TRY(radix_varying_bits(ctx, d_keys, n, &varying));
    int in_b = 0;
    PMRD_TRY(radix_sort_pairs(ctx, d_keys, d_vals, d_keys_tmp, d_vals_tmp, n, varying, &in_b));
    *out_in_tmp = in_b;
    return PMRD_OK;
}

int32_t pmrd_sort_pairs(pmrd_ctx* ctx, const uint64_t* h_keys, const uint32_t* h_vals, int64_t n, uint64_t* h_keys_out,
                        uint32_t* h_vals_out) {
    ENTER(ctx);
    PMRD_ARG(ctx, n >= 0, "n");
    if (n == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_keys && h_keys_out, "null buffer");
    DevBuf ka, kb, va, vb;
    PMRD_TRY(up(ctx, ka, h_keys, (size_t)n * 8));
    PMRD_CUDA(ctx, kb.alloc(ctx->stream, (size_t)n * 8));
    if (h_vals) {
        PMRD_TRY(up(ctx, va, h_vals, (size_t)n * 4));
        PMRD_CUDA(ctx, vb.alloc(ctx->stream, (size_t)n * 4));
    }
    uint64_t varying = 0;
    PMRD_TRY(radix_varying_bits(ctx, ka.as<uint64_t>(), n, &varying));
    int in_b = 0;
    PMRD_TRY(radix_sort_pairs(ctx, ka.as<uint64_t>(), h_vals ? va.as<uint32_t>() : nullptr, kb.as<uint64_t>(),
                              h_vals ? vb.as<uint32_t>() : nullptr, n, varying, &in_b));
    PMRD_TRY(down(ctx, h_keys_out, in_b ? kb.p : ka.p, (size_t)n * 8));
    if (h_vals && h_vals_out) PMRD_TRY(down(ctx, h_vals_out, in_b ? vb.p : va.p, (size_t)n * 4));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- collapse --------------
int32_t pmrd_sort_packed_cells(pmrd_ctx* ctx, const uint64_t* h_rec, int64_t n, const int64_t* h_cell_tile_start, int64_t n_cells,
                               int32_t key_bits, uint64_t* h_out) {
    ENTER(ctx);
    PMRD_ARG(ctx, n >= 0 && n_cells >= 0 && (n == 0 || (h_rec && h_out && h_cell_tile_start)), "null buffer");
    if (n == 0 || n_cells == 0) return PMRD_OK;
    PMRD_ARG(ctx, n % RS_TILE == 0 && h_cell_tile_start[0] == 0 && h_cell_tile_start[n_cells] * RS_TILE == n, "cells must be whole tiles covering [0, n)");
    DevBuf a, b, ts;
    PMRD_TRY(up(ctx, a, h_rec, (size_t)n * 8));
    PMRD_CUDA(ctx, b.alloc(ctx->stream, (size_t)n * 8));
    PMRD_TRY(up(ctx, ts, h_cell_tile_start, (size_t)(n_cells + 1) * 8));
    int in_b = 0;
    PMRD_TRY(radix_sort_pairs_segmented(ctx, a.as<uint64_t>(), nullptr, b.as<uint64_t>(), nullptr, n, ts.as<int64_t>(), n_cells, key_bits, &in_b));
    PMRD_TRY(down(ctx, h_out, in_b ? b.p : a.p, (size_t)n * 8));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

int32_t pmrd_collapse_reads_dev(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp,
                                uint32_t* d_payload_tmp, int64_t n_reads, const pmrd_umi_params* params,
                                pmrd_family_table* d_fam, int64_t capacity_f, int64_t* n_families,
                                pmrd_group_table* d_grp, int64_t capacity_g, int64_t* n_groups) {
    ENTER(ctx);
    PMRD_ARG(ctx, d_fam && n_families, "family table");
    return k5_collapse(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n_reads, params, d_fam, capacity_f, n_families,
                       d_grp, capacity_g, n_groups);
}

int32_t pmrd_owner_partition_dev(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp, uint32_t* d_payload_tmp,
                                 int64_t n_reads, int32_t log2_world, int64_t* h_counts, int32_t* in_tmp) {
    ENTER(ctx);
    return k18_owner_partition(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n_reads, log2_world, h_counts, in_tmp);
}

int32_t pmrd_owner_rekey_dev(pmrd_ctx* ctx, uint64_t* d_keys, int64_t n_reads, int32_t log2_world) {
    ENTER(ctx);
    return k18_owner_rekey(ctx, d_keys, n_reads, log2_world);
}

int32_t pmrd_owner_tiles_dev(pmrd_ctx* ctx, const uint64_t* d_keys, int64_t n_reads, int32_t log2_world, uint32_t* d_tile_off, int64_t* h_counts) {
    ENTER(ctx);
    return k18_owner_tiles(ctx, d_keys, n_reads, log2_world, d_tile_off, h_counts);
}

int32_t pmrd_owner_scatter_peer_dev(pmrd_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_payload, int64_t n_reads, int32_t log2_world,
                                    const uint32_t* d_tile_off, uint64_t* const* peer_keys, uint32_t* const* peer_payload, const int64_t* h_base) {
    ENTER(ctx);
    return k18_owner_scatter_peer(ctx, d_keys, d_payload, n_reads, log2_world, d_tile_off, peer_keys, peer_payload, h_base);
}

int32_t pmrd_peer_alloc(pmrd_ctx* ctx, int64_t bytes, void** d_ptr, uint8_t* handle64) {
    ENTER(ctx);
    return k18_peer_alloc(ctx, bytes, d_ptr, handle64);
}

int32_t pmrd_peer_open(pmrd_ctx* ctx, const uint8_t* handle64, void** d_ptr) {
    ENTER(ctx);
    return k18_peer_open(ctx, handle64, d_ptr);
}

int32_t pmrd_peer_close(pmrd_ctx* ctx, void* d_ptr) {
    ENTER(ctx);
    return k18_peer_close(ctx, d_ptr);
}

int32_t pmrd_peer_free(pmrd_ctx* ctx, void* d_ptr) {
    ENTER(ctx);
    return k18_peer_free(ctx, d_ptr);
}

int32_t pmrd_collapse_reads(pmrd_ctx* ctx, const uint64_t* h_keys, const uint32_t* h_payload, int64_t n_reads,
                            const pmrd_umi_params* params, pmrd_family_table* h_fam, int64_t capacity_f,
                            int64_t* n_families, pmrd_group_table* h_grp, int64_t capacity_g, int64_t* n_groups) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_reads >= 0, "n_reads");
    PMRD_ARG(ctx, h_fam && n_families && params, "null argument");
    *n_families = 0;
    if (n_groups) *n_groups = 0;
    if (n_reads == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_keys && h_payload, "null buffer");
    const size_t n = (size_t)n_reads;
    DevBuf ka, kb, pa, pb, fk, fs, fa, fc, fq, fn, ff, gg, gf, gc, ga, gt;
    PMRD_TRY(up(ctx, ka, h_keys, n * 8));
    PMRD_TRY(up(ctx, pa, h_payload, n * 4));
    PMRD_CUDA(ctx, kb.alloc(ctx->stream, n * 8));
    PMRD_CUDA(ctx, pb.alloc(ctx->stream, n * 4));
    const bool want_fam = h_fam->key != nullptr;          // NULL columns: only the group table is wanted (no family row is even computed
                                                           // when the group-major path applies)
    PMRD_ARG(ctx, want_fam || (h_grp && n_groups), "neither a family table nor a group table was asked for");
    pmrd_family_table dfam{};
    if (want_fam) {
        PMRD_CUDA(ctx, fk.alloc(ctx->stream, n * 8));
        PMRD_CUDA(ctx, fs.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, fa.alloc(ctx->stream, n * 4));
        PMRD_CUDA(ctx, fc.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, fq.alloc(ctx->stream, n * 4));
        PMRD_CUDA(ctx, fn.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, ff.alloc(ctx->stream, n * 4));
        dfam = pmrd_family_table{fk.as<uint64_t>(), fs.as<uint32_t>(), fa.as<uint32_t>(), fc.as<uint32_t>(),
                                 fq.as<uint32_t>(), fn.as<uint32_t>(), ff.as<uint32_t>()};
    }
    pmrd_group_table dgrp{};
    const bool want_groups = h_grp && n_groups;
    if (want_groups) {
        PMRD_CUDA(ctx, gg.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, gf.alloc(ctx->stream, n * 4));
        PMRD_CUDA(ctx, gc.alloc(ctx->stream, n * 16)); PMRD_CUDA(ctx, ga.alloc(ctx->stream, n * 8));
        PMRD_CUDA(ctx, gt.alloc(ctx->stream, n * 8));
        dgrp = pmrd_group_table{gg.as<uint32_t>(), gf.as<uint32_t>(), gc.as<uint32_t>(), ga.as<uint64_t>(), gt.as<uint64_t>()};
    }
    int64_t nf = 0, ng = 0;
    PMRD_TRY(k5_collapse(ctx, ka.as<uint64_t>(), pa.as<uint32_t>(), kb.as<uint64_t>(), pb.as<uint32_t>(), n_reads, params,
                         &dfam, n_reads, &nf, want_groups ? &dgrp : nullptr, n_reads, want_groups ? &ng : nullptr));
    *n_families = nf;
    if (n_groups) *n_groups = ng;
    if ((want_fam && nf > capacity_f) || (want_groups && ng > capacity_g)) {
        snprintf(ctx->err, sizeof(ctx->err), "output capacity too small: %lld families, %lld groups", (long long)nf, (long long)ng);
        return PMRD_ERR_CAPACITY;
    }
    if (want_fam) PMRD_ARG(ctx, h_fam->size && h_fam->n_alt && h_fam->consensus && h_fam->qsum && h_fam->n_best && h_fam->flags, "family table columns");
    if (!want_fam) nf = 0;                                 // (nothing to download; *n_families already holds the count)
    PMRD_TRY(down(ctx, h_fam->key, dfam.key, (size_t)nf * 8));
    PMRD_TRY(down(ctx, h_fam->size, dfam.size, (size_t)nf * 4));
    PMRD_TRY(down(ctx, h_fam->n_alt, dfam.n_alt, (size_t)nf * 4));
    PMRD_TRY(down(ctx, h_fam->consensus, dfam.consensus, (size_t)nf * 4));
    PMRD_TRY(down(ctx, h_fam->qsum, dfam.qsum, (size_t)nf * 4));
    PMRD_TRY(down(ctx, h_fam->n_best, dfam.n_best, (size_t)nf * 4));
    PMRD_TRY(down(ctx, h_fam->flags, dfam.flags, (size_t)nf * 4));
    if (want_groups) {
        PMRD_TRY(down(ctx, h_grp->group, dgrp.group, (size_t)ng * 4));
        PMRD_TRY(down(ctx, h_grp->family_count, dgrp.family_count, (size_t)ng * 4));
        PMRD_TRY(down(ctx, h_grp->consensus_count, dgrp.consensus_count, (size_t)ng * 16));
        PMRD_TRY(down(ctx, h_grp->alt_count, dgrp.alt_count, (size_t)ng * 8));
        PMRD_TRY(down(ctx, h_grp->total_count, dgrp.total_count, (size_t)ng * 8));
    }
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

int32_t pmrd_umi_hamming(pmrd_ctx* ctx, const uint32_t* h_a, const uint32_t* h_b, int64_t n, int32_t umi_len, int32_t* h_d) {
    ENTER(ctx);
    PMRD_ARG(ctx, n >= 0 && umi_len >= 1 && umi_len <= 16, "n / umi_len");
    if (n == 0) return PMRD_OK;
    DevBuf a, b, d;
    PMRD_TRY(up(ctx, a, h_a, (size_t)n * 4));
    PMRD_TRY(up(ctx, b, h_b, (size_t)n * 4));
    PMRD_CUDA(ctx, d.alloc(ctx->stream, (size_t)n * 4));
    PMRD_TRY(k6_hamming(ctx, a.as<uint32_t>(), b.as<uint32_t>(), n, d.as<int32_t>()));
    PMRD_TRY(down(ctx, h_d, d.p, (size_t)n * 4));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}


// ---------------------------------------------------------------- K14 -------------------
int32_t pmrd_fisher_exact(pmrd_ctx* ctx, const int64_t* h_table, int64_t m, double* h_odds_ratio, double* h_p) {
    ENTER(ctx);
    PMRD_ARG(ctx, m >= 0, "fisher: m");
    if (m == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_table && h_odds_ratio && h_p, "fisher: null buffer");
    for (int64_t i = 0; i < 4 * m; ++i) PMRD_ARG(ctx, h_table[i] >= 0, "All values in `table` must be nonnegative.");
    DevBuf t, o, p;
    PMRD_TRY(up(ctx, t, h_table, (size_t)m * 32));
    PMRD_CUDA(ctx, o.alloc(ctx->stream, (size_t)m * 8));
    PMRD_CUDA(ctx, p.alloc(ctx->stream, (size_t)m * 8));
    PMRD_TRY(k14_fisher(ctx, t.as<int64_t>(), m, o.as<double>(), p.as<double>()));
    PMRD_TRY(down(ctx, h_odds_ratio, o.p, (size_t)m * 8));
    PMRD_TRY(down(ctx, h_p, p.p, (size_t)m * 8));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- K13 -------------------
int32_t pmrd_fastq_scan(pmrd_ctx* ctx, const uint8_t* h_bytes, int64_t n_bytes, int64_t capacity, pmrd_fastq_records* h_out,
                        int64_t* n_records) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_bytes >= 0 && n_records && (n_bytes == 0 || h_bytes), "fastq: buffer");
    *n_records = 0;
    if (n_bytes == 0) return PMRD_OK;
    PMRD_ARG(ctx, capacity >= 0 && h_out, "fastq: outputs");
    const size_t c = (size_t)(capacity > 0 ? capacity : 1);
    DevBuf buf, st, key, ho, hl, uo, ul, so, sl, qo, ql, qs;
    PMRD_TRY(up(ctx, buf, h_bytes, (size_t)n_bytes));
    PMRD_CUDA(ctx, st.alloc(ctx->stream, c * 4)); PMRD_CUDA(ctx, key.alloc(ctx->stream, c * 8));
    PMRD_CUDA(ctx, ho.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, hl.alloc(ctx->stream, c * 4));
    PMRD_CUDA(ctx, uo.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, ul.alloc(ctx->stream, c * 4));
    PMRD_CUDA(ctx, so.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, sl.alloc(ctx->stream, c * 4));
    PMRD_CUDA(ctx, qo.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, ql.alloc(ctx->stream, c * 4));
    PMRD_CUDA(ctx, qs.alloc(ctx->stream, c * 8));
    pmrd_fastq_records d{st.as<int32_t>(), key.as<uint64_t>(), ho.as<int64_t>(), hl.as<int32_t>(), uo.as<int64_t>(), ul.as<int32_t>(),
                         so.as<int64_t>(), sl.as<int32_t>(), qo.as<int64_t>(), ql.as<int32_t>(), qs.as<int64_t>()};
    PMRD_TRY(k13_fastq_scan(ctx, buf.as<uint8_t>(), n_bytes, capacity, &d, n_records));
    const size_t r = (size_t)*n_records;
    PMRD_TRY(down(ctx, h_out->status, d.status, r * 4)); PMRD_TRY(down(ctx, h_out->umi_key, d.umi_key, r * 8));
    PMRD_TRY(down(ctx, h_out->hdr_off, d.hdr_off, r * 8)); PMRD_TRY(down(ctx, h_out->hdr_len, d.hdr_len, r * 4));
    PMRD_TRY(down(ctx, h_out->umi_off, d.umi_off, r * 8)); PMRD_TRY(down(ctx, h_out->umi_len, d.umi_len, r * 4));
    PMRD_TRY(down(ctx, h_out->seq_off, d.seq_off, r * 8)); PMRD_TRY(down(ctx, h_out->seq_len, d.seq_len, r * 4));
    PMRD_TRY(down(ctx, h_out->qual_off, d.qual_off, r * 8)); PMRD_TRY(down(ctx, h_out->qual_len, d.qual_len, r * 4));
    PMRD_TRY(down(ctx, h_out->qual_sum, d.qual_sum, r * 8));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

int32_t pmrd_fastq_families_of(pmrd_ctx* ctx, const uint64_t* h_umi_key, const int64_t* h_qual_sum, const int32_t* h_qual_len,
                               int64_t n_reads, pmrd_fastq_families* h_out, int64_t capacity, int64_t* n_families) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_families != nullptr && n_reads >= 0 && n_reads < (1ll << 31), "fastq families: n_reads");
    *n_families = 0;
    if (n_reads == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_umi_key && h_qual_sum && h_qual_len && h_out && capacity >= 0, "fastq families: buffers");
    const size_t n = (size_t)n_reads, c = (size_t)(capacity > 0 ? capacity : 1);
    DevBuf keys, qs, ql, fk, ff, fb, fs, fq, fn;
    PMRD_TRY(up(ctx, keys, h_umi_key, n * 8));
    PMRD_TRY(up(ctx, qs, h_qual_sum, n * 8));
    PMRD_TRY(up(ctx, ql, h_qual_len, n * 4));
    PMRD_CUDA(ctx, fk.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, ff.alloc(ctx->stream, c * 4)); PMRD_CUDA(ctx, fb.alloc(ctx->stream, c * 4));
    PMRD_CUDA(ctx, fs.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, fq.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, fn.alloc(ctx->stream, c * 8));
    pmrd_fastq_families d{fk.as<uint64_t>(), ff.as<uint32_t>(), fb.as<uint32_t>(), fs.as<int64_t>(), fq.as<int64_t>(), fn.as<int64_t>()};
    PMRD_TRY(k13_fastq_families(ctx, keys.as<uint64_t>(), n_reads, qs.as<int64_t>(), ql.as<int32_t>(), &d, capacity, n_families));
    const size_t F = (size_t)*n_families;
    PMRD_TRY(down(ctx, h_out->umi_key, d.umi_key, F * 8)); PMRD_TRY(down(ctx, h_out->first_read, d.first_read, F * 4));
    PMRD_TRY(down(ctx, h_out->best_read, d.best_read, F * 4)); PMRD_TRY(down(ctx, h_out->family_size, d.family_size, F * 8));
    PMRD_TRY(down(ctx, h_out->qual_sum, d.qual_sum, F * 8)); PMRD_TRY(down(ctx, h_out->n_bases, d.n_bases, F * 8));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- K12 -------------------
int32_t pmrd_bootstrap_metrics(pmrd_ctx* ctx, const uint8_t* h_y_true, const double* h_y_score, int64_t n,
                               const int64_t* h_indices, int64_t n_boot, uint64_t stream_id, double* h_point,
                               double* h_auc, double* h_ap) {
    ENTER(ctx);
    PMRD_ARG(ctx, n > 0 && h_y_true && h_y_score, "bootstrap metrics: empty input");
    PMRD_ARG(ctx, n_boot >= 0 && (n_boot == 0 || (h_auc && h_ap)), "bootstrap metrics: outputs");
    const int want_point = h_point != nullptr;
    const int64_t n_res = n_boot + want_point;
    if (n_res == 0) return PMRD_OK;
    DevBuf y, sc, ix, auc, ap;
    PMRD_TRY(up(ctx, y, h_y_true, (size_t)n));
    PMRD_TRY(up(ctx, sc, h_y_score, (size_t)n * 8));
    if (h_indices && n_boot > 0) PMRD_TRY(up(ctx, ix, h_indices, (size_t)n_boot * (size_t)n * 8));
    PMRD_CUDA(ctx, auc.alloc(ctx->stream, (size_t)n_res * 8));
    PMRD_CUDA(ctx, ap.alloc(ctx->stream, (size_t)n_res * 8));
    PMRD_TRY(k12_bootstrap_metrics(ctx, y.as<uint8_t>(), sc.as<double>(), n,
                                   (h_indices && n_boot > 0) ? (const int64_t*)ix.as<int64_t>() : (const int64_t*)nullptr, n_boot,
                                   want_point, stream_id, auc.as<double>(), ap.as<double>()));
    PMRD_TRY(down(ctx, h_auc, auc.p, (size_t)n_boot * 8));
    PMRD_TRY(down(ctx, h_ap, ap.p, (size_t)n_boot * 8));
    if (want_point) {
        PMRD_TRY(down(ctx, h_point, auc.as<double>() + n_boot, 8));
        PMRD_TRY(down(ctx, h_point + 1, ap.as<double>() + n_boot, 8));
    }
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- K1 --------------------
int32_t pmrd_simulate_cells(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_depth,
                            int64_t n_cells, int32_t max_family_size, pmrd_cell_table* h_out) {
    return pmrd_simulate_cells_stratified(ctx, h_cell_id, h_af, h_depth, nullptr, n_cells, max_family_size, h_out);
}

int32_t pmrd_simulate_cells_stratified(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_depth,
                                       const double* h_bg_multiplier, int64_t n_cells, int32_t max_family_size,
                                       pmrd_cell_table* h_out) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_cells >= 0, "n_cells");
    if (n_cells == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_cell_id && h_af && h_depth && h_out, "null buffer");
    const size_t c = (size_t)n_cells;
    DevBuf id, af, dp, bg, mfs, nt, nfp, mq, mult;
    PMRD_TRY(up(ctx, id, h_cell_id, c * 8));
    PMRD_TRY(up(ctx, af, h_af, c * 8));
    PMRD_TRY(up(ctx, dp, h_depth, c * 8));
    if (h_bg_multiplier) PMRD_TRY(up(ctx, mult, h_bg_multiplier, c * 8));
    PMRD_CUDA(ctx, bg.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, mfs.alloc(ctx->stream, c * 8));
    PMRD_CUDA(ctx, nt.alloc(ctx->stream, c * 8)); PMRD_CUDA(ctx, nfp.alloc(ctx->stream, c * 8));
    PMRD_CUDA(ctx, mq.alloc(ctx->stream, c * 8));
    pmrd_cell_table d{bg.as<double>(), mfs.as<int64_t>(), nt.as<int64_t>(), nfp.as<int64_t>(), mq.as<double>()};
    PMRD_TRY(k1_simulate_cells(ctx, id.as<int64_t>(), af.as<double>(), dp.as<int64_t>(),
                               h_bg_multiplier ? (const double*)mult.as<double>() : (const double*)nullptr, n_cells, max_family_size, &d));
    PMRD_TRY(down(ctx, h_out->background_rate, d.background_rate, c * 8));
    PMRD_TRY(down(ctx, h_out->mean_family_size, d.mean_family_size, c * 8));
    PMRD_TRY(down(ctx, h_out->n_true_variants, d.n_true_variants, c * 8));
    PMRD_TRY(down(ctx, h_out->n_false_positives, d.n_false_positives, c * 8));
    PMRD_TRY(down(ctx, h_out->mean_quality, d.mean_quality, c * 8));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- K2 (G-D) --------------
struct GdCellHost {
    uint64_t cell_id;
    double af, bg, lambda;
};

int32_t pmrd_collapse_families(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                               const double* h_background_rate, const double* h_mean_family_size, int64_t n_cells,
                               const pmrd_umi_params* params, int32_t variant, int64_t* h_row_offsets,
                               pmrd_gd_family_table* h_out, int64_t capacity) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_cells >= 0 && params && h_row_offsets, "arguments");
    PMRD_ARG(ctx, variant == 1 || variant == 2, "variant must be 1 (G-D v1) or 2 (G-D v2)");
    h_row_offsets[0] = 0;
    if (n_cells == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_cell_id && h_af && h_n_families && h_background_rate && h_mean_family_size, "null buffer");
    const int TILE = 1024;
    int64_t n_tiles = 0;
    for (int64_t i = 0; i < n_cells; ++i) {
        PMRD_ARG(ctx, h_n_families[i] >= 0 && h_n_families[i] < (1ll << 31), "n_families out of range");
        n_tiles += (h_n_families[i] + TILE - 1) / TILE;
    }
    PMRD_ARG(ctx, n_tiles < (1ll << 31), "too many tiles; chunk the cells");
    GdCellHost* hc = (GdCellHost*)malloc(sizeof(GdCellHost) * (size_t)n_cells);
    int32_t* tcell = (int32_t*)malloc(4 * (size_t)(n_tiles + 1) * 2);
    int64_t* first_tile = (int64_t*)malloc(8 * (size_t)(n_cells + 1));
    uint64_t* tbase = (uint64_t*)malloc(8 * (size_t)(n_tiles + 1));
    if (!hc || !tcell || !first_tile || !tbase) {
        free(hc); free(tcell); free(first_tile); free(tbase);
        PMRD_ARG(ctx, false, "out of host memory");
    }
    int32_t* tfam0 = tcell + (n_tiles + 1);
    int64_t t = 0;
    for (int64_t i = 0; i < n_cells; ++i) {
        hc[i].cell_id = (uint64_t)h_cell_id[i];
        hc[i].af = h_af[i];
        hc[i].bg = h_background_rate[i];
        // collapse.py:80 mean_size = max(3, mean_family_size)   ("collapse 2.py":56: poisson(5))
        hc[i].lambda = variant == 1 ? 5.0 : (h_mean_family_size[i] > 3.0 ? h_mean_family_size[i] : 3.0);
        first_tile[i] = t;
        for (int64_t f0 = 0; f0 < h_n_families[i]; f0 += TILE) {
            tcell[t] = (int32_t)i;
            tfam0[t] = (int32_t)f0;
            ++t;
        }
    }
    first_tile[n_cells] = t;
    int32_t st = PMRD_OK;
    {
        DevBuf dc, dtc, dtf, dnf, dbase, oc, ofid, ofs, oq, oa, ofl;
        auto run = [&]() -> int32_t {
            PMRD_TRY(up(ctx, dc, hc, sizeof(GdCellHost) * (size_t)n_cells));
            PMRD_TRY(up(ctx, dtc, tcell, 4 * (size_t)n_tiles));
            PMRD_TRY(up(ctx, dtf, tfam0, 4 * (size_t)n_tiles));
            PMRD_TRY(up(ctx, dnf, h_n_families, 8 * (size_t)n_cells));
            PMRD_CUDA(ctx, dbase.alloc(ctx->stream, 8 * (size_t)(n_tiles + 1)));
            PMRD_CUDA(ctx, cudaMemsetAsync(dbase.p, 0, 8 * (size_t)(n_tiles + 1), ctx->stream));
            PMRD_TRY(k2_gd_families(ctx, dtc.as<int32_t>(), dtf.as<int32_t>(), n_tiles, dc.p, dnf.as<int64_t>(), params, variant,
                                    dbase.as<uint64_t>(), nullptr));
            PMRD_CUDA(ctx, cudaMemcpyAsync(tbase, dbase.p, 8 * (size_t)(n_tiles + 1), cudaMemcpyDeviceToHost, ctx->stream));
            PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (int64_t i = 0; i <= n_cells; ++i) h_row_offsets[i] = (int64_t)tbase[first_tile[i]];
            const int64_t R = (int64_t)tbase[n_tiles];
            if (!h_out) return PMRD_OK;
            if (R > capacity) {
                snprintf(ctx->err, sizeof(ctx->err), "family capacity %lld < %lld rows", (long long)capacity, (long long)R);
                return PMRD_ERR_CAPACITY;
            }
            if (R == 0) return PMRD_OK;
            const size_t r = (size_t)R;
            PMRD_CUDA(ctx, oc.alloc(ctx->stream, r * 4)); PMRD_CUDA(ctx, ofid.alloc(ctx->stream, r * 4));
            PMRD_CUDA(ctx, ofs.alloc(ctx->stream, r * 4)); PMRD_CUDA(ctx, oq.alloc(ctx->stream, r * 8));
            PMRD_CUDA(ctx, oa.alloc(ctx->stream, r * 8)); PMRD_CUDA(ctx, ofl.alloc(ctx->stream, r));
            pmrd_gd_family_table d{oc.as<int32_t>(), ofid.as<int32_t>(), ofs.as<int32_t>(), oq.as<double>(), oa.as<double>(), ofl.as<uint8_t>()};
            PMRD_TRY(k2_gd_families(ctx, dtc.as<int32_t>(), dtf.as<int32_t>(), n_tiles, dc.p, dnf.as<int64_t>(), params, variant,
                                    dbase.as<uint64_t>(), &d));
            PMRD_TRY(down(ctx, h_out->cell, d.cell, r * 4));
            PMRD_TRY(down(ctx, h_out->family_id, d.family_id, r * 4));
            PMRD_TRY(down(ctx, h_out->family_size, d.family_size, r * 4));
            PMRD_TRY(down(ctx, h_out->quality_score, d.quality_score, r * 8));
            PMRD_TRY(down(ctx, h_out->consensus_agreement, d.consensus_agreement, r * 8));
            PMRD_TRY(down(ctx, h_out->flags, d.flags, r));
            PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return PMRD_OK;
        };
        st = run();
        cudaStreamSynchronize(ctx->stream);
    }
    free(hc); free(tcell); free(first_tile); free(tbase);
    return st;
}

// ---------------------------------------------------------------- K10 -------------------
int32_t pmrd_error_model(pmrd_ctx* ctx, int64_t n_neg, int64_t n_var, int32_t n_boot, uint64_t stream_id,
                         double* h_rate64, double* h_ci_lo64, double* h_ci_hi64) {
    ENTER(ctx);
    PMRD_ARG(ctx, h_rate64 && h_ci_lo64 && h_ci_hi64, "null buffer");
    DevBuf r, lo, hi;
    PMRD_CUDA(ctx, r.alloc(ctx->stream, 512)); PMRD_CUDA(ctx, lo.alloc(ctx->stream, 512)); PMRD_CUDA(ctx, hi.alloc(ctx->stream, 512));
    PMRD_TRY(k10_error_model(ctx, n_neg, n_var, n_boot, stream_id, r.as<double>(), lo.as<double>(), hi.as<double>()));
    PMRD_TRY(down(ctx, h_rate64, r.p, 512));
    PMRD_TRY(down(ctx, h_ci_lo64, lo.p, 512));
    PMRD_TRY(down(ctx, h_ci_hi64, hi.p, 512));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- read-level simulate ---
int32_t pmrd_simulate_reads_count(pmrd_ctx* ctx, const int64_t* h_cell_id, const int64_t* h_n_families, int64_t n_cells,
                                  const pmrd_read_sim_params* sp, int64_t* h_read_offsets) {
    ENTER(ctx);
    PMRD_ARG(ctx, h_read_offsets != nullptr, "h_read_offsets");
    h_read_offsets[0] = 0;
    if (n_cells <= 0) return PMRD_OK;
    ReadPlan plan;
    int32_t st = reads_plan_build(ctx, h_cell_id, nullptr, h_n_families, n_cells, 0, sp, 0, &plan, h_read_offsets);
    cudaStreamSynchronize(ctx->stream);
    return st;
}

int32_t pmrd_simulate_reads_dev(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                                int64_t n_cells, uint32_t group_base, const pmrd_read_sim_params* sp, uint64_t* d_keys,
                                uint32_t* d_payload, int64_t capacity, int64_t* n_reads) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_reads != nullptr, "n_reads");
    *n_reads = 0;
    if (n_cells <= 0) return PMRD_OK;
    ReadPlan plan;
    const int cell_major = sp && sp->shuffle == 2 && sp->hop_rate == 0.0;   // the plan's own order (see pmrd.h)
    PMRD_TRY(reads_plan_build(ctx, h_cell_id, h_af, h_n_families, n_cells, group_base, sp, cell_major, &plan, nullptr));
    *n_reads = plan.n_reads;
    if (plan.n_reads > capacity) {
        snprintf(ctx->err, sizeof(ctx->err), "read capacity %lld < %lld reads", (long long)capacity, (long long)plan.n_reads);
        return PMRD_ERR_CAPACITY;
    }
    if (cell_major) PMRD_TRY(reads_plan_generate_cell_major(ctx, &plan, d_keys, d_payload));
    else PMRD_TRY(reads_plan_generate(ctx, &plan, d_keys, d_payload));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the plan's scratch dies with this scope
    return PMRD_OK;
}

int32_t pmrd_simulate_reads(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                            int64_t n_cells, uint32_t group_base, const pmrd_read_sim_params* sp, uint64_t* h_keys,
                            uint32_t* h_payload, int64_t capacity, int64_t* n_reads) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_reads != nullptr, "n_reads");
    *n_reads = 0;
    if (n_cells <= 0) return PMRD_OK;
    ReadPlan plan;
    const int cell_major = sp && sp->shuffle == 2 && sp->hop_rate == 0.0;   // the plan's own order (see pmrd.h)
    PMRD_TRY(reads_plan_build(ctx, h_cell_id, h_af, h_n_families, n_cells, group_base, sp, cell_major, &plan, nullptr));
    *n_reads = plan.n_reads;
    if (plan.n_reads > capacity) {
        snprintf(ctx->err, sizeof(ctx->err), "read capacity %lld < %lld reads", (long long)capacity, (long long)plan.n_reads);
        return PMRD_ERR_CAPACITY;
    }
    if (plan.n_reads == 0) return PMRD_OK;
    DevBuf k, p;
    PMRD_CUDA(ctx, k.alloc(ctx->stream, (size_t)plan.n_reads * 8));
    PMRD_CUDA(ctx, p.alloc(ctx->stream, (size_t)plan.n_reads * 4));
    if (cell_major) PMRD_TRY(reads_plan_generate_cell_major(ctx, &plan, k.as<uint64_t>(), p.as<uint32_t>()));
    else PMRD_TRY(reads_plan_generate(ctx, &plan, k.as<uint64_t>(), p.as<uint32_t>()));
    PMRD_TRY(down(ctx, h_keys, k.p, (size_t)plan.n_reads * 8));
    PMRD_TRY(down(ctx, h_payload, p.p, (size_t)plan.n_reads * 4));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- pipeline / plan -------
int32_t pmrd_plan_create(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                         int64_t n_cells, const pmrd_read_sim_params* sp, const pmrd_umi_params* up_, double neg_af_max,
                         int32_t test_type, int32_t alternative, double alpha, pmrd_plan** out) {
    ENTER(ctx);
    return plan_create(ctx, h_cell_id, h_af, h_n_families, n_cells, sp, up_, neg_af_max, test_type, alternative, alpha, nullptr, out);
}
int32_t pmrd_plan_create_ex(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                            int64_t n_cells, const pmrd_read_sim_params* sp, const pmrd_umi_params* up_, double neg_af_max,
                            int32_t test_type, int32_t alternative, double alpha, const pmrd_cell_knobs* knobs, pmrd_plan** out) {
    ENTER(ctx);
    return plan_create(ctx, h_cell_id, h_af, h_n_families, n_cells, sp, up_, neg_af_max, test_type, alternative, alpha, knobs, out);
}
int32_t pmrd_plan_group_rates(pmrd_ctx* ctx, pmrd_plan* plan, double* h_rates, int64_t capacity) {
    ENTER(ctx);
    PMRD_ARG(ctx, plan != nullptr && h_rates != nullptr, "plan / buffer");
    return plan_group_rates(ctx, plan, h_rates, capacity);
}
int32_t pmrd_plan_run(pmrd_ctx* ctx, pmrd_plan* plan) {
    ENTER(ctx);
    PMRD_ARG(ctx, plan != nullptr, "plan");
    return plan_run(ctx, plan);
}
int32_t pmrd_plan_run_simulate(pmrd_ctx* ctx, pmrd_plan* plan, int64_t chunk) {
    ENTER(ctx);
    PMRD_ARG(ctx, plan != nullptr, "plan");
    return plan_run_simulate(ctx, plan, chunk);
}
int32_t pmrd_plan_run_collapse(pmrd_ctx* ctx, pmrd_plan* plan, int64_t chunk) {
    ENTER(ctx);
    PMRD_ARG(ctx, plan != nullptr, "plan");
    return plan_run_collapse(ctx, plan, chunk);
}
int64_t pmrd_plan_n_chunks(const pmrd_plan* plan) { return plan_n_chunks(plan); }
int64_t pmrd_plan_chunk_reads(const pmrd_plan* plan, int64_t chunk) { return plan_chunk_reads(plan, chunk); }
int32_t pmrd_plan_info(const pmrd_plan* plan, int64_t* out8) { return plan_info(plan, out8); }
int32_t pmrd_plan_run_call(pmrd_ctx* ctx, pmrd_plan* plan) {
    ENTER(ctx);
    PMRD_ARG(ctx, plan != nullptr, "plan");
    return plan_run_call(ctx, plan);
}
int32_t pmrd_plan_results(pmrd_ctx* ctx, pmrd_plan* plan, pmrd_pipeline_result* h_out, double* h_mean_error_rate, int64_t* h_totals3) {
    ENTER(ctx);
    PMRD_ARG(ctx, plan != nullptr, "plan");
    return plan_results(ctx, plan, h_out, h_mean_error_rate, h_totals3);
}
int64_t pmrd_plan_n_reads(const pmrd_plan* plan) { return plan_n_reads(plan); }
int64_t pmrd_plan_n_families(const pmrd_plan* plan) { return plan_n_families(plan); }
void pmrd_plan_destroy(pmrd_plan* plan) {
    if (plan) {
        cudaDeviceSynchronize();
        plan_destroy(plan);
    }
}

int32_t pmrd_pipeline_cells(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                            int64_t n_cells, const pmrd_read_sim_params* sp, const pmrd_umi_params* up_, double neg_af_max,
                            int32_t test_type, int32_t alternative, double alpha, pmrd_pipeline_result* h_out,
                            double* h_mean_error_rate) {
    ENTER(ctx);
    PMRD_ARG(ctx, h_out != nullptr, "out");
    pmrd_plan* plan = nullptr;
    PMRD_TRY(plan_create(ctx, h_cell_id, h_af, h_n_families, n_cells, sp, up_, neg_af_max, test_type, alternative, alpha, nullptr, &plan));
    int32_t st = plan_run(ctx, plan);
    if (st == PMRD_OK) st = plan_results(ctx, plan, h_out, h_mean_error_rate, nullptr);
    cudaStreamSynchronize(ctx->stream);
    plan_destroy(plan);
    return st;
}

// call stage over per-cell counts gathered from several GPUs: the plan's own kernels
int32_t pmrd_call_counts(pmrd_ctx* ctx, const int64_t* h_n_total, const int64_t* h_n_variants, const uint8_t* h_is_negative,
                         int64_t n_cells, uint64_t stream_id, int32_t test_type, int32_t alternative, double alpha,
                         double* h_p_value, double* h_p_adjusted, uint8_t* h_significant, double* h_mean_error_rate) {
    ENTER(ctx);
    PMRD_ARG(ctx, n_cells > 0 && h_n_total && h_n_variants && h_is_negative, "arguments");
    PMRD_ARG(ctx, test_type == PMRD_TEST_POISSON || test_type == PMRD_TEST_BINOMIAL, "Unknown test type");
    const size_t c = (size_t)n_cells;
    DevBuf tot, var, neg, mean, negc, p, q, sig;
    PMRD_TRY(up(ctx, tot, h_n_total, c * 8));
    PMRD_TRY(up(ctx, var, h_n_variants, c * 8));
    PMRD_TRY(up(ctx, neg, h_is_negative, c));
    PMRD_CUDA(ctx, mean.alloc(ctx->stream, 8));
    PMRD_CUDA(ctx, negc.alloc(ctx->stream, 16));
    PMRD_CUDA(ctx, p.alloc(ctx->stream, c * 8));
    PMRD_CUDA(ctx, q.alloc(ctx->stream, c * 8));
    PMRD_CUDA(ctx, sig.alloc(ctx->stream, c));
    PMRD_TRY(call_counts_dev(ctx, tot.as<int64_t>(), var.as<int64_t>(), neg.as<uint8_t>(), n_cells, stream_id, test_type, alternative,
                             alpha, mean.as<double>(), negc.as<int64_t>(), p.as<double>(), q.as<double>(), sig.as<uint8_t>(), nullptr));
    PMRD_TRY(down(ctx, h_p_value, p.p, c * 8));
    PMRD_TRY(down(ctx, h_p_adjusted, q.p, c * 8));
    PMRD_TRY(down(ctx, h_significant, sig.p, c));
    PMRD_TRY(down(ctx, h_mean_error_rate, mean.p, 8));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- K15 / K16 -------------
int32_t pmrd_feature_table(pmrd_ctx* ctx, const int64_t* h_family_size, const double* h_quality_score,
                           const double* h_consensus_agreement, const double* h_allele_fraction, int64_t n, double* h_out) {
    ENTER(ctx);
    PMRD_ARG(ctx, n >= 0, "n");
    if (n == 0) return PMRD_OK;
    PMRD_ARG(ctx, h_family_size && h_quality_score && h_consensus_agreement && h_allele_fraction && h_out, "null buffer");
    DevBuf sz, q, a, f, out;
    PMRD_TRY(up(ctx, sz, h_family_size, (size_t)n * 8));
    PMRD_TRY(up(ctx, q, h_quality_score, (size_t)n * 8));
    PMRD_TRY(up(ctx, a, h_consensus_agreement, (size_t)n * 8));
    PMRD_TRY(up(ctx, f, h_allele_fraction, (size_t)n * 8));
    PMRD_CUDA(ctx, out.alloc(ctx->stream, (size_t)n * 48));
    PMRD_TRY(k15_gd_features(ctx, sz.as<int64_t>(), q.as<double>(), a.as<double>(), f.as<double>(), n, out.as<double>()));
    PMRD_TRY(down(ctx, h_out, out.p, (size_t)n * 48));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

int32_t pmrd_qc_family_stats(pmrd_ctx* ctx, const uint32_t* h_family_size, const uint32_t* h_flags, uint32_t flag_mask, int64_t n,
                             int32_t n_bins, uint64_t* h_hist, uint64_t* h_counts2) {
    ENTER(ctx);
    PMRD_ARG(ctx, n >= 0 && h_hist && h_counts2 && (n == 0 || h_family_size), "arguments");
    DevBuf sz, fl, hist, cnt;
    PMRD_TRY(up(ctx, sz, h_family_size, (size_t)n * 4));
    if (h_flags) PMRD_TRY(up(ctx, fl, h_flags, (size_t)n * 4));
    PMRD_CUDA(ctx, hist.alloc(ctx->stream, (size_t)(n_bins > 0 ? n_bins : 1) * 8));
    PMRD_CUDA(ctx, cnt.alloc(ctx->stream, 16));
    PMRD_TRY(k16_qc_family(ctx, sz.as<uint32_t>(), h_flags ? fl.as<uint32_t>() : nullptr, flag_mask, n, n_bins,
                           hist.as<unsigned long long>(), cnt.as<unsigned long long>()));
    PMRD_TRY(down(ctx, h_hist, hist.p, (size_t)n_bins * 8));
    PMRD_TRY(down(ctx, h_counts2, cnt.p, 16));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- K11 -------------------
int32_t pmrd_fit_lod_batch(pmrd_ctx* ctx, const double* h_af, const double* h_hits, int32_t n, int32_t n_series, double target,
                           int32_t n_boot, const uint64_t* h_stream_ids, double* h_out3) {
    ENTER(ctx);
    PMRD_ARG(ctx, h_af && h_hits && h_stream_ids && h_out3 && n >= 1 && n_series >= 1, "arguments");
    DevBuf af, hit, ids, out;
    PMRD_TRY(up(ctx, af, h_af, (size_t)n * 8));
    PMRD_TRY(up(ctx, hit, h_hits, (size_t)n * (size_t)n_series * 8));
    PMRD_TRY(up(ctx, ids, h_stream_ids, (size_t)n_series * 8));
    PMRD_CUDA(ctx, out.alloc(ctx->stream, (size_t)n_series * 24));
    PMRD_TRY(k11_fit_lod_batch(ctx, af.as<double>(), hit.as<double>(), n, n_series, target, n_boot, ids.as<uint64_t>(), out.as<double>()));
    PMRD_TRY(down(ctx, h_out3, out.p, (size_t)n_series * 24));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

int32_t pmrd_fit_lod(pmrd_ctx* ctx, const double* h_af, const double* h_hit, int32_t n, double target, int32_t n_boot,
                     uint64_t stream_id, double* h_out3) {
    return pmrd_fit_lod_batch(ctx, h_af, h_hit, n, 1, target, n_boot, &stream_id, h_out3);
}

}  // extern "C"
