// api.cu -- the extern "C" boundary (include/pmrd.h): context management and the
// host-buffer wrappers that stage data through device memory around the K* kernels.
#include <math.h>
#include <new>

#include "kernels.cuh"
#include "radix_sort.cuh"

char g_pmrd_create_err[512] = "";

int32_t collapse_upload_weights(pmrd_ctx* ctx, const double* h_w64);
int32_t k6_hamming(pmrd_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, int64_t n, int32_t* d_d);

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
    cudaStreamDestroy(ctx->own_stream);
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

}  // extern "C"

// ---------------------------------------------------------------- staging helpers -------
#define ENTER(ctx)                              \
    if (!(ctx)) return PMRD_ERR_ARG;            \
    (ctx)->err[0] = 0;                          \
    PMRD_CUDA(ctx, cudaSetDevice((ctx)->device))

static int32_t up(pmrd_ctx* ctx, DevBuf& b, const void* h, size_t bytes) {
    PMRD_CUDA(ctx, b.alloc(ctx->stream, bytes));
    if (bytes) PMRD_CUDA(ctx, cudaMemcpyAsync(b.p, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PMRD_OK;
}
static int32_t down(pmrd_ctx* ctx, void* h, const void* d, size_t bytes) {
    if (bytes && h) PMRD_CUDA(ctx, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PMRD_OK;
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
int32_t pmrd_collapse_reads_dev(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp,
                                uint32_t* d_payload_tmp, int64_t n_reads, const pmrd_umi_params* params,
                                pmrd_family_table* d_fam, int64_t capacity_f, int64_t* n_families,
                                pmrd_group_table* d_grp, int64_t capacity_g, int64_t* n_groups) {
    ENTER(ctx);
    PMRD_ARG(ctx, d_fam && n_families, "family table");
    return k5_collapse(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n_reads, params, d_fam, capacity_f, n_families,
                       d_grp, capacity_g, n_groups);
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
    PMRD_CUDA(ctx, fk.alloc(ctx->stream, n * 8));
    PMRD_CUDA(ctx, fs.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, fa.alloc(ctx->stream, n * 4));
    PMRD_CUDA(ctx, fc.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, fq.alloc(ctx->stream, n * 4));
    PMRD_CUDA(ctx, fn.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, ff.alloc(ctx->stream, n * 4));
    pmrd_family_table dfam{fk.as<uint64_t>(), fs.as<uint32_t>(), fa.as<uint32_t>(), fc.as<uint32_t>(),
                           fq.as<uint32_t>(), fn.as<uint32_t>(), ff.as<uint32_t>()};
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
    if (nf > capacity_f || (want_groups && ng > capacity_g)) {
        snprintf(ctx->err, sizeof(ctx->err), "output capacity too small: %lld families, %lld groups", (long long)nf, (long long)ng);
        return PMRD_ERR_CAPACITY;
    }
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

}  // extern "C"
