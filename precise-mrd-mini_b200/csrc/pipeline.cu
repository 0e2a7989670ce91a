// pipeline.cu -- simulate -> collapse -> call for a batch of cells with reads resident in HBM.
// This is the bench "step" and the engine under eval-lob / eval-lod / eval-contamination:
//   reads.cu  (K2/K3)  generate reads of every cell of the batch, scrambled
//   collapse  (K4-K7)  radix sort -> families -> weighted consensus -> per-cell counts
//   K10-lite           error rate from the negative-control cells exactly as
//                      fit_error_model + call_mrd combine them: mean over 64 contexts of
//                      (n_var/n_neg) * U(0.5,2)      (error_model.py:153-156, call.py:178)
//   K8, K9             two-sided tail p per cell, BH over the batch (call.py:182-229)
#include "kernels.cuh"
#include "rng.cuh"
#include <new>
#include <stdlib.h>

#include "reads.cuh"
#include "fused.cuh"

int32_t k5_collapse_exact_nosync(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp,
                                 uint32_t* d_payload_tmp, int64_t n_reads, uint64_t varying, const pmrd_umi_params* params,
                                 const pmrd_family_table* d_fam, int64_t capacity_f, const pmrd_group_table* d_grp,
                                 int64_t capacity_g, uint32_t* d_fam_start, uint32_t* d_grp_start, uint32_t* d_counts,
                                 const int64_t* d_cell_tile_start, int64_t n_seg_cells, int seg_key_bits);

int32_t k5_collapse_counts_small(pmrd_ctx* ctx, const uint64_t* d_rec, const ReadCell* d_cells, int64_t n_cells, int64_t max_cell_reads,
                                 const pmrd_umi_params* params, int64_t cell0, const uint8_t* d_alt_allele, int64_t* d_o_fam,
                                 int64_t* d_o_tot, int64_t* d_o_var);

int32_t k5_collapse_counts_segmented(pmrd_ctx* ctx, uint64_t* d_rec, uint64_t* d_rec_tmp, int64_t n_out,
                                     const pmrd_umi_params* params, const int64_t* d_cell_tile_start, const int32_t* d_tile_cell,
                                     int64_t n_cells,
                                     int sort_bits, int umi_bits, int64_t cell0, const uint8_t* d_alt_allele, int64_t* d_o_fam,
                                     int64_t* d_o_tot, int64_t* d_o_var, uint32_t* d_hist0);

// scatter the group table into per-cell arrays (cells without reads keep zeros)
__global__ void __launch_bounds__(256)
cell_counts_kernel(const uint32_t* __restrict__ grp_id, const uint32_t* __restrict__ grp_cons /*[G*4]*/,
                   const uint32_t* __restrict__ d_ngroups, int64_t cell0, int64_t n_cells,
                   const uint8_t* __restrict__ alt_allele, int64_t* __restrict__ n_total, int64_t* __restrict__ n_variants) {
    // group id = cell index inside the chunk; outputs are indexed by cell0 + group
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int64_t)*d_ngroups || g >= n_cells) return;
    const int64_t c = (int64_t)grp_id[g];
    if (c < 0 || c >= n_cells) return;
    const uint32_t* cc = grp_cons + g * 4;
    n_total[cell0 + c] = (int64_t)cc[0] + cc[1] + cc[2] + cc[3];
    n_variants[cell0 + c] = cc[alt_allele[cell0 + c]];
}

// per-cell family counts: number of family-table rows per group
__global__ void __launch_bounds__(256)
cell_family_count_kernel(const uint64_t* __restrict__ fam_key, const uint32_t* __restrict__ d_nfam, int64_t cell0,
                         int64_t n_cells, unsigned long long* __restrict__ n_families) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // (families whose UMI is the padding sentinel of a tile-aligned cell are not families)
    const bool valid = f < (int64_t)*d_nfam && !READS_IS_PAD_UMI((uint32_t)fam_key[f]);
    const uint32_t g = valid ? (uint32_t)(fam_key[f] >> 32) : 0xffffffffu;
    const uint32_t m = __match_any_sync(0xffffffffu, g);
    if (valid && (m & lanemask_lt()) == 0) {
        const int64_t c = (int64_t)g;
        if (c >= 0 && c < n_cells) atomicAdd(&n_families[cell0 + c], (unsigned long long)__popc(m));
    }
}

// one block per CALL GROUP (cells [gs[g], gs[g+1]); gs == NULL: one group of all cells):
// rate = sum(n_var)/sum(n_tot) over the group's negative cells; mean of 64 context rates
__global__ void __launch_bounds__(256)
error_rate_kernel(uint64_t seed, uint64_t stream_id, const uint64_t* __restrict__ group_stream /* [G] or NULL */,
                  const int64_t* __restrict__ gs /* [G+1] or NULL */, const int64_t* __restrict__ n_total,
                  const int64_t* __restrict__ n_variants, const uint8_t* __restrict__ is_negative, int64_t n_cells,
                  double* __restrict__ mean_rate /*[G]*/, int64_t* __restrict__ neg_counts /*[2G]*/, int32_t* __restrict__ cell_group /*[C] or NULL*/) {
    __shared__ unsigned long long s_n[8], s_v[8];
    __shared__ double s_mod[64];
    const int g = blockIdx.x;
    const int64_t lo = gs ? gs[g] : 0, hi = gs ? gs[g + 1] : n_cells;
    if (group_stream) stream_id = group_stream[g];
    unsigned long long n = 0, v = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        if (cell_group) cell_group[i] = g;
        if (is_negative[i]) {
            n += (unsigned long long)n_total[i];
            v += (unsigned long long)n_variants[i];
        }
    }
    n = warp_sum(n);
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) { s_n[threadIdx.x >> 5] = n; s_v[threadIdx.x >> 5] = v; }
    if (threadIdx.x < 64) {
        const Philox ph(seed);
        const uint4 b = ph(0, threadIdx.x, (uint32_t)stream_id, (PMRD_STREAM_ERRMODEL << 24) | (uint32_t)((stream_id >> 32) & 0xffffffu));
        s_mod[threadIdx.x] = 0.5 + 1.5 * u53(b.x, b.y);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long nn = 0, vv = 0;
        for (int w = 0; w < 8; ++w) { nn += s_n[w]; vv += s_v[w]; }
        const double vr = nn > 0 ? (double)vv / (double)nn : 0.0;
        // error_rate.mean() over the 64 contexts (call.py:178), summed in context order
        double acc = 0.0;
        for (int c = 0; c < 64; ++c) acc += vr * s_mod[c];
        mean_rate[g] = acc / 64.0;
        neg_counts[2 * g] = (int64_t)nn;
        neg_counts[2 * g + 1] = (int64_t)vv;
    }
}

// ---------------------------------------------------------------- plan -------------------
// A plan owns every buffer of one grid of cells, so that run() only enqueues kernels: no
// allocation, no host sync, no host<->device copy inside the step.  The grid is cut into
// CHUNKS of consecutive cells (<= 2^chunk_log2 reads each, default 2^30: 17 GB of record buffers) that stream through
// ONE set of read / family buffers sized for the largest chunk -- 180 GB of HBM is plenty, but
// a chunk must stay below 2^31 reads (32-bit positions in the sort) and small chunks keep
// the working set of a radix pass closer to L2.  Per-cell counts land in grid-wide arrays;
// the error model, the tail tests and BH then run ONCE over all cells, as call_mrd does.
#include <vector>

struct PlanChunk {
    ReadPlan reads;
    int64_t cell0 = 0, n_cells = 0, fcap = 0;
    uint64_t varying = 0;
};

struct pmrd_plan {
    std::vector<PlanChunk*> chunks;
    std::vector<FusedChunk*> fchunks;   // PMRD_ENGINE_FUSED: reads are never written (fused.cu)
    DevBuf fscratch;                    // candidate-duplicate list + its sort buffers, sized for the largest chunk
    bool fused = false;
    pmrd_umi_params up;
    double neg_af_max, alpha;
    int32_t test_type, alternative;
    int64_t n_cells = 0, n_reads = 0, n_fam = 0;
    int64_t max_reads = 0, max_fcap = 0, max_cells = 0;
    int sort_bits = 16;   // UMI bits ordered by the radix sort; the consensus kernel resolves the rest inside its tile
                          // (16: two passes + cell_consensus_peel_kernel, the default; 24: three passes + cell_consensus_bulk_kernel)
    uint64_t stream_id = 0;
    uint64_t seed = 0;    // the context seed the layout was sized for (read counts are a function of it)
    DevBuf keys, keys_tmp, pl, pl_tmp;
    DevBuf tile_hist;     // [max tiles][256]: first-pass digit histogram the generator counts as a by-product (cell-major path)
    bool gen_hist = true;
    DevBuf fk, fs, fa, fc, fq, fn, ff, fam_start;
    DevBuf gg, gf, gc, ga, gt, grp_start, counts, totals;
    DevBuf alt, neg, negc, mean_rate;
    DevBuf o_reads, o_fam, o_tot, o_var, o_p, o_q, o_sig;
    DevBuf bh_ws;         // BH workspace (sort buffers + scratch): the call stage of a step allocates nothing
    // call groups (pmrd_cell_knobs.call_group_start): error rate, tail tests and BH per group of cells
    int64_t n_groups = 1;
    DevBuf gs, gstream, cell_group;   // [G+1] cell ranges, [G] stream ids (a group's first cell id), [C] group of every cell
    // general-path (hopping) chunks size these separately from the cell-major ones
    int64_t max_reads_gen = 0, max_fcap_gen = 0, max_cells_gen = 0;
    ~pmrd_plan() {
        for (PlanChunk* c : chunks) delete c;
        for (FusedChunk* c : fchunks) delete c;
    }
};

static uint64_t varying_mask_for(int64_t n_cells) {
    uint64_t m = 0xffffffull;  // 24 UMI bits
    int gb = 0;
    while ((1ll << gb) < n_cells) ++gb;
    if (gb > 0) m |= ((1ull << gb) - 1ull) << 32;
    return m;
}

__global__ void add_counts_kernel(const uint32_t* __restrict__ counts, unsigned long long* __restrict__ totals) {
    totals[0] += counts[0];
    totals[1] += counts[1];
}

int32_t plan_create(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families_in,
                    int64_t n_cells, const pmrd_read_sim_params* sp, const pmrd_umi_params* up, double neg_af_max,
                    int32_t test_type, int32_t alternative, double alpha, const pmrd_cell_knobs* knobs, pmrd_plan** out) {
    PMRD_ARG(ctx, out != nullptr, "out");
    *out = nullptr;
    PMRD_ARG(ctx, n_cells > 0 && n_cells < (1ll << 30), "n_cells");
    PMRD_ARG(ctx, h_cell_id && h_n_families_in && sp, "null cell arrays / params");
    // per-cell knobs -> host-side CellExtra rows + effective family counts (own + cross-sample contaminants)
    std::vector<CellExtra> extra;
    std::vector<int64_t> nf_eff;
    const int64_t* h_n_families = h_n_families_in;
    if (knobs && (knobs->hop_rate || knobs->collision_rate || knobs->cross_proportion)) {
        PMRD_ARG(ctx, !knobs->cross_proportion || knobs->cross_af, "cross_proportion needs cross_af");
        PMRD_ARG(ctx, !knobs->hop_rate || (knobs->hop_pool_start && knobs->hop_pool_size), "hop_rate needs hop_pool_start / hop_pool_size");
        extra.resize((size_t)n_cells);
        nf_eff.resize((size_t)n_cells);
        for (int64_t i = 0; i < n_cells; ++i) {
            CellExtra& x = extra[(size_t)i];
            x.hop_rate = knobs->hop_rate ? (float)knobs->hop_rate[i] : (float)sp->hop_rate;
            x.collision_rate = knobs->collision_rate ? (float)knobs->collision_rate[i] : (float)sp->collision_rate;
            x.pool_start = 0; x.pool_n = 0;
            if (knobs->hop_rate && knobs->hop_rate[i] > 0.0) {
                const int64_t ps = knobs->hop_pool_start[i], pn = knobs->hop_pool_size[i];
                PMRD_ARG(ctx, ps >= 0 && pn >= 1 && ps <= i && i < ps + pn && ps + pn <= n_cells, "hop pool must be a cell range containing the cell");
                x.pool_start = (uint32_t)ps; x.pool_n = (uint32_t)pn;
            }
            const int64_t own = h_n_families_in[i] > 0 ? h_n_families_in[i] : 0;
            x.n_own = -1; x.af2 = 0.0;
            int64_t add = 0;
            if (knobs->cross_proportion && knobs->cross_proportion[i] > 0.0) {
                add = (int64_t)((double)own * knobs->cross_proportion[i]);          // contamination.py:288
                x.n_own = own; x.af2 = knobs->cross_af[i];
            }
            nf_eff[(size_t)i] = own + add;
        }
        h_n_families = nf_eff.data();
    }
    const CellExtra* h_extra = extra.empty() ? nullptr : extra.data();
    auto cell_hops = [&](int64_t i) { return h_extra ? h_extra[i].hop_rate > 0.0f : sp->hop_rate > 0.0; };
    // a chunk may end before cell i unless that would split a hop pool or mix hopping and non-hopping cells is avoided
    auto pool_open_at = [&](int64_t i) { return h_extra && i > 0 && h_extra[i].pool_n > 0 && (int64_t)h_extra[i].pool_start < i; };
    PMRD_ARG(ctx, up && up->mode == PMRD_MODE_WEIGHTED && up->cluster == PMRD_CLUSTER_EXACT, "pipeline needs WEIGHTED/EXACT umi params");
    PMRD_ARG(ctx, test_type == PMRD_TEST_POISSON || test_type == PMRD_TEST_BINOMIAL, "Unknown test type");
    PMRD_ARG(ctx, sp->chunk_log2 == 0 || (sp->chunk_log2 >= 16 && sp->chunk_log2 <= 30), "chunk_log2 must be 0 or in [16, 30]");
    pmrd_plan* P = new (std::nothrow) pmrd_plan();
    PMRD_ARG(ctx, P != nullptr, "out of host memory");
    P->up = *up; P->neg_af_max = neg_af_max; P->alpha = alpha; P->test_type = test_type; P->alternative = alternative;
    P->n_cells = n_cells;
    P->fused = sp->engine == PMRD_ENGINE_FUSED;
    {
        bool any_hop = false;
        for (int64_t i = 0; i < n_cells; ++i) any_hop |= cell_hops(i);
        if (P->fused && any_hop) {
            delete P;
            PMRD_ARG(ctx, false, "the fused engine keeps a read inside its cell: hop_rate > 0 needs PMRD_ENGINE_MATERIALISED");
        }
    }
    P->stream_id = (uint64_t)h_cell_id[0];
    P->seed = ctx->seed;
    if (const char* sb = getenv("PMRD_SORT_BITS")) {   // 24: order every UMI bit (3 radix passes); 16: the default; 8: one pass (slow peel)
        const int v = atoi(sb);
        if (v == 8 || v == 16 || v == 24) P->sort_bits = v;
    }
    // cut into chunks by expected reads (mean size with 25% head room; exact counts follow)
    const double budget = (double)(1ll << (sp->chunk_log2 ? sp->chunk_log2 : 30));
    const double per_family = (sp->mean_family_size + 1.0) * 1.25;
    int64_t c0 = 0;
    int32_t st = PMRD_OK;
    while (c0 < n_cells && st == PMRD_OK) {
        double acc = 0.0;
        int64_t c1 = c0;
        while (c1 < n_cells) {
            const double add = (double)(h_n_families[c1] > 0 ? h_n_families[c1] : 0) * per_family;
            // a chunk is either all hopping cells (general path) or none (cell-major path), and never splits a hop pool
            if (c1 > c0 && cell_hops(c1) != cell_hops(c0)) break;
            if (c1 > c0 && acc + add > budget && !pool_open_at(c1)) break;
            acc += add;
            ++c1;
        }
        if (P->fused) {
            FusedChunk* fc = new (std::nothrow) FusedChunk();
            if (!fc) { st = PMRD_ERR_ARG; snprintf(ctx->err, sizeof(ctx->err), "out of host memory"); break; }
            P->fchunks.push_back(fc);
            st = fused_chunk_build(ctx, h_cell_id + c0, h_af ? h_af + c0 : nullptr, h_n_families + c0, c1 - c0, c0, sp, up, fc,
                                   h_extra ? h_extra + c0 : nullptr);
            if (st != PMRD_OK) break;
            P->n_reads += fc->n_reads;
            P->n_fam += fc->n_fam;
            if (fc->n_cand > P->max_fcap) P->max_fcap = fc->n_cand;
            c0 = c1;
            continue;
        }
        PlanChunk* ch = new (std::nothrow) PlanChunk();
        if (!ch) { st = PMRD_ERR_ARG; snprintf(ctx->err, sizeof(ctx->err), "out of host memory"); break; }
        P->chunks.push_back(ch);
        ch->cell0 = c0;
        ch->n_cells = c1 - c0;
        st = reads_plan_build(ctx, h_cell_id + c0, h_af ? h_af + c0 : nullptr, h_n_families + c0, c1 - c0, 0, sp, /*segmented=*/1,
                              &ch->reads, nullptr, h_extra ? h_extra + c0 : nullptr, c0);
        if (st != PMRD_OK) break;
        if (ch->reads.n_out >= (1ll << 31) - 8192) {
            snprintf(ctx->err, sizeof(ctx->err), "a chunk of %lld cells holds %lld reads (>= 2^31): lower chunk_log2 or split the cell",
                     (long long)ch->n_cells, (long long)ch->reads.n_reads);
            st = PMRD_ERR_ARG;
            break;
        }
        // UMI merging can only lower the family count; hopped reads can open new (cell, UMI) families
        ch->fcap = !ch->reads.segmented ? ch->reads.n_reads : ch->reads.n_fam + ch->n_cells /* padding sentinels */;
        if (ch->fcap < 1) ch->fcap = 1;
        if (!ch->reads.segmented) {
            if (ch->reads.n_out > P->max_reads_gen) P->max_reads_gen = ch->reads.n_out;
            if (ch->fcap > P->max_fcap_gen) P->max_fcap_gen = ch->fcap;
            if (ch->n_cells > P->max_cells_gen) P->max_cells_gen = ch->n_cells;
        }
        ch->varying = varying_mask_for(ch->n_cells);
        P->n_reads += ch->reads.n_reads;
        P->n_fam += ch->reads.n_fam;
        if (ch->reads.n_out > P->max_reads) P->max_reads = ch->reads.n_out;
        if (ch->fcap > P->max_fcap) P->max_fcap = ch->fcap;
        if (ch->n_cells > P->max_cells) P->max_cells = ch->n_cells;
        c0 = c1;
    }
    if (st != PMRD_OK) { delete P; return st; }
    const size_t n = (size_t)(P->max_reads > 0 ? P->max_reads : 1);
    const size_t c = (size_t)n_cells;
    cudaError_t e = cudaSuccess;
    auto A = [&](DevBuf& b, size_t bytes) { if (e == cudaSuccess) e = b.alloc(ctx->stream, bytes); };
    if (P->fused) A(P->fscratch, fused_scratch_bytes(P->max_fcap));
    else { A(P->keys, n * 8); A(P->keys_tmp, n * 8); }
    if (const char* gh = getenv("PMRD_GEN_HIST")) P->gen_hist = atoi(gh) != 0;   // 0: the sort counts its first digit itself
    if (P->fused) P->gen_hist = false;
    if (P->gen_hist) A(P->tile_hist, (n / READS_SEG_TILE + 1) * 256 * sizeof(uint32_t));
    if (P->max_reads_gen > 0) {   // only the general (non cell-major) path carries payload arrays and the family / group tables
        const size_t ng = (size_t)P->max_reads_gen;
        A(P->pl, ng * 4); A(P->pl_tmp, ng * 4);
    }
    const size_t f = (size_t)(P->max_fcap_gen > 0 ? P->max_fcap_gen : 1), g = (size_t)(P->max_cells_gen > 0 ? P->max_cells_gen : 1);
    if (P->max_reads_gen > 0) {
        A(P->fk, f * 8); A(P->fs, f * 4); A(P->fa, f * 4); A(P->fc, f * 4); A(P->fq, f * 4); A(P->fn, f * 4); A(P->ff, f * 4);
        A(P->fam_start, (f + 1) * 4);
        A(P->gg, g * 4); A(P->gf, g * 4); A(P->gc, g * 16); A(P->ga, g * 8); A(P->gt, g * 8); A(P->grp_start, (g + 1) * 4);
    }
    // call groups
    std::vector<int64_t> h_gs;
    if (knobs && knobs->call_group_start && knobs->n_call_groups > 0) {
        P->n_groups = knobs->n_call_groups;
        h_gs.assign(knobs->call_group_start, knobs->call_group_start + P->n_groups + 1);
        bool ok = h_gs.front() == 0 && h_gs.back() == n_cells;
        for (int64_t gi = 0; gi < P->n_groups && ok; ++gi)
            ok = h_gs[(size_t)gi + 1] > h_gs[(size_t)gi] && (P->n_groups == 1 || h_gs[(size_t)gi + 1] - h_gs[(size_t)gi] <= 4096);
        if (!ok) { delete P; PMRD_ARG(ctx, false, "call_group_start must partition the cells into non-empty ranges of <= 4096 cells"); }
    } else {
        h_gs = {0, n_cells};
    }
    const size_t G = (size_t)P->n_groups;
    A(P->counts, 8); A(P->totals, 16); A(P->alt, c); A(P->neg, c); A(P->negc, 16 * G); A(P->mean_rate, 8 * G);
    A(P->gs, 8 * (G + 1)); A(P->gstream, 8 * G); A(P->cell_group, 4 * c);
    A(P->o_reads, c * 8); A(P->o_fam, c * 8); A(P->o_tot, c * 8); A(P->o_var, c * 8); A(P->o_p, c * 8); A(P->o_q, c * 8); A(P->o_sig, c);
    A(P->bh_ws, bh_workspace_bytes((int64_t)c));
    uint8_t* h_alt = (uint8_t*)malloc(c * 2);
    if (e == cudaSuccess && h_alt) {
        uint8_t* h_neg = h_alt + c;
        std::vector<uint64_t> h_gstream(G);
        for (size_t gi = 0; gi < G; ++gi) {            // negative controls per call group (error_model.py:144-149)
            const int64_t lo = h_gs[gi], hi = h_gs[gi + 1];
            h_gstream[gi] = (uint64_t)h_cell_id[lo];
            bool any_neg = false;
            double min_af = INFINITY;
            for (int64_t i = lo; i < hi; ++i) {
                h_alt[i] = (uint8_t)pmrd_cell_alt_allele((uint64_t)h_cell_id[i]);
                const double a = h_af ? h_af[i] : 0.0;
                h_neg[i] = a <= neg_af_max;
                any_neg |= h_neg[i] != 0;
                if (a < min_af) min_af = a;
            }
            if (!any_neg)  // fallback: the lowest-AF samples
                for (int64_t i = lo; i < hi; ++i) h_neg[i] = (h_af ? h_af[i] : 0.0) == min_af;
        }
        e = cudaMemcpyAsync(P->alt.p, h_alt, c, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(P->gs.p, h_gs.data(), 8 * (G + 1), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(P->gstream.p, h_gstream.data(), 8 * G, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);      // (h_gstream leaves scope)
        if (e == cudaSuccess) e = cudaMemcpyAsync(P->neg.p, h_neg, c, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(P->totals.p, 0, 16, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(P->o_tot.p, 0, c * 8, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(P->o_var.p, 0, c * 8, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(P->o_fam.p, 0, c * 8, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    const bool oom = h_alt == nullptr;
    free(h_alt);
    if (e != cudaSuccess || oom) {
        snprintf(ctx->err, sizeof(ctx->err), "plan_create: %s", oom ? "out of host memory" : cudaGetErrorString(e));
        delete P;
        return oom ? PMRD_ERR_ARG : PMRD_ERR_CUDA;
    }
    *out = P;
    return PMRD_OK;
}

// ---- stages of one chunk; each only enqueues kernels --------------------------------------
static int32_t chunk_simulate(pmrd_ctx* ctx, pmrd_plan* P, PlanChunk* ch) {
    // family sizes -> read offsets -> reads (scrambled).  The plan build has just run the first two for
    // this seed (it needs the read counts to size the buffers): the first step after a build does not
    // repeat them; every later step does the full work.
    if (!(ch->reads.sizes_fresh && ch->reads.sizes_seed == ctx->seed)) PMRD_TRY(reads_plan_sizes(ctx, &ch->reads));
    ch->reads.sizes_fresh = false;
    // cell-major chunks are generated as packed 8-byte records (payload:32 | umi:32)
    return reads_plan_generate(ctx, &ch->reads, P->keys.as<uint64_t>(), ch->reads.segmented ? nullptr : P->pl.as<uint32_t>(),
                               P->gen_hist ? P->tile_hist.as<uint32_t>() : nullptr);
}

static int32_t chunk_collapse(pmrd_ctx* ctx, pmrd_plan* P, PlanChunk* ch) {
    if (ch->reads.segmented == 2)      // compact chunk: every cell sorted and voted inside one block
        return k5_collapse_counts_small(ctx, P->keys.as<uint64_t>(), ch->reads.cells.as<ReadCell>(), ch->n_cells, ch->reads.max_cell_reads,
                                        &P->up, ch->cell0, (const uint8_t*)P->alt.as<uint8_t>(), P->o_fam.as<int64_t>(),
                                        P->o_tot.as<int64_t>(), P->o_var.as<int64_t>());
    if (ch->reads.segmented) {
        // cell-major input: per-cell 3-pass sort, then ONE kernel from sorted reads to per-cell counts
        // (packed 8-byte records: P->keys / P->keys_tmp are the ping-pong pair, no payload arrays)
        return k5_collapse_counts_segmented(ctx, P->keys.as<uint64_t>(), P->keys_tmp.as<uint64_t>(), ch->reads.n_out, &P->up,
                                            (const int64_t*)ch->reads.cell_tile_start.as<int64_t>(),
                                            (const int32_t*)ch->reads.tile_cell.as<int32_t>(), ch->n_cells, P->sort_bits, 24,
                                            ch->cell0, (const uint8_t*)P->alt.as<uint8_t>(), P->o_fam.as<int64_t>(),
                                            P->o_tot.as<int64_t>(), P->o_var.as<int64_t>(),
                                            P->gen_hist ? P->tile_hist.as<uint32_t>() : nullptr);
    }
    pmrd_family_table fam{P->fk.as<uint64_t>(), P->fs.as<uint32_t>(), P->fa.as<uint32_t>(), P->fc.as<uint32_t>(),
                          P->fq.as<uint32_t>(), P->fn.as<uint32_t>(), P->ff.as<uint32_t>()};
    pmrd_group_table grp{P->gg.as<uint32_t>(), P->gf.as<uint32_t>(), P->gc.as<uint32_t>(), P->ga.as<uint64_t>(), P->gt.as<uint64_t>()};
    PMRD_CUDA(ctx, cudaMemsetAsync(P->counts.p, 0, 8, ctx->stream));
    if (ch->reads.n_reads > 0)
        PMRD_TRY(k5_collapse_exact_nosync(ctx, P->keys.as<uint64_t>(), P->pl.as<uint32_t>(), P->keys_tmp.as<uint64_t>(),
                                          P->pl_tmp.as<uint32_t>(), ch->reads.n_out, ch->varying, &P->up, &fam, ch->fcap, &grp,
                                          ch->n_cells, P->fam_start.as<uint32_t>(), P->grp_start.as<uint32_t>(),
                                          P->counts.as<uint32_t>(), nullptr, 0, 0));
    return PMRD_OK;
}

// per-cell counts of the chunk into the grid-wide arrays (K7); the segmented path already did it
static int32_t chunk_counts(pmrd_ctx* ctx, pmrd_plan* P, PlanChunk* ch) {
    if (ch->reads.segmented) return PMRD_OK;
    PMRD_LAUNCH(ctx, cell_counts_kernel, pmrd_blocks(ch->n_cells, 256), 256, 0, (const uint32_t*)P->gg.as<uint32_t>(),
                (const uint32_t*)P->gc.as<uint32_t>(), (const uint32_t*)(P->counts.as<uint32_t>() + 1), ch->cell0, ch->n_cells,
                (const uint8_t*)P->alt.as<uint8_t>(), P->o_tot.as<int64_t>(), P->o_var.as<int64_t>());
    PMRD_LAUNCH(ctx, cell_family_count_kernel, pmrd_blocks(ch->fcap, 256), 256, 0, (const uint64_t*)P->fk.as<uint64_t>(),
                (const uint32_t*)P->counts.as<uint32_t>(), ch->cell0, ch->n_cells, (unsigned long long*)P->o_fam.as<int64_t>());
    PMRD_LAUNCH(ctx, add_counts_kernel, 1, 1, 0, (const uint32_t*)P->counts.as<uint32_t>(), P->totals.as<unsigned long long>());
    return PMRD_OK;
}

static int32_t plan_reset_outputs(pmrd_ctx* ctx, pmrd_plan* P) {
    const size_t c = (size_t)P->n_cells;
    PMRD_CUDA(ctx, cudaMemsetAsync(P->o_tot.p, 0, 8 * c, ctx->stream));
    PMRD_CUDA(ctx, cudaMemsetAsync(P->o_var.p, 0, 8 * c, ctx->stream));
    PMRD_CUDA(ctx, cudaMemsetAsync(P->o_fam.p, 0, 8 * c, ctx->stream));
    PMRD_CUDA(ctx, cudaMemsetAsync(P->totals.p, 0, 16, ctx->stream));
    return PMRD_OK;
}

// call stage over ALL cells: error rate from the negative controls, tail p per cell, BH (K8, K9).
// Device arrays in, device arrays out; shared by the plan and by pmrd_call_counts (the multi-GPU finish:
// every rank runs it over the all-gathered counts, so the table is the single-GPU one bit for bit).
int32_t call_counts_dev(pmrd_ctx* ctx, const int64_t* d_tot, const int64_t* d_var, const uint8_t* d_neg, int64_t n_cells,
                        uint64_t stream_id, int32_t test_type, int32_t alternative, double alpha, double* d_mean_rate,
                        int64_t* d_negc, double* d_p, double* d_q, uint8_t* d_sig, void* d_bh_workspace) {
    PMRD_LAUNCH(ctx, error_rate_kernel, 1, 256, 0, ctx->seed, stream_id, (const uint64_t*)nullptr, (const int64_t*)nullptr, d_tot, d_var, d_neg,
                n_cells, d_mean_rate, d_negc, (int32_t*)nullptr);
    PMRD_TRY(k8_pvalues(ctx, d_var, d_tot, d_mean_rate, 0, n_cells, test_type, alternative, d_p));
    PMRD_TRY(k9_bh(ctx, d_p, n_cells, alpha, d_sig, d_q, d_bh_workspace));
    return PMRD_OK;
}

int32_t plan_run_call(pmrd_ctx* ctx, pmrd_plan* P) {
    if (P->n_groups <= 1 && P->n_cells > 4096)
        return call_counts_dev(ctx, P->o_tot.as<int64_t>(), P->o_var.as<int64_t>(), P->neg.as<uint8_t>(), P->n_cells, P->stream_id,
                               P->test_type, P->alternative, P->alpha, P->mean_rate.as<double>(), P->negc.as<int64_t>(),
                               P->o_p.as<double>(), P->o_q.as<double>(), P->o_sig.as<uint8_t>(), P->bh_ws.p);
    // several call groups: each is what ONE call_mrd of the reference sees -- its own error rate, its own BH.
    // (A single group of <= 4096 cells takes this path too: BH by one block that ranks by counting -- 3 launches for
    // the whole call stage instead of the 45 of a radix sort over 62 key bits; bit-equal to it.)
    PMRD_LAUNCH(ctx, error_rate_kernel, (int)P->n_groups, 256, 0, ctx->seed, (uint64_t)0, (const uint64_t*)P->gstream.as<uint64_t>(),
                (const int64_t*)P->gs.as<int64_t>(), (const int64_t*)P->o_tot.as<int64_t>(), (const int64_t*)P->o_var.as<int64_t>(),
                (const uint8_t*)P->neg.as<uint8_t>(), P->n_cells, P->mean_rate.as<double>(), P->negc.as<int64_t>(), P->cell_group.as<int32_t>());
    PMRD_TRY(k8_pvalues(ctx, P->o_var.as<int64_t>(), P->o_tot.as<int64_t>(), P->mean_rate.as<double>(), 0, P->n_cells, P->test_type,
                        P->alternative, P->o_p.as<double>(), (const int32_t*)P->cell_group.as<int32_t>()));
    return k9_bh_groups(ctx, P->o_p.as<double>(), (const int64_t*)P->gs.as<int64_t>(), P->n_groups, P->alpha, P->o_sig.as<uint8_t>(), P->o_q.as<double>());
}

// Stage entry points for per-stage timing.  With more than one chunk the read buffers are
// reused, so simulate / collapse of different chunks cannot be separated: these run the
// stage for chunk `which` only (bench.py times chunk 0), plan_run() is the real step.
// the cell layout (reads per cell, tiles, buffer sizes) is a function of the seed the plan was built with
#define PLAN_SEED_GUARD(ctx, P) PMRD_ARG(ctx, (ctx)->seed == (P)->seed, "the plan was built for another seed: create a new plan after pmrd_set_seed")

int32_t plan_run_simulate(pmrd_ctx* ctx, pmrd_plan* P, int64_t which) {
    PLAN_SEED_GUARD(ctx, P);
    PMRD_ARG(ctx, !P->fused, "the fused engine has no separate simulate / collapse stages");
    PMRD_ARG(ctx, which >= 0 && which < (int64_t)P->chunks.size(), "chunk index");
    return chunk_simulate(ctx, P, P->chunks[(size_t)which]);
}
int32_t plan_run_collapse(pmrd_ctx* ctx, pmrd_plan* P, int64_t which) {
    PMRD_ARG(ctx, which >= 0 && which < (int64_t)P->chunks.size(), "chunk index");
    PlanChunk* ch = P->chunks[(size_t)which];
    if (which == 0) PMRD_TRY(plan_reset_outputs(ctx, P));
    PMRD_TRY(chunk_collapse(ctx, P, ch));
    return chunk_counts(ctx, P, ch);
}

int32_t plan_run(pmrd_ctx* ctx, pmrd_plan* P) {
    PLAN_SEED_GUARD(ctx, P);
    if (P->fused) {
        // every cell's counts are WRITTEN by its block (no reset needed); the replay of the candidate duplicates adds to them
        for (FusedChunk* fc : P->fchunks)
            PMRD_TRY(fused_chunk_run(ctx, fc, &P->up, P->fscratch.p, (const uint8_t*)P->alt.as<uint8_t>(), P->o_reads.as<int64_t>(),
                                     P->o_fam.as<int64_t>(), P->o_tot.as<int64_t>(), P->o_var.as<int64_t>()));
        return plan_run_call(ctx, P);
    }
    PMRD_TRY(plan_reset_outputs(ctx, P));
    for (PlanChunk* ch : P->chunks) {
        PMRD_TRY(chunk_simulate(ctx, P, ch));
        PMRD_TRY(chunk_collapse(ctx, P, ch));
        PMRD_TRY(chunk_counts(ctx, P, ch));
    }
    return plan_run_call(ctx, P);
}

__global__ void __launch_bounds__(256) sum_i64_kernel(const int64_t* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long a = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a += (unsigned long long)v[i];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, a);
}

int32_t plan_results(pmrd_ctx* ctx, pmrd_plan* P, const pmrd_pipeline_result* h, double* h_mean_rate, int64_t* h_totals /*[3]: reads, families, groups*/) {
    const size_t c = (size_t)P->n_cells;
    for (PlanChunk* ch : P->chunks) PMRD_TRY(reads_plan_cell_counts(ctx, &ch->reads, P->o_reads.as<int64_t>() + ch->cell0));
    // (fused engine: the cell kernel writes the per-cell read counts itself)
    auto D = [&](void* dst, const DevBuf& b, size_t bytes) -> cudaError_t {
        return dst ? cudaMemcpyAsync(dst, b.p, bytes, cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
    };
    if (h) {
        PMRD_CUDA(ctx, D(h->n_reads, P->o_reads, c * 8)); PMRD_CUDA(ctx, D(h->n_families, P->o_fam, c * 8));
        PMRD_CUDA(ctx, D(h->n_total, P->o_tot, c * 8)); PMRD_CUDA(ctx, D(h->n_variants, P->o_var, c * 8));
        PMRD_CUDA(ctx, D(h->p_value, P->o_p, c * 8)); PMRD_CUDA(ctx, D(h->p_adjusted, P->o_q, c * 8));
        PMRD_CUDA(ctx, D(h->significant, P->o_sig, c));
    }
    PMRD_CUDA(ctx, D(h_mean_rate, P->mean_rate, 8));
    unsigned long long tot[2] = {0, 0};
    // families formed = sum of the per-cell counts (the family table also holds padding sentinels)
    PMRD_CUDA(ctx, cudaMemsetAsync(P->totals.p, 0, 8, ctx->stream));
    PMRD_LAUNCH(ctx, sum_i64_kernel, 64, 256, 0, (const int64_t*)P->o_fam.as<int64_t>(), P->n_cells, P->totals.as<unsigned long long>());
    PMRD_CUDA(ctx, cudaMemcpyAsync(tot, P->totals.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_totals) {
        h_totals[0] = P->n_reads;
        h_totals[1] = (int64_t)tot[0];
        int64_t groups = (int64_t)tot[1];
        for (PlanChunk* ch : P->chunks)
            if (ch->reads.segmented) groups += ch->reads.n_groups_with_reads;
        for (FusedChunk* fc : P->fchunks) groups += fc->n_cells_with_reads;
        h_totals[2] = groups;
    }
    return PMRD_OK;
}

// mean error rate of every call group (after a run); synchronises
int32_t plan_group_rates(pmrd_ctx* ctx, pmrd_plan* P, double* h_rates, int64_t cap) {
    PMRD_ARG(ctx, cap >= P->n_groups, "capacity < number of call groups");
    PMRD_CUDA(ctx, cudaMemcpyAsync(h_rates, P->mean_rate.p, 8 * (size_t)P->n_groups, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}

void plan_destroy(pmrd_plan* P) { delete P; }
int64_t plan_n_reads(const pmrd_plan* P) { return P ? P->n_reads : 0; }
int64_t plan_n_families(const pmrd_plan* P) { return P ? P->n_fam : 0; }
int64_t plan_n_chunks(const pmrd_plan* P) { return P ? (int64_t)(P->fused ? P->fchunks.size() : P->chunks.size()) : 0; }
int64_t plan_chunk_reads(const pmrd_plan* P, int64_t which) {
    if (P && P->fused) return (which >= 0 && which < (int64_t)P->fchunks.size()) ? P->fchunks[(size_t)which]->n_reads : 0;
    return (P && which >= 0 && which < (int64_t)P->chunks.size()) ? P->chunks[(size_t)which]->reads.n_reads : 0;
}

int32_t plan_info(const pmrd_plan* P, int64_t* o) {
    if (!P || !o) return PMRD_ERR_ARG;
    if (P->fused) {   // no records, no radix passes; [1] = candidate duplicates replayed per step, [5] = 12 B per candidate
        int64_t cand = 0;
        for (const FusedChunk* fc : P->fchunks) cand += fc->n_cand;
        o[0] = P->n_reads; o[1] = cand; o[2] = P->n_fam; o[3] = (int64_t)P->fchunks.size(); o[4] = 0; o[5] = 12; o[6] = P->max_fcap; o[7] = 2;
        return PMRD_OK;
    }
    int64_t n_out = 0, max_out = 0, seg = 1, passes = 0;
    for (const PlanChunk* ch : P->chunks) {
        n_out += ch->reads.n_out;
        if (ch->reads.n_out > max_out) max_out = ch->reads.n_out;
        if (!ch->reads.segmented) seg = 0;
    }
    if (seg) passes = (P->sort_bits + 7) / 8;
    else if (!P->chunks.empty()) {
        uint64_t v = P->chunks[0]->varying;
        int b = 0;
        while (b < 64) {
            if (!((v >> b) & 1ull)) { ++b; continue; }
            int w = 64 - b < 8 ? 64 - b : 8;
            while (w > 1 && !((v >> (b + w - 1)) & 1ull)) --w;
            ++passes;
            b += w;
        }
    }
    o[0] = P->n_reads; o[1] = n_out; o[2] = P->n_fam; o[3] = (int64_t)P->chunks.size(); o[4] = passes; o[5] = seg ? 8 : 12;
    o[6] = max_out; o[7] = seg;
    return PMRD_OK;
}
