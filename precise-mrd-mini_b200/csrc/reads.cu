// reads.cu -- read-level generator (K2 read level + K3 contamination injection).
//
// Restates SyntheticReadGenerator.generate_reads_for_site (src/precise_mrd/io.py:79-123):
//   family size  NegBinom(r=2, p=r/(r+mean)) + 1 (io.py:71-77)   [family_model 0]
//             or clip(Poisson(mean), 1, max) (north_star config 5) [family_model 1]
//   UMI          12 random bases = 24 random bits (io.py:60-63)
//   has_variant  U < allele_fraction (io.py:97); contamination flip (io.py:100-101)
//   per read     allele: variant family -> alt w.p. 0.9, else ref w.p. 0.95 (io.py:105-108)
//                quality max(10, min(40, int(N(30,5)))) (io.py:65-69), strand, read_pos in [10,150)
// and, at read level, what sim/contamination.py:230-278 does to the sample table:
//   index hopping      each read re-keyed to another cell of the batch w.p. hop_rate
//   barcode collision  a family takes the previous family's UMI w.p. collision_rate
// Draws are keyed (seed; stream READS, cell, family, read).  Reads are written through a
// two-level affine permutation (inside 4096-slot windows, and of the windows themselves) so
// the input to the sort is scrambled while stores stay L2-coalescible.
#include "kernels.cuh"
#include "rng.cuh"
#include "scan.cuh"
#include "reads.cuh"

__device__ __forceinline__ int find_cell(const int64_t* __restrict__ fam_offsets, int n_cells, int64_t f) {
    int lo = 0, hi = n_cells;  // fam_offsets[lo] <= f < fam_offsets[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (fam_offsets[mid] <= f) lo = mid;
        else hi = mid;
    }
    return lo;
}

struct FamDraw {
    int size;
    uint32_t umi;
    bool has_variant;
};

__device__ __forceinline__ FamDraw draw_family(const Philox& ph, const ReadCell& c, uint32_t fam, const pmrd_read_sim_params& sp) {
    const uint32_t c2 = (uint32_t)c.cell_id, c3 = (PMRD_STREAM_READS << 24) | (uint32_t)((c.cell_id >> 32) & 0xffffffu);
    const uint4 b = ph(0xffffffffu, fam, c2, c3);
    FamDraw d;
    const double u = u53(b.x, b.y);
    if (sp.family_model == 0) {
        const double p = 2.0 / (2.0 + sp.mean_family_size);
        d.size = sample_negbin_r2(u, p) + 1;
    } else {
        int s = sample_poisson_small(u, sp.mean_family_size);
        d.size = s < 1 ? 1 : s;
    }
    if (d.size > sp.max_family_size) d.size = sp.max_family_size;
    d.umi = b.z & 0xffffffu;
    // io.py:97 / :100-101
    const uint4 b2 = ph(0xfffffffeu, fam, c2, c3);
    d.has_variant = u53(b.w, b2.x) < c.af;
    if (sp.contamination_rate > 0.0 && u53(b2.y, b2.z) < sp.contamination_rate) d.has_variant = !d.has_variant;
    if (sp.collision_rate > 0.0 && fam > 0 && u24(b2.w) < (float)sp.collision_rate) {
        // barcode collision: share the previous molecule's UMI (two molecules, one barcode)
        const uint4 pb = ph(0xffffffffu, fam - 1, c2, c3);
        d.umi = pb.z & 0xffffffu;
    }
    return d;
}

__global__ void __launch_bounds__(256)
reads_size_kernel(uint64_t seed, const ReadCell* __restrict__ cells, const int64_t* __restrict__ fam_offsets, int n_cells,
                  int64_t n_fam, pmrd_read_sim_params sp, uint32_t* __restrict__ fam_size) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_fam) return;
    const Philox ph(seed);
    const int ci = find_cell(fam_offsets, n_cells, f);
    const ReadCell c = cells[ci];
    fam_size[f] = (uint32_t)draw_family(ph, c, (uint32_t)(f - c.fam_offset), sp).size;
}

#define SHUF_WINDOW 4096u

__device__ __forceinline__ int64_t shuffle_slot(int64_t x, int64_t n_reads, int shuffle) {
    if (!shuffle) return x;
    const int64_t n_full = n_reads / SHUF_WINDOW;  // complete windows; the tail window stays in place
    int64_t w = x / SHUF_WINDOW;
    uint32_t l = (uint32_t)(x % SHUF_WINDOW);
    if (w >= n_full) return x;
    l = (l * 2654435761u + 0x9e37u * (uint32_t)w) & (SHUF_WINDOW - 1);  // odd multiplier: bijection mod 4096
    // window permutation: multiply by a prime that does not divide n_full
    const int64_t A = 1000003;
    if (n_full % A != 0) w = (w * A + 12345) % n_full;
    return w * SHUF_WINDOW + l;
}

__global__ void __launch_bounds__(256)
reads_gen_kernel(uint64_t seed, const ReadCell* __restrict__ cells, const int64_t* __restrict__ fam_offsets, int n_cells,
                 int64_t n_fam, pmrd_read_sim_params sp, const uint64_t* __restrict__ read_offset /*[n_fam]*/,
                 int64_t n_reads, uint64_t* __restrict__ keys, uint32_t* __restrict__ payload) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_fam) return;
    const Philox ph(seed);
    const int ci = find_cell(fam_offsets, n_cells, f);
    const ReadCell c = cells[ci];
    const uint32_t fam = (uint32_t)(f - c.fam_offset);
    const FamDraw d = draw_family(ph, c, fam, sp);
    const uint32_t c2 = (uint32_t)c.cell_id, c3 = (PMRD_STREAM_READS << 24) | (uint32_t)((c.cell_id >> 32) & 0xffffffu);
    const int64_t off = (int64_t)read_offset[f];
    const uint64_t key_home = ((uint64_t)c.group << 32) | d.umi;
    for (int j = 0; j < d.size; ++j) {
        const uint4 r = ph((uint32_t)j, fam, c2, c3);
        // io.py:105-108
        const float ua = u24(r.x);
        uint32_t allele;
        if (d.has_variant) allele = ua < 0.9f ? c.alt_allele : c.ref_allele;
        else allele = ua < 0.95f ? c.ref_allele : c.alt_allele;
        // io.py:65-69: max(10, min(40, int(normal(30, 5))))  -- int() truncates toward zero
        const float z = sample_normal2f(u24(r.y), u24(r.z)).x;
        int q = (int)(30.0f + 5.0f * z);
        q = q < 10 ? 10 : (q > 40 ? 40 : q);
        const uint32_t strand = r.w & 1u;
        const uint32_t rpos = 10u + ((r.w >> 1) & 0xffffu) % 140u;  // io.py:119 randint(10, 150)
        uint32_t pl = allele | ((uint32_t)q << 2) | (strand << 8) | (rpos << 9) | ((allele == c.alt_allele ? 1u : 0u) << 31);
        uint64_t key = key_home;
        if (sp.hop_rate > 0.0 && n_cells > 1) {
            // index hopping: the read is assigned to another sample's index (contamination.py:230-254)
            const uint4 h = ph(0x80000000u | (uint32_t)j, fam, c2, c3);
            if (u24(h.x) < (float)sp.hop_rate) {
                const uint32_t other = (uint32_t)ci + 1u + h.y % (uint32_t)(n_cells - 1);
                const ReadCell oc = cells[other % (uint32_t)n_cells];
                key = ((uint64_t)oc.group << 32) | d.umi;
                pl = (pl & 0x7fffffffu) | ((allele == oc.alt_allele ? 1u : 0u) << 31);
            }
        }
        const int64_t dst = shuffle_slot(off + j, n_reads, sp.shuffle);
        keys[dst] = key;
        payload[dst] = pl;
    }
}

// ---------------------------------------------------------------- host side --------------
// ReadPlan: everything that depends only on (seed, cells, params) and is computed once --
// the device cell table, the family CSR and the total read count (a pure function of the
// Philox key, so it can be known before the timed step allocates anything).
int32_t reads_plan_build(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                         int64_t n_cells, uint32_t group_base, const pmrd_read_sim_params* sp, ReadPlan* plan,
                         int64_t* h_read_offsets /* optional [C+1] */) {
    PMRD_ARG(ctx, sp != nullptr && sp->mean_family_size > 0.0 && sp->max_family_size >= 1, "read sim params");
    PMRD_ARG(ctx, n_cells >= 0 && n_cells < (1ll << 31), "n_cells");
    plan->sp = *sp;
    plan->n_cells = n_cells;
    ReadCell* hc = (ReadCell*)malloc(sizeof(ReadCell) * (size_t)(n_cells + 1));
    int64_t* ho = (int64_t*)malloc(8 * (size_t)(n_cells + 1));
    if (!hc || !ho) { free(hc); free(ho); PMRD_ARG(ctx, false, "out of host memory"); }
    int64_t nf = 0;
    for (int64_t i = 0; i < n_cells; ++i) {
        hc[i].cell_id = (uint64_t)h_cell_id[i];
        hc[i].af = h_af ? h_af[i] : 0.0;
        hc[i].fam_offset = nf;
        hc[i].group = group_base + (uint32_t)i;
        hc[i].ref_allele = (uint32_t)(hc[i].cell_id & 3u);
        hc[i].alt_allele = (hc[i].ref_allele + 1u + (uint32_t)((hc[i].cell_id >> 2) % 3u)) & 3u;
        hc[i].pad = 0;
        ho[i] = nf;
        nf += h_n_families[i] > 0 ? h_n_families[i] : 0;
    }
    ho[n_cells] = nf;
    plan->n_fam = nf;
    plan->n_reads = 0;
    cudaError_t e = plan->cells.alloc(ctx->stream, sizeof(ReadCell) * (size_t)(n_cells + 1));
    if (e == cudaSuccess) e = plan->fam_offsets.alloc(ctx->stream, 8 * (size_t)(n_cells + 1));
    if (e == cudaSuccess) e = plan->fam_size.alloc(ctx->stream, 4 * (size_t)(nf + 1));
    if (e == cudaSuccess) e = plan->read_offset.alloc(ctx->stream, 8 * (size_t)(nf + 1));
    if (e == cudaSuccess) e = plan->total.alloc(ctx->stream, 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(plan->cells.p, hc, sizeof(ReadCell) * (size_t)n_cells, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(plan->fam_offsets.p, ho, 8 * (size_t)(n_cells + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    free(hc);
    if (e != cudaSuccess) {
        free(ho);
        snprintf(ctx->err, sizeof(ctx->err), "read plan: %s", cudaGetErrorString(e));
        return PMRD_ERR_CUDA;
    }
    int32_t st = reads_plan_sizes(ctx, plan);
    if (st == PMRD_OK && nf > 0) {
        uint64_t tot = 0;
        e = cudaMemcpyAsync(&tot, plan->total.p, 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "read plan: %s", cudaGetErrorString(e)); st = PMRD_ERR_CUDA; }
        plan->n_reads = (int64_t)tot;
    }
    if (st == PMRD_OK && h_read_offsets) {
        // per-cell CSR of reads (family-major layout): read_offset at each cell's first family
        uint64_t* tmp = (uint64_t*)malloc(8 * (size_t)(nf + 1));
        if (!tmp) { free(ho); PMRD_ARG(ctx, false, "out of host memory"); }
        e = cudaSuccess;
        if (nf > 0) e = cudaMemcpy(tmp, plan->read_offset.p, 8 * (size_t)nf, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "%s", cudaGetErrorString(e)); st = PMRD_ERR_CUDA; }
        else {
            for (int64_t i = 0; i < n_cells; ++i) h_read_offsets[i] = ho[i] < nf ? (int64_t)tmp[ho[i]] : plan->n_reads;
            h_read_offsets[n_cells] = plan->n_reads;
        }
        free(tmp);
    }
    free(ho);
    return st;
}

// family sizes + read offsets (device only, no sync): part of every simulate step
int32_t reads_plan_sizes(pmrd_ctx* ctx, ReadPlan* plan) {
    if (plan->n_fam <= 0) return PMRD_OK;
    PMRD_LAUNCH(ctx, reads_size_kernel, pmrd_blocks(plan->n_fam, 256), 256, 0, ctx->seed, plan->cells.as<ReadCell>(),
                plan->fam_offsets.as<int64_t>(), (int)plan->n_cells, plan->n_fam, plan->sp, plan->fam_size.as<uint32_t>());
    return exclusive_scan<uint32_t, uint64_t>(ctx, plan->fam_size.as<uint32_t>(), plan->read_offset.as<uint64_t>(), plan->n_fam,
                                              plan->total.as<uint64_t>());
}

// generate the reads (device only, no sync); capacity must be >= plan->n_reads
int32_t reads_plan_generate(pmrd_ctx* ctx, ReadPlan* plan, uint64_t* d_keys, uint32_t* d_payload) {
    if (plan->n_fam <= 0) return PMRD_OK;
    PMRD_LAUNCH(ctx, reads_gen_kernel, pmrd_blocks(plan->n_fam, 256), 256, 0, ctx->seed, plan->cells.as<ReadCell>(),
                plan->fam_offsets.as<int64_t>(), (int)plan->n_cells, plan->n_fam, plan->sp,
                (const uint64_t*)plan->read_offset.as<uint64_t>(), plan->n_reads, d_keys, d_payload);
    return PMRD_OK;
}
