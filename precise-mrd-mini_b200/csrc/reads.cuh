// reads.cuh -- host-visible pieces of the read-level generator (reads.cu).
#pragma once
#include "common.cuh"
#include <math.h>

struct ReadCell {
    uint64_t cell_id;
    double af;
    int64_t fam_offset;  // first global family index of this cell in the batch
    int64_t read_start;  // first read of the cell in family-major (unpadded) numbering
    int64_t out_start;   // where the cell's reads start in the output arrays (tile-aligned when segmented)
    uint32_t n_reads;    // reads of the cell
    uint32_t group;
    uint32_t ref_allele, alt_allele;
    int64_t af_lt, af_ge;  // has_variant threshold words (see draw_family)
    float inv_full;        // segmented layout: 1 / (complete 4096-read windows) when the window permutation applies, else 0
    uint32_t tile0;        // segmented layout: out_start / 4096
};

// per-cell contamination knobs (sim/contamination.py:230-311 at read level), kept OUT of ReadCell: the hot kernels load a
// ReadCell per block / family and must not carry fields only a contamination sweep reads.  One row per cell of a chunk,
// present only when the plan was created with pmrd_cell_knobs.
struct CellKnobDev {
    float hop_rate;        // each read is re-keyed to another cell of the cell's hop pool w.p. hop_rate
    float collision_rate;  // a molecule takes the previous molecule's UMI w.p. collision_rate
    uint32_t pool_start, pool_n;   // hop pool: cells [pool_start, pool_start + pool_n) of the chunk
    uint32_t n_own;        // molecules [0, n_own) are the sample's own; [n_own, ...) come from a contaminating sample ...
    uint32_t pad0;
    double af2;            // ... with this allele fraction (cross-sample contamination, contamination.py:280-311)
    int64_t af2_lt, af2_ge;
};

// host-side per-cell overrides handed to the plan builders (NULL: every cell takes the plan-wide parameters)
struct CellExtra {
    float hop_rate, collision_rate;
    uint32_t pool_start, pool_n;   // relative to the first cell of the call
    int64_t n_own;                 // own molecules (the rest of the cell's families are cross-sample contaminants); < 0: all
    double af2;
};

#define READS_SEG_TILE 4096 /* == RS_TILE: cells are padded to whole sort tiles when segmented */
/* padding reads carry a UMI with the top byte set (no 12-mer reaches it) and spread low bits, so
 * that they never pile up into one long run */
#define READS_IS_PAD_UMI(u) (((u) >> 24) == 0xffu)

struct ReadPlan {
    DevBuf knobs;   // CellKnobDev[n_cells] or empty (no per-cell knobs)
    DevBuf cells, fam_offsets, fam_size, fam_word, fam_key, read_offset, total, cell_tile_start, size_cdf, size_lut, blk_cell, tile_cell;
    bool has_size_cdf = false;  // Poisson sizes with mean <= 16: inverted against an fp32 cdf table
    pmrd_read_sim_params sp;
    int64_t n_cells = 0;
    int64_t n_fam = 0;
    int64_t n_reads = 0;      // reads generated
    int64_t n_out = 0;        // slots in the output arrays (== n_reads unless segmented: + padding)
    int64_t n_groups_with_reads = 0;  // cells that own at least one read
    int segmented = 0;        // 1: cell-major, tile-padded layout (per-cell 3-pass sort); 2: cell-major COMPACT layout (every
                              // cell <= 4096 reads, no padding: one block sorts and votes a cell in shared memory)
    int64_t max_cell_reads = 0;
    bool sizes_fresh = false; // the build just computed sizes / offsets for `sizes_seed`: the first step reuses them
    uint64_t sizes_seed = 0;
};

// allele convention of the generator: ref = cell_id & 3, alt = one of the other three
static inline uint32_t pmrd_cell_ref_allele(uint64_t cell_id) { return (uint32_t)(cell_id & 3u); }
static inline uint32_t pmrd_cell_alt_allele(uint64_t cell_id) {
    return (pmrd_cell_ref_allele(cell_id) + 1u + (uint32_t)((cell_id >> 2) % 3u)) & 3u;
}

// uniform in (0,1) from 53 bits (common.cuh u53) decides has_variant: words w of the first 32 bits that settle
// u53(w, .) < af on their own -- w < af_lt: true whatever the low word; w >= af_ge: false whatever the low word
static inline void reads_af_thresholds(double a, int64_t* lt, int64_t* ge) {
    int64_t g = a <= 0.0 ? 0 : (a >= 1.0 ? (1ll << 32) : (int64_t)(a * 4294967296.0));
    while (g > 0 && !(u53((uint32_t)(g - 1), 0xffffffffu) < a)) --g;
    while (g < (1ll << 32) && u53((uint32_t)g, 0xffffffffu) < a) ++g;
    *lt = g;
    while (g < (1ll << 32) && u53((uint32_t)g, 0u) < a) ++g;
    *ge = g;
}
static inline void reads_init_cell(ReadCell* c, uint64_t cell_id, double af, int64_t fam_offset, uint32_t group) {
    c->cell_id = cell_id;
    c->af = af;
    c->fam_offset = fam_offset;
    c->group = group;
    c->ref_allele = (uint32_t)(cell_id & 3u);
    c->alt_allele = (c->ref_allele + 1u + (uint32_t)((cell_id >> 2) % 3u)) & 3u;
    c->read_start = 0;
    c->out_start = 0;
    c->n_reads = 0;
    c->inv_full = 0.0f;
    c->tile0 = 0;
    reads_af_thresholds(af, &c->af_lt, &c->af_ge);
}

// device row of a cell's knobs (`first` = index of the chunk's first cell in the call, `n_chunk` its cell count: hop pools
// are re-based to the chunk and must lie inside it)
static inline void reads_fill_knob(CellKnobDev* k, const ReadCell* c, const CellExtra* x, int64_t first, int64_t n_chunk) {
    k->hop_rate = x->hop_rate;
    k->collision_rate = x->collision_rate;
    k->pool_start = 0;
    k->pool_n = (uint32_t)n_chunk;
    if (x->pool_n > 0) { k->pool_start = (uint32_t)((int64_t)x->pool_start - first); k->pool_n = x->pool_n; }
    k->n_own = 0xffffffffu;
    k->pad0 = 0;
    k->af2 = c->af;
    k->af2_lt = c->af_lt;
    k->af2_ge = c->af_ge;
    if (x->n_own >= 0) {
        k->n_own = (uint32_t)x->n_own;
        k->af2 = x->af2;
        reads_af_thresholds(x->af2, &k->af2_lt, &k->af2_ge);
    }
}

// Poisson family sizes with mean <= 16 are inverted against an fp32 cdf table: F[k] = P(X <= k), k < 96, by the pmf
// recurrence, padded so the search stops; lut[b] = #{k : F[k] < u} for the smallest uniform of bucket b (u24 of the
// words b << 20 ..).  Returns whether the tables apply.
static inline bool reads_build_size_tables(const pmrd_read_sim_params* sp, float h_cdf[128], uint8_t h_lut[4096]) {
    if (!(sp->family_model == 1 && sp->mean_family_size <= 16.0)) return false;
    const float lam = (float)sp->mean_family_size;
    float p = expf(-lam), F = p;
    for (int k = 0; k < 128; ++k) {
        if (k > 0 && k < 96) { p *= lam / (float)k; F += p; }
        h_cdf[k] = k < 96 ? F : 2.0f;
    }
    for (uint32_t b = 0; b < 4096u; ++b) {
        const float u = u24(b << 20);
        int k = 0;
        while (k < 127 && h_cdf[k] < u) ++k;
        h_lut[b] = (uint8_t)k;
    }
    return true;
}

int32_t reads_plan_build(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                         int64_t n_cells, uint32_t group_base, const pmrd_read_sim_params* sp, int segmented, ReadPlan* plan,
                         int64_t* h_read_offsets, const CellExtra* h_extra = nullptr, int64_t first_cell = 0);
int32_t reads_plan_sizes(pmrd_ctx* ctx, ReadPlan* plan);
int32_t reads_plan_generate(pmrd_ctx* ctx, ReadPlan* plan, uint64_t* d_keys, uint32_t* d_payload, uint32_t* d_tile_hist = nullptr);
int32_t reads_plan_generate_cell_major(pmrd_ctx* ctx, ReadPlan* plan, uint64_t* d_keys, uint32_t* d_payload);
int32_t reads_plan_cell_counts(pmrd_ctx* ctx, ReadPlan* plan, int64_t* d_out);
int32_t reads_upload_tables(pmrd_ctx* ctx);
