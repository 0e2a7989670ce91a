// reads_dev.cuh -- device pieces of the read-level generator shared by reads.cu (materialised reads) and fused.cu
// (reads voted where they are drawn, never written): the per-family draw, the per-read draw and the quality tables.
// Both kernels MUST derive a read from the same Philox blocks in the same way -- the fused path is checked against
// the oracle's collapse of the very reads pmrd_simulate_reads materialises.
#pragma once
#include "kernels.cuh"
#include "rng.cuh"
#include "reads.cuh"

__device__ __forceinline__ int find_cell_in(const int64_t* __restrict__ fam_offsets, int lo, int hi, int64_t f) {
    while (hi - lo > 1) {  // fam_offsets[lo] <= f < fam_offsets[hi]
        const int mid = (lo + hi) >> 1;
        if (fam_offsets[mid] <= f) lo = mid;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int find_cell(const int64_t* __restrict__ fam_offsets, int n_cells, int64_t f) {
    return find_cell_in(fam_offsets, 0, n_cells, f);
}
// A block owns READS_FAM_BLOCK consecutive families.  blk_cell[b] = cell of the block's first family
// (built once on the host with the plan), so the per-family search runs over the cells
// [blk_cell[b], blk_cell[b + 1]] -- almost always one or two -- instead of all of them.
#define READS_FAM_BLOCK 256

struct FamDraw {
    int size;
    uint32_t umi;
    uint32_t key;      // key of the family's per-read Philox2x32 stream
    bool has_variant;
};

// s_cdf: Poisson cdf table in shared memory (128 floats, padded with 2.0f) or nullptr
// KNOBS / kn: the cell's row of per-cell knobs; KNOBS = false (plan-wide knobs only) is the hot configuration and compiles
// to exactly the code it had before per-cell knobs existed
template <bool KNOBS = false>
__device__ __forceinline__ FamDraw draw_family(const Philox& ph, const ReadCell& c, uint32_t fam, const pmrd_read_sim_params& sp,
                                               const float* s_cdf, const uint8_t* s_lut = nullptr, const CellKnobDev* kn = nullptr) {
    const uint32_t c2 = (uint32_t)c.cell_id, c3 = (PMRD_STREAM_READS << 24) | (uint32_t)((c.cell_id >> 32) & 0xffffffu);
    const uint4 b = ph(0xffffffffu, fam, c2, c3);
    FamDraw d;
    if (s_cdf) {
        // Poisson(mean <= 16) by inversion of a 24-bit uniform against the fp32 cdf table the plan
        // built on the host (the sizes only steer a simulation): size = #{k : F[k] < u}
        const float uf = u24(b.x);
        int k = 0;
        if (s_lut) {
            // 4096-bucket inverse of the table: the count for the bucket's smallest uniform, then a walk that almost
            // never takes a step (a bucket holds a cdf step with probability ~10 / 4096)
            k = s_lut[b.x >> 20];
            while (s_cdf[k] < uf) ++k;
        } else {
#pragma unroll
            for (int step = 64; step > 0; step >>= 1) k += (s_cdf[k + step - 1] < uf) ? step : 0;
        }
        d.size = k < 1 ? 1 : k;
    } else {
        const double u = u53(b.x, b.y);
        if (sp.family_model == 0) {
            const double p = 2.0 / (2.0 + sp.mean_family_size);
            d.size = sample_negbin_r2(u, p) + 1;
        } else {
            const int s = sample_poisson_small(u, sp.mean_family_size);
            d.size = s < 1 ? 1 : s;
        }
    }
    if (d.size > sp.max_family_size) d.size = sp.max_family_size;
    d.umi = b.z & 0xffffffu;
    d.key = b.y ^ (b.z >> 24);
    // io.py:97: has_variant = u53(b.w, b2.x) < af.  The top 32 bits settle it unless b.w sits on the
    // cell's threshold word (probability 2^-32), so the second Philox block is drawn only then, or
    // when a contamination knob needs its other words (io.py:100-101).
    if (!KNOBS) {
        // plan-wide knobs only
        const bool extras = sp.contamination_rate > 0.0 || sp.collision_rate > 0.0;
        const int64_t hw = (int64_t)b.w;
        if (!extras && (hw < c.af_lt || hw >= c.af_ge)) {
            d.has_variant = hw < c.af_lt;
            return d;
        }
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
    // per-cell knobs (pmrd_cell_knobs): the same draws, the cell's own collision rate, and molecules [n_own, ...) drawn
    // against the contaminating sample's allele fraction
    const CellKnobDev k = *kn;
    const bool extras = sp.contamination_rate > 0.0 || k.collision_rate > 0.0f;
    const bool own = fam < k.n_own;
    const int64_t lt = own ? c.af_lt : k.af2_lt, ge = own ? c.af_ge : k.af2_ge;
    const int64_t hw = (int64_t)b.w;
    if (!extras && (hw < lt || hw >= ge)) {
        d.has_variant = hw < lt;
        return d;
    }
    const uint4 b2 = ph(0xfffffffeu, fam, c2, c3);
    d.has_variant = u53(b.w, b2.x) < (own ? c.af : k.af2);
    if (sp.contamination_rate > 0.0 && u53(b2.y, b2.z) < sp.contamination_rate) d.has_variant = !d.has_variant;
    if (k.collision_rate > 0.0f && fam > 0 && u24(b2.w) < k.collision_rate) {
        const uint4 pb = ph(0xffffffffu, fam - 1, c2, c3);
        d.umi = pb.z & 0xffffffu;
    }
    return d;
}

// cdf of q = max(10, min(40, int(N(30,5)))) in units of 2^-24: c_qual_thr[k] = P(q <= 10 + k), k = 0..29
static __constant__ uint32_t c_qual_thr[32];
// 4096-bucket inverse of that cdf: c_qual_lut[b] = #{k < 30 : T[k] <= b << 12}; a bucket holds a
// threshold with probability 30 / 4096, so the walk after the lookup almost never takes a step
#define QUAL_LUT_BITS 12
static __device__ __align__(16) uint8_t c_qual_lut[1 << QUAL_LUT_BITS];   // global, not __constant__: every thread copies its own 16 bytes

// (every translation unit that includes this header owns a copy of the two tables and uploads it itself)
static inline int32_t reads_tables_upload_here(pmrd_ctx* ctx) {
    uint32_t h[32];
    uint8_t lut[1 << QUAL_LUT_BITS];          // (on the stack: contexts of different threads may build at once)
    for (int k = 0; k < 32; ++k) {
        // int() truncates toward zero and 30 + 5 z > 0 here, so q <= 10 + k  <=>  30 + 5 z < 11 + k
        const double zc = ((double)(11 + k) - 30.0) / 5.0;
        const double cdf = 0.5 * erfc(-zc / sqrt(2.0));
        h[k] = k < 30 ? (uint32_t)llround(cdf * 16777216.0) : 0xffffffffu;
    }
    for (uint32_t b = 0; b < (1u << QUAL_LUT_BITS); ++b) {
        int cnt = 0, cnt_end = 0;
        for (int k = 0; k < 30; ++k) {
            cnt += h[k] <= (b << (24 - QUAL_LUT_BITS));
            cnt_end += h[k] <= (((b + 1u) << (24 - QUAL_LUT_BITS)) - 1u);
        }
        // bit 7: a threshold lies inside the bucket, the count has to be finished by a walk (30 of 4096 buckets)
        lut[b] = (uint8_t)(cnt | (cnt_end != cnt ? 0x80 : 0));
    }
    PMRD_CUDA(ctx, cudaMemcpyToSymbolAsync(c_qual_thr, h, sizeof(h), 0, cudaMemcpyHostToDevice, ctx->stream));
    PMRD_CUDA(ctx, cudaMemcpyToSymbolAsync(c_qual_lut, lut, sizeof(lut), 0, cudaMemcpyHostToDevice, ctx->stream));
    return PMRD_OK;
}


// The per-read draws of io.py:105-119 from the 64 bits of one Philox2x32-10 block (key = the family's stream key,
// counter = (cell | read index, family)): returns the payload word  allele[1:0] qual[7:2] strand[8] read_pos[16:9]
// is_alt[31].  s_qlut / s_qthr: the block's shared-memory copies of c_qual_lut / c_qual_thr.
__device__ __forceinline__ uint32_t draw_read_payload(uint32_t key, uint32_t ctr0, uint32_t ctr1, uint32_t word,
                                                      const uint8_t* s_qlut, const uint32_t* s_qthr) {
    const uint32_t ref = (word >> 26) & 3u, alt = (word >> 28) & 3u;
    const bool has_variant = (word >> 24) & 1u;
    // 64 random bits per read: Philox2x32-10 keyed by the family, counter = (cell | read, family)
    const uint2 r2 = philox2x32(key, ctr0, ctr1);
    const uint32_t ra = r2.x, rb = r2.y;
    // io.py:105-108 (16-bit uniform: thresholds 0.9 / 0.95 are hit to 7e-6)
    const uint32_t ua = rb & 0xffffu;
    uint32_t allele;
    if (has_variant) allele = ua < 58982u ? alt : ref;
    else allele = ua < 62259u ? ref : alt;
    // io.py:65-69: max(10, min(40, int(normal(30, 5)))) by inversion of its exact 31-point cdf
    // (thresholds in units of 2^-24, built on the host from erfc): q = 10 + #{k : u >= T[k]}
    // (4096-bucket table, then a walk that almost never takes a step; T[30] = T[31] = 2^32 - 1 stops it)
    const uint32_t uq = ra >> 8;
    uint32_t qi = s_qlut[uq >> (24 - QUAL_LUT_BITS)];
    if (qi & 0x80u) {                                    // (rare) the bucket holds a cdf step
        qi &= 0x7fu;
        while (uq >= s_qthr[qi]) ++qi;
    }
    const int q = 10 + (int)qi;
    const uint32_t strand = ra & 1u;
    const uint32_t rpos = 10u + (((rb >> 16) * 140u) >> 16);     // io.py:119 randint(10, 150)
    return allele | ((uint32_t)q << 2) | (strand << 8) | (rpos << 9) | ((allele == alt ? 1u : 0u) << 31);
}
