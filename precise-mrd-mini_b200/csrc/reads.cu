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
#include "reads_dev.cuh"
#include <stdlib.h>

template <bool KNOBS>
__global__ void __launch_bounds__(256)
reads_size_kernel(uint64_t seed, const ReadCell* __restrict__ cells, const int64_t* __restrict__ fam_offsets, int n_cells,
                  int64_t n_fam, pmrd_read_sim_params sp, const float* __restrict__ size_cdf /* [128] or NULL */,
                  const uint8_t* __restrict__ size_lut /* [4096] or NULL */, const int32_t* __restrict__ blk_cell,
                  uint32_t* __restrict__ fam_size, uint32_t* __restrict__ fam_word, uint32_t* __restrict__ fam_key,
                  const CellKnobDev* __restrict__ knobs /* [n_cells] or NULL */) {
    __shared__ float s_cdf[128];
    __shared__ __align__(16) uint8_t s_lut[4096];
    static_assert(READS_FAM_BLOCK * 16 == 4096, "one 16-byte LUT slice per thread");
    const int64_t f = (int64_t)blockIdx.x * READS_FAM_BLOCK + threadIdx.x;
    if (size_cdf && threadIdx.x < 128) s_cdf[threadIdx.x] = size_cdf[threadIdx.x];
    if (size_cdf && size_lut) reinterpret_cast<uint4*>(s_lut)[threadIdx.x] = reinterpret_cast<const uint4*>(size_lut)[threadIdx.x];
    __syncthreads();
    if (f >= n_fam) return;
    const Philox ph(seed);
    const int ci = find_cell_in(fam_offsets, blk_cell[blockIdx.x], blk_cell[blockIdx.x + 1] + 1, f);
    const ReadCell c = cells[ci];
    const FamDraw d = draw_family<KNOBS>(ph, c, (uint32_t)(f - c.fam_offset), sp, size_cdf ? s_cdf : nullptr,
                                         (size_cdf && size_lut) ? s_lut : nullptr, KNOBS ? knobs + ci : nullptr);
    fam_size[f] = (uint32_t)d.size;
    fam_word[f] = d.umi | ((d.has_variant ? 1u : 0u) << 24);   // the generator re-reads this instead of re-drawing
    fam_key[f] = d.key;
}

#define SHUF_WINDOW 4096u

int32_t reads_upload_tables(pmrd_ctx* ctx) { return reads_tables_upload_here(ctx); }

// bijection of [0, n): an affine permutation inside each complete 4096-slot window and an
// affine permutation w -> (97 w + 13) mod n_full of the complete windows themselves (97 is prime:
// a bijection unless it divides n_full); the tail window stays in place.  inv_full = 1 / n_full as
// fp32 when the window permutation applies (n_full <= 2^16, so 97 w + 13 < 2^24 is exact in fp32
// and the quotient estimate is off by at most one), else 0.
#define SHUF_MUL 97u
__host__ __device__ __forceinline__ float shuffle_inv(uint32_t n) {
    const uint32_t n_full = n / SHUF_WINDOW;
    return (n_full > 1u && n_full <= 65536u && n_full % SHUF_MUL != 0u) ? 1.0f / (float)n_full : 0.0f;
}
// destination of complete window w (< n_full) under the window permutation
__device__ __forceinline__ uint32_t shuffle_window(uint32_t w, uint32_t n_full, float inv_full) {
    if (inv_full == 0.0f) return w;
    const uint32_t y = w * SHUF_MUL + 13u;
    int32_t r = (int32_t)(y - (uint32_t)((float)y * inv_full) * n_full);
    r += r < 0 ? (int32_t)n_full : 0;
    r -= r >= (int32_t)n_full ? (int32_t)n_full : 0;
    return (uint32_t)r;
}
// per-cell read counts from the family-major read offsets
__global__ void __launch_bounds__(256)
cell_read_count_kernel(const uint64_t* __restrict__ read_offset, const int64_t* __restrict__ fam_offsets, int64_t n_cells,
                       int64_t n_fam, const uint64_t* __restrict__ total /* reads of the batch (device) */, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const int64_t n_reads = (int64_t)total[0];
    const int64_t f0 = fam_offsets[i], f1 = fam_offsets[i + 1];
    const int64_t r0 = f0 < n_fam ? (int64_t)read_offset[f0] : n_reads;
    const int64_t r1 = f1 < n_fam ? (int64_t)read_offset[f1] : n_reads;
    out[i] = r1 - r0;
}

// One warp generates the reads of 32 consecutive families, FLATTENED: lane l first draws
// family (warp_base + l) and parks its attributes in shared memory; the warp then walks the
// contiguous range of reads those families own, one read per lane per step, so lanes stay
// busy whatever the family sizes are (a per-family loop leaves ~45 % of the lanes idle).
//
// Where a read goes: read x of its cell (family-major numbering) belongs to window x / 4096, and a
// window lands as a whole in one output tile.  A family touches at most two windows unless it is longer
// than a window, so the lane that parks the family also works out the two destination tiles once
// (tile_a, tile_b; bit 31 = "this window is not scrambled") and a read only picks one of them.
#define SLOT_IDENT 0x80000000u
struct __align__(16) FamSlot {
    uint32_t local_off;   // first read of the family, relative to the warp's first read
    uint32_t cell_off;    // first read of the family, relative to its cell (or to the chunk)
    uint32_t tile_a;      // destination tile of window cell_off / 4096
    uint32_t tile_b;      // ... and of the next window
    uint32_t key;         // Philox2x32 key of the family's read stream
    uint32_t ctr_hi;      // low 16 bits of the cell id << 16 (counter word 0 = ctr_hi | read index)
    uint32_t fam;         // family index inside the cell (counter word 1)
    uint32_t word;        // umi[23:0] | has_variant[24] | ref[27:26] | alt[29:28]
    uint32_t group;
    int32_t ci;
    uint32_t cell_n;      // reads in that cell (or chunk): range of the shuffle (slow path, hopping)
    uint32_t tile0;       // first output tile of the cell
};

// destination tile (| SLOT_IDENT) of source window w of a cell with n reads whose output starts at tile0
__device__ __forceinline__ uint32_t window_tile(uint32_t w, uint32_t n, float inv_full, int shuffle, uint32_t tile0) {
    const uint32_t n_full = n / SHUF_WINDOW;
    if (!shuffle || w >= n_full) return (tile0 + w) | SLOT_IDENT;
    return tile0 + shuffle_window(w, n_full, inv_full);
}

template <bool KNOBS>
__global__ void __launch_bounds__(256)
reads_gen_kernel(uint64_t seed, const ReadCell* __restrict__ cells, const int64_t* __restrict__ fam_offsets, int n_cells,
                 int64_t n_fam, pmrd_read_sim_params sp, const uint64_t* __restrict__ read_offset /*[n_fam]*/,
                 const uint32_t* __restrict__ fam_size, const uint32_t* __restrict__ fam_word, const uint32_t* __restrict__ fam_key,
                 const int32_t* __restrict__ blk_cell, int64_t n_reads, int segmented, uint64_t* __restrict__ keys,
                 uint32_t* __restrict__ payload, uint32_t* __restrict__ tile_hist /* [n_tiles][256], zeroed, or NULL */,
                 const CellKnobDev* __restrict__ knobs /* [n_cells] or NULL: hop rates and pools per cell */) {
    __shared__ FamSlot s_slot[8][32];
    __shared__ uint32_t s_qthr[32];
    __shared__ __align__(16) uint8_t s_qlut[1 << QUAL_LUT_BITS];
    if (threadIdx.x < 32) s_qthr[threadIdx.x] = c_qual_thr[threadIdx.x];
    static_assert((1 << QUAL_LUT_BITS) == 16 * READS_FAM_BLOCK, "one 16-byte LUT slice per thread");
    reinterpret_cast<uint4*>(s_qlut)[threadIdx.x] = reinterpret_cast<const uint4*>(c_qual_lut)[threadIdx.x];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t f = (int64_t)blockIdx.x * READS_FAM_BLOCK + threadIdx.x;
    const bool valid = f < n_fam;
    int64_t off = 0;
    uint32_t size = 0;
    if (valid) {
        const int ci = find_cell_in(fam_offsets, blk_cell[blockIdx.x], blk_cell[blockIdx.x + 1] + 1, f);
        const ReadCell c = cells[ci];
        off = (int64_t)read_offset[f];
        size = fam_size[f];
        const uint32_t fw = fam_word[f];
        FamSlot& S = s_slot[warp][lane];
        S.fam = (uint32_t)(f - c.fam_offset);
        S.key = fam_key[f];
        S.ctr_hi = (uint32_t)c.cell_id << 16;
        S.word = fw | (c.ref_allele << 26) | (c.alt_allele << 28);
        S.group = c.group;
        S.ci = ci;
        float inv_full;
        if (segmented) {
            S.cell_off = (uint32_t)(off - c.read_start);
            S.cell_n = c.n_reads;
            S.tile0 = c.tile0;
            inv_full = c.inv_full;                   // (per cell: worked out on the host with the plan)
        } else {
            S.cell_off = (uint32_t)off;
            S.cell_n = (uint32_t)n_reads;
            S.tile0 = 0u;
            inv_full = shuffle_inv(S.cell_n);
        }
        const uint32_t w0 = S.cell_off / SHUF_WINDOW;
        if (segmented == 2) {
            // compact layout (cells of <= 4096 reads, back to back, no padding): the cell's reads are scrambled by an
            // odd-multiplier affine map inside its largest power-of-two prefix, the rest stays in place
            S.tile0 = (uint32_t)c.out_start;
            S.tile_a = (sp.shuffle && c.n_reads > 1u) ? 1u << (31 - __clz((int)c.n_reads)) : 0u;
            S.tile_b = 0u;
        } else {
            S.tile_a = window_tile(w0, S.cell_n, inv_full, sp.shuffle, S.tile0);
            // (the next window's tile only matters to a family that reaches into it: 1 in 800 at five reads per family)
            S.tile_b = S.cell_off % SHUF_WINDOW + size > SHUF_WINDOW ? window_tile(w0 + 1u, S.cell_n, inv_full, sp.shuffle, S.tile0) : 0u;
        }
        if (tile_hist) {
            // By-product for the collapse stage: the first radix pass needs, per 4096-slot tile of the
            // output, the histogram of the low UMI byte.  A window of the family-major read order lands
            // in ONE tile and a family's reads share their UMI, so the whole histogram costs one
            // reduction per (family, window) -- 0.2 per read -- instead of a pass over the 8 B/read.
            uint32_t x = S.cell_off, rem = size;
            uint32_t* const col = tile_hist + (fw & 0xffu);
            while (rem) {
                const uint32_t w = x / SHUF_WINDOW;
                const uint32_t room = SHUF_WINDOW - x % SHUF_WINDOW;
                const uint32_t take = rem < room ? rem : room;
                const uint32_t t = w == w0 ? S.tile_a : (w == w0 + 1u ? S.tile_b : window_tile(w, S.cell_n, inv_full, sp.shuffle, S.tile0));
                atomicAdd(col + (size_t)(t & ~SLOT_IDENT) * 256, take);
                x += take;
                rem -= take;
            }
        }
    }
    const int64_t r0 = __shfl_sync(0xffffffffu, off, 0);
    // sorted ascending over the lanes; lanes past the last family never own a read
    const uint32_t loc = valid ? (uint32_t)(off - r0) : 0xffffffffu;
    s_slot[warp][lane].local_off = loc;
    // reads owned by the warp: [0, total)
    const uint32_t end = valid ? loc + size : 0u;
    uint32_t total = end;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total = max(total, __shfl_xor_sync(0xffffffffu, total, o));
    __syncthreads();
    const uint32_t le = lanemask_lt() | (1u << lane);
    const bool hopping = sp.hop_rate > 0.0 && n_cells > 1;
    uint32_t before = 0;                             // families whose first read lies before this step
    for (uint32_t k0 = 0; k0 < total; k0 += 32) {
        const uint32_t k = k0 + lane;
        // owner of read k = (number of families whose first read is <= k) - 1.  Lane s still holds
        // family s's first read `loc` (ascending, distinct: every family has >= 1 read): the
        // families that start inside this step are marked in a 32-bit mask by one warp OR-reduction,
        // those that started before it are the running sum of the earlier steps' marks, and a popcount finishes the job.
        const uint32_t rel = loc - k0;
        const uint32_t heads = __reduce_or_sync(0xffffffffu, rel < 32u ? 1u << rel : 0u);
        const uint32_t before_now = before;
        before += __popc(heads);
        if (k >= total) continue;
        const uint32_t lo = before_now + __popc(heads & le) - 1u;
        const FamSlot& S = s_slot[warp][lo];
        const uint4 sa = *reinterpret_cast<const uint4*>(&S.local_off);   // local_off, cell_off, tile_a, tile_b
        const uint4 sb = *reinterpret_cast<const uint4*>(&S.key);         // key, ctr_hi, fam, word
        const uint32_t j = k - sa.x;
        const uint32_t word = sb.w;
        const uint32_t umi = word & 0xffffffu, alt = (word >> 28) & 3u;
        uint32_t pl = draw_read_payload(sb.x, sb.y | j, sb.z, word, s_qlut, s_qthr);
        const uint32_t allele = pl & 3u;
        uint32_t group = 0;
        if (payload) group = S.group;
        if (hopping) {
            // index hopping: the read is assigned to another sample's index (contamination.py:230-254)
            const Philox ph(seed);
            const uint32_t c2 = (uint32_t)cells[S.ci].cell_id;
            const uint32_t c3 = (PMRD_STREAM_READS << 24) | (uint32_t)((cells[S.ci].cell_id >> 32) & 0xffffffu);
            const uint4 h = ph(0x80000000u | j, sb.z, c2, c3);  // (a block of its own: hopping is rare)
            // per-cell rate and hop pool = cells [pool_start, pool_start + pool_n) when the plan carries knobs, else the whole batch
            const float rate = KNOBS ? knobs[S.ci].hop_rate : (float)sp.hop_rate;
            const uint32_t p0 = KNOBS ? knobs[S.ci].pool_start : 0u, pn = KNOBS ? knobs[S.ci].pool_n : (uint32_t)n_cells;
            if (pn > 1u && u24(h.x) < rate) {
                const uint32_t other = p0 + ((uint32_t)S.ci - p0 + 1u + h.y % (pn - 1u)) % pn;
                const ReadCell oc = cells[other];
                group = oc.group;
                pl = (pl & 0x7fffffffu) | ((allele == oc.alt_allele ? 1u : 0u) << 31);
            }
        }
        // slot: window x / 4096 goes to one tile; inside a scrambled window the slot is an odd-multiplier affine map
        uint32_t x = sa.y + j;
        int64_t dst;
        if (segmented == 2) {
            if (x < sa.z) x = (x * 2654435761u + 0x9e37u) & (sa.z - 1u);      // sa.z: power-of-two prefix of the cell
            dst = (int64_t)S.tile0 + x;                                        // S.tile0: the cell's first slot
        } else {
            const uint32_t w = x / SHUF_WINDOW, dw = w - sa.y / SHUF_WINDOW;
            uint32_t t = dw == 0u ? sa.z : sa.w;
            if (dw > 1u) t = window_tile(w, S.cell_n, shuffle_inv(S.cell_n), sp.shuffle, S.tile0);   // a family longer than a window
            uint32_t l = x % SHUF_WINDOW;
            if (!(t & SLOT_IDENT)) l = (l * 2654435761u + 0x9e37u * w) & (SHUF_WINDOW - 1);
            dst = (int64_t)(t & ~SLOT_IDENT) * SHUF_WINDOW + l;
        }
        if (payload) {
            keys[dst] = ((uint64_t)group << 32) | umi;
            payload[dst] = pl;
        } else {
            // packed 8-byte record (cell-major layout: the cell is implied by the position)
            keys[dst] = ((uint64_t)pl << 32) | umi;
        }
    }
}

// segmented layout: fill each cell's padding (up to the next 4096 boundary) with sentinel
// reads that sort to the end of the cell and are ignored downstream
__global__ void __launch_bounds__(256)
reads_pad_kernel(const ReadCell* __restrict__ cells, int n_cells, const int64_t* __restrict__ cell_tile_start,
                 uint64_t* __restrict__ keys, uint32_t* __restrict__ payload, uint32_t* __restrict__ tile_hist) {
    __shared__ uint32_t s_h[256];
    const int ci = blockIdx.x;
    if (ci >= n_cells) return;
    const ReadCell c = cells[ci];
    const int64_t lo = c.out_start + c.n_reads, hi = cell_tile_start[ci + 1] * READS_SEG_TILE;
    if (lo >= hi) return;
    s_h[threadIdx.x] = 0u;
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint32_t pad_umi = 0xff000000u | (((uint32_t)i * 40503u) & 0xffffffu);
        if (payload) {
            keys[i] = ((uint64_t)c.group << 32) | pad_umi;
            payload[i] = 0u;
        } else {
            keys[i] = pad_umi;
        }
        if (tile_hist) atomicAdd(&s_h[pad_umi & 0xffu], 1u);
    }
    if (!tile_hist) return;
    __syncthreads();
    // the padding lies in the cell's last tile (hi - lo < READS_SEG_TILE)
    if (s_h[threadIdx.x]) atomicAdd(&tile_hist[(cell_tile_start[ci + 1] - 1) * 256 + threadIdx.x], s_h[threadIdx.x]);
}

// ---------------------------------------------------------------- host side --------------
// ReadPlan: everything that depends only on (seed, cells, params) and is computed once --
// the device cell table, the family CSR and the total read count (a pure function of the
// Philox key, so it can be known before the timed step allocates anything).
int32_t reads_plan_build(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families,
                         int64_t n_cells, uint32_t group_base, const pmrd_read_sim_params* sp, int segmented, ReadPlan* plan,
                         int64_t* h_read_offsets /* optional [C+1] */, const CellExtra* h_extra /* optional [C] */, int64_t first_cell) {
    PMRD_ARG(ctx, sp != nullptr && sp->mean_family_size > 0.0 && sp->max_family_size >= 1, "read sim params");
    PMRD_ARG(ctx, n_cells >= 0 && n_cells < (1ll << 31), "n_cells");
    plan->sp = *sp;
    plan->n_cells = n_cells;
    bool any_hop = sp->hop_rate > 0.0;
    if (h_extra) {
        any_hop = false;
        for (int64_t i = 0; i < n_cells; ++i) {
            any_hop |= h_extra[i].hop_rate > 0.0f;
            if (h_extra[i].pool_n > 0)
                PMRD_ARG(ctx, (int64_t)h_extra[i].pool_start >= first_cell && (int64_t)h_extra[i].pool_start + h_extra[i].pool_n <= first_cell + n_cells,
                         "a hop pool must lie inside one chunk of cells");
        }
    }
    plan->sp.hop_rate = any_hop ? (sp->hop_rate > 0.0 ? sp->hop_rate : 1.0) : 0.0;   // (the kernel's switch; the rates are per cell)
    plan->segmented = (segmented && !any_hop) ? 1 : 0;   // hopped reads leave their cell's region
    ReadCell* hc = (ReadCell*)malloc(sizeof(ReadCell) * (size_t)(n_cells + 1));
    int64_t* ho = (int64_t*)malloc(8 * (size_t)(n_cells + 1));
    if (!hc || !ho) { free(hc); free(ho); PMRD_ARG(ctx, false, "out of host memory"); }
    int64_t nf = 0;
    for (int64_t i = 0; i < n_cells; ++i) {
        reads_init_cell(&hc[i], (uint64_t)h_cell_id[i], h_af ? h_af[i] : 0.0, nf, group_base + (uint32_t)i);
        ho[i] = nf;
        nf += h_n_families[i] > 0 ? h_n_families[i] : 0;
    }
    ho[n_cells] = nf;
    plan->n_fam = nf;
    plan->n_reads = 0;
    cudaError_t e = plan->cells.alloc(ctx->stream, sizeof(ReadCell) * (size_t)(n_cells + 1));
    if (e == cudaSuccess) e = plan->fam_offsets.alloc(ctx->stream, 8 * (size_t)(n_cells + 1));
    if (e == cudaSuccess) e = plan->fam_size.alloc(ctx->stream, 4 * (size_t)(nf + 1));
    if (e == cudaSuccess) e = plan->fam_word.alloc(ctx->stream, 4 * (size_t)(nf + 1));
    if (e == cudaSuccess) e = plan->fam_key.alloc(ctx->stream, 4 * (size_t)(nf + 1));
    if (e == cudaSuccess) e = plan->read_offset.alloc(ctx->stream, 8 * (size_t)(nf + 1));
    if (e == cudaSuccess) e = plan->total.alloc(ctx->stream, 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(plan->cells.p, hc, sizeof(ReadCell) * (size_t)n_cells, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(plan->fam_offsets.p, ho, 8 * (size_t)(n_cells + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = plan->cell_tile_start.alloc(ctx->stream, 8 * (size_t)(n_cells + 1));
    CellKnobDev* h_kn = nullptr;
    if (h_extra && n_cells > 0) {          // per-cell knobs ride in their own array (see reads.cuh)
        h_kn = (CellKnobDev*)malloc(sizeof(CellKnobDev) * (size_t)n_cells);
        if (!h_kn) { free(hc); free(ho); PMRD_ARG(ctx, false, "out of host memory"); }
        for (int64_t i = 0; i < n_cells; ++i) reads_fill_knob(&h_kn[i], &hc[i], &h_extra[i], first_cell, n_cells);
        if (e == cudaSuccess) e = plan->knobs.alloc(ctx->stream, sizeof(CellKnobDev) * (size_t)n_cells);
        if (e == cudaSuccess) e = cudaMemcpyAsync(plan->knobs.p, h_kn, sizeof(CellKnobDev) * (size_t)n_cells, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        free(h_kn);
    }
    // cell of the first family of every READS_FAM_BLOCK-family block (+ a closing entry)
    const int64_t n_blk = (nf + READS_FAM_BLOCK - 1) / READS_FAM_BLOCK;
    int32_t* h_blk = (int32_t*)malloc(4 * (size_t)(n_blk + 1));
    if (!h_blk) { free(hc); free(ho); PMRD_ARG(ctx, false, "out of host memory"); }
    {
        int64_t c = 0;
        for (int64_t b = 0; b < n_blk; ++b) {
            const int64_t f = b * READS_FAM_BLOCK;
            while (c + 1 < n_cells && ho[c + 1] <= f) ++c;     // last cell whose first family is <= f
            h_blk[b] = (int32_t)c;
        }
        h_blk[n_blk] = (int32_t)(n_cells > 0 ? n_cells - 1 : 0);
    }
    if (e == cudaSuccess) e = plan->blk_cell.alloc(ctx->stream, 4 * (size_t)(n_blk + 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(plan->blk_cell.p, h_blk, 4 * (size_t)(n_blk + 1), cudaMemcpyHostToDevice, ctx->stream);
    float h_cdf[128];
    uint8_t h_lut[4096];
    plan->has_size_cdf = reads_build_size_tables(sp, h_cdf, h_lut);
    if (plan->has_size_cdf) {
        if (e == cudaSuccess) e = plan->size_cdf.alloc(ctx->stream, sizeof(h_cdf));
        if (e == cudaSuccess) e = cudaMemcpyAsync(plan->size_cdf.p, h_cdf, sizeof(h_cdf), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = plan->size_lut.alloc(ctx->stream, sizeof(h_lut));
        if (e == cudaSuccess) e = cudaMemcpyAsync(plan->size_lut.p, h_lut, sizeof(h_lut), cudaMemcpyHostToDevice, ctx->stream);
    }
    // (no host sync here: copies from pageable memory are staged before cudaMemcpyAsync returns, and everything
    // below is ordered on the stream)
    free(h_blk);
    if (e != cudaSuccess) {
        free(hc); free(ho);
        snprintf(ctx->err, sizeof(ctx->err), "read plan: %s", cudaGetErrorString(e));
        return PMRD_ERR_CUDA;
    }
    int32_t st = reads_plan_sizes(ctx, plan);
    plan->sizes_fresh = st == PMRD_OK;
    plan->sizes_seed = ctx->seed;
    int64_t* h_cnt = (int64_t*)malloc(8 * (size_t)(n_cells + 1));
    int64_t* h_tile = (int64_t*)malloc(8 * (size_t)(n_cells + 1));
    if (!h_cnt || !h_tile) { free(hc); free(ho); free(h_cnt); free(h_tile); PMRD_ARG(ctx, false, "out of host memory"); }
    if (st == PMRD_OK) {
        // exact read count per cell (a pure function of the Philox key), gathered on the device
        // (ONE host sync per build: the total and the per-cell counts come back together)
        uint64_t tot = 0;
        DevBuf cnt;
        e = cnt.alloc(ctx->stream, 8 * (size_t)(n_cells + 1));
        if (e == cudaSuccess && nf == 0) e = cudaMemsetAsync(plan->total.p, 0, 8, ctx->stream);
        if (e == cudaSuccess && n_cells > 0) {
            cell_read_count_kernel<<<pmrd_blocks(n_cells, 256), 256, 0, ctx->stream>>>(
                plan->read_offset.as<uint64_t>(), plan->fam_offsets.as<int64_t>(), n_cells, nf, plan->total.as<uint64_t>(), cnt.as<int64_t>());
            ctx->launches++;
            e = cudaMemcpyAsync(h_cnt, cnt.p, 8 * (size_t)n_cells, cudaMemcpyDeviceToHost, ctx->stream);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&tot, plan->total.p, 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        plan->n_reads = (int64_t)tot;
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "read plan: %s", cudaGetErrorString(e)); st = PMRD_ERR_CUDA; }
    }
    if (st == PMRD_OK && plan->segmented) {
        // every cell far below a tile: the compact layout (no padding) and the one-block-per-cell collapse
        int64_t mx = 0;
        for (int64_t i = 0; i < n_cells; ++i) mx = h_cnt[i] > mx ? h_cnt[i] : mx;
        plan->max_cell_reads = mx;
        const char* sc = getenv("PMRD_SMALL_CELLS");
        if ((!sc || atoi(sc) != 0) && mx <= READS_SEG_TILE && plan->n_reads < (1ll << 32)) plan->segmented = 2;
    }
    if (st == PMRD_OK) {
        int64_t rs = 0, tile = 0;
        for (int64_t i = 0; i < n_cells; ++i) {
            if (h_cnt[i] >= (1ll << 32)) { st = PMRD_ERR_ARG; snprintf(ctx->err, sizeof(ctx->err), "a cell holds >= 2^32 reads"); break; }
            hc[i].read_start = rs;
            hc[i].n_reads = (uint32_t)h_cnt[i];
            plan->n_groups_with_reads += h_cnt[i] > 0;
            h_tile[i] = tile;
            hc[i].out_start = plan->segmented == 1 ? tile * READS_SEG_TILE : rs;
            hc[i].tile0 = (uint32_t)tile;
            hc[i].inv_full = shuffle_inv((uint32_t)h_cnt[i]);
            if (h_read_offsets) h_read_offsets[i] = rs;
            rs += h_cnt[i];
            if (plan->segmented == 1) tile += (h_cnt[i] + READS_SEG_TILE - 1) / READS_SEG_TILE;
        }
        h_tile[n_cells] = tile;
        if (h_read_offsets) h_read_offsets[n_cells] = rs;
        plan->n_out = plan->segmented == 1 ? tile * READS_SEG_TILE : plan->n_reads;
        e = cudaMemcpyAsync(plan->cells.p, hc, sizeof(ReadCell) * (size_t)n_cells, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(plan->cell_tile_start.p, h_tile, 8 * (size_t)(n_cells + 1), cudaMemcpyHostToDevice, ctx->stream);
        // cell of every tile (the consensus kernel would otherwise binary-search it, one thread per block)
        int32_t* h_tc = (int32_t*)malloc(4 * (size_t)(tile + 1));
        if (!h_tc) e = cudaErrorMemoryAllocation;
        else {
            for (int64_t i = 0; i < n_cells; ++i)
                for (int64_t t = h_tile[i]; t < h_tile[i + 1]; ++t) h_tc[t] = (int32_t)i;
            if (e == cudaSuccess) e = plan->tile_cell.alloc(ctx->stream, 4 * (size_t)(tile + 1));
            if (e == cudaSuccess) e = cudaMemcpyAsync(plan->tile_cell.p, h_tc, 4 * (size_t)tile, cudaMemcpyHostToDevice, ctx->stream);
        }
        free(h_tc);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "read plan: %s", cudaGetErrorString(e)); st = PMRD_ERR_CUDA; }
    }
    free(hc); free(ho); free(h_cnt); free(h_tile);
    return st;
}

// family sizes + read offsets (device only, no sync): part of every simulate step
int32_t reads_plan_sizes(pmrd_ctx* ctx, ReadPlan* plan) {
    if (plan->n_fam <= 0) return PMRD_OK;
#define SIZE_ARGS ctx->seed, plan->cells.as<ReadCell>(), plan->fam_offsets.as<int64_t>(), (int)plan->n_cells, plan->n_fam, plan->sp,                       \
                  plan->has_size_cdf ? (const float*)plan->size_cdf.as<float>() : (const float*)nullptr,                                                  \
                  plan->has_size_cdf ? (const uint8_t*)plan->size_lut.as<uint8_t>() : (const uint8_t*)nullptr,                                            \
                  (const int32_t*)plan->blk_cell.as<int32_t>(), plan->fam_size.as<uint32_t>(), plan->fam_word.as<uint32_t>(), plan->fam_key.as<uint32_t>(), \
                  (const CellKnobDev*)plan->knobs.as<CellKnobDev>()
    if (plan->knobs.p) PMRD_LAUNCH_NAMED_X(ctx, "reads_size_kernel", reads_size_kernel<true>, pmrd_blocks(plan->n_fam, READS_FAM_BLOCK), READS_FAM_BLOCK, 0, SIZE_ARGS);
    else PMRD_LAUNCH_NAMED_X(ctx, "reads_size_kernel", reads_size_kernel<false>, pmrd_blocks(plan->n_fam, READS_FAM_BLOCK), READS_FAM_BLOCK, 0, SIZE_ARGS);
#undef SIZE_ARGS
    return exclusive_scan<uint32_t, uint64_t>(ctx, plan->fam_size.as<uint32_t>(), plan->read_offset.as<uint64_t>(), plan->n_fam,
                                              plan->total.as<uint64_t>());
}

// generate the reads (device only, no sync); capacity must be >= plan->n_out
// d_tile_hist (optional, segmented packed layout only): [n_out / 4096][256] histogram of the low UMI byte per
// output tile, produced as a by-product -- the collapse stage then skips its first counting pass
int32_t reads_plan_generate(pmrd_ctx* ctx, ReadPlan* plan, uint64_t* d_keys, uint32_t* d_payload, uint32_t* d_tile_hist) {
    if (plan->n_fam <= 0) return PMRD_OK;
    if (plan->segmented != 1) d_tile_hist = nullptr;
    if (d_tile_hist) PMRD_CUDA(ctx, cudaMemsetAsync(d_tile_hist, 0, (size_t)(plan->n_out / READS_SEG_TILE) * 256 * sizeof(uint32_t), ctx->stream));
#define GEN_ARGS ctx->seed, plan->cells.as<ReadCell>(), plan->fam_offsets.as<int64_t>(), (int)plan->n_cells, plan->n_fam, plan->sp,                       \
                 (const uint64_t*)plan->read_offset.as<uint64_t>(), (const uint32_t*)plan->fam_size.as<uint32_t>(),                                     \
                 (const uint32_t*)plan->fam_word.as<uint32_t>(), (const uint32_t*)plan->fam_key.as<uint32_t>(),                                         \
                 (const int32_t*)plan->blk_cell.as<int32_t>(), plan->n_reads, plan->segmented, d_keys, d_payload, d_tile_hist,                          \
                 (const CellKnobDev*)plan->knobs.as<CellKnobDev>()
    if (plan->knobs.p) PMRD_LAUNCH_NAMED_X(ctx, "reads_gen_kernel", reads_gen_kernel<true>, pmrd_blocks(plan->n_fam, READS_FAM_BLOCK), READS_FAM_BLOCK, 0, GEN_ARGS);
    else PMRD_LAUNCH_NAMED_X(ctx, "reads_gen_kernel", reads_gen_kernel<false>, pmrd_blocks(plan->n_fam, READS_FAM_BLOCK), READS_FAM_BLOCK, 0, GEN_ARGS);
#undef GEN_ARGS
    if (plan->segmented == 1 && plan->n_out > plan->n_reads)
        PMRD_LAUNCH(ctx, reads_pad_kernel, (int)plan->n_cells, 256, 0, plan->cells.as<ReadCell>(), (int)plan->n_cells,
                    (const int64_t*)plan->cell_tile_start.as<int64_t>(), d_keys, d_payload, d_tile_hist);
    return PMRD_OK;
}

// cell-major packed records (as the plan generates them) -> compact (key, payload) arrays in the SAME order:
// cell after cell, each cell's reads in slot order, padding dropped.  One block per cell.
__global__ void __launch_bounds__(256)
reads_unpack_kernel(const ReadCell* __restrict__ cells, int n_cells, const uint64_t* __restrict__ rec, uint64_t* __restrict__ keys,
                    uint32_t* __restrict__ payload) {
    const int ci = blockIdx.x;
    if (ci >= n_cells) return;
    const ReadCell c = cells[ci];
    for (uint32_t i = threadIdx.x; i < c.n_reads; i += blockDim.x) {
        const uint64_t r = rec[c.out_start + i];
        keys[c.read_start + i] = ((uint64_t)c.group << 32) | (uint32_t)r;
        payload[c.read_start + i] = (uint32_t)(r >> 32);
    }
}

// generate in the plan's cell-major layout and hand the reads back compacted, in that order (tests / inspection)
int32_t reads_plan_generate_cell_major(pmrd_ctx* ctx, ReadPlan* plan, uint64_t* d_keys, uint32_t* d_payload) {
    if (plan->n_fam <= 0 || plan->n_reads <= 0) return PMRD_OK;
    PMRD_ARG(ctx, plan->segmented, "cell-major generation needs a segmented plan (hop_rate == 0)");
    DevBuf rec;
    PMRD_CUDA(ctx, rec.alloc(ctx->stream, (size_t)plan->n_out * 8));
    PMRD_TRY(reads_plan_generate(ctx, plan, rec.as<uint64_t>(), nullptr, nullptr));
    PMRD_LAUNCH(ctx, reads_unpack_kernel, (int)plan->n_cells, 256, 0, plan->cells.as<ReadCell>(), (int)plan->n_cells,
                (const uint64_t*)rec.as<uint64_t>(), d_keys, d_payload);
    return PMRD_OK;
}

// per-cell read counts into d_out[n_cells] (not part of the timed step)
int32_t reads_plan_cell_counts(pmrd_ctx* ctx, ReadPlan* plan, int64_t* d_out) {
    if (plan->n_cells <= 0) return PMRD_OK;
    PMRD_LAUNCH(ctx, cell_read_count_kernel, pmrd_blocks(plan->n_cells, 256), 256, 0, (const uint64_t*)plan->read_offset.as<uint64_t>(),
                (const int64_t*)plan->fam_offsets.as<int64_t>(), plan->n_cells, plan->n_fam, (const uint64_t*)plan->total.as<uint64_t>(), d_out);
    return PMRD_OK;
}
