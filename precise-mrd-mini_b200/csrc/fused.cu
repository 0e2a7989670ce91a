// fused.cu -- the grid path that never materialises a read (SURVEY.md §7.1 step 7, §8(d) "fused grid path").
//
// One block owns one (AF, depth, replicate) cell and goes from Philox counters straight to the cell's three
// counts (families formed, consensus families, variant consensus families): every read is drawn, voted and
// forgotten in registers / shared memory.  It replaces, per cell, the whole chain
//   SyntheticReadGenerator.generate_reads_for_site   src/precise_mrd/io.py:79-123
//   UMIProcessor.group_umis_by_site / exact UMI dict   src/precise_mrd/umi.py:97-115
//   UMIProcessor.call_consensus / filter_family_size   src/precise_mrd/umi.py:146-203
//   UMIProcessor.get_consensus_counts                  src/precise_mrd/umi.py:241-266
// i.e. the body of the triple loops at src/precise_mrd/eval/lod.py:125-150 and sim/contamination.py:85-228.
//
// Results are those of the materialised path on the family-major read order (pmrd_simulate_reads, shuffle = 0):
// the reads come from the same Philox blocks (reads_dev.cuh) and a family's vote is the same sequential fp64
// accumulation.  What a sort would have found -- two molecules of a cell that drew the SAME 24-bit UMI (about
// n^2 / 2^25 pairs in a cell of n molecules) form ONE family whose reads are voted together -- is found without
// one:
//   pass 1  every molecule's UMI goes through two independent hash bitmaps in shared memory ("seen" / "seen twice");
//   pass 2  a molecule whose two slots were both hit twice is a CANDIDATE duplicate: it is not voted here, its
//           (cell, umi, family) triple is appended to a small list;
//   replay  the list is sorted by (cell, umi); each run of equal keys -- a true duplicate group, or a false
//           positive of the filter on its own -- is voted by one thread that re-draws its reads from the Philox
//           counters in family order.  True duplicates always share both slots, so none is missed.
// Bound: INT32 issue (Philox): no HBM traffic beyond the candidate list (12 B per candidate, a few % of the
// families) and the per-cell counts.
#include "reads_dev.cuh"
#include "radix_sort.cuh"
#include "fused.cuh"
#include "vote.cuh"
#include <new>
#include <stdlib.h>

static __constant__ double c_wtab_f[64];   // 10^(q/10), q = 0..63 (host libm pow, as CPython computes it)

int32_t fused_upload_tables(pmrd_ctx* ctx, const double* h_w64) {
    PMRD_CUDA(ctx, cudaMemcpyToSymbolAsync(c_wtab_f, h_w64, 64 * sizeof(double), 0, cudaMemcpyHostToDevice, ctx->stream));
    return reads_tables_upload_here(ctx);
}

// ---------------------------------------------------------------- votes -----------------
// The walk of umi.py:152-174 over stashed reads (byte = allele[1:0] | quality << 2).  The four per-allele sums
// live in a shared-memory column of the owning thread (acc[a * STRIDE]): the allele picks the address.
// MAJ (consensus_threshold > 0.5): the winner must hold a strict majority of the weight, so two alleles tied for
// the best weight can never pass and the reference's first-seen tie-break cannot matter: accumulate only.
template <int STRIDE, bool MAJ>
struct FoldVote {
    double* acc;
    double total = 0.0;
    uint32_t order = 0, seen = 0, n_seen = 0;
    __device__ __forceinline__ explicit FoldVote(double* column) : acc(column) {
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a * STRIDE] = 0.0;
    }
    __device__ __forceinline__ void add(uint32_t byte, const double* wt) {
        const uint32_t a = byte & 3u;
        const double w = wt[(byte >> 2) & 63u];              // 0.0 below the quality threshold: adds nothing, sets no bit
        total = __dadd_rn(total, w);
        double* p = acc + a * STRIDE;
        *p = __dadd_rn(*p, w);
        if (!MAJ) {
            const uint32_t bit = (__double2hiint(w) != 0 ? 1u : 0u) << a;
            const uint32_t fresh = bit & ~seen;
            order |= fresh ? n_seen << (2 * a) : 0u;
            n_seen += fresh ? 1u : 0u;
            seen |= bit;
        }
    }
    __device__ __forceinline__ int result(double cthr) const {
        if (total == 0.0) return -1;                         // no read passed the quality threshold
        int best = -1;
        double bw = -1.0;
        uint32_t brank = 4;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double wa = acc[a * STRIDE];
            if (MAJ) {
                if (wa > bw) { bw = wa; best = a; }
            } else {
                if (!((seen >> a) & 1u)) continue;
                const uint32_t rk = (order >> (2 * a)) & 3u;
                if (wa > bw || (wa == bw && rk < brank)) { bw = wa; best = a; brank = rk; }
            }
        }
        return (__ddiv_rn(bw, total) < cthr) ? -1 : best;
    }
};

// ---------------------------------------------------------------- the cell kernel --------
#define FUSED_STASH 256     /* reads stashed per warp between two folds */
struct __align__(16) FusedSlot {
    uint32_t loc;      // first read of the family in the warp's flattened read range
    uint32_t key;      // Philox2x32 key of the family's read stream
    uint32_t word;     // umi[23:0] | has_variant[24] | ref[27:26] | alt[29:28]
    uint32_t fam;      // family index inside the cell (counter word 1)
};

// FUSED_K hash functions into ONE pair of bitmaps ("seen", "seen twice") of 2^(filt_log2 + 1) slots: a Bloom filter.  A
// molecule is a candidate duplicate when ALL its slots were hit twice; two molecules with the same UMI share all of them.
// With 8+ slots per molecule and k = 4 the false-positive rate of the largest cells is ~1 % (two separate bitmaps with one
// hash each, the first version: 3 %) for the same shared memory.
#ifndef FUSED_K
#define FUSED_K 4
#endif
__device__ __forceinline__ uint32_t fused_hash(uint32_t umi, int j, int log2) {
    const uint32_t mul[6] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu, 0x165667B1u, 0xD3A2646Du};
    return (umi * mul[j]) >> (32 - log2);          // (the UMI's 24 bits, spread by an odd multiplier: the top bits)
}

// MODE 0: count the candidates of every cell (plan build: sizes the candidate list; nothing else is written)
// MODE 1: the step
template <int THREADS, bool MAJ, int MODE, bool KNOBS>
__global__ void __launch_bounds__(THREADS)
fused_cell_kernel(uint64_t seed, const ReadCell* __restrict__ cells, int n_cells, pmrd_read_sim_params sp, pmrd_umi_params P,
                  const float* __restrict__ size_cdf, const uint8_t* __restrict__ size_lut, int filt_log2,
                  uint64_t* __restrict__ cand_key, uint32_t* __restrict__ cand_fam, unsigned long long* __restrict__ cand_cursor,
                  int64_t cell0, const uint8_t* __restrict__ alt_allele, unsigned long long* __restrict__ o_reads,
                  unsigned long long* __restrict__ o_fam, unsigned long long* __restrict__ o_tot, unsigned long long* __restrict__ o_var,
                  unsigned long long* __restrict__ o_cand, const CellKnobDev* __restrict__ knobs /* [n_cells] or NULL */) {
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char fz_raw[];
    // layout: [4 bitmaps][allele sums 4 x THREADS doubles][weights][slots][stash][tables]
    const int slot_log2 = filt_log2 + 1;                              // slots of the seen / dup bitmaps
    const uint32_t words = (1u << slot_log2) / 32u;
    uint32_t* const s_seen = reinterpret_cast<uint32_t*>(fz_raw);
    uint32_t* const s_dup = s_seen + words;
    double* const s_sum = reinterpret_cast<double*>(s_dup + words);
    double* const s_wt = s_sum + 4 * THREADS;
    FusedSlot* const s_slot = reinterpret_cast<FusedSlot*>(s_wt + 64);
    uint8_t* const s_stash = reinterpret_cast<uint8_t*>(s_slot + THREADS);
    uint8_t* const s_qlut = s_stash + WARPS * FUSED_STASH;
    uint8_t* const s_lut = s_qlut + (1 << QUAL_LUT_BITS);
    float* const s_cdf = reinterpret_cast<float*>(s_lut + 4096);
    uint32_t* const s_qthr = reinterpret_cast<uint32_t*>(s_cdf + 128);
    unsigned long long* const s_acc = reinterpret_cast<unsigned long long*>(s_qthr + 32);   // [5]

    const int cell = blockIdx.x;
    if (cell >= n_cells) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = lanemask_lt();
    const ReadCell c = cells[cell];
    const uint32_t n_fam = (uint32_t)(cells[cell + 1].fam_offset - c.fam_offset);
    const bool have_cdf = size_cdf != nullptr;

    for (uint32_t i = threadIdx.x; i < 2u * words; i += THREADS) s_seen[i] = 0u;
    if (threadIdx.x < 64) s_wt[threadIdx.x] = (int)threadIdx.x >= P.quality_threshold ? c_wtab_f[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) s_qthr[threadIdx.x] = c_qual_thr[threadIdx.x];
    if (threadIdx.x < 5) s_acc[threadIdx.x] = 0ull;
    for (uint32_t i = threadIdx.x; i < (1u << QUAL_LUT_BITS) / 16u; i += THREADS)
        reinterpret_cast<uint4*>(s_qlut)[i] = reinterpret_cast<const uint4*>(c_qual_lut)[i];
    if (have_cdf) {
        for (uint32_t i = threadIdx.x; i < 4096u / 16u; i += THREADS) reinterpret_cast<uint4*>(s_lut)[i] = reinterpret_cast<const uint4*>(size_lut)[i];
        if (threadIdx.x < 128) s_cdf[threadIdx.x] = size_cdf[threadIdx.x];
    }
    __syncthreads();
    const Philox ph(seed);
    const float* const cdf = have_cdf ? s_cdf : nullptr;
    const uint8_t* const lut = have_cdf ? s_lut : nullptr;
    const CellKnobDev* const kn = KNOBS ? knobs + cell : nullptr;

    // ---- pass 1: every molecule's UMI into the two filters --------------------------------
    for (uint32_t f = threadIdx.x; f < n_fam; f += THREADS) {
        const FamDraw d = draw_family<KNOBS>(ph, c, f, sp, cdf, lut, kn);
#pragma unroll
        for (int j = 0; j < FUSED_K; ++j) {
            const uint32_t a = fused_hash(d.umi, j, slot_log2);
            const uint32_t ba = 1u << (a & 31u);
            if (atomicOr(&s_seen[a >> 5], ba) & ba) atomicOr(&s_dup[a >> 5], ba);
        }
    }
    __syncthreads();

    // ---- pass 2: 32 molecules per warp and round, their reads flattened over the lanes ------
    uint32_t n_reads = 0, n_famc = 0, n_tot = 0, n_var = 0, n_cand = 0;
    const uint32_t alt = alt_allele ? alt_allele[cell0 + cell] : c.alt_allele;
    const uint32_t ctr_hi = (uint32_t)c.cell_id << 16;
    uint8_t* const stash = s_stash + warp * FUSED_STASH;
    FusedSlot* const slots = s_slot + warp * 32;
    for (uint32_t f0 = warp * 32u; f0 < n_fam; f0 += THREADS) {
        const uint32_t f = f0 + lane;
        const bool valid = f < n_fam;
        FamDraw d;
        d.size = 0; d.umi = 0; d.key = 0; d.has_variant = false;
        bool cand = false;
        if (valid) {
            d = draw_family<KNOBS>(ph, c, f, sp, cdf, lut, kn);
            cand = true;
#pragma unroll
            for (int j = 0; j < FUSED_K; ++j) {
                const uint32_t a = fused_hash(d.umi, j, slot_log2);
                cand = cand && ((s_dup[a >> 5] >> (a & 31u)) & 1u);
            }
            n_reads += (uint32_t)d.size;
            n_cand += cand ? 1u : 0u;
        }
        const uint32_t cmask = __ballot_sync(0xffffffffu, cand);
        if (MODE == 0) continue;
        if (cmask) {                                          // append the candidates (one cursor bump per warp)
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cand_cursor, (unsigned long long)__popc(cmask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (cand) {
                const unsigned long long pos = base + __popc(cmask & lt);
                cand_key[pos] = ((uint64_t)(uint32_t)cell << 32) | d.umi;
                cand_fam[pos] = f;
            }
        }
        const bool votes = valid && !cand;
        const uint32_t size = votes ? (uint32_t)d.size : 0u;
        // exclusive prefix of the sizes over the lanes; the molecules that vote are compacted into the slot list
        uint32_t incl = size;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t loc = votes ? incl - size : 0xffffffffu;      // (every molecule has >= 1 read: locs are distinct)
        const uint32_t vmask = __ballot_sync(0xffffffffu, votes);
        __syncwarp();                                                 // the previous round's slots / stash are no longer read
        if (votes) {
            FusedSlot& S = slots[__popc(vmask & lt)];
            S.loc = loc; S.key = d.key; S.fam = f;
            S.word = d.umi | ((d.has_variant ? 1u : 0u) << 24) | (c.ref_allele << 26) | (c.alt_allele << 28);
        }
        __syncwarp();
        // (measured on the C2 grid: shared-memory columns 17.1 ms, two register sums with predicated DADDs 18.3 ms, one DADD
        // on a selected sum 18.5 ms -- here the fold runs on the few lanes that own a family and the FP64 pipe is the scarce one)
        FoldVote<THREADS, MAJ> v(s_sum + threadIdx.x);
        const uint32_t le = lt | (1u << lane);
        for (uint32_t w0 = 0; w0 < total; w0 += FUSED_STASH) {
            const uint32_t w1 = min(total, w0 + FUSED_STASH);
            for (uint32_t k0 = w0; k0 < w1; k0 += 32) {
                const uint32_t k = k0 + lane;
                // owner of read k = (number of voting molecules whose first read is <= k) - 1  (reads.cu: same scheme)
                const uint32_t rel = loc - k0;
                const uint32_t before = __popc(__ballot_sync(0xffffffffu, loc < k0));
                const uint32_t heads = __reduce_or_sync(0xffffffffu, rel < 32u ? 1u << rel : 0u);
                if (k < w1) {
                    const uint4 s = *reinterpret_cast<const uint4*>(&slots[before + __popc(heads & le) - 1u]);   // loc, key, word, fam
                    const uint32_t pl = draw_read_payload(s.y, ctr_hi | (k - s.x), s.w, s.z, s_qlut, s_qthr);
                    stash[k - w0] = (uint8_t)pl;                      // allele[1:0] | quality << 2: all the vote needs
                }
            }
            __syncwarp();
            // every voting lane folds the reads of ITS molecule that lie in this window, in read order (umi.py:161-164)
            if (votes) {
                const uint32_t a = max(loc, w0), b = min(loc + size, w1);
                for (uint32_t k = a; k < b; ++k) v.add(stash[k - w0], s_wt);
            }
            __syncwarp();
        }
        if (votes) {
            ++n_famc;
            if ((int)size >= P.min_family_size) {
                const int best = v.result(P.consensus_threshold);
                if (best >= 0) { ++n_tot; n_var += (uint32_t)best == alt ? 1u : 0u; }
            }
        }
    }
    n_reads = warp_sum(n_reads); n_cand = warp_sum(n_cand);
    if (MODE == 1) { n_famc = warp_sum(n_famc); n_tot = warp_sum(n_tot); n_var = warp_sum(n_var); }
    if (lane == 0) {
        atomicAdd(&s_acc[0], (unsigned long long)n_reads);
        atomicAdd(&s_acc[4], (unsigned long long)n_cand);
        if (MODE == 1) {
            atomicAdd(&s_acc[1], (unsigned long long)n_famc);
            atomicAdd(&s_acc[2], (unsigned long long)n_tot);
            atomicAdd(&s_acc[3], (unsigned long long)n_var);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (o_reads) o_reads[cell0 + cell] = s_acc[0];
        if (o_cand) o_cand[cell] = s_acc[4];
        if (MODE == 1) {
            // the replay kernel ADDS the candidate groups of this cell: it runs after this kernel on the same stream
            o_fam[cell0 + cell] = s_acc[1];
            o_tot[cell0 + cell] = s_acc[2];
            o_var[cell0 + cell] = s_acc[3];
        }
    }
}

static size_t fused_smem_bytes(int threads, int filt_log2) {
    return (size_t)4 * ((1u << filt_log2) / 8u) + (size_t)4 * threads * 8 + 64 * 8 + (size_t)threads * sizeof(FusedSlot) +
           (size_t)(threads / 32) * FUSED_STASH + (1 << QUAL_LUT_BITS) + 4096 + 128 * 4 + 32 * 4 + 5 * 8 + 64;
}

// ---------------------------------------------------------------- replay -----------------
// Sorted candidate list: a run of equal (cell, umi) keys is one family.  The thread at the head of a run re-draws
// the reads of its molecules in family order and votes them as one family.
template <bool MAJ, bool KNOBS>
__global__ void __launch_bounds__(128)
fused_replay_kernel(uint64_t seed, const ReadCell* __restrict__ cells, pmrd_read_sim_params sp, pmrd_umi_params P,
                    const float* __restrict__ size_cdf, const uint8_t* __restrict__ size_lut, const uint64_t* __restrict__ key,
                    const uint32_t* __restrict__ fam, int64_t n, int64_t cell0, const uint8_t* __restrict__ alt_allele,
                    unsigned long long* __restrict__ o_fam, unsigned long long* __restrict__ o_tot, unsigned long long* __restrict__ o_var,
                    const CellKnobDev* __restrict__ knobs) {
    __shared__ double s_sum[4 * 128];
    __shared__ double s_wt[64];
    __shared__ float s_cdf[128];
    __shared__ __align__(16) uint8_t s_lut[4096];
    __shared__ __align__(16) uint8_t s_qlut[1 << QUAL_LUT_BITS];
    __shared__ uint32_t s_qthr[32];
    const bool have_cdf = size_cdf != nullptr;
    if (threadIdx.x < 64) s_wt[threadIdx.x] = (int)threadIdx.x >= P.quality_threshold ? c_wtab_f[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) s_qthr[threadIdx.x] = c_qual_thr[threadIdx.x];
    for (uint32_t i = threadIdx.x; i < (1u << QUAL_LUT_BITS) / 16u; i += 128) reinterpret_cast<uint4*>(s_qlut)[i] = reinterpret_cast<const uint4*>(c_qual_lut)[i];
    if (have_cdf) {
        for (uint32_t i = threadIdx.x; i < 4096u / 16u; i += 128) reinterpret_cast<uint4*>(s_lut)[i] = reinterpret_cast<const uint4*>(size_lut)[i];
        s_cdf[threadIdx.x] = size_cdf[threadIdx.x];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = key[i];
    if (i > 0 && key[i - 1] == k) return;                   // not the head of its run
    int64_t hi = i + 1;
    while (hi < n && key[hi] == k) ++hi;
    const int cell = (int)(k >> 32);
    const ReadCell c = cells[cell];
    const Philox ph(seed);
    const uint32_t ctr_hi = (uint32_t)c.cell_id << 16;
    FoldVote<128, MAJ> v(s_sum + threadIdx.x);
    uint32_t size_all = 0;
    // the run's molecules in family order (the sort is stable on the key only; runs hold two or three members)
    int64_t last = -1;
    for (int64_t m = i; m < hi; ++m) {
        uint32_t f = 0xffffffffu;
        for (int64_t j = i; j < hi; ++j) {
            const uint32_t fj = fam[j];
            if ((int64_t)fj > last && fj < f) f = fj;
        }
        last = (int64_t)f;
        const FamDraw d = draw_family<KNOBS>(ph, c, f, sp, have_cdf ? s_cdf : nullptr, have_cdf ? s_lut : nullptr, KNOBS ? knobs + cell : nullptr);
        const uint32_t word = d.umi | ((d.has_variant ? 1u : 0u) << 24) | (c.ref_allele << 26) | (c.alt_allele << 28);
        for (uint32_t r = 0; r < (uint32_t)d.size; ++r) v.add(draw_read_payload(d.key, ctr_hi | r, f, word, s_qlut, s_qthr) & 0xffu, s_wt);
        size_all += (uint32_t)d.size;
    }
    const uint32_t alt = alt_allele ? alt_allele[cell0 + cell] : c.alt_allele;
    atomicAdd(&o_fam[cell0 + cell], 1ull);
    if ((int)size_all >= P.min_family_size) {
        const int best = v.result(P.consensus_threshold);
        if (best >= 0) {
            atomicAdd(&o_tot[cell0 + cell], 1ull);
            if ((uint32_t)best == alt) atomicAdd(&o_var[cell0 + cell], 1ull);
        }
    }
}

// ---------------------------------------------------------------- host side --------------
template <int THREADS, bool MAJ, int MODE, bool KNOBS>
static int32_t fused_launch(pmrd_ctx* ctx, FusedChunk* ch, const pmrd_umi_params* up, uint64_t* ck, uint32_t* cf, unsigned long long* cursor,
                            const uint8_t* d_alt, int64_t* o_reads, int64_t* o_fam, int64_t* o_tot, int64_t* o_var, unsigned long long* o_cand) {
    const size_t smem = fused_smem_bytes(THREADS, ch->filt_log2);
    PMRD_CUDA(ctx, cudaFuncSetAttribute(fused_cell_kernel<THREADS, MAJ, MODE, KNOBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PMRD_LAUNCH_NAMED(ctx, MODE ? "fused_cell_kernel" : "fused_count_kernel", (fused_cell_kernel<THREADS, MAJ, MODE, KNOBS>), (int)ch->n_cells, THREADS, smem,
                      ctx->seed, (const ReadCell*)ch->cells.as<ReadCell>(), (int)ch->n_cells, ch->sp, *up,
                      ch->has_size_cdf ? (const float*)ch->size_cdf.as<float>() : (const float*)nullptr,
                      ch->has_size_cdf ? (const uint8_t*)ch->size_lut.as<uint8_t>() : (const uint8_t*)nullptr, ch->filt_log2, ck, cf, cursor,
                      ch->cell0, d_alt, (unsigned long long*)o_reads, (unsigned long long*)o_fam, (unsigned long long*)o_tot,
                      (unsigned long long*)o_var, o_cand, ch->knobs.p ? (const CellKnobDev*)ch->knobs.as<CellKnobDev>() : (const CellKnobDev*)nullptr);
    return PMRD_OK;
}

template <int MODE>
static int32_t fused_dispatch(pmrd_ctx* ctx, FusedChunk* ch, const pmrd_umi_params* up, uint64_t* ck, uint32_t* cf, unsigned long long* cursor,
                              const uint8_t* d_alt, int64_t* o_reads, int64_t* o_fam, int64_t* o_tot, int64_t* o_var, unsigned long long* o_cand) {
    const bool maj = up->consensus_threshold > 0.5;
    const bool kn = ch->knobs.p != nullptr;
#define FUSED_ARGS ctx, ch, up, ck, cf, cursor, d_alt, o_reads, o_fam, o_tot, o_var, o_cand
#define FUSED_GO(T)                                                                                                            \
    if (kn) return maj ? fused_launch<T, true, MODE, true>(FUSED_ARGS) : fused_launch<T, false, MODE, true>(FUSED_ARGS);         \
    return maj ? fused_launch<T, true, MODE, false>(FUSED_ARGS) : fused_launch<T, false, MODE, false>(FUSED_ARGS)
    if (ch->threads == 1024) { FUSED_GO(1024); }
    if (ch->threads == 256) { FUSED_GO(256); }
    FUSED_GO(128);
#undef FUSED_ARGS
#undef FUSED_GO
}

// cells + tables of one chunk on the device; counts its candidates (one sync) so that the step's list is sized exactly
int32_t fused_chunk_build(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families, int64_t n_cells,
                          int64_t cell0, const pmrd_read_sim_params* sp, const pmrd_umi_params* up, FusedChunk* ch, const CellExtra* h_extra) {
    ch->sp = *sp;
    ch->n_cells = n_cells;
    ch->cell0 = cell0;
    ReadCell* hc = (ReadCell*)malloc(sizeof(ReadCell) * (size_t)(n_cells + 1));
    PMRD_ARG(ctx, hc != nullptr, "out of host memory");
    int64_t nf = 0, max_f = 0;
    for (int64_t i = 0; i < n_cells; ++i) {
        reads_init_cell(&hc[i], (uint64_t)h_cell_id[i], h_af ? h_af[i] : 0.0, nf, (uint32_t)i);
        if (h_extra && h_extra[i].hop_rate > 0.0f) { free(hc); PMRD_ARG(ctx, false, "the fused engine keeps a read inside its cell: hop_rate > 0 needs the materialised engine"); }
        const int64_t f = h_n_families[i] > 0 ? h_n_families[i] : 0;
        nf += f;
        if (f > max_f) max_f = f;
    }
    memset(&hc[n_cells], 0, sizeof(ReadCell));
    hc[n_cells].fam_offset = nf;
    ch->n_fam = nf;
    if (max_f >= (1ll << 32)) { free(hc); PMRD_ARG(ctx, false, "a cell holds >= 2^32 families"); }
    ch->threads = max_f > 4096 ? 1024 : (max_f > 512 ? 256 : 128);
    int lg = 10;
    while (lg < 18 && (1ll << lg) < 8 * max_f) ++lg;          // >= 8 slots per molecule, at most 2^18 per filter (4 x 32 KB)
    ch->filt_log2 = lg;
    float h_cdf[128];
    uint8_t h_lut[4096];
    ch->has_size_cdf = reads_build_size_tables(sp, h_cdf, h_lut);
    cudaError_t e = ch->cells.alloc(ctx->stream, sizeof(ReadCell) * (size_t)(n_cells + 1));
    CellKnobDev* h_kn = nullptr;
    if (h_extra && n_cells > 0) {
        h_kn = (CellKnobDev*)malloc(sizeof(CellKnobDev) * (size_t)n_cells);
        if (!h_kn) { free(hc); PMRD_ARG(ctx, false, "out of host memory"); }
        for (int64_t i = 0; i < n_cells; ++i) reads_fill_knob(&h_kn[i], &hc[i], &h_extra[i], cell0, n_cells);
        if (e == cudaSuccess) e = ch->knobs.alloc(ctx->stream, sizeof(CellKnobDev) * (size_t)n_cells);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ch->knobs.p, h_kn, sizeof(CellKnobDev) * (size_t)n_cells, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ch->cells.p, hc, sizeof(ReadCell) * (size_t)(n_cells + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && ch->has_size_cdf) {
        e = ch->size_cdf.alloc(ctx->stream, sizeof(h_cdf));
        if (e == cudaSuccess) e = cudaMemcpyAsync(ch->size_cdf.p, h_cdf, sizeof(h_cdf), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = ch->size_lut.alloc(ctx->stream, sizeof(h_lut));
        if (e == cudaSuccess) e = cudaMemcpyAsync(ch->size_lut.p, h_lut, sizeof(h_lut), cudaMemcpyHostToDevice, ctx->stream);
    }
    DevBuf cnt, reads;
    if (e == cudaSuccess) e = cnt.alloc(ctx->stream, 8 * (size_t)(n_cells + 1));
    if (e == cudaSuccess) e = reads.alloc(ctx->stream, 8 * (size_t)(n_cells + 1));
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);      // hc / tables were staged from the stack and the heap
    free(hc);
    free(h_kn);
    if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "fused plan: %s", cudaGetErrorString(e)); return PMRD_ERR_CUDA; }
    ch->n_cand = 0;
    ch->n_reads = 0;
    if (n_cells > 0 && nf > 0) {
        // (o_reads is indexed cell0 + cell in the kernel: hand it a pointer shifted back by cell0)
        PMRD_TRY(fused_dispatch<0>(ctx, ch, up, nullptr, nullptr, nullptr, nullptr, reads.as<int64_t>() - cell0, nullptr, nullptr, nullptr,
                                   cnt.as<unsigned long long>()));
        int64_t* h = (int64_t*)malloc(16 * (size_t)n_cells);
        PMRD_ARG(ctx, h != nullptr, "out of host memory");
        e = cudaMemcpyAsync(h, cnt.p, 8 * (size_t)n_cells, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h + n_cells, reads.p, 8 * (size_t)n_cells, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess)
            for (int64_t i = 0; i < n_cells; ++i) { ch->n_cand += h[i]; ch->n_reads += h[n_cells + i]; ch->n_cells_with_reads += h[n_cells + i] > 0; }
        free(h);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "fused plan: %s", cudaGetErrorString(e)); return PMRD_ERR_CUDA; }
    }
    return PMRD_OK;
}

size_t fused_scratch_bytes(int64_t n_cand) {
    const size_t n = (size_t)(n_cand > 0 ? n_cand : 1);
    return 2 * ((n * 8 + 255) & ~(size_t)255) + 2 * ((n * 4 + 255) & ~(size_t)255) + 256 + radix_scratch_bytes((int64_t)n);
}

// one step of a chunk: cell kernel -> sort the candidate list -> replay.  Only enqueues kernels.
int32_t fused_chunk_run(pmrd_ctx* ctx, FusedChunk* ch, const pmrd_umi_params* up, void* d_scratch, const uint8_t* d_alt,
                        int64_t* o_reads, int64_t* o_fam, int64_t* o_tot, int64_t* o_var) {
    if (ch->n_cells <= 0 || ch->n_fam <= 0) return PMRD_OK;
    const size_t n = (size_t)(ch->n_cand > 0 ? ch->n_cand : 1);
    const size_t a = (n * 8 + 255) & ~(size_t)255, b = (n * 4 + 255) & ~(size_t)255;
    unsigned char* w = (unsigned char*)d_scratch;
    uint64_t* ka = (uint64_t*)w; w += a;
    uint64_t* kb = (uint64_t*)w; w += a;
    uint32_t* va = (uint32_t*)w; w += b;
    uint32_t* vb = (uint32_t*)w; w += b;
    unsigned long long* cursor = (unsigned long long*)w; w += 256;
    uint32_t* rs = (uint32_t*)w;
    PMRD_CUDA(ctx, cudaMemsetAsync(cursor, 0, 8, ctx->stream));
    PMRD_TRY(fused_dispatch<1>(ctx, ch, up, ka, va, cursor, d_alt, o_reads, o_fam, o_tot, o_var, nullptr));
    if (ch->n_cand <= 0) return PMRD_OK;
    // (cell, umi) keys: 24 UMI bits + the bits of the chunk's cell index -- known on the host, no reduction needed
    uint64_t varying = 0xffffffull;
    int gb = 0;
    while ((1ll << gb) < ch->n_cells) ++gb;
    if (gb > 0) varying |= ((1ull << gb) - 1ull) << 32;
    int in_b = 0;
    PMRD_TRY(radix_sort_pairs(ctx, ka, va, kb, vb, ch->n_cand, varying, &in_b, rs));
    const uint64_t* sk = in_b ? kb : ka;
    const uint32_t* sf = in_b ? vb : va;
    const float* cdf = ch->has_size_cdf ? (const float*)ch->size_cdf.as<float>() : (const float*)nullptr;
    const uint8_t* lut = ch->has_size_cdf ? (const uint8_t*)ch->size_lut.as<uint8_t>() : (const uint8_t*)nullptr;
    const CellKnobDev* kn = (const CellKnobDev*)ch->knobs.as<CellKnobDev>();
#define REPLAY_ARGS pmrd_blocks(ch->n_cand, 128), 128, 0, ctx->seed, (const ReadCell*)ch->cells.as<ReadCell>(), ch->sp, *up, cdf, lut, sk, sf, ch->n_cand, \
                    ch->cell0, d_alt, (unsigned long long*)o_fam, (unsigned long long*)o_tot, (unsigned long long*)o_var, kn
    const bool maj = up->consensus_threshold > 0.5;
    if (kn && maj) PMRD_LAUNCH_NAMED_X(ctx, "fused_replay_kernel", (fused_replay_kernel<true, true>), REPLAY_ARGS);
    else if (kn) PMRD_LAUNCH_NAMED_X(ctx, "fused_replay_kernel", (fused_replay_kernel<false, true>), REPLAY_ARGS);
    else if (maj) PMRD_LAUNCH_NAMED_X(ctx, "fused_replay_kernel", (fused_replay_kernel<true, false>), REPLAY_ARGS);
    else PMRD_LAUNCH_NAMED_X(ctx, "fused_replay_kernel", (fused_replay_kernel<false, false>), REPLAY_ARGS);
#undef REPLAY_ARGS
    return PMRD_OK;
}
