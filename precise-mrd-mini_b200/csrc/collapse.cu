// collapse.cu -- K4..K7 for the read-level path: group reads into UMI families, vote a
// consensus per family, summarise per (sample, locus) group.
//
//   grouping   : stable radix sort of the packed key (radix_sort.cu); EXACT UMI match
//                (src/precise_mrd/umi.py:111-115, rust_ext/src/lib.rs:7-22) or first-fit
//                Hamming clustering to the representative (precise-mrd-mini/src/mrd/umi.py:29-40)
//   consensus  : WEIGHTED  = quality-weighted vote w = 10^(q/10)       (umi.py:146-179, lib.rs:25-74)
//                MAJORITY  = per-column majority, first-seen wins ties  (mrd/umi.py:43-54)
//                COUNT     = REF/ALT read counts per UMI                (G-C fixture)
//   filter     : min/max family size on the RAW read count, after consensus (umi.py:181-203)
//   group table: kept families, consensus counts per allele (umi.py:241-266), alt/total reads
//                (mrd/umi.py:86-99)
//
// Bit-exactness: the reference accumulates fp64 weights sequentially in read order
// (umi.py:161-164); the sort is stable, one thread walks one family in that order with
// the same IEEE additions, and the weight table is computed on the host with the same
// libm pow() Python uses, so the >= consensus_threshold decision is identical.
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "reads.cuh"
#include "vote.cuh"

__constant__ double c_wtab[64];  // 10^(q/10), q = 0..63

int32_t collapse_upload_weights(pmrd_ctx* ctx, const double* h_w64) {
    PMRD_CUDA(ctx, cudaMemcpyToSymbolAsync(c_wtab, h_w64, 64 * sizeof(double), 0, cudaMemcpyHostToDevice, ctx->stream));
    return PMRD_OK;
}

// ---------------------------------------------------------------- segments --------------
// Segment starts of a sorted key array, with NO host round trip: the element count may
// live on the device (d_n) and the number of segments found stays there (d_count), so the
// whole collapse can be enqueued -- or captured in a CUDA graph -- without a sync.
#define SEG_THREADS 256
#define SEG_ITEMS 8
#define SEG_TILE (SEG_THREADS * SEG_ITEMS)

__device__ __forceinline__ bool seg_is_head(const uint64_t* __restrict__ keys, int64_t i, int shift) {
    return i == 0 || (keys[i] >> shift) != (keys[i - 1] >> shift);
}

__global__ void __launch_bounds__(SEG_THREADS)
seg_count_kernel(const uint64_t* __restrict__ keys, int64_t n_cap, const uint32_t* __restrict__ d_n, int shift,
                 uint32_t* __restrict__ tile_heads) {
    __shared__ uint32_t sw[SEG_THREADS / 32];
    const int64_t n = d_n ? (int64_t)*d_n : n_cap;
    const int64_t base = (int64_t)blockIdx.x * SEG_TILE;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; ++j) {
        const int64_t i = base + j * SEG_THREADS + threadIdx.x;
        if (i < n) c += seg_is_head(keys, i, shift) ? 1u : 0u;
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SEG_THREADS / 32; ++w) t += sw[w];
        tile_heads[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(SEG_THREADS)
seg_write_kernel(const uint64_t* __restrict__ keys, int64_t n_cap, const uint32_t* __restrict__ d_n, int shift,
                 const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ total, int64_t seg_cap,
                 uint32_t* __restrict__ seg_start, uint32_t* __restrict__ d_count) {
    __shared__ uint32_t sw[33];
    const int64_t n = d_n ? (int64_t)*d_n : n_cap;
    // blocked arrangement so that ranks follow memory order
    const int64_t base = (int64_t)blockIdx.x * SEG_TILE + (int64_t)threadIdx.x * SEG_ITEMS;
    uint32_t flags = 0, c = 0;
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; ++j) {
        const int64_t i = base + j;
        if (i < n && seg_is_head(keys, i, shift)) {
            flags |= 1u << j;
            ++c;
        }
    }
    uint32_t tot;
    uint32_t r = tile_base[blockIdx.x] + block_exclusive_scan<uint32_t>(c, sw, tot);
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; ++j)
        if (flags & (1u << j)) {
            if ((int64_t)r < seg_cap) seg_start[r] = (uint32_t)(base + j);
            ++r;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t t = *total;
        if ((int64_t)t <= seg_cap) seg_start[t] = (uint32_t)n;
        *d_count = t;
    }
}

// seg_start holds seg_cap+1 entries; d_count (device) receives the number of segments
// (which may exceed seg_cap: the caller checks).
static int32_t find_segments_dev(pmrd_ctx* ctx, const uint64_t* d_keys, int64_t n_cap, const uint32_t* d_n, int shift,
                                 int64_t seg_cap, uint32_t* d_seg_start, uint32_t* d_count) {
    if (n_cap <= 0) {
        PMRD_CUDA(ctx, cudaMemsetAsync(d_count, 0, 4, ctx->stream));
        PMRD_CUDA(ctx, cudaMemsetAsync(d_seg_start, 0, 4, ctx->stream));
        return PMRD_OK;
    }
    const int n_tiles = (int)((n_cap + SEG_TILE - 1) / SEG_TILE);
    DevBuf th, tot;
    PMRD_CUDA(ctx, th.alloc(ctx->stream, (size_t)n_tiles * 4));
    PMRD_CUDA(ctx, tot.alloc(ctx->stream, 4));
    PMRD_LAUNCH(ctx, seg_count_kernel, n_tiles, SEG_THREADS, 0, d_keys, n_cap, d_n, shift, th.as<uint32_t>());
    PMRD_TRY((exclusive_scan<uint32_t, uint32_t>(ctx, th.as<uint32_t>(), th.as<uint32_t>(), n_tiles, tot.as<uint32_t>())));
    PMRD_LAUNCH(ctx, seg_write_kernel, n_tiles, SEG_THREADS, 0, d_keys, n_cap, d_n, shift, (const uint32_t*)th.as<uint32_t>(),
                (const uint32_t*)tot.as<uint32_t>(), seg_cap, d_seg_start, d_count);
    return PMRD_OK;
}

static int32_t read_count(pmrd_ctx* ctx, const uint32_t* d_count, int64_t* h) {
    uint32_t v = 0;
    PMRD_CUDA(ctx, cudaMemcpyAsync(&v, d_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *h = v;
    return PMRD_OK;
}

// ---------------------------------------------------------------- K5 consensus ----------
struct FamOut {
    uint64_t* key;
    uint32_t *size, *n_alt, *consensus, *qsum, *n_best, *flags;
};

// One family's row from its reads in input order; pl_at(i) = payload of read i of the family.
// wt: the 64-entry weight table 10^(q/10) in SHARED memory (the lanes of a warp look up different q, which a
// __constant__ table would serialise).
template <int MODE, typename PlAt>
__device__ __forceinline__ void family_row(uint32_t size, PlAt pl_at, const pmrd_umi_params& P, const double* wt, uint32_t& n_alt,
                                           uint32_t& consensus, uint32_t& qsum_best, uint32_t& n_best, uint32_t& flags) {
    n_alt = 0; consensus = 0; qsum_best = 0; n_best = 0; flags = 0;
    if (MODE == PMRD_MODE_WEIGHTED) {
        double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0, total = 0.0;
        uint32_t order = 0, seen = 0, n_seen = 0;  // order: 2 bits per allele = rank of first appearance
        for (uint32_t r = 0; r < size; ++r) {
            const uint32_t pl = pl_at(r);
            n_alt += pl >> 31;
            const uint32_t a = pl & 3u, q = (pl >> 2) & 63u;
            if ((int)q >= P.quality_threshold) {
                const double w = wt[q];
                // same sequential IEEE additions as umi.py:161-164
                w0 = __dadd_rn(w0, a == 0 ? w : 0.0);
                w1 = __dadd_rn(w1, a == 1 ? w : 0.0);
                w2 = __dadd_rn(w2, a == 2 ? w : 0.0);
                w3 = __dadd_rn(w3, a == 3 ? w : 0.0);
                total = __dadd_rn(total, w);
                if (!((seen >> a) & 1u)) {
                    seen |= 1u << a;
                    order |= n_seen << (2 * a);
                    ++n_seen;
                }
            }
        }
        if (n_seen > 0) {
            // max(allele_weights, key=allele_weights.get): first-inserted wins ties (dict order)
            int best = -1;
            double bw = -1.0;
            uint32_t brank = 4;
            const double ws[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (!((seen >> a) & 1u)) continue;
                const uint32_t rk = (order >> (2 * a)) & 3u;
                if (ws[a] > bw || (ws[a] == bw && rk < brank)) {
                    bw = ws[a];
                    best = a;
                    brank = rk;
                }
            }
            const double frac = __ddiv_rn(bw, total);
            if (total != 0.0 && !(frac < P.consensus_threshold)) {
                flags |= PMRD_FAM_HAS_CONSENSUS;
                consensus = (uint32_t)best;
                // consensus_quality = mean quality of the winner's qualifying reads (umi.py:177): a second, integer-only walk
                // over the family instead of eight conditional counters in the first one
                for (uint32_t r = 0; r < size; ++r) {
                    const uint32_t pl = pl_at(r);
                    const uint32_t q = (pl >> 2) & 63u;
                    if ((pl & 3u) == (uint32_t)best && (int)q >= P.quality_threshold) { qsum_best += q; ++n_best; }
                }
            }
        }
    } else if (MODE == PMRD_MODE_MAJORITY) {
        // per column: 4 x 16-bit counters packed in a u64, first-seen ranks in a u32
        uint64_t cnt[10];
        uint32_t ord[10];
#pragma unroll
        for (int c = 0; c < 10; ++c) { cnt[c] = 0; ord[c] = 0; }
        for (uint32_t r = 0; r < size; ++r) {
            const uint32_t pl = pl_at(r);
            n_alt += pl >> 31;
#pragma unroll
            for (int c = 0; c < 10; ++c) {
                const uint32_t b = (pl >> (18 - 2 * c)) & 3u;  // base c, base 0 in the top bits
                const uint64_t old = (cnt[c] >> (16 * b)) & 0xffffull;
                if (old == 0) {
                    // ord[c]: bits [7:0] ranks (2 bits per base), bits [10:8] = number seen
                    const uint32_t ns = (ord[c] >> 8) & 7u;
                    ord[c] = (ord[c] & ~(0x700u)) | ((ns + 1) << 8) | (ns << (2 * b));
                }
                if (old < 0xffffull) cnt[c] += 1ull << (16 * b);
            }
        }
        uint32_t seq = 0;
#pragma unroll
        for (int c = 0; c < 10; ++c) {
            uint32_t best = 0, bc = 0, brank = 4;
#pragma unroll
            for (uint32_t b = 0; b < 4; ++b) {
                const uint32_t cb = (uint32_t)((cnt[c] >> (16 * b)) & 0xffffull);
                const uint32_t rk = (ord[c] >> (2 * b)) & 3u;
                if (cb > bc || (cb == bc && cb > 0 && rk < brank)) {
                    bc = cb;
                    best = b;
                    brank = rk;
                }
            }
            seq |= best << (18 - 2 * c);
        }
        consensus = seq;
        flags |= PMRD_FAM_HAS_CONSENSUS;
    } else {
        for (uint32_t r = 0; r < size; ++r) n_alt += pl_at(r) >> 31;
        flags |= PMRD_FAM_HAS_CONSENSUS;
    }
    if ((int64_t)size >= (int64_t)P.min_family_size) flags |= PMRD_FAM_KEPT;
    if ((int64_t)size > (int64_t)P.max_family_size) flags |= PMRD_FAM_OVERSIZE;
}

template <int MODE, bool GATHER>
__global__ void __launch_bounds__(128)
family_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ payload_or_idx,
              const uint64_t* __restrict__ keys_orig, const uint32_t* __restrict__ payload_orig,
              const uint32_t* __restrict__ fam_start, const uint32_t* __restrict__ d_nfam, int64_t fam_cap,
              pmrd_umi_params P, uint32_t sentinel_umi, FamOut out) {
    __shared__ double s_wt[64];
    if (threadIdx.x < 64) s_wt[threadIdx.x] = c_wtab[threadIdx.x];
    __syncthreads();
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= (int64_t)*d_nfam || f >= fam_cap) return;
    const uint32_t lo = fam_start[f], hi = fam_start[f + 1];
    const uint32_t size = hi - lo;
    uint64_t key = keys[lo];
    if (GATHER) key = (key & 0xffffffff00000000ull) | (keys_orig[payload_or_idx[lo]] & 0xffffffffull);
    uint32_t n_alt, consensus, qsum_best, n_best, flags;
    family_row<MODE>(size, [&](uint32_t r) -> uint32_t { return GATHER ? payload_orig[payload_or_idx[lo + r]] : payload_or_idx[lo + r]; },
                     P, s_wt, n_alt, consensus, qsum_best, n_best, flags);
    if (sentinel_umi != 0u && (((uint32_t)key) >> 24) == 0xffu) flags = 0;  // padding of a tile-aligned cell
    out.key[f] = key;
    out.size[f] = size;
    out.n_alt[f] = n_alt;
    out.consensus[f] = consensus;
    out.qsum[f] = qsum_best;
    out.n_best[f] = n_best;
    out.flags[f] = flags;
}

// ---------------------------------------------------------------- fused consensus+count -
// Plan pipeline, cell-major tile-padded input: one block per 4096-read tile (a tile lies inside
// ONE cell).  Finds the family heads of the tile, votes each family that STARTS in the tile (one
// thread per family, reads in input order, same IEEE additions as family_kernel) and reduces
// straight to the three per-cell counts call_mrd needs -- no family table, no segment arrays,
// no group table: the 12 B/read are read once.
struct Vote {
    double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0, total = 0.0;
    uint32_t order = 0, seen = 0, n_seen = 0;
    // wt[q] = 10^(q/10) for q >= quality_threshold, else 0.0 (shared memory: the lanes of a warp look up
    // different q, which a __constant__ table would serialise).  A read below the threshold adds +0.0 --
    // the identity on these non-negative sums -- and sets no bit, so the walk is branch-free and still
    // performs exactly the IEEE additions of the reference loop, in the same order.
    __device__ __forceinline__ void add(uint32_t pl, const double* wt) {
        const uint32_t a = pl & 3u;
        const double w = wt[(pl >> 2) & 63u];
        total = __dadd_rn(total, w);
        asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
            "setp.eq.u32 p0, %4, 0;\n\tsetp.eq.u32 p1, %4, 1;\n\tsetp.eq.u32 p2, %4, 2;\n\tsetp.eq.u32 p3, %4, 3;\n\t"
            "@p0 add.rn.f64 %0, %0, %5;\n\t@p1 add.rn.f64 %1, %1, %5;\n\t"
            "@p2 add.rn.f64 %2, %2, %5;\n\t@p3 add.rn.f64 %3, %3, %5;\n\t}"
            : "+d"(w0), "+d"(w1), "+d"(w2), "+d"(w3) : "r"(a), "d"(w));
        const uint32_t bit = (__double2hiint(w) != 0 ? 1u : 0u) << a;   // w is 0.0 or >= 1
        const uint32_t fresh = bit & ~seen;
        order |= fresh ? n_seen << (2 * a) : 0u;
        n_seen += fresh ? 1u : 0u;
        seen |= bit;
    }
    // consensus allele, or -1 (umi.py:166-174)
    __device__ __forceinline__ int result(double cthr) const {
        if (n_seen == 0 || total == 0.0) return -1;
        int best = -1;
        double bw = -1.0;
        uint32_t brank = 4;
        const double ws[4] = {w0, w1, w2, w3};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (!((seen >> a) & 1u)) continue;
            const uint32_t rk = (order >> (2 * a)) & 3u;
            if (ws[a] > bw || (ws[a] == bw && rk < brank)) { bw = ws[a]; best = a; brank = rk; }
        }
        return (__ddiv_rn(bw, total) < cthr) ? -1 : best;
    }
};

#define CC_THREADS 256
#define CC_TILE 4096

// The same walk with the four per-allele sums kept in shared memory (acc[a * CC_THREADS], one
// column per thread): "sum[a] += w" is then one load, one DADD and one store instead of four
// DADDs and four 64-bit selects -- the allele picks the address, not the instruction.
struct VoteS {
    double* acc;
    double total = 0.0;
    uint32_t order = 0, seen = 0, n_seen = 0;
    __device__ __forceinline__ explicit VoteS(double* column) : acc(column) {
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a * CC_THREADS] = 0.0;
    }
    __device__ __forceinline__ void add(uint32_t pl, const double* wt) {
        const uint32_t a = pl & 3u;
        const double w = wt[(pl >> 2) & 63u];
        total = __dadd_rn(total, w);
        double* p = acc + a * CC_THREADS;
        *p = __dadd_rn(*p, w);
        const uint32_t bit = (__double2hiint(w) != 0 ? 1u : 0u) << a;   // w is 0.0 or >= 1
        const uint32_t fresh = bit & ~seen;
        order |= fresh ? n_seen << (2 * a) : 0u;
        n_seen += fresh ? 1u : 0u;
        seen |= bit;
    }
    __device__ __forceinline__ int result(double cthr) const {
        if (n_seen == 0 || total == 0.0) return -1;
        int best = -1;
        double bw = -1.0;
        uint32_t brank = 4;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (!((seen >> a) & 1u)) continue;
            const double wa = acc[a * CC_THREADS];
            const uint32_t rk = (order >> (2 * a)) & 3u;
            if (wa > bw || (wa == bw && rk < brank)) { bw = wa; best = a; brank = rk; }
        }
        return (__ddiv_rn(bw, total) < cthr) ? -1 : best;
    }
};
// consensus_threshold > 0.5: the winning allele must hold a strict majority of the weight, so two alleles tied for
// the best weight can never pass (each holds at most half) and the reference's tie-break -- the allele seen first
// wins -- cannot change the outcome.  The walk then only accumulates: no seen mask, no first-seen order.
struct VoteM {
    double* acc;
    double total = 0.0;
    __device__ __forceinline__ explicit VoteM(double* column) : acc(column) {
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a * CC_THREADS] = 0.0;
    }
    __device__ __forceinline__ void add(uint32_t pl, const double* wt) {
        const double w = wt[(pl >> 2) & 63u];
        total = __dadd_rn(total, w);
        double* p = acc + (pl & 3u) * CC_THREADS;
        *p = __dadd_rn(*p, w);
    }
    __device__ __forceinline__ int result(double cthr) const {
        if (total == 0.0) return -1;                 // every read was below the quality threshold
        int best = 0;
        double bw = acc[0];
#pragma unroll
        for (int a = 1; a < 4; ++a) {
            const double wa = acc[a * CC_THREADS];
            if (wa > bw) { bw = wa; best = a; }
        }
        return (__ddiv_rn(bw, total) < cthr) ? -1 : best;
    }
};
#define CC_ITEMS (CC_TILE / CC_THREADS)
#define CC_SKEW(i) ((i) + ((i) >> 5))

// Input: packed 8-byte records  rec = payload:32 | umi:32  in cell-major, tile-padded order,
// each cell sorted on the low `run_mask` bits of the UMI.  `run_mask` selects the UMI bits the
// sort ordered; a RUN is a maximal sequence of reads that agree on those bits and holds the
// reads of the (few) families whose UMIs differ only in the unsorted high bits, interleaved
// but each in input order (the sort is stable).  One thread owns a run and peels its families.
template <bool FULL>
__global__ void __launch_bounds__(CC_THREADS)
cell_consensus_kernel(const uint64_t* __restrict__ rec, const int64_t* __restrict__ cell_tile_start,
                      const int32_t* __restrict__ tile_cell /* cell of every tile (host-built) or NULL */, int n_cells,
                      pmrd_umi_params P, uint32_t run_mask, int64_t cell0, const uint8_t* __restrict__ alt_allele,
                      unsigned long long* __restrict__ o_fam, unsigned long long* __restrict__ o_tot,
                      unsigned long long* __restrict__ o_var) {
    // (s_umi is skewed by one word per 32: the head pass reads 16 CONSECUTIVE words per thread, which
    // unskewed is a 16-way bank conflict on every load -- it was 60 % of the kernel's shared-memory wavefronts)
    __shared__ __align__(16) uint32_t s_umi[CC_TILE + CC_TILE / 32];   // split halves: the vote walks only the payloads (32-bit, half the
    __shared__ uint32_t s_pl[CC_TILE];    // shared-memory traffic and bank conflicts of walking 64-bit records)
    __shared__ uint16_t s_head[CC_TILE];
    __shared__ uint32_t sw[33];
    __shared__ unsigned long long s_acc[3];
    __shared__ double s_wt[64];
    __shared__ uint32_t s_bkt[32];
    __shared__ int s_cell;
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * CC_TILE;
    if (threadIdx.x < 3) s_acc[threadIdx.x] = 0ull;
    if (threadIdx.x < 32) s_bkt[threadIdx.x] = 0u;
    if (threadIdx.x < 64) s_wt[threadIdx.x] = (int)threadIdx.x >= P.quality_threshold ? c_wtab[threadIdx.x] : 0.0;
    if (tile_cell) {
        if (threadIdx.x == 0) s_cell = tile_cell[tile];
    } else if (threadIdx.x == 0) {   // which cell owns this tile
        int lo = 0, hi = n_cells;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cell_tile_start[mid] <= tile) lo = mid; else hi = mid;
        }
        s_cell = lo;
    }
#pragma unroll
    for (int j = 0; j < CC_ITEMS; ++j) {
        const int i = j * CC_THREADS + threadIdx.x;
        const uint64_t v = rec[base + i];
        s_umi[CC_SKEW(i)] = (uint32_t)v;
        s_pl[i] = (uint32_t)(v >> 32);
    }
    __syncthreads();
    const int cell = s_cell;
    const int64_t cell_lo = cell_tile_start[cell] * CC_TILE, cell_hi = cell_tile_start[cell + 1] * CC_TILE;
    // heads, blocked so that the compacted list follows memory order
    uint32_t flags = 0;
    const int i0 = threadIdx.x * CC_ITEMS;
#pragma unroll
    for (int j = 0; j < CC_ITEMS; ++j) {
        const int i = i0 + j;
        bool head;
        if (i > 0) head = ((s_umi[CC_SKEW(i)] ^ s_umi[CC_SKEW(i - 1)]) & run_mask) != 0u;
        else head = base == cell_lo || ((((uint32_t)rec[base - 1]) ^ s_umi[0]) & run_mask) != 0u;
        flags |= (head ? 1u : 0u) << j;
    }
    uint32_t n_heads;
    uint32_t r = block_exclusive_scan<uint32_t>((uint32_t)__popc(flags), sw, n_heads);
    // compaction: walk the set bits only (a thread owns ~3 heads of its 16 records)
    while (flags) {
        const int j = __ffs((int)flags) - 1;
        flags &= flags - 1u;
        // FULL: bit 15 marks a run of padding reads (s_umi is recycled below)
        const uint32_t pad = (FULL && (s_umi[CC_SKEW(i0 + j)] >> 24) == 0xffu) ? 0x8000u : 0u;
        s_head[r++] = (uint16_t)((uint32_t)(i0 + j) | pad);
    }
    __syncthreads();
    // FULL: the UMI halves are not needed past this point; part of their 16 KB now holds the bucketed head order
    uint16_t* const s_order = reinterpret_cast<uint16_t*>(s_umi + 2 * CC_THREADS * 4);   // the other 8 KB: [CC_TILE]
    if (FULL) {
        // One thread votes one family, and a warp is as slow as its longest family (sizes are Poisson:
        // a plain head-order assignment keeps 17 of 32 lanes busy).  Bucket the heads by family length,
        // longest first, padding runs last, so that the lanes of a warp walk families of equal length.
        auto bucket = [&](uint32_t h) -> uint32_t {
            const uint32_t hv = s_head[h];
            const uint32_t hi = (h + 1 < n_heads) ? (uint32_t)(s_head[h + 1] & 0xfffu) : (uint32_t)CC_TILE;
            const uint32_t len = hi - (hv & 0xfffu);
            return (hv & 0x8000u) ? 31u : 30u - (len < 30u ? len : 30u);
        };
        for (uint32_t h = threadIdx.x; h < n_heads; h += CC_THREADS) atomicAdd(&s_bkt[bucket(h)], 1u);
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t c = s_bkt[threadIdx.x];
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (threadIdx.x >= (uint32_t)o) incl += t;
            }
            s_bkt[threadIdx.x] = incl - c;
        }
        __syncthreads();
        for (uint32_t h = threadIdx.x; h < n_heads; h += CC_THREADS) s_order[atomicAdd(&s_bkt[bucket(h)], 1u)] = (uint16_t)h;
        __syncthreads();
    }
    const uint32_t alt = alt_allele[cell0 + cell];
    uint32_t n_fam = 0, n_tot = 0, n_var = 0;
    for (uint32_t hh = threadIdx.x; hh < n_heads; hh += CC_THREADS) {
        const uint32_t h = FULL ? (uint32_t)s_order[hh] : hh;
        const uint32_t hv = s_head[h];
        const int lo = (int)(hv & 0xfffu);
        const int hi = (h + 1 < n_heads) ? (int)(s_head[h + 1] & 0xfffu) : CC_TILE;
        const int n_in = hi - lo;
        // the tile's last run may continue into the next tile(s) of the same cell
        int64_t g_hi = base + CC_TILE;
        if (h + 1 == n_heads) {
            const uint32_t rk = (FULL ? (uint32_t)rec[base + lo] : s_umi[CC_SKEW(lo)]) & run_mask;
            while (g_hi < cell_hi && ((((uint32_t)rec[g_hi]) & run_mask) == rk)) ++g_hi;
        }
        const int len = n_in + (int)(g_hi - (base + CC_TILE));
        auto umi_at = [&](int i) -> uint32_t { return i < n_in ? s_umi[CC_SKEW(lo + i)] : (uint32_t)rec[base + CC_TILE + (i - n_in)]; };
        auto pl_at = [&](int i) -> uint32_t { return i < n_in ? s_pl[lo + i] : (uint32_t)(rec[base + CC_TILE + (i - n_in)] >> 32); };
        if (FULL) {
            // the sort ordered every UMI bit: a run IS a family
            if (hv & 0x8000u) continue;                          // padding of the tile-aligned cell
            // generated reads carry the cell's ref or alt allele only: two sums in registers (vote.cuh); result 1 = the
            // tested (alt) allele, 0 = the other one, -1 = no consensus
            int best = -1;
            if (P.consensus_threshold > 0.5) {                   // (uniform) strict majority: accumulate only
                FoldVote2<true> v;
                for (int c = 0; c < n_in; ++c) v.add(s_pl[lo + c], alt, s_wt);
                for (int c = n_in; c < len; ++c) v.add(pl_at(c), alt, s_wt);
                if (len >= P.min_family_size) best = v.result(P.consensus_threshold);
            } else {
                FoldVote2<false> v;
                for (int c = 0; c < n_in; ++c) v.add(s_pl[lo + c], alt, s_wt);
                for (int c = n_in; c < len; ++c) v.add(pl_at(c), alt, s_wt);
                if (len >= P.min_family_size) best = v.result(P.consensus_threshold);
            }
            ++n_fam;
            if (best >= 0) { ++n_tot; n_var += best == 1; }
            continue;
        }
        // peel families: read a opens a family iff no earlier read of the run carries its UMI.
        // The UMIs already peeled are remembered in four registers (a run rarely holds more
        // families than that); beyond four the check falls back to scanning the run.
        uint32_t seen0 = 0, seen1 = 0, seen2 = 0, seen3 = 0;
        int n_seen_umi = 0;
        for (int a = 0; a < len; ++a) {
            const uint32_t umi = umi_at(a);
            if ((umi >> 24) == 0xffu) continue;                  // padding of the tile-aligned cell
            bool first;
            if (n_seen_umi <= 4) {
                first = !((n_seen_umi > 0 && umi == seen0) || (n_seen_umi > 1 && umi == seen1) ||
                          (n_seen_umi > 2 && umi == seen2) || (n_seen_umi > 3 && umi == seen3));
            } else {
                first = true;
                for (int b = 0; b < a && first; ++b) first = umi_at(b) != umi;
            }
            if (!first) continue;
            if (n_seen_umi == 0) seen0 = umi;
            else if (n_seen_umi == 1) seen1 = umi;
            else if (n_seen_umi == 2) seen2 = umi;
            else if (n_seen_umi == 3) seen3 = umi;
            ++n_seen_umi;
            Vote v;
            uint32_t size = 0;
            for (int c = a; c < len; ++c)
                if (umi_at(c) == umi) { v.add(pl_at(c), s_wt); ++size; }
            ++n_fam;
            if ((int)size >= P.min_family_size) {
                const int best = v.result(P.consensus_threshold);
                if (best >= 0) { ++n_tot; n_var += (uint32_t)best == alt; }
            }
        }
    }
    n_fam = warp_sum(n_fam); n_tot = warp_sum(n_tot); n_var = warp_sum(n_var);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_acc[0], (unsigned long long)n_fam);
        atomicAdd(&s_acc[1], (unsigned long long)n_tot);
        atomicAdd(&s_acc[2], (unsigned long long)n_var);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_acc[0]) atomicAdd(&o_fam[cell0 + cell], s_acc[0]);
        if (s_acc[1]) atomicAdd(&o_tot[cell0 + cell], s_acc[1]);
        if (s_acc[2]) atomicAdd(&o_var[cell0 + cell], s_acc[2]);
    }
}

// ---------------------------------------------------------------- fused consensus+count, bulk-copy form -------
// The same tile -> per-cell counts step (FULL sort only) with the tile brought in by the copy engine: one elected thread
// arms an mbarrier with the tile's 32 KB and issues ONE 1-D bulk async copy (cp.async.bulk.shared::cluster.global, SASS:
// UBLKCP), every thread waits on the barrier's phase.  The 4096 LDG.64 + 8192 STS.32 of the load/split loop and their
// address arithmetic leave the LSU -- the pipe ncu names as this kernel's limiter.  Records stay interleaved (umi | payload)
// in shared memory; the head pass is warp-striped (a warp's 64-bit loads of 32 consecutive records are conflict-free,
// neighbours compared through one shuffle) and leaves the one byte of every payload the vote needs (allele | quality)
// in a dense array for the walk.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(CC_THREADS)
cell_consensus_bulk_kernel(const uint64_t* __restrict__ rec, const int64_t* __restrict__ cell_tile_start, const int32_t* __restrict__ tile_cell,
                           pmrd_umi_params P, int64_t cell0, const uint8_t* __restrict__ alt_allele, unsigned long long* __restrict__ o_fam,
                           unsigned long long* __restrict__ o_tot, unsigned long long* __restrict__ o_var) {
    constexpr int WARPS = CC_THREADS / 32, PER_WARP = CC_TILE / WARPS, ITERS = PER_WARP / 32;
    __shared__ __align__(128) uint64_t s_rec[CC_TILE];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint8_t s_plb[CC_TILE];
    __shared__ uint16_t s_head[CC_TILE];
    uint16_t* const s_order = reinterpret_cast<uint16_t*>(s_rec);    // (the records are not read any more once the heads are known)
    __shared__ uint32_t s_wcount[WARPS];
    __shared__ unsigned long long s_acc[3];
    __shared__ double s_wt[64];
    __shared__ uint32_t s_bkt[32];
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * CC_TILE;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = lanemask_lt();
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(CC_TILE * 8)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(s_rec)), "l"(rec + base), "r"((uint32_t)(CC_TILE * 8)), "r"(bar) : "memory");
    }
    // (overlaps the copy)
    if (threadIdx.x < 3) s_acc[threadIdx.x] = 0ull;
    if (threadIdx.x < 32) s_bkt[threadIdx.x] = 0u;
    if (threadIdx.x < 64) s_wt[threadIdx.x] = (int)threadIdx.x >= P.quality_threshold ? c_wtab[threadIdx.x] : 0.0;
    const int cell = tile_cell[tile];
    const int64_t cell_lo = cell_tile_start[cell] * CC_TILE, cell_hi = cell_tile_start[cell + 1] * CC_TILE;
    const uint32_t prev_tile_umi = base == cell_lo ? 0u : (uint32_t)rec[base - 1];
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    // ---- heads, warp-striped: warp w owns records [w * 512, (w + 1) * 512)
    uint32_t hm[ITERS], pm[ITERS];
    uint32_t n_mine = 0;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int i = (int)warp * PER_WARP + it * 32 + (int)lane;
        const uint64_t r = s_rec[i];
        const uint32_t umi = (uint32_t)r;
        s_plb[i] = (uint8_t)(r >> 32);                                  // allele[1:0] | quality << 2
        uint32_t prev = __shfl_up_sync(0xffffffffu, umi, 1);
        if (lane == 0) prev = i > 0 ? (uint32_t)s_rec[i - 1] : prev_tile_umi;
        const bool head = (i == 0 && base == cell_lo) || umi != prev;
        hm[it] = __ballot_sync(0xffffffffu, head);
        pm[it] = __ballot_sync(0xffffffffu, (umi >> 24) == 0xffu);        // padding of the tile-aligned cell
        n_mine += __popc(hm[it]);
    }
    if (lane == 0) s_wcount[warp] = n_mine;
    __syncthreads();
    uint32_t hbase = 0, n_heads = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const uint32_t c = s_wcount[w];
        hbase += w < (int)warp ? c : 0u;
        n_heads += c;
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        if ((hm[it] >> lane) & 1u) {
            const int i = (int)warp * PER_WARP + it * 32 + (int)lane;
            s_head[hbase + __popc(hm[it] & lt)] = (uint16_t)((uint32_t)i | (((pm[it] >> lane) & 1u) ? 0x8000u : 0u));
        }
        hbase += __popc(hm[it]);
    }
    const uint32_t last_umi = (uint32_t)s_rec[CC_TILE - 1];           // the UMI of the tile's last run (it may continue in the next tile)
    __syncthreads();
    // ---- heads bucketed by family length (longest first, padding runs last): the lanes of a warp walk equal lengths
    auto bucket = [&](uint32_t h) -> uint32_t {
        const uint32_t hv = s_head[h];
        const uint32_t hi = (h + 1 < n_heads) ? (uint32_t)(s_head[h + 1] & 0xfffu) : (uint32_t)CC_TILE;
        const uint32_t len = hi - (hv & 0xfffu);
        return (hv & 0x8000u) ? 31u : 30u - (len < 30u ? len : 30u);
    };
    for (uint32_t h = threadIdx.x; h < n_heads; h += CC_THREADS) atomicAdd(&s_bkt[bucket(h)], 1u);
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t c = s_bkt[threadIdx.x];
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (threadIdx.x >= (uint32_t)o) incl += t;
        }
        s_bkt[threadIdx.x] = incl - c;
    }
    __syncthreads();
    for (uint32_t h = threadIdx.x; h < n_heads; h += CC_THREADS) s_order[atomicAdd(&s_bkt[bucket(h)], 1u)] = (uint16_t)h;
    __syncthreads();
    const uint32_t alt = alt_allele[cell0 + cell];
    uint32_t n_fam = 0, n_tot = 0, n_var = 0;
    for (uint32_t hh = threadIdx.x; hh < n_heads; hh += CC_THREADS) {
        const uint32_t h = (uint32_t)s_order[hh];
        const uint32_t hv = s_head[h];
        if (hv & 0x8000u) continue;                                      // padding
        const int lo = (int)(hv & 0xfffu);
        const int hi = (h + 1 < n_heads) ? (int)(s_head[h + 1] & 0xfffu) : CC_TILE;
        const int n_in = hi - lo;
        // the tile's last run may continue into the next tile(s) of the same cell
        int64_t g_hi = base + CC_TILE;
        if (h + 1 == n_heads) {
            while (g_hi < cell_hi && (uint32_t)rec[g_hi] == last_umi) ++g_hi;
        }
        const int len = n_in + (int)(g_hi - (base + CC_TILE));
        int best = -1;
        if (P.consensus_threshold > 0.5) {
            FoldVote2<true> v;
            for (int c = 0; c < n_in; ++c) v.add(s_plb[lo + c], alt, s_wt);
            for (int c = n_in; c < len; ++c) v.add((uint32_t)(rec[base + CC_TILE + (c - n_in)] >> 32), alt, s_wt);
            if (len >= P.min_family_size) best = v.result(P.consensus_threshold);
        } else {
            FoldVote2<false> v;
            for (int c = 0; c < n_in; ++c) v.add(s_plb[lo + c], alt, s_wt);
            for (int c = n_in; c < len; ++c) v.add((uint32_t)(rec[base + CC_TILE + (c - n_in)] >> 32), alt, s_wt);
            if (len >= P.min_family_size) best = v.result(P.consensus_threshold);
        }
        ++n_fam;
        if (best >= 0) { ++n_tot; n_var += best == 1; }
    }
    n_fam = warp_sum(n_fam); n_tot = warp_sum(n_tot); n_var = warp_sum(n_var);
    if (lane == 0) {
        atomicAdd(&s_acc[0], (unsigned long long)n_fam);
        atomicAdd(&s_acc[1], (unsigned long long)n_tot);
        atomicAdd(&s_acc[2], (unsigned long long)n_var);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_acc[0]) atomicAdd(&o_fam[cell0 + cell], s_acc[0]);
        if (s_acc[1]) atomicAdd(&o_tot[cell0 + cell], s_acc[1]);
        if (s_acc[2]) atomicAdd(&o_var[cell0 + cell], s_acc[2]);
    }
}

// ---------------------------------------------------------------- fused consensus+count after a 16-BIT sort ---------
// Grouping, not sorting, is what the vote needs: after TWO radix passes a cell is ordered on the low 16 UMI bits and a
// RUN (equal low bits) holds the reads of the one to three families that differ only in the unsorted high byte,
// interleaved, each still in input order (the passes are stable).  The third pass over HBM (16 B per record + its
// counting pass) is replaced by work inside this kernel's shared-memory tile:
//   * the head pass also leaves the high byte of every record in a dense array and marks the runs in which it changes;
//   * a PURE run (the majority: 2^16 possible values for at most 5e4 families per cell) is one family, voted as in
//     cell_consensus_bulk_kernel;
//   * a MIXED run is swept: the first sweep folds two families at once (the first read's and the first that differs),
//     branch-free -- a read of another family adds +0.0 -- and notes the two smallest other high bytes present (ordered
//     from the run's first byte, cyclically: families are counted, their order is free); those rare third and fourth
//     families become (run, byte) items of a work list that the block sweeps afterwards, one item per thread.  Pure and
//     mixed runs are bucketed apart (and by length), so the lanes of a warp do the same work;
//   * the tile's last run may continue in the next tile(s): it is swept through global memory by its owner.
// Padding records (UMI 0xffffffff) share their low bits with real UMIs: a change between padding and reads opens a run
// too, so padding forms runs of its own (skipped) exactly as under the full sort.
struct CpSmem {
    uint64_t rec[CC_TILE];             // the tile; recycled as the bucketed run order (uint16) once the heads are known
    uint16_t head[CC_TILE];
    uint16_t hp[CC_TILE];              // high UMI byte << 8 | payload byte (allele[1:0] | quality << 2)
    uint8_t mixed[CC_TILE];            // [h] != 0: run h holds more than one distinct UMI (plain stores of 1: no atomics)
    double wt[64];
    unsigned long long acc[3];
    unsigned long long bar;
    uint32_t bkt[64];
    uint32_t wcount[CC_THREADS / 32];
    uint32_t n_items;
};

template <bool MAJ>
__device__ __forceinline__ void cp_finish(const FoldVote2<MAJ>& v, int size, const pmrd_umi_params& P, uint32_t& n_fam, uint32_t& n_tot, uint32_t& n_var) {
    ++n_fam;
    if (size >= P.min_family_size) {
        const int best = v.result(P.consensus_threshold);
        if (best >= 0) { ++n_tot; n_var += best == 1; }
    }
}
// Sweeps of a mixed run are BRANCH-FREE: a read of another family adds +0.0 (FoldVote2::add_w), so the lanes of a warp,
// whose runs interleave their families differently, stay converged.  The first sweep folds TWO families at once (the
// first read's and the first one that differs: most mixed runs hold exactly two); what is left is taken one family per
// sweep in increasing (hi - t0) mod 256.
// one family per sweep, in increasing (hi - t0) mod 256 from `key` on (key1 was folded by the first sweep); once = only `key`
template <bool MAJ, typename HpAt>
__device__ __forceinline__ void cp_sweep_chain(int len, HpAt hp_at, uint32_t t0, uint32_t key, uint32_t key1, bool once, uint32_t alt,
                                               const double* wt, const pmrd_umi_params& P, uint32_t& n_fam, uint32_t& n_tot, uint32_t& n_var) {
    while (key != 256u) {
        FoldVote2<MAJ> v;
        int size = 0;
        uint32_t nxt = 256u;
        for (int c = 0; c < len; ++c) {
            const uint32_t x = hp_at(c);
            const uint32_t k = ((x >> 8) - t0) & 0xffu;
            const bool m = k == key;
            v.add_w(x, alt, m ? wt[(x >> 2) & 63u] : 0.0);
            size += m;
            nxt = (k > key && k != key1) ? min(nxt, k) : nxt;
        }
        cp_finish<MAJ>(v, size, P, n_fam, n_tot, n_var);
        if (once) break;
        key = nxt;
    }
}
// first sweep: folds the first read's family and the first family that differs; n1 < n2: the two smallest other keys (256: none)
template <bool MAJ, typename HpAt>
__device__ __forceinline__ void cp_sweep_first(int len, HpAt hp_at, uint32_t alt, const double* wt, const pmrd_umi_params& P, uint32_t& key1,
                                               uint32_t& n1, uint32_t& n2, uint32_t& n_fam, uint32_t& n_tot, uint32_t& n_var) {
    const uint32_t t0 = hp_at(0) >> 8;
    key1 = 256u;                                            // (256: the run is pure -- only the tile's last run can be)
    for (int c = 1; c < len; ++c) {
        const uint32_t k = ((hp_at(c) >> 8) - t0) & 0xffu;
        if (k) { key1 = k; break; }
    }
    n1 = n2 = 256u;
    FoldVote2<MAJ> a, b;
    int sa = 0, sb = 0;
    for (int c = 0; c < len; ++c) {
        const uint32_t x = hp_at(c);
        const uint32_t k = ((x >> 8) - t0) & 0xffu;
        const double w = wt[(x >> 2) & 63u];
        const bool ma = k == 0u, mb = k == key1;
        a.add_w(x, alt, ma ? w : 0.0);
        b.add_w(x, alt, mb ? w : 0.0);
        sa += ma;
        sb += mb;
        const uint32_t kk = (ma || mb) ? 256u : k;          // a third family's key, else neutral
        const uint32_t hi = max(n1, kk);
        n1 = min(n1, kk);
        n2 = hi == n1 ? n2 : min(n2, hi);                   // (hi == n1: kk repeats the smallest key, or both are 256)
    }
    cp_finish<MAJ>(a, sa, P, n_fam, n_tot, n_var);
    if (key1 != 256u) cp_finish<MAJ>(b, sb, P, n_fam, n_tot, n_var);
}

template <bool MAJ>
__global__ void __launch_bounds__(CC_THREADS)
cell_consensus_peel_kernel(const uint64_t* __restrict__ rec, const int64_t* __restrict__ cell_tile_start, const int32_t* __restrict__ tile_cell,
                           pmrd_umi_params P, int64_t cell0, const uint8_t* __restrict__ alt_allele, unsigned long long* __restrict__ o_fam,
                           unsigned long long* __restrict__ o_tot, unsigned long long* __restrict__ o_var) {
    constexpr int WARPS = CC_THREADS / 32, PER_WARP = CC_TILE / WARPS, ITERS = PER_WARP / 32;
    extern __shared__ __align__(128) unsigned char cp_raw[];
    CpSmem& S = *reinterpret_cast<CpSmem*>(cp_raw);
    uint16_t* const s_order = reinterpret_cast<uint16_t*>(S.rec);                       // [CC_TILE]: 8 KB
    uint32_t* const s_items = reinterpret_cast<uint32_t*>(S.rec) + CC_TILE / 2;         // behind it: (run, key) items, at most one per read
    const int64_t tile = blockIdx.x;
    const int64_t base = tile * CC_TILE;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = lanemask_lt();
    const uint32_t bar = smem_u32(&S.bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(CC_TILE * 8)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(S.rec)), "l"(rec + base), "r"((uint32_t)(CC_TILE * 8)), "r"(bar) : "memory");
    }
    // (overlaps the copy)
    if (threadIdx.x < 3) S.acc[threadIdx.x] = 0ull;
    if (threadIdx.x == 3) S.n_items = 0u;
    if (threadIdx.x < 64) S.bkt[threadIdx.x] = 0u;
    if (threadIdx.x < 64) S.wt[threadIdx.x] = (int)threadIdx.x >= P.quality_threshold ? c_wtab[threadIdx.x] : 0.0;
    reinterpret_cast<uint4*>(S.mixed)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
    static_assert(CC_TILE == 16 * CC_THREADS, "one 16-byte slice of the run flags per thread");
    const int cell = tile_cell[tile];
    const int64_t cell_lo = cell_tile_start[cell] * CC_TILE, cell_hi = cell_tile_start[cell + 1] * CC_TILE;
    const uint32_t prev_tile_umi = base == cell_lo ? 0u : (uint32_t)rec[base - 1];
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    // ---- heads of the 16-bit runs, warp-striped: warp w owns records [w * 512, (w + 1) * 512)
    uint32_t hm[ITERS], pm[ITERS], im[ITERS];
    uint32_t n_mine = 0;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int i = (int)warp * PER_WARP + it * 32 + (int)lane;
        const uint64_t r = S.rec[i];
        const uint32_t umi = (uint32_t)r;
        const bool pad = (umi >> 24) == 0xffu;
        S.hp[i] = (uint16_t)(((umi >> 8) & 0xff00u) | ((uint32_t)(r >> 32) & 0xffu));
        uint32_t prev = __shfl_up_sync(0xffffffffu, umi, 1);
        if (lane == 0) prev = i > 0 ? (uint32_t)S.rec[i - 1] : prev_tile_umi;
        const bool head = (i == 0 && base == cell_lo) || ((umi ^ prev) & 0xffffu) != 0u || pad != ((prev >> 24) == 0xffu);
        hm[it] = __ballot_sync(0xffffffffu, head);
        pm[it] = __ballot_sync(0xffffffffu, pad);
        im[it] = __ballot_sync(0xffffffffu, !head && umi != prev);    // the run changes its high byte here
        n_mine += __popc(hm[it]);
    }
    if (lane == 0) S.wcount[warp] = n_mine;
    const uint32_t last_lo = (uint32_t)S.rec[CC_TILE - 1] & 0xffffu;    // low bits of the tile's last run (it may continue in the next tile)
    __syncthreads();
    uint32_t hbase = 0, n_heads = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const uint32_t c = S.wcount[w];
        hbase += w < (int)warp ? c : 0u;
        n_heads += c;
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const uint32_t before = __popc(hm[it] & lt);
        if ((hm[it] >> lane) & 1u) {
            const int i = (int)warp * PER_WARP + it * 32 + (int)lane;
            S.head[hbase + before] = (uint16_t)((uint32_t)i | (((pm[it] >> lane) & 1u) ? 0x8000u : 0u));
        } else if ((im[it] >> lane) & 1u) {
            const uint32_t run = hbase + before;                       // heads up to this record; its run is number run - 1
            if (run > 0u) S.mixed[run - 1u] = 1;
        }
        hbase += __popc(hm[it]);
    }
    __syncthreads();
    // ---- runs bucketed: mixed runs first (by length, longest first), then pure runs (the same), padding runs last
    auto bucket = [&](uint32_t h) -> uint32_t {
        const uint32_t hv = S.head[h];
        const uint32_t hi = (h + 1 < n_heads) ? (uint32_t)(S.head[h + 1] & 0xfffu) : (uint32_t)CC_TILE;
        const uint32_t len = hi - (hv & 0xfffu);
        if (hv & 0x8000u) return 63u;
        const bool mixed = h + 1 == n_heads || S.mixed[h] != 0;
        return (mixed ? 0u : 31u) + 30u - (len < 30u ? len : 30u);
    };
    constexpr int BK_KEEP = 5;                                // a thread owns n_heads / 256 runs: their buckets (6 bits each) stay in a register
    uint32_t bk = 0u;
#pragma unroll
    for (int j = 0; j < BK_KEEP; ++j) {
        const uint32_t h = threadIdx.x + j * CC_THREADS;
        if (h < n_heads) {
            const uint32_t b = bucket(h);
            bk |= b << (6 * j);
            atomicAdd(&S.bkt[b], 1u);
        }
    }
    for (uint32_t h = threadIdx.x + BK_KEEP * CC_THREADS; h < n_heads; h += CC_THREADS) atomicAdd(&S.bkt[bucket(h)], 1u);
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t c0 = S.bkt[threadIdx.x], c1 = S.bkt[32 + threadIdx.x];
        uint32_t i0 = c0, i1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
            if (threadIdx.x >= (uint32_t)o) { i0 += t0; i1 += t1; }
        }
        const uint32_t tot0 = __shfl_sync(0xffffffffu, i0, 31);
        S.bkt[threadIdx.x] = i0 - c0;
        S.bkt[32 + threadIdx.x] = tot0 + i1 - c1;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < BK_KEEP; ++j) {
        const uint32_t h = threadIdx.x + j * CC_THREADS;
        if (h < n_heads) s_order[atomicAdd(&S.bkt[(bk >> (6 * j)) & 63u], 1u)] = (uint16_t)h;
    }
    for (uint32_t h = threadIdx.x + BK_KEEP * CC_THREADS; h < n_heads; h += CC_THREADS) s_order[atomicAdd(&S.bkt[bucket(h)], 1u)] = (uint16_t)h;
    __syncthreads();
    const uint32_t alt = alt_allele[cell0 + cell];
    uint32_t n_fam = 0, n_tot = 0, n_var = 0;
    for (uint32_t hh = threadIdx.x; hh < n_heads; hh += CC_THREADS) {
        const uint32_t h = (uint32_t)s_order[hh];
        const uint32_t hv = S.head[h];
        if (hv & 0x8000u) continue;                                      // padding
        const int lo = (int)(hv & 0xfffu);
        const int hi = (h + 1 < n_heads) ? (int)(S.head[h + 1] & 0xfffu) : CC_TILE;
        const int n_in = hi - lo;
        if (h + 1 == n_heads) {
            // the tile's last run: it may continue into the next tile(s) of the same cell
            int64_t g_hi = base + CC_TILE;
            for (; g_hi < cell_hi; ++g_hi) {
                const uint32_t u = (uint32_t)rec[g_hi];
                if ((u & 0xffffu) != last_lo || (u >> 24) == 0xffu) break;
            }
            const int len = n_in + (int)(g_hi - (base + CC_TILE));
            const uint64_t* __restrict__ more = rec + base + CC_TILE - n_in;
            auto hp_at = [&](int c) -> uint32_t {
                if (c < n_in) return S.hp[lo + c];
                const uint64_t r = more[c];
                return (((uint32_t)r >> 8) & 0xff00u) | ((uint32_t)(r >> 32) & 0xffu);
            };
            uint32_t key1, n1, n2;
            cp_sweep_first<MAJ>(len, hp_at, alt, S.wt, P, key1, n1, n2, n_fam, n_tot, n_var);
            cp_sweep_chain<MAJ>(len, hp_at, hp_at(0) >> 8, n1, key1, false, alt, S.wt, P, n_fam, n_tot, n_var);
        } else if (S.mixed[h] != 0) {
            const uint16_t* const ph = S.hp + lo;
            uint32_t key1, n1, n2;
            cp_sweep_first<MAJ>(n_in, [&](int c) -> uint32_t { return ph[c]; }, alt, S.wt, P, key1, n1, n2, n_fam, n_tot, n_var);
            // a third (fourth, ...) family: not swept here, where most lanes have none, but queued as (run, key) items
            if (n1 != 256u) {
                const uint32_t cnt = n2 != 256u ? 2u : 1u;
                const uint32_t at = atomicAdd(&S.n_items, cnt);
                s_items[at] = h | (n1 << 12) | (key1 << 20);
                if (cnt == 2u) s_items[at + 1] = h | (n2 << 12) | (key1 << 20) | 0x10000000u;     // ... and whatever follows n2
            }
        } else {
            FoldVote2<MAJ> v;
            for (int c = 0; c < n_in; ++c) v.add(S.hp[lo + c], alt, S.wt);
            int best = -1;
            if (n_in >= P.min_family_size) best = v.result(P.consensus_threshold);
            ++n_fam;
            if (best >= 0) { ++n_tot; n_var += best == 1; }
        }
    }
    __syncthreads();
    for (uint32_t w = threadIdx.x; w < S.n_items; w += CC_THREADS) {
        const uint32_t it = s_items[w];
        const uint32_t h = it & 0xfffu;
        const int lo = (int)(S.head[h] & 0xfffu), n_in = (int)(S.head[h + 1] & 0xfffu) - lo;      // (never the tile's last run)
        const uint16_t* const ph = S.hp + lo;
        cp_sweep_chain<MAJ>(n_in, [&](int c) -> uint32_t { return ph[c]; }, (uint32_t)ph[0] >> 8, (it >> 12) & 0xffu, (it >> 20) & 0xffu,
                            !(it & 0x10000000u), alt, S.wt, P, n_fam, n_tot, n_var);
    }
    n_fam = warp_sum(n_fam); n_tot = warp_sum(n_tot); n_var = warp_sum(n_var);
    if (lane == 0) {
        atomicAdd(&S.acc[0], (unsigned long long)n_fam);
        atomicAdd(&S.acc[1], (unsigned long long)n_tot);
        atomicAdd(&S.acc[2], (unsigned long long)n_var);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (S.acc[0]) atomicAdd(&o_fam[cell0 + cell], S.acc[0]);
        if (S.acc[1]) atomicAdd(&o_tot[cell0 + cell], S.acc[1]);
        if (S.acc[2]) atomicAdd(&o_var[cell0 + cell], S.acc[2]);
    }
}

// ---------------------------------------------------------------- small cells: one block, no HBM round trip ----
// Cells of at most SC_THREADS * ITEMS reads (the C5 panel: ~1250 reads per (sample, locus) cell) are not worth a
// tile-padded trip through the global radix passes.  Their records sit back to back (compact layout, no padding);
// one block loads a cell, sorts it by the 24 UMI bits with three ranked passes that never leave shared memory,
// finds the family heads and votes them -- the 8 bytes per read are read once and nothing but the three per-cell
// counts is written.  Ranking and vote are those of the tile kernels (ballot peer masks, stable; VoteM / VoteS).
#define SC_THREADS 256
template <int ITEMS>
__global__ void __launch_bounds__(SC_THREADS)
small_cell_kernel(const uint64_t* __restrict__ rec, const ReadCell* __restrict__ cells, int n_cells, pmrd_umi_params P, int64_t cell0,
                  const uint8_t* __restrict__ alt_allele, unsigned long long* __restrict__ o_fam, unsigned long long* __restrict__ o_tot,
                  unsigned long long* __restrict__ o_var, int sort_bits /* 16: two ranked passes + peel; 24: three */) {
    constexpr int CAP = SC_THREADS * ITEMS, WARPS = SC_THREADS / 32, PER_WARP = CAP / WARPS;
    __shared__ __align__(16) uint64_t s_rec[CAP];
    __shared__ __align__(16) unsigned char s_raw[4 * SC_THREADS * sizeof(double)];   // ranking counters, then the votes' allele sums
    static_assert(sizeof(s_raw) >= WARPS * 256 * sizeof(uint32_t), "counters fit the shared scratch");
    uint32_t (*s_cnt)[256] = reinterpret_cast<uint32_t (*)[256]>(s_raw);
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t sw[33];
    __shared__ double s_wt[64];
    __shared__ unsigned long long s_acc[3];
    const int cell = blockIdx.x;
    if (cell >= n_cells) return;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    const int n = (int)cells[cell].n_reads;
    const uint64_t* __restrict__ src = rec + cells[cell].out_start;
    if (threadIdx.x < 3) s_acc[threadIdx.x] = 0ull;
    if (threadIdx.x < 64) s_wt[threadIdx.x] = (int)threadIdx.x >= P.quality_threshold ? c_wtab[threadIdx.x] : 0.0;
    // warp-striped: item i of lane l = warp * PER_WARP + i * 32 + l (memory order = item-major, lane-minor: stable)
    uint64_t key[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int loc = (int)warp * PER_WARP + i * 32 + (int)lane;
        key[i] = loc < n ? src[loc] : 0xffffffffull;        // filler: UMI word 0xFFFFFFFF sorts last and is no 12-mer
    }
    // (sort_bits = 16: a cell of <= 4096 reads has < 1000 families over 65 536 values of the low 16 bits, so a run of equal low
    // bits is one family but for ~1 % of the runs -- those are separated by the vote below, cp_sweep_first / _chain, and the
    // third ranked pass is not made)
#pragma unroll 1
    for (int shift = 0; shift < sort_bits; shift += 8) {
        for (int i = threadIdx.x; i < WARPS * 256; i += SC_THREADS) (&s_cnt[0][0])[i] = 0u;
        __syncthreads();
        uint32_t rank2[ITEMS / 2 > 0 ? (ITEMS + 1) / 2 : 1];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = ((uint32_t)key[i] >> shift) & 0xffu;
            const uint32_t m = rs_peer_mask(d, 0xffffffffu);
            const uint32_t before = __popc(m & lt);
            const int leader = __ffs((int)m) - 1;
            uint32_t old = 0;
            if (before == 0) old = atomicAdd(&s_cnt[warp][d], (uint32_t)__popc(m));
            old = __shfl_sync(0xffffffffu, old, leader);
            const uint32_t r = old + before;
            if (i & 1) rank2[i >> 1] |= r << 16;
            else rank2[i >> 1] = r;
        }
        __syncthreads();
        {   // per digit: exclusive prefix over the warps, then over the digits
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = s_cnt[w][threadIdx.x];
                s_cnt[w][threadIdx.x] = run;
                run += c;
            }
            uint32_t tot;
            s_base[threadIdx.x] = block_exclusive_scan<uint32_t>(run, sw, tot);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = ((uint32_t)key[i] >> shift) & 0xffu;
            const uint32_t r = (i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu;
            s_rec[s_base[d] + s_cnt[warp][d] + r] = key[i];
        }
        __syncthreads();
        if (shift + 8 < sort_bits) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = s_rec[warp * PER_WARP + i * 32 + lane];
        }
    }
    const uint32_t run_mask = sort_bits >= 24 ? 0xffffffffu : (1u << sort_bits) - 1u;
    // sorted: real reads occupy [0, n).  The family heads are compacted into a list first (a warp's 32 consecutive
    // records hold ~6 heads: voting in place would leave 26 lanes idle), then one thread votes one family.
    constexpr bool LIST = ITEMS <= 12;                       // (the list must fit the 48 KB of static shared memory)
    __shared__ uint16_t s_list[LIST ? CAP : 1];
    __shared__ uint32_t s_nheads;
    if (threadIdx.x == 0) s_nheads = 0u;
    __syncthreads();
    const uint32_t alt = alt_allele[cell0 + cell];
    uint32_t n_fam = 0, n_tot = 0, n_var = 0;
    auto vote_family = [&](int i) {
        const uint32_t umi = (uint32_t)s_rec[i];
        int len = 1;
        bool pure = true;
        for (; i + len < n; ++len) {
            const uint32_t u = (uint32_t)s_rec[i + len];
            if ((u ^ umi) & run_mask) break;
            pure &= u == umi;
        }
        if (!pure) {
            // several families in the run (they differ in the unsorted high byte): swept as in cell_consensus_peel_kernel
            auto hp_at = [&](int c) -> uint32_t {
                const uint64_t r = s_rec[i + c];
                return (((uint32_t)r >> 8) & 0xff00u) | ((uint32_t)(r >> 32) & 0xffu);
            };
            uint32_t key1, n1, n2;
            if (P.consensus_threshold > 0.5) {
                cp_sweep_first<true>(len, hp_at, alt, s_wt, P, key1, n1, n2, n_fam, n_tot, n_var);
                cp_sweep_chain<true>(len, hp_at, hp_at(0) >> 8, n1, key1, false, alt, s_wt, P, n_fam, n_tot, n_var);
            } else {
                cp_sweep_first<false>(len, hp_at, alt, s_wt, P, key1, n1, n2, n_fam, n_tot, n_var);
                cp_sweep_chain<false>(len, hp_at, hp_at(0) >> 8, n1, key1, false, alt, s_wt, P, n_fam, n_tot, n_var);
            }
            return;
        }
        int best = -1;
        if (P.consensus_threshold > 0.5) {
            FoldVote2<true> v;
            for (int c = 0; c < len; ++c) v.add((uint32_t)(s_rec[i + c] >> 32), alt, s_wt);
            if (len >= P.min_family_size) best = v.result(P.consensus_threshold);
        } else {
            FoldVote2<false> v;
            for (int c = 0; c < len; ++c) v.add((uint32_t)(s_rec[i + c] >> 32), alt, s_wt);
            if (len >= P.min_family_size) best = v.result(P.consensus_threshold);
        }
        ++n_fam;
        if (best >= 0) { ++n_tot; n_var += best == 1; }
    };
    if (LIST) {
        for (int i0 = 0; i0 < n; i0 += SC_THREADS) {
            const int i = i0 + (int)threadIdx.x;
            const bool head = i < n && (i == 0 || (((uint32_t)s_rec[i - 1] ^ (uint32_t)s_rec[i]) & run_mask) != 0u);
            const uint32_t hm = __ballot_sync(0xffffffffu, head);
            uint32_t pos = 0;
            if (lane == 0 && hm) pos = atomicAdd(&s_nheads, (uint32_t)__popc(hm));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (head) s_list[pos + __popc(hm & lt)] = (uint16_t)i;
        }
        __syncthreads();
        const int n_heads = (int)s_nheads;
        for (int h = threadIdx.x; h < n_heads; h += SC_THREADS) vote_family((int)s_list[h]);
    } else {
        for (int i = threadIdx.x; i < n; i += SC_THREADS)
            if (i == 0 || (((uint32_t)s_rec[i - 1] ^ (uint32_t)s_rec[i]) & run_mask) != 0u) vote_family(i);
    }
    n_fam = warp_sum(n_fam); n_tot = warp_sum(n_tot); n_var = warp_sum(n_var);
    if (lane == 0) {
        atomicAdd(&s_acc[0], (unsigned long long)n_fam);
        atomicAdd(&s_acc[1], (unsigned long long)n_tot);
        atomicAdd(&s_acc[2], (unsigned long long)n_var);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        o_fam[cell0 + cell] = s_acc[0];
        o_tot[cell0 + cell] = s_acc[1];
        o_var[cell0 + cell] = s_acc[2];
    }
}

// cells of a compact chunk (every cell <= max_cell_reads <= 4096 reads), records back to back at cells[c].out_start
int32_t k5_collapse_counts_small(pmrd_ctx* ctx, const uint64_t* d_rec, const ReadCell* d_cells, int64_t n_cells, int64_t max_cell_reads,
                                 const pmrd_umi_params* params, int64_t cell0, const uint8_t* d_alt_allele, int64_t* d_o_fam,
                                 int64_t* d_o_tot, int64_t* d_o_var) {
    if (n_cells <= 0) return PMRD_OK;
    PMRD_ARG(ctx, max_cell_reads <= 4096, "small-cell collapse: a cell holds more than 4096 reads");
    unsigned long long *f = (unsigned long long*)d_o_fam, *t = (unsigned long long*)d_o_tot, *v = (unsigned long long*)d_o_var;
    const char* sb = getenv("PMRD_SMALL_SORT_BITS");                // 24: three ranked passes, every run is a family
    const int sort_bits = sb && atoi(sb) == 24 ? 24 : 16;
    // records per thread: the smallest that holds the largest cell (a slot beyond the cell's reads still costs a rank)
#define SC_CASE(I) \
    if (max_cell_reads <= (int64_t)SC_THREADS * (I)) {                                                                          \
        PMRD_LAUNCH_NAMED(ctx, "small_cell_kernel", small_cell_kernel<I>, (int)n_cells, SC_THREADS, 0, d_rec, d_cells, (int)n_cells, \
                          *params, cell0, d_alt_allele, f, t, v, sort_bits);                                                     \
        return PMRD_OK;                                                                                                          \
    }
    SC_CASE(1) SC_CASE(2) SC_CASE(3) SC_CASE(4) SC_CASE(5) SC_CASE(6) SC_CASE(7) SC_CASE(8) SC_CASE(10) SC_CASE(12) SC_CASE(14) SC_CASE(16)
#undef SC_CASE
    return PMRD_OK;
}

// segmented sort of packed records + fused consensus/count; outputs accumulate into the per-cell arrays
int32_t k5_collapse_counts_segmented(pmrd_ctx* ctx, uint64_t* d_rec, uint64_t* d_rec_tmp, int64_t n_out,
                                     const pmrd_umi_params* params, const int64_t* d_cell_tile_start, const int32_t* d_tile_cell,
                                     int64_t n_cells, int sort_bits, int umi_bits, int64_t cell0, const uint8_t* d_alt_allele,
                                     int64_t* d_o_fam, int64_t* d_o_tot, int64_t* d_o_var, uint32_t* d_hist0) {
    if (n_out <= 0) return PMRD_OK;
    PMRD_ARG(ctx, n_out % CC_TILE == 0, "segmented collapse: input must be whole tiles");
    int in_b = 0;
    PMRD_TRY(radix_sort_pairs_segmented(ctx, d_rec, nullptr, d_rec_tmp, nullptr, n_out, d_cell_tile_start, n_cells, sort_bits, &in_b, d_hist0));
    const uint64_t* sr = in_b ? d_rec_tmp : d_rec;
    const uint32_t run_mask = sort_bits >= 32 ? 0xffffffffu : ((1u << sort_bits) - 1u);
    const int grid = (int)(n_out / CC_TILE);
    const char* cb = getenv("PMRD_CC_BULK");                 // 0: the LSU-loaded form of the consensus kernel
    if (sort_bits == 16 && umi_bits == 24 && d_tile_cell && (!cb || atoi(cb) != 0)) {
        // two passes ordered the low 16 bits: the consensus kernel resolves the high byte inside its tile
        PMRD_CUDA(ctx, cudaFuncSetAttribute(cell_consensus_peel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CpSmem)));
        PMRD_CUDA(ctx, cudaFuncSetAttribute(cell_consensus_peel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CpSmem)));
        if (params->consensus_threshold > 0.5)
            PMRD_LAUNCH_NAMED(ctx, "cell_consensus_peel_kernel", cell_consensus_peel_kernel<true>, grid, CC_THREADS, sizeof(CpSmem), sr, d_cell_tile_start,
                              d_tile_cell, *params, cell0, d_alt_allele, (unsigned long long*)d_o_fam, (unsigned long long*)d_o_tot,
                              (unsigned long long*)d_o_var);
        else
            PMRD_LAUNCH_NAMED(ctx, "cell_consensus_peel_kernel", cell_consensus_peel_kernel<false>, grid, CC_THREADS, sizeof(CpSmem), sr, d_cell_tile_start,
                              d_tile_cell, *params, cell0, d_alt_allele, (unsigned long long*)d_o_fam, (unsigned long long*)d_o_tot,
                              (unsigned long long*)d_o_var);
    } else if (sort_bits >= umi_bits && d_tile_cell && (!cb || atoi(cb) != 0)) {
        PMRD_LAUNCH_NAMED(ctx, "cell_consensus_kernel<true>", cell_consensus_bulk_kernel, grid, CC_THREADS, 0, sr, d_cell_tile_start, d_tile_cell, *params,
                          cell0, d_alt_allele, (unsigned long long*)d_o_fam, (unsigned long long*)d_o_tot, (unsigned long long*)d_o_var);
    } else if (sort_bits >= umi_bits) {
        PMRD_LAUNCH(ctx, cell_consensus_kernel<true>, grid, CC_THREADS, 0, sr, d_cell_tile_start, d_tile_cell, (int)n_cells, *params, 0xffffffffu,
                    cell0, d_alt_allele, (unsigned long long*)d_o_fam, (unsigned long long*)d_o_tot, (unsigned long long*)d_o_var);
    } else {
        PMRD_LAUNCH(ctx, cell_consensus_kernel<false>, grid, CC_THREADS, 0, sr, d_cell_tile_start, d_tile_cell, (int)n_cells, *params, run_mask,
                    cell0, d_alt_allele, (unsigned long long*)d_o_fam, (unsigned long long*)d_o_tot, (unsigned long long*)d_o_var);
    }
    return PMRD_OK;
}

// ---------------------------------------------------------------- K7 group table --------
__global__ void __launch_bounds__(128)
group_kernel(const uint64_t* __restrict__ fam_key, const uint32_t* __restrict__ fam_size, const uint32_t* __restrict__ fam_nalt,
             const uint32_t* __restrict__ fam_cons, const uint32_t* __restrict__ fam_flags,
             const uint32_t* __restrict__ grp_start, const uint32_t* __restrict__ d_ngroups, int64_t grp_cap, int weighted,
             pmrd_group_table G) {
    __shared__ unsigned long long s[4][8];  // [warp][fc, alt, tot, c0..c3]
    int64_t n_groups = (int64_t)*d_ngroups;
    if (n_groups > grp_cap) n_groups = grp_cap;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const uint32_t lo = grp_start[g], hi = grp_start[g + 1];
        unsigned long long fc = 0, alt = 0, tot = 0, c[4] = {0, 0, 0, 0};
        for (uint32_t f = lo + threadIdx.x; f < hi; f += blockDim.x) {
            const uint32_t fl = fam_flags[f];
            if (fl & PMRD_FAM_KEPT) {
                ++fc;
                alt += fam_nalt[f];
                tot += fam_size[f];
                if (weighted && (fl & PMRD_FAM_HAS_CONSENSUS)) {
                    const uint32_t a = fam_cons[f] & 3u;
                    c[0] += a == 0; c[1] += a == 1; c[2] += a == 2; c[3] += a == 3;
                }
            }
        }
        fc = warp_sum(fc); alt = warp_sum(alt); tot = warp_sum(tot);
#pragma unroll
        for (int a = 0; a < 4; ++a) c[a] = warp_sum(c[a]);
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) {
            s[w][0] = fc; s[w][1] = alt; s[w][2] = tot;
            s[w][3] = c[0]; s[w][4] = c[1]; s[w][5] = c[2]; s[w][6] = c[3];
        }
        __syncthreads();
        if (threadIdx.x < 7) {
            unsigned long long v = 0;
            for (int ww = 0; ww < 4; ++ww) v += s[ww][threadIdx.x];
            if (threadIdx.x == 0) { G.family_count[g] = (uint32_t)v; G.group[g] = (uint32_t)(fam_key[lo] >> 32); }
            else if (threadIdx.x == 1) G.alt_count[g] = v;
            else if (threadIdx.x == 2) G.total_count[g] = v;
            else G.consensus_count[g * 4 + (threadIdx.x - 3)] = (uint32_t)v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- K6 first-fit ----------
__device__ __forceinline__ int umi_hamming(uint32_t a, uint32_t b) {
    const uint32_t x = a ^ b;
    return __popc((x | (x >> 1)) & 0x55555555u);
}

__global__ void __launch_bounds__(256) hamming_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n, int32_t* __restrict__ d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = umi_hamming(a[i], b[i]);
}

// One warp per group.  Reads are visited in input order; a read joins the FIRST cluster
// whose representative (the UMI that opened it) is within max_distance, else opens a new
// one (mrd/umi.py:29-40).  reps[] (scratch, n entries) holds each group's representatives
// in its own [lo,hi) slice.  Writes new sort keys group:32 | cluster index.
__global__ void __launch_bounds__(128)
firstfit_kernel(const uint64_t* __restrict__ keys_sorted_by_group, const uint32_t* __restrict__ idx,
                const uint64_t* __restrict__ keys_orig, const uint32_t* __restrict__ grp_start, int64_t n_groups,
                int max_distance, uint32_t* __restrict__ reps, uint64_t* __restrict__ keys_out) {
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= n_groups) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lo = grp_start[g], hi = grp_start[g + 1];
    const uint64_t ghi = keys_sorted_by_group[lo] & 0xffffffff00000000ull;
    uint32_t n_rep = 0;
    for (uint32_t r = lo; r < hi; ++r) {
        const uint32_t umi = (uint32_t)keys_orig[idx[r]];
        uint32_t found = 0xffffffffu;
        for (uint32_t c0 = 0; c0 < n_rep; c0 += 32) {
            const uint32_t c = c0 + lane;
            const bool hit = c < n_rep && umi_hamming(umi, reps[lo + c]) <= max_distance;
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                found = c0 + (uint32_t)__ffs((int)m) - 1u;
                break;
            }
        }
        if (found == 0xffffffffu) {
            found = n_rep;
            if (lane == 0) reps[lo + n_rep] = umi;
            ++n_rep;
            __syncwarp();
        }
        if (lane == 0) keys_out[r] = ghi | found;
    }
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) groupkey_kernel(const uint64_t* __restrict__ k, uint64_t* __restrict__ o, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = k[i] & 0xffffffff00000000ull;
}

int32_t k6_hamming(pmrd_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, int64_t n, int32_t* d_d) {
    if (n <= 0) return PMRD_OK;
    PMRD_LAUNCH(ctx, hamming_kernel, pmrd_blocks(n, 256), 256, 0, d_a, d_b, n, d_d);
    return PMRD_OK;
}

// ---------------------------------------------------------------- K6 greedy --------------
// Deterministic greedy Hamming clustering (src/precise_mrd/umi_clustering.py:59-101 cluster_umis_greedy, the
// defined-order form of the loop at src/precise_mrd/umi.py:117-144): per group the UNIQUE UMIs are ordered by
// (-count, umi); the first unused one seeds a cluster and absorbs every unused UMI within max_distance of the
// seed (no chaining through absorbed UMIs).  Clusters are numbered in seed order.
#define GREEDY_NONE 0xffffffffu
// unique-UMI table of the (group, umi)-sorted reads: sort key group:32 | (2^32 - 1 - count), value = unique index.
// The uniques are already in (group, umi) order, so a STABLE sort on this key yields (group, -count, umi).
__global__ void __launch_bounds__(256)
greedy_uniq_kernel(const uint64_t* __restrict__ sk, const uint32_t* __restrict__ useg, int64_t n_uniq, uint64_t* __restrict__ ukey,
                   uint32_t* __restrict__ uidx) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_uniq) return;
    const uint32_t count = useg[u + 1] - useg[u];
    ukey[u] = (sk[useg[u]] & 0xffffffff00000000ull) | (uint64_t)(0xffffffffu - count);
    uidx[u] = (uint32_t)u;
}
// UMIs in seed order (coalesced for the clustering loop) + cluster ids cleared
__global__ void __launch_bounds__(256)
greedy_gather_kernel(const uint64_t* __restrict__ sk, const uint32_t* __restrict__ useg, const uint32_t* __restrict__ ord, int64_t n_uniq,
                     uint32_t* __restrict__ sumi, uint32_t* __restrict__ scl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_uniq) return;
    sumi[i] = (uint32_t)sk[useg[ord[i]]];
    scl[i] = GREEDY_NONE;
}
// one block per group: seeds in order (sequential by nature), absorption spread over the threads
__global__ void __launch_bounds__(256)
greedy_cluster_kernel(const uint32_t* __restrict__ sumi, uint32_t* __restrict__ scl, const uint32_t* __restrict__ gstart, int64_t n_groups,
                      int max_distance) {
    const int64_t g = blockIdx.x;
    if (g >= n_groups) return;
    const uint32_t lo = gstart[g], hi = gstart[g + 1];
    uint32_t n_clusters = 0;
    for (uint32_t i = lo; i < hi; ++i) {
        // every entry before i is used (it was a seed or was absorbed), so only [i, hi) can still join
        if (((volatile uint32_t*)scl)[i] != GREEDY_NONE) continue;          // uniform: written before the last barrier
        const uint32_t seed = sumi[i];
        for (uint32_t j = i + threadIdx.x; j < hi; j += blockDim.x)
            if (scl[j] == GREEDY_NONE && umi_hamming(seed, sumi[j]) <= max_distance) scl[j] = n_clusters;
        ++n_clusters;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256)
greedy_scatter_kernel(const uint32_t* __restrict__ ord, const uint32_t* __restrict__ scl, int64_t n_uniq, uint32_t* __restrict__ ucluster) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_uniq) ucluster[ord[i]] = scl[i];
}
// every read, back at its ORIGINAL position, gets the key group:32 | cluster:32 -- a stable sort of those keys then
// lists each cluster's reads in input order (umi.py:137-140)
__global__ void __launch_bounds__(256)
greedy_assign_kernel(const uint64_t* __restrict__ sk, const uint32_t* __restrict__ si, int64_t n, const uint32_t* __restrict__ useg,
                     int64_t n_uniq, const uint32_t* __restrict__ ucluster, uint64_t* __restrict__ keys_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int64_t lo = 0, hi = n_uniq;                   // useg[lo] <= r < useg[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)useg[mid] <= r) lo = mid; else hi = mid;
    }
    keys_out[si[r]] = (sk[r] & 0xffffffff00000000ull) | ucluster[lo];
}

// ---------------------------------------------------------------- external reads: group-major fast path ----
// pmrd_collapse_reads on reads in ARBITRARY order (UMIProcessor.process_reads src/precise_mrd/umi.py:205-285,
// group_umis_fast rust_ext/src/lib.rs:7-109, mrd/umi.py:57-115 EXACT clustering) when the group ids are dense
// (< 2^22) and no group holds more than 4096 reads -- a panel of loci, the shape of BASELINE.json configs[4].
// The full-key sort (6 passes of 12 + 12 B for a 17-bit group and a 24-bit UMI) becomes
//   count      reads per group id (one pass over the keys, L2 atomics)             -> eligibility + group starts
//   partition  stable LSD radix passes over the GROUP bits only (2-3 passes)        -> reads group-major, input order kept
//   group_rows one block per group: loads the group, sorts it on the UMI bits inside shared memory (ranked passes of the
//              tile kernels, stable), finds the family heads and votes every family -> family rows, written where
//              the group's reads start (a group has at most as many families as reads)
//   compact    scan of the per-group family counts, rows moved to their final place -> the family table, sorted by key
// The consensus arithmetic is family_row(), the one family_kernel runs.
// group boundaries of the group-sorted keys: start[g] = first read of group g, cnt[g] = its read count (both zeroed
// beforehand: ids without reads keep 0).  One coalesced pass over the keys -- the per-read L2 atomics of a histogram over
// 1e5 ids cost four times as much (0.81 vs 0.2 ms per 1.25e8 reads).
__global__ void __launch_bounds__(256)
grp_bounds_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ start, uint32_t* __restrict__ cnt) {
    // 4 keys per thread and step (two 16-byte loads): the kernel only reads and is paced by the bytes it keeps in
    // flight (one key per thread: 0.41 ms for 1 GB; this form: 0.19 ms).  A key's neighbours come from the thread's own
    // registers or the next lanes' by shuffle; only the keys around a warp's 128-key span are read twice (L1 hits).
    const int64_t span = (int64_t)gridDim.x * blockDim.x * 4;
    const int lane = threadIdx.x & 31;
    const bool aligned = (reinterpret_cast<uintptr_t>(keys) & 15u) == 0u;       // (a caller's device array may start anywhere)
    for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 - lane * 4 < n; i0 += span) {
        uint32_t g[4];
        if (aligned && i0 + 3 < n) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(keys + i0), b = *reinterpret_cast<const ulonglong2*>(keys + i0 + 2);
            g[0] = (uint32_t)(a.x >> 32); g[1] = (uint32_t)(a.y >> 32); g[2] = (uint32_t)(b.x >> 32); g[3] = (uint32_t)(b.y >> 32);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = i0 + j < n ? (uint32_t)(keys[i0 + j] >> 32) : 0xffffffffu;
        }
        uint32_t prev = __shfl_up_sync(0xffffffffu, g[3], 1), next = __shfl_down_sync(0xffffffffu, g[0], 1);
        if (lane == 0) prev = i0 > 0 && i0 - 1 < n ? (uint32_t)(keys[i0 - 1] >> 32) : 0xffffffffu;
        if (lane == 31) next = i0 + 4 < n ? (uint32_t)(keys[i0 + 4] >> 32) : 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = i0 + j;
            if (i >= n) break;
            const uint32_t p = j ? g[j - 1] : prev, q = j < 3 ? g[j + 1] : next;
            if (i == 0 || p != g[j]) start[g[j]] = (uint32_t)i;
            if (i == n - 1 || q != g[j]) cnt[g[j]] = (uint32_t)(i + 1);
        }
    }
}
// end -> count, largest group, non-empty flag per id (scanned into the group-table row of every id)
__global__ void __launch_bounds__(256)
grp_max_kernel(const uint32_t* __restrict__ start, uint32_t* __restrict__ cnt, int64_t n_ids, uint32_t* __restrict__ nonempty,
               uint32_t* __restrict__ out /*[2]: max size, groups with reads*/) {
    uint32_t mx = 0, ne = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ids; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t e = cnt[i];
        const uint32_t c = e ? e - start[i] : 0u;
        cnt[i] = c;
        nonempty[i] = c > 0 ? 1u : 0u;
        mx = max(mx, c);
        ne += c > 0;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    ne = warp_sum(ne);
    if ((threadIdx.x & 31) == 0) { atomicMax(&out[0], mx); atomicAdd(&out[1], ne); }
}

#define GR_THREADS 256
// ROWS = true: family rows (tmp, then compacted) for a caller that wants the family table.  ROWS = false: the group's row
// of the GROUP table is reduced right here (umi.py:241-266, mrd/umi.py:86-99) -- no family row is ever written.
template <int ITEMS, int MODE, bool ROWS>
__global__ void __launch_bounds__(GR_THREADS)
group_rows_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ payload, const uint32_t* __restrict__ grp_start,
                  const uint32_t* __restrict__ grp_cnt, int n_rounds, pmrd_umi_params P, FamOut tmp, uint32_t* __restrict__ fam_cnt,
                  const uint32_t* __restrict__ grp_row /* ROWS = false: row of every id in the group table */, pmrd_group_table G,
                  unsigned long long* __restrict__ n_fam_total) {
    constexpr int CAP = GR_THREADS * ITEMS, WARPS = GR_THREADS / 32, PER_WARP = CAP / WARPS;
    extern __shared__ __align__(16) unsigned char gr_raw[];
    uint64_t* const s_rec = reinterpret_cast<uint64_t*>(gr_raw);                 // payload:32 | umi:32
    uint32_t (*s_cnt)[256] = reinterpret_cast<uint32_t (*)[256]>(s_rec + CAP);
    uint32_t* const s_base = &s_cnt[0][0] + WARPS * 256;
    uint32_t* const sw = s_base + 256;                                           // [33]
    uint16_t* const s_list = reinterpret_cast<uint16_t*>(sw + 36);                // [CAP]
    __shared__ uint32_t s_nheads;
    __shared__ double s_wt[64];
    const uint32_t g = blockIdx.x;
    const int n = (int)grp_cnt[g];
    if (n == 0) { if (ROWS && threadIdx.x == 0) fam_cnt[g] = 0u; return; }
    if (threadIdx.x < 64) s_wt[threadIdx.x] = c_wtab[threadIdx.x];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    const uint32_t lo = grp_start[g];
    uint64_t key[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int loc = (int)warp * PER_WARP + i * 32 + (int)lane;
        // filler: UMI word 0xFFFFFFFF sorts last; a REAL UMI of that value sits before the fillers and stays there (stable)
        key[i] = loc < n ? (((uint64_t)payload[lo + loc] << 32) | (uint32_t)keys[lo + loc]) : 0xffffffffull;
    }
    if (threadIdx.x == 0) s_nheads = 0u;
    // ROWS = false (only the group's row is wanted, and a row does not depend on the order of the families): from three
    // rounds up the LAST ranked round is not made.  A run of equal sorted bits is then one family but for the ~1 % of runs
    // in which two UMIs share their low 16 (24) bits; the thread that owns such a run puts it in order itself (stable
    // insertion sort of a handful of records) before it votes.
    const bool partial = !ROWS && n_rounds >= 3 && n_rounds < 16;     // (n_rounds + 16: every round is made -- PMRD_GROUP_PARTIAL=0)
    n_rounds &= 15;
    if (partial) --n_rounds;
    const uint32_t run_mask = partial ? (1u << (8 * n_rounds)) - 1u : 0xffffffffu;
#pragma unroll 1
    for (int rd = 0; rd < n_rounds; ++rd) {
        const int shift = rd * 8;
        for (int i = threadIdx.x; i < WARPS * 256; i += GR_THREADS) (&s_cnt[0][0])[i] = 0u;
        __syncthreads();
        uint32_t rank2[(ITEMS + 1) / 2];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = ((uint32_t)key[i] >> shift) & 0xffu;
            const uint32_t m = rs_peer_mask(d, 0xffffffffu);
            const uint32_t before = __popc(m & lt);
            const int leader = __ffs((int)m) - 1;
            uint32_t old = 0;
            if (before == 0) old = atomicAdd(&s_cnt[warp][d], (uint32_t)__popc(m));
            old = __shfl_sync(0xffffffffu, old, leader);
            const uint32_t r = old + before;
            if (i & 1) rank2[i >> 1] |= r << 16;
            else rank2[i >> 1] = r;
        }
        __syncthreads();
        {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = s_cnt[w][threadIdx.x];
                s_cnt[w][threadIdx.x] = run;
                run += c;
            }
            uint32_t tot;
            s_base[threadIdx.x] = block_exclusive_scan<uint32_t>(run, sw, tot);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = ((uint32_t)key[i] >> shift) & 0xffu;
            const uint32_t r = (i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu;
            s_rec[s_base[d] + s_cnt[warp][d] + r] = key[i];
        }
        __syncthreads();
        if (rd + 1 < n_rounds) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = s_rec[warp * PER_WARP + i * 32 + lane];
        }
    }
    if (n_rounds == 0) {                          // one UMI in the whole input: nothing to order
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) s_rec[warp * PER_WARP + i * 32 + lane] = key[i];
        __syncthreads();
    }
    // family heads, compacted in order: warp w owns the sorted records [w * PER_WARP, (w + 1) * PER_WARP); the head masks
    // of its ITEMS 32-record steps stay in registers while the eight warp totals are prefixed
    uint32_t hm[ITEMS];
    uint32_t n_mine = 0;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int i = (int)warp * PER_WARP + it * 32 + (int)lane;
        const bool head = i < n && (i == 0 || (((uint32_t)s_rec[i - 1] ^ (uint32_t)s_rec[i]) & run_mask) != 0u);
        hm[it] = __ballot_sync(0xffffffffu, head);
        n_mine += __popc(hm[it]);
    }
    if (lane == 0) sw[warp] = n_mine;
    __syncthreads();
    uint32_t hbase = 0, n_all = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const uint32_t c = sw[w];
        hbase += w < (int)warp ? c : 0u;
        n_all += c;
    }
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        if ((hm[it] >> lane) & 1u) s_list[hbase + __popc(hm[it] & lt)] = (uint16_t)((int)warp * PER_WARP + it * 32 + (int)lane);
        hbase += __popc(hm[it]);
    }
    if (threadIdx.x == 0) s_nheads = n_all;
    __syncthreads();
    const int n_heads = (int)s_nheads;
    const uint64_t ghi = keys[lo] & 0xffffffff00000000ull;
    uint32_t a_fc = 0, a_alt = 0, a_tot = 0, a_c0 = 0, a_c1 = 0, a_c2 = 0, a_c3 = 0;      // ROWS = false: this thread's share of the group row
    uint32_t a_nfam = 0;
    for (int h = threadIdx.x; h < n_heads; h += GR_THREADS) {
        int i = (int)s_list[h];
        const int run_end = h + 1 < n_heads ? (int)s_list[h + 1] : n;
        int end = run_end;
        if (partial) {
            // the first family of the run; if it is not the whole run, order the run on the full UMI and take its families in turn
            const uint32_t u0 = (uint32_t)s_rec[i];
            int f1 = i + 1;
            while (f1 < run_end && (uint32_t)s_rec[f1] == u0) ++f1;
            if (f1 < run_end) {
                for (int a = i + 1; a < run_end; ++a) {
                    const uint64_t rec = s_rec[a];
                    int b = a;
                    while (b > i && (uint32_t)s_rec[b - 1] > (uint32_t)rec) { s_rec[b] = s_rec[b - 1]; --b; }
                    s_rec[b] = rec;
                }
                f1 = i + 1;
                const uint32_t u1 = (uint32_t)s_rec[i];
                while (f1 < run_end && (uint32_t)s_rec[f1] == u1) ++f1;
            }
            end = f1;
        }
    next_family:
        ++a_nfam;
        uint32_t n_alt, consensus, qsum_best, n_best, flags;
        family_row<MODE>((uint32_t)(end - i), [&](uint32_t r) -> uint32_t { return (uint32_t)(s_rec[i + r] >> 32); }, P, s_wt, n_alt,
                         consensus, qsum_best, n_best, flags);
        if (ROWS) {
            const size_t o = (size_t)lo + h;
            tmp.key[o] = ghi | (uint32_t)s_rec[i];
            tmp.size[o] = (uint32_t)(end - i);
            tmp.n_alt[o] = n_alt;
            tmp.consensus[o] = consensus;
            tmp.qsum[o] = qsum_best;
            tmp.n_best[o] = n_best;
            tmp.flags[o] = flags;
        } else if (flags & PMRD_FAM_KEPT) {                      // exactly what group_kernel sums over the family table
            ++a_fc;
            a_alt += n_alt;
            a_tot += (uint32_t)(end - i);
            if (MODE == PMRD_MODE_WEIGHTED && (flags & PMRD_FAM_HAS_CONSENSUS)) {
                const uint32_t a = consensus & 3u;
                a_c0 += a == 0; a_c1 += a == 1; a_c2 += a == 2; a_c3 += a == 3;
            }
        }
        if (partial && end < run_end) {                          // the run's next family
            i = end;
            const uint32_t u = (uint32_t)s_rec[i];
            end = i + 1;
            while (end < run_end && (uint32_t)s_rec[end] == u) ++end;
            goto next_family;
        }
    }
    if (ROWS) {
        if (threadIdx.x == 0) fam_cnt[g] = (uint32_t)n_heads;
        return;
    }
    __shared__ uint32_t s_red[8];
    if (threadIdx.x < 8) s_red[threadIdx.x] = 0u;
    __syncthreads();
    a_fc = warp_sum(a_fc); a_alt = warp_sum(a_alt); a_tot = warp_sum(a_tot);
    a_c0 = warp_sum(a_c0); a_c1 = warp_sum(a_c1); a_c2 = warp_sum(a_c2); a_c3 = warp_sum(a_c3);
    a_nfam = warp_sum(a_nfam);
    if (lane == 0) {
        atomicAdd(&s_red[0], a_fc); atomicAdd(&s_red[1], a_alt); atomicAdd(&s_red[2], a_tot);
        atomicAdd(&s_red[3], a_c0); atomicAdd(&s_red[4], a_c1); atomicAdd(&s_red[5], a_c2); atomicAdd(&s_red[6], a_c3);
        atomicAdd(&s_red[7], a_nfam);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t row = grp_row[g];
        G.group[row] = g;
        G.family_count[row] = s_red[0];
        G.alt_count[row] = s_red[1];
        G.total_count[row] = s_red[2];
        G.consensus_count[row * 4 + 0] = s_red[3]; G.consensus_count[row * 4 + 1] = s_red[4];
        G.consensus_count[row * 4 + 2] = s_red[5]; G.consensus_count[row * 4 + 3] = s_red[6];
        atomicAdd(n_fam_total, (unsigned long long)s_red[7]);
    }
}

// rows of group g: tmp[grp_start[g] .. + fam_cnt[g])  ->  out[fam_base[g] ..)
__global__ void __launch_bounds__(128)
group_rows_compact_kernel(const uint32_t* __restrict__ grp_start, const uint32_t* __restrict__ fam_cnt, const uint32_t* __restrict__ fam_base,
                          int64_t fam_cap, FamOut tmp, FamOut out) {
    const uint32_t g = blockIdx.x;
    const uint32_t n = fam_cnt[g], src = grp_start[g], dst = fam_base[g];
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        if ((int64_t)(dst + k) >= fam_cap) break;
        out.key[dst + k] = tmp.key[src + k];
        out.size[dst + k] = tmp.size[src + k];
        out.n_alt[dst + k] = tmp.n_alt[src + k];
        out.consensus[dst + k] = tmp.consensus[src + k];
        out.qsum[dst + k] = tmp.qsum[src + k];
        out.n_best[dst + k] = tmp.n_best[src + k];
        out.flags[dst + k] = tmp.flags[src + k];
    }
}

static size_t group_rows_smem(int items) {
    const size_t cap = (size_t)GR_THREADS * items;
    return cap * 8 + (GR_THREADS / 32) * 256 * 4 + 256 * 4 + 36 * 4 + cap * 2 + 16;
}

template <int MODE, bool ROWS>
static int32_t launch_group_rows(pmrd_ctx* ctx, int64_t max_group, int64_t n_ids, const uint64_t* keys, const uint32_t* payload,
                                 const uint32_t* grp_start, const uint32_t* grp_cnt, int n_rounds, const pmrd_umi_params& P, FamOut tmp,
                                 uint32_t* fam_cnt, const uint32_t* grp_row, const pmrd_group_table& G, unsigned long long* n_fam_total) {
#define GR_CASE(I)                                                                                                                      \
    if (max_group <= (int64_t)GR_THREADS * (I)) {                                                                                       \
        const size_t smem = group_rows_smem(I);                                                                                         \
        PMRD_CUDA(ctx, cudaFuncSetAttribute(group_rows_kernel<I, MODE, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        PMRD_LAUNCH_NAMED(ctx, "group_rows_kernel", (group_rows_kernel<I, MODE, ROWS>), (int)n_ids, GR_THREADS, smem, keys, payload,     \
                          grp_start, grp_cnt, n_rounds, P, tmp, fam_cnt, grp_row, G, n_fam_total);                                        \
        return PMRD_OK;                                                                                                                 \
    }
    GR_CASE(1) GR_CASE(2) GR_CASE(4) GR_CASE(6) GR_CASE(8) GR_CASE(12) GR_CASE(16)
#undef GR_CASE
    return PMRD_ERR_ARG;
}

template <bool ROWS>
static int32_t launch_group_rows_mode(pmrd_ctx* ctx, int mode, int64_t max_group, int64_t n_ids, const uint64_t* keys, const uint32_t* payload,
                                      const uint32_t* grp_start, const uint32_t* grp_cnt, int n_rounds, const pmrd_umi_params& P, FamOut tmp,
                                      uint32_t* fam_cnt, const uint32_t* grp_row, const pmrd_group_table& G, unsigned long long* n_fam_total) {
    if (mode == PMRD_MODE_WEIGHTED)
        return launch_group_rows<PMRD_MODE_WEIGHTED, ROWS>(ctx, max_group, n_ids, keys, payload, grp_start, grp_cnt, n_rounds, P, tmp, fam_cnt, grp_row, G, n_fam_total);
    if (mode == PMRD_MODE_MAJORITY)
        return launch_group_rows<PMRD_MODE_MAJORITY, ROWS>(ctx, max_group, n_ids, keys, payload, grp_start, grp_cnt, n_rounds, P, tmp, fam_cnt, grp_row, G, n_fam_total);
    return launch_group_rows<PMRD_MODE_COUNT, ROWS>(ctx, max_group, n_ids, keys, payload, grp_start, grp_cnt, n_rounds, P, tmp, fam_cnt, grp_row, G, n_fam_total);
}

// ---------------------------------------------------------------- driver ----------------
template <bool GATHER>
static int32_t launch_family(pmrd_ctx* ctx, int mode, const uint64_t* keys, const uint32_t* pl_or_idx,
                             const uint64_t* keys_orig, const uint32_t* payload_orig, const uint32_t* fam_start,
                             const uint32_t* d_nfam, int64_t fam_cap, const pmrd_umi_params& P, FamOut out,
                             uint32_t sentinel_umi = 0u) {
    const int grid = pmrd_blocks(fam_cap, 128);
    if (mode == PMRD_MODE_WEIGHTED) {
        PMRD_LAUNCH(ctx, (family_kernel<PMRD_MODE_WEIGHTED, GATHER>), grid, 128, 0, keys, pl_or_idx, keys_orig, payload_orig, fam_start, d_nfam, fam_cap, P, sentinel_umi, out);
    } else if (mode == PMRD_MODE_MAJORITY) {
        PMRD_LAUNCH(ctx, (family_kernel<PMRD_MODE_MAJORITY, GATHER>), grid, 128, 0, keys, pl_or_idx, keys_orig, payload_orig, fam_start, d_nfam, fam_cap, P, sentinel_umi, out);
    } else {
        PMRD_LAUNCH(ctx, (family_kernel<PMRD_MODE_COUNT, GATHER>), grid, 128, 0, keys, pl_or_idx, keys_orig, payload_orig, fam_start, d_nfam, fam_cap, P, sentinel_umi, out);
    }
    return PMRD_OK;
}

// Sync-free core.  `varying` = mask of key bits that differ (0: find out, costs one sync).
// d_counts[0] = families, d_counts[1] = groups (device).  EXACT clustering only.
int32_t k5_collapse_exact_nosync(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp,
                                 uint32_t* d_payload_tmp, int64_t n_reads, uint64_t varying, const pmrd_umi_params* params,
                                 const pmrd_family_table* d_fam, int64_t capacity_f, const pmrd_group_table* d_grp,
                                 int64_t capacity_g, uint32_t* d_fam_start /*[capacity_f+1]*/,
                                 uint32_t* d_grp_start /*[capacity_g+1]*/, uint32_t* d_counts,
                                 const int64_t* d_cell_tile_start /* non-null: cell-major, tile-padded input */,
                                 int64_t n_seg_cells, int seg_key_bits) {
    const pmrd_umi_params P = *params;
    int in_b = 0;
    if (d_cell_tile_start) {
        // every cell is sorted on its own over the UMI bits only: 3 passes for a 12-mer UMI
        PMRD_TRY(radix_sort_pairs_segmented(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n_reads, d_cell_tile_start,
                                            n_seg_cells, seg_key_bits, &in_b));
    } else {
        if (varying == 0) PMRD_TRY(radix_varying_bits(ctx, d_keys, n_reads, &varying));
        PMRD_TRY(radix_sort_pairs(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n_reads, varying, &in_b));
    }
    const uint64_t* sk = in_b ? d_keys_tmp : d_keys;
    const uint32_t* sp = in_b ? d_payload_tmp : d_payload;
    PMRD_TRY(find_segments_dev(ctx, sk, n_reads, nullptr, 0, capacity_f, d_fam_start, d_counts));
    FamOut out{d_fam->key, d_fam->size, d_fam->n_alt, d_fam->consensus, d_fam->qsum, d_fam->n_best, d_fam->flags};
    PMRD_TRY(launch_family<false>(ctx, P.mode, sk, sp, nullptr, nullptr, d_fam_start, d_counts, capacity_f, P, out,
                                  d_cell_tile_start ? 0xffffffffu : 0u));
    if (d_grp) {
        PMRD_TRY(find_segments_dev(ctx, d_fam->key, capacity_f, d_counts, 32, capacity_g, d_grp_start, d_counts + 1));
        int grid = (int)(capacity_g < (int64_t)ctx->sm_count * 32 ? capacity_g : (int64_t)ctx->sm_count * 32);
        if (grid < 1) grid = 1;
        PMRD_LAUNCH(ctx, group_kernel, grid, 128, 0, (const uint64_t*)d_fam->key, (const uint32_t*)d_fam->size,
                    (const uint32_t*)d_fam->n_alt, (const uint32_t*)d_fam->consensus, (const uint32_t*)d_fam->flags,
                    (const uint32_t*)d_grp_start, (const uint32_t*)(d_counts + 1), capacity_g,
                    P.mode == PMRD_MODE_WEIGHTED ? 1 : 0, *d_grp);
    }
    return PMRD_OK;
}

// EXACT clustering, external reads: the group-major fast path (see above).  *done = false: not eligible (sparse or
// huge group ids, a group of more than 4096 reads): nothing was changed, the caller runs the full-key sort.
#define GR_MAX_IDS (1ll << 22)
#define GR_MAX_GROUP 4096
static int32_t k5_collapse_groups_fast(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp, uint32_t* d_payload_tmp,
                                       int64_t n_reads, const pmrd_umi_params& P, const pmrd_family_table* d_fam, int64_t fcap,
                                       const pmrd_group_table* d_grp, int64_t gcap, uint32_t* d_grp_start, uint32_t* d_counts,
                                       bool* done, bool* in_tmp, int64_t* nf_out, int64_t* ng_out) {
    *done = false;
    *in_tmp = false;
    static_assert(GR_MAX_GROUP == GR_THREADS * 16, "the largest group_rows_kernel instance holds GR_MAX_GROUP reads");
    const char* env = getenv("PMRD_GROUP_FAST");
    if (env && atoi(env) == 0) return PMRD_OK;
    uint64_t varying = 0, all_or = 0;
    PMRD_TRY(radix_varying_bits(ctx, d_keys, n_reads, &varying, &all_or));
    const int64_t n_ids = (int64_t)(all_or >> 32) + 1;
    if (n_ids > GR_MAX_IDS || n_ids > n_reads + 1024) return PMRD_OK;
    // stable partition by group (LSD radix over the group bits only: reads keep their input order inside a group)
    int in_b = 0;
    const char* e9 = getenv("PMRD_PARTITION_BITS");            // 8: 8-bit digits only
    PMRD_TRY(radix_sort_pairs(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n_reads, varying & 0xffffffff00000000ull, &in_b, nullptr,
                              e9 && atoi(e9) == 8 ? 8 : 9));
    *in_tmp = in_b != 0;                 // (whatever happens next, this is where the reads are now)
    const uint64_t* sk = in_b ? d_keys_tmp : d_keys;
    const uint32_t* sp = in_b ? d_payload_tmp : d_payload;
    DevBuf cnt, start, mx, fcnt, nonempty, grow;
    PMRD_CUDA(ctx, cnt.alloc(ctx->stream, (size_t)(n_ids + 1) * 4));
    PMRD_CUDA(ctx, start.alloc(ctx->stream, (size_t)(n_ids + 1) * 4));
    PMRD_CUDA(ctx, nonempty.alloc(ctx->stream, (size_t)(n_ids + 1) * 4));
    PMRD_CUDA(ctx, grow.alloc(ctx->stream, (size_t)(n_ids + 1) * 4));
    PMRD_CUDA(ctx, mx.alloc(ctx->stream, 8));
    PMRD_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, (size_t)(n_ids + 1) * 4, ctx->stream));
    PMRD_CUDA(ctx, cudaMemsetAsync(start.p, 0, (size_t)(n_ids + 1) * 4, ctx->stream));
    PMRD_CUDA(ctx, cudaMemsetAsync(mx.p, 0, 8, ctx->stream));
    int grid = (int)((n_reads + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    PMRD_LAUNCH(ctx, grp_bounds_kernel, grid, 256, 0, sk, n_reads, start.as<uint32_t>(), cnt.as<uint32_t>());
    int g2 = (int)((n_ids + 255) / 256);
    if (g2 > ctx->sm_count * 8) g2 = ctx->sm_count * 8;
    PMRD_LAUNCH(ctx, grp_max_kernel, g2, 256, 0, (const uint32_t*)start.as<uint32_t>(), cnt.as<uint32_t>(), n_ids, nonempty.as<uint32_t>(), mx.as<uint32_t>());
    uint32_t h_mx[2] = {0, 0};
    PMRD_CUDA(ctx, cudaMemcpyAsync(h_mx, mx.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // a group beyond one block's capacity: the caller goes on with the full-key sort, FROM the partitioned arrays (a stable
    // sort of them gives what it gives of the original order: the partition kept the input order inside every group)
    if (h_mx[0] > GR_MAX_GROUP) return PMRD_OK;
    *done = true;
    int ubits = 0;
    while (ubits < 32 && (varying & 0xffffffffull) >> ubits) ++ubits;
    int n_rounds = (ubits + 7) / 8;
    const bool rows = d_fam && d_fam->key;
    const char* gp = getenv("PMRD_GROUP_PARTIAL");              // 0: the counts-only form makes every ranked round too
    if (gp && atoi(gp) == 0) n_rounds += 16;
    if (!rows) {
        // group table only: every block reduces its group's row; row of an id = number of non-empty ids before it
        PMRD_ARG(ctx, d_grp != nullptr, "neither a family table nor a group table was asked for");
        *ng_out = (int64_t)h_mx[1];
        if (*ng_out > gcap) { *nf_out = 0; return PMRD_OK; }             // the caller reports the capacity error
        PMRD_TRY((exclusive_scan<uint32_t, uint32_t>(ctx, nonempty.as<uint32_t>(), grow.as<uint32_t>(), n_ids, (uint32_t*)nullptr)));
        DevBuf tot;
        PMRD_CUDA(ctx, tot.alloc(ctx->stream, 8));
        PMRD_CUDA(ctx, cudaMemsetAsync(tot.p, 0, 8, ctx->stream));
        PMRD_TRY(launch_group_rows_mode<false>(ctx, P.mode, h_mx[0], n_ids, sk, sp, start.as<uint32_t>(), cnt.as<uint32_t>(), n_rounds, P, FamOut{},
                                               nullptr, grow.as<uint32_t>(), *d_grp, tot.as<unsigned long long>()));
        unsigned long long h_tot = 0;
        PMRD_CUDA(ctx, cudaMemcpyAsync(&h_tot, tot.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *nf_out = (int64_t)h_tot;
        return PMRD_OK;
    }
    DevBuf tk, ts, ta, tc, tq, tn, tf;
    const size_t n = (size_t)n_reads;
    PMRD_CUDA(ctx, tk.alloc(ctx->stream, n * 8)); PMRD_CUDA(ctx, ts.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, ta.alloc(ctx->stream, n * 4));
    PMRD_CUDA(ctx, tc.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, tq.alloc(ctx->stream, n * 4)); PMRD_CUDA(ctx, tn.alloc(ctx->stream, n * 4));
    PMRD_CUDA(ctx, tf.alloc(ctx->stream, n * 4));
    PMRD_CUDA(ctx, fcnt.alloc(ctx->stream, (size_t)(n_ids + 1) * 4));
    FamOut tmp{tk.as<uint64_t>(), ts.as<uint32_t>(), ta.as<uint32_t>(), tc.as<uint32_t>(), tq.as<uint32_t>(), tn.as<uint32_t>(), tf.as<uint32_t>()};
    PMRD_TRY(launch_group_rows_mode<true>(ctx, P.mode, h_mx[0], n_ids, sk, sp, start.as<uint32_t>(), cnt.as<uint32_t>(), n_rounds, P, tmp,
                                          fcnt.as<uint32_t>(), nullptr, pmrd_group_table{}, nullptr));
    // family counts -> row offsets; the total is the family count of the call
    DevBuf fbase;
    PMRD_CUDA(ctx, fbase.alloc(ctx->stream, (size_t)(n_ids + 1) * 4));
    PMRD_TRY((exclusive_scan<uint32_t, uint32_t>(ctx, fcnt.as<uint32_t>(), fbase.as<uint32_t>(), n_ids, d_counts)));
    FamOut out{d_fam->key, d_fam->size, d_fam->n_alt, d_fam->consensus, d_fam->qsum, d_fam->n_best, d_fam->flags};
    PMRD_LAUNCH(ctx, group_rows_compact_kernel, (int)n_ids, 128, 0, (const uint32_t*)start.as<uint32_t>(), (const uint32_t*)fcnt.as<uint32_t>(),
                (const uint32_t*)fbase.as<uint32_t>(), fcap, tmp, out);
    PMRD_TRY(read_count(ctx, d_counts, nf_out));
    *ng_out = 0;
    if (d_grp && *nf_out <= fcap) {
        PMRD_TRY(find_segments_dev(ctx, d_fam->key, *nf_out, nullptr, 32, gcap, d_grp_start, d_counts + 1));
        int ggrid = (int)(gcap < (int64_t)ctx->sm_count * 32 ? gcap : (int64_t)ctx->sm_count * 32);
        if (ggrid < 1) ggrid = 1;
        PMRD_LAUNCH(ctx, group_kernel, ggrid, 128, 0, (const uint64_t*)d_fam->key, (const uint32_t*)d_fam->size,
                    (const uint32_t*)d_fam->n_alt, (const uint32_t*)d_fam->consensus, (const uint32_t*)d_fam->flags,
                    (const uint32_t*)d_grp_start, (const uint32_t*)(d_counts + 1), gcap, P.mode == PMRD_MODE_WEIGHTED ? 1 : 0, *d_grp);
        PMRD_TRY(read_count(ctx, d_counts + 1, ng_out));
    }
    return PMRD_OK;
}

int32_t k5_collapse(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp, uint32_t* d_payload_tmp,
                    int64_t n_reads, const pmrd_umi_params* params, const pmrd_family_table* d_fam, int64_t capacity_f,
                    int64_t* n_families, const pmrd_group_table* d_grp, int64_t capacity_g, int64_t* n_groups) {
    PMRD_ARG(ctx, params != nullptr, "params");
    const pmrd_umi_params P = *params;
    PMRD_ARG(ctx, P.mode >= 0 && P.mode <= 2, "mode");
    PMRD_ARG(ctx, P.cluster == PMRD_CLUSTER_EXACT || P.cluster == PMRD_CLUSTER_FIRSTFIT || P.cluster == PMRD_CLUSTER_GREEDY, "cluster");
    PMRD_ARG(ctx, ((uintptr_t)d_keys & 15) == 0 && ((uintptr_t)d_keys_tmp & 15) == 0, "key buffers must be 16-byte aligned");
    *n_families = 0;
    if (n_groups) *n_groups = 0;
    if (n_reads <= 0) return PMRD_OK;
    PMRD_ARG(ctx, n_reads < (1ll << 31), "n_reads must be < 2^31 per call (chunk the input)");
    const bool want_groups = d_grp && n_groups;
    if (!(d_fam && d_fam->key)) capacity_f = n_reads;            // no family table asked for: no capacity to respect
    const int64_t fcap = capacity_f < n_reads ? capacity_f : n_reads;
    const int64_t gcap = want_groups ? (capacity_g < n_reads ? capacity_g : n_reads) : 0;

    DevBuf fam_start, grp_start, counts;
    PMRD_CUDA(ctx, fam_start.alloc(ctx->stream, (size_t)(fcap + 1) * 4));
    PMRD_CUDA(ctx, grp_start.alloc(ctx->stream, (size_t)(gcap + 1) * 4));
    PMRD_CUDA(ctx, counts.alloc(ctx->stream, 8));
    PMRD_CUDA(ctx, cudaMemsetAsync(counts.p, 0, 8, ctx->stream));
    int64_t nf = 0, ng = 0;

    bool fast_done = false, in_tmp = false;
    const bool want_rows = d_fam && d_fam->key;
    if (P.cluster == PMRD_CLUSTER_EXACT)
        PMRD_TRY(k5_collapse_groups_fast(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n_reads, P, d_fam, fcap, want_groups ? d_grp : nullptr,
                                         gcap, grp_start.as<uint32_t>(), counts.as<uint32_t>(), &fast_done, &in_tmp, &nf, &ng));
    if (in_tmp) {      // the partition left the reads in the tmp buffers: they are the input of whatever follows
        uint64_t* tk = d_keys; d_keys = d_keys_tmp; d_keys_tmp = tk;
        uint32_t* tp = d_payload; d_payload = d_payload_tmp; d_payload_tmp = tp;
    }
    // the other paths build the group table FROM the family table: a caller that did not ask for the family rows gets
    // them computed into scratch all the same
    DevBuf xk, xs, xa, xc, xq, xn, xf;
    pmrd_family_table scratch_fam;
    if (!fast_done && !want_rows) {
        const size_t fn = (size_t)fcap;
        PMRD_CUDA(ctx, xk.alloc(ctx->stream, fn * 8)); PMRD_CUDA(ctx, xs.alloc(ctx->stream, fn * 4)); PMRD_CUDA(ctx, xa.alloc(ctx->stream, fn * 4));
        PMRD_CUDA(ctx, xc.alloc(ctx->stream, fn * 4)); PMRD_CUDA(ctx, xq.alloc(ctx->stream, fn * 4)); PMRD_CUDA(ctx, xn.alloc(ctx->stream, fn * 4));
        PMRD_CUDA(ctx, xf.alloc(ctx->stream, fn * 4));
        scratch_fam = pmrd_family_table{xk.as<uint64_t>(), xs.as<uint32_t>(), xa.as<uint32_t>(), xc.as<uint32_t>(), xq.as<uint32_t>(),
                                        xn.as<uint32_t>(), xf.as<uint32_t>()};
        d_fam = &scratch_fam;
    }
    if (fast_done) {
        // (family table and group table are in place; the counts were read back)
    } else if (P.cluster == PMRD_CLUSTER_EXACT) {
        PMRD_TRY(k5_collapse_exact_nosync(ctx, d_keys, d_payload, d_keys_tmp, d_payload_tmp, n_reads, 0, params, d_fam, fcap,
                                          want_groups ? d_grp : nullptr, gcap, fam_start.as<uint32_t>(), grp_start.as<uint32_t>(),
                                          counts.as<uint32_t>(), nullptr, 0, 0));
        PMRD_TRY(read_count(ctx, counts.as<uint32_t>(), &nf));
        if (want_groups) PMRD_TRY(read_count(ctx, counts.as<uint32_t>() + 1, &ng));
    } else {
        // FIRSTFIT / GREEDY: reads get a new key group:32 | cluster:32, a stable sort by that key lists every
        // cluster's reads in input order, and the consensus kernel gathers the payloads through the index.
        DevBuf gk, idx, idx_tmp, reps, grp_start0, cnt0;
        PMRD_CUDA(ctx, gk.alloc(ctx->stream, (size_t)n_reads * 8));
        PMRD_CUDA(ctx, idx.alloc(ctx->stream, (size_t)n_reads * 4));
        PMRD_CUDA(ctx, idx_tmp.alloc(ctx->stream, (size_t)n_reads * 4));
        PMRD_CUDA(ctx, reps.alloc(ctx->stream, (size_t)n_reads * 4));
        PMRD_CUDA(ctx, grp_start0.alloc(ctx->stream, (size_t)(n_reads + 1) * 4));
        PMRD_CUDA(ctx, cnt0.alloc(ctx->stream, 4));
        const int g256 = pmrd_blocks(n_reads, 256);
        uint64_t varying = 0;
        const uint64_t* fk = nullptr;      // sorted (group | cluster) keys ...
        const uint32_t* fi = nullptr;      // ... and the original index of every read, cluster by cluster in input order
        if (P.cluster == PMRD_CLUSTER_FIRSTFIT) {
            // (1) stable sort read indices by group, (2) sequential first-fit per group, (3) stable sort by (group, cluster)
            PMRD_LAUNCH(ctx, iota_kernel, g256, 256, 0, idx.as<uint32_t>(), n_reads);
            PMRD_LAUNCH(ctx, groupkey_kernel, g256, 256, 0, (const uint64_t*)d_keys, gk.as<uint64_t>(), n_reads);
            PMRD_TRY(radix_varying_bits(ctx, gk.as<uint64_t>(), n_reads, &varying));
            int in_b = 0;
            PMRD_TRY(radix_sort_pairs(ctx, gk.as<uint64_t>(), idx.as<uint32_t>(), d_keys_tmp, idx_tmp.as<uint32_t>(), n_reads, varying, &in_b));
            uint64_t* sk = in_b ? d_keys_tmp : gk.as<uint64_t>();
            uint32_t* si = in_b ? idx_tmp.as<uint32_t>() : idx.as<uint32_t>();
            uint64_t* ok = in_b ? gk.as<uint64_t>() : d_keys_tmp;
            uint32_t* oi = in_b ? idx.as<uint32_t>() : idx_tmp.as<uint32_t>();
            int64_t ng0 = 0;
            PMRD_TRY(find_segments_dev(ctx, sk, n_reads, nullptr, 32, n_reads, grp_start0.as<uint32_t>(), cnt0.as<uint32_t>()));
            PMRD_TRY(read_count(ctx, cnt0.as<uint32_t>(), &ng0));
            // new keys (group | cluster) are written into `ok`, at the same positions as sk/si
            PMRD_LAUNCH(ctx, firstfit_kernel, pmrd_blocks(ng0 * 32, 128), 128, 0, (const uint64_t*)sk, (const uint32_t*)si,
                        (const uint64_t*)d_keys, (const uint32_t*)grp_start0.as<uint32_t>(), ng0, P.max_distance,
                        reps.as<uint32_t>(), ok);
            // sort (ok, si) by cluster bits; group bits are already in order and constant-digit passes are skipped
            PMRD_TRY(radix_varying_bits(ctx, ok, n_reads, &varying));
            int in_b2 = 0;
            PMRD_TRY(radix_sort_pairs(ctx, ok, si, sk, oi, n_reads, varying, &in_b2));
            fk = in_b2 ? sk : ok;
            fi = in_b2 ? oi : si;
        } else {
            // GREEDY (umi_clustering.py:59-101).  (1) stable sort read indices by the FULL key: exact-UMI runs;
            // (2) unique UMIs per group ordered by (-count, umi); (3) sequential seeding per group; (4) every read
            // gets group | cluster at its original position; (5) stable sort by that key.
            PMRD_LAUNCH(ctx, iota_kernel, g256, 256, 0, idx.as<uint32_t>(), n_reads);
            PMRD_CUDA(ctx, cudaMemcpyAsync(gk.p, d_keys, (size_t)n_reads * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            PMRD_TRY(radix_varying_bits(ctx, gk.as<uint64_t>(), n_reads, &varying));
            int in_b = 0;
            PMRD_TRY(radix_sort_pairs(ctx, gk.as<uint64_t>(), idx.as<uint32_t>(), d_keys_tmp, idx_tmp.as<uint32_t>(), n_reads, varying, &in_b));
            uint64_t* sk = in_b ? d_keys_tmp : gk.as<uint64_t>();
            uint32_t* si = in_b ? idx_tmp.as<uint32_t>() : idx.as<uint32_t>();
            uint64_t* ok = in_b ? gk.as<uint64_t>() : d_keys_tmp;
            uint32_t* oi = in_b ? idx.as<uint32_t>() : idx_tmp.as<uint32_t>();
            uint32_t* useg = grp_start0.as<uint32_t>();
            int64_t nu = 0;
            PMRD_TRY(find_segments_dev(ctx, sk, n_reads, nullptr, 0, n_reads, useg, cnt0.as<uint32_t>()));
            PMRD_TRY(read_count(ctx, cnt0.as<uint32_t>(), &nu));
            DevBuf uk, uk_tmp, ui, ui_tmp, gst, gcnt, sumi, scl;
            PMRD_CUDA(ctx, uk.alloc(ctx->stream, (size_t)nu * 8));
            PMRD_CUDA(ctx, uk_tmp.alloc(ctx->stream, (size_t)nu * 8));
            PMRD_CUDA(ctx, ui.alloc(ctx->stream, (size_t)nu * 4));
            PMRD_CUDA(ctx, ui_tmp.alloc(ctx->stream, (size_t)nu * 4));
            PMRD_CUDA(ctx, gst.alloc(ctx->stream, (size_t)(nu + 1) * 4));
            PMRD_CUDA(ctx, gcnt.alloc(ctx->stream, 4));
            PMRD_CUDA(ctx, sumi.alloc(ctx->stream, (size_t)nu * 4));
            PMRD_CUDA(ctx, scl.alloc(ctx->stream, (size_t)nu * 4));
            const int u256 = pmrd_blocks(nu, 256);
            PMRD_LAUNCH(ctx, greedy_uniq_kernel, u256, 256, 0, (const uint64_t*)sk, (const uint32_t*)useg, nu, uk.as<uint64_t>(), ui.as<uint32_t>());
            uint64_t uvar = 0;
            PMRD_TRY(radix_varying_bits(ctx, uk.as<uint64_t>(), nu, &uvar));
            int in_u = 0;
            PMRD_TRY(radix_sort_pairs(ctx, uk.as<uint64_t>(), ui.as<uint32_t>(), uk_tmp.as<uint64_t>(), ui_tmp.as<uint32_t>(), nu, uvar, &in_u));
            const uint64_t* suk = in_u ? uk_tmp.as<uint64_t>() : uk.as<uint64_t>();
            const uint32_t* ord = in_u ? ui_tmp.as<uint32_t>() : ui.as<uint32_t>();
            int64_t ng0 = 0;
            PMRD_TRY(find_segments_dev(ctx, suk, nu, nullptr, 32, nu, gst.as<uint32_t>(), gcnt.as<uint32_t>()));
            PMRD_TRY(read_count(ctx, gcnt.as<uint32_t>(), &ng0));
            PMRD_LAUNCH(ctx, greedy_gather_kernel, u256, 256, 0, (const uint64_t*)sk, (const uint32_t*)useg, ord, nu, sumi.as<uint32_t>(), scl.as<uint32_t>());
            PMRD_LAUNCH(ctx, greedy_cluster_kernel, (int)ng0, 256, 0, (const uint32_t*)sumi.as<uint32_t>(), scl.as<uint32_t>(),
                        (const uint32_t*)gst.as<uint32_t>(), ng0, P.max_distance);
            uint32_t* ucluster = reps.as<uint32_t>();          // [nu] <= [n_reads]
            PMRD_LAUNCH(ctx, greedy_scatter_kernel, u256, 256, 0, ord, (const uint32_t*)scl.as<uint32_t>(), nu, ucluster);
            PMRD_LAUNCH(ctx, greedy_assign_kernel, g256, 256, 0, (const uint64_t*)sk, (const uint32_t*)si, n_reads, (const uint32_t*)useg, nu,
                        (const uint32_t*)ucluster, ok);
            PMRD_LAUNCH(ctx, iota_kernel, g256, 256, 0, oi, n_reads);
            PMRD_TRY(radix_varying_bits(ctx, ok, n_reads, &varying));
            int in_b2 = 0;
            PMRD_TRY(radix_sort_pairs(ctx, ok, oi, sk, si, n_reads, varying, &in_b2));
            fk = in_b2 ? sk : ok;
            fi = in_b2 ? si : oi;
            PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // the unique-UMI scratch is released at scope exit
        }
        PMRD_TRY(find_segments_dev(ctx, fk, n_reads, nullptr, 0, fcap, fam_start.as<uint32_t>(), counts.as<uint32_t>()));
        PMRD_TRY(read_count(ctx, counts.as<uint32_t>(), &nf));
        FamOut out{d_fam->key, d_fam->size, d_fam->n_alt, d_fam->consensus, d_fam->qsum, d_fam->n_best, d_fam->flags};
        if (nf <= fcap) {
            PMRD_TRY(launch_family<true>(ctx, P.mode, fk, fi, d_keys, d_payload, fam_start.as<uint32_t>(), counts.as<uint32_t>(), fcap, P, out));
            if (want_groups) {
                PMRD_TRY(find_segments_dev(ctx, d_fam->key, nf, nullptr, 32, gcap, grp_start.as<uint32_t>(), counts.as<uint32_t>() + 1));
                PMRD_TRY(read_count(ctx, counts.as<uint32_t>() + 1, &ng));
                int grid = (int)(gcap < (int64_t)ctx->sm_count * 32 ? gcap : (int64_t)ctx->sm_count * 32);
                if (grid < 1) grid = 1;
                PMRD_LAUNCH(ctx, group_kernel, grid, 128, 0, (const uint64_t*)d_fam->key, (const uint32_t*)d_fam->size,
                            (const uint32_t*)d_fam->n_alt, (const uint32_t*)d_fam->consensus, (const uint32_t*)d_fam->flags,
                            (const uint32_t*)grp_start.as<uint32_t>(), (const uint32_t*)(counts.as<uint32_t>() + 1), gcap,
                            P.mode == PMRD_MODE_WEIGHTED ? 1 : 0, *d_grp);
            }
        }
        PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *n_families = nf;
    if (n_groups) *n_groups = ng;
    if (nf > capacity_f || (want_groups && ng > capacity_g)) {
        snprintf(ctx->err, sizeof(ctx->err), "capacity too small: %lld families (cap %lld), %lld groups (cap %lld)",
                 (long long)nf, (long long)capacity_f, (long long)ng, (long long)capacity_g);
        return PMRD_ERR_CAPACITY;
    }
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PMRD_OK;
}
