// radix_sort.cuh -- stable LSD radix sort of (u64 key, u32 value) pairs (kernel K4).
//
// Per digit pass (<= 8 bits):   upsweep   : per-tile digit histogram (128-bit key loads)
//                               scan      : exclusive scan of hist[digit][tile]
//                               downsweep : warp match-any ranking -> shared-memory
//                                           reorder -> coalesced per-digit run stores
// Digit windows are planned over the bits that actually vary in the input (one OR/AND
// reduction up front), so a 12-mer UMI (24 bits) + 2 000 cells (11 bits) costs 5 passes,
// not 8.  No inter-block waiting anywhere (safe under MPS / time-slicing).
#pragma once
#include "common.cuh"
#include "scan.cuh"

#define RS_THREADS 512
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS) /* 4096 pairs per tile */
#define RS_MAX_BINS 256

#ifdef __CUDACC__
// Peer mask of a warp-wide digit: the lanes holding the same 8-bit digit as the caller.
// Eight ballots, each written so that ONE predicate feeds both the vote and the select -- ptxas then
// moves all eight digit bits into predicates with a single R2P and spends three instructions per bit
// (VOTE, SEL, LOP3).  The C++ form `m &= bit ? bal : ~bal` costs six per bit, and MATCH.ANY on the
// whole digit saturates the ADU pipe (ncu: 88 %).  Measured on B200, direct downsweep, 2.13e9 records x 3
// passes: C++ ballots + MATCH on the top 3 bits 34.5 ms; this form 28.3 ms (MATCH on the top 1-2 bits: same;
// 3 bits 29.1; 4 bits 31.1).
template <int BITS = 8>
__device__ __forceinline__ uint32_t rs_peer_mask(uint32_t d, uint32_t m) {
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t, bal, x;\n\t"
                     "and.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
                     "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
                     "selp.b32 x, 0, 0xffffffff, p;\n\t"
                     "lop3.b32 %0, %0, bal, x, 0x60;\n\t}"      // m & (bal ^ x)
                     : "+r"(m) : "r"(d), "r"(1u << b));
    }
    return m;
}

#endif

struct RadixPlan {
    int n_passes;
    int shift[8];
    int bits[8];
};

// h_or (optional): the OR of all keys (an upper bound of every key field)
int32_t radix_varying_bits(pmrd_ctx* ctx, const uint64_t* d_keys, int64_t n, uint64_t* h_varying, uint64_t* h_or = nullptr);
RadixPlan radix_make_plan(uint64_t varying, int max_digit_bits = 8);
// Sorts; on return *in_b tells whether the result lives in (keys_b, vals_b).  vals may be
// NULL (keys only).  `varying` = mask of key bits that differ across the input (pass ~0ull
// to sort on every bit).
// d_scratch (optional): radix_scratch_bytes(n) bytes owned by the caller -- the sort then allocates nothing.
// max_digit_bits = 9: digit windows of up to 9 bits (512 bins; scratch is sized for them either way)
int32_t radix_sort_pairs(pmrd_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                         int64_t n, uint64_t varying, int* in_b, uint32_t* d_scratch = nullptr, int max_digit_bits = 8);
size_t radix_scratch_bytes(int64_t n);
int32_t radix_sort_pairs_segmented(pmrd_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                                   int64_t n, const int64_t* d_cell_tile_start, int64_t n_cells, int key_bits, int* in_b,
                                   uint32_t* d_hist0 = nullptr);
