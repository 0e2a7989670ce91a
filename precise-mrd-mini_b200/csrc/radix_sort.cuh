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

struct RadixPlan {
    int n_passes;
    int shift[8];
    int bits[8];
};

int32_t radix_varying_bits(pmrd_ctx* ctx, const uint64_t* d_keys, int64_t n, uint64_t* h_varying);
RadixPlan radix_make_plan(uint64_t varying);
// Sorts; on return *in_b tells whether the result lives in (keys_b, vals_b).  vals may be
// NULL (keys only).  `varying` = mask of key bits that differ across the input (pass ~0ull
// to sort on every bit).
int32_t radix_sort_pairs(pmrd_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                         int64_t n, uint64_t varying, int* in_b);
int32_t radix_sort_pairs_segmented(pmrd_ctx* ctx, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                                   int64_t n, const int64_t* d_cell_tile_start, int64_t n_cells, int key_bits, int* in_b,
                                   uint32_t* d_hist0 = nullptr);
