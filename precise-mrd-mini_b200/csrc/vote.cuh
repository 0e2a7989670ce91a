// vote.cuh -- the quality-weighted consensus vote (umi.py:146-174) for reads that can only carry TWO alleles.
#pragma once
#include "common.cuh"

// The generator only ever emits a cell's ref or alt allele (io.py:105-108), so a vote over GENERATED reads (the plan's
// cell-major kernels, the fused engine) needs two sums -- the tested allele and the other one -- both in registers: no
// shared-memory accumulator columns, no address arithmetic.  Same IEEE additions in the same order as the four-sum
// walk (an allele that never occurs contributes no addition there either), hence the same decisions.  `x` is a
// payload word or its low byte: allele[1:0] | quality << 2.
template <bool MAJ>
struct FoldVote2  {
    double total = 0.0, wref = 0.0, walt = 0.0;
    uint32_t first = 0;            // !MAJ: 1 = ref voted first, 2 = alt voted first (the dict-order tie-break of umi.py:166)
    __device__ __forceinline__ void add(uint32_t byte, uint32_t alt, const double* wt) {
        add_w(byte, alt, wt[(byte >> 2) & 63u]);             // 0.0 below the quality threshold: adds nothing
    }
    // the same with the weight handed in: +0.0 is the identity on these non-negative sums and sets no state, so a caller
    // can pass 0.0 for a read that belongs to another family instead of branching around the call
    __device__ __forceinline__ void add_w(uint32_t byte, uint32_t alt, double w) {
        const bool is_alt = (byte & 3u) == alt;
        total = __dadd_rn(total, w);
#ifdef PMRD_FOLD2_SELECT
        // one DADD on the selected sum, written back by selects (the FP64 pipe issues a warp DADD over two cycles)
        const double acc = __dadd_rn(is_alt ? walt : wref, w);
        walt = is_alt ? acc : walt;
        wref = is_alt ? wref : acc;
#else
        if (is_alt) walt = __dadd_rn(walt, w);
        else wref = __dadd_rn(wref, w);
#endif
        if (!MAJ) first = (first == 0u && __double2hiint(w) != 0) ? (is_alt ? 2u : 1u) : first;
    }
    // 1 = consensus is the alt allele, 0 = the ref allele, -1 = none (umi.py:166-174)
    __device__ __forceinline__ int result(double cthr) const {
        if (total == 0.0) return -1;                         // no read passed the quality threshold
        bool alt_wins;
        if (MAJ) alt_wins = walt > wref;                     // (a tie cannot pass a threshold above one half)
        else alt_wins = walt > wref || (walt == wref && first == 2u);
        const double bw = alt_wins ? walt : wref;
        if (bw != total && __ddiv_rn(bw, total) < cthr) return -1;      // bw == total: the fraction is exactly 1 >= cthr
        return alt_wins ? 1 : 0;
    }
};

