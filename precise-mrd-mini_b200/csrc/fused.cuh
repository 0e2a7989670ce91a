// fused.cuh -- host-visible pieces of the no-materialise grid path (fused.cu).
#pragma once
#include "common.cuh"
#include "reads.cuh"

struct FusedChunk {
    DevBuf knobs;                       // CellKnobDev[n_cells] or empty
    DevBuf cells, size_cdf, size_lut;   // ReadCell[n_cells + 1] (the closing entry carries the family total), size tables
    pmrd_read_sim_params sp;
    bool has_size_cdf = false;
    int64_t n_cells = 0, cell0 = 0, n_fam = 0;
    int64_t n_reads = 0;                // reads drawn per step (a function of the seed: counted at build)
    int64_t n_cand = 0;                 // candidate duplicates per step (exact: sizes the list and its sort)
    int64_t n_cells_with_reads = 0;
    int threads = 1024;                 // block size: 1024 (cells > 4096 families), 256, 128
    int filt_log2 = 18;                 // slots per hash bitmap
};

int32_t fused_upload_tables(pmrd_ctx* ctx, const double* h_w64);
int32_t fused_chunk_build(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af, const int64_t* h_n_families, int64_t n_cells,
                          int64_t cell0, const pmrd_read_sim_params* sp, const pmrd_umi_params* up, FusedChunk* ch,
                          const CellExtra* h_extra = nullptr /* [n_cells], already sliced to the chunk */);
size_t fused_scratch_bytes(int64_t n_cand);
int32_t fused_chunk_run(pmrd_ctx* ctx, FusedChunk* ch, const pmrd_umi_params* up, void* d_scratch, const uint8_t* d_alt,
                        int64_t* o_reads, int64_t* o_fam, int64_t* o_tot, int64_t* o_var);
