// kernels.cuh -- internal (device-pointer) entry points shared between translation units.
#pragma once
#include "common.cuh"

// pvalues.cu
int32_t k8_pvalues(pmrd_ctx* ctx, const int64_t* d_k, const int64_t* d_n, const double* d_rate, int64_t rate_stride,
                   int64_t m, int32_t test_type, int32_t alternative, double* d_p, const int32_t* d_rate_index = nullptr);
int32_t k9_bh_groups(pmrd_ctx* ctx, const double* d_p, const int64_t* d_group_start, int64_t n_groups, double alpha, uint8_t* d_rejected,
                     double* d_q);
int32_t k9_bh(pmrd_ctx* ctx, const double* d_p, int64_t m, double alpha, uint8_t* d_rejected, double* d_q, void* d_workspace = nullptr);
size_t bh_workspace_bytes(int64_t m);

// call.cu
int32_t k7_call(pmrd_ctx* ctx, const int64_t* d_seg_offsets, int64_t n_segments, const uint8_t* d_is_variant,
                const double* d_quality, const double* d_agreement, double mean_error_rate, int32_t test_type,
                double alpha, const pmrd_call_table* d_out);

// collapse.cu
int32_t k5_collapse(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp, uint32_t* d_payload_tmp,
                    int64_t n_reads, const pmrd_umi_params* params, const pmrd_family_table* d_fam, int64_t capacity_f,
                    int64_t* n_families, const pmrd_group_table* d_grp, int64_t capacity_g, int64_t* n_groups);

int32_t k14_fisher(pmrd_ctx* ctx, const int64_t* d_table, int64_t m, double* d_odds, double* d_p);

// fastq.cu
int32_t k13_fastq_scan(pmrd_ctx* ctx, const uint8_t* d_buf, int64_t n_bytes, int64_t capacity, const pmrd_fastq_records* d_out,
                       int64_t* h_n_records);
int32_t k13_fastq_families(pmrd_ctx* ctx, uint64_t* d_keys, int64_t n, const int64_t* d_qual_sum, const int32_t* d_qual_len,
                           const pmrd_fastq_families* d_out, int64_t capacity, int64_t* h_n_fam);

// metrics.cu
int32_t k12_bootstrap_metrics(pmrd_ctx* ctx, const uint8_t* d_y, const double* d_score, int64_t n, const int64_t* d_indices,
                              int64_t n_boot, int want_point, uint64_t stream_id, double* d_auc, double* d_ap);

// simulate.cu
int32_t k1_simulate_cells(pmrd_ctx* ctx, const int64_t* d_cell_id, const double* d_af, const int64_t* d_depth,
                          const double* d_bg_mult, int64_t n_cells, int32_t max_family_size, const pmrd_cell_table* d_out);

// exchange.cu (K18): external reads held by several GPUs -- stable cut into one run per owner rank, dense ids after the exchange
int32_t k18_owner_partition(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp, uint32_t* d_payload_tmp,
                            int64_t n, int32_t log2_world, int64_t* h_counts, int32_t* in_tmp);
int32_t k18_owner_rekey(pmrd_ctx* ctx, uint64_t* d_keys, int64_t n, int32_t log2_world);
int32_t k18_owner_tiles(pmrd_ctx* ctx, const uint64_t* d_keys, int64_t n, int32_t log2_world, uint32_t* d_tile_off, int64_t* h_counts);
int32_t k18_owner_scatter_peer(pmrd_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_payload, int64_t n, int32_t log2_world,
                               const uint32_t* d_tile_off, uint64_t* const* peer_keys, uint32_t* const* peer_payload, const int64_t* h_base);
int32_t k18_peer_alloc(pmrd_ctx* ctx, int64_t bytes, void** d_ptr, uint8_t* handle64);
int32_t k18_peer_open(pmrd_ctx* ctx, const uint8_t* handle64, void** d_ptr);
int32_t k18_peer_close(pmrd_ctx* ctx, void* d_ptr);
int32_t k18_peer_free(pmrd_ctx* ctx, void* d_ptr);
