/*
 * pmrd.h -- C ABI of libpmrd_b200.so: the B200-native simulate -> UMI-collapse -> call
 * hot path of precise-mrd (reference: altalanta/precise-mrd-mini).
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns an int32
 * status (0 = ok, negative = error; text via pmrd_last_error) and never throws.  One
 * context = one CUDA device + one CUDA stream; contexts are thread-compatible, not
 * thread-safe.  Pointers named h_* are HOST buffers (the call copies host<->device on
 * the context's stream and synchronises before returning); pointers named d_* are DEVICE
 * buffers on the context's device (the call only enqueues work on the context's stream;
 * the caller synchronises).  All output buffers are caller-allocated.
 *
 * The reference has no FFI for this path (its one native component, the pyo3/rayon crate
 * `precise_mrd_ext`, is never imported: SURVEY.md §2.2).  Each function below cites the
 * reference interface (file:line under /root/reference) whose arithmetic it replaces;
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 */
#ifndef PMRD_H
#define PMRD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMRD_ABI_VERSION 1

/* status codes */
#define PMRD_OK 0
#define PMRD_ERR_CUDA (-1)      /* a CUDA runtime call failed (text in pmrd_last_error) */
#define PMRD_ERR_ARG (-2)       /* invalid argument                                     */
#define PMRD_ERR_CAPACITY (-3)  /* caller-provided output capacity too small            */
#define PMRD_ERR_NO_DEVICE (-4) /* no usable sm_100 device                              */

typedef struct pmrd_ctx pmrd_ctx;

/* ---------------------------------------------------------------- context ---------- */
int32_t pmrd_abi_version(void);
/* device: CUDA ordinal.  seed: Philox key (config.seed; cli.py:63 makes the CLI seed
 * authoritative).  Fails with PMRD_ERR_NO_DEVICE when no GPU is visible -- there is no
 * CPU fallback anywhere in this library. */
int32_t pmrd_create(int32_t device, uint64_t seed, pmrd_ctx** out);
void pmrd_destroy(pmrd_ctx* ctx);
const char* pmrd_last_error(const pmrd_ctx* ctx); /* ctx may be NULL: last create error */
int32_t pmrd_set_seed(pmrd_ctx* ctx, uint64_t seed);
/* Run on an externally owned stream (e.g. torch.cuda.current_stream().cuda_stream);
 * NULL restores the context's own stream. */
int32_t pmrd_set_stream(pmrd_ctx* ctx, void* cuda_stream);
int32_t pmrd_sync(pmrd_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py: gpu_launches) */
int64_t pmrd_launch_count(const pmrd_ctx* ctx);
/* high-water mark (bytes) of the device's stream-ordered allocator pool: `precise-mrd performance` (cli.py:408-438, peak_memory_mb) */
int64_t pmrd_peak_memory_bytes(pmrd_ctx* ctx);
/* Per-kernel device timing: while enabled, every launch is bracketed by CUDA events on the
 * launching stream.  pmrd_profile_report synchronises and writes "<kernel> <launches>
 * <total_ms>\n" lines into buf (returns bytes written) and clears the records. */
int32_t pmrd_profile_enable(pmrd_ctx* ctx, int32_t on);
int64_t pmrd_profile_report(pmrd_ctx* ctx, char* buf, int64_t cap);

/* ---------------------------------------------------------------- K8: tail p-values - */
/* test_type */
#define PMRD_TEST_POISSON 0  /* call.py:14-25  poisson_test(observed, expected)          */
#define PMRD_TEST_BINOMIAL 1 /* call.py:28-38  binomial_test(successes, trials, p)       */
/* alternative */
#define PMRD_ALT_TWO_SIDED 0 /* G-D: 2*min(cdf(k), sf(k-1)) / scipy binomtest "minlike"  */
#define PMRD_ALT_GREATER 1   /* G-B: stats.py:47-133  1 - cdf(k-1) / binomtest 'greater' */
#define PMRD_ALT_LESS 2      /* G-B: stats.py:61-63   cdf(k)       / binomtest 'less'    */

/* p[i] = test(k[i], n[i], rate[i]).  Poisson: expected = n[i]*rate[i] is formed on the
 * device exactly as call.py:179 does (one IEEE multiply); pass n=NULL to give `rate`
 * directly as the expectation.  rate_stride 0 broadcasts rate[0]. */
int32_t pmrd_pvalues(pmrd_ctx* ctx, const int64_t* h_k, const int64_t* h_n, const double* h_rate,
                     int64_t rate_stride, int64_t m, int32_t test_type, int32_t alternative,
                     double* h_p);
int32_t pmrd_pvalues_dev(pmrd_ctx* ctx, const int64_t* d_k, const int64_t* d_n, const double* d_rate,
                         int64_t rate_stride, int64_t m, int32_t test_type, int32_t alternative,
                         double* d_p);

/* ---------------------------------------------------------------- K9: BH FDR -------- */
/* call.py:41-79 benjamini_hochberg_correction(p_values, alpha) -> (rejected, adjusted).
 * Device radix sort of the fp64 bit patterns + reverse running-min scan.  m == 0 is ok. */
int32_t pmrd_bh(pmrd_ctx* ctx, const double* h_p, int64_t m, double alpha, uint8_t* h_rejected,
                double* h_q);
int32_t pmrd_bh_dev(pmrd_ctx* ctx, const double* d_p, int64_t m, double alpha, uint8_t* d_rejected,
                    double* d_q);

/* ---------------------------------------------------------------- K7+K8+K9: call ---- */
/* Per-sample result of call_mrd (call.py:161-229), one row per segment. */
typedef struct pmrd_call_table {
    int64_t* n_variants;       /* [S] sum(is_variant)                      call.py:171 */
    int64_t* n_total;          /* [S] rows in segment                      call.py:172 */
    double* variant_fraction;  /* [S] n_variants / n_total                              */
    double* expected_variants; /* [S] n_total * mean_error_rate            call.py:179 */
    double* fold_change;       /* [S] k/expected | inf | 1.0               call.py:190 */
    double* p_value;           /* [S]                                      call.py:182 */
    double* mean_quality;      /* [S] mean(quality_score)                  call.py:196 */
    double* mean_consensus;    /* [S] mean(consensus_agreement)            call.py:197 */
    double* p_adjusted;        /* [S] BH                                   call.py:220 */
    uint8_t* significant;      /* [S] BH rejected (== variant_call)  */
} pmrd_call_table;

/* seg_offsets[S+1]: rows of segment s are [seg_offsets[s], seg_offsets[s+1]) -- the
 * caller groups rows by sample in first-appearance order (call.py:161).  is_variant /
 * quality / agreement have n_rows = seg_offsets[S] entries.  quality/agreement may be
 * NULL (means are then NaN).  All table pointers are host (pmrd_call) or device
 * (pmrd_call_dev) arrays of S entries. */
int32_t pmrd_call(pmrd_ctx* ctx, const int64_t* h_seg_offsets, int64_t n_segments,
                  const uint8_t* h_is_variant, const double* h_quality, const double* h_agreement,
                  double mean_error_rate, int32_t test_type, double alpha, pmrd_call_table* h_out);
int32_t pmrd_call_dev(pmrd_ctx* ctx, const int64_t* d_seg_offsets, int64_t n_segments,
                      const uint8_t* d_is_variant, const double* d_quality, const double* d_agreement,
                      double mean_error_rate, int32_t test_type, double alpha, pmrd_call_table* d_out);

/* ---------------------------------------------------------------- K4: radix sort ---- */
/* Stable LSD radix sort of (u64 key, u32 value) pairs; digit passes whose digit is
 * constant over the whole input are skipped.  Exposed because BH, collapse and the
 * read generator all use it and the tests pin it separately. */
int32_t pmrd_sort_pairs(pmrd_ctx* ctx, const uint64_t* h_keys, const uint32_t* h_vals, int64_t n,
                        uint64_t* h_keys_out, uint32_t* h_vals_out);
int32_t pmrd_sort_pairs_dev(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, int64_t n,
                            uint64_t* d_keys_tmp, uint32_t* d_vals_tmp, int32_t* out_in_tmp);

/* ---------------------------------------------------------------- K4-K7: collapse --- */
/* Read record (SoA, 12 B):
 *   key     u64 = group:32 | umi:32      group = dense (sample, locus) code, by convention
 *                                        sample:12 | locus:20; umi = 2 bit/base, <=16-mer
 *   payload u32, by mode:
 *     PMRD_MODE_WEIGHTED  allele[1:0] qual[7:2] strand[8] read_pos[16:9] is_alt[31]
 *     PMRD_MODE_MAJORITY  sequence[19:0] (10-mer, 2 bit/base, base 0 in the top bits) is_alt[31]
 *     PMRD_MODE_COUNT     is_alt[31] only
 */
#define PMRD_MODE_WEIGHTED 0 /* G-B UMIProcessor.call_consensus       src/precise_mrd/umi.py:146-179, rust_ext/src/lib.rs:25-74 */
#define PMRD_MODE_MAJORITY 1 /* G-A _consensus_base                   precise-mrd-mini/src/mrd/umi.py:43-54 */
#define PMRD_MODE_COUNT 2    /* G-C REF/ALT read counts per UMI       test_results/run/collapsed_umis.parquet */

#define PMRD_CLUSTER_EXACT 0    /* exact UMI match   umi.py:111-115, lib.rs:7-22                  */
#define PMRD_CLUSTER_FIRSTFIT 1 /* first-fit Hamming <= d to the representative  mrd/umi.py:29-40 */
#define PMRD_CLUSTER_GREEDY 2   /* greedy Hamming <= d around seeds taken in (-count, umi) order:
                                   umi_clustering.py:59-101, the defined-order form of umi.py:117-144 */

typedef struct pmrd_umi_params {
    int32_t min_family_size;     /* families below are dropped (after consensus, on raw size: umi.py:181-203) */
    int32_t max_family_size;     /* > this: reported via the oversize flag (reference down-samples with the global RNG) */
    int32_t quality_threshold;   /* reads with qual < this do not vote            umi.py:152 */
    double consensus_threshold;  /* w_best / w_total must be >= this              umi.py:172 */
    int32_t mode;                /* PMRD_MODE_*                                               */
    int32_t cluster;             /* PMRD_CLUSTER_*                                            */
    int32_t max_distance;        /* Hamming radius for FIRSTFIT / GREEDY                      */
    int32_t umi_len;             /* bases in the UMI (<=16)                                   */
} pmrd_umi_params;

/* family flags */
#define PMRD_FAM_HAS_CONSENSUS 1u /* call_consensus returned an allele                  */
#define PMRD_FAM_KEPT 2u          /* size >= min_family_size                            */
#define PMRD_FAM_OVERSIZE 4u      /* size > max_family_size                             */

typedef struct pmrd_family_table {
    uint64_t* key;       /* [F] group:32 | representative umi:32; sorted ascending (EXACT) or
                            group-sorted then cluster-creation order (FIRSTFIT: the UMI that opened
                            the cluster; GREEDY: seed order, the UMI of the cluster's first read) */
    uint32_t* size;      /* [F] reads in the family                                       */
    uint32_t* n_alt;     /* [F] reads with is_alt set                                     */
    uint32_t* consensus; /* [F] WEIGHTED: allele code; MAJORITY: 20-bit consensus sequence */
    uint32_t* qsum;      /* [F] WEIGHTED: sum of qualities of voting reads that carry the
                                consensus allele (consensus_quality = qsum / n_best)      */
    uint32_t* n_best;    /* [F] WEIGHTED: number of those reads                           */
    uint32_t* flags;     /* [F] PMRD_FAM_*                                                */
} pmrd_family_table;

typedef struct pmrd_group_table {
    uint32_t* group;          /* [G] group id (ascending)                                        */
    uint32_t* family_count;   /* [G] kept families                                               */
    uint32_t* consensus_count;/* [G*4] WEIGHTED: kept families with consensus allele a (umi.py:241-266) */
    uint64_t* alt_count;      /* [G] sum n_alt over kept families   (mrd/umi.py:97)              */
    uint64_t* total_count;    /* [G] sum size over kept families    (mrd/umi.py:98)              */
} pmrd_group_table;

/* Collapse reads into UMI families and per-group (locus) summaries.
 *   EXACT:    any input order; reads of a family vote in input order (stable sort).
 *   FIRSTFIT: reads must arrive with each group's reads in the order the reference's
 *             groupby would iterate them (input order within group is preserved).
 * capacity_f / capacity_g: entries available in the output tables; *n_families /
 * *n_groups receive the counts (PMRD_ERR_CAPACITY if too small; n <= n_reads always
 * suffices).  The family table lists ALL families (flags tell kept/consensus), so that
 * the host can apply either generation's filtering order.  h_fam->key == NULL: the family rows are
 * not downloaded (only *n_families and the group table leave the device).  Large pageable host buffers
 * move through a ring of pinned staging buffers filled by a few host threads (csrc/staging.cu), so the
 * call runs at the PCIe / host-memcpy rate instead of the driver's single-threaded pageable path. */
int32_t pmrd_collapse_reads(pmrd_ctx* ctx, const uint64_t* h_keys, const uint32_t* h_payload,
                            int64_t n_reads, const pmrd_umi_params* params,
                            pmrd_family_table* h_fam, int64_t capacity_f, int64_t* n_families,
                            pmrd_group_table* h_grp, int64_t capacity_g, int64_t* n_groups);
/* Device-resident variant: d_keys/d_payload are clobbered (used as sort ping-pong with
 * the *_tmp buffers).  n_families/n_groups are host pointers written after a sync.
 * d_keys and d_keys_tmp must be 16-byte aligned (any cudaMalloc'd array is; a view that starts
 * at an odd element is refused with PMRD_ERR_ARG): the kernels read keys two at a time. */
int32_t pmrd_collapse_reads_dev(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload,
                                uint64_t* d_keys_tmp, uint32_t* d_payload_tmp, int64_t n_reads,
                                const pmrd_umi_params* params, pmrd_family_table* d_fam,
                                int64_t capacity_f, int64_t* n_families, pmrd_group_table* d_grp,
                                int64_t capacity_g, int64_t* n_groups);

/* External reads held by SEVERAL GPUs (one process per GPU): the exchange step of SURVEY.md section 8(e).
 * The reference collapses one process's reads (umi.py:205-285, lib.rs:7-109); with the reads of a sample striped
 * over 2^k ranks every group (locus) gets the owner rank  group mod 2^k.
 *   pmrd_owner_partition_dev: STABLE partition of this rank's reads into one contiguous run per owner (owner 0 first);
 *     h_counts[r] = reads bound for rank r; *in_tmp = 1 when the result sits in the *_tmp arrays.  The host then moves the
 *     runs (torch.distributed all_to_all_single over NCCL: INTEGRATION.md section 5); the runs arrive in source-rank order,
 *     so an owner sees the reads of a family in the order of the concatenated input and its table is the single-GPU one.
 *   pmrd_owner_rekey_dev: after the exchange a rank holds only groups = rank (mod 2^k); group >>= k makes the ids dense
 *     again for pmrd_collapse_reads_dev (global id = local << k | rank).
 * 16-byte aligned key arrays, as for pmrd_collapse_reads_dev. */
int32_t pmrd_owner_partition_dev(pmrd_ctx* ctx, uint64_t* d_keys, uint32_t* d_payload, uint64_t* d_keys_tmp,
                                 uint32_t* d_payload_tmp, int64_t n_reads, int32_t log2_world, int64_t* h_counts,
                                 int32_t* in_tmp);
int32_t pmrd_owner_rekey_dev(pmrd_ctx* ctx, uint64_t* d_keys, int64_t n_reads, int32_t log2_world);

/* The same exchange as ONE kernel over peer memory (up to 8 ranks of one NVSwitch node): instead of a partitioned copy that
 * NCCL then moves, the cut stores every read straight into its owner's receive arrays over NVLink.
 *   pmrd_peer_alloc / pmrd_peer_free: a receive array other processes can map (cudaMalloc + cudaIpcGetMemHandle); the
 *     64-byte handle is what the ranks exchange.  pmrd_peer_open / pmrd_peer_close: map / unmap a peer's array.
 *   pmrd_owner_tiles_dev: per 4096-read tile, where the tile's reads of each owner start inside this rank's run for that
 *     owner: d_tile_off[owner * n_tiles + tile], n_tiles = ceil(n_reads / 4096); h_counts[r] = reads bound for rank r.
 *   pmrd_owner_scatter_peer_dev: the cut itself; read i goes, with its group already divided by 2^k (no pmrd_owner_rekey_dev), to
 *       peer_keys[owner][ h_base[owner] + d_tile_off[owner][tile] + stable rank of i among the tile's reads of that owner ],
 *     h_base[owner] = reads that LOWER ranks send to that owner (the host all-gathers the count matrix), so an owner holds
 *     the runs in source-rank order, as all_to_all would deliver them.  peer_keys / peer_payload: HOST arrays of the 2^k
 *     device pointers (this rank's own arrays included).  The caller orders the ranks around the call (every owner's
 *     previous contents consumed before; every writer finished after). */
int32_t pmrd_peer_alloc(pmrd_ctx* ctx, int64_t bytes, void** d_ptr, uint8_t* handle64);
int32_t pmrd_peer_free(pmrd_ctx* ctx, void* d_ptr);
int32_t pmrd_peer_open(pmrd_ctx* ctx, const uint8_t* handle64, void** d_ptr);
int32_t pmrd_peer_close(pmrd_ctx* ctx, void* d_ptr);
int32_t pmrd_owner_tiles_dev(pmrd_ctx* ctx, const uint64_t* d_keys, int64_t n_reads, int32_t log2_world, uint32_t* d_tile_off,
                             int64_t* h_counts);
int32_t pmrd_owner_scatter_peer_dev(pmrd_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_payload, int64_t n_reads,
                                    int32_t log2_world, const uint32_t* d_tile_off, uint64_t* const* peer_keys,
                                    uint32_t* const* peer_payload, const int64_t* h_base);

/* K4 on the plan's cell-major layout: cell c = whole 4096-record tiles [tile_start[c], tile_start[c+1]) of packed 8-byte records;
 * every cell is sorted on its own, stable LSD over the low key_bits (<= 32) bits, the bits above are carried along -- the
 * staged / direct scatter kernels the fused plan sorts its generated reads with (umi.py:97-115 grouping), exposed so that
 * their ranking can be checked on arbitrary, adversarial digit distributions. */
int32_t pmrd_sort_packed_cells(pmrd_ctx* ctx, const uint64_t* h_rec, int64_t n, const int64_t* h_cell_tile_start,
                               int64_t n_cells, int32_t key_bits, uint64_t* h_out);

/* umi_edit_distance (lib.rs:77-90, umi.py:90-95): Hamming on 2-bit packed UMIs of equal
 * length.  d[i] = hamming(a[i], b[i]). */
int32_t pmrd_umi_hamming(pmrd_ctx* ctx, const uint32_t* h_a, const uint32_t* h_b, int64_t n,
                         int32_t umi_len, int32_t* h_d);

/* ---------------------------------------------------------------- K1: simulate cells */
/* simulate_reads (src/precise_mrd/simulate.py:133-180): one row per (AF, depth, rep)
 * cell.  Philox4x32-10 keyed by ctx seed, counter (stream=1, cell, 0, draw).
 * cell_id[i] is the global cell index (so results are identical under any sharding). */
typedef struct pmrd_cell_table {
    double* background_rate;     /* [C] U(1e-5, 1e-3)                     simulate.py:144 */
    int64_t* mean_family_size;   /* [C] clip(Poisson(5), 1, max_fs)       simulate.py:147-148 */
    int64_t* n_true_variants;    /* [C] Binom(depth, af)                  simulate.py:152 */
    int64_t* n_false_positives;  /* [C] Binom(depth - n_true, bg)         simulate.py:155-158 */
    double* mean_quality;        /* [C] clip(N(25,5), 10, 40)             simulate.py:161-162 */
} pmrd_cell_table;
int32_t pmrd_simulate_cells(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af,
                            const int64_t* h_depth, int64_t n_cells, int32_t max_family_size,
                            pmrd_cell_table* h_out);
/* StratifiedAnalyzer._simulate_with_context (src/precise_mrd/eval/stratified.py:180-203): the same
 * draws with a per-cell trinucleotide-context multiplier m (CpG 1.5, CHG 1.2, CHH 1.0, NpN 0.8):
 * background_rate *= m, and n_false_positives ~ Binom(depth - n_true, background_rate * m) -- the
 * reference applies m twice (:193 and :199).  h_bg_multiplier == NULL: m = 1 everywhere. */
int32_t pmrd_simulate_cells_stratified(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af,
                                       const int64_t* h_depth, const double* h_bg_multiplier,
                                       int64_t n_cells, int32_t max_family_size, pmrd_cell_table* h_out);

/* ---------------------------------------------------------------- K14: strand bias ---- */
/* QualityFilter.test_strand_bias (src/precise_mrd/filters.py:50-100): scipy.stats.fisher_exact(
 * [[c00, c01], [c10, c11]], 'two-sided') on m tables, h_table = [m][4] row-major.  Sample odds
 * ratio c00 c11 / (c10 c01) (inf when c10 c01 == 0; NaN and p = 1 when a row or column is empty);
 * p = sum of the hypergeometric point masses no larger than the observed one (relative slack 1e-14). */
int32_t pmrd_fisher_exact(pmrd_ctx* ctx, const int64_t* h_table, int64_t m, double* h_odds_ratio, double* h_p);

/* ---------------------------------------------------------------- K13: FASTQ ingest -- */
/* FASTQReader.read_fastq + the per-UMI grouping of process_fastq_to_dataframe
 * (src/precise_mrd/fastq.py:66-110, :176-220).  pmrd_fastq_scan takes the file's bytes (already
 * decompressed) and returns one row per 4-line record, in file order; offsets index the caller's
 * buffer, so strings are sliced on the host only for the rows that are kept. */
#define PMRD_FQ_END 0        /* stripped header empty: read_fastq stops here (fastq.py:85-86)       */
#define PMRD_FQ_MALFORMED 1  /* another line empty after strip(): record skipped (:93-94)           */
#define PMRD_FQ_NO_UMI 2     /* valid record, none of the three header patterns matched (:28-32)    */
#define PMRD_FQ_OK 3         /* valid record with a UMI of <= 12 letters                            */
#define PMRD_FQ_LONG_UMI 4   /* valid record, "UMI:" capture longer than 12 letters (not packable)  */
typedef struct pmrd_fastq_records {
    int32_t* status;    /* [R] PMRD_FQ_*                                                           */
    uint64_t* umi_key;  /* [R] letters 5 bit each (A=1), left-aligned in bits 63..4, length in 3..0 */
    int64_t* hdr_off;   /* [R] stripped header line (incl. '@')                                    */
    int32_t* hdr_len;
    int64_t* umi_off;   /* [R] the regex capture                                                   */
    int32_t* umi_len;
    int64_t* seq_off;   /* [R] stripped sequence line                                              */
    int32_t* seq_len;
    int64_t* qual_off;  /* [R] stripped quality line                                               */
    int32_t* qual_len;
    int64_t* qual_sum;  /* [R] sum of (byte - 33) over the quality line (:64)                      */
} pmrd_fastq_records;
/* *n_records = ceil(lines / 4); PMRD_ERR_CAPACITY (with *n_records set) if capacity is smaller. */
int32_t pmrd_fastq_scan(pmrd_ctx* ctx, const uint8_t* h_bytes, int64_t n_bytes, int64_t capacity,
                        pmrd_fastq_records* h_out, int64_t* n_records);

/* Families of the kept reads: group by UMI key, rows ordered by first appearance (dict order,
 * fastq.py:186-196).  h_qual_sum / h_qual_len / h_umi_key are per kept read, in file order. */
typedef struct pmrd_fastq_families {
    uint64_t* umi_key;     /* [F]                                                        */
    uint32_t* first_read;  /* [F] index (into the kept reads) of the family's first read   */
    uint32_t* best_read;   /* [F] first read with the highest mean quality (:213 argmax) */
    int64_t* family_size;  /* [F]                                                        */
    int64_t* qual_sum;     /* [F] over all bases of all reads (:210,:222)                */
    int64_t* n_bases;      /* [F]                                                        */
} pmrd_fastq_families;
int32_t pmrd_fastq_families_of(pmrd_ctx* ctx, const uint64_t* h_umi_key, const int64_t* h_qual_sum,
                               const int32_t* h_qual_len, int64_t n_reads, pmrd_fastq_families* h_out,
                               int64_t capacity, int64_t* n_families);

/* ---------------------------------------------------------------- K12: metrics ------- */
/* roc_auc_score / average_precision (src/precise_mrd/metrics.py:18-33, sklearn definitions:
 * trapezoidal AUC over the distinct thresholds, AP = sum (R_i - R_{i-1}) P_i; one class only ->
 * 0.5 / mean(y_true)) of (y_true, y_score) under n_boot resamples -- bootstrap_metric,
 * metrics.py:48-118.  h_indices [n_boot * n] are the resample rows exactly as the reference
 * draws them up-front (rng.choice(n, n, replace=True), :76-79); NULL: Philox draws keyed
 * (seed, stream_id, resample, draw).  h_point (optional) receives {auc, ap} of the sample itself. */
int32_t pmrd_bootstrap_metrics(pmrd_ctx* ctx, const uint8_t* h_y_true, const double* h_y_score, int64_t n,
                               const int64_t* h_indices, int64_t n_boot, uint64_t stream_id,
                               double* h_point /* [2] or NULL */, double* h_auc /* [n_boot] */,
                               double* h_ap /* [n_boot] */);

/* ---------------------------------------------------------------- K2: G-D families -- */
/* collapse_umis (src/precise_mrd/collapse.py:70-118): per (cell, family) Philox draws.
 * variant 2 = G-D v2 constants, 1 = G-D v1 ("collapse 2.py").  Rows of cell i are
 * written contiguously; row_offsets[C+1] (host) receives the CSR offsets.  Pass
 * h_out == NULL to only count (row_offsets still filled). */
typedef struct pmrd_gd_family_table {
    int32_t* cell;          /* [R] index into the input cell arrays                      */
    int32_t* family_id;     /* [R] family index within the cell (gaps where skipped)     */
    int32_t* family_size;   /* [R]                                                       */
    double* quality_score;  /* [R]                                                       */
    double* consensus_agreement; /* [R]                                                  */
    uint8_t* flags;         /* [R] bit0 passes_quality, bit1 passes_consensus, bit2 is_variant */
} pmrd_gd_family_table;
int32_t pmrd_collapse_families(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af,
                               const int64_t* h_n_families, const double* h_background_rate,
                               const double* h_mean_family_size, int64_t n_cells,
                               const pmrd_umi_params* params, int32_t variant,
                               int64_t* h_row_offsets, pmrd_gd_family_table* h_out, int64_t capacity);

/* ---------------------------------------------------------------- K10: error model -- */
/* fit_error_model (src/precise_mrd/error_model.py:139-186): rate_c = n_var/n_neg *
 * U(0.5,2) for 64 contexts; CI = percentiles 2.5/97.5 of n_boot binomial-equivalent
 * bootstrap resamples (Binom(n_neg, n_var/n_neg)/n_neg * modifier).  n_neg <= 10 uses
 * the reference's (0.5x, 2x) fallback; n_neg == 0 draws U(1e-5,1e-3).  stream_id
 * separates independent fits (e.g. the cell index in LoD grids). */
int32_t pmrd_error_model(pmrd_ctx* ctx, int64_t n_neg, int64_t n_var, int32_t n_boot,
                         uint64_t stream_id, double* h_rate64, double* h_ci_lo64, double* h_ci_hi64);

/* ---------------------------------------------------------------- read-level simulate */
/* G-B SyntheticReadGenerator.generate_reads_for_site (src/precise_mrd/io.py:79-123)
 * re-keyed on Philox (seed; stream=3, cell, family, read): family size
 * NegBinom(r=2, mean)+1 (family_model 0) or clip(Poisson(mean),1,max) (family_model 1),
 * 12-mer UMI, has_variant ~ Bernoulli(af), optional contamination flip, per-read allele
 * (0.9 concordance / 0.05 error), quality int(N(30,5)) in [10,40], strand, read_pos in
 * [10,150).  Index hopping (re-key group bits of Binom(n_reads, hop) reads) and barcode
 * collisions (re-key UMI bits of Binom(n_families, coll) families onto the previous
 * family's UMI) implement sim/contamination.py:230-278 at read level.
 * Reads are emitted tile-shuffled (Philox permutation) so that grouping is real work. */
typedef struct pmrd_read_sim_params {
    double mean_family_size;
    int32_t family_model;       /* 0 negative binomial r=2 (io.py:71-77), 1 Poisson clipped */
    int32_t max_family_size;
    double contamination_rate;  /* io.py:100-101 */
    double hop_rate;            /* contamination.py:230-254 */
    double collision_rate;      /* contamination.py:256-278 */
    int32_t shuffle;            /* 0: family-major order; 1: tile-shuffled; 2 (pmrd_simulate_reads*): the
                                 * cell-major order of the fused plan -- cell after cell, each cell window-
                                 * shuffled inside its own region -- so that order-sensitive results (weighted
                                 * sums, first-seen tie-breaks) can be checked against the plan read for read */
    int32_t chunk_log2;         /* pipeline plans: cells are streamed in chunks of <= 2^chunk_log2 reads (0 = 30) */
    int32_t engine;             /* pipeline plans: PMRD_ENGINE_MATERIALISED (0) reads are written, sorted and voted
                                 * (the headline path: the sort does real work on window-shuffled input);
                                 * PMRD_ENGINE_FUSED (1) reads are voted where they are drawn and never written
                                 * (eval/lod.py:125-150, sim/contamination.py:85-228 grids; results = the materialised
                                 * path on family-major read order; hop_rate must be 0) */
    int32_t reserved0;          /* must be 0 (the library's own per-cell-knobs switch) */
} pmrd_read_sim_params;
#define PMRD_ENGINE_MATERIALISED 0
#define PMRD_ENGINE_FUSED 1

/* Counts reads: h_read_offsets[C+1] CSR of reads per cell (family-major layout). */
int32_t pmrd_simulate_reads_count(pmrd_ctx* ctx, const int64_t* h_cell_id, const int64_t* h_n_families,
                                  int64_t n_cells, const pmrd_read_sim_params* sp,
                                  int64_t* h_read_offsets);
/* Generates reads into device buffers (n_reads = total from the count call).  group id of
 * cell i = group_base + i. */
int32_t pmrd_simulate_reads_dev(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af,
                                const int64_t* h_n_families, int64_t n_cells, uint32_t group_base,
                                const pmrd_read_sim_params* sp, uint64_t* d_keys, uint32_t* d_payload,
                                int64_t capacity, int64_t* n_reads);
int32_t pmrd_simulate_reads(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af,
                            const int64_t* h_n_families, int64_t n_cells, uint32_t group_base,
                            const pmrd_read_sim_params* sp, uint64_t* h_keys, uint32_t* h_payload,
                            int64_t capacity, int64_t* n_reads);

/* ---------------------------------------------------------------- fused pipeline ---- */
/* simulate -> collapse -> call for a batch of cells, reads never leaving the device:
 * the bench "step" and the engine under eval-lod / eval-lob / eval-contamination.
 * Per cell: k = kept consensus families carrying the alt allele, n = kept consensus
 * families; error rate from the negative-control cells (af <= neg_af_max) of the batch
 * exactly as fit_error_model/call_mrd do (rate = mean over 64 context rates);
 * p (two-sided Poisson/binomial), BH over the batch. */
typedef struct pmrd_pipeline_result {
    int64_t* n_reads;     /* [C] reads generated for the cell                    */
    int64_t* n_families;  /* [C] families formed (after UMI merging)             */
    int64_t* n_total;     /* [C] kept consensus families                         */
    int64_t* n_variants;  /* [C] of which alt                                    */
    double* p_value;      /* [C]                                                 */
    double* p_adjusted;   /* [C]                                                 */
    uint8_t* significant; /* [C]                                                 */
} pmrd_pipeline_result;

int32_t pmrd_pipeline_cells(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af,
                            const int64_t* h_n_families, int64_t n_cells,
                            const pmrd_read_sim_params* sp, const pmrd_umi_params* up,
                            double neg_af_max, int32_t test_type, int32_t alternative, double alpha,
                            pmrd_pipeline_result* h_out, double* h_mean_error_rate);

/* The call stage of the pipeline on its own, over per-cell counts: error rate from the negative
 * controls (mean over 64 contexts of (sum n_variants / sum n_total) * U(0.5,2), error_model.py:153-156,
 * call.py:178), tail p per cell (call.py:182-197), BH over the batch (call.py:200-212).  This is what
 * every rank runs after the one all-gather of per-cell counts of a grid sharded over several GPUs
 * (SURVEY.md 8e): the same kernels as pmrd_pipeline_cells, so the table equals the single-GPU one bit
 * for bit.  stream_id = the grid's first cell id (keys the 64 context modifiers). */
int32_t pmrd_call_counts(pmrd_ctx* ctx, const int64_t* h_n_total, const int64_t* h_n_variants,
                         const uint8_t* h_is_negative, int64_t n_cells, uint64_t stream_id,
                         int32_t test_type, int32_t alternative, double alpha, double* h_p_value,
                         double* h_p_adjusted, uint8_t* h_significant, double* h_mean_error_rate);

/* The same pipeline as a prepared plan: create() uploads the cell table, sizes and
 * allocates every buffer once; run() only enqueues kernels on the context's stream (no
 * allocation, no host<->device copy, no sync -- safe to time with events or capture in a
 * CUDA graph); results() copies the per-cell table back and synchronises.
 * h_totals3 = {reads generated, families formed, groups (cells) with reads}. */
typedef struct pmrd_plan pmrd_plan;

/* Optional per-cell knobs of a plan: the contamination stress sweep of sim/contamination.py:85-228 (index hopping x
 * barcode collisions x cross-sample contamination, each over rates x AF x depth x replicates) as ONE plan instead of one
 * pipeline per (rate, AF, depth).  Every pointer may be NULL (the plan-wide pmrd_read_sim_params value applies).
 *   hop_rate / hop_pool_start / hop_pool_size   a read is re-keyed, w.p. hop_rate[c], to another cell of its hop pool =
 *        cells [hop_pool_start[c], + hop_pool_size[c]) of this call (the samples that share a flow-cell lane:
 *        contamination.py:230-254).  Cells with hop_rate > 0 take the general path (full-key sort); pools must be
 *        contiguous and are never split over chunks.
 *   collision_rate      a molecule takes the previous molecule's UMI w.p. collision_rate[c] (contamination.py:256-278)
 *   cross_proportion / cross_af   int(n_families[c] * cross_proportion[c]) EXTRA molecules are mixed into the cell from a
 *        contaminating sample of allele fraction cross_af[c] (contamination.py:280-311: n_contam = int(n * proportion),
 *        contam_af = min(0.1, 10 af)); they carry their own UMIs and reads and are collapsed with the cell's own
 *   call_group_start[G + 1]   cells [s[g], s[g+1]) are CALLED together: error rate from the group's own negative
 *        controls, BH inside the group -- what one call_mrd of the reference sees (call.py:146-229).  NULL: one group.
 *        Groups hold at most 4096 cells when G > 1. */
typedef struct pmrd_cell_knobs {
    const double* hop_rate;
    const int64_t* hop_pool_start;
    const int64_t* hop_pool_size;
    const double* collision_rate;
    const double* cross_proportion;
    const double* cross_af;
    const int64_t* call_group_start;
    int64_t n_call_groups;
} pmrd_cell_knobs;
int32_t pmrd_plan_create_ex(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af,
                            const int64_t* h_n_families, int64_t n_cells, const pmrd_read_sim_params* sp,
                            const pmrd_umi_params* up, double neg_af_max, int32_t test_type,
                            int32_t alternative, double alpha, const pmrd_cell_knobs* knobs, pmrd_plan** out);
int32_t pmrd_plan_create(pmrd_ctx* ctx, const int64_t* h_cell_id, const double* h_af,
                         const int64_t* h_n_families, int64_t n_cells, const pmrd_read_sim_params* sp,
                         const pmrd_umi_params* up, double neg_af_max, int32_t test_type,
                         int32_t alternative, double alpha, pmrd_plan** out);
/* mean error rate of every call group of the last run (h_rates[capacity >= G]); synchronises */
int32_t pmrd_plan_group_rates(pmrd_ctx* ctx, pmrd_plan* plan, double* h_rates, int64_t capacity);
int32_t pmrd_plan_run(pmrd_ctx* ctx, pmrd_plan* plan);
int32_t pmrd_plan_results(pmrd_ctx* ctx, pmrd_plan* plan, pmrd_pipeline_result* h_out,
                          double* h_mean_error_rate, int64_t* h_totals3);
int64_t pmrd_plan_n_reads(const pmrd_plan* plan);
int64_t pmrd_plan_n_families(const pmrd_plan* plan);
int64_t pmrd_plan_n_chunks(const pmrd_plan* plan);
int64_t pmrd_plan_chunk_reads(const pmrd_plan* plan, int64_t chunk);
/* Byte-accounting facts for roofline reporting: out[0] reads generated, [1] record slots
 * written (reads + tile padding), [2] families drawn, [3] chunks, [4] radix passes per chunk,
 * [5] bytes per read record in the sort (8 packed / 12 key+payload), [6] largest chunk (slots),
 * [7] 1 if the cell-major (segmented) path is used. */
int32_t pmrd_plan_info(const pmrd_plan* plan, int64_t* out8);
void pmrd_plan_destroy(pmrd_plan* plan);
/* Stage-by-stage entry points, for per-stage timing (bench.py): the stage for ONE chunk
 * (chunks share the read buffers, so stages of different chunks cannot be separated). */
int32_t pmrd_plan_run_simulate(pmrd_ctx* ctx, pmrd_plan* plan, int64_t chunk); /* sizes + scan + reads        */
int32_t pmrd_plan_run_collapse(pmrd_ctx* ctx, pmrd_plan* plan, int64_t chunk); /* sort + families + cell counts */
int32_t pmrd_plan_run_call(pmrd_ctx* ctx, pmrd_plan* plan);                    /* error rate + p + BH, all cells */

/* ---------------------------------------------------------------- K15: feature table */
/* FeatureEngineer.extract_features (src/precise_mrd/advanced_stats.py:33-83): the matrix the ML /
 * Bayesian callers train on, from the columns of the collapsed-family table.  h_out [n][6] row-major =
 * family_size, quality_score, consensus_agreement, allele_fraction, family_size / (quality_score + 1),
 * consensus_agreement * allele_fraction -- single IEEE operations, bit-equal to the pandas expressions. */
int32_t pmrd_feature_table(pmrd_ctx* ctx, const int64_t* h_family_size, const double* h_quality_score,
                           const double* h_consensus_agreement, const double* h_allele_fraction, int64_t n,
                           double* h_out);

/* ---------------------------------------------------------------- K16: QC reductions */
/* QCAnalyzer.analyze_family_size_distribution / analyze_consensus_efficiency (src/precise_mrd/qc.py:
 * 54-97, 185-215) over a family table: h_hist[min(size, n_bins-1)] = number of families of that size (the
 * last bin collects every larger one), h_counts2[0] = families whose flags meet flag_mask (e.g.
 * PMRD_FAM_HAS_CONSENSUS; h_flags may be NULL), h_counts2[1] = largest size.  Every statistic of the
 * report (count, mean, std, percentiles, skewness, kurtosis, Counter, below / above threshold) is exact
 * integer arithmetic on this histogram (precise_mrd/qc.py). */
int32_t pmrd_qc_family_stats(pmrd_ctx* ctx, const uint32_t* h_family_size, const uint32_t* h_flags, uint32_t flag_mask,
                             int64_t n, int32_t n_bins, uint64_t* h_hist, uint64_t* h_counts2);

/* ---------------------------------------------------------------- K11: LoD fit ------ */
/* LODAnalyzer._fit_detection_curve + _bootstrap_lod_ci (src/precise_mrd/eval/lod.py:
 * 351-408): 2-parameter logistic on log10(af) by Levenberg-Marquardt from (1,1), solved
 * at `target`; n_boot Philox resamples of the (af, hit) pairs refit in parallel.
 * out[0]=lod, out[1]=ci_lo (P2.5), out[2]=ci_hi (P97.5); NaN where the reference is NaN. */
int32_t pmrd_fit_lod(pmrd_ctx* ctx, const double* h_af, const double* h_hit, int32_t n, double target,
                     int32_t n_boot, uint64_t stream_id, double* h_out3);
/* The same for n_series detection curves over the same af grid in ONE launch (the per-depth loop of
 * LODAnalyzer.estimate_lod, eval/lod.py:147-160): h_hits [n_series][n], one stream id per curve,
 * h_out3 [n_series][3].  A fit is a latency-bound sequential iteration, so the curves of an analysis are
 * fitted side by side; results equal n_series separate pmrd_fit_lod calls bit for bit. */
int32_t pmrd_fit_lod_batch(pmrd_ctx* ctx, const double* h_af, const double* h_hits, int32_t n, int32_t n_series,
                           double target, int32_t n_boot, const uint64_t* h_stream_ids, double* h_out3);

#ifdef __cplusplus
}
#endif
#endif /* PMRD_H */
