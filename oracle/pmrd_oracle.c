/*
 * oracle/pmrd_oracle.c -- plain-C restatement of the reference's read-level hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Used as (a) a second, independent checker of the collapse
 * arithmetic (tests/test_oracle_golden.py pins it to the reference-generated G-B fixture) and
 * (b) the CPU baseline bench.py times on the GPU box's host cores ("kind": "port").  The
 * product never links or loads this file.
 *
 * What it restates (paths under /root/reference):
 *   generate   SyntheticReadGenerator.generate_reads_for_site   src/precise_mrd/io.py:79-123
 *   group      UMIProcessor.group_umis_by_site / exact UMI dict  src/precise_mrd/umi.py:97-115
 *              (== group_umis_fast, rust_ext/src/lib.rs:7-22: a hash map of read-index lists)
 *   consensus  UMIProcessor.call_consensus                       src/precise_mrd/umi.py:146-179
 *   filter     UMIProcessor.filter_family_size                   src/precise_mrd/umi.py:181-203
 *   count      UMIProcessor.get_consensus_counts                 src/precise_mrd/umi.py:241-266
 *   test       poisson_test (two-sided)                          src/precise_mrd/call.py:14-25
 *   FDR        benjamini_hochberg_correction                     src/precise_mrd/call.py:41-79
 * The reference is single-threaded Python; this port is C with OpenMP over cells, i.e. a
 * deliberately STRONG baseline.  RNG: splitmix64-seeded xoshiro256** per cell (the reference
 * uses MT19937; the draws are distributionally the same, not bit-equal).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FAM_HAS_CONSENSUS 1u
#define FAM_KEPT 2u
#define FAM_OVERSIZE 4u

/* ------------------------------------------------------------------ rng ----------------- */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {
    uint64_t* s = r->s;
    const uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static inline void rng_seed(rng_t* r, uint64_t seed) {
    for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed);
}
static inline double rng_u(rng_t* r) { return ((double)(rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline double rng_normal(rng_t* r) { return sqrt(-2.0 * log(rng_u(r))) * cos(6.283185307179586 * rng_u(r)); }

static int draw_poisson(rng_t* r, double lam) {
    double u = rng_u(r), p = exp(-lam), F = p;
    int k = 0;
    while (u > F && k < 4096) { ++k; p *= lam / k; F += p; }
    return k;
}
static int draw_negbin_r2(rng_t* r, double p) { /* io.py:71-77 */
    double u = rng_u(r), q = 1.0 - p, f = p * p, F = f;
    int k = 0;
    while (u > F && k < 65535) { ++k; f *= q * (double)(k + 1) / (double)k; F += f; }
    return k;
}

/* ------------------------------------------------------------------ collapse ------------ */
/* One group's reads (umi[i], payload[i]), i in input order.  Exact-UMI hash grouping with
 * per-family read lists in insertion order, then weighted consensus + filter.
 * Outputs one record per family in FIRST-APPEARANCE order (dict order, as the reference). */
typedef struct {
    uint32_t umi, size, n_alt, consensus, qsum, n_best, flags;
} fam_rec;

static int64_t collapse_group(const uint32_t* umi, const uint32_t* payload, int64_t n, int min_fs, int max_fs, int qthr,
                              double cthr, const double* wtab, fam_rec* out, int32_t* work /* 2n + table */, int64_t table_size) {
    int32_t* next = work;             /* next read of the same family */
    int32_t* tail = work + n;         /* per family: last read */
    int32_t* table = work + 2 * n;    /* open addressing: family index + 1 */
    memset(table, 0, sizeof(int32_t) * (size_t)table_size);
    int32_t* head = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int64_t n_fam = 0;
    const uint64_t mask = (uint64_t)table_size - 1;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t h = ((uint64_t)umi[i] * 0x9E3779B97F4A7C15ull) >> 20;
        for (;;) {
            h &= mask;
            const int32_t e = table[h];
            if (e == 0) {
                table[h] = (int32_t)(n_fam + 1);
                head[n_fam] = (int32_t)i;
                tail[n_fam] = (int32_t)i;
                out[n_fam].umi = umi[i];
                ++n_fam;
                break;
            }
            if (out[e - 1].umi == umi[i]) {
                next[tail[e - 1]] = (int32_t)i;
                tail[e - 1] = (int32_t)i;
                break;
            }
            ++h;
        }
        next[i] = -1;
    }
    for (int64_t f = 0; f < n_fam; ++f) {
        double w[4] = {0, 0, 0, 0}, total = 0.0;
        uint32_t qs[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0}, first[4] = {9, 9, 9, 9}, n_seen = 0, size = 0, n_alt = 0;
        for (int32_t i = head[f]; i >= 0; i = next[i]) {
            const uint32_t pl = payload[i];
            ++size;
            n_alt += pl >> 31;
            const uint32_t a = pl & 3u, q = (pl >> 2) & 63u;
            if ((int)q < qthr) continue;               /* umi.py:152 */
            const double wt = wtab[q];                 /* 10 ** (q / 10.0), umi.py:161 */
            w[a] += wt;
            total += wt;
            qs[a] += q;
            cnt[a] += 1;
            if (first[a] == 9) first[a] = n_seen++;
        }
        fam_rec* r = &out[f];
        r->size = size; r->n_alt = n_alt; r->consensus = 0; r->qsum = 0; r->n_best = 0; r->flags = 0;
        if (n_seen > 0 && total != 0.0) {
            int best = -1;
            for (int a = 0; a < 4; ++a) {              /* max(): first inserted wins ties */
                if (first[a] == 9) continue;
                if (best < 0 || w[a] > w[best] || (w[a] == w[best] && first[a] < first[best])) best = a;
            }
            if (!(w[best] / total < cthr)) {           /* umi.py:172 */
                r->flags |= FAM_HAS_CONSENSUS;
                r->consensus = (uint32_t)best;
                r->qsum = qs[best];
                r->n_best = cnt[best];
            }
        }
        if ((int)size >= min_fs) r->flags |= FAM_KEPT; /* umi.py:186, on the raw size */
        if ((int)size > max_fs) r->flags |= FAM_OVERSIZE;
    }
    free(head);
    return n_fam;
}

static int64_t pow2_at_least(int64_t x) {
    int64_t p = 16;
    while (p < x) p <<= 1;
    return p;
}

static int cmp_key(const void* a, const void* b) {
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* Whole-array entry point for the golden-vector test: keys = group:32|umi:32, any order.
 * Families are returned sorted by key (the order of the CUDA path). */
int64_t oracle_collapse_weighted(const uint64_t* keys, const uint32_t* payload, int64_t n, int min_fs, int max_fs,
                                 int qthr, double cthr, uint64_t* fam_key, uint32_t* fam_size, uint32_t* fam_nalt,
                                 uint32_t* fam_cons, uint32_t* fam_qsum, uint32_t* fam_nbest, uint32_t* fam_flags) {
    double wtab[64];
    for (int q = 0; q < 64; ++q) wtab[q] = pow(10.0, (double)q / 10.0);
    /* bucket reads by group, keeping input order */
    uint64_t* gids = (uint64_t*)malloc(8 * (size_t)(n + 1));
    for (int64_t i = 0; i < n; ++i) gids[i] = keys[i] >> 32;
    uint64_t* ug = (uint64_t*)malloc(8 * (size_t)(n + 1));
    memcpy(ug, gids, 8 * (size_t)n);
    qsort(ug, (size_t)n, 8, cmp_key);
    int64_t n_groups = 0;
    for (int64_t i = 0; i < n; ++i)
        if (i == 0 || ug[i] != ug[i - 1]) ug[n_groups++] = ug[i];
    int64_t total_f = 0;
    uint32_t* u = (uint32_t*)malloc(4 * (size_t)(n + 1));
    uint32_t* p = (uint32_t*)malloc(4 * (size_t)(n + 1));
    fam_rec* rec = (fam_rec*)malloc(sizeof(fam_rec) * (size_t)(n + 1));
    uint64_t* order = (uint64_t*)malloc(8 * (size_t)(n + 1));
    for (int64_t g = 0; g < n_groups; ++g) {
        int64_t m = 0;
        for (int64_t i = 0; i < n; ++i)
            if (gids[i] == ug[g]) { u[m] = (uint32_t)keys[i]; p[m] = payload[i]; ++m; }
        const int64_t ts = pow2_at_least(2 * m);
        int32_t* work = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * m + ts));
        const int64_t nf = collapse_group(u, p, m, min_fs, max_fs, qthr, cthr, wtab, rec, work, ts);
        free(work);
        /* sort this group's families by umi */
        for (int64_t f = 0; f < nf; ++f) order[f] = ((uint64_t)rec[f].umi << 32) | (uint64_t)f;
        qsort(order, (size_t)nf, 8, cmp_key);
        for (int64_t j = 0; j < nf; ++j) {
            const fam_rec* r = &rec[order[j] & 0xffffffffu];
            fam_key[total_f] = (ug[g] << 32) | r->umi;
            fam_size[total_f] = r->size; fam_nalt[total_f] = r->n_alt; fam_cons[total_f] = r->consensus;
            fam_qsum[total_f] = r->qsum; fam_nbest[total_f] = r->n_best; fam_flags[total_f] = r->flags;
            ++total_f;
        }
    }
    free(gids); free(ug); free(u); free(p); free(rec); free(order);
    return total_f;
}

/* External reads of a panel (dense group ids < n_ids, any order): the group table of UMIProcessor.process_reads +
 * get_consensus_counts (umi.py:205-266).  Reads are bucketed by group in input order (a counting sort of the read
 * indices), then every group goes through collapse_group, groups spread over the OpenMP team.  This is the CPU
 * baseline of bench.py --workload external and a checker of the device group table at sizes the whole-array
 * entry point above (O(groups x reads)) cannot reach. */
int64_t oracle_collapse_groups(const uint64_t* keys, const uint32_t* payload, int64_t n, int64_t n_ids, int min_fs, int max_fs,
                               int qthr, double cthr, uint32_t* family_count /*[n_ids]: kept*/, uint32_t* consensus_count /*[n_ids*4]*/,
                               uint32_t* families_formed /*[n_ids]*/, int n_threads) {
    double wtab[64];
    for (int q = 0; q < 64; ++q) wtab[q] = pow(10.0, (double)q / 10.0);
    int64_t* start = (int64_t*)calloc((size_t)(n_ids + 1), 8);
    for (int64_t i = 0; i < n; ++i) start[(keys[i] >> 32) + 1]++;
    for (int64_t g = 0; g < n_ids; ++g) start[g + 1] += start[g];
    int64_t* cur = (int64_t*)malloc(8 * (size_t)(n_ids + 1));
    memcpy(cur, start, 8 * (size_t)(n_ids + 1));
    uint32_t* u = (uint32_t*)malloc(4 * (size_t)(n + 1));
    uint32_t* p = (uint32_t*)malloc(4 * (size_t)(n + 1));
    for (int64_t i = 0; i < n; ++i) {              /* stable: input order inside a group */
        const int64_t d = cur[keys[i] >> 32]++;
        u[d] = (uint32_t)keys[i];
        p[d] = payload[i];
    }
    int64_t total = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : total)
    {
        int64_t cap = 0;
        fam_rec* rec = NULL;
        int32_t* work = NULL;
#pragma omp for schedule(dynamic, 64)
        for (int64_t g = 0; g < n_ids; ++g) {
            const int64_t m = start[g + 1] - start[g];
            family_count[g] = 0; families_formed[g] = 0;
            consensus_count[4 * g] = consensus_count[4 * g + 1] = consensus_count[4 * g + 2] = consensus_count[4 * g + 3] = 0;
            if (m == 0) continue;
            const int64_t ts = pow2_at_least(2 * m);
            if (m > cap) {
                cap = 2 * m;
                free(rec); free(work);
                rec = (fam_rec*)malloc(sizeof(fam_rec) * (size_t)(cap + 1));
                work = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * cap + pow2_at_least(2 * cap)));
            }
            const int64_t nf = collapse_group(u + start[g], p + start[g], m, min_fs, max_fs, qthr, cthr, wtab, rec, work, ts);
            families_formed[g] = (uint32_t)nf;
            total += nf;
            for (int64_t f = 0; f < nf; ++f) {
                if (!(rec[f].flags & FAM_KEPT)) continue;         /* umi.py:186 */
                family_count[g]++;
                if (rec[f].flags & FAM_HAS_CONSENSUS) consensus_count[4 * g + (rec[f].consensus & 3u)]++;   /* umi.py:246-251 */
            }
        }
        free(rec); free(work);
    }
    free(start); free(cur); free(u); free(p);
    return total;
}

/* ------------------------------------------------------------------ statistics ---------- */
/* Regularised incomplete gamma by series / continued fraction (the textbook pair the
 * reference reaches through scipy.special.pdtr / pdtrc). */
static double gamma_p_series(double a, double x) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 100000; ++n) {
        ap += 1.0; del *= x / ap; sum += del;
        if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return sum * exp(-x + a * log(x) - lgamma(a));
}
static double gamma_q_cf(double a, double x) {
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b; if (fabs(d) < tiny) d = tiny;
        c = b + an / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return exp(-x + a * log(x) - lgamma(a)) * h;
}
static void gamma_pq(double a, double x, double* P, double* Q) {
    if (x <= 0.0) { *P = 0.0; *Q = 1.0; return; }
    if (x < a + 1.0) { *P = gamma_p_series(a, x); *Q = 1.0 - *P; }
    else { *Q = gamma_q_cf(a, x); *P = 1.0 - *Q; }
}
double oracle_poisson_two_sided(int64_t k, double mu) { /* call.py:14-25 */
    if (!(mu > 0.0)) return 1.0;
    double P, Q, left, right;
    gamma_pq((double)k + 1.0, mu, &P, &Q);   /* cdf(k) = Q(k+1, mu) */
    left = Q;
    if (k <= 0) right = 1.0;
    else { gamma_pq((double)k, mu, &P, &Q); right = P; } /* sf(k-1) = P(k, mu) */
    const double p = 2.0 * (left < right ? left : right);
    return p < 1.0 ? p : 1.0;
}

typedef struct { double p; int64_t i; } pi_t;
static int cmp_pi(const void* a, const void* b) {
    const pi_t* x = (const pi_t*)a; const pi_t* y = (const pi_t*)b;
    if (x->p < y->p) return -1;
    if (x->p > y->p) return 1;
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}
void oracle_bh(const double* p, int64_t m, double alpha, uint8_t* rejected, double* q) { /* call.py:41-79 */
    if (m <= 0) return;
    pi_t* s = (pi_t*)malloc(sizeof(pi_t) * (size_t)m);
    for (int64_t i = 0; i < m; ++i) { s[i].p = p[i]; s[i].i = i; rejected[i] = 0; }
    qsort(s, (size_t)m, sizeof(pi_t), cmp_pi);
    for (int64_t i = m - 1; i >= 0; --i)
        if (s[i].p <= (double)(i + 1) / (double)m * alpha) {
            for (int64_t j = 0; j <= i; ++j) rejected[s[j].i] = 1;
            break;
        }
    double run = INFINITY;
    for (int64_t i = m - 1; i >= 0; --i) {
        double a = s[i].p * (double)m / (double)(i + 1);
        if (a < run) run = a;
        double v = run < 0.0 ? 0.0 : (run > 1.0 ? 1.0 : run);
        q[s[i].i] = v;
    }
    free(s);
}

/* ------------------------------------------------------------------ pipeline ------------ */
/* simulate -> collapse -> count for n_cells cells (one site per cell), OpenMP over cells.
 * Returns the number of reads generated.  n_total / n_variants as the CUDA pipeline defines
 * them: kept families with a consensus, and those whose consensus is the cell's alt allele. */
int64_t oracle_pipeline_cells(const int64_t* cell_id, const double* af, const int64_t* n_families, int64_t n_cells,
                              uint64_t seed, double mean_family_size, int family_model, int max_family_size, int min_fs,
                              int qthr, double cthr, int64_t* n_total, int64_t* n_variants, int64_t* n_reads_cell,
                              int n_threads) {
    double wtab[64];
    for (int q = 0; q < 64; ++q) wtab[q] = pow(10.0, (double)q / 10.0);
    int64_t total_reads = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_reads)
    for (int64_t c = 0; c < n_cells; ++c) {
        rng_t r;
        rng_seed(&r, seed ^ ((uint64_t)cell_id[c] * 0xD1342543DE82EF95ull));
        const int64_t nf = n_families[c] > 0 ? n_families[c] : 0;
        const uint32_t ref = (uint32_t)(cell_id[c] & 3), alt = (ref + 1u + (uint32_t)((cell_id[c] >> 2) % 3)) & 3u;
        int64_t cap = (int64_t)((double)nf * (mean_family_size + 2.0) * 1.3) + 1024, n = 0;
        uint32_t* umi = (uint32_t*)malloc(4 * (size_t)cap);
        uint32_t* pl = (uint32_t*)malloc(4 * (size_t)cap);
        for (int64_t f = 0; f < nf; ++f) {
            int size;
            if (family_model == 0) size = draw_negbin_r2(&r, 2.0 / (2.0 + mean_family_size)) + 1;
            else { size = draw_poisson(&r, mean_family_size); if (size < 1) size = 1; }
            if (size > max_family_size) size = max_family_size;
            const uint32_t u = (uint32_t)(rng_next(&r) >> 40);           /* 12 random bases, io.py:60-63 */
            const int has_variant = rng_u(&r) < af[c];                    /* io.py:97 */
            if (n + size > cap) {
                cap = (n + size) * 2;
                umi = (uint32_t*)realloc(umi, 4 * (size_t)cap);
                pl = (uint32_t*)realloc(pl, 4 * (size_t)cap);
            }
            for (int j = 0; j < size; ++j) {
                const double ua = rng_u(&r);
                uint32_t allele;
                if (has_variant) allele = ua < 0.9 ? alt : ref;           /* io.py:105-108 */
                else allele = ua < 0.95 ? ref : alt;
                int q = (int)(30.0 + 5.0 * rng_normal(&r));               /* io.py:65-69 */
                q = q < 10 ? 10 : (q > 40 ? 40 : q);
                const uint64_t w = rng_next(&r);
                const uint32_t strand = (uint32_t)(w & 1), rpos = 10u + (uint32_t)((w >> 1) % 140u);
                umi[n] = u;
                pl[n] = allele | ((uint32_t)q << 2) | (strand << 8) | (rpos << 9) | ((allele == alt ? 1u : 0u) << 31);
                ++n;
            }
        }
        const int64_t ts = pow2_at_least(2 * n);
        int32_t* work = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * n + ts));
        fam_rec* rec = (fam_rec*)malloc(sizeof(fam_rec) * (size_t)(n + 1));
        const int64_t got = collapse_group(umi, pl, n, min_fs, max_family_size, qthr, cthr, wtab, rec, work, ts);
        int64_t nt = 0, nv = 0;
        for (int64_t f = 0; f < got; ++f) {                                /* umi.py:241-266 */
            if ((rec[f].flags & (FAM_KEPT | FAM_HAS_CONSENSUS)) == (FAM_KEPT | FAM_HAS_CONSENSUS)) {
                ++nt;
                nv += rec[f].consensus == alt;
            }
        }
        n_total[c] = nt;
        n_variants[c] = nv;
        if (n_reads_cell) n_reads_cell[c] = n;
        total_reads += n;
        free(work); free(rec); free(umi); free(pl);
    }
    return total_reads;
}

/* error rate of fit_error_model + call_mrd (error_model.py:153-156, call.py:178) and the
 * per-cell two-sided Poisson test + BH: the "call" leg of the CPU baseline. */
void oracle_call_cells(const int64_t* n_total, const int64_t* n_variants, const uint8_t* is_negative, int64_t n_cells,
                       uint64_t seed, double alpha, double* mean_rate_out, double* p, double* q, uint8_t* significant) {
    int64_t nn = 0, vv = 0;
    for (int64_t i = 0; i < n_cells; ++i)
        if (is_negative[i]) { nn += n_total[i]; vv += n_variants[i]; }
    const double vr = nn > 0 ? (double)vv / (double)nn : 0.0;
    rng_t r;
    rng_seed(&r, seed ^ 0xABCDEF12345ull);
    double acc = 0.0;
    for (int c = 0; c < 64; ++c) acc += vr * (0.5 + 1.5 * rng_u(&r));
    const double rate = acc / 64.0;
    *mean_rate_out = rate;
    for (int64_t i = 0; i < n_cells; ++i) p[i] = oracle_poisson_two_sided(n_variants[i], (double)n_total[i] * rate);
    oracle_bh(p, n_cells, alpha, significant, q);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
