// fastq.cu -- K13: FASTQ ingest (src/precise_mrd/fastq.py:15-220).
//
// The reference walks the file in Python: readline() x 4, three regexes on the header,
// ord(c) - 33 per quality character, a dict of lists keyed by UMI.  Here the file's bytes go to
// the device once and
//   fq_count / fq_fill    find the line ends (byte-parallel count, scan, fill);
//   fq_record_kernel      one thread per 4-line record: strip() each line, apply the reference's
//                         UMI patterns to the header IN ITS ORDER
//                             UMI:([A-Z]+)      _([A-Z]{8,12})_      :([A-Z]{8,12})$
//                         pack the UMI (5 bits per letter, <= 12 letters, length in the low
//                         nibble -- so equal keys <=> equal strings), sum the Phred+33 scores;
//   fq_family_*           after a stable radix sort of (UMI key, read index): one thread per
//                         family -- size, total quality, total bases, and the first read whose
//                         mean quality is the family's maximum (np.argmax, fastq.py:213).
// Record semantics follow read_fastq (fastq.py:66-110): records are consecutive groups of four
// lines; the first record whose stripped header is empty ends the file; a record with any other
// empty line is skipped.
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

#define FQ_CHUNK 64

__device__ __forceinline__ bool fq_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }
__device__ __forceinline__ bool fq_upper(uint8_t c) { return c >= 'A' && c <= 'Z'; }

// A thread owns one 64-byte chunk (the device buffer is 256-byte aligned, chunks are 64-aligned): four 16-byte loads,
// newlines found four bytes at a time with the SIMD byte compare (the byte loop was 64 loads per thread).
#define FQ_NL4 0x0a0a0a0au
__global__ void __launch_bounds__(256)
fq_count_kernel(const uint8_t* __restrict__ buf, int64_t n, uint32_t* __restrict__ counts) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t lo = c * FQ_CHUNK;
    if (lo >= n) return;
    uint32_t k = 0;
    if (lo + FQ_CHUNK <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(buf + lo);
#pragma unroll
        for (int j = 0; j < FQ_CHUNK / 16; ++j) {
            const uint4 v = p[j];
            k += __popc(__vcmpeq4(v.x, FQ_NL4) & 0x01010101u) + __popc(__vcmpeq4(v.y, FQ_NL4) & 0x01010101u) +
                 __popc(__vcmpeq4(v.z, FQ_NL4) & 0x01010101u) + __popc(__vcmpeq4(v.w, FQ_NL4) & 0x01010101u);
        }
    } else {
        for (int64_t i = lo; i < n; ++i) k += buf[i] == '\n';
    }
    counts[c] = k;
}

__global__ void __launch_bounds__(256)
fq_fill_kernel(const uint8_t* __restrict__ buf, int64_t n, const uint32_t* __restrict__ offs, int64_t* __restrict__ line_end) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t lo = c * FQ_CHUNK;
    if (lo >= n) return;
    uint32_t k = offs[c];
    if (lo + FQ_CHUNK <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(buf + lo);
#pragma unroll
        for (int j = 0; j < FQ_CHUNK / 16; ++j) {
            const uint4 v = p[j];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t m = __vcmpeq4(w[q], FQ_NL4) & 0x01010101u;     // bit 8b set <=> byte b is a newline
                while (m) {
                    const int b = (__ffs((int)m) - 1) >> 3;
                    m &= m - 1u;
                    line_end[k++] = lo + j * 16 + q * 4 + b;
                }
            }
        }
    } else {
        for (int64_t i = lo; i < n; ++i)
            if (buf[i] == '\n') line_end[k++] = i;
    }
}

struct FqLine { int64_t lo, hi; };   // stripped [lo, hi)

__device__ __forceinline__ FqLine fq_line(const uint8_t* __restrict__ buf, const int64_t* __restrict__ line_end, int64_t n_lines,
                                          int64_t n_bytes, int64_t k) {
    FqLine L{0, 0};
    if (k >= n_lines) return L;
    L.lo = k == 0 ? 0 : line_end[k - 1] + 1;
    L.hi = line_end[k] < n_bytes ? line_end[k] : n_bytes;
    while (L.lo < L.hi && fq_space(buf[L.lo])) ++L.lo;
    while (L.hi > L.lo && fq_space(buf[L.hi - 1])) --L.hi;
    return L;
}

// the three patterns of fastq.py:28-32, tried in order; returns the capture [*u_lo, *u_hi) or false
__device__ bool fq_extract_umi(const uint8_t* __restrict__ s, int64_t lo, int64_t hi, int64_t* u_lo, int64_t* u_hi) {
    // 1. UMI:([A-Z]+) -- leftmost "UMI:" that is followed by a capital letter; greedy run
    for (int64_t i = lo; i + 4 < hi; ++i)
        if (s[i] == 'U' && s[i + 1] == 'M' && s[i + 2] == 'I' && s[i + 3] == ':' && fq_upper(s[i + 4])) {
            int64_t e = i + 4;
            while (e < hi && fq_upper(s[e])) ++e;
            *u_lo = i + 4; *u_hi = e;
            return true;
        }
    // 2. _([A-Z]{8,12})_ -- leftmost '_' whose following run of capitals has 8..12 letters and ends in '_'
    for (int64_t i = lo; i < hi; ++i)
        if (s[i] == '_') {
            int64_t e = i + 1;
            while (e < hi && e - (i + 1) <= 12 && fq_upper(s[e])) ++e;
            const int64_t len = e - (i + 1);
            if (len >= 8 && len <= 12 && e < hi && s[e] == '_') { *u_lo = i + 1; *u_hi = e; return true; }
        }
    // 3. :([A-Z]{8,12})$ -- the header ends in ':' + 8..12 capitals
    {
        int64_t b = hi;
        while (b > lo && hi - b <= 12 && fq_upper(s[b - 1])) --b;
        const int64_t len = hi - b;
        if (len >= 8 && len <= 12 && b > lo && s[b - 1] == ':') { *u_lo = b; *u_hi = hi; return true; }
        // a longer run of capitals can still END in a valid capture if a ':' sits inside it? no: ':' is not a capital
    }
    return false;
}

__global__ void __launch_bounds__(128)
fq_record_kernel(const uint8_t* __restrict__ buf, int64_t n_bytes, const int64_t* __restrict__ line_end, int64_t n_lines,
                 int64_t n_records, pmrd_fastq_records out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_records) return;
    const FqLine h = fq_line(buf, line_end, n_lines, n_bytes, 4 * r), s = fq_line(buf, line_end, n_lines, n_bytes, 4 * r + 1),
                 p = fq_line(buf, line_end, n_lines, n_bytes, 4 * r + 2), q = fq_line(buf, line_end, n_lines, n_bytes, 4 * r + 3);
    int32_t status;
    uint64_t key = 0;
    int64_t qsum = 0, u_lo = 0, u_hi = 0;
    if (h.hi == h.lo) status = PMRD_FQ_END;                                          // fastq.py:85-86
    else if (s.hi == s.lo || p.hi == p.lo || q.hi == q.lo) status = PMRD_FQ_MALFORMED;   // :93-94
    else {
        status = PMRD_FQ_NO_UMI;
        if (fq_extract_umi(buf, h.lo, h.hi, &u_lo, &u_hi)) {
            const int64_t len = u_hi - u_lo;
            if (len > 12) status = PMRD_FQ_LONG_UMI;
            else {
                status = PMRD_FQ_OK;
                for (int64_t i = u_lo; i < u_hi; ++i) key = (key << 5) | (uint64_t)(buf[i] - 'A' + 1);
                key = (key << (5 * (12 - len))) << 4 | (uint64_t)len;             // left-aligned letters, then the length
            }
        }
        for (int64_t i = q.lo; i < q.hi; ++i) qsum += (int64_t)buf[i] - 33;          // :64 Phred+33
    }
    out.status[r] = status;
    out.umi_key[r] = key;
    out.hdr_off[r] = h.lo; out.hdr_len[r] = (int32_t)(h.hi - h.lo);
    out.umi_off[r] = u_lo; out.umi_len[r] = (int32_t)(u_hi - u_lo);
    out.seq_off[r] = s.lo; out.seq_len[r] = (int32_t)(s.hi - s.lo);
    out.qual_off[r] = q.lo; out.qual_len[r] = (int32_t)(q.hi - q.lo);
    out.qual_sum[r] = qsum;
}

int32_t k13_fastq_scan(pmrd_ctx* ctx, const uint8_t* d_buf, int64_t n_bytes, int64_t capacity, const pmrd_fastq_records* d_out,
                       int64_t* h_n_records) {
    *h_n_records = 0;
    if (n_bytes <= 0) return PMRD_OK;
    const int64_t n_chunks = (n_bytes + FQ_CHUNK - 1) / FQ_CHUNK;
    DevBuf counts, offs, tot, ends;
    PMRD_CUDA(ctx, counts.alloc(ctx->stream, (size_t)n_chunks * 4));
    PMRD_CUDA(ctx, offs.alloc(ctx->stream, (size_t)n_chunks * 4));
    PMRD_CUDA(ctx, tot.alloc(ctx->stream, 4));
    PMRD_ARG(ctx, n_bytes < (1ll << 40), "fastq: buffer too large");
    PMRD_LAUNCH(ctx, fq_count_kernel, pmrd_blocks(n_chunks, 256), 256, 0, d_buf, n_bytes, counts.as<uint32_t>());
    PMRD_TRY((exclusive_scan<uint32_t, uint32_t>(ctx, counts.as<uint32_t>(), offs.as<uint32_t>(), n_chunks, tot.as<uint32_t>())));
    uint32_t n_nl = 0;
    uint8_t last = 0;
    PMRD_CUDA(ctx, cudaMemcpyAsync(&n_nl, tot.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaMemcpyAsync(&last, d_buf + n_bytes - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t n_lines = (int64_t)n_nl + (last != '\n' ? 1 : 0);              // an unterminated last line still counts
    PMRD_CUDA(ctx, ends.alloc(ctx->stream, (size_t)(n_lines + 1) * 8));
    PMRD_LAUNCH(ctx, fq_fill_kernel, pmrd_blocks(n_chunks, 256), 256, 0, d_buf, n_bytes, (const uint32_t*)offs.as<uint32_t>(),
                ends.as<int64_t>());
    if (last != '\n') PMRD_CUDA(ctx, cudaMemcpyAsync(ends.as<int64_t>() + n_nl, &n_bytes, 8, cudaMemcpyHostToDevice, ctx->stream));
    const int64_t n_records = (n_lines + 3) / 4;
    *h_n_records = n_records;
    if (n_records > capacity) {
        snprintf(ctx->err, sizeof(ctx->err), "fastq: %lld records exceed the capacity %lld", (long long)n_records, (long long)capacity);
        return PMRD_ERR_CAPACITY;
    }
    if (n_records > 0)
        PMRD_LAUNCH(ctx, fq_record_kernel, pmrd_blocks(n_records, 128), 128, 0, d_buf, n_bytes, (const int64_t*)ends.as<int64_t>(),
                    n_lines, n_records, *d_out);
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                              // `n_bytes` above is read from the host stack
    return PMRD_OK;
}

// ---------------------------------------------------------------- families --------------
__global__ void __launch_bounds__(256)
fq_head_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
fq_start_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ excl, int64_t n, uint32_t* __restrict__ start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) start[excl[i]] = (uint32_t)i;
}
// one thread per family, reads in file order (the sort is stable)
__global__ void __launch_bounds__(128)
fq_family_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ read_of, const uint32_t* __restrict__ start,
                 uint32_t n_fam, int64_t n, const int64_t* __restrict__ qual_sum, const int32_t* __restrict__ qual_len,
                 pmrd_fastq_families out) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_fam) return;
    const int64_t lo = start[f], hi = f + 1 < n_fam ? (int64_t)start[f + 1] : n;
    int64_t tot = 0, bases = 0;
    double best = -1.0;
    uint32_t best_read = read_of[lo];
    for (int64_t i = lo; i < hi; ++i) {
        const uint32_t r = read_of[i];
        const int64_t qs = qual_sum[r], ql = qual_len[r];
        tot += qs; bases += ql;
        const double m = __ddiv_rn((double)qs, (double)ql);          // np.mean of the read's scores (fastq.py:105)
        if (m > best) { best = m; best_read = r; }                   // np.argmax: first maximum (fastq.py:213)
    }
    out.umi_key[f] = keys[lo];
    out.first_read[f] = read_of[lo];
    out.best_read[f] = best_read;
    out.family_size[f] = (int64_t)(hi - lo);
    out.qual_sum[f] = tot;
    out.n_bases[f] = bases;
}

__global__ void __launch_bounds__(256) fq_iota_kernel(uint32_t* __restrict__ v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) fq_widen_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
// out[i] = in[perm[i]] for the six family columns
__global__ void __launch_bounds__(256)
fq_permute_kernel(pmrd_fastq_families in, const uint32_t* __restrict__ perm, pmrd_fastq_families out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = perm[i];
    out.umi_key[i] = in.umi_key[j];
    out.first_read[i] = in.first_read[j];
    out.best_read[i] = in.best_read[j];
    out.family_size[i] = in.family_size[j];
    out.qual_sum[i] = in.qual_sum[j];
    out.n_bases[i] = in.n_bases[j];
}

// d_keys: UMI key of every kept read, in file order (clobbered).  Families come back ordered by
// their first read -- the insertion order of the reference's dict (fastq.py:186-196).
int32_t k13_fastq_families(pmrd_ctx* ctx, uint64_t* d_keys, int64_t n, const int64_t* d_qual_sum, const int32_t* d_qual_len,
                           const pmrd_fastq_families* d_out, int64_t capacity, int64_t* h_n_fam) {
    *h_n_fam = 0;
    if (n <= 0) return PMRD_OK;
    DevBuf kb, va, vb, flag, excl, tot, start;
    PMRD_CUDA(ctx, kb.alloc(ctx->stream, (size_t)n * 8));
    PMRD_CUDA(ctx, va.alloc(ctx->stream, (size_t)n * 4));
    PMRD_CUDA(ctx, vb.alloc(ctx->stream, (size_t)n * 4));
    PMRD_LAUNCH(ctx, fq_iota_kernel, pmrd_blocks(n, 256), 256, 0, va.as<uint32_t>(), n);
    uint64_t varying = 0;
    PMRD_TRY(radix_varying_bits(ctx, d_keys, n, &varying));
    int in_b = 0;
    PMRD_TRY(radix_sort_pairs(ctx, d_keys, va.as<uint32_t>(), kb.as<uint64_t>(), vb.as<uint32_t>(), n, varying, &in_b));
    const uint64_t* keys = in_b ? kb.as<uint64_t>() : d_keys;
    const uint32_t* read_of = in_b ? vb.as<uint32_t>() : va.as<uint32_t>();
    PMRD_CUDA(ctx, flag.alloc(ctx->stream, (size_t)n * 4));
    PMRD_CUDA(ctx, excl.alloc(ctx->stream, (size_t)n * 4));
    PMRD_CUDA(ctx, tot.alloc(ctx->stream, 4));
    PMRD_LAUNCH(ctx, fq_head_kernel, pmrd_blocks(n, 256), 256, 0, keys, n, flag.as<uint32_t>());
    PMRD_TRY((exclusive_scan<uint32_t, uint32_t>(ctx, flag.as<uint32_t>(), excl.as<uint32_t>(), n, tot.as<uint32_t>())));
    uint32_t n_fam = 0;
    PMRD_CUDA(ctx, cudaMemcpyAsync(&n_fam, tot.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PMRD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *h_n_fam = n_fam;
    if ((int64_t)n_fam > capacity) {
        snprintf(ctx->err, sizeof(ctx->err), "fastq: %u families exceed the capacity %lld", n_fam, (long long)capacity);
        return PMRD_ERR_CAPACITY;
    }
    const size_t F = n_fam;
    DevBuf fk, ff, fb, fs, fq, fn, oa, ob, pa, pb;
    PMRD_CUDA(ctx, start.alloc(ctx->stream, F * 4));
    PMRD_CUDA(ctx, fk.alloc(ctx->stream, F * 8)); PMRD_CUDA(ctx, ff.alloc(ctx->stream, F * 4)); PMRD_CUDA(ctx, fb.alloc(ctx->stream, F * 4));
    PMRD_CUDA(ctx, fs.alloc(ctx->stream, F * 8)); PMRD_CUDA(ctx, fq.alloc(ctx->stream, F * 8)); PMRD_CUDA(ctx, fn.alloc(ctx->stream, F * 8));
    pmrd_fastq_families t{fk.as<uint64_t>(), ff.as<uint32_t>(), fb.as<uint32_t>(), fs.as<int64_t>(), fq.as<int64_t>(), fn.as<int64_t>()};
    PMRD_LAUNCH(ctx, fq_start_kernel, pmrd_blocks(n, 256), 256, 0, (const uint32_t*)flag.as<uint32_t>(),
                (const uint32_t*)excl.as<uint32_t>(), n, start.as<uint32_t>());
    PMRD_LAUNCH(ctx, fq_family_kernel, pmrd_blocks(n_fam, 128), 128, 0, keys, read_of, (const uint32_t*)start.as<uint32_t>(), n_fam, n,
                d_qual_sum, d_qual_len, t);
    // families by first read: sort (first_read -> family) and permute the columns
    PMRD_CUDA(ctx, oa.alloc(ctx->stream, F * 8)); PMRD_CUDA(ctx, ob.alloc(ctx->stream, F * 8));
    PMRD_CUDA(ctx, pa.alloc(ctx->stream, F * 4)); PMRD_CUDA(ctx, pb.alloc(ctx->stream, F * 4));
    PMRD_LAUNCH(ctx, fq_widen_kernel, pmrd_blocks(n_fam, 256), 256, 0, (const uint32_t*)t.first_read, oa.as<uint64_t>(), (int64_t)n_fam);
    PMRD_LAUNCH(ctx, fq_iota_kernel, pmrd_blocks(n_fam, 256), 256, 0, pa.as<uint32_t>(), (int64_t)n_fam);
    PMRD_TRY(radix_varying_bits(ctx, oa.as<uint64_t>(), n_fam, &varying));
    PMRD_TRY(radix_sort_pairs(ctx, oa.as<uint64_t>(), pa.as<uint32_t>(), ob.as<uint64_t>(), pb.as<uint32_t>(), n_fam, varying, &in_b));
    PMRD_LAUNCH(ctx, fq_permute_kernel, pmrd_blocks(n_fam, 256), 256, 0, t, (const uint32_t*)(in_b ? pb.as<uint32_t>() : pa.as<uint32_t>()),
                *d_out, (int64_t)n_fam);
    return PMRD_OK;
}
