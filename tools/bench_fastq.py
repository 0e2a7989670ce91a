"""Throughput of the FASTQ ingest kernels (K13) on a synthetic file: records/s and file GB/s through
``Context.fastq_scan`` (host bytes in, host record table out) with the kernel share from the profiler.
Secondary evidence for SURVEY.md §8f-1; not part of bench.py's contract line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
from precise_mrd import _native as nv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(1)
L = 100
umis = rng.integers(0, 4, (n // 5 + 1, 12), dtype=np.uint8)
fam = rng.integers(0, len(umis), n)
bases = np.frombuffer(b"ACGT", dtype=np.uint8)
seq = bases[rng.integers(0, 4, (n, L), dtype=np.uint8)]
qual = (rng.integers(2, 41, (n, L)) + 33).astype(np.uint8)
rows = []
umi_txt = bases[umis[fam]]
for i in range(n):          # (host-side file synthesis: not timed)
    rows.append(b"@r%d UMI:%s\n%s\n+\n%s\n" % (i, umi_txt[i].tobytes(), seq[i].tobytes(), qual[i].tobytes()))
data = b"".join(rows)
ctx = nv.Context(device=0, seed=1)
rec = ctx.fastq_scan(data)
assert len(rec["umi_key"]) == n, (len(rec["umi_key"]), n)
ctx.profile_enable(True)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    rec = ctx.fastq_scan(data)
dt = (time.perf_counter() - t0) / reps
prof = ctx.profile_report()
ctx.profile_enable(False)
kern_ms = sum(v[1] for v in prof.values()) / reps
t0 = time.perf_counter()
fams = ctx.fastq_families(rec["umi_key"], rec["qual_sum"], rec["qual_len"])
dt_f = time.perf_counter() - t0
print(json.dumps({"records": n, "file_bytes": len(data), "scan_s_end_to_end": dt, "records_per_s_end_to_end": n / dt,
                  "file_GBps_end_to_end": len(data) / dt / 1e9, "scan_kernels_ms": kern_ms,
                  "file_GBps_kernels": len(data) / (kern_ms * 1e-3) / 1e9, "families": int(len(fams["family_size"])), "families_s": dt_f,
                  "kernels_ms": {k: round(v[1] / reps, 3) for k, v in prof.items()}}))
