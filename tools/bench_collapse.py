"""Throughput of the general collapse path (`pmrd_collapse_reads`: external reads in arbitrary order, 12-byte
records, full-key sort, family + group tables) -- the reference's own benchmark shape (1e6 reads, SURVEY.md §6:
grouping 2.3 s + consensus 1.8 s on its CPU path) and a larger one.  Secondary evidence; not bench.py's line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
from precise_mrd import _native as nv  # noqa: E402

ctx = nv.Context(device=0, seed=5)
P = nv.UmiParams(min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6,
                 mode=nv.MODE_WEIGHTED, cluster=nv.CLUSTER_EXACT, max_distance=0, umi_len=12)
out = []
for n, n_loci in ((1_000_000, 1000), (50_000_000, 100_000)):
    rng = np.random.default_rng(n)
    fam = rng.integers(0, n // 5, n)
    umi = (fam.astype(np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFF)
    keys = ((fam % n_loci).astype(np.uint64) << np.uint64(32)) | umi
    q = rng.integers(10, 41, n).astype(np.uint32)
    pl = (rng.random(n) < 0.03).astype(np.uint32) | (q << np.uint32(2))
    ctx.collapse_reads(keys, pl, P)                      # warm-up at full size: the stream-ordered pool grows once
    t0 = time.perf_counter()
    fam_t, grp_t = ctx.collapse_reads(keys, pl, P)
    dt = time.perf_counter() - t0
    ctx.profile_enable(True)
    ctx.collapse_reads(keys, pl, P)
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    k_ms = sum(v[1] for v in prof.values())
    out.append({"reads": n, "loci": n_loci, "families": int(len(fam_t["size"])), "end_to_end_s": dt, "reads_per_s_end_to_end": n / dt,
                "kernels_ms": k_ms, "reads_per_s_kernels": n / (k_ms * 1e-3),
                "top_kernels_ms": {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:4]}})
print(json.dumps(out))
