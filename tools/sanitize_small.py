"""A small pass over every hot kernel family, sized for compute-sanitizer (racecheck / memcheck): a cell-major plan with
multi-tile cells (staged + direct scatter, counting, cell scan, consensus), a compact-layout plan (small_cell_kernel), the
fused engine (filters, stash, replay), the general collapse and the group-major collapse."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
from precise_mrd import _native as nv  # noqa: E402

ctx = nv.Context(device=0, seed=3)
up = nv.UmiParams(min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6, mode=nv.MODE_WEIGHTED,
                  cluster=nv.CLUSTER_EXACT, max_distance=0, umi_len=12)


def sp(**kw):
    d = dict(mean_family_size=5.0, family_model=1, max_family_size=1000, contamination_rate=0.0, hop_rate=0.0, collision_rate=0.0,
             shuffle=1, chunk_log2=0)
    d.update(kw)
    return nv.ReadSimParams(**d)


cid = np.arange(6, dtype=np.int64) + 50
af = np.array([0.0, 0.0, 0.01, 0.05, 0.2, 0.001])
r1 = ctx.pipeline_cells(cid, af, np.array([3000, 900, 2500, 1200, 40, 5000]), sp(), up)                      # tile-padded cells
r2 = ctx.pipeline_cells(cid, af, np.array([300, 90, 250, 120, 40, 500]), sp(), up)                           # compact layout
r3 = ctx.pipeline_cells(cid, af, np.array([3000, 900, 2500, 1200, 40, 5000]), sp(engine=nv.ENGINE_FUSED), up)
r4 = ctx.pipeline_cells(cid, af, np.array([3000, 900, 2500, 1200, 40, 5000]), sp(hop_rate=0.02), up)         # general path
assert (r1["n_total"] == r3["n_total"]).all() and (r1["n_variants"] == r3["n_variants"]).all()
rng = np.random.default_rng(1)
n = 60000
keys = (rng.integers(0, 40, n, dtype=np.uint64) << np.uint64(32)) | rng.integers(0, 3000, n, dtype=np.uint64)
pl = (rng.integers(0, 4, n).astype(np.uint32)) | (rng.integers(10, 41, n).astype(np.uint32) << np.uint32(2))
ctx.collapse_reads(keys, pl, up)                                                                             # group-major path
keys2 = (rng.integers(0, 3, n, dtype=np.uint64) << np.uint64(32)) | rng.integers(0, 3000, n, dtype=np.uint64)
ctx.collapse_reads(keys2, pl, up)                                                                            # groups > 4096 reads: full-key sort
print("sanitize_small ok", int(r1["n_total"].sum()), int(r2["n_total"].sum()), int(r4["n_total"].sum()))
