"""Where the end-to-end step (precise_mrd.grid.run_cells on the C2 grid) spends host time.  Dev helper."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from precise_mrd import _native as nv  # noqa: E402
from precise_mrd import grid  # noqa: E402

cfg, label, cell_id, af, depth = bench.workload("default", 0, 1)
ctx = nv.Context(device=0, seed=int(cfg.seed))
sim = grid.read_sim_params(cfg, mean_family_size=5.0, family_model=1)
for _ in range(2):
    grid.run_cells(cfg, cell_id, af, depth, ctx=ctx, sim=sim)
t0 = time.perf_counter()
plan = nv.Plan(ctx, cell_id, af, depth, sim, grid.umi_params(cfg))
ctx.sync()
t1 = time.perf_counter()
plan.run()
ctx.sync()
t2 = time.perf_counter()
res = plan.results()
t3 = time.perf_counter()
plan.close()
t4 = time.perf_counter()
print(f"plan create {1e3*(t1-t0):.2f} ms, run {1e3*(t2-t1):.2f} ms, results {1e3*(t3-t2):.2f} ms, close {1e3*(t4-t3):.2f} ms")
pr = cProfile.Profile()
pr.enable()
grid.run_cells(cfg, cell_id, af, depth, ctx=ctx, sim=sim)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
