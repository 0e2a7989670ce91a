#!/usr/bin/env python
"""Dev probe: per-kernel times of one plan step with the generator's window shuffle on / off."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "precise-mrd-mini_b200"), str(ROOT)]
import torch  # noqa: E402

import bench  # noqa: E402
from precise_mrd import _native as nv, grid  # noqa: E402

cfg, label, cell_id, af, depth = bench.workload("default", 0, 1)
ctx = nv.Context(device=0, seed=int(cfg.seed))
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for shuffle in (True, False):
    sim = grid.read_sim_params(cfg, mean_family_size=5.0, family_model=1, chunk_log2=0, shuffle=shuffle)
    plan = nv.Plan(ctx, cell_id, af, depth, sim, grid.umi_params(cfg), neg_af_max=1e-4, test_type=0, alpha=0.05)
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    plan.run()
    torch.cuda.synchronize()
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    print("shuffle", shuffle, {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:7]})
    plan.close()
