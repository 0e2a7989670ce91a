"""Two B200s, one process per GPU over NCCL: the eval-lod grid sharded across the ranks must give
the single-GPU table bit for bit.  Skipped on a one-GPU box; the sharding and
gathering logic itself is covered on CPU ranks (gloo) in test_distributed_cpu.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _grid():
    from precise_mrd.grid import grid_cells

    cid, af, depth, _ = grid_cells([0.0, 1e-4, 1e-3, 5e-3, 1e-2], [1000, 5000, 20000], 6)
    return cid, af, depth


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    from precise_mrd import _native as nv
    from precise_mrd.config import load_config
    from precise_mrd.distributed import run_cells_sharded
    from precise_mrd.eval import LODAnalyzer

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = load_config(ROOT / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs" / "smoke.yaml")
    cid, af, depth = _grid()
    out = run_cells_sharded(cfg, cid, af, depth, neg_af_max=0.0, alternative=nv.ALT_GREATER)
    an = LODAnalyzer(cfg, np.random.default_rng(7), engine="grid")
    lod = an.estimate_lod(depth_values=[1000, 5000], n_replicates=8)
    q.put((rank, {k: np.asarray(v).tolist() for k, v in out.items() if k != "negative_control"},
           {str(d): r["hit_rates"] + [r["lod_af"]] for d, r in lod["depth_results"].items()}))
    dist.destroy_process_group()


def test_lod_grid_on_two_gpus_equals_one_gpu(native):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    from precise_mrd import grid
    from precise_mrd.config import load_config
    from precise_mrd.eval import LODAnalyzer

    world, port = 2, _free_port()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg = load_config(ROOT / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs" / "smoke.yaml")
    cid, af, depth = _grid()
    one = grid.run_cells(cfg, cid, af, depth, neg_af_max=0.0, alternative=native.ALT_GREATER)
    lod1 = LODAnalyzer(cfg, np.random.default_rng(7), engine="grid").estimate_lod(depth_values=[1000, 5000], n_replicates=8)
    want_lod = {str(d): r["hit_rates"] + [r["lod_af"]] for d, r in lod1["depth_results"].items()}
    for rank, out, lod in res:
        for k in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
            assert (np.asarray(out[k]) == one[k]).all(), (rank, k)
        assert out["mean_error_rate"] == one["mean_error_rate"]
        assert str(lod) == str(want_lod)          # (NaN-safe: an all-zero hit vector fits to NaN on both)
