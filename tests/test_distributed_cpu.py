"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: cell sharding, the one all-gather,
and rank-independence of the assembled table."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_cells_partitions_and_balances():
    from precise_mrd.distributed import shard_cells

    depth = np.tile([1000, 5000, 10000, 50000, 100000], 375)          # the eval-lod grid: 1875 cells
    for world in (1, 2, 4, 8):
        shards = [shard_cells(depth, r, world) for r in range(world)]
        allc = np.sort(np.concatenate(shards))
        assert (allc == np.arange(len(depth))).all()                     # a partition
        loads = np.array([depth[s].sum() for s in shards], float)
        assert loads.max() / loads.mean() < 1.01                         # LPT: within 1 % of perfect balance
    assert (shard_cells(depth, 0, 2) == shard_cells(depth, 0, 2)).all()  # deterministic


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from precise_mrd.config import load_config
    from precise_mrd.distributed import run_cells_sharded
    from precise_mrd.grid import grid_cells

    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = load_config(ROOT / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs" / "smoke.yaml")
    cid, af, depth, _ = grid_cells([0.01, 0.001, 0.0001], [1000, 5000, 20000], 7)
    calls = []

    def fake_compute(c, a, f):          # stands in for the device plan: a pure function of the GLOBAL cell id
        calls.append(len(c))
        return {"n_total": (f * 8) // 10 + c % 7, "n_variants": (c * 13) % 11, "n_families": f - c % 3, "n_reads": f * 5 + c}

    out = run_cells_sharded(cfg, cid, af, depth, compute=fake_compute)
    q.put((rank, calls[0], out["n_total"].tolist(), out["n_variants"].tolist(), out["n_reads"].tolist(), out["negative_control"].tolist()))
    dist.destroy_process_group()


def test_sharded_grid_equals_single_rank_gloo():
    import torch.multiprocessing as mp

    from precise_mrd.grid import grid_cells

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cid, af, depth, _ = grid_cells([0.01, 0.001, 0.0001], [1000, 5000, 20000], 7)
    want_total = ((depth * 8) // 10 + cid % 7).tolist()
    for rank, n_mine, n_total, n_var, n_reads, neg in res:
        assert n_total == want_total                                   # every rank holds the full table
        assert n_var == ((cid * 13) % 11).tolist() and n_reads == (depth * 5 + cid).tolist()
        assert neg == (af <= 1e-4).tolist()
    assert res[0][1] + res[1][1] == len(cid) and min(res[0][1], res[1][1]) > 0   # the work really was split
