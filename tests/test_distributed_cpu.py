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


# ---------------------------------------------------------------- the exchange of external reads
class _NumpyExchangeOps:
    """Test double of the exchange's device steps (csrc/exchange.cu, the group-major collapse) on CPU tensors: NumPy
    for the cut and the re-keying, the C oracle for the per-group table."""

    def owner_partition(self, keys, payload, keys_tmp, payload_tmp, log2_world):
        k, p = keys.numpy().view(np.uint64), payload.numpy().view(np.uint32)
        dest = ((k >> np.uint64(32)) & np.uint64((1 << log2_world) - 1)).astype(np.int64)
        order = np.argsort(dest, kind="stable")
        keys_tmp.numpy().view(np.uint64)[:] = k[order]
        payload_tmp.numpy().view(np.uint32)[:] = p[order]
        return np.bincount(dest, minlength=1 << log2_world).tolist(), True

    def owner_rekey(self, keys, log2_world):
        k = keys.numpy().view(np.uint64)
        k[:] = (((k >> np.uint64(32)) >> np.uint64(log2_world)) << np.uint64(32)) | (k & np.uint64(0xFFFFFFFF))

    def collapse_groups(self, keys, payload, keys_tmp, payload_tmp, params, capacity_g):
        from oracle import cport

        k, p = keys.numpy().view(np.uint64), payload.numpy().view(np.uint32)
        og = cport.collapse_groups(k, p, capacity_g, *params)
        ids = np.unique((k >> np.uint64(32)).astype(np.int64))
        return og["total_families"], {"group": ids, "family_count": og["family_count"][ids], "consensus_count": og["consensus_count"][ids]}


def _striped_reads(rank, n, n_ids):
    rng = np.random.default_rng(100 + rank)
    g = rng.integers(0, n_ids, n, dtype=np.uint64)
    g[g % np.uint64(11) == 3] += np.uint64(1)                              # some ids stay empty
    umi = rng.integers(0, 40, n, dtype=np.uint64) * np.uint64(2654435761) % np.uint64(1 << 24)
    q = rng.choice([5, 10, 20, 30, 30, 37, 40], n).astype(np.uint32)
    pl = (rng.integers(0, 4, n).astype(np.uint32)) | (q << np.uint32(2))
    return (g << np.uint64(32)) | umi, pl


def _exchange_worker(rank, world, port, q, n_ids):
    sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from precise_mrd.distributed import collapse_reads_exchanged

    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys, pl = _striped_reads(rank, 30_000 + 1000 * rank, n_ids)
    tk, tp = torch.from_numpy(keys.view(np.int64).copy()), torch.from_numpy(pl.view(np.int32).copy())
    full = collapse_reads_exchanged(None, tk, tp, (3, 1000, 20, 0.6), n_ids, ops=_NumpyExchangeOps())
    tk, tp = torch.from_numpy(keys.view(np.int64).copy()), torch.from_numpy(pl.view(np.int32).copy())
    own = collapse_reads_exchanged(None, tk, tp, (3, 1000, 20, 0.6), n_ids, ops=_NumpyExchangeOps(), gather=False)
    q.put((rank, {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in full.items()},
           {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in own.items()}))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_exchanged_collapse_equals_the_collapse_of_the_concatenated_reads_gloo(world):
    """Reads of one sample striped over the ranks: owner cut -> all-to-all -> owner-side collapse must give the table of
    the ranks' reads concatenated in rank order (the family vote adds in read order, so the ORDER the runs arrive in
    matters), on every rank; with gather=False each rank holds exactly its own groups."""
    import torch.multiprocessing as mp

    sys.path.insert(0, str(ROOT))
    from oracle import cport

    n_ids, port = 1000, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q, n_ids)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = [_striped_reads(r, 30_000 + 1000 * r, n_ids) for r in range(world)]
    keys, pl = np.concatenate([a for a, _ in parts]), np.concatenate([b for _, b in parts])
    og = cport.collapse_groups(keys, pl, n_ids + 1, 3, 1000, 20, 0.6)
    ids = np.unique((keys >> np.uint64(32)).astype(np.int64))
    for rank, full, own in res:
        assert full["group"] == ids.tolist() and full["n_families"] == og["total_families"]
        assert full["family_count"] == og["family_count"][ids].tolist()
        assert full["consensus_count"] == og["consensus_count"][ids].tolist()
        mine = ids[ids % world == rank]
        assert own["group"] == mine.tolist() and own["family_count"] == og["family_count"][mine].tolist()
    assert sum(o["n_families"] for _, _, o in res) == og["total_families"]


def test_exchanged_collapse_refuses_three_ranks_and_cpu_reads_without_ops():
    import torch

    from precise_mrd.distributed import collapse_reads_exchanged

    with pytest.raises(RuntimeError, match="CUDA tensors"):           # no CPU fallback: the product path is the GPU's
        collapse_reads_exchanged(None, torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int32), None, 8)


def test_exchanged_collapse_single_process_is_the_plain_collapse():
    """No process group: the exchange degenerates to the local collapse (one owner, nothing sent)."""
    import torch

    sys.path.insert(0, str(ROOT))
    from oracle import cport
    from precise_mrd.distributed import collapse_reads_exchanged

    keys, pl = _striped_reads(0, 20_000, 300)
    t = {}
    tab = collapse_reads_exchanged(None, torch.from_numpy(keys.view(np.int64).copy()), torch.from_numpy(pl.view(np.int32).copy()),
                                   (3, 1000, 20, 0.6), 300, ops=_NumpyExchangeOps(), timings=t)
    og = cport.collapse_groups(keys, pl, 301, 3, 1000, 20, 0.6)
    ids = np.unique((keys >> np.uint64(32)).astype(np.int64))
    assert (tab["group"] == ids).all() and tab["n_families"] == og["total_families"]
    assert (tab["family_count"] == og["family_count"][ids]).all() and (tab["consensus_count"] == og["consensus_count"][ids]).all()
    assert t == {}                                   # (device timings only exist for CUDA tensors)
