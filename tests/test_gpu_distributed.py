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


# ---------------------------------------------------------------- the exchange of external reads
def _striped_reads(rank, n, n_ids):
    rng = np.random.default_rng(100 + rank)
    g = rng.integers(0, n_ids, n, dtype=np.uint64)
    g[g % np.uint64(11) == 3] += np.uint64(1)                              # some ids stay empty
    umi = rng.integers(0, 60, n, dtype=np.uint64) * np.uint64(2654435761) % np.uint64(1 << 24)
    q = rng.choice([5, 10, 20, 30, 30, 37, 40], n).astype(np.uint32)
    pl = (rng.integers(0, 4, n).astype(np.uint32)) | (q << np.uint32(2)) | (rng.integers(0, 1 << 20, n).astype(np.uint32) << np.uint32(8))
    return (g << np.uint64(32)) | umi, pl


def _exchange_worker(rank, world, port, q, n_ids, n_per_rank):
    sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    from precise_mrd import _native as nv
    from precise_mrd.distributed import collapse_reads_exchanged

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = nv.Context(device=rank, seed=7)
    P = nv.UmiParams(min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6, mode=nv.MODE_WEIGHTED,
                     cluster=nv.CLUSTER_EXACT, max_distance=0, umi_len=12)
    keys, pl = _striped_reads(rank, n_per_rank + 1001 * rank, n_ids)
    dev = torch.device("cuda", rank)
    out = {}
    from precise_mrd.distributed import PeerExchange

    peer = PeerExchange(ctx, 2 * (n_per_rank + 1001 * world))          # receive arrays of all ranks, mapped everywhere
    for gather in (True, False):
        tk, tp = torch.from_numpy(keys.view(np.int64)).to(dev), torch.from_numpy(pl.view(np.int32)).to(dev)
        t = {}
        tab = collapse_reads_exchanged(ctx, tk, tp, P, n_ids, gather=gather, timings=t)
        out[gather] = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in tab.items()}
        out[gather]["timings"] = t
    # the same through ONE cut-and-store kernel over peer memory instead of partitioned copy + NCCL all-to-all (twice: the
    # receive arrays are reused)
    for it in range(2):
        tk, tp = torch.from_numpy(keys.view(np.int64)).to(dev), torch.from_numpy(pl.view(np.int32)).to(dev)
        tab = collapse_reads_exchanged(ctx, tk, tp, P, n_ids, gather=True, peer=peer)
        out[("peer", it)] = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in tab.items()}
    with pytest.raises(RuntimeError, match="capacity"):
        small = PeerExchange(ctx, 1000)
        try:
            tk, tp = torch.from_numpy(keys.view(np.int64)).to(dev), torch.from_numpy(pl.view(np.int32)).to(dev)
            collapse_reads_exchanged(ctx, tk, tp, P, n_ids, peer=small)
        finally:
            small.close()
    peer.close()
    launched = ctx.launch_count
    q.put((rank, out[True], out[False], launched, out[("peer", 0)], out[("peer", 1)]))
    ctx.close() if hasattr(ctx, "close") else None
    dist.destroy_process_group()


def test_exchanged_collapse_on_gpus_equals_one_gpu(native, ctx):
    """External reads of one sample striped over 2 (and, when the box has them, 4) GPUs: owner cut on the device
    (pmrd_owner_partition_dev) -> all-to-all over NCCL -> pmrd_collapse_reads_dev at the owner.  Every rank's gathered
    table must equal the single-GPU table of the reads concatenated in rank order, which in turn equals the C oracle's."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    from oracle import cport

    n_ids, n_per_rank = 5000, 400_000
    for world in [w for w in (2, 4) if torch.cuda.device_count() >= w]:
        port = _free_port()
        mpc = mp.get_context("spawn")
        q = mpc.Queue()
        procs = [mpc.Process(target=_exchange_worker, args=(r, world, port, q, n_ids, n_per_rank)) for r in range(world)]
        for p in procs:
            p.start()
        res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        parts = [_striped_reads(r, n_per_rank + 1001 * r, n_ids) for r in range(world)]
        keys, pl = np.concatenate([a for a, _ in parts]), np.concatenate([b for _, b in parts])
        P = native.UmiParams(min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6, mode=native.MODE_WEIGHTED,
                             cluster=native.CLUSTER_EXACT, max_distance=0, umi_len=12)
        info, one = ctx.collapse_reads(keys, pl, P, want_families=False)
        og = cport.collapse_groups(keys, pl, n_ids + 1, 3, 1000, 20, 0.6)
        assert (one["family_count"] == og["family_count"][one["group"]]).all() and info["n_families"] == og["total_families"]
        for rank, full, own, launched, peer0, peer1 in res:
            for pv in (peer0, peer1):                                   # the peer-memory form gives the very same table
                assert pv["group"] == full["group"] and pv["family_count"] == full["family_count"], (world, rank)
                assert pv["consensus_count"] == full["consensus_count"] and pv["n_families"] == full["n_families"]
            assert full["group"] == one["group"].tolist() and full["n_families"] == info["n_families"], (world, rank)
            assert full["family_count"] == one["family_count"].tolist()
            assert full["consensus_count"] == one["consensus_count"].tolist()
            mine = one["group"] % world == rank
            assert own["group"] == one["group"][mine].tolist() and own["family_count"] == one["family_count"][mine].tolist()
            t = full["timings"]
            assert t["received_reads"] > 0 and t["sent_reads"] > 0 and launched > 0       # reads really changed hands, our kernels ran
        assert sum(r_[2]["n_families"] for r_ in res) == info["n_families"]
