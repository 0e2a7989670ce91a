"""Multi-GPU plumbing: one process per GPU, cells sharded across ranks, one small all-gather.

The generated grids have no data-path exchange: (AF, depth, replicate) cells are independent and every
Philox stream is keyed by the GLOBAL cell id, so a cell gives the same counts whichever rank runs
it.  A rank simulates / collapses / counts its own cells; the only collective is
an all-gather of the per-cell ``(n_total, n_variants, n_families, n_reads)`` rows (32 B per
cell: 60 KB for the eval-lod grid), after which every rank holds the full table and runs the
error model, the tail tests and BH redundantly -- exactly what ``call_mrd`` does on one process.

EXTERNAL reads are the one case with a real exchange (third bullet): when the reads of a sample are
striped over the ranks in arrival order, the families of a locus are spread over all of them.
``collapse_reads_exchanged`` gives every locus an owner rank, cuts each rank's reads into one run per owner on the
device (``pmrd_owner_partition_dev``), moves the runs with one all-to-all over NCCL and lets each owner collapse what
it received with the single-GPU kernels."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np


def shard_cells(weights: Sequence[float], rank: int, world: int) -> np.ndarray:
    """Indices of the cells rank ``rank`` owns: longest-processing-time greedy on the per-cell
    weight (the family count), so that the 100k-family cells spread evenly.  Deterministic, and
    the shards of all ranks partition ``range(len(weights))``."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.argsort(-w, kind="stable")
    load = np.zeros(world)
    owner = np.empty(len(w), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))          # ties -> lowest rank: deterministic
        owner[i] = r
        load[r] += w[i]
    return np.flatnonzero(owner == rank)


def all_gather_rows(local_index: np.ndarray, local_rows: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """Scatter every rank's ``local_rows[k]`` to row ``local_index[k]`` of a ``(n_total, C)`` int64
    table that all ranks receive.  Uses the initialised ``torch.distributed`` group (NCCL on GPUs,
    gloo on CPU); with no group it is the identity."""
    import torch
    import torch.distributed as dist

    local_rows = np.ascontiguousarray(local_rows, dtype=np.int64)
    cols = local_rows.shape[1]
    out = np.zeros((n_total, cols), dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out[local_index] = local_rows
        return out
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(local_index)], dtype=torch.int64, device=dev), group=group)
    m = int(max(int(c.item()) for c in counts))
    pad = np.full((m, cols + 1), -1, dtype=np.int64)          # column 0: global row index (-1 = padding)
    pad[: len(local_index), 0] = local_index
    pad[: len(local_index), 1:] = local_rows
    bufs = [torch.empty((m, cols + 1), dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(pad).to(dev), group=group)
    for b in bufs:
        a = b.cpu().numpy()
        keep = a[:, 0] >= 0
        out[a[keep, 0]] = a[keep, 1:]
    return out


def world_size() -> int:
    """Ranks of the initialised default process group (1 when torch.distributed is not in use)."""
    import sys

    if "torch" not in sys.modules:      # nobody imported torch, so no process group exists: do not pay its import
        return 1
    try:
        import torch.distributed as dist
    except Exception:          # pragma: no cover
        return 1
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def run_cells_sharded(config, cell_id, allele_fraction, n_families, neg_af_max: float = 1e-4, alternative: int = 0,
                      sim=None, compute=None, ctx=None, finish: bool = True) -> Dict[str, np.ndarray]:
    """``precise_mrd.grid.run_cells`` over all ranks of the current process group: each rank runs
    its shard of cells through the fused device plan, the per-cell counts are all-gathered, and
    every rank finishes the call stage (error rate, p-values, BH) over the whole grid with the
    plan's own kernels (``pmrd_call_counts``) -- the table equals the single-GPU one bit for bit.

    ``compute(cell_id, af, n_families) -> {"n_total", "n_variants", "n_families", "n_reads"}`` is
    injectable so that the sharding / gathering logic can be tested on CPU ranks (gloo); with an
    injected ``compute`` (or ``finish=False``) only the gathered counts are returned."""
    import torch.distributed as dist

    from . import _native as nv
    from .config import section

    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    cell_id = np.asarray(cell_id, dtype=np.int64)
    af = np.asarray(allele_fraction, dtype=np.float64)
    nf = np.asarray(n_families, dtype=np.int64)
    st = section(config, "stats")
    if st["test_type"] not in nv.TEST_TYPES:
        raise ValueError(f"Unknown test type: {st['test_type']}")      # call.py:187
    if len(cell_id) == 0:                                              # empty grid: an empty table on every rank
        out = {k: np.zeros(0, np.int64) for k in ("n_total", "n_variants", "n_families", "n_reads")}
        out["negative_control"] = np.zeros(0, bool)
        if compute is None and finish:
            out.update({"p_value": np.zeros(0), "p_adjusted": np.zeros(0), "significant": np.zeros(0, bool),
                        "variant_call": np.zeros(0, bool), "mean_error_rate": 0.0})
        return out
    mine = shard_cells(nf, rank, world)
    on_device = compute is None
    if on_device:
        from . import grid

        ctx = ctx or nv.default_context(int(config.seed))

        def compute(c, a, f):          # counts only: the rank's own p-values are discarded
            return grid.run_cells(config, c, a, f, ctx=ctx, sim=sim, neg_af_max=neg_af_max, alternative=alternative)

    loc = compute(cell_id[mine], af[mine], nf[mine]) if len(mine) else {k: np.zeros(0, np.int64) for k in ("n_total", "n_variants", "n_families", "n_reads")}
    rows = np.stack([loc["n_total"], loc["n_variants"], loc["n_families"], loc["n_reads"]], axis=1) if len(mine) else np.zeros((0, 4), np.int64)
    table = all_gather_rows(mine, rows, len(cell_id))
    out = {"n_total": table[:, 0], "n_variants": table[:, 1], "n_families": table[:, 2], "n_reads": table[:, 3]}
    # call stage over the full grid, identically on every rank (call.py:178-229)
    neg = af <= neg_af_max
    if not neg.any():
        neg = af == af.min()
    out["negative_control"] = neg
    if on_device and finish:
        out.update(ctx.call_counts(out["n_total"], out["n_variants"], neg, int(cell_id[0]), nv.TEST_TYPES[st["test_type"]],
                                   alternative, float(st["alpha"])))
        out["variant_call"] = out["significant"]
    return out


class _DeviceExchangeOps:
    """The exchange's device steps on torch CUDA tensors, through the C-ABI (``csrc/exchange.cu``, ``collapse.cu``)."""

    def __init__(self, ctx):
        self.ctx = ctx

    @staticmethod
    def _after_torch(t):
        # torch (copies, NCCL) wrote these arrays on ITS stream; the library launches on the context's
        import torch

        torch.cuda.current_stream(t.device).synchronize()

    def owner_partition(self, keys, payload, keys_tmp, payload_tmp, log2_world):
        self._after_torch(keys)
        return self.ctx.owner_partition_dev(keys.data_ptr(), payload.data_ptr(), keys_tmp.data_ptr(), payload_tmp.data_ptr(),
                                            keys.numel(), log2_world)

    def owner_rekey(self, keys, log2_world):
        self._after_torch(keys)
        self.ctx.owner_rekey_dev(keys.data_ptr(), keys.numel(), log2_world)

    def collapse_groups(self, keys, payload, keys_tmp, payload_tmp, params, capacity_g):
        """Per-group table of the reads in ``keys`` / ``payload`` (clobbered): ``(n_families, {column: ndarray})``."""
        import torch

        dev = keys.device
        grp = {"group": torch.empty(capacity_g, dtype=torch.int32, device=dev), "family_count": torch.empty(capacity_g, dtype=torch.int32, device=dev),
               "consensus_count": torch.empty(capacity_g * 4, dtype=torch.int32, device=dev),
               "alt_count": torch.empty(capacity_g, dtype=torch.int64, device=dev), "total_count": torch.empty(capacity_g, dtype=torch.int64, device=dev)}
        nf, ng = self.ctx.collapse_reads_dev(keys.data_ptr(), payload.data_ptr(), keys_tmp.data_ptr(), payload_tmp.data_ptr(), keys.numel(),
                                             params, None, 0, {k: v.data_ptr() for k, v in grp.items()}, capacity_g)
        torch.cuda.synchronize(dev)
        out = {"group": grp["group"][:ng].cpu().numpy().astype(np.int64), "family_count": grp["family_count"][:ng].cpu().numpy().astype(np.uint32),
               "consensus_count": grp["consensus_count"][:ng * 4].cpu().numpy().astype(np.uint32).reshape(ng, 4)}
        return nf, out


class _RawArray:
    """A device array known by pointer only (a peer-mapped receive array): what the device steps need of a tensor."""

    def __init__(self, ptr: int, n: int, device):
        self._ptr, self._n, self.device = int(ptr), int(n), device

    def data_ptr(self) -> int:
        return self._ptr

    def numel(self) -> int:
        return self._n


class PeerExchange:
    """Receive arrays of every rank, mapped into every rank: the set-up of the peer-memory exchange (one NVSwitch node,
    at most 8 ranks, one process per GPU).  Each rank allocates ``capacity_reads`` x (8 + 4) bytes with
    ``pmrd_peer_alloc``, the 64-byte IPC handles go round once (``all_gather_object``) and every rank maps the others'
    arrays (``pmrd_peer_open``).  ``collapse_reads_exchanged(..., peer=this)`` then cuts AND moves the reads with one
    kernel (``pmrd_owner_scatter_peer_dev``) instead of a partitioned copy + NCCL all-to-all."""

    def __init__(self, ctx, capacity_reads: int, group=None):
        import torch.distributed as dist

        self.ctx, self.group, self.capacity = ctx, group, int(capacity_reads)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8 or self.world & (self.world - 1):
            raise ValueError(f"PeerExchange serves 2, 4 or 8 ranks of one node, not {self.world}")
        self.kptr, kh = ctx.peer_alloc(self.capacity * 8)
        self.pptr, ph = ctx.peer_alloc(self.capacity * 4)
        handles = [None] * self.world
        dist.all_gather_object(handles, (kh, ph), group=group)
        self.peer_keys = [self.kptr if r == self.rank else ctx.peer_open(h[0]) for r, h in enumerate(handles)]
        self.peer_payload = [self.pptr if r == self.rank else ctx.peer_open(h[1]) for r, h in enumerate(handles)]
        dist.barrier(group)

    def close(self) -> None:
        import torch.distributed as dist

        if self.ctx is None:
            return
        dist.barrier(self.group)                      # nobody still writes into (or reads from) a mapped array
        for r in range(self.world):
            if r != self.rank:
                self.ctx.peer_close(self.peer_keys[r])
                self.ctx.peer_close(self.peer_payload[r])
        dist.barrier(self.group)                      # every mapping is gone before the memory is
        self.ctx.peer_free(self.kptr)
        self.ctx.peer_free(self.pptr)
        self.ctx = None


def exchange_splits(send_counts: Sequence[int], group=None, device=None) -> Tuple[list, list]:
    """``(send, recv)`` split sizes of the all-to-all: ``recv[r]`` = what rank ``r`` cut off for this rank."""
    import torch
    import torch.distributed as dist

    send = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return [int(v) for v in send.tolist()], [int(v) for v in recv.tolist()]


def collapse_reads_exchanged(ctx, keys, payload, params, n_ids: int, group=None, ops=None, gather: bool = True,
                             timings: Optional[dict] = None, peer: Optional[PeerExchange] = None) -> Dict[str, np.ndarray]:
    """``Context.collapse_reads(..., want_families=False)`` for reads held by ALL ranks of the process group.

    ``keys`` (int64: ``group << 32 | umi``) and ``payload`` (int32) are this rank's reads as torch tensors on its GPU,
    in arrival order; group ids are dense, ``< n_ids``.  Owner of a group = ``group mod world`` (world a power of two).
    Returns the per-group table (``group``, ``family_count``, ``consensus_count``) and ``n_families`` -- of ALL groups on
    every rank when ``gather`` (one all-gather of 24 B per group), else of the rank's own groups -- equal, bit for bit,
    to the single-GPU table of the ranks' reads concatenated in rank order.  Replaces umi.py:205-285 for that case.

    ``peer``: a ``PeerExchange`` of the group -- the cut then stores every read straight into its owner's receive array
    over NVLink (one kernel instead of partitioned copy + NCCL all-to-all); same table.

    ``ops`` injects the device steps (tests run the exchange logic on CPU ranks over gloo with a NumPy double); without
    it the reads must be CUDA tensors and the steps are the library's kernels."""
    import torch
    import torch.distributed as dist

    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    if world & (world - 1):
        raise ValueError(f"collapse_reads_exchanged needs a power-of-two number of ranks, not {world}")
    log2w = world.bit_length() - 1
    if ops is None:
        if not (keys.is_cuda and payload.is_cuda):
            raise RuntimeError("collapse_reads_exchanged: reads must be CUDA tensors (the exchange and the collapse run on the GPU)")
        ops = _DeviceExchangeOps(ctx)
    n = int(keys.numel())
    fused = peer is not None and world > 1
    ev = None
    if timings is not None and keys.is_cuda:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
    if fused:
        # ---- cut + exchange as one kernel over peer memory
        n_tiles = (n + 4095) // 4096
        tile_off = torch.empty(max(1, world * n_tiles), dtype=torch.int32, device=keys.device)
        ops._after_torch(keys)
        counts = ctx.owner_tiles_dev(keys.data_ptr(), n, log2w, tile_off.data_ptr()) if n else [0] * world
        mat = torch.empty(world * world, dtype=torch.int64, device=keys.device)
        dist.all_gather_into_tensor(mat, torch.tensor(counts, dtype=torch.int64, device=keys.device), group=group)
        mat = mat.view(world, world).cpu().numpy()                 # mat[r][o]: reads rank r sends to owner o
        base = mat[:rank].sum(axis=0).tolist()
        n_recv = int(mat[:, rank].sum())
        if int(mat.sum(axis=0).max()) > peer.capacity:             # (every rank sees the whole matrix: all of them raise, none waits)
            o = int(mat.sum(axis=0).argmax())
            raise RuntimeError(f"PeerExchange capacity {peer.capacity} reads < {int(mat[:, o].sum())} reads bound for rank {o}")
        if ev:
            ev[1].record()
        dist.barrier(group)                                        # every owner has consumed what its arrays held
        if n:
            ctx.owner_scatter_peer_dev(keys.data_ptr(), payload.data_ptr(), n, log2w, tile_off.data_ptr(), peer.peer_keys, peer.peer_payload, base)
        torch.cuda.synchronize(keys.device)
        dist.barrier(group)                                        # every writer has finished
        rk, rp = _RawArray(peer.kptr, n_recv, keys.device), _RawArray(peer.pptr, n_recv, keys.device)
    else:
        keys_tmp, payload_tmp = torch.empty_like(keys), torch.empty_like(payload)
        counts, in_tmp = ops.owner_partition(keys, payload, keys_tmp, payload_tmp, log2w) if n else ([0] * world, False)
        sk, sp = (keys_tmp, payload_tmp) if in_tmp else (keys, payload)
        if ev:
            ev[1].record()
    if fused:
        pass                                                        # (the reads are already at their owners)
    elif world > 1:
        send, recv = exchange_splits(counts, group, keys.device)
        n_recv = sum(recv)
        rk, rp = torch.empty(n_recv, dtype=keys.dtype, device=keys.device), torch.empty(n_recv, dtype=payload.dtype, device=keys.device)
        dist.all_to_all_single(rk, sk[:n], output_split_sizes=recv, input_split_sizes=send, group=group)
        dist.all_to_all_single(rp, sp[:n], output_split_sizes=recv, input_split_sizes=send, group=group)
    else:
        rk, rp, n_recv = sk, sp, n
    if ev:
        ev[2].record()
    n_local_ids = (int(n_ids) - rank + world - 1) // world if n_ids > rank else 0
    if n_recv:
        if not fused:                                               # (the peer-memory cut stores the dense ids itself)
            ops.owner_rekey(rk, log2w)
        rkt = torch.empty(n_recv, dtype=torch.int64, device=keys.device)
        rpt = torch.empty(n_recv, dtype=torch.int32, device=keys.device)
        nf, tab = ops.collapse_groups(rk, rp, rkt, rpt, params, n_local_ids + 16)
    else:
        nf, tab = 0, {"group": np.zeros(0, np.int64), "family_count": np.zeros(0, np.uint32), "consensus_count": np.zeros((0, 4), np.uint32)}
    if ev:
        ev[3].record()
        torch.cuda.synchronize(keys.device)
        timings.update({"partition_ms": ev[0].elapsed_time(ev[1]), "exchange_ms": ev[1].elapsed_time(ev[2]), "collapse_ms": ev[2].elapsed_time(ev[3]),
                        "sent_reads": n - counts[rank] if n else 0, "received_reads": n_recv})
    tab["group"] = tab["group"] * world + rank                                  # local id -> global id
    if not gather or world == 1:
        tab["n_families"] = int(nf)
        return tab
    rows = np.concatenate([tab["family_count"].astype(np.int64)[:, None], tab["consensus_count"].astype(np.int64),
                           np.ones((len(tab["group"]), 1), np.int64)], axis=1)   # last column: the group is present
    full = all_gather_rows(tab["group"], rows, int(n_ids), group)
    present = full[:, 5] != 0
    tot = torch.tensor([int(nf)], dtype=torch.int64, device=keys.device)
    dist.all_reduce(tot, group=group)
    return {"group": np.flatnonzero(present).astype(np.int64), "family_count": full[present, 0].astype(np.uint32),
            "consensus_count": full[present, 1:5].astype(np.uint32), "n_families": int(tot.item())}
