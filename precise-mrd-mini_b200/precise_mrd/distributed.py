"""Multi-GPU plumbing: one process per GPU, cells sharded across ranks, one small all-gather.

The path has no data-path exchange: (AF, depth, replicate) cells are independent and every
Philox stream is keyed by the GLOBAL cell id, so a cell gives the same counts whichever rank runs
it.  A rank simulates / collapses / counts its own cells; the only collective is
an all-gather of the per-cell ``(n_total, n_variants, n_families, n_reads)`` rows (32 B per
cell: 60 KB for the eval-lod grid), after which every rank holds the full table and runs the
error model, the tail tests and BH redundantly -- exactly what ``call_mrd`` does on one process."""
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
