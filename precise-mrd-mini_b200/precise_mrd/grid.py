"""Read-level grid engine: simulate -> collapse -> call for a grid of (AF, depth, replicate)
cells with the reads never leaving the GPU.  This is what ``eval-lob / eval-lod / eval-loq /
eval-contamination`` and ``bench.py`` drive; it is the device-resident form of the L2 call
chain ``simulate_reads -> collapse_umis -> fit_error_model -> call_mrd``
(``src/precise_mrd/cli.py:103-119``, ``eval/lod.py:63-68``) with the read-level arithmetic of
``io.py`` / ``umi.py`` that ``north_star`` describes."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _native as nv
from .config import section


def grid_cells(allele_fractions: Sequence[float], umi_depths: Sequence[int], n_replicates: int,
               rep_offset: int = 0, rep_stride: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """``np.meshgrid(AF, depth, rep, indexing='ij')`` flattened, AF-major (``simulate.py:133-141``).

    Returns ``(cell_id, af, depth, replicate)``.  ``cell_id = (af_i * n_depth + depth_i) *
    rep_stride + rep`` is the GLOBAL cell index that keys the Philox stream, so a cell's reads
    are the same whichever rank or batch it lands in; ``rep_offset`` shifts the replicate
    range (multi-GPU sharding) and ``rep_stride`` (default: ``rep_offset + n_replicates``) is
    the global replicate count."""
    af = np.asarray(allele_fractions, dtype=np.float64)
    depth = np.asarray(umi_depths, dtype=np.int64)
    rep = np.arange(rep_offset, rep_offset + n_replicates, dtype=np.int64)
    stride = int(rep_stride if rep_stride is not None else rep_offset + n_replicates)
    a, d, r = np.meshgrid(np.arange(len(af)), np.arange(len(depth)), rep, indexing="ij")
    a, d, r = a.ravel(), d.ravel(), r.ravel()
    cell_id = (a * len(depth) + d) * stride + r
    return cell_id.astype(np.int64), af[a], depth[d], r


def umi_params(config, mode: int = nv.MODE_WEIGHTED, cluster: int = nv.CLUSTER_EXACT) -> nv.UmiParams:
    u = section(config, "umi")
    return nv.UmiParams(min_family_size=int(u["min_family_size"]), max_family_size=int(u["max_family_size"]),
                        quality_threshold=int(u["quality_threshold"]), consensus_threshold=float(u["consensus_threshold"]),
                        mode=mode, cluster=cluster, max_distance=int(u.get("max_edit_distance", 0) or 0), umi_len=12)


def read_sim_params(config=None, mean_family_size: float = 5.0, family_model: int = 1, hop_rate: float = 0.0,
                    collision_rate: float = 0.0, contamination_rate: float = 0.0, shuffle: bool = True,
                    chunk_log2: int = 0, fused: bool = False) -> nv.ReadSimParams:
    max_fs = int(section(config, "umi")["max_family_size"]) if config is not None else 1000
    return nv.ReadSimParams(mean_family_size=mean_family_size, family_model=family_model, max_family_size=max_fs,
                            contamination_rate=contamination_rate, hop_rate=hop_rate, collision_rate=collision_rate,
                            shuffle=1 if shuffle else 0, chunk_log2=chunk_log2,
                            engine=nv.ENGINE_FUSED if fused else nv.ENGINE_MATERIALISED)


def run_cells(config, cell_id, allele_fraction, n_families, ctx: Optional[nv.Context] = None,
              sim: Optional[nv.ReadSimParams] = None, neg_af_max: float = 1e-4,
              alternative: int = nv.ALT_TWO_SIDED, knobs: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
    """One pass of the hot path over a batch of cells given as HOST arrays; returns the
    per-cell call table as HOST arrays (the C ABI stages both directions).  ``knobs``: per-cell
    contamination knobs and call groups (``pmrd_cell_knobs``)."""
    st = section(config, "stats")
    if st["test_type"] not in nv.TEST_TYPES:
        raise ValueError(f"Unknown test type: {st['test_type']}")      # call.py:187
    ctx = ctx or nv.default_context(int(config.seed))
    sim = sim or read_sim_params(config)
    res = ctx.pipeline_cells(cell_id, allele_fraction, n_families, sim, umi_params(config), neg_af_max=neg_af_max,
                             test_type=nv.TEST_TYPES[st["test_type"]], alternative=alternative, alpha=float(st["alpha"]), knobs=knobs)
    res["variant_call"] = res["significant"]
    return res
