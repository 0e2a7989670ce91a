"""simulate_reads -- drop-in for ``src/precise_mrd/simulate.py:14-189`` (G-D v2, vectorised
branch ``:133-180``).  One row per (AF, depth, replicate) cell; the five per-cell draws run in
kernel K1 (``csrc/simulate.cu``) on a Philox stream keyed by the cell index."""
from __future__ import annotations

from typing import Optional

import numpy as np
import pandas as pd

from ._device import stage_context
from .config import section


def simulate_reads(config, rng: np.random.Generator, output_path: Optional[str] = None, use_cache: bool = True,
                   chunked_processing: bool = False, background_multiplier: Optional[float] = None) -> pd.DataFrame:
    """Same signature as the reference.  ``use_cache`` / ``chunked_processing`` are accepted and
    ignored: the reference's 1-hour TTL cache silently changes later draws and
    the chunked branch only exists to bound host memory.  ``background_multiplier`` (not in the
    reference signature) is the trinucleotide-context factor of ``eval/stratified.py:180-203``,
    applied inside the kernel."""
    sim = section(config, "simulation")
    umi = section(config, "umi")
    afs, depths, n_rep = sim["allele_fractions"], sim["umi_depths"], int(sim["n_replicates"])
    af_g, depth_g, rep_g = np.meshgrid(np.asarray(afs, dtype=float), np.asarray(depths, dtype=np.int64),
                                       np.arange(n_rep), indexing="ij")
    af_flat, depth_flat, rep_flat = af_g.ravel(), depth_g.ravel(), rep_g.ravel()
    total = af_flat.shape[0]
    print(f"  Simulating {total:,} samples with {len(afs)} AF levels, {len(depths)} depths, {n_rep} replicates")
    ctx = stage_context(config, rng)
    sample_ids = np.arange(total, dtype=np.int64)
    cells = ctx.simulate_cells(sample_ids, af_flat, depth_flat, int(umi["max_family_size"]),
                               bg_multiplier=background_multiplier)
    df = pd.DataFrame({
        "sample_id": sample_ids,
        "allele_fraction": af_flat,
        "target_depth": depth_flat,
        "replicate": rep_flat,
        "n_families": depth_flat,
        "n_true_variants": cells["n_true_variants"],
        "n_false_positives": cells["n_false_positives"],
        "background_rate": cells["background_rate"],
        "mean_family_size": cells["mean_family_size"],
        "mean_quality": cells["mean_quality"],
        "config_hash": config.config_hash(),
    })
    if output_path:
        df.to_parquet(output_path, index=False)
    return df
