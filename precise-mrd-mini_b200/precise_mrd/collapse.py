"""collapse_umis -- drop-in for ``src/precise_mrd/collapse.py:12-125`` (G-D v2).

In the G-D generation "collapse" is a per-family Monte-Carlo generator, not a grouping
(SURVEY.md §0.3-1): for each of ``n_families`` it draws a family size, a quality, a consensus
agreement and a variant flag.  Kernel K2 (``csrc/simulate.cu``) makes those draws on a Philox
stream keyed by (sample, family), compacts away the families below ``min_family_size`` and
returns the table in the reference's row order.  The read-level grouping the name suggests is
``precise_mrd.umi.UMIProcessor`` / ``precise_mrd.grid``."""
from __future__ import annotations

from typing import Optional

import numpy as np
import pandas as pd

from . import _native as nv
from ._device import stage_context
from .config import section

COLUMNS = ["sample_id", "family_id", "family_size", "quality_score", "consensus_agreement", "passes_quality",
           "passes_consensus", "is_variant", "allele_fraction"]


def collapse_umis(reads_df: pd.DataFrame, config, rng: np.random.Generator, output_path: Optional[str] = None,
                  is_fastq_data: bool = False, variant: int = 2) -> pd.DataFrame:
    umi = section(config, "umi")
    if is_fastq_data:
        # collapse.py:46-64: FASTQ rows arrive already grouped per UMI family
        df = pd.DataFrame({
            "sample_id": reads_df["sample_id"].values,
            "family_id": 0,
            "family_size": reads_df["family_size"].values,
            "quality_score": reads_df["mean_quality"].values,
            "consensus_agreement": reads_df["consensus_agreement"].values,
            "passes_quality": reads_df["passes_quality"].values,
            "passes_consensus": reads_df["passes_consensus"].values,
            "is_variant": False,
            "allele_fraction": 0.0,
        }, columns=COLUMNS)
    elif len(reads_df) == 0:
        df = pd.DataFrame()
    else:
        params = nv.UmiParams(min_family_size=int(umi["min_family_size"]), max_family_size=int(umi["max_family_size"]),
                              quality_threshold=int(umi["quality_threshold"]),
                              consensus_threshold=float(umi["consensus_threshold"]), mode=nv.MODE_COUNT,
                              cluster=nv.CLUSTER_EXACT, max_distance=0, umi_len=12)
        ctx = stage_context(config, rng)
        sample_id = reads_df["sample_id"].values
        # the Philox stream is keyed by the ROW position, so rows that contamination.py renamed
        # ('hopped_3', 'contam_0') or duplicated still get independent draws
        cell_key = np.arange(len(reads_df), dtype=np.int64)
        af = reads_df["allele_fraction"].values.astype(np.float64)
        offs, fam = ctx.collapse_families(cell_key, af, reads_df["n_families"].values.astype(np.int64),
                                          reads_df["background_rate"].values.astype(np.float64),
                                          reads_df["mean_family_size"].values.astype(np.float64), params, variant=variant)
        cell = fam["cell"]
        fl = fam["flags"]
        df = pd.DataFrame({
            "sample_id": sample_id[cell],
            "family_id": fam["family_id"].astype(np.int64),
            "family_size": fam["family_size"].astype(np.int64),
            "quality_score": fam["quality_score"],
            "consensus_agreement": fam["consensus_agreement"],
            "passes_quality": (fl & 1).astype(bool),
            "passes_consensus": ((fl >> 1) & 1).astype(bool),
            "is_variant": ((fl >> 2) & 1).astype(bool),
            "allele_fraction": af[cell],
        }, columns=COLUMNS)
    if output_path:
        df.to_parquet(output_path, index=False)
    return df
