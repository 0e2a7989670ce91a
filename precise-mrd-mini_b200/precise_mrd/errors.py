"""Error / contamination model -- the part of ``src/precise_mrd/errors.py`` on the read-level
path: background rates from negative controls (``:29-67``) and the read-flip
contamination of a read table (``:94-137``).  The same flip, applied per family while reads are
being generated, is the ``contamination_rate`` knob of the device generator
(``pmrd_read_sim_params``, ``precise_mrd.io``); this module is the table-level form for reads
that already exist."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import pandas as pd

from .context import ContextAnalyzer

_BASES = np.array(["A", "C", "G", "T"])


class ErrorModel:
    """Model sequencing errors and contamination for MRD analysis."""

    def __init__(self, base_error_rate: float = 1e-4, contamination_rate: float = 1e-3):
        self.base_error_rate = base_error_rate
        self.contamination_rate = contamination_rate
        self.context_analyzer = ContextAnalyzer()
        self.context_error_rates: Dict[str, Dict[str, float]] = {}
        self.contamination_matrix: Dict = {}

    def estimate_background_errors(self, negative_control_data: pd.DataFrame, min_depth: int = 100) -> Dict[str, Dict[str, float]]:
        """``errors.py:29-67``: per normalised (context, mutation), the mean of 1 / consensus_count
        over the negative-control rows with depth >= min_depth that carry a non-reference allele."""
        deep = negative_control_data[negative_control_data["consensus_count"] >= min_depth]
        acc: Dict[str, Dict[str, list]] = {}
        ctx = deep["context"].astype(str) if "context" in deep else pd.Series("NNN", index=deep.index)
        for c, r, a, n in zip(ctx, deep["ref"].astype(str), deep["allele"].astype(str), deep["consensus_count"]):
            if a != r:
                nc, nr, na = self.context_analyzer.normalize_context(c, r, a)
                acc.setdefault(nc, {}).setdefault(f"{nr}>{na}", []).append(1.0 / n)
        self.context_error_rates = {c: {k: float(np.mean(v)) for k, v in m.items()} for c, m in acc.items()}
        return self.context_error_rates

    def simulate_contamination_events(self, reads_data: pd.DataFrame, contamination_rate: Optional[float] = None,
                                      rng: Optional[np.random.Generator] = None) -> pd.DataFrame:
        """``errors.py:94-137``: ``int(n * rate)`` reads chosen without replacement have their allele
        flipped -- reference reads to a random other base, non-reference reads back to the
        reference -- and every row gets ``is_contaminated``.  ``rng`` replaces the reference's use
        of the global ``np.random`` state (default: a generator seeded 0)."""
        rate = self.contamination_rate if contamination_rate is None else contamination_rate
        out = reads_data.copy()
        n = len(reads_data)
        n_flip = int(n * rate)
        if n_flip == 0:
            return out
        rng = np.random.default_rng(0) if rng is None else rng
        pick = rng.choice(n, size=n_flip, replace=False)
        allele = out["allele"].astype(str).to_numpy().copy()
        ref = out["ref"].astype(str).to_numpy()
        is_ref = allele[pick] == ref[pick]
        # a random base among the three that differ from the reference
        ref_code = np.searchsorted(_BASES, ref[pick])
        other = _BASES[(ref_code + 1 + rng.integers(0, 3, size=n_flip)) & 3]
        allele[pick] = np.where(is_ref, other, ref[pick])
        flag = np.zeros(n, dtype=bool)
        flag[pick] = True
        out["allele"] = allele
        out["is_contaminated"] = flag
        return out
