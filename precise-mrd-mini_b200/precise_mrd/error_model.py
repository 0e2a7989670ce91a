"""fit_error_model -- drop-in for ``src/precise_mrd/error_model.py:53-198``.

64 trinucleotide contexts; ``error_rate_c = mean(is_variant | negative control) * U(0.5, 2)``;
CI = percentiles 2.5 / 97.5 of 100 bootstrap resamples.  Resampling ``n`` rows with replacement
and taking ``mean(is_variant)`` is exactly ``Binomial(n, k/n) / n``, so kernel K10 draws the 6 400
resampled rates directly instead of materialising 6 400 pandas frames (16 s of the reference's
20 s smoke run).  No on-disk cache: the reference's cache key ignores the seed and the data and
so silently changes results on a rerun."""
from __future__ import annotations

from typing import Optional

import numpy as np
import pandas as pd

from ._device import stage_context

CONTEXTS = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT"]   # AAA, AAC, ... TTT (error_model.py:120-137)


def negative_controls(collapsed_df: pd.DataFrame) -> np.ndarray:
    """error_model.py:144-149: AF <= 1e-4, else the rows at the minimum AF."""
    af = collapsed_df["allele_fraction"].values
    mask = af <= 0.0001
    if not mask.any() and len(af):
        mask = af == af.min()
    return mask


def fit_error_model(collapsed_df: pd.DataFrame, config, rng: np.random.Generator, output_path: Optional[str] = None,
                    use_cache: bool = True, cache_dir: str = ".cache", use_advanced_stats: bool = False) -> pd.DataFrame:
    if use_advanced_stats:
        raise NotImplementedError("the Bayesian error model is outside the B200 hot path (SURVEY.md §2.5)")
    if len(collapsed_df):
        mask = negative_controls(collapsed_df)
        n_neg = int(mask.sum())
        n_var = int(collapsed_df["is_variant"].values[mask].sum())
    else:
        n_neg = n_var = 0
    ctx = stage_context(config, rng)
    rate, lo, hi = ctx.error_model(n_neg, n_var, 100, stream_id=0)
    df = pd.DataFrame({
        "trinucleotide_context": CONTEXTS,
        "error_rate": rate,
        "ci_lower": lo,
        "ci_upper": hi,
        "n_observations": n_neg,
        "config_hash": config.config_hash(),
    })
    if output_path:
        df.to_parquet(output_path, index=False)
    return df
