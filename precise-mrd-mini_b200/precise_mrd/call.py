"""MRD calling -- drop-in for ``src/precise_mrd/call.py`` (statistical branch).

``call_mrd`` runs K7 (per-sample segmented counts and fp64 means), K8 (two-sided Poisson /
binomial tail p-values) and K9 (Benjamini-Hochberg: device radix sort + reverse running-min
scan) through ``pmrd_call``; the scalar helpers go through ``pmrd_pvalues`` / ``pmrd_bh``.
Adds the ``variant_call`` column every ``eval-*`` consumer indexes but no reference producer
emits: ``variant_call == significant``."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import pandas as pd

from . import _native as nv
from .config import section


def poisson_test(observed: int, expected: float) -> float:
    """call.py:14-25 two-tailed Poisson test."""
    return float(nv.default_context().pvalues([int(observed)], None, [float(expected)], nv.TEST_POISSON)[0])


def binomial_test(successes: int, trials: int, p: float) -> float:
    """call.py:28-38 two-tailed binomial test (scipy binomtest 'two-sided')."""
    return float(nv.default_context().pvalues([int(successes)], [int(trials)], [float(p)], nv.TEST_BINOMIAL)[0])


def benjamini_hochberg_correction(p_values: np.ndarray, alpha: float) -> Tuple[np.ndarray, np.ndarray]:
    """call.py:41-79.  Empty in -> (bool[0], float[0])."""
    p_values = np.asarray(p_values, dtype=float)
    if len(p_values) == 0:
        return np.array([], dtype=bool), np.array([])
    return nv.default_context().bh(p_values, float(alpha))


def call_mrd(collapsed_df: pd.DataFrame, error_model_df: pd.DataFrame, config, rng: np.random.Generator,
             output_path: Optional[str] = None, use_ml_calling: bool = False) -> pd.DataFrame:
    if use_ml_calling:
        raise NotImplementedError("ML-based calling is outside the B200 hot path (SURVEY.md §2.5)")
    st = section(config, "stats")
    test_type, alpha, fdr_method = st["test_type"], float(st["alpha"]), st["fdr_method"]
    if test_type not in nv.TEST_TYPES:
        raise ValueError(f"Unknown test type: {test_type}")                     # call.py:187
    if len(collapsed_df) == 0:
        return pd.DataFrame()                                                   # call.py:214-215
    # call.py:161: samples in first-appearance order; rows of a sample need not be contiguous
    codes, uniques = pd.factorize(collapsed_df["sample_id"], sort=False)
    is_variant = collapsed_df["is_variant"].values.astype(np.uint8)
    quality = collapsed_df["quality_score"].values.astype(np.float64)
    agreement = collapsed_df["consensus_agreement"].values.astype(np.float64)
    af = collapsed_df["allele_fraction"].values
    if (np.diff(codes) < 0).any():
        order = np.argsort(codes, kind="stable")
        codes, is_variant, quality, agreement, af = codes[order], is_variant[order], quality[order], agreement[order], af[order]
    seg = np.concatenate([[0], np.flatnonzero(np.diff(codes)) + 1, [len(codes)]]).astype(np.int64)
    mean_error_rate = float(error_model_df["error_rate"].mean())                # call.py:178
    ctx = nv.default_context()
    out = ctx.call(seg, is_variant, quality, agreement, mean_error_rate, nv.TEST_TYPES[test_type], alpha)
    df = pd.DataFrame({
        "sample_id": np.asarray(uniques),
        "allele_fraction": af[seg[:-1]],
        "n_variants": out["n_variants"],
        "n_total": out["n_total"],
        "variant_fraction": out["variant_fraction"],
        "expected_variants": out["expected_variants"],
        "fold_change": out["fold_change"],
        "p_value": out["p_value"],
        "mean_quality": out["mean_quality"],
        "mean_consensus": out["mean_consensus"],
        "test_type": test_type,
        "config_hash": config.config_hash(),
        "p_adjusted": out["p_adjusted"],
        "significant": out["significant"],
        "alpha": alpha,
        "fdr_method": fdr_method,
    })
    df["variant_call"] = df["significant"]
    if output_path:
        df.to_parquet(output_path, index=False)
    return df
