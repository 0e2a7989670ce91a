"""MRD calling -- drop-in for ``src/precise_mrd/call.py`` (statistical branch).

``call_mrd`` dispatches to a ``VariantCaller`` plugin (``models/``); the statistical one runs K7
(per-sample segmented counts and fp64 means), K8 (two-sided Poisson / binomial tail p-values)
and K9 (Benjamini-Hochberg: device radix sort + reverse running-min scan) through ``pmrd_call``;
the scalar helpers go through ``pmrd_pvalues`` / ``pmrd_bh``.  The result carries the
``variant_call`` column every ``eval-*`` consumer indexes but no reference producer emits
(SURVEY.md §0.3-2): ``variant_call == significant``."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import pandas as pd

from . import _native as nv
from .models import B200StatisticalVariantCaller, VariantCaller


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


def get_variant_caller(config, use_ml_calling: bool = False, ml_model_type: str = "ensemble",
                       use_deep_learning: bool = False, dl_model_type: str = "cnn_lstm") -> VariantCaller:
    """Factory of ``precise_mrd/call.py:21-34``.  The ML / deep-learning callers are outside the
    B200 hot path; a third-party ``VariantCaller`` subclass can be used directly."""
    if use_ml_calling or use_deep_learning:
        raise NotImplementedError("ML-based calling is outside the B200 hot path (SURVEY.md §2.5)")
    return B200StatisticalVariantCaller(config)


def call_mrd(collapsed_df: pd.DataFrame, error_model_df: pd.DataFrame, config, rng: np.random.Generator,
             output_path: Optional[str] = None, use_ml_calling: bool = False, ml_model_type: str = "ensemble",
             use_deep_learning: bool = False, dl_model_type: str = "cnn_lstm") -> pd.DataFrame:
    """Dispatch to the pluggable caller (train is a no-op for the statistical one), write parquet
    iff ``output_path``."""
    caller = get_variant_caller(config, use_ml_calling, ml_model_type, use_deep_learning, dl_model_type)
    caller.train(collapsed_df, rng)
    df = caller.predict(collapsed_df, error_model_df)
    if output_path:
        df.to_parquet(output_path, index=False)
    return df
