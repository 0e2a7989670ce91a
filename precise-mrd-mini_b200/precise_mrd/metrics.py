"""Performance metrics -- ``src/precise_mrd/metrics.py:18-316``.

ROC-AUC / average precision follow sklearn's definitions (trapezoidal AUC over distinct
thresholds; AP = sum (R_n - R_{n-1}) P_n) and run on the device (kernel K12,
``pmrd_bootstrap_metrics``): one sort of the scores, then every bootstrap resample is a histogram
over the threshold groups and one block-wide walk.  The bootstrap draws all resample indices up
front from the caller's generator (``metrics.py:76-79``), so the result is a pure function of the
generator state and resample-for-resample the reference's [SURVEY.md §8(f)-2]."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import pandas as pd

from . import _native as nv


def _as_binary(y_true) -> np.ndarray:
    y = np.asarray(y_true)
    return (y != 0).astype(np.uint8)


def roc_auc_score(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """metrics.py:18-24 (sklearn ``roc_auc_score``; single-class input -> 0.5), kernel K12."""
    if len(y_true) == 0:
        return 0.5
    pt, _, _ = nv.default_context().bootstrap_metrics(_as_binary(y_true), np.asarray(y_score, dtype=float))
    return float(pt[0])


def average_precision(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """metrics.py:27-33 (sklearn ``average_precision_score``; no positives -> mean(y_true) = 0), kernel K12."""
    if len(y_true) == 0:
        return 0.0
    pt, _, _ = nv.default_context().bootstrap_metrics(_as_binary(y_true), np.asarray(y_score, dtype=float))
    return float(pt[1])


def bootstrap_metric(y_true, y_score, metric_func, n_bootstrap: int = 1000, rng: Optional[np.random.Generator] = None,
                     n_jobs: int = -1) -> Dict[str, float]:
    """metrics.py:48-118.  The resample rows are drawn up-front from the caller's generator exactly
    as the reference does (``rng.choice(n, n, replace=True)`` per resample, ``:76-79``), so the
    generator ends in the same state and every resample is the reference's; all ``n_bootstrap``
    AUC / AP evaluations are then ONE device call (kernel K12).  A ``metric_func`` other than this
    module's two is a Python callable and is evaluated per resample on the host, as in the
    reference; ``n_jobs`` is accepted and ignored."""
    if rng is None:
        rng = np.random.default_rng(42)
    y_true, y_score = np.asarray(y_true), np.asarray(y_score)
    n = len(y_true)
    idx = [rng.choice(n, size=n, replace=True) for _ in range(n_bootstrap)]
    if metric_func in (roc_auc_score, average_precision) and n > 0 and n_bootstrap > 0:
        _, auc, ap = nv.default_context().bootstrap_metrics(_as_binary(y_true), y_score.astype(float), indices=np.stack(idx),
                                                            point=False)
        s = auc if metric_func is roc_auc_score else ap
    else:
        scores = []
        for ii in idx:
            try:
                scores.append(metric_func(y_true[ii], y_score[ii]))
            except Exception:
                continue
        s = np.array(scores)
    if len(s) == 0:
        return {"mean": 0.0, "lower": 0.0, "upper": 0.0, "std": 0.0}
    return {"mean": float(np.mean(s)), "lower": float(np.percentile(s, 2.5)), "upper": float(np.percentile(s, 97.5)),
            "std": float(np.std(s))}


def calibration_analysis(y_true: np.ndarray, y_prob: np.ndarray, n_bins: int = 10):
    y_true, y_prob = np.asarray(y_true), np.asarray(y_prob)
    bins = np.linspace(0, 1, n_bins + 1)
    out = []
    for i in range(n_bins):
        lo, hi = bins[i], bins[i + 1]
        in_bin = (y_prob >= lo) & ((y_prob <= hi) if i == n_bins - 1 else (y_prob < hi))   # metrics.py:144-147
        if in_bin.sum() == 0:
            continue
        out.append({"bin": i, "lower": float(lo), "upper": float(hi), "count": int(in_bin.sum()),
                    "event_rate": float(np.mean(y_true[in_bin])), "confidence": float(np.mean(y_prob[in_bin]))})
    return out


def calculate_metrics(calls_df: pd.DataFrame, rng: np.random.Generator, n_bootstrap: int = 1000, n_jobs: int = -1,
                      config=None, use_advanced_ci: bool = False, run_validation: bool = False) -> Dict[str, Any]:
    if len(calls_df) == 0:
        return {"roc_auc": 0.0, "average_precision": 0.0, "detected_cases": 0, "total_cases": 0, "calibration": []}
    y_true = (calls_df["allele_fraction"] > 0.001).astype(int).values if "allele_fraction" in calls_df else np.zeros(len(calls_df), int)
    if "variant_fraction" in calls_df:
        y_score = calls_df["variant_fraction"].values.astype(float)
    elif "p_value" in calls_df:
        y_score = 1.0 - calls_df["p_value"].values
    else:
        y_score = rng.random(len(calls_df))
    roc_ci = bootstrap_metric(y_true, y_score, roc_auc_score, n_bootstrap, rng, n_jobs)
    ap_ci = bootstrap_metric(y_true, y_score, average_precision, n_bootstrap, rng, n_jobs)
    if "significant" in calls_df:
        detected = int(np.sum(calls_df["significant"].values & (y_true == 1)))
        total = int(np.sum(y_true))
    else:
        detected, total = 0, len(calls_df)
    return {
        "roc_auc": float(roc_auc_score(y_true, y_score)), "roc_auc_ci": roc_ci,
        "average_precision": float(average_precision(y_true, y_score)), "average_precision_ci": ap_ci,
        "brier_score": float(np.mean((y_score - y_true) ** 2)),
        "detected_cases": detected, "total_cases": total,
        "calibration": calibration_analysis(y_true, y_score),
        "statistical_validation": None,
    }
