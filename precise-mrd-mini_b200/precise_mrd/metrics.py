"""Performance metrics -- ``src/precise_mrd/metrics.py:18-316``.

ROC-AUC / average precision follow sklearn's definitions (trapezoidal AUC over distinct
thresholds; AP = sum (R_n - R_{n-1}) P_n) in vectorised numpy; the bootstrap draws all resample
indices up front from the caller's generator (``metrics.py:76-79``) so the result is a pure
function of the generator state.  [SURVEY.md §8(f)-2 lists the device version as the next row.]"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np
import pandas as pd


def _binary_curve(y_true: np.ndarray, y_score: np.ndarray):
    order = np.argsort(-y_score, kind="mergesort")
    y_true, y_score = y_true[order], y_score[order]
    distinct = np.flatnonzero(np.diff(y_score)) if len(y_score) > 1 else np.array([], dtype=int)
    idx = np.r_[distinct, len(y_true) - 1]
    tps = np.cumsum(y_true)[idx].astype(float)
    fps = (1 + idx - tps).astype(float)
    return tps, fps


def roc_auc_score(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """metrics.py:18-24: single-class input -> 0.5."""
    y_true, y_score = np.asarray(y_true).astype(int), np.asarray(y_score, dtype=float)
    if len(np.unique(y_true)) != 2:
        return 0.5
    tps, fps = _binary_curve(y_true, y_score)
    tpr, fpr = np.r_[0.0, tps] / tps[-1], np.r_[0.0, fps] / fps[-1]
    return float(np.trapezoid(tpr, fpr)) if hasattr(np, "trapezoid") else float(np.trapz(tpr, fpr))


def average_precision(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """metrics.py:27-33: degenerate input -> mean(y_true)."""
    y_true, y_score = np.asarray(y_true).astype(int), np.asarray(y_score, dtype=float)
    if y_true.sum() == 0 or len(y_true) == 0:
        return float(np.mean(y_true)) if len(y_true) else 0.0
    tps, fps = _binary_curve(y_true, y_score)
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def bootstrap_metric(y_true, y_score, metric_func, n_bootstrap: int = 1000, rng: Optional[np.random.Generator] = None,
                     n_jobs: int = -1) -> Dict[str, float]:
    if rng is None:
        rng = np.random.default_rng(42)
    y_true, y_score = np.asarray(y_true), np.asarray(y_score)
    n = len(y_true)
    idx = [rng.choice(n, size=n, replace=True) for _ in range(n_bootstrap)]       # metrics.py:76-79
    scores = []
    for ii in idx:
        try:
            scores.append(metric_func(y_true[ii], y_score[ii]))
        except Exception:
            continue
    if not scores:
        return {"mean": 0.0, "lower": 0.0, "upper": 0.0, "std": 0.0}
    s = np.array(scores)
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
