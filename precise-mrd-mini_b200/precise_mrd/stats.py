"""G-B statistical testing -- drop-in for ``src/precise_mrd/stats.py:30-288``
(``TestResult``, ``StatisticalTester``): per-locus ONE-SIDED tests + FDR correction.

Tail probabilities are kernel K8 (``pmrd_pvalues`` with ``PMRD_ALT_GREATER`` / ``PMRD_ALT_LESS``
/ ``PMRD_ALT_TWO_SIDED``), Benjamini-Hochberg is K9 (``pmrd_bh``); ``test_multiple_variants``
sends the whole locus table through ONE launch of each instead of a Python loop of scipy calls.
The exact (Clopper-Pearson) proportion interval scipy attaches to ``binomtest`` is found by
bisection over the same device tails.  ``multipletests`` (statsmodels, absent here) is restated
for the three methods the reference maps to; 'bonferroni' and 'holm' are O(m) host arithmetic
on the device p-values."""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd

from . import _native as nv

_ALT = {"greater": nv.ALT_GREATER, "less": nv.ALT_LESS, "two-sided": nv.ALT_TWO_SIDED}


@dataclass
class TestResult:
    """Container for statistical test results (``stats.py:19-27``)."""
    __test__ = False   # not a pytest class
    statistic: float
    pvalue: float
    alternative: str
    method: str
    effect_size: Optional[float] = None
    confidence_interval: Optional[Tuple[float, float]] = None


def _tails(k, n, rate, test: int, alternative: str) -> np.ndarray:
    if alternative not in _ALT:
        raise ValueError("alternative must be 'greater', 'less', or 'two-sided'")
    return nv.default_context().pvalues(k, n, rate, test, _ALT[alternative])


def _proportion_ci(k: int, n: int, alternative: str, confidence_level: float = 0.95) -> Tuple[float, float]:
    """scipy ``BinomTestResult.proportion_ci(method='exact')``: Clopper-Pearson bounds, i.e. the
    p at which the one-sided tail of the observed count equals alpha (alpha/2 when two-sided).
    The tails are monotone in p, so 64 bisection steps on the device tail kernel converge to
    the last bit."""
    alpha = 1.0 - confidence_level
    if alternative == "two-sided":
        alpha /= 2.0

    def solve(side: str) -> float:
        # lower bound: P(X >= k | p) = alpha (increasing in p); upper: P(X <= k | p) = alpha (decreasing)
        lo, hi = 0.0, 1.0
        for _ in range(64):
            mid = 0.5 * (lo + hi)
            t = float(_tails([k], [n], [mid], nv.TEST_BINOMIAL, "greater" if side == "lower" else "less")[0])
            if (t < alpha) == (side == "lower"):
                lo = mid
            else:
                hi = mid
        return 0.5 * (lo + hi)

    low = 0.0 if (alternative == "less" or k == 0) else solve("lower")
    high = 1.0 if (alternative == "greater" or k == n) else solve("upper")
    return (low, high)


class StatisticalTester:
    """Perform statistical hypothesis testing for variant calling."""

    def __init__(self, test_type: str = "poisson", alpha: float = 0.05, fdr_method: str = "benjamini_hochberg"):
        self.test_type = test_type
        self.alpha = alpha
        self.fdr_method = fdr_method
        if test_type not in ["poisson", "binomial"]:
            raise ValueError("test_type must be 'poisson' or 'binomial'")

    # ------------------------------------------------------------------ scalar tests
    def poisson_test(self, observed: int, expected: float, alternative: str = "greater") -> TestResult:
        """``stats.py:47-90``.  The reference's two-sided form (2 cdf(k) if k <= mu else
        2 (1 - cdf(k-1)), capped at 1) equals 2 min(cdf(k), sf(k-1)) capped at 1: whenever the
        two differ both exceed 1 (a Poisson median is >= floor(mu))."""
        if expected <= 0:
            return TestResult(statistic=observed, pvalue=1.0, alternative=alternative, method="poisson")
        pvalue = float(_tails([int(observed)], None, [float(expected)], nv.TEST_POISSON, alternative)[0])
        effect = float(np.log(observed + 1e-8) - np.log(expected + 1e-8)) if expected > 1e-12 else float("nan")
        return TestResult(statistic=observed, pvalue=pvalue, alternative=alternative, method="poisson", effect_size=effect)

    def binomial_test(self, successes: int, trials: int, expected_prob: float, alternative: str = "greater") -> TestResult:
        """``stats.py:92-133`` (scipy ``binomtest``; statistic = k / n; exact proportion CI)."""
        if trials <= 0:
            return TestResult(statistic=successes, pvalue=1.0, alternative=alternative, method="binomial")
        if successes < 0 or successes > trials:
            raise ValueError("k must be an integer such that 0 <= k <= n")       # scipy's message
        if not 0.0 <= expected_prob <= 1.0:
            raise ValueError("p must be in range [0,1]")
        pvalue = float(_tails([int(successes)], [int(trials)], [float(expected_prob)], nv.TEST_BINOMIAL, alternative)[0])
        observed_prob = successes / trials
        if 1e-12 < expected_prob < (1 - 1e-12):
            odds = float((np.log(observed_prob + 1e-12) - np.log(1 - observed_prob + 1e-12))
                         - (np.log(expected_prob + 1e-12) - np.log(1 - expected_prob + 1e-12)))
        else:
            odds = float(np.clip(observed_prob / (expected_prob + 1e-12), 1e-12, 1e12))
        return TestResult(statistic=observed_prob, pvalue=pvalue, alternative=alternative, method="binomial",
                          effect_size=odds, confidence_interval=_proportion_ci(int(successes), int(trials), alternative))

    def test_variant_significance(self, observed_alt: int, total_depth: int, expected_error_rate: float,
                                  context: Optional[str] = None) -> TestResult:
        """``stats.py:135-157``: one-sided (greater) test of the alt count against the context's rate."""
        if self.test_type == "poisson":
            return self.poisson_test(observed_alt, expected_error_rate * total_depth, alternative="greater")
        elif self.test_type == "binomial":
            return self.binomial_test(observed_alt, total_depth, expected_error_rate, alternative="greater")
        raise ValueError(f"Unknown test type: {self.test_type}")

    # ------------------------------------------------------------------ multiple testing
    def multiple_testing_correction(self, pvalues: List[float], method: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
        """``stats.py:159-185`` -> statsmodels ``multipletests(pvalues, alpha, method)``."""
        method = self.fdr_method if method is None else method
        method = {"benjamini_hochberg": "fdr_bh"}.get(method, method)
        p = np.asarray(pvalues, dtype=float)
        m = len(p)
        if m == 0:
            return np.array([], dtype=bool), np.array([])
        if method == "fdr_bh":
            return nv.default_context().bh(p, float(self.alpha))
        if method == "bonferroni":
            return p <= self.alpha / m, np.minimum(p * m, 1.0)
        if method == "holm":
            order = np.argsort(p, kind="stable")
            sp, steps = p[order], np.arange(m, 0, -1)
            fail = np.nonzero(sp > self.alpha / steps)[0]
            reject_s = np.arange(m) < (fail.min() if len(fail) else m)
            corr_s = np.minimum(np.maximum.accumulate(sp * steps), 1.0)
            reject, corr = np.empty(m, dtype=bool), np.empty(m)
            reject[order], corr[order] = reject_s, corr_s
            return reject, corr
        raise ValueError(f"method not recognized: {method}")

    def test_multiple_variants(self, variant_data: pd.DataFrame, error_rates: Dict[str, float]) -> pd.DataFrame:
        """``stats.py:187-288``: per-locus one-sided tests + FDR correction, same validation, same
        output columns; the loci go through K8 / K9 as arrays."""
        if variant_data.empty:
            return pd.DataFrame()
        missing = [c for c in ["alt_count", "total_depth"] if c not in variant_data.columns]
        if missing:
            raise ValueError(f"Missing required columns: {missing}")
        if not pd.api.types.is_numeric_dtype(variant_data["alt_count"]):
            raise ValueError("alt_count column must be numeric")
        if not pd.api.types.is_numeric_dtype(variant_data["total_depth"]):
            raise ValueError("total_depth column must be numeric")
        if (variant_data["alt_count"] < 0).any():
            raise ValueError("alt_count cannot be negative")
        if (variant_data["total_depth"] <= 0).any():
            raise ValueError("total_depth must be positive")
        if (variant_data["alt_count"] > variant_data["total_depth"]).any():
            raise ValueError("alt_count cannot exceed total_depth")
        m = len(variant_data)
        context = variant_data["context"].tolist() if "context" in variant_data else ["default"] * m
        rate = np.empty(m)
        for i, c in enumerate(context):
            r = error_rates.get(c, 1e-4)
            rate[i] = r if isinstance(r, (int, float)) and 0 <= r <= 1 else 1e-4      # :234-236 fallback
        k = variant_data["alt_count"].values.astype(np.int64)
        n = variant_data["total_depth"].values.astype(np.int64)
        if self.test_type == "poisson":
            expected = rate * n                                                       # :144
            p = _tails(k, None, expected, nv.TEST_POISSON, "greater")
            p = np.where(expected <= 0, 1.0, p)                                       # :54-60
            with np.errstate(divide="ignore"):
                effect = np.where(expected > 1e-12, np.log(k + 1e-8) - np.log(expected + 1e-8), np.nan)
            statistic = k.astype(float)
        else:
            p = _tails(k, n, rate, nv.TEST_BINOMIAL, "greater")
            obs = k / n
            inside = (rate > 1e-12) & (rate < 1 - 1e-12)
            with np.errstate(divide="ignore", invalid="ignore"):
                logit = (np.log(obs + 1e-12) - np.log(1 - obs + 1e-12)) - (np.log(rate + 1e-12) - np.log(1 - rate + 1e-12))
            effect = np.where(inside, logit, np.clip(obs / (rate + 1e-12), 1e-12, 1e12))
            statistic = obs
        df = pd.DataFrame({
            "site_key": variant_data["site_key"].values if "site_key" in variant_data else "",
            "context": context, "observed_alt": k, "total_depth": n, "expected_rate": rate,
            "test_statistic": statistic, "pvalue": p, "effect_size": effect, "method": self.test_type,
        })
        try:
            rejected, qvalues = self.multiple_testing_correction(p)
        except Exception as e:  # stats.py:270-276: fall back to uncorrected calls
            logging.warning(f"Multiple testing correction failed: {e}")
            rejected, qvalues = p < self.alpha, p
        df["qvalue"] = qvalues
        df["significant"] = rejected
        df["significant_uncorrected"] = df["pvalue"] < self.alpha
        return df
