"""Quality filters -- drop-in for ``src/precise_mrd/filters.py:19-342`` (``FilterResult``,
``QualityFilter``); SURVEY.md §8(f)-3.

The two statistical filters run on the device: Fisher's exact strand-bias test is kernel K14
(``pmrd_fisher_exact``), the end-repair enrichment test is the one-sided binomial tail of K8
(``pmrd_pvalues``).  ``filter_variant_dataframe`` sends the whole variant table through ONE launch
of each instead of a Python loop of scipy calls; the threshold comparisons and reason strings
around them are host plumbing.  Scalar methods keep the reference's signatures."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd

from . import _native as nv


@dataclass
class FilterResult:
    """Container for filter results."""
    passed: bool
    reason: Optional[str] = None
    statistic: Optional[float] = None
    pvalue: Optional[float] = None


class QualityFilter:
    """Apply quality filters to variant calls."""

    def __init__(self, strand_bias_threshold: float = 0.01, end_repair_filter: bool = True, end_repair_distance: int = 10,
                 end_repair_contexts: Optional[List[str]] = None, min_alt_count: int = 2, min_total_depth: int = 10,
                 min_qual_score: int = 20):
        self.strand_bias_threshold = strand_bias_threshold
        self.end_repair_filter = end_repair_filter
        self.end_repair_distance = end_repair_distance
        self.end_repair_contexts = end_repair_contexts or ["G>T", "C>A"]
        self.min_alt_count = min_alt_count
        self.min_total_depth = min_total_depth
        self.min_qual_score = min_qual_score

    # ------------------------------------------------------------------ batched device tests
    def _strand_bias(self, alt_f, alt_r, ref_f, ref_r):
        """Vectorised filters.py:50-100 -> (passed, reason, odds_ratio, pvalue); the last two are NaN
        where the table holds fewer than four reads (no test is run there)."""
        alt_f, alt_r, ref_f, ref_r = (np.asarray(v, dtype=np.int64) for v in (alt_f, alt_r, ref_f, ref_r))
        tables = np.stack([ref_f, ref_r, alt_f, alt_r], axis=1)                 # rows: ref, alt; cols: forward, reverse
        odds, p = nv.default_context().fisher_exact(tables)
        enough = tables.sum(axis=1) >= 4
        total_alt = alt_f + alt_r
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = alt_f / total_alt
        biased = p < self.strand_bias_threshold
        biased &= np.where(total_alt > 0, (ratio > 0.95) | (ratio < 0.05), True)  # :83-88 extreme-imbalance gate
        biased &= enough
        reason = np.where(~enough, "insufficient_data_for_strand_bias_test", np.where(biased, "strand_bias", None))
        return ~biased, reason, np.where(enough, odds, np.nan), np.where(enough, p, np.nan)

    def _end_repair(self, is_context, n_near, n_total, read_length: int = 150):
        """Vectorised filters.py:102-166 for variants whose mutation type is an end-repair context."""
        n_near, n_total = np.asarray(n_near, dtype=np.int64), np.asarray(n_total, dtype=np.int64)
        tested = np.asarray(is_context, dtype=bool) & (n_total > 0) & bool(self.end_repair_filter)
        expected = 2 * self.end_repair_distance / read_length
        p = np.full(len(n_total), np.nan)
        if tested.any():
            p[tested] = nv.default_context().pvalues(n_near[tested], n_total[tested], [expected], nv.TEST_BINOMIAL, nv.ALT_GREATER)
        with np.errstate(divide="ignore", invalid="ignore"):
            observed = n_near / n_total
        artifact = tested & (p < 0.05) & (observed > 2 * expected)
        return ~artifact, np.where(artifact, "end_repair_artifact", None), np.where(tested, observed, np.nan), p

    # ------------------------------------------------------------------ scalar API
    def test_strand_bias(self, alt_forward: int, alt_reverse: int, ref_forward: int, ref_reverse: int) -> FilterResult:
        passed, reason, odds, p = self._strand_bias([alt_forward], [alt_reverse], [ref_forward], [ref_reverse])
        if reason[0] == "insufficient_data_for_strand_bias_test":
            return FilterResult(passed=True, reason=reason[0])
        return FilterResult(passed=bool(passed[0]), reason=reason[0], statistic=float(odds[0]), pvalue=float(p[0]))

    def test_end_repair_artifact(self, ref: str, alt: str, read_positions: List[int], read_length: int = 150) -> FilterResult:
        if not self.end_repair_filter or f"{ref}>{alt}" not in self.end_repair_contexts or len(read_positions) == 0:
            return FilterResult(passed=True)
        pos = np.asarray(read_positions)
        near = int(((pos <= self.end_repair_distance) | (pos >= read_length - self.end_repair_distance)).sum())
        passed, reason, stat, p = self._end_repair([True], [near], [len(pos)], read_length)
        return FilterResult(passed=bool(passed[0]), reason=reason[0], statistic=float(stat[0]), pvalue=float(p[0]))

    def filter_by_depth(self, alt_count: int, total_depth: int) -> FilterResult:
        if alt_count < self.min_alt_count:
            return FilterResult(passed=False, reason="insufficient_alt_count", statistic=alt_count)
        if total_depth < self.min_total_depth:
            return FilterResult(passed=False, reason="insufficient_total_depth", statistic=total_depth)
        return FilterResult(passed=True)

    def filter_by_quality(self, quality_scores: List[int]) -> FilterResult:
        if len(quality_scores) == 0:
            return FilterResult(passed=False, reason="no_quality_scores")
        mean_quality = np.mean(quality_scores)
        if mean_quality < self.min_qual_score:
            return FilterResult(passed=False, reason="low_quality", statistic=mean_quality)
        return FilterResult(passed=True)

    def apply_all_filters(self, variant_data: Dict[str, Any]) -> Dict[str, FilterResult]:
        """filters.py:205-243: which filters run depends on which keys the variant carries."""
        out = {"depth": self.filter_by_depth(alt_count=variant_data.get("alt_count", 0), total_depth=variant_data.get("total_depth", 0))}
        if "quality_scores" in variant_data:
            out["quality"] = self.filter_by_quality(quality_scores=variant_data["quality_scores"])
        if all(k in variant_data for k in ("alt_forward", "alt_reverse", "ref_forward", "ref_reverse")):
            out["strand_bias"] = self.test_strand_bias(variant_data["alt_forward"], variant_data["alt_reverse"],
                                                       variant_data["ref_forward"], variant_data["ref_reverse"])
        if all(k in variant_data for k in ("ref", "alt", "read_positions")):
            out["end_repair"] = self.test_end_repair_artifact(variant_data["ref"], variant_data["alt"], variant_data["read_positions"])
        return out

    def summarize_filters(self, filter_results: Dict[str, FilterResult]) -> Dict[str, Any]:
        passed = [n for n, r in filter_results.items() if r.passed]
        failed = [n for n, r in filter_results.items() if not r.passed]
        return {"all_filters_passed": len(failed) == 0, "n_filters_applied": len(filter_results), "n_filters_passed": len(passed),
                "n_filters_failed": len(failed), "passed_filters": passed, "failed_filters": failed,
                "failure_reasons": [r.reason for r in filter_results.values() if not r.passed and r.reason]}

    # ------------------------------------------------------------------ table API
    def filter_variant_dataframe(self, variants_df: pd.DataFrame) -> pd.DataFrame:
        """filters.py:262-296, column for column -- including its quirk that ``apply_all_filters``
        names the strand test 'strand_bias' and the table column is 'strand_bias_filter', so all
        four filters land in their columns."""
        if variants_df.empty:
            return variants_df.copy()
        df = variants_df.copy()
        n = len(df)
        cols = ["depth_filter", "quality_filter", "strand_bias_filter", "end_repair_filter"]
        passed = {c: np.ones(n, dtype=bool) for c in cols}
        reason = {c: np.full(n, None, dtype=object) for c in cols}
        alt = df["alt_count"].values if "alt_count" in df else np.zeros(n, np.int64)
        depth = df["total_depth"].values if "total_depth" in df else np.zeros(n, np.int64)
        low_alt, low_depth = alt < self.min_alt_count, depth < self.min_total_depth
        passed["depth_filter"] = ~(low_alt | low_depth)
        reason["depth_filter"] = np.where(low_alt, "insufficient_alt_count", np.where(low_depth, "insufficient_total_depth", None))
        if "quality_scores" in df:
            mean_q = np.array([np.mean(q) if len(q) else np.nan for q in df["quality_scores"]])
            empty = np.array([len(q) == 0 for q in df["quality_scores"]])
            low = mean_q < self.min_qual_score
            passed["quality_filter"] = ~(empty | low)
            reason["quality_filter"] = np.where(empty, "no_quality_scores", np.where(low, "low_quality", None))
        if all(c in df for c in ("alt_forward", "alt_reverse", "ref_forward", "ref_reverse")):
            ok, why, _, _ = self._strand_bias(df["alt_forward"].values, df["alt_reverse"].values, df["ref_forward"].values,
                                              df["ref_reverse"].values)
            passed["strand_bias_filter"] = ok
            reason["strand_bias_filter"] = np.where(ok, None, why)           # reasons are recorded for failures only (:288-289)
        if all(c in df for c in ("ref", "alt", "read_positions")):
            is_ctx = np.array([f"{r}>{a}" in self.end_repair_contexts for r, a in zip(df["ref"], df["alt"])])
            near = np.array([int(((np.asarray(p) <= self.end_repair_distance) | (np.asarray(p) >= 150 - self.end_repair_distance)).sum())
                             for p in df["read_positions"]], dtype=np.int64)
            tot = np.array([len(p) for p in df["read_positions"]], dtype=np.int64)
            ok, why, _, _ = self._end_repair(is_ctx, near, tot)
            passed["end_repair_filter"] = ok
            reason["end_repair_filter"] = why
        for c in cols:
            df[c] = passed[c]
            df[f"{c}_reason"] = reason[c]
        filter_cols = [c for c in df.columns if c.endswith("_filter") and not c.endswith("_reason")]
        df["all_filters_passed"] = df[filter_cols].all(axis=1)
        return df

    def generate_filter_report(self, filtered_variants: pd.DataFrame) -> Dict[str, Any]:
        """filters.py:298-342."""
        if filtered_variants.empty:
            return {"total_variants": 0}
        total = len(filtered_variants)
        report: Dict[str, Any] = {"total_variants": total,
                                  "variants_passing_all_filters": filtered_variants["all_filters_passed"].sum(),
                                  "overall_pass_rate": filtered_variants["all_filters_passed"].mean()}
        for col in [c for c in filtered_variants.columns if c.endswith("_filter") and not c.endswith("_reason")]:
            name = col.replace("_filter", "")
            n_passed = filtered_variants[col].sum()
            report[f"{name}_pass_rate"] = n_passed / total
            report[f"{name}_n_passed"] = n_passed
            report[f"{name}_n_failed"] = total - n_passed
        reasons: Dict[str, int] = {}
        for col in [c for c in filtered_variants.columns if c.endswith("_reason")]:
            for r in filtered_variants[col].dropna():
                reasons[r] = reasons.get(r, 0) + 1
        report["failure_reasons"] = reasons
        return report
