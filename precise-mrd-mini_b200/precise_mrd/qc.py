"""Quality-control guard-rails (mirror of the reference's ``src/precise_mrd/qc.py:22-396``).

``QCThresholds`` / ``SiteQCResult`` / ``QCAnalyzer`` keep the reference's names, arguments, dictionary
keys and messages.  The one reduction that scales with the data -- the family-size distribution and the
consensus-efficiency counts over the family table -- runs on the device (kernel K16,
``pmrd_qc_family_stats``): an exact size histogram plus the count of families with a consensus.  Every
statistic of the report is then exact integer arithmetic on that histogram (count, min, max, mean,
median, quartiles, the Counter, below / above threshold) or a few floating-point operations on it (std,
skewness, kurtosis: equal to NumPy / SciPy to rounding).  The per-site depth tables have one row per
(site, allele) -- tens of rows -- and stay host-side pandas, as in the reference.

``report_from_tables`` goes straight from the device collapse tables (``Context.collapse_reads``) to the
report, without the reference's list-of-dicts detour."""
from __future__ import annotations

import json
import logging
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Union

import numpy as np
import pandas as pd

from . import _native as nv

QC_BINS = 4096          # exact sizes 0 .. 4094; larger families fall back to an exact host count of the tail


@dataclass
class QCThresholds:                                  # qc.py:22-29
    min_total_umi_families: int = 1000
    max_contamination_rate: float = 0.05
    min_depth_per_locus: int = 100
    min_family_size: int = 3
    max_family_size_outlier: int = 1000
    min_consensus_rate: float = 0.8


@dataclass
class SiteQCResult:                                  # qc.py:32-43
    site_key: str
    total_families: int
    consensus_families: int
    consensus_rate: float
    mean_family_size: float
    depth_distribution: Dict[str, float]
    contamination_estimate: float
    passes_qc: bool
    failed_criteria: List[str]


def _size_histogram(sizes: np.ndarray, has_consensus: Optional[np.ndarray] = None):
    """(hist {size: count} as two int arrays, n_consensus) of a family table, reduced on the device."""
    sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
    flags = None if has_consensus is None else np.ascontiguousarray(np.asarray(has_consensus, dtype=bool).astype(np.uint32))
    hist, n_cons, vmax = nv.default_context().qc_family_stats(sizes, flags, 1, QC_BINS)
    values = np.flatnonzero(hist[:-1]).astype(np.int64)
    counts = hist[values].astype(np.int64)
    if hist[-1]:                                     # families beyond the histogram: exact tail on the host
        tail_v, tail_c = np.unique(sizes[sizes >= QC_BINS - 1].astype(np.int64), return_counts=True)
        values, counts = np.concatenate([values, tail_v]), np.concatenate([counts, tail_c])
    return values, counts, int(n_cons), int(vmax)


def _order_stat(values: np.ndarray, cum: np.ndarray, k: int) -> int:
    """k-th smallest (0-based) of the multiset given by (values, cumulative counts)."""
    return int(values[int(np.searchsorted(cum, k, side="right"))])


def _percentile(values, cum, n: int, q: float) -> float:
    """numpy.percentile(method='linear') on the multiset, including NumPy's lerp form."""
    pos = q / 100.0 * (n - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    t = pos - lo
    a, b = float(_order_stat(values, cum, lo)), float(_order_stat(values, cum, hi))
    d = b - a
    return b - d * (1.0 - t) if t >= 0.5 else a + d * t


class QCAnalyzer:
    """qc.py:46-396."""

    def __init__(self, thresholds: Optional[QCThresholds] = None):
        self.thresholds = thresholds or QCThresholds()
        self.logger = logging.getLogger(__name__)

    # ------------------------------------------------------------------ family sizes (K16)
    def analyze_family_size_distribution(self, family_sizes) -> Dict[str, Union[float, int, Dict]]:
        if family_sizes is None or len(family_sizes) == 0:
            return {"error": "No family sizes provided"}
        values, counts, _, _ = _size_histogram(np.asarray(family_sizes))
        return self._size_stats(values, counts, order=np.asarray(family_sizes))

    def _size_stats(self, values: np.ndarray, counts: np.ndarray, order: Optional[np.ndarray] = None) -> Dict[str, Any]:
        n = int(counts.sum())
        cum = np.cumsum(counts)
        s1 = int((values * counts).sum())
        mean = s1 / n
        dev = values.astype(np.float64) - mean
        m2 = float((counts * dev**2).sum() / n)
        lo_mid, hi_mid = _order_stat(values, cum, (n - 1) // 2), _order_stat(values, cum, n // 2)
        out: Dict[str, Any] = {
            "count": n, "mean": float(mean), "median": float((lo_mid + hi_mid) / 2.0), "std": float(np.sqrt(m2)),
            "min": int(values[0]), "max": int(values[-1]),
            "q25": _percentile(values, cum, n, 25), "q75": _percentile(values, cum, n, 75),
        }
        if n > 1:
            m3 = float((counts * dev**3).sum() / n)
            m4 = float((counts * dev**4).sum() / n)
            # scipy.stats.skew / kurtosis (biased, Fisher): NaN when the data are constant
            with np.errstate(invalid="ignore", divide="ignore"):
                out["skewness"] = float(m3 / m2**1.5) if m2 > 0 else float("nan")
                out["kurtosis"] = float(m4 / m2**2 - 3.0) if m2 > 0 else float("nan")
        if order is not None:                        # Counter keeps first-appearance order
            _, first = np.unique(order, return_index=True)
            keys = [int(order[i]) for i in np.sort(first)]
        else:
            keys = [int(v) for v in values]
        by = {int(v): int(c) for v, c in zip(values, counts)}
        out["size_distribution"] = {k: by[k] for k in keys}
        issues = []
        if out["mean"] < 2:
            issues.append("low_mean_family_size")
        if out["max"] > 100:
            issues.append("extremely_large_families")
        if len(by) == 1:
            issues.append("no_size_variation")
        out["potential_issues"] = issues
        return out

    # ------------------------------------------------------------------ depth (host: one row per site x allele)
    def calculate_depth_metrics(self, consensus_data: pd.DataFrame) -> Dict[str, Union[float, int]]:
        if consensus_data.empty:
            return {"error": "No consensus data provided"}
        site_depths = consensus_data.groupby("site_key")["consensus_count"].sum()
        th = self.thresholds
        m = {
            "n_sites": len(site_depths), "total_depth": int(site_depths.sum()), "mean_depth_per_site": float(site_depths.mean()),
            "median_depth_per_site": float(site_depths.median()), "min_depth_per_site": int(site_depths.min()),
            "max_depth_per_site": int(site_depths.max()), "std_depth_per_site": float(site_depths.std()),
            "sites_below_min_depth": int((site_depths < th.min_depth_per_locus).sum()),
            "fraction_sites_below_min_depth": float((site_depths < th.min_depth_per_locus).mean()),
        }
        if len(site_depths) > 1:
            m["depth_uniformity_cv"] = float(site_depths.std() / site_depths.mean())
            med = site_depths.median()
            m["fraction_within_2fold_median"] = float(((site_depths >= med / 2) & (site_depths <= med * 2)).mean())
        return m

    def estimate_contamination_metrics(self, sample_data: pd.DataFrame, reference_patterns=None) -> Dict[str, float]:
        out = {"estimated_contamination_rate": 0.0, "cross_sample_evidence": 0.0, "unexpected_allele_fraction": 0.0}
        if sample_data.empty:
            return out
        total = len(sample_data)
        depth = sample_data.groupby("site_key")["consensus_count"].transform("sum")
        ref0 = sample_data.groupby("site_key")["ref"].transform("first")                   # qc.py:165 (.iloc[0] of the site)
        vaf = sample_data["consensus_count"] / depth
        unexpected = int(((sample_data["allele"] != ref0) & (vaf >= 0.001) & (vaf <= 0.05)).sum())
        out["unexpected_allele_fraction"] = unexpected / total if total > 0 else 0.0
        out["estimated_contamination_rate"] = min(out["unexpected_allele_fraction"], 0.1)
        return out

    # ------------------------------------------------------------------ consensus efficiency (K16)
    def analyze_consensus_efficiency(self, families_data: List[Dict[str, Any]]) -> Dict[str, float]:
        if not families_data:
            return {"error": "No families data provided"}
        sizes = np.fromiter((f.get("family_size", 0) for f in families_data), dtype=np.int64, count=len(families_data))
        has = np.fromiter((f.get("consensus_allele") is not None for f in families_data), dtype=bool, count=len(families_data))
        values, counts, n_cons, _ = _size_histogram(sizes, has)
        return self._efficiency(values, counts, n_cons)

    def _efficiency(self, values, counts, n_cons: int) -> Dict[str, float]:
        th = self.thresholds
        total = int(counts.sum())
        below = int(counts[values < th.min_family_size].sum())
        above = int(counts[values > th.max_family_size_outlier].sum())
        eff = total - below - above
        return {
            "total_families": total, "consensus_families": int(n_cons), "consensus_rate": n_cons / total if total > 0 else 0.0,
            "families_below_min_size": below, "families_above_max_size": above,
            "fraction_below_min_size": below / total if total > 0 else 0.0,
            "fraction_above_max_size": above / total if total > 0 else 0.0,
            "effective_families": eff, "effective_family_rate": eff / total if total > 0 else 0.0,
        }

    # ------------------------------------------------------------------ guard-rails
    def validate_clinical_guardrails(self, qc_metrics: Dict[str, Any]) -> Dict[str, Any]:
        th = self.thresholds
        res = {"sample_valid": True, "failed_criteria": [], "warnings": []}
        total = qc_metrics.get("total_families", 0)
        if total < th.min_total_umi_families:
            res["sample_valid"] = False
            res["failed_criteria"].append(f"insufficient_total_families: {total} < {th.min_total_umi_families}")
        rate = qc_metrics.get("estimated_contamination_rate", 0.0)
        if rate > th.max_contamination_rate:
            res["sample_valid"] = False
            res["failed_criteria"].append(f"high_contamination: {rate:.4f} > {th.max_contamination_rate}")
        cr = qc_metrics.get("consensus_rate", 0.0)
        if cr < th.min_consensus_rate:
            res["warnings"].append(f"low_consensus_rate: {cr:.4f} < {th.min_consensus_rate}")
        md = qc_metrics.get("mean_depth_per_site", 0)
        if md < th.min_depth_per_locus:
            res["warnings"].append(f"low_mean_depth: {md:.1f} < {th.min_depth_per_locus}")
        return res

    def generate_site_qc_report(self, site_key: str, consensus_data: pd.DataFrame, families_data: List[Dict[str, Any]]) -> SiteQCResult:
        site_cons = consensus_data[consensus_data["site_key"] == site_key]
        site_fams = [f for f in families_data if f.get("site_key") == site_key]
        return self._site_result(site_key, site_cons, len(site_fams),
                                 sum(1 for f in site_fams if f.get("consensus_allele") is not None),
                                 float(np.mean([f.get("family_size", 0) for f in site_fams])) if site_fams else 0.0)

    def _site_result(self, site_key, site_cons: pd.DataFrame, total: int, n_cons: int, mean_size: float) -> SiteQCResult:
        th = self.thresholds
        rate = n_cons / total if total > 0 else 0.0
        if not site_cons.empty:
            dist = {"total_depth": site_cons["consensus_count"].sum(), "n_alleles": len(site_cons),
                    "max_allele_depth": site_cons["consensus_count"].max()}
        else:
            dist = {"total_depth": 0, "n_alleles": 0, "max_allele_depth": 0}
        failed = []
        if total < th.min_depth_per_locus:
            failed.append("insufficient_families")
        if rate < th.min_consensus_rate:
            failed.append("low_consensus_rate")
        if mean_size < th.min_family_size:
            failed.append("low_mean_family_size")
        return SiteQCResult(site_key=site_key, total_families=total, consensus_families=n_cons, consensus_rate=rate,
                            mean_family_size=mean_size, depth_distribution=dist, contamination_estimate=0.0,
                            passes_qc=len(failed) == 0, failed_criteria=failed)

    def generate_comprehensive_qc_report(self, consensus_data: pd.DataFrame, families_data: List[Dict[str, Any]],
                                         simulation_metadata: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        sizes = [f.get("family_size", 0) for f in families_data]
        overall = {**self.analyze_family_size_distribution(sizes), **self.calculate_depth_metrics(consensus_data),
                   **self.estimate_contamination_metrics(consensus_data), **self.analyze_consensus_efficiency(families_data)}
        sites = [self.generate_site_qc_report(k, consensus_data, families_data).__dict__
                 for k in consensus_data["site_key"].unique()[:10]]            # qc.py:349: the first 10 sites
        return self._assemble(overall, sites, simulation_metadata)

    def _assemble(self, overall, sites, simulation_metadata):
        passing = sum(1 for s in sites if s["passes_qc"])
        return {"overall_metrics": overall, "clinical_validation": self.validate_clinical_guardrails(overall),
                "site_qc_summary": {"total_sites_analyzed": len(sites), "sites_passing_qc": passing,
                                    "site_pass_rate": passing / len(sites) if sites else 0.0},
                "detailed_site_qc": sites, "simulation_metadata": simulation_metadata or {}}

    # ------------------------------------------------------------------ from the device collapse tables
    def report_from_tables(self, fam: Dict[str, np.ndarray], grp: Dict[str, np.ndarray], site_keys: Optional[List[str]] = None,
                           ref_allele: Optional[np.ndarray] = None, simulation_metadata: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        """The comprehensive report straight from ``Context.collapse_reads`` output (weighted mode): ``fam`` the
        family table (key = group << 32 | umi, size, flags), ``grp`` the group table (group, family_count,
        consensus_count [G, 4]).  ``site_keys[g]`` names group g (default ``"site<g>"``), ``ref_allele[g]`` its
        reference allele code (default 0).  Same dictionary as ``generate_comprehensive_qc_report``."""
        g_id = np.asarray(grp["group"], dtype=np.int64)
        name = (lambda g: site_keys[g]) if site_keys is not None else (lambda g: f"site{g}")
        ref = np.zeros(int(g_id.max()) + 1 if len(g_id) else 1, np.int64) if ref_allele is None else np.asarray(ref_allele, dtype=np.int64)
        rows = [{"site_key": name(int(g)), "allele": "ACGT"[a], "ref": "ACGT"[int(ref[int(g)])], "consensus_count": int(c)}
                for g, cc in zip(g_id, np.asarray(grp["consensus_count"])) for a, c in enumerate(cc) if c > 0]
        consensus = pd.DataFrame(rows, columns=["site_key", "allele", "ref", "consensus_count"])
        sizes = np.asarray(fam["size"])
        has = (np.asarray(fam["flags"]) & nv.FAM_HAS_CONSENSUS) != 0
        values, counts, n_cons, _ = _size_histogram(sizes, has)
        overall = {**self._size_stats(values, counts, order=sizes), **self.calculate_depth_metrics(consensus),
                   **self.estimate_contamination_metrics(consensus), **self._efficiency(values, counts, n_cons)}
        fam_group = (np.asarray(fam["key"]) >> np.uint64(32)).astype(np.int64)
        sites = []
        for key in consensus["site_key"].unique()[:10]:
            g = next(int(x) for x in g_id if name(int(x)) == key)
            sel = fam_group == g
            sites.append(self._site_result(key, consensus[consensus["site_key"] == key], int(sel.sum()), int(has[sel].sum()),
                                           float(sizes[sel].mean()) if sel.any() else 0.0).__dict__)
        return self._assemble(overall, sites, simulation_metadata)

    def export_qc_metrics(self, qc_report: Dict[str, Any], filepath: str) -> None:        # qc.py:373-396
        def conv(o):
            if isinstance(o, np.integer):
                return int(o)
            if isinstance(o, np.floating):
                return float(o)
            if isinstance(o, np.ndarray):
                return o.tolist()
            if isinstance(o, dict):
                return {k: conv(v) for k, v in o.items()}
            if isinstance(o, list):
                return [conv(v) for v in o]
            return o

        with open(filepath, "w") as f:
            json.dump(conv(qc_report), f, indent=2)
