"""Stratified power and calibration -- drop-in for ``src/precise_mrd/eval/stratified.py``
(``StratifiedAnalyzer``, ``run_stratified_analysis``; ``north_star`` config 5's "stratified
trinucleotide-context power").

Same class, method names, defaults and output files as the reference.  Each (context, depth, AF,
replicate) cell is one ``simulate_reads -> collapse_umis -> fit_error_model -> call_mrd`` chain
of CUDA stage calls; the context's background multiplier (CpG 1.5, CHG 1.2, CHH 1.0, NpN 0.8,
``stratified.py:186-191``) is applied inside kernel K1 (``pmrd_simulate_cells_stratified``) --
twice to the false-positive rate, as the reference does (``:193`` and ``:199``).

Positions taken on reference defects:

* ``hash(context) % 1000`` (``:61``) is salted per interpreter process, so the reference's
  replicate seeds are not reproducible; here the context offset is ``crc32(context) % 1000``.
* ``calls_df.get('context', 'NpN') == context`` (``:76``) indexes the frame with a Python bool and
  raises ``KeyError`` on every non-empty call table, and ``variant_call`` is never produced
  (Appendix A #1).  Here every call of a context-specific run belongs to that context, and
  ``variant_call == significant``.
* ``calls_df.get('variant_call', False).tolist()`` (``:145-148``) fails for the same reason; the
  labels are the ``variant_call`` column."""
from __future__ import annotations

import json
import zlib
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import pandas as pd

from ..call import call_mrd
from ..collapse import collapse_umis
from ..config import PipelineConfig
from ..error_model import fit_error_model
from ..simulate import simulate_reads

CONTEXT_ERROR_MULTIPLIER = {"CpG": 1.5, "CHG": 1.2, "CHH": 1.0, "NpN": 0.8}   # stratified.py:186-191


def context_seed_offset(context: str) -> int:
    """Deterministic stand-in for the reference's per-process ``hash(context) % 1000``."""
    return zlib.crc32(context.encode()) % 1000


class StratifiedAnalyzer:
    """Analyzer for stratified power and calibration analysis."""

    def __init__(self, config: PipelineConfig, rng: np.random.Generator):
        self.config = config
        self.rng = rng
        self.power_results: Optional[Dict[str, Any]] = None
        self.calibration_results: Optional[Dict[str, Any]] = None

    # ------------------------------------------------------------------ power
    def analyze_stratified_power(self, af_values: List[float] = [0.001, 0.005, 0.01, 0.05],
                                 depth_values: List[int] = [1000, 5000, 10000],
                                 contexts: List[str] = ["CpG", "CHG", "CHH", "NpN"],
                                 n_replicates: int = 50) -> Dict[str, Any]:
        print("Running stratified power analysis...")
        power: Dict[str, Any] = {}
        for context in contexts:
            print(f"  Analyzing context: {context}")
            power[context] = {}
            for depth in depth_values:
                power[context][depth] = {}
                for af in af_values:
                    rates = []
                    for rep in range(n_replicates):
                        run_rng = np.random.default_rng(self.config.seed + rep * 1000 + context_seed_offset(context))
                        cfg = self._create_context_config(af, depth, context)
                        reads_df = self._simulate_with_context(cfg, run_rng, context)
                        collapsed_df = collapse_umis(reads_df, cfg, run_rng)
                        error_model = fit_error_model(collapsed_df, cfg, run_rng)
                        calls_df = call_mrd(collapsed_df, error_model, cfg, run_rng)
                        if len(calls_df) > 0:
                            rates.append(float((calls_df["variant_call"] == True).sum()) / max(1, len(calls_df)))  # noqa: E712
                        else:
                            rates.append(0.0)
                    power[context][depth][af] = {
                        "mean_detection_rate": float(np.mean(rates)), "std_detection_rate": float(np.std(rates)),
                        "detection_rates": rates, "n_replicates": n_replicates,
                    }
                    print(f"    {context} @ depth={depth}, AF={af:.0e}: {np.mean(rates):.3f} ± {np.std(rates):.3f}")
        self.power_results = {"stratified_results": power, "af_values": af_values, "depth_values": depth_values,
                              "contexts": contexts, "config_hash": self.config.config_hash()}
        return self.power_results

    # ------------------------------------------------------------------ calibration
    def analyze_calibration_by_bins(self, af_values: List[float] = [0.001, 0.005, 0.01, 0.05],
                                    depth_values: List[int] = [1000, 5000, 10000], n_bins: int = 10,
                                    n_replicates: int = 100) -> Dict[str, Any]:
        print(f"Running calibration analysis with {n_bins} bins...")
        rows = []
        for depth in depth_values:
            print(f"  Processing depth: {depth}")
            for af in af_values:
                probs: List[float] = []
                labels: List[bool] = []
                for rep in range(n_replicates):
                    run_rng = np.random.default_rng(self.config.seed + rep * 2000 + int(af * 1e6) + depth)   # :128-130
                    cfg = self._create_calibration_config(af, depth)
                    reads_df = simulate_reads(cfg, run_rng)
                    collapsed_df = collapse_umis(reads_df, cfg, run_rng)
                    error_model = fit_error_model(collapsed_df, cfg, run_rng)
                    calls_df = call_mrd(collapsed_df, error_model, cfg, run_rng)
                    if len(calls_df) > 0:
                        probs.extend((1.0 - calls_df["p_value"]).tolist())          # :144 p-value as inverted score
                        labels.extend(calls_df["variant_call"].astype(bool).tolist())
                if probs:
                    m = self._compute_calibration_metrics(np.array(probs), np.array(labels), n_bins)
                    rows.append({"depth": depth, "af": af, "ece": m["ece"], "max_ce": m["max_ce"],
                                 "bin_accuracies": m["bin_accuracies"], "bin_confidences": m["bin_confidences"],
                                 "bin_counts": m["bin_counts"], "n_samples": len(probs)})
        self.calibration_results = {"calibration_data": rows, "af_values": af_values, "depth_values": depth_values,
                                    "n_bins": n_bins, "config_hash": self.config.config_hash()}
        return self.calibration_results

    # ------------------------------------------------------------------ helpers
    def _simulate_with_context(self, config: PipelineConfig, rng: np.random.Generator, context: str) -> pd.DataFrame:
        """stratified.py:180-203: context tag + background multiplier (unknown context: 1.0)."""
        reads_df = simulate_reads(config, rng, background_multiplier=CONTEXT_ERROR_MULTIPLIER.get(context, 1.0))
        reads_df["context"] = context
        return reads_df

    def _cell_config(self, tag: str, af: float, depth: int) -> PipelineConfig:
        sim = self.config.simulation
        return PipelineConfig(run_id=f"{self.config.run_id}_{tag}", seed=self.config.seed,
                              simulation=type(sim)(allele_fractions=[af], umi_depths=[depth], n_replicates=1,
                                                   n_bootstrap=sim.n_bootstrap),
                              umi=self.config.umi, stats=self.config.stats, lod=self.config.lod)

    def _create_context_config(self, af: float, depth: int, context: str) -> PipelineConfig:     # :205-220
        return self._cell_config(f"context_{context}", af, depth)

    def _create_calibration_config(self, af: float, depth: int) -> PipelineConfig:               # :222-237
        return self._cell_config("calibration", af, depth)

    def _compute_calibration_metrics(self, predicted_probs: np.ndarray, true_labels: np.ndarray,
                                     n_bins: int) -> Dict[str, Any]:
        """ECE / MCE over ``n_bins`` equal-width bins ``(lo, hi]`` (stratified.py:239-287): a
        prediction of exactly 0 falls in no bin."""
        edges = np.linspace(0, 1, n_bins + 1)
        acc, conf, cnt = [], [], []
        ece = max_ce = 0.0
        for lo, hi in zip(edges[:-1], edges[1:]):
            in_bin = (predicted_probs > lo) & (predicted_probs <= hi)
            share = in_bin.mean()
            if share > 0:
                a, c = true_labels[in_bin].mean(), predicted_probs[in_bin].mean()
                ce = abs(c - a)
                ece += ce * share
                max_ce = max(max_ce, ce)
                acc.append(float(a)); conf.append(float(c)); cnt.append(int(in_bin.sum()))
            else:
                acc.append(0.0); conf.append(0.0); cnt.append(0)
        return {"ece": float(ece), "max_ce": float(max_ce), "bin_accuracies": acc, "bin_confidences": conf, "bin_counts": cnt}

    # ------------------------------------------------------------------ reports
    def generate_stratified_reports(self, output_dir: str = "reports") -> None:
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        if self.power_results:
            payload = dict(self.power_results)
            payload["schema_version"] = "1.0.0"
            with open(out / "power_by_stratum.json", "w") as f:
                json.dump(payload, f, indent=2)
            print(f"Stratified power results saved to {out / 'power_by_stratum.json'}")
        if self.calibration_results:
            pd.DataFrame(self.calibration_results["calibration_data"]).to_csv(out / "calibration_by_bin.csv", index=False)
            payload = dict(self.calibration_results)
            payload["schema_version"] = "1.0.0"
            with open(out / "calibration_by_bin.json", "w") as f:
                json.dump(payload, f, indent=2)
            print(f"Calibration by bin results saved to {out / 'calibration_by_bin.csv'}")


def run_stratified_analysis(config: PipelineConfig, rng: np.random.Generator,
                            output_dir: str = "reports") -> Tuple[Dict[str, Any], Dict[str, Any]]:
    """stratified.py:316-338: power, then calibration, then the three report files."""
    analyzer = StratifiedAnalyzer(config, rng)
    power = analyzer.analyze_stratified_power()
    calibration = analyzer.analyze_calibration_by_bins()
    analyzer.generate_stratified_reports(output_dir)
    return power, calibration
