"""LoB / LoD / LoQ -- drop-in for ``src/precise_mrd/eval/lod.py`` (``LODAnalyzer``).

Two engines behind the reference's class and method signatures:

* ``engine="stages"`` (default, the PARITY mode): the reference's own triple loop
  (``lod.py:125-150``) -- one ``simulate_reads -> collapse_umis -> fit_error_model -> call_mrd``
  pipeline per (depth, AF, replicate) cell, seeded ``default_rng(seed + rep*1000 + int(af*1e6))``
  exactly as the reference does -- except that every stage is a CUDA kernel call, so a cell costs
  milliseconds instead of 3-7 s.  It reproduces the reference's behaviour including its
  degenerate outputs on its own grids (LoB 0, LoD NaN, LoQ None: SURVEY.md Appendix A #12).
* ``engine="grid"`` (the read-level mode ``north_star`` describes): all cells of the grid go
  through ONE device plan (``precise_mrd.grid``) on the fused engine: Philox read generation and
  quality-weighted consensus in one kernel, no read ever written (duplicate UMIs are found by hash
  bitmaps and replayed; results = the radix-sort collapse on family-major read order),
  blank-derived error rate, tail test per cell.  Cell
  keys are built from the INDICES (analysis, depth_i, af_i, rep), which removes the reference's
  seed collisions across AF/replicate pairs and depths.

The logistic detection-curve fit and its 200-resample bootstrap run on the device (kernel K11,
a statement of MINPACK lmdif: same optimum, same fallback-to-interpolation decisions)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import pandas as pd

from .. import _native as nv
from ..call import call_mrd
from ..collapse import collapse_umis
from ..config import PipelineConfig
from ..error_model import fit_error_model
from ..simulate import simulate_reads

ANALYSIS_LOB, ANALYSIS_LOD, ANALYSIS_LOQ = 1, 2, 3


class LODAnalyzer:
    """Analyzer for Limit of Blank, Detection, and Quantification."""

    def __init__(self, config: PipelineConfig, rng: np.random.Generator, engine: str = "stages"):
        if engine not in ("stages", "grid"):
            raise ValueError("engine must be 'stages' or 'grid'")
        self.config = config
        self.rng = rng
        self.engine = engine
        self.lob_results: Optional[Dict[str, Any]] = None
        self.lod_results: Optional[Dict[str, Any]] = None
        self.loq_results: Optional[Dict[str, Any]] = None

    # ------------------------------------------------------------------ helpers
    def _cell_config(self, run_id: str, af: float, depth: int) -> PipelineConfig:
        sim = self.config.simulation
        return PipelineConfig(run_id=f"{self.config.run_id}_{run_id}", seed=self.config.seed,
                              simulation=type(sim)(allele_fractions=[af], umi_depths=[depth], n_replicates=1,
                                                   n_bootstrap=sim.n_bootstrap),
                              umi=self.config.umi, stats=self.config.stats, lod=self.config.lod)

    def _create_blank_config(self) -> PipelineConfig:          # lod.py:317-332
        return self._cell_config("blank", 0.0, self.config.simulation.umi_depths[0])

    def _create_lod_config(self, af: float, depth: int) -> PipelineConfig:   # lod.py:334-349
        return self._cell_config("lod", af, depth)

    @staticmethod
    def _run_pipeline(cfg: PipelineConfig, run_rng: np.random.Generator) -> pd.DataFrame:
        reads_df = simulate_reads(cfg, run_rng)
        collapsed_df = collapse_umis(reads_df, cfg, run_rng)
        error_model = fit_error_model(collapsed_df, cfg, run_rng)
        return call_mrd(collapsed_df, error_model, cfg, run_rng)

    def _grid_cells(self, analysis: int, afs, depths, n_rep: int, n_blank_rep: int):
        """Cells of a read-level grid + per-depth blanks (the negative controls).
        cell_id = ((analysis*64 + depth_i)*64 + af_i)*2^20 + rep: indices, never seeds."""
        ids, af_l, dp_l = [], [], []
        for di, d in enumerate(depths):
            for ai, a in enumerate(list(afs) + [None]):
                reps = n_rep if a is not None else n_blank_rep
                slot = ai if a is not None else 63
                base = ((analysis * 64 + di) * 64 + slot) << 20
                ids.append(base + np.arange(reps, dtype=np.int64))
                af_l.append(np.full(reps, 0.0 if a is None else a))
                dp_l.append(np.full(reps, d, dtype=np.int64))
        return np.concatenate(ids), np.concatenate(af_l), np.concatenate(dp_l)

    def _run_grid(self, cell_id, af, depth) -> Dict[str, np.ndarray]:
        from .. import grid

        ctx = nv.default_context(int(self.config.seed))
        # one-sided (greater) test: a read-level "detection" is an EXCESS of alt families over the
        # blank-derived background (G-B stats.py:135-157); the two-sided G-D test also fires when a
        # cell sits significantly BELOW the U(0.5,2)-inflated expectation
        from .. import distributed

        # the FUSED engine (csrc/fused.cu): a cell goes from Philox counters to its counts without a read ever being written
        sim = grid.read_sim_params(self.config, fused=True)
        if distributed.world_size() > 1:
            # one process per GPU: cells sharded by family count, ONE all-gather of the per-cell counts, then
            # the call stage over the whole grid on every rank -- same table as on one GPU
            return distributed.run_cells_sharded(self.config, cell_id, af, depth, neg_af_max=0.0, alternative=nv.ALT_GREATER, ctx=ctx, sim=sim)
        return grid.run_cells(self.config, cell_id, af, depth, ctx=ctx, sim=sim, neg_af_max=0.0, alternative=nv.ALT_GREATER)

    # ------------------------------------------------------------------ LoB
    def estimate_lob(self, n_blank_runs: int = 100) -> Dict[str, Any]:
        print(f"Estimating LoB with {n_blank_runs} blank runs...")
        if self.engine == "stages":
            blank_config = self._create_blank_config()
            meas = []
            for i in range(n_blank_runs):
                run_rng = np.random.default_rng(self.config.seed + i)              # lod.py:60
                calls_df = self._run_pipeline(blank_config, run_rng)
                meas.append(int((calls_df["variant_call"] == True).sum()))         # noqa: E712  lod.py:71
            meas = np.array(meas)
        else:
            depth = self.config.simulation.umi_depths[0]
            cid = (ANALYSIS_LOB << 32) + np.arange(n_blank_runs, dtype=np.int64)
            res = self._run_grid(cid, np.zeros(n_blank_runs), np.full(n_blank_runs, depth, dtype=np.int64))
            meas = (res["p_value"] <= float(self.config.stats.alpha)).astype(int)
        self.lob_results = {
            "lob_value": float(np.percentile(meas, 95)), "blank_mean": float(np.mean(meas)), "blank_std": float(np.std(meas)),
            "blank_measurements": [int(v) for v in meas], "n_blank_runs": int(n_blank_runs), "percentile": 95,
            "config_hash": self.config.config_hash(),
        }
        print(f"LoB estimated: {self.lob_results['lob_value']:.3f}")
        return self.lob_results

    # ------------------------------------------------------------------ LoD
    def estimate_lod(self, af_range: Tuple[float, float] = (1e-4, 1e-2), depth_values: List[int] = [1000, 5000, 10000],
                     n_replicates: int = 50, target_detection_rate: float = 0.95, alpha: float = 0.05,
                     beta: float = 0.05) -> Dict[str, Any]:
        print(f"Estimating LoD across AF range {af_range} at depths {depth_values}...")
        af_values = np.logspace(np.log10(af_range[0]), np.log10(af_range[1]), 15)      # lod.py:121
        hits: Dict[int, List[float]] = {}
        if self.engine == "stages":
            for depth in depth_values:
                hits[depth] = []
                for af in af_values:
                    det = 0
                    for rep in range(n_replicates):
                        run_rng = np.random.default_rng(self.config.seed + rep * 1000 + int(af * 1e6))   # lod.py:136
                        calls_df = self._run_pipeline(self._create_lod_config(float(af), int(depth)), run_rng)
                        det += int((calls_df["variant_call"] == True).sum() > 0)   # noqa: E712
                    hits[depth].append(det / n_replicates)
        else:
            cid, af, dp = self._grid_cells(ANALYSIS_LOD, af_values, depth_values, n_replicates, max(n_replicates, 25))
            res = self._run_grid(cid, af, dp)
            detected = res["p_value"] <= float(self.config.stats.alpha)
            for depth in depth_values:
                hits[depth] = [float(detected[(dp == depth) & (af == a)].mean()) for a in af_values]
        lod_results = {}
        # all depth curves are fitted (and bootstrapped) by ONE device launch; the salts are drawn in depth order,
        # exactly as a per-depth loop of _fit_with_ci would draw them
        ids = [(int(self.rng.integers(0, 2**31 - 1)) << 8) | (di & 0xFF) for di in range(len(depth_values))]
        fits = nv.default_context(int(self.config.seed)).fit_lod_batch(
            list(af_values), [hits[d] for d in depth_values], target_detection_rate, 200, ids)
        for di, depth in enumerate(depth_values):
            lod_af, lo, hi = fits[di]
            lod_results[depth] = {
                "lod_af": float(lod_af), "lod_ci_lower": float(lo), "lod_ci_upper": float(hi),
                "af_values": [float(x) for x in af_values], "hit_rates": [float(h) for h in hits[depth]],
                "target_detection_rate": target_detection_rate, "n_replicates": n_replicates,
            }
            print(f"    LoD at depth {depth}: {lod_af:.2e} AF [{lo:.2e}, {hi:.2e}]")
        self.lod_results = {"depth_results": lod_results, "af_range": af_range, "target_detection_rate": target_detection_rate,
                            "alpha": alpha, "beta": beta, "config_hash": self.config.config_hash()}
        return self.lod_results

    # ------------------------------------------------------------------ LoQ
    def estimate_loq(self, af_range: Tuple[float, float] = (1e-4, 1e-2), depth_values: List[int] = [1000, 5000, 10000],
                     n_replicates: int = 50, cv_threshold: float = 0.20,
                     abs_error_threshold: Optional[float] = None) -> Dict[str, Any]:
        print(f"Estimating LoQ with CV threshold {cv_threshold:.1%}...")
        af_values = np.logspace(np.log10(af_range[0]), np.log10(af_range[1]), 12)      # lod.py:209
        est: Dict[Tuple[int, float], np.ndarray] = {}
        if self.engine == "stages":
            for depth in depth_values:
                for af in af_values:
                    vals = []
                    for rep in range(n_replicates):
                        run_rng = np.random.default_rng(self.config.seed + rep * 2000 + int(af * 1e6))   # lod.py:224
                        calls_df = self._run_pipeline(self._create_lod_config(float(af), int(depth)), run_rng)
                        vals.append(float((calls_df["variant_call"] == True).sum()) / len(calls_df) if len(calls_df) else 0.0)  # noqa: E712
                    est[(depth, float(af))] = np.array(vals)
        else:
            cid, af, dp = self._grid_cells(ANALYSIS_LOQ, af_values, depth_values, n_replicates, max(n_replicates, 25))
            res = self._run_grid(cid, af, dp)
            # read level: the quantity being quantified is the alt-family fraction of the cell
            frac = res["n_variants"] / np.maximum(res["n_total"], 1)
            for depth in depth_values:
                for a in af_values:
                    est[(depth, float(a))] = frac[(dp == depth) & (af == a)]
        loq_results = {}
        for depth in depth_values:
            cvs, errs, tested = [], [], []
            for af in af_values:
                v = est[(depth, float(af))]
                if len(v) > 1:
                    mean_af, std_af = float(np.mean(v)), float(np.std(v))
                    cvs.append(std_af / mean_af if mean_af > 0 else float("inf"))      # lod.py:245-247
                    errs.append(abs(mean_af - float(af)))
                    tested.append(float(af))
            loq_cv = next((tested[i] for i, cv in enumerate(cvs) if cv <= cv_threshold), None)
            loq_abs = None
            if abs_error_threshold is not None:
                loq_abs = next((tested[i] for i, e in enumerate(errs) if e <= abs_error_threshold), None)
            loq_results[depth] = {"loq_af_cv": loq_cv, "loq_af_abs_error": loq_abs, "af_values": tested, "cv_values": cvs,
                                  "abs_errors": errs, "cv_threshold": cv_threshold, "abs_error_threshold": abs_error_threshold,
                                  "n_replicates": n_replicates}
        self.loq_results = {"depth_results": loq_results, "af_range": af_range, "cv_threshold": cv_threshold,
                            "abs_error_threshold": abs_error_threshold, "config_hash": self.config.config_hash()}
        return self.loq_results

    # ------------------------------------------------------------------ curve fit (K11)
    def _fit_with_ci(self, af_values, hit_rates, target_rate, n_bootstrap, stream_id=0):
        salt = int(self.rng.integers(0, 2**31 - 1))     # one draw from the analyzer's generator keys the resampling
        ctx = nv.default_context(int(self.config.seed))
        return ctx.fit_lod(af_values, hit_rates, target_rate, n_bootstrap, stream_id=(salt << 8) | (stream_id & 0xFF))

    def _fit_detection_curve(self, af_values: List[float], hit_rates: List[float], target_rate: float = 0.95) -> float:
        """lod.py:351-383: logistic LM fit on log10(AF); interpolation fallback when curve_fit fails."""
        return nv.default_context(int(self.config.seed)).fit_lod(af_values, hit_rates, target_rate, 0)[0]

    def _bootstrap_lod_ci(self, af_values: List[float], hit_rates: List[float], target_rate: float,
                          n_bootstrap: int = 200) -> Tuple[float, float]:
        """lod.py:385-408."""
        _, lo, hi = self._fit_with_ci(af_values, hit_rates, target_rate, n_bootstrap)
        return lo, hi

    # ------------------------------------------------------------------ reports
    def generate_reports(self, output_dir: str = "reports") -> None:
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        if self.lob_results:
            payload = dict(self.lob_results)
            payload["schema_version"] = "1.0.0"
            with open(out / "lob.json", "w") as f:
                json.dump(payload, f, indent=2)
        if self.lod_results:
            rows = [{"depth": d, "lod_af": r["lod_af"], "lod_ci_lower": r["lod_ci_lower"], "lod_ci_upper": r["lod_ci_upper"],
                     "target_detection_rate": r["target_detection_rate"], "n_replicates": r["n_replicates"]}
                    for d, r in self.lod_results["depth_results"].items()]
            pd.DataFrame(rows).to_csv(out / "lod_table.csv", index=False)
            from ..reporting import write_png_heatmap

            write_png_heatmap([r["hit_rates"] for r in self.lod_results["depth_results"].values()], out / "lod_curves.png")
        if self.loq_results:
            rows = [{"depth": d, "loq_af_cv": r["loq_af_cv"], "loq_af_abs_error": r["loq_af_abs_error"],
                     "cv_threshold": r["cv_threshold"], "abs_error_threshold": r["abs_error_threshold"],
                     "n_replicates": r["n_replicates"]} for d, r in self.loq_results["depth_results"].items()]
            pd.DataFrame(rows).to_csv(out / "loq_table.csv", index=False)


def estimate_lob(config, rng, n_blank_runs: int = 100, engine: str = "stages"):
    return LODAnalyzer(config, rng, engine).estimate_lob(n_blank_runs)


def estimate_lod(config, rng, af_range=(1e-4, 1e-2), depth_values=[1000, 5000, 10000], n_replicates: int = 50, engine: str = "stages"):
    return LODAnalyzer(config, rng, engine).estimate_lod(af_range, depth_values, n_replicates)


def estimate_loq(config, rng, af_range=(1e-4, 1e-2), depth_values=[1000, 5000, 10000], n_replicates: int = 50,
                 cv_threshold: float = 0.20, engine: str = "stages"):
    return LODAnalyzer(config, rng, engine).estimate_loq(af_range, depth_values, n_replicates, cv_threshold)
