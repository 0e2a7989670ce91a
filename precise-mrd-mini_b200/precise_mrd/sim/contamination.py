"""Contamination stress test -- drop-in for ``src/precise_mrd/sim/contamination.py``
(``ContaminationSimulator``, ``run_contamination_stress_test``).

* ``engine="stages"`` (parity mode): the reference's sample-table perturbations
  (``contamination.py:230-311``) followed by the stage pipeline per (rate, AF, depth, rep), with
  the reference's seed arithmetic; every stage is a kernel call.  Because each inner run holds ONE
  row, ``Binom(1, rate)`` is almost always 0 and the perturbation is a near no-op -- exactly as
  in the reference.
* ``engine="grid"``: the read-level meaning ``north_star`` gives these knobs -- index hopping
  re-keys a Binomial(n_reads, hop) subset of READS to another sample's index, barcode collisions
  make a Binomial(n_families, rate) subset of FAMILIES share another molecule's UMI, cross-sample
  contamination mixes ``int(n_families * proportion)`` molecules of a 10x-AF sample (their own UMIs
  and reads) into the cell (``contamination.py:280-311``) -- and the WHOLE sweep (1 560 cells + their
  blanks: every kind x rate x AF x depth x replicate) runs as ONE device plan with per-cell knobs
  (``pmrd_plan_create_ex``): a (kind, rate, AF, depth) condition is a *call group* (own negative
  controls, own error rate -- what one ``call_mrd`` of the reference sees) and, for hopping, a *hop
  pool* (the samples that share a lane)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd

from .. import _native as nv
from ..call import call_mrd
from ..collapse import collapse_umis
from ..config import PipelineConfig
from ..error_model import fit_error_model
from ..simulate import simulate_reads


class ContaminationSimulator:
    def __init__(self, config: PipelineConfig, rng: np.random.Generator, engine: str = "stages"):
        if engine not in ("stages", "grid"):
            raise ValueError("engine must be 'stages' or 'grid'")
        self.config, self.rng, self.engine = config, rng, engine
        self.contamination_results: Optional[Dict[str, Any]] = None

    # ------------------------------------------------------------------ public
    def simulate_contamination_effects(self, hop_rates: List[float] = [0.0, 0.001, 0.002, 0.005, 0.01],
                                       barcode_collision_rates: List[float] = [0.0, 0.0001, 0.0005, 0.001],
                                       cross_sample_proportions: List[float] = [0.0, 0.01, 0.05, 0.1],
                                       af_test_values: List[float] = [0.001, 0.005, 0.01],
                                       depth_values: List[int] = [1000, 5000], n_replicates: int = 20) -> Dict[str, Any]:
        print("Running contamination stress testing...")
        results: Dict[str, Any] = {}
        self._grid_det = None
        if self.engine == "grid":
            self._grid_det = self._grid_sweep_all({"hop": hop_rates, "collision": barcode_collision_rates, "cross": cross_sample_proportions},
                                                  af_test_values, depth_values, n_replicates)
        results["index_hopping"] = self._sweep("hop", hop_rates, af_test_values, depth_values, n_replicates, 1000)
        results["barcode_collisions"] = self._sweep("collision", barcode_collision_rates, af_test_values, depth_values, n_replicates, 2000)
        results["cross_sample_contamination"] = self._sweep("cross", cross_sample_proportions, af_test_values, depth_values, n_replicates, 3000)
        results["sensitivity_matrix"] = self._create_sensitivity_matrix(results, af_test_values, depth_values)
        self.contamination_results = results
        return results

    # ------------------------------------------------------------------ sweeps
    def _sweep(self, kind: str, rates, af_values, depth_values, n_rep: int, rep_mult: int) -> Dict[Any, Any]:
        out: Dict[Any, Any] = {}
        for ri, rate in enumerate(rates):
            out[rate] = {}
            for ai, af in enumerate(af_values):
                out[rate][af] = {}
                for di, depth in enumerate(depth_values):
                    if self.engine == "stages":
                        n_det = [self._one_run(kind, rate, af, depth, rep, rep_mult) for rep in range(n_rep)]
                    else:
                        n_det = self._grid_runs(kind, rate, af, depth, n_rep, (ri, ai, di))
                    expected = max(1, int(af * depth * 0.8))                          # contamination.py:119
                    sens = [min(1.0, n / expected) for n in n_det]
                    cell = {"mean_sensitivity": float(np.mean(sens)), "std_sensitivity": float(np.std(sens)), "n_replicates": n_rep}
                    if kind == "hop":
                        cell["sensitivity_scores"] = sens
                    if kind == "collision":
                        fp = [max(0, n - expected) / depth for n in n_det]            # contamination.py:168-170
                        cell["mean_fp_rate"], cell["std_fp_rate"] = float(np.mean(fp)), float(np.std(fp))
                    out[rate][af][depth] = cell
        return out

    def _one_run(self, kind: str, rate: float, af: float, depth: int, rep: int, rep_mult: int) -> int:
        run_rng = np.random.default_rng(self.config.seed + rep * rep_mult + int(rate * 1e6))   # contamination.py:101,148,201
        cfg = self._create_contamination_config(af, depth)
        if kind == "hop":
            reads_df = self._simulate_with_index_hopping(cfg, run_rng, rate)
        elif kind == "collision":
            reads_df = self._simulate_with_barcode_collisions(cfg, run_rng, rate)
        else:
            reads_df = self._simulate_with_cross_contamination(cfg, run_rng, rate)
        collapsed_df = collapse_umis(reads_df, cfg, run_rng)
        error_model = fit_error_model(collapsed_df, cfg, run_rng)
        calls_df = call_mrd(collapsed_df, error_model, cfg, run_rng)
        return int((calls_df["variant_call"] == True).sum())                            # noqa: E712

    def _grid_runs(self, kind: str, rate: float, af: float, depth: int, n_rep: int, index) -> List[int]:
        return self._grid_det[(kind,) + tuple(index)]

    def _grid_sweep_all(self, rates_by_kind, af_values, depth_values, n_rep: int):
        """The whole stress sweep as ONE plan: returns {(kind, rate_i, af_i, depth_i): [detected per replicate]}."""
        from .. import grid

        ctx = nv.default_context(int(self.config.seed))
        n_blank = max(n_rep, 10)
        m = n_rep + n_blank
        cid, afs, nfam, hop, pool0, pooln, coll, crossp, crossaf, gstart, keys = [], [], [], [], [], [], [], [], [], [0], []
        # hopping conditions first (those cells take the general path; a chunk never mixes the two kinds of cells)
        for kind in ("hop", "collision", "cross"):
            kid = {"hop": 1, "collision": 2, "cross": 3}[kind]
            for ri, rate in enumerate(rates_by_kind[kind]):
                for ai, af in enumerate(af_values):
                    for di, depth in enumerate(depth_values):
                        # cell ids from INDICES (sweep, rate_i, af_i, depth_i, rep), never from seeds
                        base = ((((kid * 64 + ri) * 64 + ai) * 64 + di) << 20) + (4 << 50)
                        start = gstart[-1]
                        cid.append(base + np.arange(m, dtype=np.int64))
                        afs.append(np.concatenate([np.full(n_rep, af), np.zeros(n_blank)]))
                        nfam.append(np.full(m, depth, dtype=np.int64))
                        hop.append(np.full(m, rate if kind == "hop" else 0.0))
                        pool0.append(np.full(m, start, dtype=np.int64))
                        pooln.append(np.full(m, m, dtype=np.int64))
                        coll.append(np.full(m, rate if kind == "collision" else 0.0))
                        # contamination.py:291-299: the contaminating sample carries min(0.1, 10 AF); blanks stay clean
                        crossp.append(np.concatenate([np.full(n_rep, rate if kind == "cross" else 0.0), np.zeros(n_blank)]))
                        crossaf.append(np.full(m, min(0.1, af * 10)))
                        gstart.append(start + m)
                        keys.append((kind, ri, ai, di))
        knobs = {"hop_rate": np.concatenate(hop), "hop_pool_start": np.concatenate(pool0), "hop_pool_size": np.concatenate(pooln),
                 "collision_rate": np.concatenate(coll), "cross_proportion": np.concatenate(crossp), "cross_af": np.concatenate(crossaf),
                 "call_group_start": np.array(gstart, dtype=np.int64)}
        res = grid.run_cells(self.config, np.concatenate(cid), np.concatenate(afs), np.concatenate(nfam), ctx=ctx,
                             sim=grid.read_sim_params(self.config), neg_af_max=0.0, alternative=nv.ALT_GREATER, knobs=knobs)
        self.grid_table = res
        det = res["p_value"] <= float(self.config.stats.alpha)
        return {k: [int(v) for v in det[gstart[g]: gstart[g] + n_rep]] for g, k in enumerate(keys)}

    # ------------------------------------------------------------------ reference perturbations (stages engine)
    def _simulate_with_index_hopping(self, config, rng, hop_rate: float) -> pd.DataFrame:
        reads_df = simulate_reads(config, rng)
        if hop_rate > 0:
            n_hopped = rng.binomial(len(reads_df), hop_rate)                           # contamination.py:240
            if n_hopped > 0:
                contam = reads_df.sample(n=n_hopped, random_state=rng).copy()
                contam["background_rate"] *= 2.0
                contam["n_false_positives"] = contam["n_false_positives"] * 1.5
                contam["sample_id"] = "hopped_" + contam["sample_id"].astype(str)
                reads_df = pd.concat([reads_df.astype({"sample_id": object}), contam], ignore_index=True)
        return reads_df

    def _simulate_with_barcode_collisions(self, config, rng, collision_rate: float) -> pd.DataFrame:
        reads_df = simulate_reads(config, rng)
        if collision_rate > 0:
            n_coll = rng.binomial(len(reads_df), collision_rate)                       # contamination.py:265
            if n_coll > 0:
                idx = rng.choice(len(reads_df), size=n_coll, replace=False)
                reads_df = reads_df.astype({"n_false_positives": float})
                reads_df.loc[idx, "n_false_positives"] *= 2.0
                reads_df.loc[idx, "background_rate"] *= 1.5
                reads_df.loc[idx, "mean_quality"] *= 0.8
        return reads_df

    def _simulate_with_cross_contamination(self, config, rng, proportion: float) -> pd.DataFrame:
        reads_df = simulate_reads(config, rng)
        if proportion > 0:
            n_contam = int(len(reads_df) * proportion)                                 # contamination.py:288
            if n_contam > 0:
                contam_af = min(0.1, reads_df["allele_fraction"].iloc[0] * 10)
                contam_cfg = self._create_contamination_config(contam_af, int(reads_df["target_depth"].iloc[0]))
                contam = simulate_reads(contam_cfg, rng)
                if len(contam) > 0:
                    sub = contam.sample(n=min(n_contam, len(contam)), random_state=rng).copy()
                    sub["sample_id"] = "contam_" + sub["sample_id"].astype(str)
                    reads_df = pd.concat([reads_df.astype({"sample_id": object}), sub], ignore_index=True)
        return reads_df

    def _create_contamination_config(self, af: float, depth: int) -> PipelineConfig:   # contamination.py:313-327
        sim = self.config.simulation
        return PipelineConfig(run_id=f"{self.config.run_id}_contam", seed=self.config.seed,
                              simulation=type(sim)(allele_fractions=[af], umi_depths=[depth], n_replicates=1,
                                                   n_bootstrap=sim.n_bootstrap),
                              umi=self.config.umi, stats=self.config.stats, lod=self.config.lod)

    # ------------------------------------------------------------------ reports
    def _create_sensitivity_matrix(self, results, af_values, depth_values) -> Dict[str, Any]:
        hop = results.get("index_hopping", {})
        rates = list(hop.keys())
        m = np.zeros((len(rates), len(af_values) * len(depth_values)))
        for i, r in enumerate(rates):
            col = 0
            for af in af_values:
                for depth in depth_values:
                    m[i, col] = hop[r][af][depth]["mean_sensitivity"]
                    col += 1
        return {"index_hopping": {"matrix": m.tolist(), "hop_rates": rates,
                                  "condition_labels": [f"AF={af:.0e}, D={d}" for af in af_values for d in depth_values],
                                  "af_values": list(af_values), "depth_values": list(depth_values)}}

    def generate_contamination_reports(self, output_dir: str = "reports") -> None:
        if not self.contamination_results:
            print("No contamination results to report")
            return
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        payload = dict(self.contamination_results)
        payload["schema_version"] = "1.0.0"
        with open(out / "contam_sensitivity.json", "w") as f:
            json.dump(payload, f, indent=2)
        from ..reporting import write_png_heatmap

        write_png_heatmap(self.contamination_results["sensitivity_matrix"]["index_hopping"]["matrix"], out / "contam_heatmap.png")


def run_contamination_stress_test(config: PipelineConfig, rng: np.random.Generator, output_dir: str = "reports",
                                  engine: str = "stages") -> Dict[str, Any]:
    sim = ContaminationSimulator(config, rng, engine)
    results = sim.simulate_contamination_effects()
    sim.generate_contamination_reports(output_dir)
    return results
