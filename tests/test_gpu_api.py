"""GPU tests of the reference-facing Python API (the drop-in boundary of SURVEY.md §8b): same
signatures, same DataFrame columns / dtypes, same JSON artifacts, same error behaviour -- with
every stage executing in the CUDA library."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / "precise-mrd-mini_b200"


@pytest.fixture(scope="module")
def smoke_cfg():
    from precise_mrd.config import load_config

    cfg = load_config(PKG_DIR / "precise_mrd" / "assets" / "configs" / "smoke.yaml")
    cfg.seed = 7
    return cfg


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


@pytest.mark.parametrize("name", ["call_smoke_v1.npz", "call_smoke_v2.npz"])
def test_call_mrd_dataframe_api_on_reference_fixture(ctx, golden, smoke_cfg, name):
    """Feed the reference's own collapsed_umis + error_model tables; require its mrd_calls back."""
    from precise_mrd.call import call_mrd

    g = golden(name)
    seg = g["seg_offsets"]
    sid = np.repeat(g["sample_id"], np.diff(seg))
    collapsed = pd.DataFrame({"sample_id": sid, "is_variant": g["is_variant"].astype(bool), "quality_score": g["quality"],
                              "consensus_agreement": g["agreement"], "allele_fraction": np.repeat(g["allele_fraction"], np.diff(seg))})
    em = pd.DataFrame({"error_rate": g["error_rate"]})
    calls = call_mrd(collapsed, em, smoke_cfg, np.random.default_rng(0))
    assert list(calls.columns) == ["sample_id", "allele_fraction", "n_variants", "n_total", "variant_fraction", "expected_variants",
                                   "fold_change", "p_value", "mean_quality", "mean_consensus", "test_type", "config_hash",
                                   "p_adjusted", "significant", "alpha", "fdr_method", "variant_call"]
    assert calls["n_variants"].dtype == np.int64 and calls["significant"].dtype == bool
    for col in ("n_variants", "n_total", "significant"):                           # bit-exact columns
        assert (calls[col].values == g[f"want_{col}"]).all(), col
    # one row per sample in first-appearance order, carrying the sample's own allele fraction (call.py:161-170)
    assert (calls["sample_id"].values == g["sample_id"]).all()
    assert (calls["allele_fraction"].values == g["allele_fraction"]).all()
    for col in ("variant_fraction", "expected_variants", "fold_change", "p_value", "p_adjusted", "mean_quality", "mean_consensus"):
        assert _rel(calls[col].values, g[f"want_{col}"]).max() < 1e-12, col
    assert (calls["variant_call"] == calls["significant"]).all()
    assert (calls["test_type"] == "poisson").all() and (calls["config_hash"] == smoke_cfg.config_hash()).all()
    # interleaved rows (samples not contiguous) give the same table (call.py:161 first-appearance order)
    perm = np.random.default_rng(1).permutation(len(collapsed))
    keep_first = np.sort(np.unique(sid[perm], return_index=True)[1])
    calls2 = call_mrd(collapsed.iloc[perm].reset_index(drop=True), em, smoke_cfg, np.random.default_rng(0))
    assert list(calls2["sample_id"]) == list(sid[perm][keep_first])
    m = calls2.set_index("sample_id").loc[calls["sample_id"]]
    assert (m["n_variants"].values == calls["n_variants"].values).all() and _rel(m["p_value"].values, calls["p_value"].values).max() < 1e-12


def test_call_mrd_error_behaviour(ctx, smoke_cfg):
    import copy

    from precise_mrd.call import benjamini_hochberg_correction, binomial_test, call_mrd, poisson_test

    assert len(call_mrd(pd.DataFrame(), pd.DataFrame({"error_rate": [1e-3]}), smoke_cfg, np.random.default_rng(0))) == 0   # call.py:214
    bad = copy.deepcopy(smoke_cfg)
    bad.stats.test_type = "fisher"
    df = pd.DataFrame({"sample_id": [0], "is_variant": [True], "quality_score": [30.0], "consensus_agreement": [0.9], "allele_fraction": [0.1]})
    with pytest.raises(ValueError, match="Unknown test type"):
        call_mrd(df, pd.DataFrame({"error_rate": [1e-3]}), bad, np.random.default_rng(0))
    with pytest.raises(KeyError):
        call_mrd(df.drop(columns=["is_variant"]), pd.DataFrame({"error_rate": [1e-3]}), smoke_cfg, np.random.default_rng(0))
    # the reference's KAT table (the values the reference code really returns)
    assert poisson_test(5, 1.0) == pytest.approx(0.007319693654687426, rel=1e-12)
    assert poisson_test(1, 5.0) == pytest.approx(0.08085536398902558, rel=1e-12)
    assert poisson_test(5, 5.0) == 1.0 and poisson_test(0, 0.0) == 1.0
    assert poisson_test(10, 0.1) == pytest.approx(5.032695613540631e-17, rel=1e-12)
    assert binomial_test(10, 100, 0.01) == pytest.approx(7.631587532260653e-08, rel=1e-12)
    assert binomial_test(1, 100, 0.1) == pytest.approx(0.0006336060581826209, rel=1e-12)
    assert [binomial_test(*a) for a in [(10, 100, 0.1), (0, 0, .5), (0, 10, 0.), (1, 10, 0.), (10, 10, 1.), (9, 10, 1.)]] == [1, 1, 1, 0, 1, 0]
    rej, adj = benjamini_hochberg_correction(np.array([0.001, 0.005, 0.01, 0.03, 0.05, 0.1]), 0.05)
    assert list(rej) == [True, True, True, True, False, False] and np.allclose(adj, [0.006, 0.015, 0.02, 0.045, 0.06, 0.1], rtol=1e-15)
    rej, adj = benjamini_hochberg_correction(np.array([]), 0.05)
    assert rej.dtype == bool and len(rej) == 0 and len(adj) == 0


def test_stage_pipeline_tables_and_determinism(ctx, smoke_cfg):
    from precise_mrd import call_mrd, collapse_umis, fit_error_model, simulate_reads

    def run(seed):
        rng = np.random.default_rng(seed)
        reads = simulate_reads(smoke_cfg, rng)
        col = collapse_umis(reads, smoke_cfg, rng)
        em = fit_error_model(col, smoke_cfg, rng)
        return reads, col, em, call_mrd(col, em, smoke_cfg, rng)

    reads, col, em, calls = run(7)
    assert list(reads.columns) == ["sample_id", "allele_fraction", "target_depth", "replicate", "n_families", "n_true_variants",
                                   "n_false_positives", "background_rate", "mean_family_size", "mean_quality", "config_hash"]
    assert len(reads) == 60 and (reads["sample_id"].values == np.arange(60)).all()
    assert (reads["allele_fraction"].values[:20] == 0.01).all() and (reads["target_depth"].values[:10] == 1000).all()   # AF-major meshgrid
    assert list(col.columns) == ["sample_id", "family_id", "family_size", "quality_score", "consensus_agreement", "passes_quality",
                                 "passes_consensus", "is_variant", "allele_fraction"]
    assert col["family_size"].min() >= 3 and col["passes_quality"].dtype == bool
    # Appendix C: kept fraction ~ P[clip(Poisson(max(3, m))) >= 3] averaged over m ~ clip(Poisson(5)): 0.83 in the reference run
    assert 0.78 < len(col) / 180000 < 0.89
    assert list(em.columns) == ["trinucleotide_context", "error_rate", "ci_lower", "ci_upper", "n_observations", "config_hash"]
    assert len(em) == 64 and em["trinucleotide_context"].iloc[0] == "AAA" and em["trinucleotide_context"].iloc[-1] == "TTT"
    assert len(calls) == 60 and calls["significant"].sum() == 0       # Appendix A #12: AF only acts through AF > 0.01
    r2, c2, e2, k2 = run(7)
    pd.testing.assert_frame_equal(col, c2)
    pd.testing.assert_frame_equal(calls, k2)                          # same generator state => identical tables
    _, c3, _, _ = run(8)
    assert len(c3) != len(col) or not (c3["quality_score"].values[:100] == col["quality_score"].values[:100]).all()
    # dict-style sub-sections (tests/test_rng_api.py:151-172)
    class DictCfg:
        seed = 7
        simulation = {"allele_fractions": [0.02], "umi_depths": [500], "n_replicates": 2, "n_bootstrap": 10}
        umi = {"min_family_size": 3, "max_family_size": 1000, "quality_threshold": 20, "consensus_threshold": 0.6}
        stats = {"test_type": "binomial", "alpha": 0.05, "fdr_method": "benjamini_hochberg"}
        def config_hash(self):
            return "dictcfg"
    rng = np.random.default_rng(1)
    d = DictCfg()
    reads = simulate_reads(d, rng)
    col = collapse_umis(reads, d, rng)
    calls = call_mrd(col, fit_error_model(col, d, rng), d, rng)
    assert len(calls) == 2 and (calls["test_type"] == "binomial").all()
    assert 0.70 < col["is_variant"].mean() < 0.88     # AF 0.02 > 0.01: P[U < 0.8] x pass rates (collapse.py:104)


def test_read_level_dataframe_apis_vs_reference_golden(ctx, golden):
    from precise_mrd.umi import UMIProcessor, collapse_reads, unpack_bases

    g = golden("ga_reads.npz")
    keys, pl = g["keys"], g["payload"]
    grp = (keys >> np.uint64(32)).astype(np.int64)
    reads = pd.DataFrame({"patient_id": "P001", "chrom": "chr1", "start": grp, "pos": g["group_locus"][grp],
                          "umi": [unpack_bases(int(k) & 0xFFFF, 8) for k in keys],
                          "sequence": [unpack_bases(int(p) & 0xFFFFF, 10) for p in pl], "true_alt": (pl >> 31).astype(bool)})
    fam, loc = collapse_reads(reads, min_family_size=2, max_distance=1)
    assert len(fam) == len(g["want_fam_key"]) and len(loc) == len(g["want_locus"])
    assert (fam["family_size"].values == g["want_fam_size"]).all() and (fam["alt_count"].values == g["want_fam_alt"]).all()
    assert [unpack_bases(int(c), 10) for c in g["want_fam_consensus"]] == list(fam["consensus"])
    assert [unpack_bases(int(k) & 0xFFFF, 8) for k in g["want_fam_key"]] == list(fam["umi_representative"])
    order = {l: i for i, l in enumerate(g["want_locus"])}
    idx = [order[p] for p in loc["pos"]]
    assert (loc["alt_count"].values == g["want_locus_alt"][idx]).all() and (loc["total_count"].values == g["want_locus_total"][idx]).all()
    assert (loc["family_count"].values == g["want_locus_families"][idx]).all()

    g = golden("gb_reads.npz")
    keys, pl = g["keys"], g["payload"]
    site = (keys >> np.uint64(32)).astype(np.int64)
    reads = pd.DataFrame({"chrom": "chr" + pd.Series(site).astype(str), "pos": 1000 + site, "ref": "A",
                          "alt": ["ACGT"[int(p) & 3] for p in pl], "umi": [unpack_bases(int(k) & 0xFFFFFF, 12) for k in keys],
                          "quality": ((pl >> 2) & 63).astype(int), "strand": np.where((pl >> 8) & 1, "+", "-"),
                          "read_position": ((pl >> 9) & 255).astype(int)})
    proc = UMIProcessor(min_family_size=3, max_family_size=1000, max_edit_distance=0, quality_threshold=20, consensus_threshold=0.6)
    fams = proc.process_reads(reads)
    assert len(fams) == int(g["want_kept"])
    counts = proc.get_consensus_counts(fams)
    cc = np.zeros_like(g["want_consensus_count"])
    for _, r in counts.iterrows():
        cc[int(r.chrom[3:]), "ACGT".index(r.allele)] = r.consensus_count
    assert (cc == g["want_consensus_count"]).all()
    want_q = {int(k): q for k, q in zip(g["want_fam_key"], g["want_fam_quality"])}
    for f in fams[:500]:
        k = (int(f.site_key.split(":")[0][3:]) << 32) | int(sum("ACGT".index(b) << (2 * (11 - i)) for i, b in enumerate(f.umi)))
        if f.consensus_quality is not None:
            assert f.consensus_quality == want_q[k]
    # families come back in the reference's order: sites by first appearance, UMIs by insertion order (umi.py:97-115)
    def fkey(f):
        return (int(f.site_key.split(":")[0][3:]) << 32) | int(sum("ACGT".index(b) << (2 * (11 - i)) for i, b in enumerate(f.umi)))
    assert [fkey(f) for f in fams] == [int(k) for k in g["want_fam_key"][g["want_fam_size"] >= 3]]
    assert [f.family_size for f in fams] == [int(v) for v in g["want_fam_size"][g["want_fam_size"] >= 3]]
    # filter_family_size caps an oversize family at max_family_size reads (umi.py:190-199)
    capped = UMIProcessor(min_family_size=3, max_family_size=6, max_edit_distance=0).process_reads(reads)
    assert len(capped) == len(fams) and max(f.family_size for f in capped) == 6 < max(f.family_size for f in fams)
    assert proc.edit_distance("ACGTACGTACGT", "ACGTACGTACGA") == 1 and proc.edit_distance("ACG", "ACGT") == 4

    # the reference's DEFAULT: max_edit_distance = 1 (configs/default.yaml:16) -- greedy Hamming clusters, listed as
    # umi_clustering.py:59-101 seeds them; reference-generated vectors (oracle/make_golden.py gbgreedy)
    g = golden("gb_greedy.npz")
    keys, pl = g["keys"], g["payload"]
    site = (keys >> np.uint64(32)).astype(np.int64)
    reads = pd.DataFrame({"chrom": "chr" + pd.Series(site).astype(str), "pos": 1000 + site, "ref": "A",
                          "alt": ["ACGT"[int(p) & 3] for p in pl], "umi": [unpack_bases(int(k) & 0xFFFFFF, 12) for k in keys],
                          "quality": ((pl >> 2) & 63).astype(int), "strand": np.where((pl >> 8) & 1, "+", "-"),
                          "read_position": ((pl >> 9) & 255).astype(int)})
    default = UMIProcessor()
    assert default.max_edit_distance == 1
    fams = default.process_reads(reads)
    keep = g["want_fam_size_d1"] >= 3
    assert [fkey(f) for f in fams] == [int(k) for k in g["want_fam_key_d1"][keep]]
    assert [f.family_size for f in fams] == [int(v) for v in g["want_fam_size_d1"][keep]]
    want_allele = g["want_fam_allele_d1"][keep]
    assert [(-1 if f.consensus_allele is None else "ACGT".index(f.consensus_allele)) for f in fams] == [int(a) for a in want_allele]
    wq = g["want_fam_quality_d1"][keep]
    assert all(f.consensus_quality == q for f, q in zip(fams, wq) if f.consensus_allele is not None)
    with pytest.raises(ValueError):
        UMIProcessor(min_family_size=0)


def test_contamination_stages_engine_equals_the_reference_run(ctx, smoke_cfg):
    """C4 parity: the stages engine against the reference's OWN ContaminationSimulator run (oracle/make_golden.py
    contam -> tests/golden/contam_ref.json; src/precise_mrd/sim/contamination.py:85-228 with `variant_call` added).
    On the reference grid (AF <= 0.01) no variant is ever generated, so every sensitivity and
    false-positive rate of the reference is exactly 0 -- and must be here: same keys, same numbers."""
    import json
    from pathlib import Path

    from precise_mrd.sim import ContaminationSimulator

    ref = json.loads((Path(__file__).parent / "golden" / "contam_ref.json").read_text())
    assert ref["seed"] == smoke_cfg.seed
    got = ContaminationSimulator(smoke_cfg, np.random.default_rng(smoke_cfg.seed), engine="stages").simulate_contamination_effects(**ref["args"])
    n = 0
    for kind in ("index_hopping", "barcode_collisions", "cross_sample_contamination"):
        assert [str(r) for r in got[kind]] == list(ref["results"][kind])
        for rate, by_af in got[kind].items():
            for af, by_depth in by_af.items():
                for depth, cell in by_depth.items():
                    want = ref["results"][kind][str(rate)][str(af)][str(depth)]
                    assert set(cell) == set(want), (kind, rate, af, depth)
                    for k, v in want.items():
                        assert cell[k] == v, (kind, rate, af, depth, k, cell[k], v)
                    n += 1
    assert n == 12
    m, want_m = got["sensitivity_matrix"]["index_hopping"], ref["results"]["sensitivity_matrix"]["index_hopping"]
    assert m["matrix"] == want_m["matrix"] and m["condition_labels"] == want_m["condition_labels"]


def test_contamination_grid_engine_is_one_plan(ctx, smoke_cfg):
    """engine="grid": the whole stress sweep runs as ONE device plan (per-cell knobs, hop pools, call groups).
    Detection falls with the hop rate's dilution only mildly, rises with AF and depth; the result dict has the
    reference's shape; the launch count is that of one plan, not one per condition."""
    from precise_mrd import _native as nv
    from precise_mrd.sim import ContaminationSimulator

    c = nv.default_context(int(smoke_cfg.seed))
    sim = ContaminationSimulator(smoke_cfg, np.random.default_rng(smoke_cfg.seed), engine="grid")
    l0 = c.launch_count
    res = sim.simulate_contamination_effects(n_replicates=6)
    launches = c.launch_count - l0
    assert set(res) == {"index_hopping", "barcode_collisions", "cross_sample_contamination", "sensitivity_matrix"}
    assert list(res["index_hopping"]) == [0.0, 0.001, 0.002, 0.005, 0.01] and list(res["cross_sample_contamination"]) == [0.0, 0.01, 0.05, 0.1]
    n_groups = (5 + 4 + 4) * 3 * 2
    assert len(sim.grid_table["group_error_rate"]) == n_groups and len(sim.grid_table["p_value"]) == n_groups * (6 + 10)
    assert launches < 6 * n_groups, launches                     # one plan per condition used to cost ~40 launches each
    hop = res["index_hopping"]
    assert hop[0.0][0.01][5000]["mean_sensitivity"] >= hop[0.0][0.001][1000]["mean_sensitivity"]
    # a 10x-AF contaminant raises the alt-family count: detection at AF 1e-3 / depth 5000 does not fall with the proportion
    cross = res["cross_sample_contamination"]
    assert cross[0.1][0.001][5000]["mean_sensitivity"] >= cross[0.0][0.001][5000]["mean_sensitivity"]
    assert all("mean_fp_rate" in v for r in res["barcode_collisions"].values() for a in r.values() for v in a.values())


def test_lod_analyzer_both_engines(ctx, smoke_cfg, tmp_path):
    from precise_mrd.eval import LODAnalyzer
    from precise_mrd.validation import validate_json

    an = LODAnalyzer(smoke_cfg, np.random.default_rng(7), engine="stages")
    lob = an.estimate_lob(n_blank_runs=6)
    assert lob["lob_value"] == 0.0 and lob["blank_measurements"] == [0] * 6 and lob["percentile"] == 95
    lod = an.estimate_lod(depth_values=[1000], n_replicates=2)
    r = lod["depth_results"][1000]
    # the parity target: on the reference's own grid no variant is ever generated (AF <= 0.01), every
    # hit rate is 0, curve_fit fails, the interpolation fallback is NaN
    assert r["hit_rates"] == [0.0] * 15 and np.isnan(r["lod_af"]) and np.isnan(r["lod_ci_lower"])
    loq = an.estimate_loq(depth_values=[1000], n_replicates=2)
    assert loq["depth_results"][1000]["loq_af_cv"] is None
    an.generate_reports(str(tmp_path))
    validate_json(tmp_path / "lob.json", "lob.schema.json")
    assert list(pd.read_csv(tmp_path / "lod_table.csv").columns)[:4] == ["depth", "lod_af", "lod_ci_lower", "lod_ci_upper"]
    assert list(pd.read_csv(tmp_path / "loq_table.csv").columns)[:3] == ["depth", "loq_af_cv", "loq_af_abs_error"]
    # curve-fit helpers on supplied vectors
    af = np.logspace(-4, -2, 15)
    hits = [0, 0, 0, 0, 0, .04, 0, .08, .2, .4, .6, .8, .92, 1, 1]
    assert an._fit_detection_curve(list(af), hits, 0.95) == pytest.approx(0.006233694357173142, rel=1e-6)
    lo, hi = an._bootstrap_lod_ci(list(af), hits, 0.95, 50)
    assert 0.003 < lo < 0.0063 < hi < 0.02
    # read-level engine: detection rises with AF and depth, LoD is finite and inside the AF range
    g = LODAnalyzer(smoke_cfg, np.random.default_rng(7), engine="grid")
    lod = g.estimate_lod(af_range=(1e-3, 1e-1), depth_values=[2000, 20000], n_replicates=12)
    h_lo, h_hi = lod["depth_results"][2000]["hit_rates"], lod["depth_results"][20000]["hit_rates"]
    assert h_hi[-1] == 1.0 and np.mean(h_hi) >= np.mean(h_lo) and np.mean(h_hi[-5:]) > np.mean(h_hi[:5])
    assert 1e-3 < lod["depth_results"][20000]["lod_af"] < 1e-1
    lob = g.estimate_lob(n_blank_runs=40)
    assert 0 <= lob["blank_mean"] <= 0.5


def test_contamination_simulator_stages_engine(ctx, smoke_cfg, tmp_path):
    from precise_mrd.sim import ContaminationSimulator
    from precise_mrd.validation import validate_json

    sim = ContaminationSimulator(smoke_cfg, np.random.default_rng(7))
    res = sim.simulate_contamination_effects(hop_rates=[0.0, 0.01], barcode_collision_rates=[0.0, 0.001],
                                             cross_sample_proportions=[0.0, 0.1], af_test_values=[0.005], depth_values=[1000], n_replicates=2)
    cell = res["index_hopping"][0.01][0.005][1000]
    assert set(cell) == {"mean_sensitivity", "std_sensitivity", "sensitivity_scores", "n_replicates"}
    assert set(res["barcode_collisions"][0.001][0.005][1000]) == {"mean_sensitivity", "std_sensitivity", "mean_fp_rate", "std_fp_rate", "n_replicates"}
    assert np.array(res["sensitivity_matrix"]["index_hopping"]["matrix"]).shape == (2, 1)
    sim.generate_contamination_reports(str(tmp_path))
    validate_json(tmp_path / "contam_sensitivity.json", "contamination.schema.json")
    assert (tmp_path / "contam_heatmap.png").stat().st_size > 0
    g = ContaminationSimulator(smoke_cfg, np.random.default_rng(7), engine="grid")
    res = g.simulate_contamination_effects(hop_rates=[0.0, 0.05], barcode_collision_rates=[0.0], cross_sample_proportions=[0.0],
                                           af_test_values=[0.05], depth_values=[5000], n_replicates=6)
    assert res["index_hopping"][0.0][0.05][5000]["mean_sensitivity"] > 0


def test_cli_smoke_and_determinism_subprocess(tmp_path):
    """tests/test_artifact_contract.py:88-127: run the CLI, validate artifacts, two runs => identical manifest."""
    env = dict(os.environ, PYTHONPATH=f"{PKG_DIR}:{os.environ.get('PYTHONPATH', '')}")
    cmd = [sys.executable, "-m", "precise_mrd.cli", "--seed", "13", "smoke", "--out-dir", "data/smoke"]
    out = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    summary = json.loads(out.stdout[out.stdout.index("{"):])
    assert summary["stage"] == "smoke" and summary["seed"] == 13
    from precise_mrd.validation import validate_artifacts

    validate_artifacts(tmp_path / "reports")
    for f in ("simulated_reads.parquet", "collapsed_umis.parquet", "error_model.parquet", "mrd_calls.parquet"):
        assert (tmp_path / "data" / "smoke" / f).exists()
    calls = pd.read_parquet(tmp_path / "data" / "smoke" / "mrd_calls.parquet")
    assert "variant_call" in calls.columns and len(calls) == 60
    m1 = (tmp_path / "reports" / "hash_manifest.txt").read_text()
    out = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and (tmp_path / "reports" / "hash_manifest.txt").read_text() == m1
    out = subprocess.run([sys.executable, "-m", "precise_mrd.cli", "--seed", "13", "determinism"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "hashes-identical" in out.stdout, out.stderr[-2000:]
    out = subprocess.run([sys.executable, "-m", "precise_mrd.cli", "--seed", "14", "smoke"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    a, b = m1.splitlines()[0].split()[0], (tmp_path / "reports" / "hash_manifest.txt").read_text().splitlines()[0].split()[0]
    assert a != b                                                                   # different seed => different metrics.json
    out = subprocess.run([sys.executable, "-m", "precise_mrd.cli", "--seed", "7", "eval-lob", "--n-blank", "5"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    validate_artifacts(tmp_path / "reports")
    assert json.loads((tmp_path / "reports" / "lob.json").read_text())["n_blank_runs"] == 5
    # `precise-mrd performance` (cli.py:408-438): the report shape of the reference's PerformanceMonitor, filled with the
    # per-kernel CUDA-event times of one smoke run
    out = subprocess.run([sys.executable, "-m", "precise_mrd.cli", "--seed", "7", "performance"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Performance Report:" in out.stdout and "Timing Statistics:" in out.stdout
    rep = json.loads(out.stdout[out.stdout.index("{"):])
    assert set(rep) >= {"timing_statistics", "peak_memory_mb", "total_functions_tracked"}
    ts = rep["timing_statistics"]
    assert rep["total_functions_tracked"] == len(ts) >= 5 and rep["peak_memory_mb"] > 0
    for name, st in ts.items():
        assert set(st) == {"calls", "total_time", "avg_time", "min_time", "max_time"}
        assert st["calls"] >= 1 and 0 <= st["min_time"] <= st["avg_time"] <= st["max_time"] <= st["total_time"] + 1e-12, name
    assert any(k.startswith("gd_") for k in ts) and any(k.startswith("pvalues_kernel") for k in ts)
