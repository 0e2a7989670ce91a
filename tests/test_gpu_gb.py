"""GPU tests of the G-B surface and the stratified / plugin rows of SURVEY.md §8a (a11, a12, a15)
and §8b (``VariantCaller`` seam): every tail, BH and draw runs in the CUDA library; expectations
are vectors produced by RUNNING the reference (``tests/golden/gb_stats.json``) or, for the
Monte-Carlo generators, its distributions."""
import json
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
from scipy import stats as sps

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / "precise-mrd-mini_b200"
GOLD = json.loads((ROOT / "tests" / "golden" / "gb_stats.json").read_text())
ALPHA = 1e-3


@pytest.fixture(scope="module")
def smoke_cfg():
    from precise_mrd.config import load_config

    cfg = load_config(PKG_DIR / "precise_mrd" / "assets" / "configs" / "smoke.yaml")
    cfg.seed = 7
    return cfg


def _close(got, want, poisson_upper, n=0):
    # stats.py:66,74 form the Poisson upper tail as 1 - cdf(k-1) (absolute error ~1e-16); the device
    # evaluates the survival function -- see tests/test_gb_host.py.  Beyond n = 2e4 SciPy's incomplete
    # beta is itself only good to ~1e-11 (mpmath check in tests/test_gpu_parity.py): 2e-11 there.
    return got == pytest.approx(want, rel=1e-12 if n <= 20000 else 2e-11, abs=1e-15 if poisson_upper else 1e-300)


def test_statistical_tester_scalar_tests_match_reference(ctx):
    from precise_mrd.stats import StatisticalTester, TestResult

    for r in GOLD["scalar"]:
        t = StatisticalTester(test_type=r["test"])
        if r["test"] == "poisson":
            res = t.poisson_test(r["k"], r["rate"] * r["n"], alternative=r["alternative"])
        else:
            res = t.binomial_test(r["k"], r["n"], r["rate"], alternative=r["alternative"])
        assert isinstance(res, TestResult) and res.method == r["test"] and res.alternative == r["alternative"]
        assert _close(res.pvalue, r["pvalue"], r["test"] == "poisson" and r["alternative"] != "less", r["n"]), r
        assert res.statistic == pytest.approx(r["statistic"], rel=1e-15)
        if r["effect_size"] is None:
            assert res.effect_size is None
        else:
            assert res.effect_size == pytest.approx(r["effect_size"], rel=1e-12, abs=1e-12)
        if r["ci"] is not None:     # exact Clopper-Pearson bounds: scipy solves them with brentq (xtol 2e-12)
            assert res.confidence_interval[0] == pytest.approx(r["ci"][0], rel=1e-8, abs=1e-11)
            assert res.confidence_interval[1] == pytest.approx(r["ci"][1], rel=1e-8, abs=1e-11)
    with pytest.raises(ValueError, match="test_type must be"):
        StatisticalTester(test_type="t-test")
    with pytest.raises(ValueError, match="alternative must be"):
        StatisticalTester().poisson_test(3, 1.0, alternative="sideways")
    assert StatisticalTester().binomial_test(0, 0, 0.1).pvalue == 1.0


@pytest.mark.parametrize("test_type", ["poisson", "binomial"])
def test_test_multiple_variants_matches_reference(ctx, test_type):
    from precise_mrd.stats import StatisticalTester

    t = GOLD["table"]
    table = pd.DataFrame({"site_key": t["site_key"], "context": t["context"], "alt_count": t["alt_count"], "total_depth": t["total_depth"]})
    want = GOLD[f"multi_{test_type}"]
    got = StatisticalTester(test_type=test_type, alpha=0.05).test_multiple_variants(table, t["rates"])
    assert list(got.columns) == list(want.keys())
    for col in ("site_key", "context", "method"):
        assert got[col].tolist() == want[col]
    for col in ("observed_alt", "total_depth", "significant", "significant_uncorrected"):
        assert (got[col].values.astype(float) == np.array(want[col])).all(), col
    np.testing.assert_allclose(got["expected_rate"], want["expected_rate"], rtol=0)
    np.testing.assert_allclose(got["test_statistic"], want["test_statistic"], rtol=1e-15)
    np.testing.assert_allclose(got["effect_size"], want["effect_size"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got["pvalue"], want["pvalue"], rtol=2e-11, atol=1e-15 if test_type == "poisson" else 0)
    np.testing.assert_allclose(got["qvalue"], want["qvalue"], rtol=2e-11, atol=4e-13 if test_type == "poisson" else 0)
    # validation errors of stats.py:205-226
    tester = StatisticalTester()
    assert tester.test_multiple_variants(pd.DataFrame(), {}).empty
    with pytest.raises(ValueError, match="Missing required columns"):
        tester.test_multiple_variants(pd.DataFrame({"alt_count": [1]}), {})
    with pytest.raises(ValueError, match="cannot exceed"):
        tester.test_multiple_variants(pd.DataFrame({"alt_count": [5], "total_depth": [3]}), {})
    with pytest.raises(ValueError, match="must be positive"):
        tester.test_multiple_variants(pd.DataFrame({"alt_count": [0], "total_depth": [0]}), {})


def test_multiple_testing_correction_methods(ctx):
    from precise_mrd.stats import StatisticalTester

    p = GOLD["mtc"]["p"]
    for method in ("benjamini_hochberg", "bonferroni", "holm"):
        rej, corr = StatisticalTester(alpha=0.05).multiple_testing_correction(p, method)
        assert rej.tolist() == GOLD["mtc"][method]["rejected"], method
        np.testing.assert_allclose(corr, GOLD["mtc"][method]["corrected"], rtol=4e-16)   # p*m/k vs p/(k/m): 1 ulp
    rej, corr = StatisticalTester().multiple_testing_correction([])
    assert len(rej) == 0 and len(corr) == 0


def test_variant_caller_plugin_seam(ctx, golden, smoke_cfg):
    """§8b: VariantCaller.__init__(config) / train -> {} / predict; get_variant_caller returns the B200 caller."""
    from precise_mrd.call import call_mrd, get_variant_caller
    from precise_mrd.models import B200StatisticalVariantCaller, StatisticalVariantCaller, VariantCaller

    g = golden("call_smoke_v2.npz")
    seg = g["seg_offsets"]
    collapsed = pd.DataFrame({"sample_id": np.repeat(g["sample_id"], np.diff(seg)), "is_variant": g["is_variant"].astype(bool),
                              "quality_score": g["quality"], "consensus_agreement": g["agreement"],
                              "allele_fraction": np.repeat(g["allele_fraction"], np.diff(seg))})
    em = pd.DataFrame({"error_rate": g["error_rate"]})
    caller = get_variant_caller(smoke_cfg, False, "ensemble", False, "cnn_lstm")
    assert isinstance(caller, B200StatisticalVariantCaller) and isinstance(caller, VariantCaller)
    assert StatisticalVariantCaller is B200StatisticalVariantCaller
    assert caller.train(collapsed, np.random.default_rng(0)) == {} and caller.model is None
    pred = caller.predict(collapsed, em)
    assert pred.equals(call_mrd(collapsed, em, smoke_cfg, np.random.default_rng(0)))
    assert (pred["n_variants"].values == g["want_n_variants"]).all() and (pred["significant"].values == g["want_significant"]).all()
    with pytest.raises(NotImplementedError):
        get_variant_caller(smoke_cfg, True, "ensemble", False, "cnn_lstm")
    with pytest.raises(TypeError):
        VariantCaller(smoke_cfg)                      # abstract

    class Veto(VariantCaller):                        # a third-party plugin against the same interface
        def predict(self, collapsed_df, error_model_df):
            return pd.DataFrame({"sample_id": collapsed_df["sample_id"].unique(), "significant": False})

    assert not Veto(smoke_cfg).predict(collapsed, em)["significant"].any()


def test_synthetic_read_generator_matches_reference_distributions(ctx):
    """io.py:79-123 on the device: NegBinom(r=2)+1 sizes, allele concordance 0.9 / 0.95, quality
    int(N(30,5)) clipped [10,40], strand, read position [10,150), family ids, site alleles."""
    from precise_mrd.io import Read, SyntheticReadGenerator, TargetSite, site_cell_id

    site = TargetSite("chr7", 55249071, "C", "T", "ACG", gene="EGFR")
    assert site.key == "chr7:55249071:C>T" and hash(site) == hash(site.key)
    cid = site_cell_id(site)
    assert cid & 3 == 1 and ((cid & 3) + 1 + (cid >> 2) % 3) & 3 == 3          # C -> T
    gen = SyntheticReadGenerator(seed=5)
    reads = gen.generate_reads_for_site(site, n_umi_families=20000, allele_fraction=0.3, mean_family_size=5.0)
    assert isinstance(reads[0], Read) and {r.ref for r in reads[:2000]} == {"C"} and {r.alt for r in reads} == {"C", "T"}
    df = pd.DataFrame([r.__dict__ for r in reads])
    assert df["family_id"].nunique() == 20000 and df["family_id"].iloc[0] == "fam_000000"
    sizes = df.groupby("family_id", sort=False).size().values
    p = 2.0 / (2.0 + 5.0)
    ks = np.arange(1, 40)
    exp = sps.nbinom.pmf(ks - 1, 2, p) * len(sizes)
    obs = np.bincount(sizes, minlength=40)[1:40].astype(float)
    keep = exp > 5
    assert sps.chisquare(obs[keep] * exp[keep].sum() / obs[keep].sum(), exp[keep])[1] > ALPHA
    fam_alt = df.groupby("family_id", sort=False)["alt"].agg(lambda s: (s == "T").mean()).values
    variant = fam_alt > 0.5
    assert abs(variant.mean() - 0.3) < 4 * np.sqrt(0.3 * 0.7 / 20000) + 0.01      # majority vote blurs tiny families
    big = sizes >= 8
    assert abs(fam_alt[big & variant].mean() - 0.9) < 0.01 and abs(fam_alt[big & ~variant].mean() - 0.05) < 0.01
    q = df["quality"].values
    assert q.min() >= 10 and q.max() <= 40
    grid = np.arange(10, 41)
    cdf = sps.norm.cdf((grid + 1 - 30) / 5.0)
    cdf[-1] = 1.0
    exp_q = np.diff(np.r_[0.0, cdf]) * len(q)
    obs_q = np.bincount(q, minlength=41)[10:41].astype(float)
    keep = exp_q > 5
    assert sps.chisquare(obs_q[keep] * exp_q[keep].sum() / obs_q[keep].sum(), exp_q[keep])[1] > ALPHA
    assert abs((df["strand"] == "+").mean() - 0.5) < 0.01
    assert df["read_position"].min() >= 10 and df["read_position"].max() <= 149
    assert all(len(u) == 12 and set(u) <= set("ACGT") for u in df["umi"].iloc[:500])
    # successive calls draw fresh reads; a new generator with the same seed repeats the sequence
    again = gen.generate_reads_for_site(site, 50)
    fresh = SyntheticReadGenerator(seed=5)
    first = fresh.generate_reads_for_site(site, n_umi_families=20000, allele_fraction=0.3, mean_family_size=5.0)
    assert [r.umi for r in first[:200]] == [r.umi for r in reads[:200]] and [r.umi for r in again[:20]] != [r.umi for r in reads[:20]]
    # contamination flips the family's variant status (io.py:100-101)
    flipped = gen.generate_reads_for_site(site, 20000, allele_fraction=0.0, contamination_rate=0.25)
    d2 = pd.DataFrame([r.__dict__ for r in flipped])
    v2 = d2.groupby("family_id", sort=False)["alt"].agg(lambda s: (s == "T").mean()).values
    assert abs((v2 > 0.5).mean() - 0.25) < 0.02
    # the packed records feed UMIProcessor directly
    from precise_mrd.umi import UMIProcessor

    fams = UMIProcessor(min_family_size=3).process_reads(first[:5000])
    assert 0 < len(fams) < 5000 and all(f.family_size >= 3 for f in fams)


def test_stratified_analyzer(ctx, smoke_cfg, tmp_path):
    """eval/stratified.py: same result structure, context multipliers, calibration bins, artifacts."""
    from precise_mrd.eval.stratified import CONTEXT_ERROR_MULTIPLIER, StratifiedAnalyzer, context_seed_offset, run_stratified_analysis
    from precise_mrd.validation import validate_json

    an = StratifiedAnalyzer(smoke_cfg, np.random.default_rng(7))
    power = an.analyze_stratified_power(af_values=[0.001, 0.05], depth_values=[1000, 5000], contexts=["CpG", "NpN"], n_replicates=6)
    assert set(power) == {"stratified_results", "af_values", "depth_values", "contexts", "config_hash"}
    cell = power["stratified_results"]["CpG"][5000][0.05]
    assert set(cell) == {"mean_detection_rate", "std_detection_rate", "detection_rates", "n_replicates"}
    assert cell["n_replicates"] == 6 and len(cell["detection_rates"]) == 6 and set(cell["detection_rates"]) <= {0.0, 1.0}
    assert cell["mean_detection_rate"] == 1.0                                      # AF 0.05 >> background: always called
    assert power["stratified_results"]["NpN"][1000][0.001]["mean_detection_rate"] <= 0.5
    # deterministic: the reference's hash(context) is replaced by a fixed offset
    again = StratifiedAnalyzer(smoke_cfg, np.random.default_rng(7)).analyze_stratified_power(
        af_values=[0.001, 0.05], depth_values=[1000, 5000], contexts=["CpG", "NpN"], n_replicates=6)
    assert again == power and context_seed_offset("CpG") == context_seed_offset("CpG") < 1000
    # the context multiplier reaches the kernel: stored rate scaled once, false positives at rate * m^2
    cfg = an._create_context_config(0.0, 50000, "CpG")
    base = [an._simulate_with_context(cfg, np.random.default_rng(s), "CHH") for s in range(200)]
    cpg = [an._simulate_with_context(cfg, np.random.default_rng(s), "CpG") for s in range(200)]
    b_rate = np.array([d["background_rate"].iloc[0] for d in base])
    c_rate = np.array([d["background_rate"].iloc[0] for d in cpg])
    np.testing.assert_allclose(c_rate, b_rate * CONTEXT_ERROR_MULTIPLIER["CpG"], rtol=1e-15)
    assert (cpg[0]["context"] == "CpG").all() and cfg.run_id.endswith("_context_CpG")
    ratio = sum(d["n_false_positives"].iloc[0] for d in cpg) / sum(d["n_false_positives"].iloc[0] for d in base)
    assert abs(ratio - 1.5 ** 2) < 0.15
    # calibration: (lo, hi] bins, ECE / MCE definitions
    m = an._compute_calibration_metrics(np.array([0.0, 0.05, 0.1, 0.95, 1.0]), np.array([0, 0, 1, 1, 1]), 10)
    assert m["bin_counts"] == [2, 0, 0, 0, 0, 0, 0, 0, 0, 2]                        # 0.0 falls in no bin; 0.1 in the first
    assert m["bin_accuracies"][0] == 0.5 and m["bin_confidences"][0] == pytest.approx(0.075)
    assert m["ece"] == pytest.approx(0.425 * 0.4 + 0.025 * 0.4) and m["max_ce"] == pytest.approx(0.425)
    calib = an.analyze_calibration_by_bins(af_values=[0.01], depth_values=[1000], n_bins=10, n_replicates=8)
    row = calib["calibration_data"][0]
    assert row["n_samples"] == 8 and len(row["bin_counts"]) == 10 and 0.0 <= row["ece"] <= 1.0
    an.generate_stratified_reports(str(tmp_path))
    validate_json(tmp_path / "power_by_stratum.json", "power.schema.json")
    validate_json(tmp_path / "calibration_by_bin.json", "calibration.schema.json")
    assert list(pd.read_csv(tmp_path / "calibration_by_bin.csv").columns)[:4] == ["depth", "af", "ece", "max_ce"]
    assert callable(run_stratified_analysis)


def test_bootstrap_metrics_kernel_matches_sklearn_resample_for_resample(ctx):
    """K12 (metrics.py:18-118): point AUC / AP and every bootstrap resample against the oracle
    (itself pinned to sklearn in the CPU suite), on tie-heavy, tie-free, one-class and tiny inputs;
    the public ``bootstrap_metric`` consumes the caller's generator exactly as the reference does."""
    from sklearn.metrics import average_precision_score
    from sklearn.metrics import roc_auc_score as sk_auc

    from oracle import port
    from precise_mrd import metrics

    rng = np.random.default_rng(11)
    cases = [(rng.integers(0, 2, 500), np.round(rng.random(500), 2)),                  # heavy ties
             (rng.integers(0, 2, 3000), rng.normal(size=3000)),                        # no ties, negative scores
             ((rng.random(20000) < 0.01).astype(int), rng.random(20000) ** 3),         # C2-sized, rare positives
             (np.array([0, 1]), np.array([0.3, 0.3])), (np.array([1, 0, 1]), np.array([-0.0, 0.0, 1.0]))]
    for y, s in cases:
        pt, _, _ = ctx.bootstrap_metrics(y, s)
        assert pt[0] == pytest.approx(sk_auc(y, s), rel=1e-12) and pt[1] == pytest.approx(average_precision_score(y, s), rel=1e-12)
        idx = np.stack([rng.choice(len(y), len(y), replace=True) for _ in range(40)])
        _, auc, ap = ctx.bootstrap_metrics(y, s, indices=idx, point=False)
        want_auc = np.array([port.roc_auc(y[i], s[i]) for i in idx])
        want_ap = np.array([port.average_precision(y[i], s[i]) for i in idx])
        np.testing.assert_allclose(auc, want_auc, rtol=1e-12)
        np.testing.assert_allclose(ap, want_ap, rtol=1e-12, atol=1e-300)
    # one class only: the reference's fallbacks (metrics.py:18-33)
    pt, auc, ap = ctx.bootstrap_metrics(np.zeros(50), rng.random(50), n_boot=8)
    assert pt.tolist() == [0.5, 0.0] and (auc == 0.5).all() and (ap == 0.0).all()
    pt, _, _ = ctx.bootstrap_metrics(np.ones(50), rng.random(50))
    assert pt.tolist() == [0.5, 1.0]
    # Philox resampling (no indices given): deterministic per (seed, stream), distribution matches numpy's
    y, s = cases[0]
    _, a1, _ = ctx.bootstrap_metrics(y, s, n_boot=2000, stream_id=3, point=False)
    _, a2, _ = ctx.bootstrap_metrics(y, s, n_boot=2000, stream_id=3, point=False)
    _, a3, _ = ctx.bootstrap_metrics(y, s, n_boot=2000, stream_id=4, point=False)
    assert (a1 == a2).all() and (a1 != a3).any()
    _, ref_scores = port.bootstrap_metric(y, s, port.roc_auc, 2000, np.random.default_rng(5))
    assert sps.ks_2samp(a1, ref_scores)[1] > ALPHA
    with pytest.raises(Exception, match="out of range"):
        ctx.bootstrap_metrics(y, s, indices=np.full((1, len(y)), len(y)), point=False)
    # public API: same generator consumption and the same summary as the reference formula
    y, s = cases[1]
    got = metrics.bootstrap_metric(y, s, metrics.roc_auc_score, 200, np.random.default_rng(3))
    want, _ = port.bootstrap_metric(y, s, port.roc_auc, 200, np.random.default_rng(3))
    assert got == pytest.approx(want, rel=1e-12)
    got = metrics.bootstrap_metric(y, s, metrics.average_precision, 200, np.random.default_rng(3))
    want, _ = port.bootstrap_metric(y, s, port.average_precision, 200, np.random.default_rng(3))
    assert got == pytest.approx(want, rel=1e-12)
    g1, g2 = np.random.default_rng(9), np.random.default_rng(9)
    metrics.bootstrap_metric(y, s, metrics.roc_auc_score, 7, g1)
    port.bootstrap_metric(y, s, port.roc_auc, 7, g2)
    assert g1.random() == g2.random()
    assert metrics.roc_auc_score(y, s) == pytest.approx(sk_auc(y, s), rel=1e-12)
    assert metrics.average_precision(y, s) == pytest.approx(average_precision_score(y, s), rel=1e-12)
    assert metrics.bootstrap_metric(y, s, lambda a, b: float(np.mean(a)), 5, np.random.default_rng(0))["mean"] > 0   # host callable


def test_fastq_ingest_matches_reference(ctx, smoke_cfg, tmp_path):
    """K13 (fastq.py): the device scan + family table on a synthetic FASTQ with every header
    convention, CRLF records, malformed records, UMI-less reads, >12-letter UMIs and a blank-line
    end -- against what the reference's own fastq.py produced for the same file."""
    import gzip

    from precise_mrd.collapse import collapse_umis
    from precise_mrd.fastq import FASTQReader, detect_umi_format, process_fastq_to_dataframe

    path = ROOT / "tests" / "golden" / "fastq_mixed.fastq"
    want = json.loads((ROOT / "tests" / "golden" / "fastq_mixed.json").read_text())
    reader = FASTQReader(path)
    reads = list(reader.read_fastq())
    assert len(reads) == len(want["reads"])
    for got, w in zip(reads, want["reads"]):
        assert (got["read_id"], got["sequence"], got["umi"], got["sequence_length"]) == (w["read_id"], w["sequence"], w["umi"], w["sequence_length"])
        assert got["mean_quality"] == w["mean_quality"] and sum(got["quality_scores"]) == w["qual_sum"]
    assert [r["read_id"] for r in reader.read_fastq(max_reads=50)] == want["reads_max_50"]
    v = reader.validate_fastq()
    for k, w in want["validate"].items():
        assert v[k] == (pytest.approx(w, rel=1e-15) if isinstance(w, float) else w), k
    for tag, mr in (("all", None), ("max_300", 300)):
        df = process_fastq_to_dataframe(path, smoke_cfg, np.random.default_rng(0), max_reads=mr)
        w = want[f"table_{tag}"]
        assert list(df.columns) == w["columns"]
        for col in ("sample_id", "umi", "family_size", "consensus_sequence", "consensus_agreement", "passes_quality", "passes_consensus"):
            assert df[col].tolist() == w[col], (tag, col)
        assert df["mean_quality"].tolist() == w["mean_quality"]                       # exact: integer sums, one division
        assert [sum(q) for q in df["quality_scores"]] == w["quality_scores_sum"]
        assert [q[:5] for q in df["quality_scores"]] == w["quality_scores_head"]
        assert (df["config_hash"] == smoke_cfg.config_hash()).all()
    assert detect_umi_format(path, 100) == want["format"]
    # gzip input and the hand-off to collapse_umis(is_fastq_data=True) (collapse.py:46-64)
    gz = tmp_path / "mixed.fastq.gz"
    gz.write_bytes(gzip.compress(path.read_bytes()))
    df_gz = process_fastq_to_dataframe(gz, smoke_cfg, np.random.default_rng(0), output_path=str(tmp_path / "fq.parquet"))
    assert df_gz["umi"].tolist() == want["table_all"]["umi"] and (tmp_path / "fq.parquet").exists()
    collapsed = collapse_umis(df_gz, smoke_cfg, np.random.default_rng(0), is_fastq_data=True)
    assert len(collapsed) == len(df_gz) and collapsed["family_size"].tolist() == want["table_all"]["family_size"]
    # scalar helpers and empty / UMI-less input
    assert reader.extract_umi("@x UMI: UMI:ACGT z") == "ACGT" and reader.extract_umi("@a_ACGTACGTA_b") == "ACGTACGTA"
    assert reader.extract_umi("@a:b:ACGTACGTAC") == "ACGTACGTAC" and reader.extract_umi("@a_ACGTACGTACGTA_") is None
    assert reader.parse_quality_string("I5!") == [40, 20, 0]
    empty = tmp_path / "empty.fastq"
    empty.write_text("")
    assert list(FASTQReader(empty).read_fastq()) == [] and detect_umi_format(empty) == "No UMI detected in sample"
    plain = tmp_path / "plain.fastq"
    plain.write_text("@r1\nACGT\n+\nIIII\n@r2\nAC\n+\nII")                            # no UMIs, no trailing newline
    assert [r["read_id"] for r in FASTQReader(plain).read_fastq()] == ["r1", "r2"]
    assert process_fastq_to_dataframe(plain, smoke_cfg, np.random.default_rng(0)).empty


def test_quality_filters_match_reference(ctx):
    """K14 Fisher strand bias + K8 end-repair enrichment behind QualityFilter (filters.py), against
    the reference's own outputs on a seeded 500-variant table (tests/golden/gb_filters.json)."""
    from scipy.stats import fisher_exact

    from precise_mrd.filters import FilterResult, QualityFilter

    g = json.loads((ROOT / "tests" / "golden" / "gb_filters.json").read_text())
    qf = QualityFilter()
    df = pd.DataFrame(g["rows"])
    out = qf.filter_variant_dataframe(df)
    for col, want in g["table"].items():
        got = [None if (v is None or (isinstance(v, float) and np.isnan(v))) else (bool(v) if isinstance(v, (bool, np.bool_)) else v)
               for v in out[col]]
        assert got == want, col
    assert list(df.columns) == [c for c in out.columns if c in df.columns]          # the input frame is not mutated
    rep = qf.generate_filter_report(out)
    for k, w in g["report"].items():
        assert (rep[k] == w) if not isinstance(w, float) else rep[k] == pytest.approx(w, rel=1e-15), k
    for r, w in zip(g["rows"], g["scalar"]):
        sb = qf.test_strand_bias(r["alt_forward"], r["alt_reverse"], r["ref_forward"], r["ref_reverse"])
        assert isinstance(sb, FilterResult) and (bool(sb.passed), sb.reason) == (w["sb"][0], w["sb"][1])
        if w["sb"][3] is not None:
            # beyond ~2e4 reads SciPy's hypergeometric tails are good to ~3e-11 only (mpmath check in DESIGN.md)
            assert sb.pvalue == pytest.approx(w["sb"][3], rel=1e-12 if r["total_depth"] <= 20000 else 5e-11, abs=1e-300)
            assert (np.isnan(sb.statistic) and np.isnan(w["sb"][2])) or sb.statistic == pytest.approx(w["sb"][2], rel=1e-15)
        er = qf.test_end_repair_artifact(r["ref"], r["alt"], r["read_positions"])
        assert (bool(er.passed), er.reason) == (w["er"][0], w["er"][1])
        if w["er"][3] is not None:
            assert er.pvalue == pytest.approx(w["er"][3], rel=1e-12, abs=1e-300) and er.statistic == pytest.approx(w["er"][2], rel=1e-15)
    # the raw kernel against scipy over a wide range of tables, incl. empty rows / columns
    rng = np.random.default_rng(8)
    tabs = np.array([rng.integers(0, s, 4) for s in rng.choice([3, 30, 300, 3000], 600)], dtype=np.int64)
    tabs[::17, 2:] = 0
    tabs[::19, 1] = 0
    odds, p = ctx.fisher_exact(tabs)
    for t, o, pv in zip(tabs, odds, p):
        w = fisher_exact(t.reshape(2, 2))
        assert pv == pytest.approx(w.pvalue, rel=1e-12, abs=1e-300), t
        assert (np.isnan(o) and np.isnan(w.statistic)) or o == w.statistic, t
    with pytest.raises(Exception, match="nonnegative"):
        ctx.fisher_exact([[1, -1, 2, 3]])
    # contract of the simple filters and the summary (filters.py:168-203, 245-260)
    assert qf.filter_by_depth(1, 100).reason == "insufficient_alt_count" and qf.filter_by_depth(5, 3).reason == "insufficient_total_depth"
    assert qf.filter_by_quality([]).reason == "no_quality_scores" and qf.filter_by_quality([10, 12]).reason == "low_quality"
    res = qf.apply_all_filters({"alt_count": 5, "total_depth": 100, "quality_scores": [30, 35]})
    assert set(res) == {"depth", "quality"} and qf.summarize_filters(res)["all_filters_passed"]
    assert QualityFilter(end_repair_filter=False).test_end_repair_artifact("G", "T", [1, 2, 3]).passed
    assert qf.filter_variant_dataframe(pd.DataFrame()).empty and qf.generate_filter_report(pd.DataFrame()) == {"total_variants": 0}


# ------------------------------------------------------------------------------ QC report + feature table (K15, K16)
def _qc_golden():
    import json

    with open(ROOT / "tests" / "golden" / "gb_qc.json") as f:
        g = json.load(f)
    fam = g["families"]
    families = [{"site_key": g["site_keys"][s], "family_size": z, "consensus_allele": (None if a == "-" else a)}
                for s, z, a in zip(fam["site"], fam["size"], fam["allele"])]
    return g, families, pd.DataFrame(g["consensus"])


def _same(got, want, path=""):
    """recursive equality: ints / strings / bools exact, floats to 1e-12 relative (NaN == NaN)"""
    if isinstance(want, dict):
        assert set(map(str, got.keys())) == set(want.keys()), path
        for k, v in want.items():
            _same({str(kk): vv for kk, vv in got.items()}[k], v, f"{path}/{k}")
    elif isinstance(want, list):
        assert len(got) == len(want), path
        for i, (a, b) in enumerate(zip(got, want)):
            _same(a, b, f"{path}[{i}]")
    elif isinstance(want, float):
        assert (np.isnan(want) and np.isnan(got)) or abs(float(got) - want) <= 1e-12 * max(1.0, abs(want)), (path, got, want)
    else:
        assert got == want or (isinstance(want, int) and int(got) == want), (path, got, want)


def test_qc_analyzer_matches_reference_report(ctx):
    from precise_mrd.qc import QCAnalyzer, QCThresholds

    g, families, consensus = _qc_golden()
    qa = QCAnalyzer(QCThresholds())
    _same(qa.analyze_family_size_distribution([f["family_size"] for f in families]), g["family_size"], "family_size")
    _same(qa.analyze_family_size_distribution([4, 4, 4]), g["family_size_const"], "family_size_const")
    _same(qa.calculate_depth_metrics(consensus), g["depth"], "depth")
    _same(qa.estimate_contamination_metrics(consensus), g["contamination"], "contamination")
    _same(qa.analyze_consensus_efficiency(families), g["efficiency"], "efficiency")
    rep = qa.generate_comprehensive_qc_report(consensus, families, {"seed": 77})
    _same(json_roundtrip(rep), g["report"], "report")
    assert qa.analyze_family_size_distribution([]) == {"error": "No family sizes provided"}
    assert qa.analyze_consensus_efficiency([]) == {"error": "No families data provided"}
    # sizes beyond the device histogram are counted exactly
    big = qa.analyze_family_size_distribution([5, 70000, 5, 4095, 4094])
    assert big["max"] == 70000 and big["size_distribution"] == {5: 2, 70000: 1, 4095: 1, 4094: 1} and big["median"] == 4094.0


def json_roundtrip(o):
    import json

    def conv(x):
        if isinstance(x, dict):
            return {str(k): conv(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [conv(v) for v in x]
        if isinstance(x, np.integer):
            return int(x)
        if isinstance(x, np.floating):
            return float(x)
        if isinstance(x, np.bool_):
            return bool(x)
        return x

    return json.loads(json.dumps(conv(o)))


def test_qc_report_from_device_collapse_tables(ctx, native):
    """report_from_tables (device family / group tables) == the list-of-dicts report on the same collapse."""
    from precise_mrd.qc import QCAnalyzer

    rng = np.random.default_rng(12)
    n, n_sites = 60000, 12
    keys = (rng.integers(0, n_sites, n, dtype=np.uint64) << np.uint64(32)) | rng.integers(0, 1500, n, dtype=np.uint64)
    q = rng.choice([12, 25, 30, 37], n).astype(np.uint32)
    pl = (rng.choice([0, 0, 0, 0, 0, 0, 0, 1, 2], n).astype(np.uint32)) | (q << np.uint32(2))
    P = native.UmiParams(min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6,
                         mode=native.MODE_WEIGHTED, cluster=native.CLUSTER_EXACT, max_distance=0, umi_len=12)
    fam, grp = ctx.collapse_reads(keys, pl, P)
    qa = QCAnalyzer()
    rep = qa.report_from_tables(fam, grp)
    group = (fam["key"] >> np.uint64(32)).astype(np.int64)
    has = (fam["flags"] & native.FAM_HAS_CONSENSUS) != 0
    families = [{"site_key": f"site{g}", "family_size": int(s), "consensus_allele": ("ACGT"[int(c)] if h else None)}
                for g, s, c, h in zip(group, fam["size"], fam["consensus"], has)]
    rows = [{"site_key": f"site{int(g)}", "allele": "ACGT"[a], "ref": "A", "consensus_count": int(c)}
            for g, cc in zip(grp["group"], grp["consensus_count"]) for a, c in enumerate(cc) if c > 0]
    want = qa.generate_comprehensive_qc_report(pd.DataFrame(rows), families)
    _same(json_roundtrip(rep), json_roundtrip(want), "report")
    assert rep["overall_metrics"]["total_families"] == len(fam["size"]) and rep["site_qc_summary"]["total_sites_analyzed"] == 10


def test_feature_table_bit_equal_to_reference(ctx):
    from precise_mrd.advanced_stats import FeatureEngineer

    g, _, _ = _qc_golden()
    fr = g["features"]["frame"]
    df = pd.DataFrame({"family_size": np.array(fr["family_size"], dtype=np.int64), "quality_score": fr["quality_score"],
                       "consensus_agreement": fr["consensus_agreement"], "allele_fraction": fr["allele_fraction"]})
    X, names = FeatureEngineer(None).extract_features(df)
    assert names == g["features"]["names"] and list(X.columns) == names
    want = np.array([[float.fromhex(v) for v in row] for row in g["features"]["matrix_hex"]])
    assert X.to_numpy().tobytes() == want.tobytes()                      # bit-exact
    X2, names2 = FeatureEngineer(None).extract_features(df[["family_size", "allele_fraction"]])
    assert names2 == g["features_partial"]["names"] and list(X2.shape) == g["features_partial"]["shape"]
    Xs, sel = FeatureEngineer(None).select_features(X, (df["allele_fraction"] > 0.001).to_numpy().astype(int), method="f_classif", k=3)
    assert len(sel) == 3 and Xs.shape == (len(df), 3)
    with pytest.raises(ValueError):
        FeatureEngineer(None).select_features(X, np.zeros(len(df)), method="nope")
