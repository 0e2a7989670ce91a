"""Pins oracle/port.py (the CPU restatement) against the reference's own outputs.

Every fixture under tests/golden/ was produced by running the reference in the build
container (oracle/make_golden.py); the oracle must reproduce them before it is trusted
as the checker of the CUDA path."""
import numpy as np
import pytest

from oracle import port


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


@pytest.mark.parametrize("name", ["call_smoke_v1.npz", "call_smoke_v2.npz"])
def test_call_stage_matches_reference_fixture(golden, name):
    g = golden(name)
    out = port.call_segments(g["seg_offsets"], g["is_variant"], g["quality"], g["agreement"],
                             float(g["mean_error_rate"]), "poisson", float(g["alpha"]))
    for col in ("n_variants", "n_total", "significant"):
        assert (out[col] == g[f"want_{col}"]).all(), col
    for col in ("variant_fraction", "expected_variants", "fold_change", "p_value", "p_adjusted"):
        assert (out[col] == g[f"want_{col}"]).all(), col          # same SciPy, same operation order: bit-equal
    for col in ("mean_quality", "mean_consensus"):
        assert _rel(out[col], g[f"want_{col}"]).max() < 1e-14, col
    assert (g["want_significant"] == False).all()


def test_kat_rows(golden):
    # the values the reference code actually returns
    assert port.poisson_test(5, 1.0) == pytest.approx(0.007319693654687426, rel=1e-15)
    assert port.poisson_test(1, 5.0) == pytest.approx(0.08085536398902558, rel=1e-15)
    assert port.poisson_test(5, 5.0) == 1.0
    assert port.poisson_test(0, 0.0) == 1.0
    assert port.poisson_test(10, 0.1) == pytest.approx(5.032695613540631e-17, rel=1e-14)
    assert port.binomial_test(10, 100, 0.01) == pytest.approx(7.631587532260653e-08, rel=1e-14)
    assert port.binomial_test(1, 100, 0.1) == pytest.approx(0.0006336060581826209, rel=1e-14)
    assert port.binomial_test(10, 100, 0.1) == 1.0
    assert [port.binomial_test(*a) for a in [(0, 0, .5), (0, 10, 0.), (1, 10, 0.), (10, 10, 1.), (9, 10, 1.)]] == [1, 1, 0, 1, 0]
    g = golden("stats_kat.npz")
    got = np.array([port.poisson_test(int(k), float(e)) for k, e in zip(g["pois_k"][:400], g["pois_e"][:400])])
    assert (got == g["pois_p"][:400]).all()
    got = np.array([port.binomial_test(int(k), int(n), float(p)) for k, n, p in zip(g["bin_k"][:300], g["bin_n"][:300], g["bin_p"][:300])])
    assert (got == g["bin_pv"][:300]).all()


def test_bh_matches_reference(golden):
    g = golden("stats_kat.npz")
    for i in range(int(g["n_bh"])):
        rej, adj = port.benjamini_hochberg(g[f"bh{i}_p"], float(g[f"bh{i}_alpha"]))
        assert (rej == g[f"bh{i}_rej"]).all(), i
        assert (adj == g[f"bh{i}_adj"]).all(), i
    rej, adj = port.benjamini_hochberg(np.array([]), 0.05)
    assert rej.shape == (0,) and adj.shape == (0,)


def test_ga_collapse_matches_reference(golden):
    g = golden("ga_reads.npz")
    fam, grp = port.collapse_firstfit(g["keys"], g["payload"], mode=1, min_fs=2, max_fs=1 << 30, max_distance=1)
    kept = (fam["flags"] & port.FAM_KEPT) != 0
    assert (fam["key"][kept] == g["want_fam_key"]).all()
    assert (fam["size"][kept] == g["want_fam_size"]).all()
    assert (fam["n_alt"][kept] == g["want_fam_alt"]).all()
    assert (fam["consensus"][kept] == g["want_fam_consensus"]).all()
    locus = g["group_locus"][grp["group"]]
    for name, col in (("want_locus_alt", "alt_count"), ("want_locus_total", "total_count"), ("want_locus_families", "family_count")):
        agg = np.zeros(locus.max() + 1, np.int64)
        np.add.at(agg, locus, grp[col].astype(np.int64))
        assert (agg[g["want_locus"]] == g[name]).all(), name


def test_gc_counts_match_reference(golden):
    g = golden("gc_reads.npz")
    fam, _ = port.collapse_exact(g["keys"], g["payload"], mode=2, min_fs=1, max_fs=1 << 30)
    order = np.argsort(g["want_fam_key"], kind="stable")
    assert (fam["key"] == g["want_fam_key"][order]).all()
    assert (fam["size"] == g["want_size"][order]).all()
    assert (fam["n_alt"] == g["want_alt"][order]).all()
    assert (fam["size"] - fam["n_alt"] == g["want_ref"][order]).all()


def test_gb_weighted_consensus_matches_reference(golden):
    g = golden("gb_reads.npz")
    mn, mx, qthr = [int(v) for v in g["params"]]
    fam, grp = port.collapse_exact(g["keys"], g["payload"], mode=0, min_fs=mn, max_fs=mx, qthr=qthr,
                                   cthr=float(g["consensus_threshold"]))
    order = np.argsort(g["want_fam_key"], kind="stable")
    assert (fam["key"] == g["want_fam_key"][order]).all()
    assert (fam["size"] == g["want_fam_size"][order]).all()
    has = (fam["flags"] & port.FAM_HAS_CONSENSUS) != 0
    want_allele = g["want_fam_allele"][order]
    assert (has == (want_allele >= 0)).all()
    assert (fam["consensus"][has] == want_allele[has]).all()
    q = fam["qsum"][has] / fam["n_best"][has]
    assert (q == g["want_fam_quality"][order][has]).all()          # mean of ints: exact
    assert int(((fam["flags"] & port.FAM_KEPT) != 0).sum()) == int(g["want_kept"])
    assert (grp["consensus_count"] == g["want_consensus_count"][grp["group"]]).all()


def check_greedy_tables_against_reference(fam, g, d, has_flag):
    """Family table of a greedy-clustered collapse (group-major, seed order) vs the reference-generated vectors."""
    assert len(fam["key"]) == len(g[f"want_fam_key_d{d}"])
    assert (fam["key"] == g[f"want_fam_key_d{d}"]).all()           # same clusters, same order, same first-read UMI
    assert (fam["size"] == g[f"want_fam_size_d{d}"]).all()
    has = (fam["flags"] & has_flag) != 0
    want_allele = g[f"want_fam_allele_d{d}"]
    assert (has == (want_allele >= 0)).all()
    assert (fam["consensus"][has] == want_allele[has]).all()
    assert (fam["qsum"][has] / fam["n_best"][has] == g[f"want_fam_quality_d{d}"][has]).all()


@pytest.mark.parametrize("d", [1, 2])
def test_gb_greedy_clustering_matches_reference(golden, d):
    """max_edit_distance > 0 (configs/default.yaml:16): the port's greedy Hamming clustering equals the reference's
    cluster_umis_greedy (umi_clustering.py:59-101) cluster for cluster, in its order, consensus included."""
    g = golden("gb_greedy.npz")
    mn, mx, qthr = [int(v) for v in g["params"]]
    fam, grp = port.collapse_greedy(g["keys"], g["payload"], 0, mn, mx, d, qthr, float(g["consensus_threshold"]))
    check_greedy_tables_against_reference(fam, g, d, port.FAM_HAS_CONSENSUS)
    assert (g[f"want_cluster_umis_d{d}"] > 1).sum() > 1000         # the fixture does merge barcodes
    # distance 0 degenerates to exact matching
    f0, g0 = port.collapse_greedy(g["keys"], g["payload"], 0, mn, mx, 0, qthr, float(g["consensus_threshold"]))
    e0, ge0 = port.collapse_exact(g["keys"], g["payload"], 0, mn, mx, qthr, float(g["consensus_threshold"]))
    assert sorted(f0["key"].tolist()) == e0["key"].tolist() and (g0["consensus_count"] == ge0["consensus_count"]).all()


def test_c_oracle_collapse_matches_reference_and_port(golden):
    """oracle/pmrd_oracle.c (the cpu_baseline port) reproduces the reference-generated G-B fixture."""
    from oracle import cport

    g = golden("gb_reads.npz")
    mn, mx, qthr = [int(v) for v in g["params"]]
    fam = cport.collapse_weighted(g["keys"], g["payload"], mn, mx, qthr, float(g["consensus_threshold"]))
    want, _ = port.collapse_exact(g["keys"], g["payload"], 0, mn, mx, qthr, float(g["consensus_threshold"]))
    for k in want:
        assert (fam[k] == want[k]).all(), k
    order = np.argsort(g["want_fam_key"], kind="stable")
    has = (fam["flags"] & port.FAM_HAS_CONSENSUS) != 0
    assert (has == (g["want_fam_allele"][order] >= 0)).all()


def test_c_oracle_statistics_and_pipeline_sanity():
    from oracle import cport

    lib = cport.load()
    rng = np.random.default_rng(4)
    for _ in range(300):
        mu = 10 ** rng.uniform(-3, 3)
        k = int(rng.poisson(mu * rng.choice([0.5, 1, 2])))
        want = port.poisson_test(k, mu)
        got = lib.oracle_poisson_two_sided(k, mu)
        assert abs(got - want) <= 1e-9 * max(want, 1e-300) + 1e-15
    p = rng.uniform(size=500) ** 3
    c = cport.call_cells(np.full(500, 1000), rng.poisson(1.0, 500), np.ones(500), seed=1)
    rej, adj = port.benjamini_hochberg(c["p_value"], 0.05)
    assert (rej == c["significant"]).all() and np.allclose(adj, c["p_adjusted"], rtol=1e-15, atol=0)
    cells = np.arange(24)
    af = np.tile([0.0, 0.1], 12)
    r1 = cport.pipeline_cells(cells, af, np.full(24, 2000), seed=3, n_threads=1)
    r2 = cport.pipeline_cells(cells, af, np.full(24, 2000), seed=3, n_threads=4)
    assert (r1["n_total"] == r2["n_total"]).all() and (r1["n_variants"] == r2["n_variants"]).all()   # thread-count independent
    assert abs(r1["total_reads"] / (24 * 2000) - 5.03) < 0.1
    frac = r1["n_variants"] / r1["n_total"]
    assert frac[af == 0].max() < 0.03 and abs(frac[af > 0].mean() - 0.1) < 0.03   # 5% per-read error leaks ~1.4% alt consensus at AF 0
