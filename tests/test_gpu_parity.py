"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libpmrd_b200.so),
against (a) the golden vectors exported from the reference and (b) the oracle on seeded
inputs.  Integer / index work is bit-exact; fp64 tails are compared at the tolerances
stated (and justified) inline."""
import numpy as np
import pytest

from oracle import port

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def _params(native, **kw):
    d = dict(min_family_size=1, max_family_size=1 << 30, quality_threshold=0, consensus_threshold=0.0,
             mode=native.MODE_COUNT, cluster=native.CLUSTER_EXACT, max_distance=0, umi_len=12)
    d.update(kw)
    return native.UmiParams(**d)


# ------------------------------------------------------------------------------ K4 sort
@pytest.mark.parametrize("n", [0, 1, 2, 31, 4095, 4096, 4097, 70_001, 1_500_000])
@pytest.mark.parametrize("pattern", ["full64", "umi24_group11", "few_distinct", "constant"])
def test_sort_pairs_stable(ctx, n, pattern):
    rng = np.random.default_rng(n * 7 + len(pattern))
    if pattern == "full64":
        keys = rng.integers(0, 2**63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    elif pattern == "umi24_group11":
        keys = (rng.integers(0, 2000, n, dtype=np.uint64) << np.uint64(32)) | rng.integers(0, 1 << 24, n, dtype=np.uint64)
    elif pattern == "few_distinct":
        keys = rng.integers(0, 7, n, dtype=np.uint64) * np.uint64(0x0101010101010101)
    else:
        keys = np.full(n, 0xDEADBEEF12345678, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32)
    ko, vo = ctx.sort_pairs(keys, vals)
    wk, wv = port.sort_pairs(keys, vals)
    assert (ko == wk).all()
    assert (vo == wv).all()          # stability: equal keys keep input order
    ko2, _ = ctx.sort_pairs(keys)    # keys-only variant
    assert (ko2 == wk).all()


# ------------------------------------------------------------------------------ K8 p-values
def test_poisson_two_sided_vs_reference_golden(ctx, golden, native):
    g = golden("stats_kat.npz")
    got = ctx.pvalues(g["pois_k"], None, g["pois_e"], native.TEST_POISSON, native.ALT_TWO_SIDED)
    want = g["pois_p"]
    rel = _rel(got, want)
    # 1e-12 relative down to p = 1e-50.  Deeper, SciPy's own igam/igamc drifts (checked against
    # mpmath, 60 digits: (k=5495, mu=8255.353) exact 1.09942091479873e-229, SciPy ...47866e-229
    # = 1.1e-11 off, this kernel ...47990e-229 = 2e-13 off; (1380, 763.79): SciPy 3.2e-12 off,
    # kernel 6e-14 off), so below 1e-50 the gate against the reference is 2e-11 and
    # test_tails_against_mpmath holds the kernel itself to 1e-12 there.
    shallow = want >= 1e-50
    assert rel[shallow].max() < 1e-12
    assert rel[~shallow & (want > 0)].max() < 2e-11
    assert (got[want == 0] < 1e-300).all()
    assert (got[want == 1.0] == 1.0).all()
    # KAT rows: edge cases are exact
    assert got[2] == 1.0 and got[3] == 1.0 and got[5] == 1.0


def test_binomial_two_sided_vs_reference_golden(ctx, golden, native):
    g = golden("stats_kat.npz")
    got = ctx.pvalues(g["bin_k"], g["bin_n"], g["bin_p"], native.TEST_BINOMIAL, native.ALT_TWO_SIDED)
    want = g["bin_pv"]
    rel = _rel(got, want)
    # scipy.stats.binomtest goes through Boost ibeta, which is itself only ~3e-12 accurate for
    # n ~ 1e5 (mpmath: binomtest(7, 88655, 6.8095e-05) exact 0.679765642907125639, SciPy
    # ...9061246 (1.5e-12 off), this kernel ...9071257 (1e-16 off)).  Gate: 1e-12 for n <= 2e4,
    # 1e-11 above.
    small = g["bin_n"] <= 20000
    assert rel[small].max() < 1e-12
    assert rel[~small].max() < 1e-11
    # special cases of call.py:30-37
    assert list(got[3:8]) == [1.0, 1.0, 0.0, 1.0, 0.0]


def test_pvalues_one_sided_and_n_times_rate(ctx, native):
    rng = np.random.default_rng(5)
    m = 2000
    n = rng.integers(1, 60000, m)
    rate = 10 ** rng.uniform(-5, -1.5, m)
    k = rng.poisson(n * rate * rng.choice([0.5, 1, 2, 5], m))
    got = ctx.pvalues(k, n, rate, native.TEST_POISSON, native.ALT_TWO_SIDED)
    want = np.array([port.poisson_test(int(a), int(b) * float(c)) for a, b, c in zip(k, n, rate)])
    assert _rel(got, want)[want >= 1e-50].max() < 1e-12
    got = ctx.pvalues(k, n, rate, native.TEST_POISSON, native.ALT_GREATER)
    want = np.array([port.poisson_greater(int(a), int(b) * float(c)) for a, b, c in zip(k, n, rate)])
    assert _rel(got, want)[want >= 1e-50].max() < 1e-12
    k = np.minimum(k, n)
    got = ctx.pvalues(k, n, rate, native.TEST_BINOMIAL, native.ALT_GREATER)
    want = np.array([port.binomial_greater(int(a), int(b), float(c)) for a, b, c in zip(k, n, rate)])
    ok = want >= 1e-50
    assert _rel(got, want)[ok & (n <= 20000)].max() < 1e-12
    assert _rel(got, want)[ok].max() < 1e-11
    # scalar rate broadcast
    got = ctx.pvalues(k[:50], n[:50], 1e-3, native.TEST_POISSON)
    want = np.array([port.poisson_test(int(a), int(b) * 1e-3) for a, b in zip(k[:50], n[:50])])
    assert _rel(got, want).max() < 1e-12
    with pytest.raises(native.NativeError):
        ctx.pvalues(k[:5], n[:5], 1e-3, 7)       # "Unknown test type" (call.py:187)


def test_tails_against_mpmath(ctx, native):
    """Where the reference (SciPy) is itself inexact, hold the kernel to the true value."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60

    def pois2(k, mu):
        mu = mp.mpf(mu)
        t = mp.e ** (-mu) * mu ** k / mp.factorial(k)
        s, j, tt = mp.mpf(0), k, t
        if k > mu:
            while tt > s * mp.mpf(10) ** -40:
                s += tt; j += 1; tt = tt * mu / j
            return min(1, 2 * min(s, 1 - s + t))
        while True:
            s += tt
            if j == 0 or tt < s * mp.mpf(10) ** -40:
                break
            tt = tt * j / mu; j -= 1
        return min(1, 2 * min(s, 1 - s + t))

    cases = [(5495, 8255.35343336275), (5620, 8299.835485743746), (1380, 763.788540437689), (3567, 2012.154107293812),
             (10, 0.1), (5, 1.0), (0, 30.0), (100000, 99000.0)]
    got = ctx.pvalues([c[0] for c in cases], None, [c[1] for c in cases], native.TEST_POISSON)
    for (k, mu), v in zip(cases, got):
        ex = pois2(k, mu)
        assert abs(mp.mpf(float(v)) - ex) / ex < mp.mpf("1e-12"), (k, mu, v)

    def binom_minlike(k, n, p):
        pm = lambda i: mp.binomial(n, i) * mp.mpf(p) ** i * (1 - mp.mpf(p)) ** (n - i)
        d = pm(k) * (1 + mp.mpf("1e-7"))
        sd = int(60 * (n * p) ** 0.5 + 60)
        return min(1, sum(f for f in (pm(i) for i in range(max(0, int(n * p) - sd), min(n, int(n * p) + sd) + 1)) if f <= d))

    cases = [(7, 88655, 6.809545321347221e-05), (16, 100767, 0.00014704516006750236), (9, 131776, 0.00017227525854530073)]
    got = ctx.pvalues([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], native.TEST_BINOMIAL)
    for (k, n, p), v in zip(cases, got):
        ex = binom_minlike(k, n, p)
        assert abs(mp.mpf(float(v)) - ex) / ex < mp.mpf("1e-12"), (k, n, p, v)


# ------------------------------------------------------------------------------ K9 BH
def test_bh_vs_reference_golden(ctx, golden):
    g = golden("stats_kat.npz")
    for i in range(int(g["n_bh"])):
        rej, adj = ctx.bh(g[f"bh{i}_p"], float(g[f"bh{i}_alpha"]))
        assert (rej == g[f"bh{i}_rej"]).all(), i
        assert (adj == g[f"bh{i}_adj"]).all(), i      # same IEEE ops in the same order: bit-equal
    rej, adj = ctx.bh(np.array([]), 0.05)
    assert rej.shape == (0,) and adj.shape == (0,)


@pytest.mark.parametrize("m", [1, 2047, 2048, 2049, 300_000])
def test_bh_vs_oracle_random(ctx, m):
    rng = np.random.default_rng(m)
    p = rng.uniform(0, 1, m) ** 4
    p[rng.integers(0, m, max(1, m // 7))] = 1.0
    for alpha in (0.0, 0.05, 1.0):
        rej, adj = ctx.bh(p, alpha)
        wrej, wadj = port.benjamini_hochberg(p, alpha)
        assert (rej == wrej).all() and (adj == wadj).all()
    # properties (tests/stat_tests/test_bh_fdr.py): monotone in p, adj >= p, idempotent decisions
    order = np.argsort(p, kind="stable")
    assert (np.diff(adj[order]) >= 0).all() and (adj >= p).all()


# ------------------------------------------------------------------------------ call stage
@pytest.mark.parametrize("name", ["call_smoke_v1.npz", "call_smoke_v2.npz"])
def test_call_stage_vs_reference_fixture(ctx, golden, native, name):
    g = golden(name)
    out = ctx.call(g["seg_offsets"], g["is_variant"], g["quality"], g["agreement"], float(g["mean_error_rate"]),
                   native.TEST_POISSON, float(g["alpha"]))
    for col in ("n_variants", "n_total", "significant"):
        assert (out[col] == g[f"want_{col}"]).all(), col             # bit-exact class
    for col in ("variant_fraction", "expected_variants", "fold_change"):
        assert (out[col] == g[f"want_{col}"]).all(), col             # single IEEE ops: also bit-exact
    for col in ("p_value", "p_adjusted", "mean_quality", "mean_consensus"):
        assert _rel(out[col], g[f"want_{col}"]).max() < 1e-12, col   # 1e-12 relative class


def test_call_stage_binomial_and_ragged(ctx, native):
    rng = np.random.default_rng(9)
    lens = np.array([1, 2, 255, 256, 257, 5000, 1, 40000, 3])
    so = np.concatenate([[0], np.cumsum(lens)])
    R = so[-1]
    iv = (rng.uniform(size=R) < 0.002).astype(np.uint8)
    q = rng.normal(25, 3, R)
    a = rng.uniform(0.7, 1, R)
    for tt, name in ((native.TEST_POISSON, "poisson"), (native.TEST_BINOMIAL, "binomial")):
        out = ctx.call(so, iv, q, a, 1.3e-3, tt, 0.05)
        want = port.call_segments(so, iv, q, a, 1.3e-3, name, 0.05)
        assert (out["n_variants"] == want["n_variants"]).all() and (out["n_total"] == want["n_total"]).all()
        assert (out["significant"] == want["significant"]).all()
        for col in ("p_value", "p_adjusted", "mean_quality", "mean_consensus", "fold_change", "expected_variants"):
            assert _rel(out[col], want[col]).max() < 1e-12, (name, col)
    # expected == 0 -> p = 1, fold_change inf / 1.0 (call.py:190-193)
    out = ctx.call(so, iv, None, None, 0.0, native.TEST_POISSON, 0.05)
    assert (out["p_value"] == 1.0).all()
    assert np.isinf(out["fold_change"][out["n_variants"] > 0]).all()
    assert (out["fold_change"][out["n_variants"] == 0] == 1.0).all()
    # determinism: fp64 means are bit-identical run to run
    o1 = ctx.call(so, iv, q, a, 1.3e-3, native.TEST_POISSON, 0.05)
    o2 = ctx.call(so, iv, q, a, 1.3e-3, native.TEST_POISSON, 0.05)
    assert (o1["mean_quality"] == o2["mean_quality"]).all() and (o1["p_adjusted"] == o2["p_adjusted"]).all()


# ------------------------------------------------------------------------------ collapse
def _assert_tables_equal(fam, grp, wfam, wgrp):
    for k in ("key", "size", "n_alt", "consensus", "qsum", "n_best", "flags"):
        assert (fam[k] == wfam[k]).all(), k
    for k in ("group", "family_count", "consensus_count", "alt_count", "total_count"):
        assert (grp[k] == wgrp[k]).all(), k


def test_collapse_ga_firstfit_majority_vs_reference(ctx, golden, native):
    g = golden("ga_reads.npz")
    P = _params(native, min_family_size=2, mode=native.MODE_MAJORITY, cluster=native.CLUSTER_FIRSTFIT, max_distance=1, umi_len=8)
    fam, grp = ctx.collapse_reads(g["keys"], g["payload"], P)
    kept = (fam["flags"] & native.FAM_KEPT) != 0
    assert (fam["key"][kept] == g["want_fam_key"]).all()          # representative UMI, reference order
    assert (fam["size"][kept] == g["want_fam_size"]).all()
    assert (fam["n_alt"][kept] == g["want_fam_alt"]).all()
    assert (fam["consensus"][kept] == g["want_fam_consensus"]).all()
    locus = g["group_locus"][grp["group"]]
    for name, col in (("want_locus_alt", "alt_count"), ("want_locus_total", "total_count"), ("want_locus_families", "family_count")):
        agg = np.zeros(locus.max() + 1, np.int64)
        np.add.at(agg, locus, grp[col].astype(np.int64))
        assert (agg[g["want_locus"]] == g[name]).all(), name
    wfam, wgrp = port.collapse_firstfit(g["keys"], g["payload"], 1, 2, 1 << 30, 1)
    _assert_tables_equal(fam, grp, wfam, wgrp)


def test_collapse_gc_counts_vs_reference(ctx, golden, native):
    g = golden("gc_reads.npz")
    fam, grp = ctx.collapse_reads(g["keys"], g["payload"], _params(native))
    order = np.argsort(g["want_fam_key"], kind="stable")
    assert (fam["key"] == g["want_fam_key"][order]).all()
    assert (fam["size"] == g["want_size"][order]).all()
    assert (fam["n_alt"] == g["want_alt"][order]).all()
    assert (fam["size"] - fam["n_alt"] == g["want_ref"][order]).all()


def test_collapse_gb_weighted_vs_reference(ctx, golden, native):
    g = golden("gb_reads.npz")
    mn, mx, qthr = [int(v) for v in g["params"]]
    P = _params(native, min_family_size=mn, max_family_size=mx, quality_threshold=qthr,
                consensus_threshold=float(g["consensus_threshold"]), mode=native.MODE_WEIGHTED)
    fam, grp = ctx.collapse_reads(g["keys"], g["payload"], P)
    order = np.argsort(g["want_fam_key"], kind="stable")
    assert (fam["key"] == g["want_fam_key"][order]).all()
    assert (fam["size"] == g["want_fam_size"][order]).all()
    has = (fam["flags"] & native.FAM_HAS_CONSENSUS) != 0
    want_allele = g["want_fam_allele"][order]
    assert (has == (want_allele >= 0)).all()                      # same >= threshold decisions, bit-exact
    assert (fam["consensus"][has] == want_allele[has]).all()
    assert (fam["qsum"][has] / fam["n_best"][has] == g["want_fam_quality"][order][has]).all()
    assert int(((fam["flags"] & native.FAM_KEPT) != 0).sum()) == int(g["want_kept"])
    assert (grp["consensus_count"] == g["want_consensus_count"][grp["group"]]).all()
    wfam, wgrp = port.collapse_exact(g["keys"], g["payload"], 0, mn, mx, qthr, float(g["consensus_threshold"]))
    _assert_tables_equal(fam, grp, wfam, wgrp)


# ------------------------------------------------------------------------------ ranking kernels under stress
@pytest.mark.parametrize("pattern", ["all_equal", "two_valued", "sorted", "reverse", "one_hot_digit", "random", "runs"])
def test_ranking_kernels_on_adversarial_digit_distributions(ctx, pattern):
    """The ballot / leader-atomic ranking of the scatter kernels (staged pass 1, direct passes 2-3, the counting and
    cell-scan kernels) is where a race would hide, and the pool refuses compute-sanitizer's racecheck: every pattern
    that maximises same-digit collisions inside a warp (all lanes one digit, two digits, long runs, already sorted,
    reverse order) is sorted 40 times on the plan's cell-major layout -- cells of 1, 2, 7 and 61 tiles -- and must come
    back, every time, as the stable per-cell sort of the input."""
    rng = np.random.default_rng(len(pattern))
    tiles = np.array([1, 2, 7, 61, 1, 3])
    ts = np.concatenate([[0], np.cumsum(tiles)]).astype(np.int64)
    n = int(ts[-1]) * 4096
    if pattern == "all_equal":
        umi = np.full(n, 0xABCDEF, np.uint64)
    elif pattern == "two_valued":
        umi = np.where(rng.random(n) < 0.5, np.uint64(0x010203), np.uint64(0xFFFEFD))
    elif pattern == "sorted":
        umi = (np.arange(n, dtype=np.uint64) // np.uint64(3)) & np.uint64(0xFFFFFF)
    elif pattern == "reverse":
        umi = ((np.uint64(n) - np.arange(n, dtype=np.uint64)) // np.uint64(5)) & np.uint64(0xFFFFFF)
    elif pattern == "one_hot_digit":                      # one byte varies at a time, the other two constant
        umi = (rng.integers(0, 256, n, dtype=np.uint64) << np.uint64(8)) | np.uint64(0x550033)
    elif pattern == "runs":                               # runs of 32-200 equal keys: whole warps on one digit
        lens = rng.integers(32, 200, n // 32)
        umi = np.repeat(rng.integers(0, 1 << 24, len(lens), dtype=np.uint64), lens)[:n]
    else:
        umi = rng.integers(0, 1 << 24, n, dtype=np.uint64)
    rec = (np.arange(n, dtype=np.uint64) << np.uint64(32)) | umi           # payload = input position: stability is visible
    want = rec.copy()
    for c in range(len(tiles)):
        lo, hi = int(ts[c]) * 4096, int(ts[c + 1]) * 4096
        want[lo:hi] = rec[lo:hi][np.argsort(umi[lo:hi], kind="stable")]
    for _ in range(40):
        got = ctx.sort_packed_cells(rec, ts, 24)
        assert (got == want).all()
    # the general (key + payload) scatter kernel and the in-block group sort on the same patterns
    keys = ((np.arange(n, dtype=np.uint64) % np.uint64(37)) << np.uint64(32)) | umi
    vals = np.arange(n, dtype=np.uint32)
    for _ in range(10):
        sk, sv = ctx.sort_pairs(keys, vals)
        order = np.argsort(keys, kind="stable")
        assert (sk == keys[order]).all() and (sv == vals[order]).all()


def _panel_reads(rng, n, n_groups, umis_per_group, umi_bits=24):
    """External reads of a panel: dense group ids, a pool of UMIs per group, arbitrary order."""
    g = rng.integers(0, n_groups, n, dtype=np.uint64)
    pool = rng.integers(0, 1 << umi_bits, (n_groups, umis_per_group), dtype=np.uint64)
    umi = pool[g.astype(np.int64), rng.integers(0, umis_per_group, n)]
    payload = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    q = rng.choice([5, 10, 20, 20, 30, 30, 30, 37, 40], n).astype(np.uint32)
    return (g << np.uint64(32)) | umi, (payload & np.uint32(0xFFFFFF03)) | (q << np.uint32(2))


def test_collapse_external_reads_group_major_path(ctx, native, monkeypatch):
    """pmrd_collapse_reads on a panel of loci in arbitrary order (dense group ids, no group above 4096 reads) takes the
    group-major path: stable partition by group, then one block sorts and votes a group in shared memory.  Tables
    against the C restatement of the reference (umi.py:97-203), and -- at a size the oracle cannot reach in seconds --
    against the full-key-sort path (PMRD_GROUP_FAST=0), for every consensus mode."""
    from oracle import cport

    rng = np.random.default_rng(77)
    keys, pl = _panel_reads(rng, 300_000, 600, 120)
    P = _params(native, min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6, mode=native.MODE_WEIGHTED)
    fam, grp = ctx.collapse_reads(keys, pl, P)
    want = cport.collapse_weighted(keys, pl, 3, 1000, 20, 0.6)
    for k in want:
        assert (fam[k] == want[k]).all(), k
    assert len(grp["group"]) == 600 and (grp["family_count"] == np.bincount((want["key"][(want["flags"] & 2) != 0] >> np.uint64(32)).astype(np.int64), minlength=600)).all()
    # group table only (no family row is written): the same table, the family count still reported
    info, grp_only = ctx.collapse_reads(keys, pl, P, want_families=False)
    assert info["n_families"] == len(want["key"])
    for k in grp:
        assert (grp_only[k] == grp[k]).all(), k
    og = cport.collapse_groups(keys, pl, 600, 3, 1000, 20, 0.6)
    assert (grp_only["family_count"] == og["family_count"]).all() and (grp_only["consensus_count"] == og["consensus_count"]).all()
    # ... and when the group-major path does not apply (a group of more than 4096 reads) the fallback gives it too
    kbig = keys.copy()
    kbig[:6000] = (np.uint64(5) << np.uint64(32)) | (kbig[:6000] & np.uint64(0xFFFFFFFF))
    f_all, g_all = ctx.collapse_reads(kbig, pl, P)
    info2, g_only2 = ctx.collapse_reads(kbig, pl, P, want_families=False)
    wf, wg = port.collapse_exact(kbig, pl, 0, 3, 1000, 20, 0.6)
    _assert_tables_equal(f_all, g_all, wf, wg)
    assert info2["n_families"] == len(wf["key"])
    for k in g_all:
        assert (g_only2[k] == g_all[k]).all(), k
    # 16-mer UMIs (32 bits, four ranked rounds), a UMI of all T (= the filler word), groups of one read, empty ids
    keys2, pl2 = _panel_reads(rng, 40_000, 3000, 4, umi_bits=32)
    keys2[:50] |= np.uint64(0xFFFFFFFF)
    keys2 = np.where((keys2 >> np.uint64(32)) % np.uint64(7) == 3, keys2 + (np.uint64(1) << np.uint64(32)), keys2)   # leave some ids empty
    for mode in (0, 1, 2):
        P2 = _params(native, min_family_size=2, max_family_size=9, quality_threshold=20, consensus_threshold=0.4, mode=mode, umi_len=16)
        fam, grp = ctx.collapse_reads(keys2, pl2, P2)
        wfam, wgrp = port.collapse_exact(keys2, pl2, mode, 2, 9, 20, 0.4)
        _assert_tables_equal(fam, grp, wfam, wgrp)
    # 1e5 loci (17 group bits): the partition takes two 9-bit digit passes (512 bins) instead of three 8-bit ones
    keys4, pl4 = _panel_reads(rng, 600_000, 100_000, 3)
    fa, ga = ctx.collapse_reads(keys4, pl4, P)
    monkeypatch.setenv("PMRD_PARTITION_BITS", "8")
    fb, gb = ctx.collapse_reads(keys4, pl4, P)
    monkeypatch.delenv("PMRD_PARTITION_BITS")
    _assert_tables_equal(fa, ga, fb, gb)
    og4 = cport.collapse_groups(keys4, pl4, 100_000, 3, 1000, 20, 0.6)
    assert (ga["family_count"] == og4["family_count"][ga["group"]]).all() and (ga["consensus_count"] == og4["consensus_count"][ga["group"]]).all()
    assert len(fa["key"]) == og4["total_families"]
    # the counts-only form makes one ranked round fewer and leaves the high UMI byte to the vote.  Adversarial input for
    # that: UMIs that share their low 16 bits by the dozen (3 low values x 80 high bytes per locus), so EVERY run is mixed
    keys5, pl5 = _panel_reads(rng, 200_000, 500, 80)
    u5 = keys5 & np.uint64(0xFFFFFFFF)
    keys5 = (keys5 & np.uint64(0xFFFFFFFF00000000)) | ((u5 % np.uint64(80)) << np.uint64(16)) | (u5 % np.uint64(3))
    for mode in (0, 1, 2):
        P5 = _params(native, min_family_size=2, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6, mode=mode)
        f5, g5 = ctx.collapse_reads(keys5, pl5, P5)
        i5, go5 = ctx.collapse_reads(keys5, pl5, P5, want_families=False)
        monkeypatch.setenv("PMRD_GROUP_PARTIAL", "0")
        i6, go6 = ctx.collapse_reads(keys5, pl5, P5, want_families=False)
        monkeypatch.delenv("PMRD_GROUP_PARTIAL")
        assert i5["n_families"] == len(f5["key"]) == i6["n_families"]
        for k in g5:
            assert (go5[k] == g5[k]).all() and (go6[k] == g5[k]).all(), (mode, k)
    og5 = cport.collapse_groups(keys5, pl5, 500, 2, 1000, 20, 0.6)
    _, go5 = ctx.collapse_reads(keys5, pl5, _params(native, min_family_size=2, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6, mode=native.MODE_WEIGHTED), want_families=False)
    assert (go5["family_count"] == og5["family_count"]).all() and (go5["consensus_count"] == og5["consensus_count"]).all()
    # 3e6 reads over 20 000 loci: both paths, every mode
    keys3, pl3 = _panel_reads(rng, 3_000_000, 20_000, 40)
    for mode in (0, 1, 2):
        P3 = _params(native, min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6, mode=mode)
        monkeypatch.setenv("PMRD_GROUP_FAST", "1")
        fa, ga = ctx.collapse_reads(keys3, pl3, P3)
        monkeypatch.setenv("PMRD_GROUP_FAST", "0")
        fb, gb = ctx.collapse_reads(keys3, pl3, P3)
        _assert_tables_equal(fa, ga, fb, gb)
        assert len(fa["key"]) > 700_000 and (np.diff(fa["key"].astype(np.int64)) > 0).all()       # sorted by (group, UMI), no duplicates


def test_collapse_reads_dev_on_device_arrays_odd_length_and_alignment_contract(ctx, native):
    """pmrd_collapse_reads_dev on a caller's device arrays: an odd read count (the tail of the wide run-boundary loads) gives
    the oracle's table; key arrays that are not 16-byte aligned (a view one element into an allocation) are refused loudly,
    as include/pmrd.h states for pmrd_collapse_reads_dev, instead of being read with misaligned vector loads."""
    torch = pytest.importorskip("torch")
    from oracle import cport

    rng = np.random.default_rng(5)
    keys, pl = _panel_reads(rng, 200_001, 900, 60)
    n = len(keys)
    P = _params(native, min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6, mode=native.MODE_WEIGHTED)
    dev = torch.device("cuda", 0)
    kbuf, ktmp = torch.zeros(n + 1, dtype=torch.int64, device=dev), torch.zeros(n + 1, dtype=torch.int64, device=dev)
    pbuf, ptmp = torch.zeros(n + 1, dtype=torch.int32, device=dev), torch.zeros(n + 1, dtype=torch.int32, device=dev)
    gcap = 900 + 16
    grp_t = {"group": torch.zeros(gcap, dtype=torch.int32, device=dev), "family_count": torch.zeros(gcap, dtype=torch.int32, device=dev),
             "consensus_count": torch.zeros(gcap * 4, dtype=torch.int32, device=dev), "alt_count": torch.zeros(gcap, dtype=torch.int64, device=dev),
             "total_count": torch.zeros(gcap, dtype=torch.int64, device=dev)}
    grp_p = {k: v.data_ptr() for k, v in grp_t.items()}
    kbuf[:n].copy_(torch.from_numpy(keys.view(np.int64)))
    pbuf[:n].copy_(torch.from_numpy(pl.view(np.int32)))
    torch.cuda.synchronize()
    nf, ng = ctx.collapse_reads_dev(kbuf.data_ptr(), pbuf.data_ptr(), ktmp.data_ptr(), ptmp.data_ptr(), n, P, None, 0, grp_p, gcap)
    torch.cuda.synchronize()
    og = cport.collapse_groups(keys, pl, 900, 3, 1000, 20, 0.6)
    assert nf == og["total_families"] and ng == 900
    assert (grp_t["family_count"][:ng].cpu().numpy().astype(np.uint32) == og["family_count"]).all()
    assert (grp_t["consensus_count"][:ng * 4].cpu().numpy().astype(np.uint32).reshape(ng, 4) == og["consensus_count"]).all()
    assert kbuf[1:].data_ptr() % 16 == 8
    with pytest.raises(native.NativeError, match="16-byte aligned"):
        ctx.collapse_reads_dev(kbuf[1:].data_ptr(), pbuf.data_ptr(), ktmp[1:].data_ptr(), ptmp.data_ptr(), n, P, None, 0, grp_p, gcap)


@pytest.mark.parametrize("log2w", [1, 2, 3])
def test_owner_cut_kernels_emulating_the_ranks_on_one_gpu(ctx, native, log2w):
    """The device steps of the multi-GPU exchange (csrc/exchange.cu) with all "ranks" on ONE GPU: every rank's reads are cut
    by owner (group mod 2^k) -- the partitioned-copy form (pmrd_owner_partition_dev) and the store-into-the-owner form
    (pmrd_owner_tiles_dev + pmrd_owner_scatter_peer_dev, here with ordinary device arrays as the "peer" arrays) -- and each
    owner must end up with exactly the concatenation, in rank order, of the ranks' reads it owns in arrival order, group ids
    divided by 2^k.  (The real thing over NCCL / IPC: tests/test_gpu_distributed.py.)"""
    torch = pytest.importorskip("torch")
    world = 1 << log2w
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(40 + log2w)
    parts = [_panel_reads(rng, 50_000 + 4097 * r + (r == 1), 3000, 40) for r in range(world)]       # odd sizes, partial tiles
    parts[world - 1] = (parts[world - 1][0][:5], parts[world - 1][1][:5])                             # one rank with 5 reads
    mask = np.uint64(world - 1)
    want_k, want_p = [], []
    for o in range(world):
        ks = [k[((k >> np.uint64(32)) & mask) == o] for k, _ in parts]
        ps = [p[((k >> np.uint64(32)) & mask) == o] for k, p in parts]
        kk = np.concatenate(ks)
        want_k.append((((kk >> np.uint64(32)) >> np.uint64(log2w)) << np.uint64(32)) | (kk & np.uint64(0xFFFFFFFF)))
        want_p.append(np.concatenate(ps))
    counts = np.array([[int((((k >> np.uint64(32)) & mask) == o).sum()) for o in range(world)] for k, _ in parts])   # [rank][owner]
    # ---- partitioned-copy form
    got = [[] for _ in range(world)]
    for r, (k, p) in enumerate(parts):
        n = len(k)
        dk, dp = torch.from_numpy(k.view(np.int64)).to(dev), torch.from_numpy(p.view(np.int32)).to(dev)
        dkt, dpt = torch.empty_like(dk), torch.empty_like(dp)
        torch.cuda.synchronize()
        c, in_tmp = ctx.owner_partition_dev(dk.data_ptr(), dp.data_ptr(), dkt.data_ptr(), dpt.data_ptr(), n, log2w)
        assert c == counts[r].tolist()
        sk, sp = (dkt, dpt) if in_tmp else (dk, dp)
        ctx.owner_rekey_dev(sk.data_ptr(), n, log2w)
        ctx.sync()
        hk, hp = sk.cpu().numpy().view(np.uint64), sp.cpu().numpy().view(np.uint32)
        at = 0
        for o in range(world):
            got[o].append((hk[at:at + c[o]], hp[at:at + c[o]]))
            at += c[o]
    for o in range(world):
        assert (np.concatenate([a for a, _ in got[o]]) == want_k[o]).all() and (np.concatenate([b for _, b in got[o]]) == want_p[o]).all(), o
    # ---- store-into-the-owner form: "peer" arrays = this GPU's own arrays
    rk = [torch.full((int(counts[:, o].sum()) + 8,), -1, dtype=torch.int64, device=dev) for o in range(world)]
    rp = [torch.full((int(counts[:, o].sum()) + 8,), -1, dtype=torch.int32, device=dev) for o in range(world)]
    for r, (k, p) in enumerate(parts):
        n = len(k)
        dk, dp = torch.from_numpy(k.view(np.int64)).to(dev), torch.from_numpy(p.view(np.int32)).to(dev)
        tile_off = torch.empty(world * ((n + 4095) // 4096), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        c = ctx.owner_tiles_dev(dk.data_ptr(), n, log2w, tile_off.data_ptr())
        assert c == counts[r].tolist()
        base = counts[:r].sum(axis=0).tolist()
        ctx.owner_scatter_peer_dev(dk.data_ptr(), dp.data_ptr(), n, log2w, tile_off.data_ptr(), [t.data_ptr() for t in rk], [t.data_ptr() for t in rp], base)
        ctx.sync()
    for o in range(world):
        m = int(counts[:, o].sum())
        assert (rk[o][:m].cpu().numpy().view(np.uint64) == want_k[o]).all() and (rp[o][:m].cpu().numpy().view(np.uint32) == want_p[o]).all(), o
        assert (rk[o][m:] == -1).all() and (rp[o][m:] == -1).all()                    # nothing written past an owner's reads


@pytest.mark.parametrize("d", [1, 2])
def test_collapse_gb_greedy_clusters_vs_reference(ctx, golden, native, d):
    """K6 greedy (max_edit_distance > 0, the reference's default): the device clusters equal the reference's
    cluster_umis_greedy (umi_clustering.py:59-101) cluster for cluster and in its order; consensus per cluster
    over the reads in input order (umi.py:137-140) bit-exact; also equal to the port on every column."""
    from test_oracle_golden import check_greedy_tables_against_reference

    g = golden("gb_greedy.npz")
    mn, mx, qthr = [int(v) for v in g["params"]]
    P = _params(native, min_family_size=mn, max_family_size=mx, quality_threshold=qthr, consensus_threshold=float(g["consensus_threshold"]),
                mode=native.MODE_WEIGHTED, cluster=native.CLUSTER_GREEDY, max_distance=d, umi_len=12)
    fam, grp = ctx.collapse_reads(g["keys"], g["payload"], P)
    check_greedy_tables_against_reference(fam, g, d, native.FAM_HAS_CONSENSUS)
    wfam, wgrp = port.collapse_greedy(g["keys"], g["payload"], 0, mn, mx, d, qthr, float(g["consensus_threshold"]))
    _assert_tables_equal(fam, grp, wfam, wgrp)


@pytest.mark.parametrize("n,n_groups,umi_space", [(1, 1, 4), (7, 2, 3), (5000, 3, 1 << 9), (40_000, 50, 1 << 12), (30_000, 1, 1 << 13)])
def test_collapse_greedy_vs_oracle_random(ctx, native, n, n_groups, umi_space):
    """Random reads, dense barcode neighbourhoods (equal-count ties, chains the greedy rule cuts, one group with
    thousands of unique UMIs): every mode, distances 0-2, against the port."""
    rng = np.random.default_rng(n + n_groups)
    groups = rng.integers(0, n_groups, n, dtype=np.uint64) * np.uint64(17)
    umis = np.where(rng.random(n) < 0.5, rng.zipf(1.6, n).astype(np.uint64) % np.uint64(umi_space),     # skewed counts ...
                    rng.integers(0, umi_space, n, dtype=np.uint64))                                  # ... + many singletons
    keys = (groups << np.uint64(32)) | umis
    payload = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    q = rng.choice([5, 10, 20, 20, 30, 30, 30, 37, 40], n).astype(np.uint32)
    wpay = (payload & np.uint32(0xFFFFFF03)) | (q << np.uint32(2))
    for mode, d in ((0, 1), (0, 2), (1, 1), (2, 0)):
        pl = wpay if mode == 0 else payload
        P = _params(native, min_family_size=2, max_family_size=50, quality_threshold=20, consensus_threshold=0.6, mode=mode,
                    cluster=native.CLUSTER_GREEDY, max_distance=d, umi_len=7)
        fam, grp = ctx.collapse_reads(keys, pl, P)
        wfam, wgrp = port.collapse_greedy(keys, pl, mode, 2, 50, d, 20, 0.6)
        _assert_tables_equal(fam, grp, wfam, wgrp)


@pytest.mark.parametrize("n", [0, 1, 5, 4097, 60_000])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_collapse_exact_vs_oracle_random(ctx, native, n, mode):
    rng = np.random.default_rng(100 * n + mode)
    groups = rng.integers(0, 9, n, dtype=np.uint64) * np.uint64(131)          # sparse group ids
    umis = rng.integers(0, max(2, n // 4), n, dtype=np.uint64) * np.uint64(2654435761) % np.uint64(1 << 24)
    keys = (groups << np.uint64(32)) | umis
    payload = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    if mode == 0:   # qualities in 0..63 with equal-quality ties being common
        q = rng.choice([5, 10, 20, 20, 30, 30, 30, 37, 40, 63], n).astype(np.uint32)
        payload = (payload & np.uint32(0xFFFFFF03)) | (q << np.uint32(2))
    P = _params(native, min_family_size=3, max_family_size=6, quality_threshold=20, consensus_threshold=0.6, mode=mode)
    fam, grp = ctx.collapse_reads(keys, payload, P)
    wfam, wgrp = port.collapse_exact(keys, payload, mode, 3, 6, 20, 0.6)
    _assert_tables_equal(fam, grp, wfam, wgrp)


def test_collapse_one_giant_family_and_firstfit_random(ctx, native):
    n = 20_000
    keys = np.full(n, (5 << 32) | 77, dtype=np.uint64)
    payload = np.random.default_rng(1).integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    P = _params(native, mode=native.MODE_WEIGHTED, quality_threshold=20, consensus_threshold=0.3, max_family_size=1000)
    fam, grp = ctx.collapse_reads(keys, payload, P)
    wfam, wgrp = port.collapse_exact(keys, payload, 0, 1, 1000, 20, 0.3)
    _assert_tables_equal(fam, grp, wfam, wgrp)
    assert fam["flags"][0] & native.FAM_OVERSIZE
    rng = np.random.default_rng(2)
    n = 6000
    keys = (rng.integers(0, 40, n, dtype=np.uint64) << np.uint64(32)) | rng.integers(0, 1 << 10, n, dtype=np.uint64)
    payload = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    for d in (0, 1, 2):
        P = _params(native, min_family_size=2, mode=native.MODE_MAJORITY, cluster=native.CLUSTER_FIRSTFIT, max_distance=d, umi_len=5)
        fam, grp = ctx.collapse_reads(keys, payload, P)
        wfam, wgrp = port.collapse_firstfit(keys, payload, 1, 2, 1 << 30, d)
        _assert_tables_equal(fam, grp, wfam, wgrp)


def test_umi_hamming(ctx):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 1 << 24, 5000, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 1 << 24, 5000, dtype=np.uint64).astype(np.uint32)
    b[:100] = a[:100]
    assert (ctx.umi_hamming(a, b, 12) == port.umi_hamming(a, b)).all()
