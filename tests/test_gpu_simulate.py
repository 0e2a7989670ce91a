"""GPU tests of the Philox generators and of the fused pipeline.

Parity class for anything that consumed the reference's PCG64 stream is DISTRIBUTIONAL
(north_star): chi-square / KS at alpha = 0.01 against the closed forms of the
reference's draws and against samples of the reference itself
(tests/golden/gd_dist.npz).  Everything downstream of the generated reads is checked
bit-exactly against the oracle run on the very same reads."""
import numpy as np
import pytest
from scipy import stats

from oracle import port

pytestmark = pytest.mark.gpu
ALPHA = 0.01


def _umi(native, **kw):
    d = dict(min_family_size=3, max_family_size=1000, quality_threshold=20, consensus_threshold=0.6,
             mode=native.MODE_WEIGHTED, cluster=native.CLUSTER_EXACT, max_distance=0, umi_len=12)
    d.update(kw)
    return native.UmiParams(**d)


def _sp(native, **kw):
    d = dict(mean_family_size=5.0, family_model=0, max_family_size=1000, contamination_rate=0.0, hop_rate=0.0,
             collision_rate=0.0, shuffle=1, chunk_log2=0)
    d.update(kw)
    return native.ReadSimParams(**d)


def _chi2_pvalue(obs, exp):
    obs, exp = np.asarray(obs, float), np.asarray(exp, float)
    keep = exp >= 5
    o = np.append(obs[keep], obs[~keep].sum())
    e = np.append(exp[keep], exp[~keep].sum())
    if e[-1] == 0:
        o, e = o[:-1], e[:-1]
    return stats.chisquare(o, e * o.sum() / e.sum()).pvalue


# ------------------------------------------------------------------------------ K1 cells
def test_simulate_cells_distributions(ctx):
    c = 200_000
    cid = np.arange(c)
    af = np.tile([0.01, 0.001, 0.0001, 0.05], c // 4)
    depth = np.tile([1000, 5000, 20000, 50000, 100000], c // 5)
    out = ctx.simulate_cells(cid, af, depth, 1000)
    # simulate.py:144 U(1e-5,1e-3); :147 clip(Poisson(5),1,max); :161 clip(N(25,5),10,40)
    assert stats.kstest(out["background_rate"], stats.uniform(1e-5, 1e-3 - 1e-5).cdf).pvalue > ALPHA
    ks = np.arange(0, 25)
    exp = stats.poisson.pmf(ks, 5) * c
    exp[1] += exp[0]; exp[0] = 0
    assert _chi2_pvalue(np.bincount(out["mean_family_size"], minlength=25)[:25], exp) > ALPHA
    q = out["mean_quality"]
    inner = (q > 10) & (q < 40)
    assert abs((q == 10).mean() - stats.norm.cdf(-3)) < 3e-4 and abs((q == 40).mean() - stats.norm.sf(3)) < 3e-4
    tn = stats.truncnorm(-3, 3, loc=25, scale=5)
    assert stats.kstest(q[inner], tn.cdf).pvalue > ALPHA
    # simulate.py:152 Binom(depth, af): probability-integral transform with a randomised tie-break
    rng = np.random.default_rng(0)
    for a, d in ((0.0001, 1000), (0.01, 5000), (0.05, 100000), (0.001, 20000)):
        sel = (af == a) & (depth == d)
        k = out["n_true_variants"][sel]
        if sel.sum() == 0:
            continue
        u = stats.binom.cdf(k - 1, d, a) + rng.uniform(size=k.size) * stats.binom.pmf(k, d, a)
        assert stats.kstest(u, "uniform").pvalue > ALPHA, (a, d)
    # :155 Binom(depth - n_true, bg): mean check at 5 sigma
    nfp, bg, nt = out["n_false_positives"], out["background_rate"], out["n_true_variants"]
    z = (nfp - (depth - nt) * bg) / np.sqrt((depth - nt) * bg * (1 - bg))
    assert abs(z.mean()) < 5 / np.sqrt(c) and abs(z.std() - 1) < 0.02
    # counter-based: a cell's draw does not depend on what else is in the batch
    sub = ctx.simulate_cells(cid[1000:1010], af[1000:1010], depth[1000:1010], 1000)
    for k_ in out:
        assert (sub[k_] == out[k_][1000:1010]).all()


# ------------------------------------------------------------------------------ K2 G-D families
@pytest.mark.parametrize("variant", [1, 2])
def test_gd_families_match_reference_distributions(ctx, golden, native, variant):
    g = golden("gd_dist.npz")
    tag = f"v{variant}"
    af, bg, mfs = g[f"{tag}_cell_af"], g[f"{tag}_cell_bg"], g[f"{tag}_cell_mfs"]
    c = af.shape[0]
    reps = 12                                 # 12 x the reference's 24 cells, different cell ids
    cid = np.arange(c * reps) + 1000
    AF, BG, MFS = np.tile(af, reps), np.tile(bg, reps), np.tile(mfs, reps)
    NF = np.full(c * reps, 4000)
    P = _umi(native)
    offs, fam = ctx.collapse_families(cid, AF, NF, BG, MFS, P, variant=variant)
    assert offs[-1] == fam["cell"].shape[0] and (np.diff(offs) >= 0).all()
    cnt_only, _ = ctx.collapse_families(cid, AF, NF, BG, MFS, P, variant=variant, count_only=True)
    assert (cnt_only == offs).all()
    assert (fam["family_size"] >= 3).all()    # collapse.py:85-86
    # rows are cell-major with ascending family ids (gaps where skipped)
    assert (np.diff(fam["cell"]) >= 0).all()
    same = np.diff(fam["cell"]) == 0
    assert (np.diff(fam["family_id"])[same] > 0).all()
    ref_fs, ref_q, ref_a, ref_fl = g[f"{tag}_family_size"], g[f"{tag}_quality"], g[f"{tag}_agreement"], g[f"{tag}_flags"]
    ref_cell = g[f"{tag}_sample"]
    # two-sample tests against the reference's own draws, per AF level (quality depends on AF in v2)
    for a in np.unique(af):
        m_ref = af[ref_cell] == a
        m_new = AF[fam["cell"]] == a
        assert stats.ks_2samp(fam["quality_score"][m_new], ref_q[m_ref]).pvalue > ALPHA, ("quality", a)
        assert stats.ks_2samp(fam["consensus_agreement"][m_new], ref_a[m_ref]).pvalue > ALPHA, ("agreement", a)
    # family size: per lambda group, chi-square new-vs-reference histogram
    lam = np.maximum(3, mfs) if variant == 2 else np.full(c, 5.0)
    for L in np.unique(lam):
        m_ref = lam[ref_cell] == L
        m_new = np.tile(lam, reps)[fam["cell"]] == L
        hr = np.bincount(ref_fs[m_ref], minlength=40)[:40].astype(float)
        hn = np.bincount(fam["family_size"][m_new], minlength=40)[:40].astype(float)
        keep = (hr + hn) >= 10
        tab = np.vstack([np.append(hr[keep], hr[~keep].sum()), np.append(hn[keep], hn[~keep].sum())])
        tab = tab[:, tab.sum(0) > 0]
        assert stats.chi2_contingency(tab)[1] > ALPHA, ("family_size", L)
        # kept fraction = P[clip(Poisson(L)) >= 3]
        kept_new = m_new.sum() / (4000 * reps * (lam == L).sum())
        assert abs(kept_new - stats.poisson.sf(2, L)) < 4e-3
    # flags: pass rates and is_variant rate (binomial z-test vs the reference's rate)
    for bit, name in ((0, "passes_quality"), (1, "passes_consensus")):
        pn, pr = ((fam["flags"] >> bit) & 1).mean(), ((ref_fl >> bit) & 1).mean()
        se = np.sqrt(pr * (1 - pr) * (1 / len(ref_fl) + 1 / len(fam["flags"])) + 1e-12)
        assert abs(pn - pr) < 4 * se + 1e-9, name
    for a in np.unique(af):
        m_ref, m_new = af[ref_cell] == a, AF[fam["cell"]] == a
        kr, kn = ((ref_fl[m_ref] >> 2) & 1).sum(), ((fam["flags"][m_new] >> 2) & 1).sum()
        tab = np.array([[kr, m_ref.sum() - kr], [kn, m_new.sum() - kn]])
        if (tab.sum(0) > 0).all():
            assert stats.fisher_exact(tab)[1] > ALPHA / 3, ("is_variant", a)   # AF acts only through AF > 0.01 (App. A #12)


# ------------------------------------------------------------------------------ K10 error model
def test_error_model_rates_and_bootstrap(ctx):
    n_neg, n_var = 50523, 31
    rate, lo, hi = ctx.error_model(n_neg, n_var, 100, stream_id=5)
    vr = n_var / n_neg
    mod = rate / vr
    assert (mod >= 0.5).all() and (mod <= 2.0).all() and mod.std() > 0.2          # error_model.py:155
    # bootstrap percentiles bracket the binomial-equivalent 2.5 / 97.5 quantiles (100 resamples: loose)
    qlo, qhi = stats.binom.ppf([0.025, 0.975], n_neg, vr) / n_neg
    assert (lo <= rate).all() and (hi >= rate * 0.999).all()
    assert abs(np.median(lo / mod) - qlo) < 0.35 * vr and abs(np.median(hi / mod) - qhi) < 0.35 * vr
    # many independent fits: modifiers are U(0.5, 2)
    mods = np.concatenate([ctx.error_model(n_neg, n_var, 8, stream_id=s)[0] / vr for s in range(40)])
    assert stats.kstest(mods, stats.uniform(0.5, 1.5).cdf).pvalue > ALPHA
    # fallbacks (error_model.py:157-159,175-177)
    r, l, h = ctx.error_model(5, 1, 100, 0)
    assert np.allclose(l, r * 0.5) and np.allclose(h, r * 2.0)
    r, l, h = ctx.error_model(0, 0, 100, 0)
    assert (r >= 1e-5).all() and (r <= 1e-3).all()


# ------------------------------------------------------------------------------ read generator
def test_read_generator_distributions(ctx, native):
    c = 64
    cid = np.arange(c) + 17
    af = np.tile([0.0, 0.05, 0.3, 0.8], c // 4)
    nf = np.full(c, 20000)
    for model in (0, 1):
        sp = _sp(native, family_model=model, shuffle=0)
        offs = ctx.simulate_reads_count(cid, nf, sp)
        keys, pl = ctx.simulate_reads(cid, af, nf, sp)
        assert keys.shape[0] == offs[-1]
        # family-major, unshuffled: runs of equal key are the families
        change = np.flatnonzero(np.diff(keys.astype(np.int64)) != 0) + 1
        starts = np.concatenate([[0], change])
        sizes = np.diff(np.concatenate([starts, [len(keys)]]))
        # (adjacent families with the same UMI in the same cell merge: negligible at 24-bit UMIs)
        ks = np.arange(1, 60)
        if model == 0:   # io.py:71-77  NegBinom(2, 2/(2+mean)) + 1
            exp = stats.nbinom.pmf(ks - 1, 2, 2 / 7.0)
        else:            # clip(Poisson(5), 1, max)
            exp = stats.poisson.pmf(ks, 5.0)
            exp[0] += stats.poisson.pmf(0, 5.0)
        assert _chi2_pvalue(np.bincount(sizes, minlength=60)[1:60], exp * len(sizes)) > ALPHA, model
        grp = (keys >> np.uint64(32)).astype(np.int64)
        umi = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
        assert grp.min() == 0 and grp.max() == c - 1 and umi.max() < (1 << 24)
        assert stats.kstest(umi[starts] / float(1 << 24), "uniform").pvalue > ALPHA          # io.py:60-63
        q = ((pl >> 2) & 63).astype(int)
        # io.py:65-69  max(10, min(40, int(N(30,5))))
        edges = np.arange(10, 42)
        cdf = stats.norm.cdf(edges, 30, 5)
        exp_q = np.diff(cdf)
        exp_q[0] += cdf[0]
        exp_q[-1] += 1 - cdf[-1]                     # q == 40 collects [40, inf)
        # int() truncation: for x in [k, k+1) -> k (x > 0 here)
        assert _chi2_pvalue(np.bincount(q, minlength=41)[10:41], exp_q[:31] * len(q)) > ALPHA
        assert abs(((pl >> 8) & 1).mean() - 0.5) < 4 * 0.5 / np.sqrt(len(pl))
        rp = ((pl >> 9) & 255).astype(int)
        assert rp.min() >= 10 and rp.max() <= 149 and _chi2_pvalue(np.bincount(rp, minlength=150)[10:150], np.full(140, len(rp) / 140)) > ALPHA
        # alleles: io.py:97,105-108
        is_alt = (pl >> 31).astype(int)
        fam_of_read = np.repeat(np.arange(len(sizes)), sizes)
        for a in (0.0, 0.05, 0.3, 0.8):
            m = af[grp] == a
            expect = a * 0.9 + (1 - a) * 0.05
            # reads within a family are correlated (shared has_variant): compare family-level
            # alt fractions with the two-component mixture instead of a per-read binomial
            fam_alt = np.bincount(fam_of_read[m], weights=is_alt[m], minlength=len(sizes))
            fam_n = np.bincount(fam_of_read[m], minlength=len(sizes))
            big = fam_n >= 8                       # large families separate the two components cleanly
            if big.sum() > 200:
                assert abs(((fam_alt[big] / fam_n[big]) > 0.5).mean() - a) < 4 * np.sqrt(max(a * (1 - a), 1e-3) / big.sum()) + 2e-3, a
            assert abs(is_alt[m].mean() - expect) < 0.01, a
        allele = (pl & 3).astype(int)
        ref = (cid & 3)[grp]
        alt = ((ref + 1 + ((cid >> 2) % 3)[grp]) & 3)
        assert ((allele == ref) | (allele == alt)).all() and ((allele == alt) == (is_alt == 1)).all()


def test_read_generator_shuffle_is_a_permutation_and_contamination(ctx, native):
    cid = np.arange(40) + 3
    af = np.full(40, 0.1)
    nf = np.full(40, 3000)
    k0, p0 = ctx.simulate_reads(cid, af, nf, _sp(native, shuffle=0))
    k1, p1 = ctx.simulate_reads(cid, af, nf, _sp(native, shuffle=1))
    assert len(k0) == len(k1) and (k0 != k1).mean() > 0.9
    o0, o1 = np.lexsort((p0, k0)), np.lexsort((p1, k1))
    assert (k0[o0] == k1[o1]).all() and (p0[o0] == p1[o1]).all()
    # index hopping: ~hop_rate of the reads change group; UMI and payload allele survive
    kh, ph = ctx.simulate_reads(cid, af, nf, _sp(native, shuffle=0, hop_rate=0.02))
    moved = (kh >> np.uint64(32)) != (k0 >> np.uint64(32))
    assert abs(moved.mean() - 0.02) < 4 * np.sqrt(0.02 * 0.98 / len(k0))
    assert ((kh & np.uint64(0xFFFFFFFF)) == (k0 & np.uint64(0xFFFFFFFF))).all() and ((ph & 0xFF) == (p0 & 0xFF)).all()
    # barcode collisions: ~rate of the families share the previous family's UMI -> fewer distinct keys
    kc, _ = ctx.simulate_reads(cid, af, nf, _sp(native, shuffle=0, collision_rate=0.05))
    lost = len(np.unique(k0)) - len(np.unique(kc))
    assert abs(lost / (40 * 3000) - 0.05) < 0.01
    # contamination flip (io.py:100-101): at af = 0 a fraction `rate` of the families turns variant
    kz, pz = ctx.simulate_reads(cid, np.zeros(40), nf, _sp(native, shuffle=0, contamination_rate=0.2))
    assert abs((pz >> 31).mean() - (0.2 * 0.9 + 0.8 * 0.05)) < 0.01


# ------------------------------------------------------------------------------ fused pipeline
@pytest.mark.parametrize("test_type", ["poisson", "binomial"])
def test_pipeline_equals_oracle_on_the_same_reads(ctx, native, test_type):
    rng = np.random.default_rng(3)
    c = 96
    cid = np.arange(c) * 7 + 11
    af = np.tile([0.0, 0.0, 1e-4, 0.01, 0.05, 0.2], c // 6)
    nf = rng.integers(50, 3000, c)
    nf[5] = 0                                            # a cell with no families at all
    sp, up = _sp(native), _umi(native)
    tt = native.TEST_TYPES[test_type]
    res = ctx.pipeline_cells(cid, af, nf, sp, up, neg_af_max=1e-4, test_type=tt, alpha=0.05)
    keys, pl = ctx.simulate_reads(cid, af, nf, sp)          # the very reads the pipeline generated
    assert res["n_reads"].sum() == len(keys) == res["totals"][0]
    fam, grp = port.collapse_exact(keys, pl, 0, 3, 1000, 20, 0.6)
    alt = ((cid & 3) + 1 + ((cid >> 2) % 3)) & 3
    n_total, n_var, n_fam = np.zeros(c, np.int64), np.zeros(c, np.int64), np.zeros(c, np.int64)
    n_total[grp["group"]] = grp["consensus_count"].sum(1)
    n_var[grp["group"]] = grp["consensus_count"][np.arange(len(grp["group"])), alt[grp["group"]]]
    np.add.at(n_fam, (fam["key"] >> np.uint64(32)).astype(np.int64), 1)
    assert (res["n_total"] == n_total).all() and (res["n_variants"] == n_var).all() and (res["n_families"] == n_fam).all()
    assert (res["n_reads"] == np.bincount((keys >> np.uint64(32)).astype(np.int64), minlength=c)).all()
    # error rate: mean over 64 contexts of (n_var/n_neg) * U(0.5, 2)   (error_model.py:153-156, call.py:178)
    neg = af <= 1e-4
    vr = n_var[neg].sum() / n_total[neg].sum()
    assert 0.5 * vr <= res["mean_error_rate"] <= 2.0 * vr
    rate = res["mean_error_rate"]
    if test_type == "poisson":
        want_p = np.array([port.poisson_test(int(k), int(n) * rate) for k, n in zip(n_var, n_total)])
    else:
        want_p = np.array([port.binomial_test(int(k), int(n), rate) for k, n in zip(n_var, n_total)])
    ok = want_p >= 1e-50
    assert np.max(np.abs(res["p_value"][ok] - want_p[ok]) / want_p[ok]) < 1e-12
    assert (res["p_value"][~ok] < 1e-49).all()
    rej, adj = port.benjamini_hochberg(res["p_value"], 0.05)
    assert (res["significant"] == rej).all() and (res["p_adjusted"] == adj).all()
    assert res["significant"][(af >= 0.05) & (nf >= 1000)].all()      # (two-sided: blanks may be flagged LOW, App. A #13)
    # deterministic and independent of batch composition: rerun, and a sub-batch containing the negatives
    res2 = ctx.pipeline_cells(cid, af, nf, sp, up, neg_af_max=1e-4, test_type=tt, alpha=0.05)
    for k_ in ("n_total", "n_variants", "p_value", "p_adjusted"):
        assert (res[k_] == res2[k_]).all()
    sub = np.arange(0, c, 2)
    res3 = ctx.pipeline_cells(cid[sub], af[sub], nf[sub], sp, up, neg_af_max=1e-4, test_type=tt, alpha=0.05)
    assert (res3["n_total"] == res["n_total"][sub]).all() and (res3["n_variants"] == res["n_variants"][sub]).all()


# ------------------------------------------------------------------------------ fused engine (no reads written)
def _counts_from_oracle(keys, pl, cid, min_fs, qthr, cthr):
    c = len(cid)
    fam, grp = port.collapse_exact(keys, pl, 0, min_fs, 1000, qthr, cthr)
    alt = ((cid & 3) + 1 + ((cid >> 2) % 3)) & 3
    n_total, n_var, n_fam = np.zeros(c, np.int64), np.zeros(c, np.int64), np.zeros(c, np.int64)
    n_total[grp["group"]] = grp["consensus_count"].sum(1)
    n_var[grp["group"]] = grp["consensus_count"][np.arange(len(grp["group"])), alt[grp["group"]]]
    np.add.at(n_fam, (fam["key"] >> np.uint64(32)).astype(np.int64), 1)
    return n_fam, n_total, n_var


@pytest.mark.parametrize("cthr,qthr,coll,model", [(0.6, 20, 0.0, 1), (0.4, 20, 0.0, 1), (0.6, 30, 0.05, 1), (0.5, 20, 0.01, 0)])
def test_fused_engine_equals_oracle_on_family_major_reads(ctx, native, cthr, qthr, coll, model):
    """PMRD_ENGINE_FUSED (reads voted where they are drawn, never written; duplicate UMIs found by two hash bitmaps
    and replayed from the Philox counters): per-cell counts bit-equal to the oracle's collapse of the reads the
    materialising generator writes in family-major order (shuffle = 0).  Cells of 0 / 1 / a few / thousands / 70 000
    molecules: the 128-, 256- and 1024-thread blocks, natural 24-bit UMI coincidences (n^2 / 2^25 pairs) and forced
    ones (collision_rate)."""
    rng = np.random.default_rng(int(cthr * 100) + qthr)
    big = np.array([0, 1, 2, 33, 700, 5000, 70000, 31, 4096, 4097, 20000, 512])
    for nf in (big, rng.integers(1, 400, 150), rng.integers(500, 4000, 24)):
        c = len(nf)
        cid = np.arange(c, dtype=np.int64) * 11 + 5 + (np.int64(9) << 33)
        af = np.tile([0.0, 0.3, 0.5, 0.05, 0.01, 1e-4], c // 6 + 1)[:c]
        kw = dict(mean_family_size=4.0 if model else 3.0, family_model=model, collision_rate=coll)
        up = _umi(native, consensus_threshold=cthr, quality_threshold=qthr, min_family_size=2)
        res = ctx.pipeline_cells(cid, af, nf, _sp(native, engine=native.ENGINE_FUSED, **kw), up)
        keys, pl = ctx.simulate_reads(cid, af, nf, _sp(native, shuffle=0, **kw))
        assert res["n_reads"].sum() == len(keys) == res["totals"][0]
        assert (res["n_reads"] == np.bincount((keys >> np.uint64(32)).astype(np.int64), minlength=c)).all()
        n_fam, n_total, n_var = _counts_from_oracle(keys, pl, cid, 2, qthr, cthr)
        assert (res["n_families"] == n_fam).all(), np.flatnonzero(res["n_families"] != n_fam)
        assert (res["n_total"] == n_total).all(), np.flatnonzero(res["n_total"] != n_total)
        assert (res["n_variants"] == n_var).all(), np.flatnonzero(res["n_variants"] != n_var)
        if nf.max() >= 20000 or coll > 0:
            assert n_fam.sum() < nf.sum()                              # molecules did share UMIs: the replay ran
        # the call stage is the materialised plan's: same counts -> same table
        rej, adj = port.benjamini_hochberg(res["p_value"], 0.05)
        assert (res["significant"] == rej).all() and (res["p_adjusted"] == adj).all()
        # deterministic, and a cell's row does not depend on the batch
        again = ctx.pipeline_cells(cid[::-1].copy(), af[::-1].copy(), nf[::-1].copy(), _sp(native, engine=native.ENGINE_FUSED, **kw), up)
        for k_ in ("n_reads", "n_families", "n_total", "n_variants"):
            assert (again[k_][::-1] == res[k_]).all(), k_


def test_fused_engine_agrees_with_the_materialised_plan_and_refuses_hopping(ctx, native):
    """Same molecules, same reads: the fused engine and the sort-based plan see each family's reads in different
    orders (family-major vs window-shuffled), which can only matter through fp64 rounding at the consensus threshold or
    a first-seen tie (threshold <= 0.5).  At the default thresholds the two tables are equal."""
    rng = np.random.default_rng(12)
    c = 64
    cid = np.arange(c, dtype=np.int64) + 300
    af = np.tile([0.0, 1e-4, 1e-3, 1e-2, 0.05, 0.2, 0.0, 0.0], c // 8)
    nf = rng.integers(100, 12000, c)
    up = _umi(native)
    a = ctx.pipeline_cells(cid, af, nf, _sp(native, family_model=1, engine=native.ENGINE_FUSED), up)
    b = ctx.pipeline_cells(cid, af, nf, _sp(native, family_model=1), up)
    for k_ in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
        assert (a[k_] == b[k_]).all(), k_
    assert a["mean_error_rate"] == b["mean_error_rate"]
    with pytest.raises(native.NativeError):
        ctx.pipeline_cells(cid, af, nf, _sp(native, engine=native.ENGINE_FUSED, hop_rate=0.01), up)


# ------------------------------------------------------------------------------ per-cell knobs: one plan for the C4 sweep
def test_one_plan_with_per_cell_knobs_equals_one_plan_per_condition(ctx, native):
    """pmrd_plan_create_ex: a contamination sweep (hop rates x collision rates, each condition = test cells + blanks) as
    ONE plan with per-cell knobs, hop pools and call groups gives, condition by condition, exactly the table of a
    separate plan with the plan-wide knob: counts, error rate, p-values, and BH inside the group (bit-equal q)."""
    rng = np.random.default_rng(21)
    conds = [("hop", 0.0), ("hop", 0.01), ("hop", 0.05), ("coll", 0.0), ("coll", 0.02), ("coll", 0.1)]
    m = 14
    up = _umi(native)
    cid, af, nf, hop, p0, pn, coll, gs, singles = [], [], [], [], [], [], [], [0], []
    for gi, (kind, rate) in enumerate(conds):
        c = (np.int64(7) << 40) + gi * 1000 + np.arange(m, dtype=np.int64)
        a = np.concatenate([np.full(8, 0.05), np.zeros(6)])
        f = rng.integers(300, 2500, m)
        start = gs[-1]
        cid.append(c); af.append(a); nf.append(f)
        hop.append(np.full(m, rate if kind == "hop" else 0.0)); p0.append(np.full(m, start)); pn.append(np.full(m, m))
        coll.append(np.full(m, rate if kind == "coll" else 0.0))
        gs.append(start + m)
        sp = _sp(native, family_model=1, hop_rate=rate if kind == "hop" else 0.0, collision_rate=rate if kind == "coll" else 0.0)
        singles.append(ctx.pipeline_cells(c, a, f, sp, up, neg_af_max=0.0, alternative=native.ALT_GREATER))
    knobs = {"hop_rate": np.concatenate(hop), "hop_pool_start": np.concatenate(p0), "hop_pool_size": np.concatenate(pn),
             "collision_rate": np.concatenate(coll), "call_group_start": np.array(gs)}
    res = ctx.pipeline_cells(np.concatenate(cid), np.concatenate(af), np.concatenate(nf), _sp(native, family_model=1), up,
                             neg_af_max=0.0, alternative=native.ALT_GREATER, knobs=knobs)
    assert len(res["group_error_rate"]) == len(conds)
    for gi, one in enumerate(singles):
        sl = slice(gs[gi], gs[gi + 1])
        for k in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
            assert (res[k][sl] == one[k]).all(), (conds[gi], k)
        assert res["group_error_rate"][gi] == one["mean_error_rate"]
        rej, adj = port.benjamini_hochberg(res["p_value"][sl], 0.05)
        assert (res["significant"][sl] == rej).all() and (res["p_adjusted"][sl] == adj).all()
    # hopping did move reads between the cells of a pool (more, smaller families), never across pools
    assert res["n_families"][gs[2]:gs[3]].sum() > np.concatenate(nf)[gs[2]:gs[3]].sum()
    assert res["n_reads"].sum() == sum(o["n_reads"].sum() for o in singles)
    # a hop pool that a chunk boundary would split, or that does not contain its cell, is refused
    bad = dict(knobs)
    bad["hop_pool_start"] = np.concatenate(p0) + 1
    with pytest.raises(native.NativeError):
        ctx.pipeline_cells(np.concatenate(cid), np.concatenate(af), np.concatenate(nf), _sp(native, family_model=1), up, knobs=bad)


@pytest.mark.parametrize("fused", [False, True])
def test_cross_sample_contamination_mixes_molecules_of_a_second_sample(ctx, native, fused):
    """Read-level cross-sample contamination (contamination.py:280-311): int(n * proportion) molecules of a sample with
    its own allele fraction are added to the cell -- own UMIs, own reads, collapsed together with the cell's.  The mixed
    cell must look like the union of a clean cell and a separate cell of the contaminating sample: read and family
    counts add up, and the variant consensus count matches the sum within binomial noise (a chi-square on the 2 x 2
    table variant / non-variant x mixed / union at alpha = 0.001)."""
    n, prop, af2 = 20000, 0.25, 0.5
    c = 40
    cid = (np.int64(11) << 40) + np.arange(c, dtype=np.int64)
    eng = dict(engine=native.ENGINE_FUSED) if fused else {}
    sp = _sp(native, family_model=1, **eng)
    up = _umi(native)
    z = np.zeros(c)
    mixed = ctx.pipeline_cells(cid, z, np.full(c, n), sp, up, knobs={"cross_proportion": np.full(c, prop), "cross_af": np.full(c, af2)})
    clean = ctx.pipeline_cells(cid, z, np.full(c, n), sp, up)
    same = ctx.pipeline_cells(cid, z, np.full(c, n), sp, up, knobs={"cross_proportion": z, "cross_af": np.full(c, af2)})
    for k in ("n_reads", "n_families", "n_total", "n_variants"):
        assert (same[k] == clean[k]).all(), k                                        # proportion 0: nothing is mixed in
    donor = ctx.pipeline_cells(cid + 5000, np.full(c, af2), np.full(c, int(n * prop)), sp, up)
    n_add = int(n * prop)
    # molecules: exactly n + int(n * prop) drawn; reads and families follow (minus the few UMI coincidences)
    assert (mixed["n_reads"] > clean["n_reads"]).all()
    assert abs(mixed["n_reads"].sum() / clean["n_reads"].sum() - (1 + prop)) < 0.01
    assert abs(mixed["n_families"].sum() - (clean["n_families"].sum() + donor["n_families"].sum())) < 0.002 * c * n
    # the cell's own molecules are untouched: the clean cell's variant families are still there, the rest is the donor's share
    a = np.array([[mixed["n_variants"].sum(), mixed["n_total"].sum() - mixed["n_variants"].sum()],
                  [clean["n_variants"].sum() + donor["n_variants"].sum(),
                   clean["n_total"].sum() + donor["n_total"].sum() - clean["n_variants"].sum() - donor["n_variants"].sum()]], float)
    exp = a.sum(1, keepdims=True) * a.sum(0, keepdims=True) / a.sum()
    chi2 = ((a - exp) ** 2 / exp).sum()
    assert chi2 < 10.83, (chi2, a)                                                   # 1 dof, alpha = 0.001
    frac = (mixed["n_variants"].sum() - clean["n_variants"].sum()) / max(1, donor["n_total"].sum())
    assert 0.35 < frac < 0.55                                                        # ~ af2 x P(alt consensus | variant molecule)
    assert n_add == 5000


@pytest.mark.parametrize("hop,coll", [(0.01, 0.0), (0.05, 0.02)])
def test_pipeline_with_index_hopping_equals_oracle(ctx, native, hop, coll):
    """C4 (eval-contamination) read-level knobs through the plan: with hop_rate > 0 reads leave their cell, so the
    plan takes the general path (12-byte records, full-key sort, family + group tables).  Same gate as the
    cell-major path: the plan's per-cell counts equal the oracle's collapse of the very reads it generated."""
    rng = np.random.default_rng(8)
    c = 60
    cid = np.arange(c) * 3 + 1000
    af = np.tile([0.0, 0.0, 0.01, 0.05, 0.2, 0.001], c // 6)
    nf = rng.integers(100, 4000, c)
    sp, up = _sp(native, hop_rate=hop, collision_rate=coll), _umi(native)
    res = ctx.pipeline_cells(cid, af, nf, sp, up, neg_af_max=1e-4, test_type=native.TEST_POISSON, alpha=0.05)
    keys, pl = ctx.simulate_reads(cid, af, nf, sp)
    assert res["n_reads"].sum() == len(keys)
    k0, _ = ctx.simulate_reads(cid, af, nf, _sp(native, collision_rate=coll))
    moved = ((keys >> np.uint64(32)) != (k0 >> np.uint64(32))).mean()
    assert abs(moved - hop) < 5 * np.sqrt(hop / len(keys)) + 1e-4              # reads did hop, at the requested rate
    fam, grp = port.collapse_exact(keys, pl, 0, 3, 1000, 20, 0.6)
    alt = ((cid & 3) + 1 + ((cid >> 2) % 3)) & 3
    n_total, n_var, n_fam = np.zeros(c, np.int64), np.zeros(c, np.int64), np.zeros(c, np.int64)
    n_total[grp["group"]] = grp["consensus_count"].sum(1)
    n_var[grp["group"]] = grp["consensus_count"][np.arange(len(grp["group"])), alt[grp["group"]]]
    np.add.at(n_fam, (fam["key"] >> np.uint64(32)).astype(np.int64), 1)
    assert (res["n_total"] == n_total).all() and (res["n_variants"] == n_var).all() and (res["n_families"] == n_fam).all()
    # hopped reads open small foreign families in the receiving cell: more families, (almost) none of them kept
    base = ctx.pipeline_cells(cid, af, nf, _sp(native, collision_rate=coll), up)
    assert res["n_families"].sum() > base["n_families"].sum() and res["n_total"].sum() <= base["n_total"].sum()
    want_p = np.array([port.poisson_test(int(k), int(n) * res["mean_error_rate"]) for k, n in zip(n_var, n_total)])
    ok = want_p >= 1e-50
    assert np.max(np.abs(res["p_value"][ok] - want_p[ok]) / want_p[ok]) < 1e-12


@pytest.mark.parametrize("cthr,qthr", [(0.4, 20), (0.5, 30), (0.51, 20), (0.9, 0)])
def test_plan_consensus_thresholds_either_side_of_one_half(ctx, native, cthr, qthr):
    """The fused consensus kernel walks with a lighter vote when consensus_threshold > 0.5 (a tie for the best weight
    can never pass, so the first-seen tie-break is moot) and with the full one otherwise: both against the oracle."""
    rng = np.random.default_rng(17)
    c = 40
    cid = np.arange(c) * 5 + 2
    af = np.tile([0.0, 0.3, 0.5, 0.05], c // 4)          # high AF: many mixed families, ties and near-threshold votes
    nf = rng.integers(200, 2500, c)
    sp, up = _sp(native, mean_family_size=2.5, family_model=1), _umi(native, consensus_threshold=cthr, quality_threshold=qthr, min_family_size=2)
    res = ctx.pipeline_cells(cid, af, nf, sp, up)
    # shuffle=2: the reads in the plan's own (cell-major) order -- the first-seen tie-break depends on it
    keys, pl = ctx.simulate_reads(cid, af, nf, _sp(native, mean_family_size=2.5, family_model=1, shuffle=2))
    k1, p1 = ctx.simulate_reads(cid, af, nf, sp)
    assert (np.sort(keys) == np.sort(k1)).all() and np.sort(pl).tobytes() == np.sort(p1).tobytes()   # the same reads, reordered
    fam, grp = port.collapse_exact(keys, pl, 0, 2, 1000, qthr, cthr)
    alt = ((cid & 3) + 1 + ((cid >> 2) % 3)) & 3
    n_total, n_var = np.zeros(c, np.int64), np.zeros(c, np.int64)
    n_total[grp["group"]] = grp["consensus_count"].sum(1)
    n_var[grp["group"]] = grp["consensus_count"][np.arange(len(grp["group"])), alt[grp["group"]]]
    assert (res["n_total"] == n_total).all() and (res["n_variants"] == n_var).all()
    assert n_total.sum() > 0 and (n_total < res["n_families"]).any()          # some families did fail the vote


def test_families_longer_than_a_shuffle_window(ctx, native, monkeypatch):
    """Families of thousands of reads span three and more 4096-read windows: the generator's per-window destination
    tiles (two precomputed, the rest worked out per read) and its by-product histogram must still describe the very
    layout the sort and the consensus walk."""
    rng = np.random.default_rng(4)
    c = 24
    cid = np.arange(c) + 90
    af = np.tile([0.0, 0.2, 0.01], c // 3)
    nf = rng.integers(3, 40, c)
    sp = _sp(native, family_model=0, mean_family_size=3000.0, max_family_size=30000)
    up = _umi(native, max_family_size=100000)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("PMRD_GEN_HIST", flag)
        out[flag] = ctx.pipeline_cells(cid, af, nf, sp, up)
    for k_ in ("n_reads", "n_families", "n_total", "n_variants", "p_value"):
        assert (out["0"][k_] == out["1"][k_]).all(), k_
    res = out["1"]
    keys, pl = ctx.simulate_reads(cid, af, nf, _sp(native, family_model=0, mean_family_size=3000.0, max_family_size=30000, shuffle=2))
    assert len(keys) == res["n_reads"].sum()
    sizes = np.unique(keys, return_counts=True)[1]
    assert sizes.max() > 3 * 4096                                  # the case is really exercised
    fam, grp = port.collapse_exact(keys, pl, 0, 3, 100000, 20, 0.6)
    alt = ((cid & 3) + 1 + ((cid >> 2) % 3)) & 3
    n_total, n_var = np.zeros(c, np.int64), np.zeros(c, np.int64)
    n_total[grp["group"]] = grp["consensus_count"].sum(1)
    n_var[grp["group"]] = grp["consensus_count"][np.arange(len(grp["group"])), alt[grp["group"]]]
    assert (res["n_total"] == n_total).all() and (res["n_variants"] == n_var).all()


@pytest.mark.parametrize("cthr", [0.6, 0.4])
def test_small_cells_are_sorted_and_voted_inside_one_block(ctx, native, monkeypatch, cthr):
    """Grids whose cells all hold <= 4096 reads (the C5 panel: ~1250 reads per locus) take the compact layout: no tile
    padding, one block sorts and votes a cell in shared memory.  Same tables as the tile-padded path, and equal to the
    oracle on the plan's own read order (first-seen tie-breaks included)."""
    rng = np.random.default_rng(31)
    c = 700
    cid = (np.int64(3) << 20) + np.arange(c)
    af = np.array([0.0, 1e-4, 0.3, 0.01])[np.arange(c) % 4]
    nf = rng.integers(0, 700, c)                     # 0 .. ~3500 reads: 2, 4, 8 and 16 records per thread
    nf[:5] = [0, 1, 2, 250, 699]
    sp = _sp(native, family_model=1, mean_family_size=5.0, collision_rate=2e-3)
    up = _umi(native, consensus_threshold=cthr, min_family_size=2)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("PMRD_SMALL_CELLS", flag)
        plan = native.Plan(ctx, cid, af, nf, sp, up)
        ctx.profile_enable(True)
        plan.run()
        prof = ctx.profile_report()
        ctx.profile_enable(False)
        out[flag] = (plan.results(), prof)
        plan.close()
    assert "small_cell_kernel" in out["1"][1] and "small_cell_kernel" not in out["0"][1]
    assert "rs_downsweep_direct_kernel" not in out["1"][1]
    a, b = out["1"][0], out["0"][0]
    for k_ in ("n_reads", "n_families"):
        assert (a[k_] == b[k_]).all(), k_
    # (the two layouts order a cell's reads differently, so only order-free columns must agree between them ...)
    if cthr > 0.5:
        for k_ in ("n_total", "n_variants", "p_value", "p_adjusted", "significant"):
            assert (a[k_] == b[k_]).all(), k_
    # ... and each must equal the oracle run on ITS OWN read order
    for flag in ("1", "0"):
        monkeypatch.setenv("PMRD_SMALL_CELLS", flag)
        keys, pl = ctx.simulate_reads(cid, af, nf, _sp(native, family_model=1, mean_family_size=5.0, collision_rate=2e-3, shuffle=2))
        fam, grp = port.collapse_exact(keys, pl, 0, 2, 1000, 20, cthr)
        alt = ((cid & 3) + 1 + ((cid >> 2) % 3)) & 3
        n_total, n_var, n_fam = np.zeros(c, np.int64), np.zeros(c, np.int64), np.zeros(c, np.int64)
        n_total[grp["group"]] = grp["consensus_count"].sum(1)
        n_var[grp["group"]] = grp["consensus_count"][np.arange(len(grp["group"])), alt[grp["group"]]]
        np.add.at(n_fam, (fam["key"] >> np.uint64(32)).astype(np.int64), 1)
        r = out[flag][0]
        assert (r["n_total"] == n_total).all() and (r["n_variants"] == n_var).all() and (r["n_families"] == n_fam).all(), flag
    # the block sorts the low 16 UMI bits (two ranked passes) and separates the ~1 % of runs that hold several families in
    # the vote; PMRD_SMALL_SORT_BITS=24 makes the third pass instead: the same table
    monkeypatch.setenv("PMRD_SMALL_CELLS", "1")
    monkeypatch.setenv("PMRD_SMALL_SORT_BITS", "24")
    plan = native.Plan(ctx, cid, af, nf, sp, up)
    plan.run()
    full_sort = plan.results()
    plan.close()
    monkeypatch.delenv("PMRD_SMALL_SORT_BITS")
    for k_ in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
        assert (full_sort[k_] == a[k_]).all(), k_
    # streamed through many compact chunks: nothing changes
    monkeypatch.setenv("PMRD_SMALL_CELLS", "1")
    plan = native.Plan(ctx, cid, af, nf, _sp(native, family_model=1, mean_family_size=5.0, collision_rate=2e-3, chunk_log2=16), up)
    assert plan.n_chunks > 5
    plan.run()
    many = plan.results()
    plan.close()
    for k_ in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
        assert (many[k_] == a[k_]).all(), k_
    # the compact layout does scramble a cell (inside its power-of-two prefix)
    monkeypatch.setenv("PMRD_SMALL_CELLS", "1")
    keys, _ = ctx.simulate_reads(cid[3:4], af[3:4], nf[3:4], _sp(native, family_model=1, shuffle=2))
    runs = 1 + int((np.diff(keys.astype(np.int64)) != 0).sum())
    assert runs > 3 * 250                                   # 250 families would form 250 runs if left family by family


@pytest.mark.parametrize("nf", [[4096, 4095, 1, 2048, 4096, 0, 3], [4097, 8191, 8192, 8193, 12288, 16384, 4096, 1, 0, 20480]])
def test_cells_exactly_at_tile_boundaries(ctx, native, monkeypatch, nf):
    """max_family_size = 1 makes every family one read, so a cell of n families holds exactly n reads: cells that fill
    whole 4096-record tiles (no padding, empty tail window), one read more, one less; the first list takes the compact
    layout (every cell <= 4096 reads), the second the tile-padded one."""
    nf = np.array(nf, dtype=np.int64)
    c = len(nf)
    cid = np.arange(c) + 500
    af = np.tile([0.0, 0.5], c)[:c]
    sp = _sp(native, family_model=1, max_family_size=1)
    up = _umi(native, min_family_size=1)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("PMRD_GEN_HIST", flag)
        out[flag] = ctx.pipeline_cells(cid, af, nf, sp, up)
    monkeypatch.setenv("PMRD_GEN_HIST", "1")
    res = out["1"]
    for k_ in ("n_reads", "n_families", "n_total", "n_variants"):
        assert (out["0"][k_] == res[k_]).all(), k_
    assert (res["n_reads"] == nf).all()
    keys, pl = ctx.simulate_reads(cid, af, nf, _sp(native, family_model=1, max_family_size=1, shuffle=2))
    assert len(keys) == nf.sum()
    fam, grp = port.collapse_exact(keys, pl, 0, 1, 1000, 20, 0.6)
    alt = ((cid & 3) + 1 + ((cid >> 2) % 3)) & 3
    n_total, n_var, n_fam = np.zeros(c, np.int64), np.zeros(c, np.int64), np.zeros(c, np.int64)
    n_total[grp["group"]] = grp["consensus_count"].sum(1)
    n_var[grp["group"]] = grp["consensus_count"][np.arange(len(grp["group"])), alt[grp["group"]]]
    np.add.at(n_fam, (fam["key"] >> np.uint64(32)).astype(np.int64), 1)
    assert (res["n_total"] == n_total).all() and (res["n_variants"] == n_var).all() and (res["n_families"] == n_fam).all()
    assert (res["n_families"] <= nf).all() and res["n_families"].sum() > 0.99 * nf.sum()      # (UMI collisions merge a few)


def test_plan_stage_entry_points_compose(ctx, native):
    cid = np.arange(32)
    af = np.tile([0.0, 0.02], 16)
    nf = np.full(32, 2000)
    plan = native.Plan(ctx, cid, af, nf, _sp(native), _umi(native))
    assert plan.n_chunks == 1 and plan.chunk_reads(0) == plan.n_reads
    plan.run_simulate(0); plan.run_collapse(0); plan.run_call()
    a = plan.results()
    plan.run()
    b = plan.results()
    plan.close()
    for k_ in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
        assert (a[k_] == b[k_]).all(), k_
    assert a["totals"][0] == plan.n_reads and a["totals"][2] == 32
    # a plan is sized for the seed it was built with (read counts depend on it): running it under another seed is refused
    plan = native.Plan(ctx, cid, af, nf, _sp(native), _umi(native))
    seed = ctx.seed
    ctx.set_seed(seed + 1)
    with pytest.raises(native.NativeError, match="another seed"):
        plan.run()
    ctx.set_seed(seed)
    plan.run()
    assert (plan.results()["n_total"] == b["n_total"]).all()
    plan.close()


def test_chunked_plan_equals_single_chunk(ctx, native):
    """Streaming the grid through small chunks must not change a single count or p-value."""
    rng = np.random.default_rng(8)
    c = 300
    cid = np.arange(c) + 5
    af = np.tile([0.0, 1e-4, 0.01, 0.05, 0.0005], c // 5)
    nf = rng.integers(100, 2500, c)
    one = ctx.pipeline_cells(cid, af, nf, _sp(native), _umi(native))
    plan = native.Plan(ctx, cid, af, nf, _sp(native, chunk_log2=16), _umi(native))
    assert plan.n_chunks > 5 and sum(plan.chunk_reads(i) for i in range(plan.n_chunks)) == plan.n_reads
    plan.run()
    many = plan.results()
    ctx.profile_enable(True)
    plan.run()
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    n_chunks = plan.n_chunks
    plan.close()
    for k_ in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
        assert (one[k_] == many[k_]).all(), k_
    assert one["mean_error_rate"] == many["mean_error_rate"] and (one["totals"] == many["totals"]).all()
    assert prof["reads_gen_kernel"][0] == n_chunks and prof["reads_gen_kernel"][1] > 0
    assert prof["bh_group_kernel"][0] == 1 and prof["pvalues_kernel"][0] == 1       # one call stage over all chunks (<= 4096 cells: BH in one block)


def test_generator_histogram_replaces_the_first_counting_pass(ctx, native, monkeypatch):
    """The generator counts the first radix digit per output tile as a by-product (one reduction per
    family and window); the sort must then give the very same tables with one upsweep fewer per chunk."""
    rng = np.random.default_rng(21)
    c = 120
    cid = np.arange(c) * 3 + 1
    af = np.tile([0.0, 1e-4, 0.01, 0.05], c // 4)
    nf = rng.integers(1, 12000, c)           # 1 .. 15 tiles per cell: tail-only cells, permuted windows, padding
    nf[7] = 0
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("PMRD_GEN_HIST", flag)
        plan = native.Plan(ctx, cid, af, nf, _sp(native, family_model=1, collision_rate=1e-3), _umi(native))
        plan.run()
        ctx.profile_enable(True)
        plan.run()
        prof = ctx.profile_report()
        ctx.profile_enable(False)
        out[flag] = (plan.results(), prof["rs_upsweep_kernel"][0], plan.n_chunks)
        plan.close()
    a, b = out["0"][0], out["1"][0]
    for k_ in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
        assert (a[k_] == b[k_]).all(), k_
    assert out["0"][1] - out["1"][1] == out["1"][2] >= 1      # one counting pass per chunk gone (BH sorts too)


# ------------------------------------------------------------------------------ K11 LoD fit
def test_fit_lod_matches_reference_golden(ctx, golden, monkeypatch):
    g = golden("lod_fit.npz")
    af = g["af"]
    for name in ("mid", "steep", "shallow"):
        lod, lo, hi = ctx.fit_lod(af, g[f"{name}_hit"], 0.95, 200, stream_id=1)
        assert abs(lod - float(g[f"{name}_lod"])) / float(g[f"{name}_lod"]) < 1e-6, name   # same LM optimum
        rlo, rhi = g[f"{name}_ci"]
        # bootstrap CI of 200 resamples of 15 points: within bootstrap noise of the reference's
        assert lo <= lod * 1.0001 and hi >= lod * 0.9999
        assert abs(np.log(lo / rlo)) < 0.35 and abs(np.log(hi / rhi)) < 0.6, (name, lo, hi, rlo, rhi)
    # the reference's own grids give all-zero hit rates -> curve_fit fails -> interp1d fallback -> NaN
    lod, lo, hi = ctx.fit_lod(af, g["zeros_hit"], 0.95, 200)
    assert np.isnan(lod) and np.isnan(lo) and np.isnan(hi)
    assert np.isnan(float(g["zeros_lod"]))
    # determinism
    assert ctx.fit_lod(af, g["mid_hit"], 0.95, 50, 9) == ctx.fit_lod(af, g["mid_hit"], 0.95, 50, 9)
    # one warp per fit (lane = curve point, 2 x 2 algebra in registers) takes the decisions of one thread per fit
    # (the general loop form): the same NaN pattern and the same optimum to within lmdif's own stopping tolerance
    # (ftol = xtol = 1.5e-8; measured: the same bits, but FMA contraction is the compiler's choice)
    rng = np.random.default_rng(5)
    curves = np.clip(np.sort(rng.random((12, len(af))), axis=1) ** rng.uniform(0.5, 6.0, (12, 1)), 0, 1)
    curves[3] = 0.0
    curves[4, :-2] = 0.0
    ids = list(range(40, 52))
    monkeypatch.setenv("PMRD_LOD_WARP", "0")
    ref = ctx.fit_lod_batch(af, curves, 0.95, 200, ids)
    monkeypatch.setenv("PMRD_LOD_WARP", "1")
    got = ctx.fit_lod_batch(af, curves, 0.95, 200, ids)
    for r, g_ in zip(ref, got):
        for a_, b_ in zip(r, g_):
            assert (np.isnan(a_) and np.isnan(b_)) or abs(a_ - b_) <= 1e-6 * abs(a_), (r, g_)
    # all curves of an analysis in one launch == one launch per curve
    names = [k[:-4] for k in g.files if k.endswith("_hit")]
    batch = ctx.fit_lod_batch(af, [g[f"{n}_hit"] for n in names], 0.95, 60, list(range(3, 3 + len(names))))
    for i, n in enumerate(names):
        assert str(batch[i]) == str(ctx.fit_lod(af, g[f"{n}_hit"], 0.95, 60, 3 + i)), n



def test_two_pass_sort_with_in_tile_peel_equals_the_full_sort(ctx, native, monkeypatch):
    """Default plan: two radix passes order the low 16 UMI bits and cell_consensus_peel_kernel separates the families
    that share a 16-bit run inside its tile.  PMRD_SORT_BITS=24 orders every bit (three passes) and votes whole runs;
    PMRD_SORT_BITS=8 leaves 16 bits to the general peel.  All three must give the same table -- for cells big enough that a
    third of the runs hold several families (50 000 families over 65 536 values), with duplicate UMIs (collision_rate),
    runs that straddle tiles, padding, and under both vote forms (consensus_threshold above and below one half)."""
    rng = np.random.default_rng(77)
    nf = np.concatenate([[50000, 41000, 1, 0, 3, 4097], rng.integers(1, 30000, 34)])
    c = len(nf)
    cid = np.arange(c) * 7 + 2
    af = np.tile([0.0, 0.02, 0.3, 0.001], c // 4)
    for cthr, minfs in ((0.6, 3), (0.5, 1), (0.3, 2)):
        out = {}
        for bits in ("24", "16", "8"):
            monkeypatch.setenv("PMRD_SORT_BITS", bits)
            plan = native.Plan(ctx, cid, af, nf, _sp(native, family_model=1, collision_rate=2e-3), _umi(native, consensus_threshold=cthr, min_family_size=minfs))
            plan.run()
            ctx.profile_enable(True)
            plan.run()
            prof = ctx.profile_report()
            ctx.profile_enable(False)
            out[bits] = (plan.results(), prof)
            plan.close()
        assert "cell_consensus_peel_kernel" in out["16"][1] and "cell_consensus_peel_kernel" not in out["24"][1]
        for bits in ("16", "8"):
            for k_ in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
                assert (out["24"][0][k_] == out[bits][0][k_]).all(), (cthr, bits, k_)
        assert out["24"][0]["n_variants"].sum() > 0
    monkeypatch.delenv("PMRD_SORT_BITS")
    plan = native.Plan(ctx, cid[:4], af[:4], nf[:4], _sp(native, family_model=1), _umi(native))
    ctx.profile_enable(True)
    plan.run()
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    plan.close()
    assert "cell_consensus_peel_kernel" in prof                     # the default
