"""Full-size (BASELINE.json configs[1] = configs/default.yaml: 20 000 cells, 4.25e8 families,
2.1e9 reads) checks of the hot path through size-independent properties -- the oracle cannot
run this size in seconds, so what is asserted is what must hold at ANY size: conservation of
reads and families, idempotence, independence of a cell's counts from chunking and from the
batch it is processed in, a sorted + stable radix sort at 5e7 pairs, and the statistical shape
of the call table."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(ctx, native):
    from pathlib import Path

    from precise_mrd import grid
    from precise_mrd.config import load_config

    root = Path(__file__).resolve().parent.parent / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs"
    cfg = load_config(root / "default.yaml")
    sim = cfg.simulation
    cell_id, af, depth, _ = grid.grid_cells(sim.allele_fractions, sim.umi_depths, sim.n_replicates)
    ctx.set_seed(int(cfg.seed))
    sp = grid.read_sim_params(cfg, mean_family_size=5.0, family_model=1)
    up = grid.umi_params(cfg)
    plan = native.Plan(ctx, cell_id, af, depth, sp, up, neg_af_max=1e-4, test_type=native.TEST_POISSON, alpha=float(cfg.stats.alpha))
    plan.run()
    res = plan.results()
    yield {"cfg": cfg, "cell_id": cell_id, "af": af, "depth": depth, "sp": sp, "up": up, "plan": plan, "res": res, "grid": grid}
    plan.close()
    ctx.set_seed(7)


def test_c2_conservation_and_shape(c2):
    res, depth, af, plan = c2["res"], c2["depth"], c2["af"], c2["plan"]
    assert len(depth) == 20000 and depth.sum() == 425_000_000
    assert plan.n_families == depth.sum() and plan.n_reads > 2.0e9
    assert sum(plan.chunk_reads(i) for i in range(plan.n_chunks)) == plan.n_reads and plan.n_chunks >= 2
    # every read is accounted to exactly one cell; families formed = distinct UMIs, at most the molecules drawn
    assert res["n_reads"].sum() == plan.n_reads == res["totals"][0]
    assert (res["n_families"] <= depth).all() and (res["n_families"] >= 0.99 * depth).all()
    assert res["totals"][1] == res["n_families"].sum() and res["totals"][2] == 20000
    # consensus families: a subset of the families formed (min family size 3 of clip(Poisson(5),1,..): ~87 % pass)
    assert (res["n_total"] <= res["n_families"]).all() and (res["n_variants"] <= res["n_total"]).all()
    frac = res["n_total"].sum() / res["n_families"].sum()
    assert 0.84 < frac < 0.90
    # mean reads per family = E[clip(Poisson(5), 1, ...)] = 5 + P(0) = 5.0067
    assert abs(res["n_reads"].sum() / depth.sum() - 5.0067) < 2e-3
    # the variant-family fraction tracks the allele fraction (alt consensus needs a variant molecule or an error run)
    vf = res["n_variants"] / np.maximum(res["n_total"], 1)
    means = [vf[af == a].mean() for a in np.sort(np.unique(af))]
    assert all(x < y for x, y in zip(means, means[1:]))                     # monotone in AF
    base = vf[af == af.min()].mean()                                        # background: error-only alt consensus
    assert 0.005 < base < 0.03
    for a in np.unique(af)[np.unique(af) >= 0.01]:
        assert 0.8 < (vf[af == a].mean() - base) / (a - af.min()) < 1.05, a   # ~one alt consensus per variant molecule
    # p-values are probabilities, BH never lowers them, significant <=> adjusted p <= alpha
    p, q = res["p_value"], res["p_adjusted"]
    assert ((p >= 0) & (p <= 1)).all() and (q >= p - 1e-18).all() and (q <= 1).all()
    assert (res["significant"].astype(bool) == (q <= 0.05)).all()
    assert res["significant"][af >= 0.05].all() and res["significant"][af >= 0.01].mean() > 0.95
    # (negative-control cells are also flagged at a sizeable rate: the reference's expected count carries a
    # U(0.5, 2) modifier and its test is two-sided -- SURVEY.md Appendix A #13 -- so only the ordering is asserted)
    assert res["significant"][af <= 1e-4].mean() < 0.6 < res["significant"][af >= 0.01].mean() and 0.0 < res["mean_error_rate"] < 0.1


def test_c2_strided_cells_equal_the_oracle(ctx, native, c2):
    """Oracle comparison AT C2 CELL SIZES: a strided sample of the real grid's cells (every depth 5k / 10k / 20k / 50k,
    every AF, 48 cells; the 50 000-family cells span 61 tiles, so their families straddle many tile boundaries and the
    four-range cell scan runs).  The plan's own reads of each cell (``simulate_reads(shuffle=2)`` = the cell-major
    order the sort sees, first-seen tie-breaks included) go through the C restatement of the reference
    (``oracle/pmrd_oracle.c``: umi.py:97-203), and the three per-cell counts of the full 20 000-cell run must be
    bit-equal to it."""
    from oracle import cport

    af, depth, cid, res, cfg = c2["af"], c2["depth"], c2["cell_id"], c2["res"], c2["cfg"]
    # cells are AF-major, then depth, then 1000 replicates: stride 419 (coprime to 1000) walks every (AF, depth) stratum
    sel = np.arange(7, len(depth), 419)[:48]
    strata = {(a, d) for a, d in zip(af[sel], depth[sel])}
    assert len(strata) >= 18 and (depth[sel] == 50000).sum() >= 8
    sp2 = c2["grid"].read_sim_params(cfg, mean_family_size=5.0, family_model=1)
    sp2.shuffle = 2
    u = cfg.umi
    checked = 0
    for lo in range(0, len(sel), 8):                       # 8 cells (~1e6 reads) per round trip
        s = sel[lo:lo + 8]
        keys, pl = ctx.simulate_reads(cid[s], af[s], depth[s], sp2)
        assert len(keys) == res["n_reads"][s].sum()
        fam = cport.collapse_weighted(keys, pl, int(u.min_family_size), int(u.max_family_size), int(u.quality_threshold),
                                      float(u.consensus_threshold))
        g = (fam["key"] >> np.uint64(32)).astype(np.int64)              # group = position of the cell in the sub-batch
        alt = ((cid[s] & 3) + 1 + ((cid[s] >> 2) % 3)) & 3
        ok = (fam["flags"] & 3) == 3                                    # has consensus and kept (>= min_family_size)
        n_fam = np.bincount(g, minlength=len(s))
        n_tot = np.bincount(g[ok], minlength=len(s))
        n_var = np.bincount(g[ok & (fam["consensus"] == alt[g])], minlength=len(s))
        assert (res["n_families"][s] == n_fam).all(), (lo, res["n_families"][s], n_fam)
        assert (res["n_total"][s] == n_tot).all(), (lo, res["n_total"][s], n_tot)
        assert (res["n_variants"][s] == n_var).all(), (lo, res["n_variants"][s], n_var)
        checked += len(keys)
    assert checked > 5_000_000


def test_c2_idempotent_and_chunking_invariant(ctx, native, c2):
    """Same plan twice -> identical table; a 4x finer chunking -> identical table."""
    plan, res = c2["plan"], c2["res"]
    plan.run()
    again = plan.results()
    keys = ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant")
    for k in keys:
        assert (again[k] == res[k]).all(), k
    sp2 = c2["grid"].read_sim_params(c2["cfg"], mean_family_size=5.0, family_model=1, chunk_log2=27)
    fine = native.Plan(ctx, c2["cell_id"], c2["af"], c2["depth"], sp2, c2["up"], neg_af_max=1e-4, test_type=native.TEST_POISSON, alpha=0.05)
    assert fine.n_chunks >= 3 * plan.n_chunks and fine.n_reads == plan.n_reads
    fine.run()
    r2 = fine.results()
    fine.close()
    for k in keys:
        assert (r2[k] == res[k]).all(), k
    assert r2["mean_error_rate"] == res["mean_error_rate"]


def test_c2_cell_counts_do_not_depend_on_the_batch(ctx, native, c2):
    """A cell's reads, families and consensus counts are a function of (seed, cell id) only: any
    sub-batch -- here every 7th cell, in reverse order -- reproduces the rows of the full run (this
    is what lets the replicate axis shard over GPUs without a collective)."""
    sel = np.arange(len(c2["depth"]))[::7][::-1].copy()
    sub = native.Plan(ctx, c2["cell_id"][sel], c2["af"][sel], c2["depth"][sel], c2["sp"], c2["up"], neg_af_max=1e-4,
                      test_type=native.TEST_POISSON, alpha=0.05)
    sub.run()
    r = sub.results()
    sub.close()
    for k in ("n_reads", "n_families", "n_total", "n_variants"):
        assert (r[k] == c2["res"][k][sel]).all(), k


def test_sort_5e7_pairs_sorted_and_stable(ctx):
    """K4 at 5e7 (key, value) pairs: output keys ascending, the multiset preserved (key-sum and
    value-xor checksums), equal keys keep their input order (values are the input positions)."""
    rng = np.random.default_rng(3)
    n = 50_000_000
    keys = rng.integers(0, 1 << 22, n, dtype=np.uint64) | (rng.integers(0, 2000, n, dtype=np.uint64) << np.uint64(32))
    vals = np.arange(n, dtype=np.uint32)
    sk, sv = ctx.sort_pairs(keys, vals)
    assert (sk[1:] >= sk[:-1]).all()
    assert int(sk.sum(dtype=np.uint64)) == int(keys.sum(dtype=np.uint64))
    assert np.bitwise_xor.reduce(sv) == np.bitwise_xor.reduce(vals)
    assert (keys[sv] == sk).all()                                          # it is a permutation of the input pairs
    same = sk[1:] == sk[:-1]
    assert (sv[1:][same] > sv[:-1][same]).all()                            # stable
