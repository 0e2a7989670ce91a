#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE in the build container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (absent on the GPU box) and
``python oracle/build_ref.py`` first.  Outputs are small packed arrays -- inputs in the
device record layout of include/pmrd.h plus the reference's own outputs -- so the GPU
parity tests can run where the reference cannot travel.

  call_smoke_v1.npz / call_smoke_v2.npz   reference fixtures data/det_a/smoke, data/smoke/smoke
                                          (collapsed_umis + error_model -> mrd_calls), re-derived
                                          here by running the reference call_mrd and checked
                                          bit-equal to the committed mrd_calls.parquet
  stats_kat.npz                           poisson_test / binomial_test / BH from call.py on the
                                          KAT rows + seeded random grids
  ga_reads.npz                            G-A reads -> mrd.umi.collapse_reads (== tmp/families.parquet)
  gc_reads.npz                            G-C reads -> committed collapsed_umis / mrd_calls tables
  gb_reads.npz                            G-B SyntheticReadGenerator reads -> UMIProcessor(max_edit_distance=0)
  gb_greedy.npz                           the same reads with barcode errors -> UMIProcessor(max_edit_distance=1, 2) with the
                                          reference's deterministic greedy clusters (umi_clustering.py:59-101)
  gb_stats.json                           G-B StatisticalTester (one-sided tests, CIs, BH/Bonferroni/Holm),
                                          ContextAnalyzer and ErrorModel on seeded inputs
  gb_qc.json                              G-B QCAnalyzer report + FeatureEngineer.extract_features on seeded tables
  gb_filters.json                         G-B QualityFilter (Fisher strand bias, end-repair, depth, quality) on a seeded table
  fastq_mixed.fastq / fastq_mixed.json    a synthetic FASTQ with every header convention and parser edge
                                          case + the reference's read list / family table / validation for it
  contam_ref.json                         ContaminationSimulator.simulate_contamination_effects (reduced sweep) of the reference
  lod_fit.npz                             LODAnalyzer._fit_detection_curve/_bootstrap_lod_ci on given vectors
  gd_dist.npz                             distribution samples from the reference simulate/collapse (KS targets)
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
OUT = HERE.parent / "tests" / "golden"
BASE = {"A": 0, "C": 1, "G": 2, "T": 3}


def pack_seq(s: str) -> int:
    v = 0
    for ch in s:
        v = (v << 2) | BASE[ch]
    return v


def _import_ref(which: str):
    for k in [k for k in sys.modules if k == "precise_mrd" or k.startswith("precise_mrd.")]:
        del sys.modules[k]
    sys.path[:] = [p for p in sys.path if "/oracle/_ref/" not in p]
    sys.path.insert(0, str(HERE / "_ref" / "stubs"))
    sys.path.insert(0, str(HERE / "_ref" / which))


# ------------------------------------------------------------------------------- call stage
def golden_call(which: str, fixture: str, out_name: str) -> None:
    _import_ref(which)
    from precise_mrd.call import call_mrd
    from precise_mrd.config import load_config

    d = REF / "data" / fixture / "smoke"
    collapsed = pd.read_parquet(d / "collapsed_umis.parquet")
    em = pd.read_parquet(d / "error_model.parquet")
    want = pd.read_parquet(d / "mrd_calls.parquet")
    cfg = load_config(REF / "configs" / "smoke.yaml")
    cfg.seed = 7
    got = call_mrd(collapsed, em, cfg, np.random.default_rng(0))
    for col in ("n_variants", "n_total", "variant_fraction", "expected_variants", "fold_change", "p_value",
                "mean_quality", "mean_consensus", "p_adjusted", "significant"):
        assert (got[col].values == want[col].values).all(), (which, col)
    sid = collapsed["sample_id"].values
    assert (np.diff(sid) >= 0).all()
    uniq, start = np.unique(sid, return_index=True)
    seg = np.concatenate([start, [len(sid)]]).astype(np.int64)
    np.savez_compressed(
        OUT / out_name,
        seg_offsets=seg,
        sample_id=uniq.astype(np.int64),
        is_variant=collapsed["is_variant"].values.astype(np.uint8),
        quality=collapsed["quality_score"].values,
        agreement=collapsed["consensus_agreement"].values,
        allele_fraction=want["allele_fraction"].values,
        error_rate=em["error_rate"].values,
        mean_error_rate=np.float64(em["error_rate"].mean()),
        alpha=np.float64(0.05),
        **{f"want_{c}": want[c].values for c in ("n_variants", "n_total", "variant_fraction", "expected_variants",
                                                  "fold_change", "p_value", "mean_quality", "mean_consensus",
                                                  "p_adjusted", "significant")},
    )
    print("  wrote", out_name, len(sid), "rows", len(uniq), "samples")


def golden_stats() -> None:
    _import_ref("gd_v1")
    from precise_mrd.call import benjamini_hochberg_correction, binomial_test, poisson_test

    rng = np.random.default_rng(20240901)
    # Poisson: KAT rows + grid across 1e-4 .. 2e5 expectations, deep tails both sides
    pk = [5, 1, 5, 0, 10, 0, 3, 100, 0]
    pe = [1.0, 5.0, 5.0, 0.0, 0.1, -1.0, 1e-4, 1e-3, 1e5]
    m = 3000
    e = 10 ** rng.uniform(-4, 5.3, m)
    k = np.maximum(0, np.round(e + rng.normal(0, 1, m) * np.sqrt(e) * rng.choice([0.3, 1, 3, 8, 20], m))).astype(np.int64)
    k[:300] = rng.integers(0, 30, 300)
    pk = np.concatenate([np.array(pk, np.int64), k])
    pe = np.concatenate([np.array(pe, float), e])
    pp = np.array([poisson_test(int(a), float(b)) for a, b in zip(pk, pe)])
    # binomial
    bk = [10, 1, 10, 0, 0, 1, 10, 9, 0, 5]
    bn = [100, 100, 100, 0, 10, 10, 10, 10, 1, 5]
    bp = [0.01, 0.1, 0.1, 0.5, 0.0, 0.0, 1.0, 1.0, 0.5, 0.5]
    m = 2500
    n = (10 ** rng.uniform(0, 5.2, m)).astype(np.int64)
    pr = 10 ** rng.uniform(-5, -0.01, m)
    mu = n * pr
    kk = np.clip(np.round(mu + rng.normal(0, 1, m) * np.sqrt(mu * (1 - pr) + 0.3) * rng.choice([0.3, 1, 3, 8], m)), 0, n).astype(np.int64)
    bk = np.concatenate([np.array(bk, np.int64), kk])
    bn = np.concatenate([np.array(bn, np.int64), n])
    bp = np.concatenate([np.array(bp, float), pr])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bpv = np.array([binomial_test(int(a), int(b), float(c)) for a, b, c in zip(bk, bn, bp)])
    # BH cases: KAT, ties, all-ones, all-zeros, single, random sizes
    cases = [np.array([0.001, 0.005, 0.01, 0.03, 0.05, 0.1]), np.array([0.5]), np.ones(7), np.zeros(5),
             np.array([0.04, 0.04, 0.04, 0.9, 0.9]), np.array([1e-300, 0.0, 1.0, 0.049999, 0.05])]
    for size in (2, 17, 60, 257, 5000, 20000):
        p = rng.uniform(0, 1, size) ** rng.choice([1, 3, 8])
        p[rng.integers(0, size, size // 10)] = 1.0
        if size > 10:
            p[rng.integers(0, size, size // 10)] = p[0]
        cases.append(p)
    alphas = [0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.05, 0.1, 0.05, 0.01, 0.05, 0.2]
    bh = {}
    for i, (p, a) in enumerate(zip(cases, alphas)):
        rej, adj = benjamini_hochberg_correction(p, a)
        bh[f"bh{i}_p"] = p
        bh[f"bh{i}_alpha"] = np.float64(a)
        bh[f"bh{i}_rej"] = rej
        bh[f"bh{i}_adj"] = adj
    np.savez_compressed(OUT / "stats_kat.npz", pois_k=pk, pois_e=pe, pois_p=pp, bin_k=bk, bin_n=bn, bin_p=bp,
                        bin_pv=bpv, n_bh=np.int64(len(cases)), **bh)
    print("  wrote stats_kat.npz", len(pk), "poisson", len(bk), "binomial", len(cases), "BH cases")


# ------------------------------------------------------------------------------- G-A
def golden_ga() -> None:
    sys.path.insert(0, str(HERE / "_ref" / "ga"))
    from mrd.umi import collapse_reads

    reads = pd.read_parquet(REF / "precise-mrd-mini" / "simulated_reads" / "reads.parquet")
    fam, loc = collapse_reads(reads, min_family_size=2, max_distance=1)
    want_f = pd.read_parquet(REF / "precise-mrd-mini" / "tmp" / "families.parquet")
    want_l = pd.read_parquet(REF / "precise-mrd-mini" / "tmp" / "collapsed.parquet")
    for c in want_f.columns:
        assert (fam[c].astype(want_f[c].dtype).values == want_f[c].values).all(), c
    for c in ("alt_count", "total_count", "family_count"):
        assert (loc[c].values == want_l[c].values).all(), c
    # groups = sorted (chrom, start, pos) -- pandas groupby order (mrd/umi.py:65)
    gkeys = sorted(set(zip(reads.chrom, reads.start, reads.pos)))
    gid = {g: i for i, g in enumerate(gkeys)}
    g = np.array([gid[t] for t in zip(reads.chrom, reads.start, reads.pos)], np.uint64)
    umi = np.array([pack_seq(u) for u in reads.umi], np.uint64)
    keys = (g << np.uint64(32)) | umi
    payload = np.array([pack_seq(s) for s in reads.sequence], np.uint32) | (reads.true_alt.values.astype(np.uint32) << np.uint32(31))
    fg = np.array([gid[t] for t in zip(fam.chrom, fam.start, fam.pos)], np.uint64)
    loci = sorted(set(zip(reads.chrom, reads.pos)))
    lid = {l: i for i, l in enumerate(loci)}
    np.savez_compressed(
        OUT / "ga_reads.npz",
        keys=keys, payload=payload,
        group_locus=np.array([lid[(c, p)] for (c, s, p) in gkeys], np.int64),
        want_fam_key=(fg << np.uint64(32)) | np.array([pack_seq(u) for u in fam.umi_representative], np.uint64),
        want_fam_size=fam.family_size.values.astype(np.uint32),
        want_fam_alt=fam.alt_count.values.astype(np.uint32),
        want_fam_consensus=np.array([pack_seq(s) for s in fam.consensus], np.uint32),
        want_locus=np.array([lid[(c, p)] for c, p in zip(loc.chrom, loc.pos)], np.int64),
        want_locus_alt=loc.alt_count.values.astype(np.int64),
        want_locus_total=loc.total_count.values.astype(np.int64),
        want_locus_families=loc.family_count.values.astype(np.int64),
    )
    print("  wrote ga_reads.npz", len(keys), "reads", len(fam), "families", len(loc), "loci")


# ------------------------------------------------------------------------------- G-C
def golden_gc() -> None:
    d = REF / "test_results" / "run"
    reads = pd.read_parquet(d / "simulated_reads.parquet")
    fam = pd.read_parquet(d / "collapsed_umis.parquet")
    calls = pd.read_parquet(d / "mrd_calls.parquet")
    gk = sorted(set(zip(reads.sample_id, reads.variant_id, reads.depth)))
    gid = {g: i for i, g in enumerate(gk)}
    g = np.array([gid[t] for t in zip(reads.sample_id, reads.variant_id, reads.depth)], np.uint64)
    umi = np.array([int(u[3:]) for u in reads.umi_id], np.uint64)
    keys = (g << np.uint64(32)) | umi
    payload = (reads.allele.values == "ALT").astype(np.uint32) << np.uint32(31)
    # shuffle reads: grouping must not depend on arrival order
    perm = np.random.default_rng(5).permutation(len(keys))
    fk = (np.array([gid[t] for t in zip(fam.sample_id, fam.variant_id, fam.depth)], np.uint64) << np.uint64(32)) | \
        np.array([int(u[3:]) for u in fam.umi_id], np.uint64)
    assert (calls.umi_id.values == fam.umi_id.values).all()
    np.savez_compressed(
        OUT / "gc_reads.npz", keys=keys[perm], payload=payload[perm],
        want_fam_key=fk, want_ref=fam.ref_reads.values.astype(np.uint32), want_alt=fam.alt_reads.values.astype(np.uint32),
        want_size=fam.family_size.values.astype(np.uint32),
        call_alt=calls.alt_reads.values.astype(np.int64), call_total=calls.total_reads.values.astype(np.int64),
        call_pvalue=calls.pvalue.values, call_detected=calls.detected.values,
    )
    print("  wrote gc_reads.npz", len(keys), "reads", len(fam), "families")


# ------------------------------------------------------------------------------- G-B
def golden_gb() -> None:
    sys.path.insert(0, str(HERE / "_ref" / "gb"))
    from precise_mrd_gb.io import SyntheticReadGenerator
    from precise_mrd_gb.umi import UMIProcessor

    gen = SyntheticReadGenerator(seed=11)
    sites = gen.generate_target_sites(24)
    reads = []
    afs = [0.0, 0.01, 0.05, 0.2, 0.5, 0.9]
    for i, s in enumerate(sites):
        reads += gen.generate_reads_for_site(s, n_umi_families=400, allele_fraction=afs[i % 6],
                                             contamination_rate=0.01 if i % 5 == 0 else 0.0, mean_family_size=5.0)
    order = np.random.default_rng(3).permutation(len(reads))
    reads = [reads[i] for i in order]
    # a low-quality-heavy tail so the quality filter and the consensus threshold both bite
    for r in reads[::7]:
        r.quality = max(10, r.quality - 12)
    proc = UMIProcessor(min_family_size=3, max_family_size=1000, max_edit_distance=0, quality_threshold=20,
                        consensus_threshold=0.6)
    # all families before the size filter (process_reads minus filter_family_size, umi.py:205-239)
    site_groups = proc.group_umis_by_site(reads)
    fams = []
    for site_key, site_reads in site_groups.items():
        for umi_reads in proc.group_umis_with_distance(site_reads):
            ca, cq = proc.call_consensus(umi_reads)
            fams.append((site_key, umi_reads[0].umi, len(umi_reads), ca, cq))
    kept = proc.process_reads(reads)
    counts = proc.get_consensus_counts(kept)
    skeys = list(site_groups.keys())
    sid = {k: i for i, k in enumerate(skeys)}
    site_alt = {f"{s.chrom}:{s.pos}:{s.ref}": s.alt for s in sites}
    g = np.array([sid[f"{r.chrom}:{r.pos}:{r.ref}"] for r in reads], np.uint64)
    keys = (g << np.uint64(32)) | np.array([pack_seq(r.umi) for r in reads], np.uint64)
    payload = np.array([
        BASE[r.alt] | (r.quality << 2) | ((1 if r.strand == "+" else 0) << 8) | (r.read_position << 9) |
        ((1 if r.alt == site_alt[f"{r.chrom}:{r.pos}:{r.ref}"] else 0) << 31) for r in reads], np.uint64).astype(np.uint32)
    cc = np.zeros((len(skeys), 4), np.int64)
    for _, row in counts.iterrows():
        cc[sid[row.site_key], BASE[row.allele]] = row.consensus_count
    np.savez_compressed(
        OUT / "gb_reads.npz", keys=keys, payload=payload,
        want_fam_key=(np.array([sid[f[0]] for f in fams], np.uint64) << np.uint64(32)) | np.array([pack_seq(f[1]) for f in fams], np.uint64),
        want_fam_size=np.array([f[2] for f in fams], np.uint32),
        want_fam_allele=np.array([-1 if f[3] is None else BASE[f[3]] for f in fams], np.int32),
        want_fam_quality=np.array([np.nan if f[4] is None else f[4] for f in fams], np.float64),
        want_kept=np.int64(len(kept)),
        want_consensus_count=cc,
        site_alt=np.array([BASE[site_alt[k]] for k in skeys], np.int32),
        params=np.array([3, 1000, 20], np.int64), consensus_threshold=np.float64(0.6),
    )
    print("  wrote gb_reads.npz", len(reads), "reads", len(fams), "families", len(kept), "kept")


# ------------------------------------------------------------------------------- G-B greedy clustering
def golden_gb_greedy() -> None:
    """UMIProcessor with max_edit_distance = 1 (configs/default.yaml:16), the Hamming clusters taken from the
    reference's DETERMINISTIC greedy form: src/precise_mrd/umi_clustering.py:59-101 (cluster_umis_greedy -- unique
    UMIs ordered by (-count, umi); the first unused UMI seeds a cluster and absorbs every unused UMI within the
    distance).  umi.py:117-144 runs the same loop over list(set(...)) -- hash-order dependent -- so its iteration
    order is DEFINED here as that sorted order.  Everything else is the reference's own code:
    cluster reads in input order (umi.py:137-140), UMIFamily.umi = first read's UMI (umi.py:225), call_consensus,
    filter_family_size."""
    import importlib.util

    sys.path.insert(0, str(HERE / "_ref" / "gb"))
    from precise_mrd_gb.io import SyntheticReadGenerator
    from precise_mrd_gb.umi import UMIProcessor

    spec = importlib.util.spec_from_file_location("ref_umi_clustering", REF / "src" / "precise_mrd" / "umi_clustering.py")
    uc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(uc)

    gen = SyntheticReadGenerator(seed=23)
    sites = gen.generate_target_sites(16)
    reads = []
    afs = [0.0, 0.05, 0.3, 0.6]
    for i, s in enumerate(sites):
        reads += gen.generate_reads_for_site(s, n_umi_families=150 if i % 4 else 600, allele_fraction=afs[i % 4],
                                             mean_family_size=4.0)
    rng = np.random.default_rng(9)
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    # sequencing errors in the barcode: one substituted base in 12 % of the reads, two in 2 % (so that clusters of
    # several UMIs, chains a-b-c the greedy rule cuts, and equal-count ties all occur)
    for r in reads:
        u = rng.random()
        n_sub = 2 if u < 0.02 else (1 if u < 0.14 else 0)
        for _ in range(n_sub):
            pos = int(rng.integers(0, len(r.umi)))
            r.umi = r.umi[:pos] + "ACGT"[(BASE[r.umi[pos]] + 1 + int(rng.integers(0, 3))) % 4] + r.umi[pos + 1:]
    for r in reads[::9]:
        r.quality = max(10, r.quality - 12)
    out = {}
    skeys = None
    for d in (1, 2):
        proc = UMIProcessor(min_family_size=3, max_family_size=1000, max_edit_distance=d, quality_threshold=20,
                            consensus_threshold=0.6)
        clusterer = uc.OptimizedUMICluster(max_edit_distance=d)
        site_groups = proc.group_umis_by_site(reads)
        skeys = list(site_groups.keys())
        fams = []
        for site_key, site_reads in site_groups.items():
            for cluster_umis in clusterer.cluster_umis_greedy([r.umi for r in site_reads]):
                members = set(cluster_umis)
                umi_reads = [r for r in site_reads if r.umi in members]          # umi.py:137-140
                ca, cq = proc.call_consensus(umi_reads)
                fams.append((site_key, umi_reads[0].umi, len(umi_reads), ca, cq, len(cluster_umis)))
        sid = {k: i for i, k in enumerate(skeys)}
        out[f"want_fam_key_d{d}"] = (np.array([sid[f[0]] for f in fams], np.uint64) << np.uint64(32)) | \
            np.array([pack_seq(f[1]) for f in fams], np.uint64)
        out[f"want_fam_size_d{d}"] = np.array([f[2] for f in fams], np.uint32)
        out[f"want_fam_allele_d{d}"] = np.array([-1 if f[3] is None else BASE[f[3]] for f in fams], np.int32)
        out[f"want_fam_quality_d{d}"] = np.array([np.nan if f[4] is None else f[4] for f in fams], np.float64)
        out[f"want_cluster_umis_d{d}"] = np.array([f[5] for f in fams], np.int32)
        print(f"    d={d}: {len(fams)} clusters, {sum(f[5] > 1 for f in fams)} of several UMIs, "
              f"{sum(f[2] >= 3 for f in fams)} kept")
    sid = {k: i for i, k in enumerate(skeys)}
    site_alt = {f"{s.chrom}:{s.pos}:{s.ref}": s.alt for s in sites}
    g = np.array([sid[f"{r.chrom}:{r.pos}:{r.ref}"] for r in reads], np.uint64)
    keys = (g << np.uint64(32)) | np.array([pack_seq(r.umi) for r in reads], np.uint64)
    payload = np.array([
        BASE[r.alt] | (r.quality << 2) | ((1 if r.strand == "+" else 0) << 8) | (r.read_position << 9) |
        ((1 if r.alt == site_alt[f"{r.chrom}:{r.pos}:{r.ref}"] else 0) << 31) for r in reads], np.uint64).astype(np.uint32)
    np.savez_compressed(OUT / "gb_greedy.npz", keys=keys, payload=payload, params=np.array([3, 1000, 20], np.int64),
                        consensus_threshold=np.float64(0.6), **out)
    print("  wrote gb_greedy.npz", len(reads), "reads")


# ------------------------------------------------------------------------------- G-B stats / context / errors
def golden_gb_stats() -> None:
    """StatisticalTester (one-sided tests, BH / Bonferroni / Holm), ContextAnalyzer and ErrorModel
    of the reference, run on seeded inputs -> tests/golden/gb_stats.json (text: it holds strings)."""
    import json

    sys.path.insert(0, str(HERE.parent))                 # the statsmodels stand-in imports oracle.port
    sys.path.insert(0, str(HERE / "_ref" / "stubs"))
    sys.path.insert(0, str(HERE / "_ref" / "gb"))
    from precise_mrd_gb.context import ContextAnalyzer
    from precise_mrd_gb.errors import ErrorModel
    from precise_mrd_gb.stats import StatisticalTester

    rng = np.random.default_rng(2024)
    out = {}
    # scalar tests, every alternative
    rows = []
    for test in ("poisson", "binomial"):
        t = StatisticalTester(test_type=test)
        for _ in range(60):
            n = int(rng.integers(1, 60000))
            rate = float(10 ** rng.uniform(-6, -1))
            k = int(min(n, rng.poisson(n * rate * rng.choice([0.3, 1.0, 3.0, 10.0]))))
            for alt in ("greater", "less", "two-sided"):
                if test == "poisson":
                    r = t.poisson_test(k, rate * n, alternative=alt)
                    ci = None
                else:
                    r = t.binomial_test(k, n, rate, alternative=alt)
                    ci = [float(r.confidence_interval.low), float(r.confidence_interval.high)]
                rows.append({"test": test, "k": k, "n": n, "rate": rate, "alternative": alt, "statistic": float(r.statistic),
                             "pvalue": float(r.pvalue), "effect_size": None if r.effect_size is None else float(r.effect_size),
                             "ci": ci})
    rows.append({"test": "poisson", "k": 3, "n": 10, "rate": 0.0, "alternative": "greater", "statistic": 3.0, "pvalue":
                 float(StatisticalTester().poisson_test(3, 0.0).pvalue), "effect_size": None, "ci": None})
    out["scalar"] = rows
    # test_multiple_variants, both test types
    contexts = ["ACG", "CCG", "TCA", "GCA", "default"]
    m = 400
    depth = rng.integers(200, 80000, m)
    ctx = rng.choice(contexts + ["ZZZ"], m)
    rates = {"ACG": 3e-4, "CCG": 1e-4, "TCA": 5e-5, "GCA": 2.5e-4, "default": 1e-4}
    lam = depth * np.array([rates.get(c, 1e-4) for c in ctx]) * rng.choice([1.0, 1.0, 1.0, 4.0, 12.0], m)
    alt_count = np.minimum(rng.poisson(lam), depth)
    table = pd.DataFrame({"site_key": [f"chr1:{1000 + i}:A" for i in range(m)], "context": ctx, "alt_count": alt_count,
                          "total_depth": depth})
    out["table"] = {"site_key": table.site_key.tolist(), "context": table.context.tolist(),
                    "alt_count": [int(v) for v in alt_count], "total_depth": [int(v) for v in depth], "rates": rates}
    for test in ("poisson", "binomial"):
        res = StatisticalTester(test_type=test, alpha=0.05).test_multiple_variants(table, rates)
        out[f"multi_{test}"] = {c: (res[c].astype(float).tolist() if res[c].dtype.kind in "fiub" else [str(v) for v in res[c]])
                                for c in res.columns}
    # multiple_testing_correction, three methods
    p = np.concatenate([rng.uniform(0, 1, 150), 10 ** rng.uniform(-9, -2, 50)])
    out["mtc"] = {"p": p.tolist()}
    for method in ("benjamini_hochberg", "bonferroni", "holm"):
        rej, q = StatisticalTester(alpha=0.05).multiple_testing_correction(p.tolist(), method)
        out["mtc"][method] = {"rejected": [bool(v) for v in rej], "corrected": [float(v) for v in q]}
    # ContextAnalyzer
    bases = "ACGT"
    norm = []
    for a in bases:
        for b in bases:
            for c in bases:
                for r in bases:
                    for al in bases:
                        norm.append([a + b + c, r, al, *ContextAnalyzer.normalize_context(a + b + c, r, al)])
    out["normalize"] = norm
    n_rows = 600
    cons = pd.DataFrame({
        "context": rng.choice([x + y + z for x in bases for y in bases for z in bases], n_rows),
        "ref": rng.choice(list(bases), n_rows), "allele": rng.choice(list(bases), n_rows),
        "consensus_count": rng.integers(1, 5000, n_rows)})
    an = ContextAnalyzer()
    rates_est = an.estimate_context_error_rates(cons)
    lookups = [["ACG", "C", "T"], ["CGT", "G", "A"], ["TTT", "T", "G"], ["AAA", "A", "C"], ["GCA", "C", "A"], ["NCN", "C", "T"]]
    out["context"] = {"table": {c: ([int(v) for v in cons[c]] if c == "consensus_count" else [str(v) for v in cons[c]]) for c in cons.columns},
                      "rates": rates_est, "lookups": [[*q, float(an.get_expected_error_rate(*q))] for q in lookups],
                      "lookup_unfitted": float(ContextAnalyzer().get_expected_error_rate("ACG", "C", "T"))}
    em = ErrorModel()
    out["background"] = {"min_depth": 100, "rates": em.estimate_background_errors(cons, min_depth=100)}
    with open(OUT / "gb_stats.json", "w") as f:
        json.dump(out, f)
    print("  wrote gb_stats.json", len(rows), "scalar rows,", m, "loci")


# ------------------------------------------------------------------------------- G-B quality filters
def golden_filters() -> None:
    """QualityFilter of the reference (Fisher strand bias, end-repair enrichment, depth, quality) on
    a seeded variant table -> tests/golden/gb_filters.json."""
    import json

    sys.path.insert(0, str(HERE / "_ref" / "gb"))
    from precise_mrd_gb.filters import QualityFilter

    rng = np.random.default_rng(404)
    m = 500
    depth = rng.integers(4, 30000, m)
    alt = np.minimum(rng.poisson(depth * rng.choice([1e-4, 1e-3, 1e-2, 0.2], m)), depth)
    af = np.array([rng.binomial(a, rng.choice([0.5, 0.5, 0.5, 0.02, 0.97])) for a in alt])
    rf = np.array([rng.binomial(d - a, 0.5) for d, a in zip(depth, alt)])
    rows = []
    for i in range(m):
        n_pos = int(alt[i])
        ends = rng.random() < 0.3
        pos = np.where(rng.random(n_pos) < (0.7 if ends else 0.13), rng.choice(np.r_[np.arange(1, 11), np.arange(140, 151)], n_pos),
                       rng.integers(11, 140, n_pos))
        ref, al = rng.choice([("G", "T"), ("C", "A"), ("C", "T"), ("A", "G")])
        q = rng.integers(8, 41, int(rng.integers(0, 12))).tolist()
        rows.append({"alt_count": int(alt[i]), "total_depth": int(depth[i]), "alt_forward": int(af[i]), "alt_reverse": int(alt[i] - af[i]),
                     "ref_forward": int(rf[i]), "ref_reverse": int(depth[i] - alt[i] - rf[i]), "ref": str(ref), "alt": str(al),
                     "read_positions": [int(v) for v in pos], "quality_scores": q})
    rows[0].update(alt_forward=1, alt_reverse=0, ref_forward=1, ref_reverse=1)          # < 4 reads: no strand test
    rows[1].update(alt_forward=0, alt_reverse=0, alt_count=0)                            # empty alt row: odds nan, p 1
    qf = QualityFilter()
    df = pd.DataFrame(rows)
    out_df = qf.filter_variant_dataframe(df)
    scal = []
    for r in rows[:120]:
        sb = qf.test_strand_bias(r["alt_forward"], r["alt_reverse"], r["ref_forward"], r["ref_reverse"])
        er = qf.test_end_repair_artifact(r["ref"], r["alt"], r["read_positions"])
        scal.append({"sb": [bool(sb.passed), sb.reason, None if sb.statistic is None else float(sb.statistic), None if sb.pvalue is None else float(sb.pvalue)],
                     "er": [bool(er.passed), er.reason, None if er.statistic is None else float(er.statistic), None if er.pvalue is None else float(er.pvalue)]})
    rep = qf.generate_filter_report(out_df)
    out = {"rows": rows, "scalar": scal,
           "table": {c: [None if (v is None or (isinstance(v, float) and np.isnan(v))) else (bool(v) if isinstance(v, (bool, np.bool_)) else v)
                         for v in out_df[c]] for c in out_df.columns if c.endswith("_filter") or c.endswith("_reason") or c == "all_filters_passed"},
           "report": {k: (int(v) if isinstance(v, (np.integer, int)) and not isinstance(v, bool) else (float(v) if isinstance(v, (float, np.floating)) else v))
                      for k, v in rep.items()}}
    with open(OUT / "gb_filters.json", "w") as f:
        json.dump(out, f)
    print("  wrote gb_filters.json", m, "variants;", int(out_df.all_filters_passed.sum()), "pass;", rep["failure_reasons"])


# ------------------------------------------------------------------------------- QC report + feature table
def golden_qc() -> None:
    """QCAnalyzer (src/precise_mrd/qc.py) and FeatureEngineer.extract_features (advanced_stats.py:33-83) of
    the reference on seeded tables -> tests/golden/gb_qc.json."""
    import json

    sys.path.insert(0, str(HERE / "_ref" / "stubs"))
    sys.path.insert(0, str(HERE / "_ref" / "gb"))
    from precise_mrd_gb.advanced_stats import FeatureEngineer
    from precise_mrd_gb.qc import QCAnalyzer, QCThresholds

    rng = np.random.default_rng(77)
    n_sites, fams, cons = 14, [], []
    for s in range(n_sites):
        key = f"chr{1 + s % 5}:{1000 + 37 * s}:{'ACGT'[s % 4]}"
        ref = "ACGT"[s % 4]
        nf = int(rng.integers(40, 900)) if s != 3 else 25
        sizes = np.clip(rng.poisson(rng.choice([1.5, 5.0, 9.0]), nf), 1, None)
        sizes[rng.random(nf) < 0.01] = int(rng.integers(101, 1500))
        counts = {}
        for f in range(nf):
            has = rng.random() < (0.93 if s != 5 else 0.6)
            al = ref if rng.random() > 0.02 else "ACGT"[(s + 1 + int(rng.integers(0, 3))) % 4]
            fams.append({"site_key": key, "family_size": int(sizes[f]), "consensus_allele": (al if has else None)})
            if has and sizes[f] >= 3:
                counts[al] = counts.get(al, 0) + 1
        for al, c in counts.items():
            cons.append({"site_key": key, "allele": al, "ref": ref, "consensus_count": int(c)})
    consensus = pd.DataFrame(cons)
    qa = QCAnalyzer(QCThresholds())

    def clean(o):
        if isinstance(o, dict):
            return {str(k): clean(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [clean(v) for v in o]
        if isinstance(o, (np.integer,)):
            return int(o)
        if isinstance(o, (np.floating,)):
            return float(o)
        if isinstance(o, np.bool_):
            return bool(o)
        return o

    keys = sorted({f["site_key"] for f in fams})
    out = {"site_keys": keys,
           "families": {"site": [keys.index(f["site_key"]) for f in fams], "size": [f["family_size"] for f in fams],
                        "allele": "".join(f["consensus_allele"] or "-" for f in fams)},
           "consensus": cons,
           "family_size": clean(qa.analyze_family_size_distribution([f["family_size"] for f in fams])),
           "family_size_const": clean(qa.analyze_family_size_distribution([4, 4, 4])),
           "depth": clean(qa.calculate_depth_metrics(consensus)),
           "contamination": clean(qa.estimate_contamination_metrics(consensus)),
           "efficiency": clean(qa.analyze_consensus_efficiency(fams)),
           "report": clean(qa.generate_comprehensive_qc_report(consensus, fams, {"seed": 77}))}
    # feature table of the collapsed-family frame (G-D columns)
    m = 400
    df = pd.DataFrame({"family_size": rng.integers(1, 40, m), "quality_score": rng.normal(28, 4, m).clip(10, 40),
                       "consensus_agreement": rng.uniform(0.6, 1.0, m), "allele_fraction": rng.choice([0.0, 1e-4, 1e-3, 0.01, 0.05], m)})
    X, names = FeatureEngineer(None).extract_features(df)
    out["features"] = {"frame": {c: [float(v) if c != "family_size" else int(v) for v in df[c]] for c in df.columns},
                       "names": list(names), "matrix_hex": [[float(v).hex() for v in row] for row in X.to_numpy()]}
    X2, names2 = FeatureEngineer(None).extract_features(df[["family_size", "allele_fraction"]])
    out["features_partial"] = {"names": list(names2), "shape": list(X2.shape)}
    with open(OUT / "gb_qc.json", "w") as f:
        json.dump(out, f)
    print("  wrote gb_qc.json", len(fams), "families,", len(cons), "consensus rows; valid:", out["report"]["clinical_validation"]["sample_valid"])


# ------------------------------------------------------------------------------- FASTQ ingest
def golden_fastq() -> None:
    """A seeded synthetic FASTQ exercising every header convention and parser edge case
    (tests/golden/fastq_mixed.fastq) + what the reference's fastq.py makes of it (fastq_mixed.json)."""
    import json

    _import_ref("gd_v2")
    from precise_mrd.config import load_config
    from precise_mrd.fastq import FASTQReader, detect_umi_format, process_fastq_to_dataframe

    rng = np.random.default_rng(77)
    bases = np.array(list("ACGT"))
    umis = ["".join(rng.choice(bases, int(rng.integers(8, 13)))) for _ in range(60)]

    def rec(i, header):
        n = int(rng.integers(30, 90))
        seq = "".join(rng.choice(bases, n))
        qual = "".join(chr(33 + int(q)) for q in np.clip(rng.normal(28, 9, n), 2, 41).astype(int))
        return [header, seq, "+", qual]

    lines = []
    for i in range(900):
        u = umis[int(rng.integers(0, len(umis)))]
        style = int(rng.integers(0, 9))
        if style <= 2:
            h = f"@READ_{i:05d} UMI:{u}"
        elif style == 3:
            h = f"@INST:{i}:FC_{u}_L001 1:N:0"
        elif style == 4:
            h = f"@INST:{i}:FC:1:{i % 97}:{u}"
        elif style == 5:
            h = f"@READ_{i:05d} no umi here"                     # no pattern matches
        elif style == 6:
            h = f"@R{i} UMI: UMI:{u} extra"                      # first 'UMI:' has no capital after it
        elif style == 7:
            h = f"@R{i}_{u}X_tail:{u}"                            # '_' run may be 9..13 letters; ':' tail decides
        else:
            h = f"  @READ_{i:05d} UMI:{u}ACGTNNNNNNNNNNNN"[: 24 + len(u)] + "  "   # padded, capture cut to <= 12 letters
        r = rec(i, h)
        if i % 131 == 5:
            r[1] = "   "                                          # malformed: whitespace-only sequence line
        if i % 171 == 9:
            r = [x + "\r" for x in r]                             # CRLF record
        lines += r
    lines += ["", "@AFTER_BLANK UMI:AAAAAAAA", "ACGT", "+", "IIII"]   # an empty header line ends the file
    text = "\n".join(lines) + "\n"
    path = OUT / "fastq_mixed.fastq"
    path.write_text(text)
    cfg = load_config(REF / "configs" / "smoke.yaml")
    out = {}
    reader = FASTQReader(path)
    reads = list(reader.read_fastq())
    out["reads"] = [{"read_id": r["read_id"], "sequence": r["sequence"], "umi": r["umi"], "mean_quality": float(r["mean_quality"]),
                     "sequence_length": r["sequence_length"], "qual_sum": int(sum(r["quality_scores"]))} for r in reads]
    out["reads_max_50"] = [r["read_id"] for r in reader.read_fastq(max_reads=50)]
    out["validate"] = {k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in reader.validate_fastq().items()}
    for tag, mr in (("all", None), ("max_300", 300)):
        df = process_fastq_to_dataframe(path, cfg, np.random.default_rng(0), max_reads=mr)
        out[f"table_{tag}"] = {"sample_id": [int(v) for v in df.sample_id], "umi": df.umi.tolist(), "family_size": [int(v) for v in df.family_size],
                               "consensus_sequence": df.consensus_sequence.tolist(), "mean_quality": [float(v) for v in df.mean_quality],
                               "quality_scores_sum": [int(sum(v)) for v in df.quality_scores],
                               "quality_scores_head": [[int(x) for x in v[:5]] for v in df.quality_scores],
                               "consensus_agreement": [float(v) for v in df.consensus_agreement],
                               "passes_quality": [bool(v) for v in df.passes_quality],
                               "passes_consensus": [bool(v) for v in df.passes_consensus], "columns": list(df.columns)}
    out["format"] = detect_umi_format(path, 100)
    with open(OUT / "fastq_mixed.json", "w") as f:
        json.dump(out, f)
    print("  wrote fastq_mixed.fastq / .json:", len(reads), "reads,", len(out["table_all"]["umi"]), "families, format", out["format"])


# ------------------------------------------------------------------------------- contamination (C4)
def golden_contam() -> None:
    """The reference's own ContaminationSimulator (src/precise_mrd/sim/contamination.py:33-311, G-D v1 snapshot) on a
    reduced sweep: 2 rates per kind x AF {0.001, 0.01} x depth 1000 x 2 replicates.  call_mrd is wrapped to add the
    `variant_call` column the reference never produces -- without it every eval-* command dies on
    a KeyError.  Output: tests/golden/contam_ref.json (the result dict, keys stringified)."""
    import json

    _import_ref("gd_v1")
    import precise_mrd.call as ref_call
    import precise_mrd.sim.contamination as ref_contam
    from precise_mrd.config import load_config

    real = ref_call.call_mrd

    def call_with_variant_call(*a, **k):
        df = real(*a, **k)
        if len(df) and "variant_call" not in df:
            df["variant_call"] = df["significant"]
        return df

    ref_contam.call_mrd = call_with_variant_call
    cfg = load_config(REF / "configs" / "smoke.yaml")
    args = dict(hop_rates=[0.0, 0.01], barcode_collision_rates=[0.0, 0.001], cross_sample_proportions=[0.0, 0.1],
                af_test_values=[0.001, 0.01], depth_values=[1000], n_replicates=2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = ref_contam.ContaminationSimulator(cfg, np.random.default_rng(cfg.seed))
        res = sim.simulate_contamination_effects(**args)

    def clean(o):
        if isinstance(o, dict):
            return {str(k): clean(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [clean(v) for v in o]
        if isinstance(o, (np.floating, np.integer)):
            return o.item()
        return o

    with open(OUT / "contam_ref.json", "w") as f:
        json.dump({"args": args, "seed": int(cfg.seed), "results": clean(res)}, f, indent=1)
    flat = [v["mean_sensitivity"] for kind in ("index_hopping", "barcode_collisions", "cross_sample_contamination")
            for r in res[kind].values() for a in r.values() for v in a.values()]
    print("  wrote contam_ref.json; mean sensitivities:", sorted(set(flat)))


# ------------------------------------------------------------------------------- LoD fit
def golden_lod_fit() -> None:
    _import_ref("gd_v1")
    from precise_mrd.config import load_config
    from precise_mrd.eval.lod import LODAnalyzer

    cfg = load_config(REF / "configs" / "smoke.yaml")
    af = np.logspace(-4, -2, 15)
    curves = {
        "mid": [0, 0, 0, 0, 0, .04, 0, .08, .2, .4, .6, .8, .92, 1, 1],
        "steep": [0, 0, 0, 0, 0, 0, 0, 0, .04, .52, .96, 1, 1, 1, 1],
        "shallow": [.04, .08, .08, .16, .2, .28, .4, .44, .56, .6, .72, .8, .84, .92, .96],
        "zeros": [0] * 15,
        "ones": [1] * 15,
    }
    out = {"af": af}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, hit in curves.items():
            an = LODAnalyzer(cfg, np.random.default_rng(123))
            lod = an._fit_detection_curve(list(af), list(hit), 0.95)
            ci = an._bootstrap_lod_ci(list(af), list(hit), 0.95, n_bootstrap=200)
            out[f"{name}_hit"] = np.array(hit, float)
            out[f"{name}_lod"] = np.float64(lod)
            out[f"{name}_ci"] = np.array(ci, float)
            print("   ", name, lod, ci)
    np.savez_compressed(OUT / "lod_fit.npz", **out)
    print("  wrote lod_fit.npz")


# ------------------------------------------------------------------------------- distributions
def golden_gd_dist() -> None:
    """Samples of the reference's G-D generator outputs: targets of the KS / chi-square gate."""
    out = {}
    for which in ("gd_v1", "gd_v2"):
        _import_ref(which)
        from precise_mrd.collapse import collapse_umis
        from precise_mrd.config import LODConfig, PipelineConfig, SimulationConfig, StatsConfig, UMIConfig
        from precise_mrd.simulate import simulate_reads

        cfg = PipelineConfig(run_id="dist", seed=1, simulation=SimulationConfig([0.02, 0.001, 0.0001], [4000], 8, 10),
                             umi=UMIConfig(3, 1000, 20, 0.6), stats=StatsConfig("poisson", 0.05, "benjamini_hochberg"),
                             lod=LODConfig(0.95, 0.95))
        rng = np.random.default_rng(99)
        kw = {"use_cache": False} if which == "gd_v2" else {}
        reads = simulate_reads(cfg, rng, **kw)
        col = collapse_umis(reads, cfg, rng)
        tag = which[-2:]
        out[f"{tag}_cell_af"] = reads.allele_fraction.values
        out[f"{tag}_cell_bg"] = reads.background_rate.values
        out[f"{tag}_cell_mfs"] = reads.mean_family_size.values.astype(float)
        out[f"{tag}_cell_ntrue"] = reads.n_true_variants.values
        out[f"{tag}_cell_nfp"] = reads.n_false_positives.values
        out[f"{tag}_cell_mq"] = reads.mean_quality.values
        out[f"{tag}_sample"] = col.sample_id.values.astype(np.int32)
        out[f"{tag}_family_size"] = col.family_size.values.astype(np.int32)
        out[f"{tag}_quality"] = col.quality_score.values.astype(np.float32)
        out[f"{tag}_agreement"] = col.consensus_agreement.values.astype(np.float32)
        out[f"{tag}_flags"] = (col.passes_quality.values.astype(np.uint8) | (col.passes_consensus.values.astype(np.uint8) << 1) |
                               (col.is_variant.values.astype(np.uint8) << 2))
        print("   ", which, len(reads), "cells", len(col), "families")
    np.savez_compressed(OUT / "gd_dist.npz", **out)
    print("  wrote gd_dist.npz")


def main() -> int:
    if not REF.exists():
        print("reference not present; golden vectors cannot be regenerated here")
        return 1
    OUT.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])
    steps = {
        "call": lambda: (golden_call("gd_v1", "det_a", "call_smoke_v1.npz"), golden_call("gd_v2", "smoke", "call_smoke_v2.npz")),
        "stats": golden_stats, "ga": golden_ga, "gc": golden_gc, "gb": golden_gb, "gbgreedy": golden_gb_greedy, "gbstats": golden_gb_stats, "filters": golden_filters, "qc": golden_qc, "fastq": golden_fastq,
        "lod": golden_lod_fit, "contam": golden_contam,
        "dist": golden_gd_dist,
    }
    for name, fn in steps.items():
        if only and name not in only:
            continue
        print("[make_golden]", name)
        fn()
    return 0


if __name__ == "__main__":
    sys.exit(main())
