"""Host-side contract tests (no GPU): config hashing, tolerant loader, cell grid order, artifact
validation, manifest format, CLI wiring, packing helpers.  Modelled on the reference's
tests/test_determinism_hashes.py, tests/test_artifact_contract.py and tests/test_rng_api.py."""
import ast
import os
import json
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "precise-mrd-mini_b200" / "precise_mrd"


def test_config_hash_matches_reference_value():
    from precise_mrd.config import load_config

    cfg = load_config(PKG / "assets" / "configs" / "smoke.yaml")
    cfg.seed = 7
    assert cfg.config_hash() == "0b42f9b03f3fd555"
    assert cfg.fastq is None and "fastq" in cfg.to_dict()


def test_default_yaml_with_unknown_keys_loads():
    from precise_mrd.config import load_config, section

    cfg = load_config(PKG / "assets" / "configs" / "default.yaml")      # carries max_edit_distance, min_alt_umi, lob_replicates
    assert cfg.simulation.n_replicates == 1000 and cfg.umi.max_family_size == 1000
    assert section(cfg, "umi")["min_family_size"] == 3
    # dict sub-sections are accepted too (tests/test_rng_api.py:151-172)
    class C:  # noqa: D401
        umi = {"min_family_size": 2}
    assert section(C(), "umi")["min_family_size"] == 2


def test_grid_cells_meshgrid_order_and_global_ids():
    from precise_mrd.grid import grid_cells

    cid, af, depth, rep = grid_cells([0.1, 0.2], [10, 20, 30], 2)
    a, d, r = np.meshgrid([0.1, 0.2], [10, 20, 30], np.arange(2), indexing="ij")
    assert (af == a.ravel()).all() and (depth == d.ravel()).all() and (rep == r.ravel()).all()   # simulate.py:133-141
    assert (cid == np.arange(12)).all()
    # sharding the replicate axis over 2 ranks covers every global cell id exactly once
    c0 = grid_cells([0.1, 0.2], [10, 20, 30], 2, rep_offset=0, rep_stride=4)[0]
    c1 = grid_cells([0.1, 0.2], [10, 20, 30], 2, rep_offset=2, rep_stride=4)[0]
    assert sorted(np.concatenate([c0, c1]).tolist()) == list(range(24))


def test_manifest_and_validation_roundtrip(tmp_path):
    from precise_mrd.determinism_utils import hash_file, write_manifest
    from precise_mrd.validation import assert_hashes_stable, validate_artifacts

    reports = tmp_path / "reports"
    reports.mkdir()
    metrics = {"schema_version": "1.0.0", "roc_auc": 0.5, "average_precision": 0.3, "detected_cases": 0, "total_cases": 2, "calibration": []}
    ctx = {"schema_version": "1.0.0", "seed": 7, "timestamp": "2024-10-03T12:00:00", "config_hash": "x", "git_sha": "y", "python_version": "3"}
    (reports / "metrics.json").write_text(json.dumps(metrics))
    (reports / "run_context.json").write_text(json.dumps(ctx))
    write_manifest([reports / "metrics.json", reports / "run_context.json"], reports / "hash_manifest.txt")
    line = (reports / "hash_manifest.txt").read_text().splitlines()[0]
    assert line == f"{hash_file(reports / 'metrics.json')}  {reports / 'metrics.json'}"           # "<sha256>  <path>"
    validate_artifacts(reports)
    (reports / "lob.json").write_text(json.dumps({"schema_version": "1.0.0", "lob_value": -1}))
    with pytest.raises(Exception):
        validate_artifacts(reports)
    (reports / "lob.json").unlink()
    (reports / "lod_table.csv").write_text("depth,lod_af\n1000,0.1\n")
    with pytest.raises(ValueError):
        validate_artifacts(reports)
    (reports / "lod_table.csv").unlink()
    write_manifest([reports / "metrics.json"], reports / "m2.txt")
    with pytest.raises(AssertionError):
        assert_hashes_stable(reports / "hash_manifest.txt", reports / "m2.txt")
    with pytest.raises(AssertionError):
        validate_artifacts(tmp_path / "nope")


def test_cli_surface():
    from click.testing import CliRunner

    from precise_mrd.cli import main

    r = CliRunner().invoke(main, ["--help"])
    for cmd in ("smoke", "determinism", "eval-lob", "eval-lod", "eval-loq", "eval-contamination"):
        assert cmd in r.output
    r = CliRunner().invoke(main, ["--config", "/nonexistent.yaml", "smoke"])
    assert r.exit_code != 0 and "Configuration file not found" in r.output                        # cli.py:51


def test_no_legacy_global_rng_in_package():
    """tests/test_rng_api.py:9-114: no np.random.<legacy> calls; stages take an explicit rng."""
    banned = {"seed", "rand", "randn", "randint", "random_sample", "choice", "normal", "uniform", "poisson", "binomial", "RandomState"}
    for path in PKG.rglob("*.py"):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and node.attr in banned:
                v = node.value
                if isinstance(v, ast.Attribute) and v.attr == "random" and isinstance(v.value, ast.Name) and v.value.id in ("np", "numpy"):
                    raise AssertionError(f"{path}: np.random.{node.attr}")


def test_pack_unpack_bases():
    from precise_mrd.umi import pack_bases, unpack_bases

    seqs = ["ACGT", "TTTT", "AAAA", "CAGT"]
    p = pack_bases(seqs)
    assert [unpack_bases(v, 4) for v in p] == seqs
    assert list(np.argsort(p)) == list(np.argsort(seqs))
    with pytest.raises(ValueError):
        pack_bases(["ACGN"])


def test_metrics_oracle_matches_sklearn_and_calibration_bins():
    """The restatement the GPU metrics kernel (K12) is checked against, pinned to sklearn itself."""
    from sklearn.metrics import average_precision_score, roc_auc_score

    from oracle import port
    from precise_mrd.metrics import calibration_analysis

    rng = np.random.default_rng(0)
    for _ in range(20):
        y = rng.integers(0, 2, 120)
        s = np.round(rng.random(120), 1)          # heavy ties
        assert abs(port.roc_auc(y, s) - roc_auc_score(y, s)) < 1e-12
        assert abs(port.average_precision(y, s) - average_precision_score(y, s)) < 1e-12
    assert port.roc_auc(np.zeros(5), np.arange(5)) == 0.5                      # metrics.py:18-24 fallback
    assert port.average_precision(np.zeros(5), np.arange(5)) == 0.0            # metrics.py:27-33 fallback
    a, _ = port.bootstrap_metric(y, s, port.roc_auc, 50, np.random.default_rng(3))
    b, _ = port.bootstrap_metric(y, s, port.roc_auc, 50, np.random.default_rng(3))
    assert a == b and a["lower"] <= a["mean"] <= a["upper"]                    # test_lob_lod_bootstrap.py:148-168
    cal = calibration_analysis(np.array([0, 1, 1]), np.array([0.05, 0.95, 1.0]))
    assert cal[-1]["bin"] == 9 and cal[-1]["count"] == 2                       # last bin closed


def test_bench_workloads_and_sharding_keys():
    """bench.py's workloads: the C2 grid shards the replicate axis with GLOBAL cell ids (so a cell's Philox stream does not
    depend on the rank), the C5 panel gives each rank one sample of 1e5 loci with the AF strata by locus mod 4."""
    import importlib
    import sys as _sys

    root = str(Path(__file__).resolve().parent.parent)
    if root not in _sys.path:
        _sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    os.environ["PMRD_BENCH_REPS"] = "3"
    try:
        ids = []
        for rank in range(2):
            cfg, label, cid, af, depth = bench.workload("default", rank, 2)
            assert len(cid) == 5 * 4 * 3 and "C2" in label
            ids.append(cid)
        assert len(np.intersect1d(ids[0], ids[1])) == 0                      # ranks own disjoint global cells
        _, _, one, _, _ = bench.workload("default", 0, 1)
        assert len(one) == 60
    finally:
        os.environ.pop("PMRD_BENCH_REPS", None)
    os.environ["PMRD_PANEL_LOCI"] = "1000"
    try:
        cfg, label, cid, af, depth = bench.workload("panel", 3, 8)
        assert len(cid) == 1000 and (depth == 250).all() and (cid >> 20 == 3).all()
        assert (af == np.array([0.0, 1e-4, 1e-3, 1e-2])[np.arange(1000) % 4]).all()
    finally:
        os.environ.pop("PMRD_PANEL_LOCI", None)


def test_bench_external_workload_and_reference_arms(tmp_path):
    """bench.py --workload external: the NumPy read generator (both arms build the same arrays), the oracle's per-locus table
    on them, the config dict shared by the two arms, and the two reference arms end to end (CPU only: `--impl reference`
    never touches a GPU)."""
    import importlib
    import json
    import subprocess
    import sys as _sys

    root = Path(__file__).resolve().parent.parent
    if str(root) not in _sys.path:
        _sys.path.insert(0, str(root))
    bench = importlib.import_module("bench")
    keys, pl = bench.external_reads(42, 50_000, 400)
    k2, p2 = bench.external_reads(42, 50_000, 400)
    assert (keys == k2).all() and (pl == p2).all() and keys.dtype == np.uint64 and pl.dtype == np.uint32
    assert (keys >> np.uint64(32)).max() < 400 and ((keys & np.uint64(0xFFFFFFFF)) < (1 << 24)).all()
    sizes = np.unique(keys, return_counts=True)[1]
    assert 4.0 < sizes.mean() < 6.5                                               # ~Poisson(5) families (zero-size ones never appear)
    q = (pl >> np.uint32(2)) & np.uint32(63)
    assert q.min() >= 10 and q.max() <= 40 and ((pl >> np.uint32(31)) == (pl & np.uint32(1))).all()
    from oracle import cport, port

    r = cport.collapse_groups(keys, pl, 400, 3, 1000, 20, 0.6)
    fam, grp = port.collapse_exact(keys, pl, 0, 3, 1000, 20, 0.6)
    assert (r["family_count"] == grp["family_count"]).all() and (r["consensus_count"] == grp["consensus_count"]).all()
    assert r["total_families"] == len(fam["key"])
    a, b = bench.external_config(1000, 10, 2), bench.external_config(1000, 10, 2)
    assert a == b and "EXTERNAL" in a["workload"]
    cfgd = bench.workload_config("default", "x", np.arange(3), np.array([5, 5, 5]), 1)
    assert cfgd == bench.workload_config("default", "x", np.arange(3), np.array([5, 5, 5]), 1) and cfgd["families_per_gpu_per_step"] == 15
    env = dict(os.environ, PMRD_REF_READS="200000", PMRD_EXT_LOCI="2000", PMRD_REF_CELLS="8", PMRD_BENCH_REPS="2")
    for extra in (["--workload", "external"], []):
        out = subprocess.run([_sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"] + extra,
                             capture_output=True, text=True, timeout=300, env=env, cwd=tmp_path)
        assert out.returncode == 0, out.stderr[-1000:]
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
        assert line["e2e"]["value"] == line["value"] and line["cpu_baseline"]["kind"] == "port"
        assert set(line) >= {"metric", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "dtype"}


def test_world_size_does_not_import_torch():
    """A CLI process must not pay torch's import (2.5 s) to learn that it is alone."""
    import subprocess
    import sys as _sys

    pkg = str(Path(__file__).resolve().parent.parent / "precise-mrd-mini_b200")
    code = ("import sys; sys.path.insert(0, %r); from precise_mrd import distributed as d; from precise_mrd.determinism_utils import "
            "set_global_seed; set_global_seed(7); assert d.world_size() == 1; assert 'torch' not in sys.modules; print('ok')" % pkg)
    out = subprocess.run([_sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-500:]
