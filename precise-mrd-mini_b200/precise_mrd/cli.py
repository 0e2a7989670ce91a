"""``precise-mrd`` command line -- the G-D v2 surface (``src/precise_mrd/cli.py:178-351``):

    precise-mrd [--seed S] [--config C] smoke [--out-dir D]
    precise-mrd ... determinism | eval-lob [--n-blank N] | eval-lod [--replicates R]
                   | eval-loq [--replicates R] | eval-contamination | eval-stratified

Group options come BEFORE the sub-command; artifacts go to ``reports/`` and ``data/`` under the
current directory; every command ends by echoing a JSON summary.  ``--engine grid`` switches the
``eval-*`` commands from the parity engine to the read-level fused device plan.

Several GPUs: launch the same command line under ``python -m torch.distributed.run --nproc-per-node N -m precise_mrd.cli
--engine grid eval-lod ...``; the group joins the NCCL process group, the grid engine shards its cells over the ranks
(``distributed.run_cells_sharded``: the table is the single-GPU one bit for bit) and rank 0 alone writes the artifacts."""
from __future__ import annotations

import json
import os
import shutil
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path
from typing import Dict, Optional

import click

from .config import LODConfig, PipelineConfig, SimulationConfig, StatsConfig, UMIConfig, load_config
from .determinism_utils import env_fingerprint, set_global_seed, write_manifest
from .utils import PipelineIO, json_ready
from .validation import assert_hashes_stable, validate_artifacts, validate_json

DEFAULT_CONFIG = Path(__file__).resolve().parent / "assets" / "configs" / "smoke.yaml"
REPORTS_DIR = Path("reports")
DATA_ROOT = Path("data")


@dataclass
class CLIContext:
    seed: int
    config_override: Optional[Path]
    engine: str = "stages"


def create_minimal_config(seed: int) -> PipelineConfig:
    return PipelineConfig(run_id="smoke_test", seed=seed,
                          simulation=SimulationConfig([0.01, 0.001, 0.0001], [1000, 5000], 10, 100),
                          umi=UMIConfig(3, 1000, 20, 0.6), stats=StatsConfig("poisson", 0.05, "benjamini_hochberg"),
                          lod=LODConfig(0.95, 0.95))


def _load_pipeline_config(config_option: Optional[Path], seed: int) -> PipelineConfig:
    if config_option:
        path = Path(config_option)
        if not path.exists():
            raise click.ClickException(f"Configuration file not found: {path}")              # cli.py:51
        try:
            config = load_config(path)
        except Exception as exc:
            raise click.ClickException(f"Failed to load configuration {path}: {exc}") from exc  # cli.py:60
    else:
        config = load_config(DEFAULT_CONFIG)
    config.seed = seed                                                                        # cli.py:63
    return config


def _run_smoke_pipeline(config: PipelineConfig, seed: int, output_dir: Path, context_dir: Optional[Path] = None) -> Dict[str, Path]:
    from .call import call_mrd
    from .collapse import collapse_umis
    from .error_model import fit_error_model
    from .metrics import calculate_metrics
    from .reporting import render_report
    from .simulate import simulate_reads

    rng = set_global_seed(seed, deterministic_ops=True)
    stage_dir = PipelineIO.ensure_dir(output_dir)
    reports = PipelineIO.ensure_dir(REPORTS_DIR)
    paths = {
        "reads": stage_dir / "simulated_reads.parquet", "collapsed": stage_dir / "collapsed_umis.parquet",
        "error_model": stage_dir / "error_model.parquet", "calls": stage_dir / "mrd_calls.parquet",
        "metrics": reports / "metrics.json", "run_context": reports / "run_context.json",
        "report": reports / "auto_report.html", "manifest": reports / "hash_manifest.txt",
    }
    reads_df = simulate_reads(config, rng, output_path=str(paths["reads"]))
    collapsed_df = collapse_umis(reads_df, config, rng, output_path=str(paths["collapsed"]))
    error_model_df = fit_error_model(collapsed_df, config, rng, output_path=str(paths["error_model"]))
    calls_df = call_mrd(collapsed_df, error_model_df, config, rng, output_path=str(paths["calls"]))
    metrics = json_ready(calculate_metrics(calls_df, rng, n_bootstrap=config.simulation.n_bootstrap if config.simulation else 100, config=config))
    metrics["schema_version"] = "1.0.0"
    run_context = {
        "schema_version": "1.0.0", "seed": seed,
        "timestamp": datetime(2024, 10, 3, 12, 0, 0).isoformat(),     # constant, for hash stability (cli.py:136)
        "config_hash": config.config_hash(),
        "cli_args": {"command": "smoke", "seed": seed, "config_run_id": config.run_id,
                     "output_dir": str(context_dir if context_dir is not None else stage_dir)},
        **env_fingerprint(),
    }
    PipelineIO.save_json(metrics, paths["metrics"])
    PipelineIO.save_json(run_context, paths["run_context"])
    render_report(calls_df, metrics, config.to_dict(), run_context, str(paths["report"]))
    write_manifest([paths["metrics"], paths["run_context"], paths["report"]], out_manifest=paths["manifest"])
    return paths


@click.group()
@click.option("--seed", default=7, show_default=True, type=int, help="Seed for deterministic runs.")
@click.option("--config", "config_path", type=click.Path(path_type=Path), help="Pipeline configuration (default: packaged smoke config).")
@click.option("--engine", type=click.Choice(["stages", "grid"]), default="stages", show_default=True,
              help="eval-*: 'stages' = reference-compatible per-cell pipelines, 'grid' = read-level fused device plan.")
@click.pass_context
def main(ctx: click.Context, seed: int, config_path: Optional[Path], engine: str) -> None:
    """Precise MRD (B200): deterministic MRD analytics with hardened artifact contracts."""
    ctx.obj = CLIContext(seed=seed, config_override=config_path, engine=engine)
    if engine == "grid" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        _join_process_group()


def _join_process_group() -> None:
    """Under ``torch.distributed.run``: one process per GPU over NCCL (the grid engine shards its cells over the ranks)."""
    import torch
    import torch.distributed as dist

    if dist.is_initialized():
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import atexit

    atexit.register(lambda: dist.is_initialized() and dist.destroy_process_group())


def _is_writer() -> bool:
    """Rank 0 of a multi-process launch (every rank holds the same tables), or the only process."""
    return int(os.environ.get("RANK", "0")) == 0 or int(os.environ.get("WORLD_SIZE", "1")) <= 1


@main.command("smoke")
@click.option("--out-dir", default=DATA_ROOT / "smoke", show_default=True, type=click.Path(path_type=Path))
@click.pass_obj
def smoke_cmd(ctx: CLIContext, out_dir: Path) -> None:
    """Run the fast deterministic smoke pipeline."""
    config = _load_pipeline_config(ctx.config_override, ctx.seed)
    artifacts = _run_smoke_pipeline(config, ctx.seed, out_dir)
    validate_artifacts(REPORTS_DIR)
    click.echo(json.dumps({"stage": "smoke", "seed": ctx.seed, "config": config.to_dict(),
                           "artifacts": {k: str(v) for k, v in artifacts.items()}}, indent=2))


@main.command("determinism")
@click.option("--out-dir", default=DATA_ROOT / "determinism", show_default=True, type=click.Path(path_type=Path))
@click.pass_obj
def determinism_cmd(ctx: CLIContext, out_dir: Path) -> None:
    """Run the smoke pipeline twice and assert identical artifacts."""
    config = _load_pipeline_config(ctx.config_override, ctx.seed)
    # both runs record the same output_dir in run_context.json, which is part of the manifest
    # (the reference records run1/ and run2/ there, so its own `determinism` can never pass)
    first = _run_smoke_pipeline(config, ctx.seed, out_dir / "run1", context_dir=out_dir)
    validate_artifacts(REPORTS_DIR)
    snap = Path(first["manifest"]).with_name("hash_manifest_run1.txt")
    shutil.copy2(first["manifest"], snap)
    second = _run_smoke_pipeline(config, ctx.seed, out_dir / "run2", context_dir=out_dir)
    validate_artifacts(REPORTS_DIR)
    assert_hashes_stable(snap, Path(second["manifest"]))
    click.echo(json.dumps({"stage": "determinism", "seed": ctx.seed, "status": "hashes-identical",
                           "artifacts": {"first_manifest": str(snap), "second_manifest": str(second["manifest"])}}, indent=2))


def _analyzer(ctx: CLIContext):
    from .eval.lod import LODAnalyzer

    config = _load_pipeline_config(ctx.config_override, ctx.seed)
    rng = set_global_seed(ctx.seed, deterministic_ops=True)
    PipelineIO.ensure_dir(REPORTS_DIR)
    return LODAnalyzer(config, rng, engine=ctx.engine)


def _echo(stage: str, ctx: CLIContext, **artifacts) -> None:
    click.echo(json.dumps({"stage": stage, "seed": ctx.seed, "engine": ctx.engine,
                           "artifacts": {k: str(v) for k, v in artifacts.items()}}, indent=2))


@main.command("eval-lob")
@click.option("--n-blank", default=50, show_default=True, type=int, help="Blank replicates.")
@click.pass_obj
def eval_lob_cmd(ctx: CLIContext, n_blank: int) -> None:
    """Estimate the Limit of Blank."""
    an = _analyzer(ctx)
    an.estimate_lob(n_blank_runs=n_blank)
    if _is_writer():
        an.generate_reports(str(REPORTS_DIR))
        _echo("eval-lob", ctx, lob=REPORTS_DIR / "lob.json")


@main.command("eval-lod")
@click.option("--replicates", default=25, show_default=True, type=int, help="Replicates per AF/depth.")
@click.pass_obj
def eval_lod_cmd(ctx: CLIContext, replicates: int) -> None:
    """Estimate the Limit of Detection."""
    an = _analyzer(ctx)
    an.estimate_lob(n_blank_runs=max(10, replicates // 2))
    an.estimate_lod(n_replicates=replicates)
    if _is_writer():
        an.generate_reports(str(REPORTS_DIR))
        _echo("eval-lod", ctx, lob=REPORTS_DIR / "lob.json", lod_table=REPORTS_DIR / "lod_table.csv", lod_curves=REPORTS_DIR / "lod_curves.png")


@main.command("eval-loq")
@click.option("--replicates", default=25, show_default=True, type=int, help="Replicates per AF/depth.")
@click.pass_obj
def eval_loq_cmd(ctx: CLIContext, replicates: int) -> None:
    """Estimate the Limit of Quantification."""
    an = _analyzer(ctx)
    an.estimate_lob(n_blank_runs=max(10, replicates // 2))
    an.estimate_loq(n_replicates=replicates)
    if _is_writer():
        an.generate_reports(str(REPORTS_DIR))
        _echo("eval-loq", ctx, lob=REPORTS_DIR / "lob.json", loq_table=REPORTS_DIR / "loq_table.csv")


@main.command("eval-contamination")
@click.pass_obj
def eval_contamination_cmd(ctx: CLIContext) -> None:
    """Index-hopping / barcode-collision / cross-sample stress sweep."""
    from .sim.contamination import run_contamination_stress_test

    config = _load_pipeline_config(ctx.config_override, ctx.seed)
    rng = set_global_seed(ctx.seed, deterministic_ops=True)
    PipelineIO.ensure_dir(REPORTS_DIR)
    run_contamination_stress_test(config, rng, output_dir=str(REPORTS_DIR), engine=ctx.engine)
    _echo("eval-contamination", ctx, contam=REPORTS_DIR / "contam_sensitivity.json", heatmap=REPORTS_DIR / "contam_heatmap.png")


@main.command("performance")
@click.option("--reset", is_flag=True, help="Reset performance monitoring data")
@click.option("--out-dir", default=DATA_ROOT / "performance", show_default=True, type=click.Path(path_type=Path))
@click.pass_obj
def performance_cmd(ctx: CLIContext, reset: bool, out_dir: Path) -> None:
    """Show performance monitoring statistics (``cli.py:408-438``).

    The reference reports its in-process ``PerformanceMonitor`` (empty in a fresh CLI process).  Here the command
    runs the smoke pipeline once with per-kernel CUDA-event timing and prints the same report shape --
    ``timing_statistics`` maps every CUDA kernel to ``calls / total_time / avg_time / min_time / max_time`` (seconds),
    ``peak_memory_mb`` is the device allocator's high-water mark."""
    from . import _native as nv

    if reset:
        nv.default_context(ctx.seed).profile_report()
        click.echo("Performance monitoring data reset.")
        return
    config = _load_pipeline_config(ctx.config_override, ctx.seed)
    dctx = nv.default_context(int(config.seed))
    dctx.profile_enable(True)
    _run_smoke_pipeline(config, ctx.seed, out_dir)
    prof = dctx.profile_report(detail=True)
    dctx.profile_enable(False)
    stats = {k: {"calls": n, "total_time": t * 1e-3, "avg_time": t * 1e-3 / max(n, 1), "min_time": mn * 1e-3, "max_time": mx * 1e-3}
             for k, (n, t, mn, mx) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    report = {"timing_statistics": stats, "peak_memory_mb": dctx.peak_memory_mb(), "total_functions_tracked": len(stats),
              "device_time_s": sum(v["total_time"] for v in stats.values()), "kernel_launches": int(sum(v["calls"] for v in stats.values()))}
    click.echo("Performance Report:")
    click.echo("=" * 50)
    if stats:
        click.echo("\nTiming Statistics:")
        for name, st in stats.items():
            click.echo(f"  {name}:")
            click.echo(f"    Calls: {st['calls']}")
            click.echo(f"    Total: {st['total_time']:.6f}s")
            click.echo(f"    Average: {st['avg_time']:.6f}s")
            click.echo(f"    Min: {st['min_time']:.6f}s")
            click.echo(f"    Max: {st['max_time']:.6f}s")
    click.echo(f"\nPeak Memory Usage: {report['peak_memory_mb']:.1f} MB")
    click.echo(f"Functions Tracked: {report['total_functions_tracked']}")
    click.echo(json.dumps(report, indent=2, default=str))


@main.command("eval-stratified")
@click.pass_obj
def eval_stratified_cmd(ctx: CLIContext) -> None:
    """Conduct stratified power and calibration analysis (cli.py:354-374)."""
    from .eval.stratified import run_stratified_analysis

    config = _load_pipeline_config(ctx.config_override, ctx.seed)
    rng = set_global_seed(ctx.seed, deterministic_ops=True)
    PipelineIO.ensure_dir(REPORTS_DIR)
    power, _ = run_stratified_analysis(config, rng, output_dir=str(REPORTS_DIR))
    validate_json(REPORTS_DIR / "power_by_stratum.json", "power.schema.json")
    validate_json(REPORTS_DIR / "calibration_by_bin.json", "calibration.schema.json")
    click.echo(json.dumps({"stage": "eval-stratified", "power_results": str(REPORTS_DIR / "power_by_stratum.json"),
                           "calibration_csv": str(REPORTS_DIR / "calibration_by_bin.csv"),
                           "contexts": power.get("contexts", []), "depths": power.get("depth_values", [])}, indent=2))


def cli() -> None:  # pragma: no cover
    main()


if __name__ == "__main__":  # pragma: no cover
    main()
