#!/usr/bin/env python
"""bench.py -- UMI reads/s through simulate -> collapse -> call on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host CPUs

Workload (config.workload): the reference's configs/default.yaml grid -- 5 AF x 4 depths x 1000
replicates = 20 000 cells, 4.25e8 UMI families, ~2.1e9 reads per GPU (C2 of SURVEY.md §8d).
One STEP = one pass of the hot path over that grid: Philox read generation -> radix sort ->
quality-weighted consensus -> per-cell counts -> error model -> fp64 tail p-values -> BH.
A "UMI read" is one member of one generated UMI family (before the min-family-size filter).

Multi-GPU: the replicate axis shards (rank r owns replicates [r*1000, (r+1)*1000)); there is no
data-path collective, per-GPU work is fixed => "scaling": "weak".  Philox streams are keyed by
the GLOBAL cell id, so a cell's result is independent of the GPU count.

Timing: CUDA events on the launching stream, W >= 3 warm-up steps, barrier + synchronize on
both sides, max over ranks.  Every chunk the kernels touch (several GB) is far larger than the
126 MB L2, so no L2 flush is needed between iterations (config.l2: "inputs larger than L2").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT / "precise-mrd-mini_b200"), str(ROOT)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "umi_reads_per_sec_simulate_collapse_call"
UNIT = "reads/s"
DTYPE = "u64 packed records (payload:32 | umi:32) through the sort, f64 consensus weights and tail tests"
B_READ_PIPELINE = 54.4  # algorithmic bytes per read of the whole path (48 + 32/f, f = 5)
HBM_FALLBACK_GBS = 6650.0


# ------------------------------------------------------------------------------- workload
def workload(name: str, rank: int, world: int):
    from precise_mrd.config import load_config
    from precise_mrd.grid import grid_cells

    cfg_dir = ROOT / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs"
    if name == "default":
        cfg = load_config(cfg_dir / "default.yaml")
        label = "configs/default.yaml C2 grid: 5 AF x 4 depths x 1000 replicates = 20000 cells per GPU"
    elif name == "smoke":
        cfg = load_config(cfg_dir / "smoke.yaml")
        label = "configs/smoke.yaml C1 grid: 3 AF x 2 depths x 10 replicates = 60 cells per GPU"
    elif name == "panel":
        # BASELINE.json configs[4]: 8 samples x 1e5 loci x 250 families per locus, Poisson(5) family
        # sizes, AF strata {0, 1e-4, 1e-3, 1e-2} by locus mod 4; one sample per GPU (the sample axis shards).  A cell is
        # a (sample, locus) pair: ~1250 reads, less than one 4096-record tile -- a secondary line, not the default.
        cfg = load_config(cfg_dir / "default.yaml")
        n_loci = int(os.environ.get("PMRD_PANEL_LOCI", 100000))
        locus = np.arange(n_loci, dtype=np.int64)
        cell_id = (np.int64(rank) << 20) + locus
        af = np.array([0.0, 1e-4, 1e-3, 1e-2])[locus % 4]
        depth = np.full(n_loci, 250, dtype=np.int64)
        return cfg, f"C5 panel: 1 sample x {n_loci} loci x 250 families per GPU (AF strata by locus mod 4)", cell_id, af, depth
    else:
        raise SystemExit(f"unknown workload {name}")
    sim = cfg.simulation
    reps = int(os.environ.get("PMRD_BENCH_REPS", sim.n_replicates))
    cell_id, af, depth, rep = grid_cells(sim.allele_fractions, sim.umi_depths, reps, rep_offset=rank * reps,
                                         rep_stride=world * reps)
    return cfg, label, cell_id, af, depth


def workload_config(name, label, cell_id, depth, world):
    """The workload-defining facts both arms can state without running anything (the reference arm prints the same dict)."""
    fused = name == "fused"
    return {"workload": label, "cells_per_gpu": int(len(cell_id)), "families_per_gpu_per_step": int(depth.sum()),
            "family_sizes": "clip(Poisson(5),1,1000)", "umi": "12-mer (24 bit)",
            "input_order": "family-major (no reads exist to be ordered)" if fused else "window-shuffled (Philox-keyed affine permutation)",
            "l2": "no read buffers; per-cell state lives in shared memory" if fused else "inputs larger than L2 (GBs of records per chunk vs 126 MB)",
            "parallelism": f"replicate axis sharded over {world} GPU(s)"}


def peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.samples, self._stop, self._t = device, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU port
def cpu_sample_cells(cell_id, af, depth, n):
    """A bounded, depth-balanced sample of the workload's cells (every k-th cell)."""
    step = max(1, len(cell_id) // n)
    idx = np.arange(0, len(cell_id), step)[:n]
    return cell_id[idx], af[idx], depth[idx]


def cpu_pipeline_once(cfg, cid, af, depth, threads):
    """The oracle port (oracle/pmrd_oracle.c): simulate -> collapse -> count -> call."""
    from oracle import cport

    t0 = time.perf_counter()
    r = cport.pipeline_cells(cid, af, depth, seed=int(cfg.seed), mean_family_size=5.0, family_model=1,
                             max_family_size=cfg.umi.max_family_size, min_fs=cfg.umi.min_family_size,
                             qthr=cfg.umi.quality_threshold, cthr=cfg.umi.consensus_threshold, n_threads=threads)
    neg = af <= 1e-4
    if not neg.any():
        neg = af == af.min()
    cport.call_cells(r["n_total"], r["n_variants"], neg, seed=int(cfg.seed), alpha=cfg.stats.alpha)
    return r["total_reads"], time.perf_counter() - t0


def host_threads() -> int:
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm
    sets the OpenMP team size explicitly instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(cfg, cell_id, af, depth, budget_s=12.0):
    threads = host_threads()
    n = 64
    reads, dt = cpu_pipeline_once(cfg, *cpu_sample_cells(cell_id, af, depth, n), threads)   # warm-up + rate estimate
    rate = reads / max(dt, 1e-6)
    per_cell = reads / n
    n = int(min(len(cell_id), max(64, budget_s * rate / per_cell)))
    reads, dt = cpu_pipeline_once(cfg, *cpu_sample_cells(cell_id, af, depth, n), threads)
    return {"value": reads / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} of {len(cell_id)} cells (every {max(1, len(cell_id) // n)}th), {reads} reads in {dt:.2f} s; "
                      "oracle/pmrd_oracle.c, OpenMP over cells"}


# ------------------------------------------------------------------------------- LoD grid
def lod_grid_wall(cfg, world=1, barrier=None):
    """Second half of BASELINE.json's metric: wall time of the eval-lod grid (C3 of SURVEY.md §8d:
    LoB blanks + 15 log-spaced AF in [1e-4, 1e-2] x depths {1k,5k,10k,50k,100k} x 25 replicates, 200-resample
    logistic bootstrap per depth) from LODAnalyzer entry to the finished tables, host wall clock."""
    import contextlib
    import copy
    import io
    import tempfile

    from precise_mrd.eval import LODAnalyzer

    out = {}
    depths = [1000, 5000, 10000, 50000, 100000]
    # world > 1: every rank enters; LODAnalyzer shards the cells over the ranks (LPT on the family count) and
    # all-gathers the per-cell counts once per grid (precise_mrd.distributed); wall = barrier to barrier
    engines = (("grid", 25), ("stages", 2)) if world == 1 else (("grid", 25),)
    from precise_mrd import _native as nv

    def analysis(engine, reps):
        c = copy.deepcopy(cfg)
        c.seed = 7
        an = LODAnalyzer(c, np.random.default_rng(7), engine=engine)
        with contextlib.redirect_stdout(io.StringIO()):
            an.estimate_lob(n_blank_runs=max(10, reps // 2))
            res = an.estimate_lod(depth_values=depths if engine == "grid" else depths[:3], n_replicates=reps)
            with tempfile.TemporaryDirectory() as td:
                an.generate_reports(td)
        return res

    for engine, reps in engines:
        if engine == "grid":
            analysis(engine, reps)                   # warm-up: allocator pool, module load
        if barrier:
            barrier()
        t0 = time.perf_counter()
        res = analysis(engine, reps)
        if barrier:
            barrier()
        wall = time.perf_counter() - t0
        n_dep = len(res["depth_results"])
        out[engine] = {"wall_s": wall, "n_gpus": world, "cells": 15 * n_dep * reps, "depths": n_dep, "replicates": reps,
                       "lod_af": {str(d): r["lod_af"] for d, r in res["depth_results"].items()}}
        if engine == "grid":
            # device-only time of the same analysis (per-kernel CUDA events on this rank), so that the host part is visible
            dctx = nv.default_context(7)
            dctx.profile_enable(True)
            analysis(engine, reps)
            prof = dctx.profile_report()
            dctx.profile_enable(False)
            out[engine]["device_ms"] = round(sum(v[1] for v in prof.values()), 3)
            out[engine]["device_launches"] = int(sum(v[0] for v in prof.values()))
            top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:4]
            out[engine]["device_top_ms"] = {k: round(v[1], 3) for k, v in top}
    out["note"] = ("grid = read-level fused device plan over the whole C3 grid (1875 cells + blanks); stages = the reference's own "
                   "per-cell simulate->collapse->fit_error_model->call_mrd loop with every stage on the GPU (parity engine; 90 cells "
                   "timed here, the reference's default 1125-cell grid scales linearly; the reference itself needs 3-7 s per cell)")
    return out


def contamination_grid_wall(cfg):
    """C4 (BASELINE.json configs[3]): the eval-contamination stress sweep -- hop {0,.001,.002,.005,.01} x AF {.001,.005,.01} x
    depth {1k,5k} x 20 reps + collisions (4 rates) + cross-sample (4 proportions) = 1 560 cells + their blanks -- through
    ContaminationSimulator(engine="grid"): ONE device plan with per-cell knobs.  Wall time and device time."""
    import contextlib
    import copy
    import io

    from precise_mrd import _native as nv
    from precise_mrd.sim import ContaminationSimulator

    c = copy.deepcopy(cfg)
    c.seed = 7

    def once():
        sim = ContaminationSimulator(c, np.random.default_rng(7), engine="grid")
        with contextlib.redirect_stdout(io.StringIO()):
            sim.simulate_contamination_effects()
        return sim

    once()                                   # warm-up
    t0 = time.perf_counter()
    sim = once()
    wall = time.perf_counter() - t0
    dctx = nv.default_context(7)
    dctx.profile_enable(True)
    l0 = dctx.launch_count
    once()
    launches = dctx.launch_count - l0
    prof = dctx.profile_report()
    dctx.profile_enable(False)
    t = sim.grid_table
    return {"wall_s": wall, "device_ms": round(sum(v[1] for v in prof.values()), 3), "device_launches": int(launches),
            "cells": int(len(t["p_value"])), "call_groups": int(len(t["group_error_rate"])), "reads": int(t["n_reads"].sum()),
            "device_top_ms": {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:4]},
            "note": "one plan: hopping conditions take the general path (12-byte records, full-key sort), the others the cell-major path"}


# ------------------------------------------------------------------------------- external reads
EXT_B_READ = 42.4   # SURVEY.md §8d without the simulate stage: sort 24 + consensus 15.2 + call 3.2 B / read


def external_reads(seed: int, n: int, n_loci: int):
    """BASELINE.json configs[4] as EXTERNAL input (one sample): n reads over n_loci loci in random
    order -- every read picks one of n / 5 molecules uniformly (family sizes ~ Poisson(5)), molecule m sits at locus
    m mod n_loci and carries a hashed 24-bit UMI; qualities uniform on 10..40, 3 % alt reads.  NumPy, so that the
    reference arm builds the very same arrays without a GPU."""
    rng = np.random.default_rng(seed)
    fam = rng.integers(0, max(1, n // 5), n)
    umi = (fam.astype(np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFF)
    keys = ((fam % n_loci).astype(np.uint64) << np.uint64(32)) | umi
    q = rng.integers(10, 41, n).astype(np.uint32)
    alt = (rng.random(n) < 0.03).astype(np.uint32)
    payload = alt | (q << np.uint32(2)) | (alt << np.uint32(31))
    return keys, payload


def external_config(n, n_loci, world):
    return {"workload": f"C5 panel as EXTERNAL reads: {n} reads per GPU over {n_loci} loci in random order (one sample per GPU) "
                        "-> pmrd_collapse_reads (UMIProcessor.process_reads + get_consensus_counts)",
            "reads_per_gpu": int(n), "loci": int(n_loci), "family_sizes": "~Poisson(5) (uniform molecule choice)", "umi": "12-mer (24 bit)",
            "input_order": "random (i.i.d. molecule choice per read)", "l2": "inputs larger than L2 (12 B x reads vs 126 MB)",
            "parallelism": f"sample axis sharded over {world} GPU(s)"}


def external_cpu(keys, payload, n_loci, cfg, threads):
    from oracle import cport

    t0 = time.perf_counter()
    r = cport.collapse_groups(keys, payload, n_loci, cfg.umi.min_family_size, cfg.umi.max_family_size, cfg.umi.quality_threshold,
                              cfg.umi.consensus_threshold, n_threads=threads)
    return r, time.perf_counter() - t0


def run_external(args, rank, world, local_rank, emit):
    """--workload external: the general collapse path on reads that arrive from outside, in arbitrary order."""
    import torch
    import torch.distributed as dist

    from precise_mrd import _native as nv
    from precise_mrd.config import load_config

    cfg = load_config(ROOT / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs" / "default.yaml")
    n = int(os.environ.get("PMRD_EXT_READS", 125_000_000))
    n_loci = int(os.environ.get("PMRD_EXT_LOCI", 100_000))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = nv.Context(device=local_rank, seed=int(cfg.seed))
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    P = nv.UmiParams(min_family_size=cfg.umi.min_family_size, max_family_size=cfg.umi.max_family_size,
                     quality_threshold=cfg.umi.quality_threshold, consensus_threshold=cfg.umi.consensus_threshold,
                     mode=nv.MODE_WEIGHTED, cluster=nv.CLUSTER_EXACT, max_distance=0, umi_len=12)
    keys, payload = external_reads(int(cfg.seed) + rank, n, n_loci)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: pmrd_collapse_reads_dev on buffers already in HBM ---------------------------
    dev = torch.device("cuda", local_rank)
    k0 = torch.from_numpy(keys.view(np.int64)).to(dev)
    p0 = torch.from_numpy(payload.view(np.int32)).to(dev)
    dk, dp, dkt, dpt = torch.empty_like(k0), torch.empty_like(p0), torch.empty_like(k0), torch.empty_like(p0)
    fcap = n
    fam_t = {"key": torch.empty(fcap, dtype=torch.int64, device=dev)}
    for name in ("size", "n_alt", "consensus", "qsum", "n_best", "flags"):
        fam_t[name] = torch.empty(fcap, dtype=torch.int32, device=dev)
    gcap = n_loci + 16
    grp_t = {"group": torch.empty(gcap, dtype=torch.int32, device=dev), "family_count": torch.empty(gcap, dtype=torch.int32, device=dev),
             "consensus_count": torch.empty(gcap * 4, dtype=torch.int32, device=dev), "alt_count": torch.empty(gcap, dtype=torch.int64, device=dev),
             "total_count": torch.empty(gcap, dtype=torch.int64, device=dev)}
    fam_p = {k: v.data_ptr() for k, v in fam_t.items()}
    grp_p = {k: v.data_ptr() for k, v in grp_t.items()}

    def step_dev(with_rows=False):
        """The per-locus table (UMIProcessor.process_reads + get_consensus_counts); with_rows: the family table as well."""
        dk.copy_(k0); dp.copy_(p0)                                   # restore the inputs (the call clobbers them): untimed
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        nf, ng = ctx.collapse_reads_dev(dk.data_ptr(), dp.data_ptr(), dkt.data_ptr(), dpt.data_ptr(), n, P, fam_p if with_rows else None, fcap, grp_p, gcap)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), nf, ng

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    launches0 = ctx.launch_count
    ms = 0.0
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.steps):
            dt, nf, ng = step_dev()
            ms += dt
        barrier()
    launches = ctx.launch_count - launches0
    ctx.profile_enable(True)
    for _ in range(args.steps):
        step_dev()
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    d_fc = grp_t["family_count"][:ng].cpu().numpy().astype(np.uint32)
    d_cc = grp_t["consensus_count"][:ng * 4].cpu().numpy().astype(np.uint32).reshape(ng, 4)
    rows_ms = min(step_dev(True)[0] for _ in range(3))               # the same call with the family table materialised as well
    assert (grp_t["family_count"][:ng].cpu().numpy().astype(np.uint32) == d_fc).all()

    # ---- end-to-end arm: host (pageable NumPy) arrays in, the per-locus table out, every step --------------
    ctx.collapse_reads(keys, payload, P, want_families=False)        # warm the allocator pool and the pinned ring
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fam_h, grp_h = ctx.collapse_reads(keys, payload, P, want_families=False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = int(keys.nbytes + payload.nbytes)
    d2h = int(sum(v.nbytes for v in grp_h.values()))
    assert (grp_h["family_count"] == d_fc).all() and (grp_h["consensus_count"] == d_cc).all()
    # with the family table as well (7 columns, 32 B per family)
    t1 = time.perf_counter()
    fam_h, _ = ctx.collapse_reads(keys, payload, P, want_families=True)
    torch.cuda.synchronize()
    e2e_fam_s = time.perf_counter() - t1
    assert len(fam_h["key"]) == nf

    # ---- N > 1: the SAME reads taken as stripes of ONE sample -> loci shared by all ranks -> the exchange of SURVEY.md §8e
    # (owner cut on the device, all-to-all over NCCL, owner-side collapse; precise_mrd.distributed.collapse_reads_exchanged)
    exchange = None
    if world in (2, 4, 8):                                            # (owner = locus mod N needs a power of two; PeerExchange serves one node)
        exchange = exchange_leg(ctx, args, k0, p0, P, n, n_loci, rank, world, local_rank, barrier)

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = [float(x) for x in t.tolist()]
    value = float(n) * world * args.steps / (ms_max * 1e-3)
    e2e_value = float(n) * world * args.steps / (e2e_ms_max * 1e-3)
    if rank == 0:
        peak, peak_src = peak_hbm()
        total_prof_ms = sum(v[1] for v in prof.values()) or 1.0
        kms = {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:10]}
        top_name, top = max(prof.items(), key=lambda kv: kv[1][1])
        # bytes one launch of each kernel moves per read (its share of the data path)
        per_launch = {"rs_downsweep_kernel<true>": 24.0, "rs_upsweep_kernel": 8.0, "group_rows_kernel": 12.0, "grp_bounds_kernel": 8.0,
                      "family_kernel": 12.0 + 32.0 / 5.0, "group_rows_compact_kernel": 64.0 / 5.0}
        b = next((v for k, v in per_launch.items() if top_name.startswith(k.split("<")[0])), None)
        avg_ms = top[1] / max(top[0], 1)
        ach = b * n / (avg_ms * 1e-3) / 1e9 if b else None
        roof = {"bound": "hbm", "kernel": top_name, "launches": int(top[0]), "avg_launch_ms": avg_ms, "share_of_step": top[1] / total_prof_ms,
                "algorithmic_bytes_per_launch": b * n if b else None, "bytes_per_record": b, "achieved": ach, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak if ach else None, "traffic": None,
                "note": "partition pass over the group bits: R 12 + W 12 B per read and launch"}
        line = {
            "metric": METRIC.replace("simulate_collapse_call", "collapse_external"), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 key (group:32 | umi:32) + u32 payload, f64 consensus weights", "data": "synthetic",
            "config": external_config(n, n_loci, world),
            "layout": {"families": int(nf), "loci_with_reads": int(ng), "path": "group-major (stable partition by locus -> one block sorts and votes a locus -> its row of the per-locus table)",
                       "with_family_table_ms_per_step": rows_ms, "with_family_table_value": float(n) * world / (rows_ms * 1e-3)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "Context.collapse_reads(keys, payload, want_families=False): pageable NumPy arrays through the pinned staging "
                           "ring (csrc/staging.cu), per-locus table back",
                    "with_family_table": {"value": float(n) / e2e_fam_s, "d2h_bytes": int(sum(v.nbytes for v in fam_h.values()))}},
            "gpu_launches": int(launches), "roofline": roof,
            "pipeline_roofline": {"algorithmic_bytes_per_read": EXT_B_READ, "achieved": value / world * EXT_B_READ / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": value / world * EXT_B_READ / 1e9 / peak},
            "kernels_ms_per_step": kms, "clocks": clocks.summary(), "reference_class": "cpu-port",
        }
        if exchange is not None:
            line["exchange"] = exchange
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            r, dt = external_cpu(keys, payload, n_loci, cfg, threads)
            assert (r["family_count"][:ng] == d_fc).all() and (r["consensus_count"][:ng] == d_cc).all(), "device table != oracle table"
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"all {n} reads once in {dt:.2f} s; oracle/pmrd_oracle.c oracle_collapse_groups (bucket by locus, hash "
                                              "the UMIs, weighted vote: umi.py:205-266), OpenMP over loci; its table equals the device's"}
            line["result_check"] = {"families": int(nf), "equals_oracle": True}
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def exchange_leg(ctx, args, k0, p0, P, n, n_loci, rank, world, local_rank, barrier):
    """Reads of one sample striped over the ranks (each rank's 1.25e8 reads cover ALL loci): every step cuts them by owner
    rank, moves 1 - 1/N of them over NVLink and collapses at the owner.  Device time per step = max over ranks (CUDA events
    around the whole call); first a small instance is checked against the single-GPU table of the concatenated reads."""
    import torch
    import torch.distributed as dist

    from precise_mrd.distributed import collapse_reads_exchanged

    dev = torch.device("cuda", local_rank)
    # weak scaling of ONE sample: N x 1.25e8 reads over N x 1e5 loci, every rank holding a stripe that covers all of them
    # (~1250 reads per locus in total, ~1250 / N of them on each rank), so an owner ends up with the single-GPU problem
    n_loci = n_loci * world
    kx, px = external_reads(77_000 + rank, n, n_loci)
    k0, p0 = torch.from_numpy(kx.view(np.int64)).to(dev), torch.from_numpy(px.view(np.int32)).to(dev)
    del kx, px
    # (1) equality on a small instance: rank 0 collapses the concatenation on its own
    ns, ids = 200_000, 4096
    ks, ps = external_reads(1000 + rank, ns, ids)
    tk, tp = torch.from_numpy(ks.view(np.int64)).to(dev), torch.from_numpy(ps.view(np.int32)).to(dev)
    allk = [torch.empty_like(tk) for _ in range(world)]
    allp = [torch.empty_like(tp) for _ in range(world)]
    dist.all_gather(allk, tk)
    dist.all_gather(allp, tp)
    got = collapse_reads_exchanged(ctx, tk.clone(), tp.clone(), P, ids)
    ok = torch.ones(1, dtype=torch.int64, device=dev)
    if rank == 0:
        info, one = ctx.collapse_reads(torch.cat(allk).cpu().numpy().view(np.uint64), torch.cat(allp).cpu().numpy().view(np.uint32), P, want_families=False)
        same = (got["n_families"] == info["n_families"] and (got["group"] == one["group"]).all() and (got["family_count"] == one["family_count"]).all()
                and (got["consensus_count"] == one["consensus_count"]).all())
        ok[0] = 1 if same else 0
    dist.broadcast(ok, 0)
    # (2) the full-size exchange, timed in both forms: partitioned copy + NCCL all-to-all, and ONE cut-and-store kernel over
    # peer memory (the owners' receive arrays mapped into every rank)
    from precise_mrd.distributed import PeerExchange

    peer = PeerExchange(ctx, int(n * 1.05) + 4096)
    got_p = collapse_reads_exchanged(ctx, tk.clone(), tp.clone(), P, ids, peer=peer)
    okp = torch.ones(1, dtype=torch.int64, device=dev)
    if not (got_p["n_families"] == got["n_families"] and (got_p["group"] == got["group"]).all() and (got_p["family_count"] == got["family_count"]).all()
            and (got_p["consensus_count"] == got["consensus_count"]).all()):
        okp[0] = 0
    dist.all_reduce(okp, op=dist.ReduceOp.MIN)
    forms = {}
    for form, pe in (("nccl", None), ("peer", peer)):
        parts = {"partition_ms": 0.0, "exchange_ms": 0.0, "collapse_ms": 0.0}
        total_ms, fams, sent = 0.0, 0, 0
        launches0 = ctx.launch_count
        for it in range(max(args.warmup, 3) + args.steps):
            dk, dp = k0.clone(), p0.clone()                          # the call clobbers its inputs: restored untimed
            barrier()
            if it == max(args.warmup, 3):
                launches0 = ctx.launch_count
            tm = {}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tab = collapse_reads_exchanged(ctx, dk, dp, P, n_loci, gather=False, timings=tm, peer=pe)
            e1.record()
            torch.cuda.synchronize()
            if it >= max(args.warmup, 3):
                total_ms += e0.elapsed_time(e1)
                for k_ in parts:
                    parts[k_] += tm[k_]
                fams, sent = tab["n_families"], tm["sent_reads"]
            del dk, dp
        t = torch.tensor([total_ms] + [parts[k_] for k_ in ("partition_ms", "exchange_ms", "collapse_ms")], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        f = torch.tensor([fams, sent], dtype=torch.int64, device=dev)
        dist.all_reduce(f)
        tot, pm, xm, cm = [float(x) / args.steps for x in t.tolist()]
        names = ("owner_partition", "all_to_all", "owner_collapse") if form == "nccl" else ("tile_offsets_and_counts", "cut_and_store_over_nvlink", "owner_collapse")
        forms[form] = {"value": float(n) * world / (tot * 1e-3), "unit": UNIT, "ms_per_step": tot, "max_over_ranks_ms": dict(zip(names, (pm, xm, cm))),
                       "reads_sent_per_step": int(f[1].item()), "bytes_sent_per_step": int(f[1].item()) * 12,
                       "exchange_GBps_per_gpu": (int(f[1].item()) / world * 12 / 1e9) / (xm * 1e-3) if xm > 0 else None,
                       "families": int(f[0].item()), "gpu_launches_per_step": (ctx.launch_count - launches0) // max(args.steps, 1)}
    peer.close()
    return {"what": "ONE sample of N x 1.25e8 reads over N x 1e5 loci, striped over the ranks in arrival order (every rank holds reads of every "
                    "locus) -> every locus has an owner rank (locus mod N) -> per-locus table of all reads, rows held by their owners.  nccl: pmrd_owner_partition_dev -> all_to_all_single -> pmrd_owner_rekey_dev + "
                    "pmrd_collapse_reads_dev; peer: pmrd_owner_tiles_dev -> pmrd_owner_scatter_peer_dev (the cut stores into the owners' "
                    "IPC-mapped receive arrays over NVLink: no staging copy, no all-to-all) -> the same owner-side collapse",
            "scaling": "weak", "value": forms["peer"]["value"], "unit": UNIT, "ms_per_step": forms["peer"]["ms_per_step"], "loci": int(n_loci), "forms": forms,
            "exchanged_equals_single": bool(ok.item()), "peer_form_equals_nccl_form": bool(okp.item()),
            "equality_instance": f"{ns} reads per rank over {ids} loci vs rank 0's single-GPU table of the concatenation"}


# ------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world, emit):
    if rank != 0:
        return 0
    if args.workload == "external":
        from precise_mrd.config import load_config

        cfg = load_config(ROOT / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs" / "default.yaml")
        n_full = int(os.environ.get("PMRD_EXT_READS", 125_000_000))
        n_loci = int(os.environ.get("PMRD_EXT_LOCI", 100_000))
        n = int(os.environ.get("PMRD_REF_READS", min(n_full, 25_000_000)))         # bounded sample: a fifth of the reads, all loci
        keys, payload = external_reads(int(cfg.seed), n, n_loci)
        threads = host_threads()
        for _ in range(args.warmup):
            external_cpu(keys, payload, n_loci, cfg, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            external_cpu(keys, payload, n_loci, cfg, threads)
        dt = time.perf_counter() - t0
        value = n * args.steps / dt
        emit({"impl": "reference", "metric": METRIC.replace("simulate_collapse_call", "collapse_external"), "value": value, "unit": UNIT,
              "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "u64 key (group:32 | umi:32) + u32 payload, f64 consensus weights",
              "data": "synthetic", "config": external_config(n_full, n_loci, args.gpus),
              "sample": f"{n} of {n_full} reads per step over all {n_loci} loci; value is a rate",
              "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{n} reads per step; oracle/pmrd_oracle.c oracle_collapse_groups (umi.py:205-266), OpenMP over loci"},
              "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
              "reference_class": "cpu-port"})
        return 0
    fused = args.workload == "fused"
    cfg, label, cell_id, af, depth = workload("default" if fused else args.workload, 0, 1)
    if fused:
        label = label.replace("C2 grid", "C2 grid, FUSED engine (reads voted where they are drawn, never written)")
    threads = host_threads()
    n = int(os.environ.get("PMRD_REF_CELLS", 256))
    cid, a, d = cpu_sample_cells(cell_id, af, depth, n)
    for _ in range(args.warmup):
        cpu_pipeline_once(cfg, cid, a, d, threads)
    t0 = time.perf_counter()
    reads = 0
    for _ in range(args.steps):
        r, _dt = cpu_pipeline_once(cfg, cid, a, d, threads)
        reads += r
    dt = time.perf_counter() - t0
    value = reads / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args.workload, label, cell_id, depth, args.gpus),
        "sample": f"{len(cid)} cells per step (every {max(1, len(cell_id) // n)}th cell of the grid: every AF x depth stratum); value is a rate",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{len(cid)} cells, {reads // args.steps} reads per step; oracle/pmrd_oracle.c (C + OpenMP "
                                   "restatement of io.py / umi.py / call.py; the Python reference cannot travel to the GPU box)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "reference_class": "cpu-port",
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------- main arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PMRD_BENCH_WORKLOAD", "default"))
    ap.add_argument("--chunk-log2", type=int, default=int(os.environ.get("PMRD_CHUNK_LOG2", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lod", action="store_true", help="skip the LoD-grid wall-time leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, rank, world, emit)

    import torch
    import torch.distributed as dist

    from precise_mrd import _native as nv
    from precise_mrd import grid

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    if args.workload == "external":
        return run_external(args, rank, world, local_rank, emit)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    fused = args.workload == "fused"
    cfg, label, cell_id, af, depth = workload("default" if fused else args.workload, rank, world)
    if fused:
        label = label.replace("C2 grid", "C2 grid, FUSED engine (reads voted where they are drawn, never written)")
    ctx = nv.Context(device=local_rank, seed=int(cfg.seed))
    # a real (non-default) stream shared by torch's events and the library's launches: the CUDA events that time the
    # step are recorded on the very stream the kernels run on (the default stream's handle is 0, which pmrd_set_stream
    # reads as "keep the context's own stream")
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    assert stream.cuda_stream != 0
    sim = grid.read_sim_params(cfg, mean_family_size=5.0, family_model=1, chunk_log2=args.chunk_log2, fused=fused)
    up = grid.umi_params(cfg)
    tt = nv.TEST_TYPES[cfg.stats.test_type]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(plan, steps):
        """K steps bracketed by CUDA events on the launching stream, barrier + synchronize on both sides."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            plan.run()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    # ---- device-resident arm: the plan owns all buffers; run() only enqueues kernels ----------
    neg_af_max = 0.0 if args.workload == "panel" else 1e-4
    plan = nv.Plan(ctx, cell_id, af, depth, sim, up, neg_af_max=neg_af_max, test_type=tt, alpha=float(cfg.stats.alpha))
    n_reads = plan.n_reads
    for _ in range(max(args.warmup, 3)):
        plan.run()
    launches0 = ctx.launch_count
    with ClockSampler(local_rank) as clocks:
        ms = timed_steps(plan, args.steps)
    launches = ctx.launch_count - launches0
    # per-kernel breakdown: a SEPARATE pass of the same K steps with per-launch CUDA events (pmrd_profile_enable) --
    # the event pairs cost device time, so they stay out of the region `value` is timed over
    ctx.profile_enable(True)
    for _ in range(args.steps):
        plan.run()
    barrier()
    prof = ctx.profile_report()
    ctx.profile_enable(False)
    res = plan.results()
    chunk_reads = [plan.chunk_reads(i) for i in range(plan.n_chunks)]
    n_chunks = plan.n_chunks
    pinfo = plan.info()
    plan.close()

    # ---- end-to-end arm: public API, host buffers in, host table out, every step ------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_cid, h_af, h_depth = pin(cell_id), pin(af), pin(depth)
    out = grid.run_cells(cfg, h_cid, h_af, h_depth, ctx=ctx, sim=sim, neg_af_max=neg_af_max)      # warm the allocator pool
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = grid.run_cells(cfg, h_cid, h_af, h_depth, ctx=ctx, sim=sim, neg_af_max=neg_af_max)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = int(h_cid.nbytes + h_af.nbytes + h_depth.nbytes)
    d2h = int(sum(v.nbytes for k, v in out.items() if isinstance(v, np.ndarray) and k != "variant_call"))
    assert (out["n_total"] == res["n_total"]).all() and (out["p_adjusted"] == res["p_adjusted"]).all()

    # ---- strong scaling (N > 1): the FIXED single-GPU grid split over the ranks ---------------------------------
    # `value` above is weak scaling (every rank runs a whole grid).  north_star also asks how one fixed grid scales:
    # the N = 1 workload's cells (same global cell ids) are dealt to the ranks in blocks of replicates, every rank
    # runs its share, time = max over ranks.  No data-path collective (per-cell counts are what a rank keeps).
    strong = None
    if world > 1 and args.workload in ("default", "fused"):
        _, _, s_cid, s_af, s_depth = workload("default", 0, 1)
        mine = np.arange(len(s_cid))[rank::world]                      # every (AF, depth) stratum on every rank
        splan = nv.Plan(ctx, s_cid[mine], s_af[mine], s_depth[mine], sim, up, neg_af_max=neg_af_max, test_type=tt, alpha=float(cfg.stats.alpha))
        for _ in range(3):
            splan.run()
        s_ms = timed_steps(splan, args.steps)
        s_reads = splan.n_reads
        splan.close()
        ts = torch.tensor([s_ms], dtype=torch.float64, device="cuda")
        rs = torch.tensor([float(s_reads)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        dist.all_reduce(rs, op=dist.ReduceOp.SUM)
        strong = {"scaling": "strong", "workload": f"the fixed N=1 grid ({len(s_cid)} cells, {int(rs.item())} reads) split over {world} GPUs, cell i -> rank i mod N",
                  "ms_per_step": float(ts.item()) / args.steps, "value": float(rs.item()) * args.steps / (float(ts.item()) * 1e-3), "unit": UNIT,
                  "note": "per-rank share incl. its own call stage; compare ms_per_step with the N=1 line's"}

    # ---- sharded == single (N > 1): driver-visible proof that the multi-GPU table is the single-GPU one ----------
    sharded_equals_single = None
    if world > 1:
        from precise_mrd import distributed
        from precise_mrd.eval import LODAnalyzer

        an = LODAnalyzer(cfg, np.random.default_rng(7), engine="grid")
        g_cid, g_af, g_dp = an._grid_cells(2, np.logspace(-4, -2, 15), [1000, 5000, 10000, 50000, 100000], 25, 25)
        ok = True
        for eng_fused in (True, False):                      # the engine LODAnalyzer(engine="grid") uses, and the sort-based one
            g_sim = grid.read_sim_params(cfg, fused=eng_fused)
            sh = distributed.run_cells_sharded(cfg, g_cid, g_af, g_dp, neg_af_max=0.0, alternative=nv.ALT_GREATER, ctx=ctx, sim=g_sim)
            if rank == 0:
                one = grid.run_cells(cfg, g_cid, g_af, g_dp, ctx=ctx, sim=g_sim, neg_af_max=0.0, alternative=nv.ALT_GREATER)
                for k in ("n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant"):
                    ok = ok and bool((np.asarray(sh[k]) == np.asarray(one[k])).all())
                ok = ok and sh["mean_error_rate"] == one["mean_error_rate"]
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        sharded_equals_single = {"equal": bool(flag.item()), "cells": int(len(g_cid)),
                                 "what": "C3 eval-lod grid (1875 cells + blanks), fused and materialised engines: run_cells_sharded over all ranks (LPT shards, one NCCL "
                                         "all-gather of per-cell counts, call stage on every rank) vs grid.run_cells on rank 0 alone; "
                                         "counts, p, q, significant, error rate compared bit for bit"}
        if not flag.item():
            raise SystemExit("bench.py: the sharded table differs from the single-GPU one")

    # ---- reduce over ranks: max time, sum of reads ---------------------------------------------
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    r = torch.tensor([float(n_reads)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = [float(x) for x in t.tolist()]
    total_reads = float(r.item())
    value = total_reads * args.steps / (ms_max * 1e-3)
    e2e_value = total_reads * args.steps / (e2e_ms_max * 1e-3)

    lod = None
    if not args.no_lod:
        lod = lod_grid_wall(cfg, world, barrier if world > 1 else None)      # every rank takes part when world > 1

    if rank == 0:
        peak, peak_src = peak_hbm()
        total_prof_ms = sum(v[1] for v in prof.values()) or 1.0
        kernels_ms = {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]}
        config = workload_config(args.workload, label, cell_id, depth, world)
        layout = {"reads_per_gpu_per_step": int(n_reads), "chunks": n_chunks}
        if fused:
            # No read is written: the bounding resource is the INT32 issue rate of Philox, so the
            # roofline is issue-slot utilisation.  Warp instructions per read come from the ncu capture of the same
            # kernel (profiles/fused_inst.json: smsp__inst_executed.sum / reads); peak = SMs x 4 schedulers x SM clock.
            inst = None
            fj = ROOT / "profiles" / "fused_inst.json"
            if fj.exists():
                try:
                    inst = float(json.loads(fj.read_text())["warp_inst_per_read"])
                except Exception:
                    inst = None
            cs = clocks.summary()
            sm_clock = (cs["sm_mhz"] or 1965.0) * 1e6
            k_ms = sum(v[1] for k, v in prof.items() if k.startswith("fused_cell_kernel")) / args.steps
            issue_peak = 148 * 4 * sm_clock / 1e9                                    # G warp-instructions / s
            ach = inst * n_reads / (k_ms * 1e-3) / 1e9 if (inst and k_ms) else None
            roof = {"bound": "issue", "kernel": "fused_cell_kernel", "launches": int(sum(v[0] for k, v in prof.items() if k.startswith("fused_cell_kernel"))),
                    "avg_launch_ms": k_ms / max(n_chunks, 1), "share_of_step": k_ms * args.steps / total_prof_ms,
                    "achieved": ach, "peak": issue_peak, "peak_source": "148 SMs x 4 warp schedulers x SM clock under load",
                    "unit": "Gwarp-inst/s", "frac": (ach / issue_peak) if ach else None,
                    "warp_inst_per_read": inst, "traffic": 12.0 * pinfo["n_slots"] / max(n_chunks, 1),
                    "note": "traffic = the candidate-duplicate list (12 B per candidate), the only per-read-scale HBM data; "
                            "HBM fraction is not this kernel's roofline"}
            layout.update({"candidate_duplicates_per_step": pinfo["n_slots"]})
            pipeline = {"algorithmic_bytes_per_read": 6.4, "achieved": value / world * 6.4 / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": value / world * 6.4 / 1e9 / peak, "note": "SURVEY.md §8d: family record only, 32 B / f"}
        else:
            # Per-kernel device time comes from CUDA events recorded in the profiled pass.  The radix scatter
            # pass exists in two variants -- staged through shared memory for the first pass (scattered generator
            # output), direct for the later passes -- which are ONE logical kernel: the dominant one.  Its
            # achieved bandwidth = algorithmic bytes of all its launches in a step / their summed device time.
            rec = pinfo["record_bytes"]
            slots = pinfo["n_slots"]                                       # record slots per step (reads + tile padding)
            passes = pinfo["sort_passes"]
            n_staged = sum(v[0] for k, v in prof.items() if k.startswith("rs_downsweep_staged_kernel")) / max(args.steps * n_chunks, 1)
            # the generator counts the first digit as a by-product (PMRD_GEN_HIST, default on): one counting pass fewer
            up_passes = passes - (0 if os.environ.get("PMRD_GEN_HIST", "1") == "0" else 1)
            groups = {
                "rs_downsweep": (["rs_downsweep_direct_kernel", "rs_downsweep_staged_kernel", "rs_downsweep_kernel<false>"], 2.0 * rec * passes),
                "rs_upsweep_kernel": (["rs_upsweep_kernel"], 8.0 * up_passes),
                "reads_gen_kernel": (["reads_gen_kernel"], float(rec)),
                "cell_consensus_kernel": (["cell_consensus_peel_kernel", "cell_consensus_kernel<true>", "cell_consensus_kernel<false>"], float(rec)),
                "small_cell_kernel": (["small_cell_kernel"], float(rec)),
            }
            table = []
            for gname, (members, bytes_per_slot) in groups.items():
                g_ms = sum(prof[m][1] for m in members if m in prof)
                g_n = sum(prof[m][0] for m in members if m in prof)
                if g_n == 0:
                    continue
                ach = bytes_per_slot * slots * args.steps / (g_ms * 1e-3) / 1e9
                table.append({"kernel": gname, "variants": [m for m in members if m in prof], "launches": g_n, "ms_per_step": g_ms / args.steps,
                              "share_of_step": g_ms / total_prof_ms, "algorithmic_bytes_per_record": bytes_per_slot, "achieved": ach,
                              "frac": ach / peak})
            table.sort(key=lambda r: -r["ms_per_step"])
            top = table[0] if table else {"kernel": "none", "variants": [], "launches": 0, "ms_per_step": 0.0, "share_of_step": 0.0,
                                          "algorithmic_bytes_per_record": 0.0, "achieved": None, "frac": None}
            kname = top["variants"][0] if top["variants"] else "none"
            roof = {"bound": "hbm", "kernel": top["kernel"], "variants": top["variants"], "launches": top["launches"],
                    "avg_launch_ms": top["ms_per_step"] * args.steps / max(top["launches"], 1), "share_of_step": top["share_of_step"],
                    "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": None, "achieved": top["achieved"], "frac": top["frac"],
                    "algorithmic_bytes_per_launch": (top["algorithmic_bytes_per_record"] * slots * args.steps / max(top["launches"], 1))
                    if top["launches"] else None,
                    "bytes_per_record": top["algorithmic_bytes_per_record"] / (passes if top["kernel"].startswith("rs_") else 1),
                    "records_per_launch": slots / max(n_chunks, 1), "staged_passes_per_chunk": n_staged,
                    "note": "reads_gen_kernel is INT32-issue bound (one Philox2x32-10 block per read), "
                            "its HBM fraction is listed for completeness only",
                    "all_hot_kernels": table}
            reads_per_launch = slots / max(n_chunks, 1)
            tj = ROOT / "profiles" / "traffic.json"
            if tj.exists():
                try:
                    per_read = json.loads(tj.read_text()).get(kname, {}).get("dram_bytes_per_read")   # per launch of one variant
                    if per_read:
                        roof["traffic"] = per_read * reads_per_launch
                except Exception:
                    pass
            layout.update({"record_bytes": pinfo["record_bytes"], "radix_passes": pinfo["sort_passes"], "counting_passes": up_passes,
                           "smallest_chunk_gb": round(min(chunk_reads) * pinfo["record_bytes"] / 1e9, 2)})
            pipeline = {"algorithmic_bytes_per_read": B_READ_PIPELINE, "achieved": value / world * B_READ_PIPELINE / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": value / world * B_READ_PIPELINE / 1e9 / peak}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": config, "layout": layout,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "precise_mrd.grid.run_cells: per step plan create (host cell tables + H2D + the sizing pass over the "
                           "families -- the first run() after a build reuses that pass's family sizes instead of redrawing them, "
                           "which saves reads_size_kernel once per e2e step) + run + D2H of the per-cell table"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "pipeline_roofline": pipeline,
            "kernels_ms_per_step": kernels_ms,
            "clocks": clocks.summary(),
            "result_check": {"significant_cells": int(res["significant"].sum()), "mean_error_rate": res["mean_error_rate"],
                             "families_formed": int(res["totals"][1])},
            "reference_class": "cpu-port",
            "tolerances": "counts / consensus / BH decisions bit-exact; p, q <= 1e-12 relative (p >= 1e-50, binomial n <= 2e4), "
                          "2e-11 / 1e-11 in the deeper tails where SciPy itself is the inexact party (mpmath evidence in tests/test_gpu_parity.py)",
        }
        if strong is not None:
            line["strong_scaling"] = strong
        if sharded_equals_single is not None:
            line["sharded_equals_single"] = sharded_equals_single["equal"]
            line["sharded_check"] = sharded_equals_single
        if lod is not None:
            line["lod_grid"] = lod
        if world == 1 and not args.no_lod:
            try:
                line["contamination_grid"] = contamination_grid_wall(cfg)
            except Exception as exc:
                line["contamination_grid"] = {"failed": str(exc)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(cfg, cell_id, af, depth)
            except Exception as exc:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
