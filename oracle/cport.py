"""ctypes door to oracle/pmrd_oracle.c (TEST INFRASTRUCTURE ONLY: tests/, smoke(), and the
cpu_baseline / --impl reference legs of bench.py)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libpmrd_oracle.so"
_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
        _lib = C.CDLL(str(LIB))
        _lib.oracle_collapse_weighted.restype = C.c_int64
        _lib.oracle_pipeline_cells.restype = C.c_int64
        _lib.oracle_collapse_groups.restype = C.c_int64
        _lib.oracle_poisson_two_sided.restype = C.c_double
        _lib.oracle_poisson_two_sided.argtypes = [C.c_int64, C.c_double]
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def collapse_weighted(keys, payload, min_fs, max_fs, qthr, cthr):
    lib = load()
    keys = np.ascontiguousarray(keys, np.uint64)
    payload = np.ascontiguousarray(payload, np.uint32)
    n = len(keys)
    out = {"key": np.zeros(n, np.uint64)}
    for k in ("size", "n_alt", "consensus", "qsum", "n_best", "flags"):
        out[k] = np.zeros(n, np.uint32)
    nf = lib.oracle_collapse_weighted(_p(keys), _p(payload), C.c_int64(n), C.c_int(min_fs), C.c_int(max_fs), C.c_int(qthr),
                                      C.c_double(cthr), *[_p(out[k]) for k in ("key", "size", "n_alt", "consensus", "qsum", "n_best", "flags")])
    return {k: v[:nf] for k, v in out.items()}


def collapse_groups(keys, payload, n_ids, min_fs, max_fs, qthr, cthr, n_threads=0):
    """Group table of external panel reads (dense group ids < n_ids): kept families, consensus counts per allele,
    families formed -- umi.py:205-266 per group, OpenMP over groups."""
    lib = load()
    keys = np.ascontiguousarray(keys, np.uint64)
    payload = np.ascontiguousarray(payload, np.uint32)
    fc, cc, ff = np.zeros(n_ids, np.uint32), np.zeros((n_ids, 4), np.uint32), np.zeros(n_ids, np.uint32)
    total = lib.oracle_collapse_groups(_p(keys), _p(payload), C.c_int64(len(keys)), C.c_int64(n_ids), C.c_int(min_fs), C.c_int(max_fs),
                                       C.c_int(qthr), C.c_double(cthr), _p(fc), _p(cc), _p(ff), C.c_int(n_threads))
    return {"family_count": fc, "consensus_count": cc, "families_formed": ff, "total_families": int(total)}


def pipeline_cells(cell_id, af, n_families, seed, mean_family_size=5.0, family_model=1, max_family_size=1000, min_fs=3,
                   qthr=20, cthr=0.6, n_threads=0):
    lib = load()
    cell_id = np.ascontiguousarray(cell_id, np.int64)
    af = np.ascontiguousarray(af, np.float64)
    nf = np.ascontiguousarray(n_families, np.int64)
    c = len(cell_id)
    n_total, n_var, n_reads = np.zeros(c, np.int64), np.zeros(c, np.int64), np.zeros(c, np.int64)
    total = lib.oracle_pipeline_cells(_p(cell_id), _p(af), _p(nf), C.c_int64(c), C.c_uint64(seed), C.c_double(mean_family_size),
                                      C.c_int(family_model), C.c_int(max_family_size), C.c_int(min_fs), C.c_int(qthr),
                                      C.c_double(cthr), _p(n_total), _p(n_var), _p(n_reads), C.c_int(n_threads))
    return {"n_total": n_total, "n_variants": n_var, "n_reads": n_reads, "total_reads": int(total)}


def call_cells(n_total, n_variants, is_negative, seed, alpha=0.05):
    lib = load()
    n_total = np.ascontiguousarray(n_total, np.int64)
    n_variants = np.ascontiguousarray(n_variants, np.int64)
    neg = np.ascontiguousarray(is_negative, np.uint8)
    c = len(n_total)
    rate = C.c_double(0)
    p, q, sig = np.zeros(c), np.zeros(c), np.zeros(c, np.uint8)
    lib.oracle_call_cells(_p(n_total), _p(n_variants), _p(neg), C.c_int64(c), C.c_uint64(seed), C.c_double(alpha),
                          C.byref(rate), _p(p), _p(q), _p(sig))
    return {"mean_error_rate": rate.value, "p_value": p, "p_adjusted": q, "significant": sig.astype(bool)}


def max_threads() -> int:
    return int(load().oracle_max_threads())
