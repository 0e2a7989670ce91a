"""ctypes binding of ``libpmrd_b200.so`` (C ABI: ``include/pmrd.h``).

This is the only door between the Python host mirror of the reference API and the CUDA
kernels.  There is **no CPU fallback**: if the shared library is missing or no B200 is
visible, every compute entry point raises :class:`NativeUnavailable`.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np

_PKG_ROOT = Path(__file__).resolve().parent.parent  # precise-mrd-mini_b200/
LIB_PATH = Path(os.environ.get("PMRD_B200_LIB", _PKG_ROOT / "lib" / "libpmrd_b200.so"))

PMRD_OK = 0
ERR_CUDA, ERR_ARG, ERR_CAPACITY, ERR_NO_DEVICE = -1, -2, -3, -4
TEST_POISSON, TEST_BINOMIAL = 0, 1
ALT_TWO_SIDED, ALT_GREATER = 0, 1
MODE_WEIGHTED, MODE_MAJORITY, MODE_COUNT = 0, 1, 2
CLUSTER_EXACT, CLUSTER_FIRSTFIT = 0, 1
FAM_HAS_CONSENSUS, FAM_KEPT, FAM_OVERSIZE = 1, 2, 4

TEST_TYPES = {"poisson": TEST_POISSON, "binomial": TEST_BINOMIAL}


class NativeUnavailable(RuntimeError):
    """libpmrd_b200.so or a B200 device is missing -- the product has no other path."""


class NativeError(RuntimeError):
    """A C-ABI call returned a non-zero status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libpmrd_b200 status {status}: {message}")
        self.status = status
        self.message = message


class UmiParams(C.Structure):
    _fields_ = [
        ("min_family_size", C.c_int32),
        ("max_family_size", C.c_int32),
        ("quality_threshold", C.c_int32),
        ("consensus_threshold", C.c_double),
        ("mode", C.c_int32),
        ("cluster", C.c_int32),
        ("max_distance", C.c_int32),
        ("umi_len", C.c_int32),
    ]


class ReadSimParams(C.Structure):
    _fields_ = [
        ("mean_family_size", C.c_double),
        ("family_model", C.c_int32),
        ("max_family_size", C.c_int32),
        ("contamination_rate", C.c_double),
        ("hop_rate", C.c_double),
        ("collision_rate", C.c_double),
        ("shuffle", C.c_int32),
        ("reserved", C.c_int32),
    ]


class CallTable(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "n_variants", "n_total", "variant_fraction", "expected_variants", "fold_change",
        "p_value", "mean_quality", "mean_consensus", "p_adjusted", "significant")]


class FamilyTable(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("key", "size", "n_alt", "consensus", "qsum", "n_best", "flags")]


class GroupTable(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("group", "family_count", "consensus_count", "alt_count", "total_count")]


class CellTable(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "background_rate", "mean_family_size", "n_true_variants", "n_false_positives", "mean_quality")]


class GdFamilyTable(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "cell", "family_id", "family_size", "quality_score", "consensus_agreement", "flags")]


class PipelineResult(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "n_reads", "n_families", "n_total", "n_variants", "p_value", "p_adjusted", "significant")]


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen the in-tree library (built by ``__graft_entry__.build()`` / ``csrc/Makefile``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeUnavailable(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). precise_mrd (B200) has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    lib.pmrd_last_error.restype = C.c_char_p
    lib.pmrd_last_error.argtypes = [C.c_void_p]
    lib.pmrd_launch_count.restype = C.c_int64
    lib.pmrd_launch_count.argtypes = [C.c_void_p]
    lib.pmrd_destroy.restype = None
    lib.pmrd_destroy.argtypes = [C.c_void_p]
    lib.pmrd_create.argtypes = [C.c_int32, C.c_uint64, C.POINTER(C.c_void_p)]
    _lib = lib
    return lib


def _ptr(a: Optional[np.ndarray]) -> C.c_void_p:
    return C.c_void_p(0) if a is None else C.c_void_p(a.ctypes.data)


def _arr(x, dtype) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=dtype)


class Context:
    """One CUDA device + one stream + one Philox seed (``pmrd_ctx``)."""

    def __init__(self, device: int = 0, seed: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.pmrd_create(C.c_int32(device), C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), C.byref(h))
        if st != PMRD_OK:
            msg = self.lib.pmrd_last_error(None).decode()
            if st == ERR_NO_DEVICE:
                raise NativeUnavailable(msg)
            raise NativeError(st, msg)
        self.h = h
        self.device = device
        self.seed = seed

    # -- plumbing ------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.pmrd_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int) -> None:
        if st != PMRD_OK:
            raise NativeError(st, self.lib.pmrd_last_error(self.h).decode())

    def set_seed(self, seed: int) -> None:
        self._check(self.lib.pmrd_set_seed(self.h, C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF)))
        self.seed = seed

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self.lib.pmrd_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self) -> None:
        self._check(self.lib.pmrd_sync(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.pmrd_launch_count(self.h))

    # -- K8 ------------------------------------------------------------------------------
    def pvalues(self, k, n, rate, test_type: int, alternative: int = ALT_TWO_SIDED) -> np.ndarray:
        k = _arr(k, np.int64)
        m = k.shape[0]
        n_a = None if n is None else _arr(n, np.int64)
        rate = np.atleast_1d(_arr(rate, np.float64))
        stride = 0 if rate.shape[0] == 1 and m != 1 else 1
        if stride == 1 and rate.shape[0] != m:
            raise ValueError("rate must be a scalar or have one entry per test")
        out = np.empty(m, dtype=np.float64)
        self._check(self.lib.pmrd_pvalues(self.h, _ptr(k), _ptr(n_a), _ptr(rate), C.c_int64(stride), C.c_int64(m),
                                          C.c_int32(test_type), C.c_int32(alternative), _ptr(out)))
        return out

    # -- K9 ------------------------------------------------------------------------------
    def bh(self, p, alpha: float) -> Tuple[np.ndarray, np.ndarray]:
        p = _arr(p, np.float64)
        m = p.shape[0]
        rej = np.zeros(m, dtype=np.uint8)
        q = np.empty(m, dtype=np.float64)
        self._check(self.lib.pmrd_bh(self.h, _ptr(p), C.c_int64(m), C.c_double(alpha), _ptr(rej), _ptr(q)))
        return rej.astype(bool), q

    # -- call ----------------------------------------------------------------------------
    def call(self, seg_offsets, is_variant, quality, agreement, mean_error_rate: float, test_type: int,
             alpha: float) -> Dict[str, np.ndarray]:
        so = _arr(seg_offsets, np.int64)
        S = so.shape[0] - 1
        iv = _arr(is_variant, np.uint8)
        q = None if quality is None else _arr(quality, np.float64)
        a = None if agreement is None else _arr(agreement, np.float64)
        out = {
            "n_variants": np.zeros(S, np.int64), "n_total": np.zeros(S, np.int64),
            "variant_fraction": np.zeros(S), "expected_variants": np.zeros(S), "fold_change": np.zeros(S),
            "p_value": np.zeros(S), "mean_quality": np.zeros(S), "mean_consensus": np.zeros(S),
            "p_adjusted": np.zeros(S), "significant": np.zeros(S, np.uint8),
        }
        t = CallTable(*[out[n].ctypes.data for n, _ in CallTable._fields_])
        self._check(self.lib.pmrd_call(self.h, _ptr(so), C.c_int64(S), _ptr(iv), _ptr(q), _ptr(a),
                                       C.c_double(mean_error_rate), C.c_int32(test_type), C.c_double(alpha), C.byref(t)))
        out["significant"] = out["significant"].astype(bool)
        return out

    # -- K4 ------------------------------------------------------------------------------
    def sort_pairs(self, keys, vals=None):
        keys = _arr(keys, np.uint64)
        n = keys.shape[0]
        v = None if vals is None else _arr(vals, np.uint32)
        ko = np.empty(n, np.uint64)
        vo = None if vals is None else np.empty(n, np.uint32)
        self._check(self.lib.pmrd_sort_pairs(self.h, _ptr(keys), _ptr(v), C.c_int64(n), _ptr(ko), _ptr(vo)))
        return ko, vo

    # -- collapse ------------------------------------------------------------------------
    def collapse_reads(self, keys, payload, params: UmiParams, want_groups: bool = True):
        keys = _arr(keys, np.uint64)
        payload = _arr(payload, np.uint32)
        n = keys.shape[0]
        if payload.shape[0] != n:
            raise ValueError("keys and payload must have the same length")
        cap = max(n, 1)
        fam = {"key": np.zeros(cap, np.uint64)}
        for name in ("size", "n_alt", "consensus", "qsum", "n_best", "flags"):
            fam[name] = np.zeros(cap, np.uint32)
        ft = FamilyTable(*[fam[nm].ctypes.data for nm, _ in FamilyTable._fields_])
        grp = {"group": np.zeros(cap, np.uint32), "family_count": np.zeros(cap, np.uint32),
               "consensus_count": np.zeros(cap * 4, np.uint32), "alt_count": np.zeros(cap, np.uint64),
               "total_count": np.zeros(cap, np.uint64)}
        gt = GroupTable(*[grp[nm].ctypes.data for nm, _ in GroupTable._fields_])
        nf, ng = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.pmrd_collapse_reads(
            self.h, _ptr(keys), _ptr(payload), C.c_int64(n), C.byref(params), C.byref(ft), C.c_int64(cap), C.byref(nf),
            C.byref(gt) if want_groups else None, C.c_int64(cap), C.byref(ng) if want_groups else None))
        F, G = nf.value, ng.value
        fam = {k: v[:F].copy() for k, v in fam.items()}
        if want_groups:
            grp = {k: (v[:G * 4].reshape(G, 4).copy() if k == "consensus_count" else v[:G].copy()) for k, v in grp.items()}
        else:
            grp = None
        return fam, grp

    def umi_hamming(self, a, b, umi_len: int) -> np.ndarray:
        a = _arr(a, np.uint32)
        b = _arr(b, np.uint32)
        d = np.empty(a.shape[0], np.int32)
        self._check(self.lib.pmrd_umi_hamming(self.h, _ptr(a), _ptr(b), C.c_int64(a.shape[0]), C.c_int32(umi_len), _ptr(d)))
        return d


_default_ctx: Optional[Context] = None


def default_context(seed: Optional[int] = None) -> Context:
    """Process-wide context on ``cuda:LOCAL_RANK`` (one process per GPU)."""
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("PMRD_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _default_ctx = Context(device=dev, seed=0 if seed is None else seed)
    if seed is not None and seed != _default_ctx.seed:
        _default_ctx.set_seed(seed)
    return _default_ctx
