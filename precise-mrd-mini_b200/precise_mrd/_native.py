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
ALT_TWO_SIDED, ALT_GREATER, ALT_LESS = 0, 1, 2
MODE_WEIGHTED, MODE_MAJORITY, MODE_COUNT = 0, 1, 2
CLUSTER_EXACT, CLUSTER_FIRSTFIT, CLUSTER_GREEDY = 0, 1, 2
ENGINE_MATERIALISED, ENGINE_FUSED = 0, 1
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
        ("chunk_log2", C.c_int32),
        ("engine", C.c_int32),          # ENGINE_MATERIALISED / ENGINE_FUSED (pipeline plans)
        ("reserved0", C.c_int32),
    ]


class CellKnobs(C.Structure):
    """``pmrd_cell_knobs``: optional per-cell contamination knobs + call groups of a plan."""
    _fields_ = [("hop_rate", C.c_void_p), ("hop_pool_start", C.c_void_p), ("hop_pool_size", C.c_void_p),
                ("collision_rate", C.c_void_p), ("cross_proportion", C.c_void_p), ("cross_af", C.c_void_p),
                ("call_group_start", C.c_void_p), ("n_call_groups", C.c_int64)]


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

    def profile_enable(self, on: bool = True) -> None:
        self._check(self.lib.pmrd_profile_enable(self.h, C.c_int32(1 if on else 0)))

    def peak_memory_mb(self) -> float:
        """High-water mark of the device's stream-ordered allocator pool, MB (``pmrd_peak_memory_bytes``)."""
        self.lib.pmrd_peak_memory_bytes.restype = C.c_int64
        self.lib.pmrd_peak_memory_bytes.argtypes = [C.c_void_p]
        return float(self.lib.pmrd_peak_memory_bytes(self.h)) / 1e6

    def profile_report(self, detail: bool = False):
        """{kernel name: (launches, total device ms)} since the last report (synchronises);
        ``detail=True``: ``(launches, total ms, min ms, max ms)``."""
        buf = C.create_string_buffer(1 << 16)
        self.lib.pmrd_profile_report.restype = C.c_int64
        self.lib.pmrd_profile_report.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        self.lib.pmrd_profile_report(self.h, buf, C.c_int64(len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms, mn, mx = line.rsplit(" ", 4)
            out[name] = (int(cnt), float(ms), float(mn), float(mx)) if detail else (int(cnt), float(ms))
        return out

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

    def sort_packed_cells(self, rec, cell_tile_start, key_bits: int = 24) -> np.ndarray:
        """``pmrd_sort_packed_cells``: per-cell stable sort of packed 8-byte records on their low ``key_bits`` bits."""
        rec = _arr(rec, np.uint64)
        ts = _arr(cell_tile_start, np.int64)
        out = np.empty_like(rec)
        self._check(self.lib.pmrd_sort_packed_cells(self.h, _ptr(rec), C.c_int64(rec.shape[0]), _ptr(ts), C.c_int64(ts.shape[0] - 1),
                                                    C.c_int32(key_bits), _ptr(out)))
        return out

    # -- collapse ------------------------------------------------------------------------
    def collapse_reads(self, keys, payload, params: UmiParams, want_groups: bool = True, want_families: bool = True):
        """``pmrd_collapse_reads``: host arrays in, family table and / or group table out (views of the
        call's output buffers, no extra copy).  ``want_families=False`` leaves the family rows on the
        device -- only the per-group table (``umi.py:241-266``) crosses PCIe."""
        keys = _arr(keys, np.uint64)
        payload = _arr(payload, np.uint32)
        n = keys.shape[0]
        if payload.shape[0] != n:
            raise ValueError("keys and payload must have the same length")
        cap = max(n, 1)
        fam = {}
        if want_families:
            fam = {"key": np.zeros(cap, np.uint64)}
            for name in ("size", "n_alt", "consensus", "qsum", "n_best", "flags"):
                fam[name] = np.zeros(cap, np.uint32)
            ft = FamilyTable(*[fam[nm].ctypes.data for nm, _ in FamilyTable._fields_])
        else:
            ft = FamilyTable(*[None for _ in FamilyTable._fields_])
        grp = {"group": np.zeros(cap, np.uint32), "family_count": np.zeros(cap, np.uint32),
               "consensus_count": np.zeros(cap * 4, np.uint32), "alt_count": np.zeros(cap, np.uint64),
               "total_count": np.zeros(cap, np.uint64)} if want_groups else {}
        gt = GroupTable(*[grp[nm].ctypes.data for nm, _ in GroupTable._fields_]) if want_groups else None
        nf, ng = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.pmrd_collapse_reads(
            self.h, _ptr(keys), _ptr(payload), C.c_int64(n), C.byref(params), C.byref(ft), C.c_int64(cap), C.byref(nf),
            C.byref(gt) if want_groups else None, C.c_int64(cap), C.byref(ng) if want_groups else None))
        F, G = nf.value, ng.value
        fam = {k: v[:F] for k, v in fam.items()} if want_families else {"n_families": F}
        if want_groups:
            grp = {k: (v[:G * 4].reshape(G, 4) if k == "consensus_count" else v[:G]) for k, v in grp.items()}
        else:
            grp = None
        return fam, grp

    def collapse_reads_dev(self, d_keys: int, d_payload: int, d_keys_tmp: int, d_payload_tmp: int, n: int, params: UmiParams,
                           d_fam: Optional[Dict[str, int]], capacity_f: int, d_grp: Optional[Dict[str, int]] = None, capacity_g: int = 0):
        """``pmrd_collapse_reads_dev``: every buffer is a DEVICE pointer (an int: ``tensor.data_ptr()``); the key /
        payload buffers are clobbered.  Returns ``(n_families, n_groups)``."""
        ft = FamilyTable(*[(d_fam[nm] if d_fam else None) for nm, _ in FamilyTable._fields_])     # d_fam = None: group table only
        gt = GroupTable(*[d_grp[nm] for nm, _ in GroupTable._fields_]) if d_grp else None
        nf, ng = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.pmrd_collapse_reads_dev(
            self.h, C.c_void_p(d_keys), C.c_void_p(d_payload), C.c_void_p(d_keys_tmp), C.c_void_p(d_payload_tmp), C.c_int64(n),
            C.byref(params), C.byref(ft), C.c_int64(capacity_f), C.byref(nf), C.byref(gt) if d_grp else None, C.c_int64(capacity_g),
            C.byref(ng) if d_grp else None))
        return nf.value, ng.value

    def owner_partition_dev(self, d_keys: int, d_payload: int, d_keys_tmp: int, d_payload_tmp: int, n: int, log2_world: int):
        """``pmrd_owner_partition_dev``: stable cut of this rank's reads (DEVICE pointers) into one run per owner rank
        (``group mod 2**log2_world``).  Returns ``(counts, in_tmp)``: reads bound for each rank, and whether the
        partitioned reads sit in the ``*_tmp`` arrays."""
        counts = (C.c_int64 * (1 << log2_world))()
        in_tmp = C.c_int32(0)
        self._check(self.lib.pmrd_owner_partition_dev(self.h, C.c_void_p(d_keys), C.c_void_p(d_payload), C.c_void_p(d_keys_tmp),
                                                      C.c_void_p(d_payload_tmp), C.c_int64(n), C.c_int32(log2_world), counts, C.byref(in_tmp)))
        return [int(c) for c in counts], bool(in_tmp.value)

    def owner_rekey_dev(self, d_keys: int, n: int, log2_world: int) -> None:
        """``pmrd_owner_rekey_dev``: ``group >>= log2_world`` on a DEVICE key array (dense ids after the exchange)."""
        self._check(self.lib.pmrd_owner_rekey_dev(self.h, C.c_void_p(d_keys), C.c_int64(n), C.c_int32(log2_world)))

    # -- peer-memory exchange (csrc/exchange.cu) ---------------------------------------------
    def peer_alloc(self, nbytes: int):
        """``pmrd_peer_alloc``: a device array other processes on the node can map.  Returns ``(device pointer, 64-byte handle)``."""
        ptr, handle = C.c_void_p(0), (C.c_uint8 * 64)()
        self._check(self.lib.pmrd_peer_alloc(self.h, C.c_int64(nbytes), C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def peer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p(0)
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        self._check(self.lib.pmrd_peer_open(self.h, buf, C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int) -> None:
        self._check(self.lib.pmrd_peer_close(self.h, C.c_void_p(ptr)))

    def peer_free(self, ptr: int) -> None:
        self._check(self.lib.pmrd_peer_free(self.h, C.c_void_p(ptr)))

    def owner_tiles_dev(self, d_keys: int, n: int, log2_world: int, d_tile_off: int):
        """``pmrd_owner_tiles_dev``: fills ``d_tile_off`` (uint32 ``[2**log2_world][ceil(n / 4096)]``), returns reads per owner."""
        counts = (C.c_int64 * (1 << log2_world))()
        self._check(self.lib.pmrd_owner_tiles_dev(self.h, C.c_void_p(d_keys), C.c_int64(n), C.c_int32(log2_world), C.c_void_p(d_tile_off), counts))
        return [int(c) for c in counts]

    def owner_scatter_peer_dev(self, d_keys: int, d_payload: int, n: int, log2_world: int, d_tile_off: int, peer_keys, peer_payload, base) -> None:
        """``pmrd_owner_scatter_peer_dev``: the cut, stored straight into the owners' (peer) arrays."""
        w = 1 << log2_world
        pk = (C.c_void_p * w)(*[C.c_void_p(int(p)) for p in peer_keys])
        pp = (C.c_void_p * w)(*[C.c_void_p(int(p)) for p in peer_payload])
        hb = (C.c_int64 * w)(*[int(b) for b in base])
        self._check(self.lib.pmrd_owner_scatter_peer_dev(self.h, C.c_void_p(d_keys), C.c_void_p(d_payload), C.c_int64(n), C.c_int32(log2_world),
                                                         C.c_void_p(d_tile_off), pk, pp, hb))

    def umi_hamming(self, a, b, umi_len: int) -> np.ndarray:
        a = _arr(a, np.uint32)
        b = _arr(b, np.uint32)
        d = np.empty(a.shape[0], np.int32)
        self._check(self.lib.pmrd_umi_hamming(self.h, _ptr(a), _ptr(b), C.c_int64(a.shape[0]), C.c_int32(umi_len), _ptr(d)))
        return d


    # -- K1 ------------------------------------------------------------------------------
    def simulate_cells(self, cell_id, af, depth, max_family_size: int, bg_multiplier=None) -> Dict[str, np.ndarray]:
        cell_id = _arr(cell_id, np.int64)
        af = _arr(af, np.float64)
        depth = _arr(depth, np.int64)
        c = cell_id.shape[0]
        mult = None if bg_multiplier is None else _arr(np.broadcast_to(bg_multiplier, (c,)), np.float64)
        out = {"background_rate": np.zeros(c), "mean_family_size": np.zeros(c, np.int64),
               "n_true_variants": np.zeros(c, np.int64), "n_false_positives": np.zeros(c, np.int64),
               "mean_quality": np.zeros(c)}
        t = CellTable(*[out[n].ctypes.data for n, _ in CellTable._fields_])
        self._check(self.lib.pmrd_simulate_cells_stratified(self.h, _ptr(cell_id), _ptr(af), _ptr(depth),
                                                            None if mult is None else _ptr(mult), C.c_int64(c),
                                                            C.c_int32(max_family_size), C.byref(t)))
        return out

    # -- K2 (G-D) ------------------------------------------------------------------------
    def collapse_families(self, cell_id, af, n_families, background_rate, mean_family_size, params: UmiParams,
                          variant: int = 2, count_only: bool = False):
        cell_id = _arr(cell_id, np.int64)
        af = _arr(af, np.float64)
        nf = _arr(n_families, np.int64)
        bg = _arr(background_rate, np.float64)
        mfs = _arr(mean_family_size, np.float64)
        c = cell_id.shape[0]
        offs = np.zeros(c + 1, np.int64)
        args = (self.h, _ptr(cell_id), _ptr(af), _ptr(nf), _ptr(bg), _ptr(mfs), C.c_int64(c), C.byref(params), C.c_int32(variant))
        self._check(self.lib.pmrd_collapse_families(*args, _ptr(offs), None, C.c_int64(0)))
        if count_only:
            return offs, None
        R = int(offs[-1])
        cap = max(R, 1)
        out = {"cell": np.zeros(cap, np.int32), "family_id": np.zeros(cap, np.int32), "family_size": np.zeros(cap, np.int32),
               "quality_score": np.zeros(cap), "consensus_agreement": np.zeros(cap), "flags": np.zeros(cap, np.uint8)}
        t = GdFamilyTable(*[out[n].ctypes.data for n, _ in GdFamilyTable._fields_])
        self._check(self.lib.pmrd_collapse_families(*args, _ptr(offs), C.byref(t), C.c_int64(cap)))
        return offs, {k: v[:R] for k, v in out.items()}

    # -- K10 -----------------------------------------------------------------------------
    def error_model(self, n_neg: int, n_var: int, n_boot: int = 100, stream_id: int = 0):
        r, lo, hi = np.zeros(64), np.zeros(64), np.zeros(64)
        self._check(self.lib.pmrd_error_model(self.h, C.c_int64(n_neg), C.c_int64(n_var), C.c_int32(n_boot),
                                              C.c_uint64(stream_id), _ptr(r), _ptr(lo), _ptr(hi)))
        return r, lo, hi

    # -- read-level simulate -------------------------------------------------------------
    def simulate_reads_count(self, cell_id, n_families, sp: ReadSimParams) -> np.ndarray:
        cell_id = _arr(cell_id, np.int64)
        nf = _arr(n_families, np.int64)
        offs = np.zeros(cell_id.shape[0] + 1, np.int64)
        self._check(self.lib.pmrd_simulate_reads_count(self.h, _ptr(cell_id), _ptr(nf), C.c_int64(cell_id.shape[0]),
                                                       C.byref(sp), _ptr(offs)))
        return offs

    def simulate_reads(self, cell_id, af, n_families, sp: ReadSimParams, group_base: int = 0):
        cell_id = _arr(cell_id, np.int64)
        af = _arr(af, np.float64)
        nf = _arr(n_families, np.int64)
        n = int(self.simulate_reads_count(cell_id, nf, sp)[-1])
        keys = np.zeros(max(n, 1), np.uint64)
        payload = np.zeros(max(n, 1), np.uint32)
        got = C.c_int64(0)
        self._check(self.lib.pmrd_simulate_reads(self.h, _ptr(cell_id), _ptr(af), _ptr(nf), C.c_int64(cell_id.shape[0]),
                                                 C.c_uint32(group_base), C.byref(sp), _ptr(keys), _ptr(payload),
                                                 C.c_int64(max(n, 1)), C.byref(got)))
        return keys[:got.value], payload[:got.value]

    # -- fused pipeline ------------------------------------------------------------------
    def pipeline_cells(self, cell_id, af, n_families, sp: ReadSimParams, up: UmiParams, neg_af_max: float = 1e-4,
                       test_type: int = TEST_POISSON, alternative: int = ALT_TWO_SIDED, alpha: float = 0.05, knobs=None):
        plan = Plan(self, cell_id, af, n_families, sp, up, neg_af_max, test_type, alternative, alpha, knobs=knobs)
        try:
            plan.run()
            return plan.results()
        finally:
            plan.close()

    def call_counts(self, n_total, n_variants, is_negative, stream_id: int, test_type: int = TEST_POISSON,
                    alternative: int = ALT_TWO_SIDED, alpha: float = 0.05) -> Dict[str, np.ndarray]:
        """The pipeline's call stage over per-cell counts (``pmrd_call_counts``): what every rank runs
        after the all-gather of a grid sharded over several GPUs."""
        tot, var = _arr(n_total, np.int64), _arr(n_variants, np.int64)
        neg = np.ascontiguousarray(np.asarray(is_negative).astype(np.uint8))
        c = tot.shape[0]
        p, q, sig, rate = np.zeros(c), np.zeros(c), np.zeros(c, np.uint8), np.zeros(1)
        self._check(self.lib.pmrd_call_counts(self.h, _ptr(tot), _ptr(var), _ptr(neg), C.c_int64(c), C.c_uint64(stream_id),
                                              C.c_int32(test_type), C.c_int32(alternative), C.c_double(alpha), _ptr(p), _ptr(q),
                                              _ptr(sig), _ptr(rate)))
        return {"p_value": p, "p_adjusted": q, "significant": sig.astype(bool), "mean_error_rate": float(rate[0])}

    # -- K15 / K16 -----------------------------------------------------------------------
    def feature_table(self, family_size, quality_score, consensus_agreement, allele_fraction) -> np.ndarray:
        """[n, 6] feature matrix of FeatureEngineer.extract_features (``pmrd_feature_table``)."""
        sz = _arr(family_size, np.int64)
        q, a, f = _arr(quality_score, np.float64), _arr(consensus_agreement, np.float64), _arr(allele_fraction, np.float64)
        n = sz.shape[0]
        if not (q.shape[0] == a.shape[0] == f.shape[0] == n):
            raise ValueError("feature columns must have one length")
        out = np.zeros((n, 6))
        self._check(self.lib.pmrd_feature_table(self.h, _ptr(sz), _ptr(q), _ptr(a), _ptr(f), C.c_int64(n), _ptr(out)))
        return out

    def qc_family_stats(self, family_size, flags=None, flag_mask: int = FAM_HAS_CONSENSUS, n_bins: int = 4096):
        """(hist[n_bins] of family sizes (last bin = larger), families meeting flag_mask, largest size)
        (``pmrd_qc_family_stats``)."""
        sz = np.ascontiguousarray(family_size, dtype=np.uint32)
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint32)
        hist, cnt = np.zeros(n_bins, np.uint64), np.zeros(2, np.uint64)
        self._check(self.lib.pmrd_qc_family_stats(self.h, _ptr(sz), _ptr(fl), C.c_uint32(flag_mask), C.c_int64(sz.shape[0]),
                                                  C.c_int32(n_bins), _ptr(hist), _ptr(cnt)))
        return hist.astype(np.int64), int(cnt[0]), int(cnt[1])

    # -- K14 -----------------------------------------------------------------------------
    def fisher_exact(self, tables) -> Tuple[np.ndarray, np.ndarray]:
        """(odds_ratio[m], p[m]) of m 2x2 tables given as [m, 4] = c00 c01 c10 c11 (two-sided)."""
        t = _arr(tables, np.int64).reshape(-1, 4)
        m = t.shape[0]
        odds, p = np.zeros(m), np.zeros(m)
        self._check(self.lib.pmrd_fisher_exact(self.h, _ptr(t), C.c_int64(m), _ptr(odds), _ptr(p)))
        return odds, p

    # -- K13 -----------------------------------------------------------------------------
    _FQ_REC = (("status", np.int32), ("umi_key", np.uint64), ("hdr_off", np.int64), ("hdr_len", np.int32), ("umi_off", np.int64),
               ("umi_len", np.int32), ("seq_off", np.int64), ("seq_len", np.int32), ("qual_off", np.int64), ("qual_len", np.int32),
               ("qual_sum", np.int64))
    _FQ_FAM = (("umi_key", np.uint64), ("first_read", np.uint32), ("best_read", np.uint32), ("family_size", np.int64),
               ("qual_sum", np.int64), ("n_bases", np.int64))

    def fastq_scan(self, data: bytes) -> Dict[str, np.ndarray]:
        """One row per 4-line FASTQ record of ``data`` (``pmrd_fastq_scan``)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        cap = buf.shape[0] // 8 + 2           # a record has >= 4 newlines; an upper bound that needs no second call
        out = {k: np.zeros(cap, dt) for k, dt in self._FQ_REC}
        t = (C.c_void_p * len(self._FQ_REC))(*[out[k].ctypes.data for k, _ in self._FQ_REC])
        n = C.c_int64(0)
        self._check(self.lib.pmrd_fastq_scan(self.h, _ptr(buf) if buf.shape[0] else None, C.c_int64(buf.shape[0]), C.c_int64(cap),
                                             C.byref(t), C.byref(n)))
        return {k: v[:n.value] for k, v in out.items()}

    def fastq_families(self, umi_key, qual_sum, qual_len) -> Dict[str, np.ndarray]:
        """Per-UMI families of the kept reads, in first-appearance order (``pmrd_fastq_families_of``)."""
        key = _arr(umi_key, np.uint64)
        qs = _arr(qual_sum, np.int64)
        ql = _arr(qual_len, np.int32)
        n = key.shape[0]
        out = {k: np.zeros(max(n, 1), dt) for k, dt in self._FQ_FAM}
        t = (C.c_void_p * len(self._FQ_FAM))(*[out[k].ctypes.data for k, _ in self._FQ_FAM])
        nf = C.c_int64(0)
        self._check(self.lib.pmrd_fastq_families_of(self.h, _ptr(key), _ptr(qs), _ptr(ql), C.c_int64(n), C.byref(t), C.c_int64(n),
                                                    C.byref(nf)))
        return {k: v[:nf.value] for k, v in out.items()}

    # -- K12 -----------------------------------------------------------------------------
    def bootstrap_metrics(self, y_true, y_score, indices=None, n_boot: int = 0, stream_id: int = 0, point: bool = True):
        """(point {auc, ap} or None, auc[n_boot], ap[n_boot]).  ``indices``: [n_boot, n] resample rows."""
        y = _arr(np.asarray(y_true) != 0, np.uint8)
        sc = _arr(y_score, np.float64)
        n = y.shape[0]
        ix = None
        if indices is not None:
            ix = _arr(indices, np.int64).reshape(-1, n)
            n_boot = ix.shape[0]
        auc, ap = np.zeros(max(n_boot, 1)), np.zeros(max(n_boot, 1))
        pt = np.zeros(2) if point else None
        self._check(self.lib.pmrd_bootstrap_metrics(self.h, _ptr(y), _ptr(sc), C.c_int64(n), _ptr(ix), C.c_int64(n_boot),
                                                    C.c_uint64(stream_id), _ptr(pt), _ptr(auc), _ptr(ap)))
        return pt, auc[:n_boot], ap[:n_boot]

    # -- K11 -----------------------------------------------------------------------------
    def fit_lod(self, af, hit, target: float = 0.95, n_boot: int = 200, stream_id: int = 0):
        """(lod, ci_lo, ci_hi) of one detection curve (``pmrd_fit_lod``)."""
        return self.fit_lod_batch(af, np.asarray(hit, dtype=np.float64)[None, :], target, n_boot, [stream_id])[0]

    def fit_lod_batch(self, af, hits, target: float = 0.95, n_boot: int = 200, stream_ids=None):
        """[(lod, ci_lo, ci_hi)] of ``hits.shape[0]`` detection curves over the same AF grid, fitted by one
        launch (``pmrd_fit_lod_batch``); equal to separate ``fit_lod`` calls bit for bit."""
        af = _arr(af, np.float64)
        hits = np.ascontiguousarray(np.asarray(hits, dtype=np.float64))
        if hits.ndim != 2 or hits.shape[1] != af.shape[0]:
            raise ValueError("hits must be [n_series, len(af)]")
        s = hits.shape[0]
        ids = np.zeros(s, np.uint64) if stream_ids is None else np.ascontiguousarray(np.asarray(stream_ids, dtype=np.uint64))
        if ids.shape[0] != s:
            raise ValueError("one stream id per series")
        out = np.zeros((s, 3))
        self._check(self.lib.pmrd_fit_lod_batch(self.h, _ptr(af), _ptr(hits), C.c_int32(af.shape[0]), C.c_int32(s), C.c_double(target),
                                                C.c_int32(n_boot), _ptr(ids), _ptr(out)))
        return [(float(r[0]), float(r[1]), float(r[2])) for r in out]


class Plan:
    """Prepared simulate -> collapse -> call pipeline over a batch of cells (``pmrd_plan``)."""

    def __init__(self, ctx: Context, cell_id, af, n_families, sp: ReadSimParams, up: UmiParams, neg_af_max: float = 1e-4,
                 test_type: int = TEST_POISSON, alternative: int = ALT_TWO_SIDED, alpha: float = 0.05,
                 knobs: Optional[Dict[str, np.ndarray]] = None):
        """``knobs`` (optional, ``pmrd_cell_knobs``): per-cell arrays ``hop_rate`` (+ ``hop_pool_start`` / ``hop_pool_size``),
        ``collision_rate``, ``cross_proportion`` (+ ``cross_af``) and ``call_group_start`` (G + 1 cell offsets)."""
        self.ctx = ctx
        self.cell_id = _arr(cell_id, np.int64)
        self.af = _arr(af, np.float64)
        self.nf = _arr(n_families, np.int64)
        self.n_cells = self.cell_id.shape[0]
        h = C.c_void_p()
        ctx.lib.pmrd_plan_n_reads.restype = C.c_int64
        ctx.lib.pmrd_plan_n_reads.argtypes = [C.c_void_p]
        ctx.lib.pmrd_plan_n_families.restype = C.c_int64
        ctx.lib.pmrd_plan_n_families.argtypes = [C.c_void_p]
        ctx.lib.pmrd_plan_destroy.restype = None
        ctx.lib.pmrd_plan_destroy.argtypes = [C.c_void_p]
        self.n_groups = 1
        if knobs:
            typ = {"hop_rate": np.float64, "hop_pool_start": np.int64, "hop_pool_size": np.int64, "collision_rate": np.float64,
                   "cross_proportion": np.float64, "cross_af": np.float64, "call_group_start": np.int64}
            unknown = set(knobs) - set(typ)
            if unknown:
                raise ValueError(f"unknown plan knobs: {sorted(unknown)}")
            self._knobs = {k: _arr(v, typ[k]) for k, v in knobs.items() if v is not None}
            for k, v in self._knobs.items():
                if k != "call_group_start" and v.shape[0] != self.n_cells:
                    raise ValueError(f"knob {k} must hold one value per cell")
            kn = CellKnobs(*[self._knobs[k].ctypes.data if k in self._knobs else None for k in typ], 0)
            if "call_group_start" in self._knobs:
                self.n_groups = int(self._knobs["call_group_start"].shape[0] - 1)
                kn.n_call_groups = self.n_groups
            ctx._check(ctx.lib.pmrd_plan_create_ex(ctx.h, _ptr(self.cell_id), _ptr(self.af), _ptr(self.nf), C.c_int64(self.n_cells),
                                                   C.byref(sp), C.byref(up), C.c_double(neg_af_max), C.c_int32(test_type),
                                                   C.c_int32(alternative), C.c_double(alpha), C.byref(kn), C.byref(h)))
        else:
            ctx._check(ctx.lib.pmrd_plan_create(ctx.h, _ptr(self.cell_id), _ptr(self.af), _ptr(self.nf), C.c_int64(self.n_cells),
                                                C.byref(sp), C.byref(up), C.c_double(neg_af_max), C.c_int32(test_type),
                                                C.c_int32(alternative), C.c_double(alpha), C.byref(h)))
        self.h = h
        self.n_reads = int(ctx.lib.pmrd_plan_n_reads(h))
        self.n_families = int(ctx.lib.pmrd_plan_n_families(h))

    def run(self) -> None:
        self.ctx._check(self.ctx.lib.pmrd_plan_run(self.ctx.h, self.h))

    def run_simulate(self, chunk: int = 0) -> None:
        self.ctx._check(self.ctx.lib.pmrd_plan_run_simulate(self.ctx.h, self.h, C.c_int64(chunk)))

    def run_collapse(self, chunk: int = 0) -> None:
        self.ctx._check(self.ctx.lib.pmrd_plan_run_collapse(self.ctx.h, self.h, C.c_int64(chunk)))

    @property
    def n_chunks(self) -> int:
        self.ctx.lib.pmrd_plan_n_chunks.restype = C.c_int64
        self.ctx.lib.pmrd_plan_n_chunks.argtypes = [C.c_void_p]
        return int(self.ctx.lib.pmrd_plan_n_chunks(self.h))

    def info(self) -> Dict[str, int]:
        o = np.zeros(8, np.int64)
        self.ctx._check(self.ctx.lib.pmrd_plan_info(self.h, _ptr(o)))
        names = ("n_reads", "n_slots", "n_families", "n_chunks", "sort_passes", "record_bytes", "max_chunk_slots", "segmented")
        return {k: int(v) for k, v in zip(names, o)}

    def chunk_reads(self, chunk: int) -> int:
        self.ctx.lib.pmrd_plan_chunk_reads.restype = C.c_int64
        self.ctx.lib.pmrd_plan_chunk_reads.argtypes = [C.c_void_p, C.c_int64]
        return int(self.ctx.lib.pmrd_plan_chunk_reads(self.h, C.c_int64(chunk)))

    def run_call(self) -> None:
        self.ctx._check(self.ctx.lib.pmrd_plan_run_call(self.ctx.h, self.h))

    def results(self) -> Dict[str, np.ndarray]:
        c = self.n_cells
        out = {"n_reads": np.zeros(c, np.int64), "n_families": np.zeros(c, np.int64), "n_total": np.zeros(c, np.int64),
               "n_variants": np.zeros(c, np.int64), "p_value": np.zeros(c), "p_adjusted": np.zeros(c),
               "significant": np.zeros(c, np.uint8)}
        t = PipelineResult(*[out[n].ctypes.data for n, _ in PipelineResult._fields_])
        rate = np.zeros(1)
        totals = np.zeros(3, np.int64)
        self.ctx._check(self.ctx.lib.pmrd_plan_results(self.ctx.h, self.h, C.byref(t), _ptr(rate), _ptr(totals)))
        out["significant"] = out["significant"].astype(bool)
        out["mean_error_rate"] = float(rate[0])
        out["totals"] = totals
        if self.n_groups > 1:
            rates = np.zeros(self.n_groups)
            self.ctx._check(self.ctx.lib.pmrd_plan_group_rates(self.ctx.h, self.h, _ptr(rates), C.c_int64(self.n_groups)))
            out["group_error_rate"] = rates
        return out

    def close(self) -> None:
        if getattr(self, "h", None):
            self.ctx.lib.pmrd_plan_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


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
