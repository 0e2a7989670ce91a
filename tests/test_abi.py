"""The C-ABI library loads and exports every symbol include/pmrd.h declares (no GPU needed:
no compute entry point is called), and compute fails loudly -- never silently -- without a GPU."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "pmrd.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmrd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from precise_mrd import _native

    lib = _native.load_library()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/pmrd.h but not exported: {missing}"
    lib.pmrd_abi_version.restype = ctypes.c_int32
    assert lib.pmrd_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch

    from precise_mrd import _native

    if torch.cuda.is_available():
        pytest.skip("GPU present: the fallback check is for the CPU box")
    with pytest.raises(_native.NativeUnavailable):
        _native.Context(device=0, seed=1)


def test_header_cites_reference_lines():
    text = (ROOT / "include" / "pmrd.h").read_text()
    for ref in ("call.py:14-25", "call.py:41-79", "umi.py:146-179", "mrd/umi.py:29-40", "simulate.py:133-180",
                "collapse.py:70-118", "error_model.py:139-186", "io.py:79-123", "eval/lod.py"):
        assert ref in text, ref
