"""Shared pytest plumbing.

* ``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol check (no GPU needed).
* ``-m gpu``     : parity tests proper -- every one calls through the C ABI of
                   ``precise-mrd-mini_b200/lib/libpmrd_b200.so`` on a real B200.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG_DIR = ROOT / "precise-mrd-mini_b200"
for p in (str(PKG_DIR), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver on the GPU box)")
    config.addinivalue_line("markers", "slow: long-running")


def load_golden(name: str):
    return np.load(GOLDEN / name, allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def ctx():
    """One native context for the whole GPU session.  Fails loudly (no skip, no fallback)
    if the library or the device is missing -- a `-m gpu` run without native code is a bug."""
    from precise_mrd import _native

    c = _native.Context(device=int(os.environ.get("LOCAL_RANK", "0")), seed=7)
    yield c
    c.close()


@pytest.fixture(scope="session")
def native():
    from precise_mrd import _native

    return _native
