"""``set_global_seed`` / ``env_fingerprint`` (``determinism_utils/seed.py:15-121``)."""
from __future__ import annotations

import os
import platform
import random
import subprocess
import sys
from typing import Any

import numpy as np


def set_global_seed(seed: int, deterministic_ops: bool = True) -> np.random.Generator:
    """Seeds Python / NumPy (and torch when present) and returns ``np.random.default_rng(seed)``.
    The returned generator is what the stage functions draw their Philox salt from."""
    os.environ["PYTHONHASHSEED"] = str(seed)      # no effect after start-up; kept for parity
    random.seed(seed)
    rng = np.random.default_rng(seed)
    # the reference also seeds torch "when present" (seed.py:40-47).  Nothing on this path draws from torch, so it is
    # seeded only if the caller has imported it -- importing it here would cost every CLI process ~2.5 s
    if "torch" in sys.modules:
        try:
            torch = sys.modules["torch"]
            torch.manual_seed(seed)
            if torch.cuda.is_available():
                torch.cuda.manual_seed_all(seed)
        except Exception:
            pass
    return rng


def env_fingerprint() -> dict[str, Any]:
    def git_sha() -> str:
        try:
            return subprocess.check_output(["git", "rev-parse", "HEAD"], text=True, timeout=5, stderr=subprocess.DEVNULL).strip()
        except Exception:
            return "unknown"

    fp: dict[str, Any] = {"python_version": sys.version, "platform": platform.platform(), "git_sha": git_sha(),
                          "numpy_version": np.__version__}
    try:
        import pandas as pd

        fp["pandas_version"] = pd.__version__
    except ImportError:
        pass
    try:
        import scipy

        fp["scipy_version"] = scipy.__version__
    except ImportError:
        pass
    # same keys as the reference's fingerprint, filled without importing torch (2.5 s): the version from the
    # package metadata, the device from the native library this path actually runs on
    try:
        from importlib import metadata

        fp["torch_version"] = metadata.version("torch")
    except Exception:
        fp["torch_version"] = None
    try:
        from .. import _native as nv

        nv.default_context()
        fp["cuda_available"] = True
    except Exception:
        fp["cuda_available"] = False
    fp["cpu_count"] = os.cpu_count() or 1
    fp["machine"] = platform.machine()
    return fp
