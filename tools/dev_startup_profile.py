"""Where a CLI process spends its start-up (imports, CUDA context, first launch).  Dev helper."""
import sys
import time
from pathlib import Path

t0 = time.perf_counter()
import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402,F401

t1 = time.perf_counter()
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "precise-mrd-mini_b200"))
import precise_mrd.cli  # noqa: E402,F401
from precise_mrd import _native as nv  # noqa: E402

t2 = time.perf_counter()
ctx = nv.Context(device=0, seed=7)
t3 = time.perf_counter()
p = ctx.pvalues(np.array([3]), np.array([1000]), 1e-3, nv.TEST_POISSON)
t4 = time.perf_counter()
print(f"numpy+pandas {t1 - t0:.2f} s, precise_mrd (click, jsonschema, yaml, scipy...) {t2 - t1:.2f} s, "
      f"pmrd_create (CUDA context, tables) {t3 - t2:.2f} s, first kernel {t4 - t3:.3f} s; torch imported: {'torch' in sys.modules}")
