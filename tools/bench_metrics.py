"""Bootstrap ROC-AUC / average precision (K12) against scikit-learn on the host, same resamples: time and agreement.
Secondary evidence for SURVEY.md §8f-2; not part of bench.py's contract line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
from precise_mrd import _native as nv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n_boot = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
rng = np.random.default_rng(3)
y = rng.random(n) < 0.4
score = np.round(np.where(y, rng.normal(0.012, 0.01, n), rng.normal(0.004, 0.006, n)).clip(0, None), 4)   # ties, like variant_fraction
idx = rng.choice(n, size=(n_boot, n), replace=True)                     # metrics.py:76-79
ctx = nv.Context(device=0, seed=3)
ctx.bootstrap_metrics(y, score, indices=idx[:8])
t0 = time.perf_counter()
pt, auc, ap = ctx.bootstrap_metrics(y, score, indices=idx)
dt = time.perf_counter() - t0
ctx.profile_enable(True)
ctx.bootstrap_metrics(y, score, indices=idx)
prof = ctx.profile_report()
ctx.profile_enable(False)
from sklearn.metrics import average_precision_score, roc_auc_score  # noqa: E402

m = min(n_boot, 50)
t0 = time.perf_counter()
ref_auc = np.array([roc_auc_score(y[idx[b]], score[idx[b]]) for b in range(m)])
ref_ap = np.array([average_precision_score(y[idx[b]], score[idx[b]]) for b in range(m)])
dt_ref = (time.perf_counter() - t0) / m * n_boot
print(json.dumps({"n": n, "n_boot": n_boot, "device_s_end_to_end": dt, "device_kernels_ms": sum(v[1] for v in prof.values()),
                  "sklearn_s_extrapolated": dt_ref, "speedup_end_to_end": dt_ref / dt,
                  "max_abs_diff_auc": float(np.abs(auc[:m] - ref_auc).max()), "max_abs_diff_ap": float(np.abs(ap[:m] - ref_ap).max()),
                  "h2d_bytes": int(idx.nbytes)}))
