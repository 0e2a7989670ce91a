"""Host profile of the per-cell stage pipeline (simulate_reads -> collapse_umis -> fit_error_model -> call_mrd).  Dev helper."""
import cProfile
import contextlib
import io
import pstats
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
from precise_mrd.config import load_config  # noqa: E402
from precise_mrd.eval import LODAnalyzer  # noqa: E402

cfg = load_config(ROOT / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs" / "smoke.yaml")
an = LODAnalyzer(cfg, np.random.default_rng(7), engine="stages")
c = an._create_lod_config(0.005, 5000)
with contextlib.redirect_stdout(io.StringIO()):          # (the stage functions print their progress, as the reference's do)
    for _ in range(3):
        an._run_pipeline(c, np.random.default_rng(1))
    t0 = time.perf_counter()
    for i in range(50):
        an._run_pipeline(c, np.random.default_rng(i))
    per = (time.perf_counter() - t0) / 50
    pr = cProfile.Profile()
    pr.enable()
    for i in range(50):
        an._run_pipeline(c, np.random.default_rng(i))
    pr.disable()
print(f"{per * 1e3:.2f} ms per pipeline")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
