"""Where the wall time of the eval-lod grid goes (host profile + device kernel time).  Dev helper."""
import contextlib
import cProfile
import io
import pstats
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
from precise_mrd import _native as nv  # noqa: E402
from precise_mrd.config import load_config  # noqa: E402
from precise_mrd.eval import LODAnalyzer  # noqa: E402

cfg = load_config(ROOT / "precise-mrd-mini_b200" / "precise_mrd" / "assets" / "configs" / "default.yaml")
cfg.seed = 7
depths = [1000, 5000, 10000, 50000, 100000]


def once():
    an = LODAnalyzer(cfg, np.random.default_rng(7), engine="grid")
    with contextlib.redirect_stdout(io.StringIO()):
        an.estimate_lob(n_blank_runs=12)
        an.estimate_lod(depth_values=depths, n_replicates=25)
        with tempfile.TemporaryDirectory() as td:
            an.generate_reports(td)


once()
ctx = nv.default_context(7)
ctx.profile_enable(True)
t0 = time.perf_counter()
once()
wall = time.perf_counter() - t0
prof = ctx.profile_report()
ctx.profile_enable(False)
print(f"wall {wall:.4f} s, device kernels {sum(v[1] for v in prof.values()):.2f} ms in {sum(v[0] for v in prof.values())} launches")
pr = cProfile.Profile()
pr.enable()
once()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
