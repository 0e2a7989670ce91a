"""Device time of the LoD detection-curve fits (K11) on the reference-generated curves of tests/golden/lod_fit.npz:
5 curves x (1 + 200) Levenberg-Marquardt refits in one launch.  Dev helper; prints the fits so variants can be diffed."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
from precise_mrd import _native as nv  # noqa: E402

g = np.load(ROOT / "tests" / "golden" / "lod_fit.npz")
names = ["mid", "steep", "shallow", "zeros", "ones"]
ctx = nv.Context(device=0, seed=7)
hits = [g[f"{n}_hit"] for n in names]
ids = [(12345 << 8) | i for i in range(len(names))]
ctx.fit_lod_batch(list(g["af"]), hits, 0.95, 200, ids)
ctx.profile_enable(True)
for _ in range(5):
    fits = ctx.fit_lod_batch(list(g["af"]), hits, 0.95, 200, ids)
prof = ctx.profile_report()
print({k: (v[0], round(v[1] / v[0], 3)) for k, v in prof.items()})
for n, f in zip(names, fits):
    print(n, [float(x).hex() for x in f], float(g[f"{n}_lod"]))
