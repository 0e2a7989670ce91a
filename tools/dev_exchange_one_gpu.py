"""The exchange's cut-and-store kernel on ONE GPU (4 emulated ranks, the "peer" arrays are this GPU's own): for ncu.  Dev helper."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "precise-mrd-mini_b200"))
sys.path.insert(0, str(ROOT))
from precise_mrd import _native as nv  # noqa: E402
from bench import external_reads  # noqa: E402

log2w, n = 2, 50_000_000
world = 1 << log2w
ctx = nv.Context(device=0, seed=7)
dev = torch.device("cuda", 0)
k, p = external_reads(5, n, 400_000)
dk, dp = torch.from_numpy(k.view(np.int64)).to(dev), torch.from_numpy(p.view(np.int32)).to(dev)
rk = [torch.empty(n, dtype=torch.int64, device=dev) for _ in range(world)]
rp = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(world)]
tile_off = torch.empty(world * ((n + 4095) // 4096), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    c = ctx.owner_tiles_dev(dk.data_ptr(), n, log2w, tile_off.data_ptr())
    ctx.owner_scatter_peer_dev(dk.data_ptr(), dp.data_ptr(), n, log2w, tile_off.data_ptr(), [t.data_ptr() for t in rk], [t.data_ptr() for t in rp], [0] * world)
    ctx.sync()
print(c)
