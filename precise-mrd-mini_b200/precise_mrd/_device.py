"""Shared plumbing of the stage functions: the process-wide native context and the rule that
turns the caller's ``np.random.Generator`` into a Philox key."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _native as nv


def stage_context(config, rng: Optional[np.random.Generator]) -> nv.Context:
    """Context keyed for ONE stage call.

    The reference threads one PCG64 stream through every stage, so each stage's output depends
    on how many words the earlier stages happened to consume.  Here a stage
    takes exactly ONE 63-bit draw from the caller's generator and uses it -- mixed with
    ``config.seed`` -- as the Philox key; every draw inside the stage is then addressed by
    (cell, family, read).  Consequences: same generator state => same output (the reference's
    determinism contract); differently seeded generators (``default_rng(seed + rep*1000 +
    int(af*1e6))`` in ``eval/lod.py:136``) => different replicates; and no stage can perturb
    another by consuming a data-dependent number of words."""
    seed = int(getattr(config, "seed", 0) or 0)
    salt = int(rng.integers(0, 2**63 - 1)) if rng is not None else 0
    key = (seed * 0x9E3779B97F4A7C15 + salt * 0xD1342543DE82EF95 + 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF
    ctx = nv.default_context()
    ctx.set_seed(key)
    return ctx
