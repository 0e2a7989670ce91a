#!/usr/bin/env python
"""Assemble the runnable pieces of the reference under ``oracle/_ref/`` (git-ignored).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this file or anything
under ``oracle/``.  The reference (``/root/reference``) is a pile of six overlapping
generations none of which imports as shipped.  This script builds
*import shims* -- directories of **symlinks** into ``/root/reference`` plus a few
stub modules for symbols/third-party packages that do not exist -- so that the
reference's own code can be executed in this container to

  * validate the restatement in ``oracle/port.py`` / ``oracle/pmrd_oracle.c``;
  * generate golden vectors (``oracle/make_golden.py`` -> ``tests/golden/``).

No reference source is copied: every module is a symlink, so ``oracle/_ref`` is useless
on the GPU box (where ``/root/reference`` does not exist) and nothing that runs there
may touch it.

Layouts produced:

  oracle/_ref/gd_v1/precise_mrd/   G-D v1: the ``* 2.py`` snapshot under un-suffixed names
  oracle/_ref/gd_v2/precise_mrd/   G-D v2: ``src/precise_mrd/*.py`` + stubs for the four
                                   missing ``performance`` symbols and the two missing
                                   ``advanced_stats`` symbols
  oracle/_ref/ga/mrd/              G-A read-level package (precise-mrd-mini/src/mrd)
  oracle/_ref/gb/precise_mrd_gb/   G-B read-level modules io/umi/context/errors/stats/filters/qc + advanced_stats/config (validation stubbed)
  oracle/_ref/stubs/               matplotlib / seaborn stand-ins (PNG writers only); statsmodels.multipletests
                                   stand-in backed by oracle/port.py
"""
from __future__ import annotations

import os
import shutil
import sys
from pathlib import Path

REF = Path(os.environ.get("PMRD_REFERENCE", "/root/reference"))
OUT = Path(__file__).resolve().parent / "_ref"


def _link(src: Path, dst: Path) -> None:
    dst.parent.mkdir(parents=True, exist_ok=True)
    if dst.is_symlink() or dst.exists():
        dst.unlink()
    dst.symlink_to(src)


def _write(dst: Path, text: str) -> None:
    dst.parent.mkdir(parents=True, exist_ok=True)
    dst.write_text(text)


def build_gd_v1() -> None:
    src = REF / "src" / "precise_mrd"
    pkg = OUT / "gd_v1" / "precise_mrd"
    for p in src.glob("* 2.py"):
        _link(p, pkg / p.name.replace(" 2.py", ".py"))
    for sub in ("determinism_utils", "eval", "sim"):
        for p in (src / sub).glob("* 2.py"):
            _link(p, pkg / sub / p.name.replace(" 2.py", ".py"))
    # determinism.py has no " 2" twin but v1's package does not import it.


def build_gd_v2() -> None:
    src = REF / "src" / "precise_mrd"
    pkg = OUT / "gd_v2" / "precise_mrd"
    for name in ("simulate", "collapse", "error_model", "call", "config", "utils", "fastq"):
        _link(src / f"{name}.py", pkg / f"{name}.py")
    _write(pkg / "__init__.py", "# shim: the reference's own __init__ fails to import (SURVEY.md §0.2)\n")
    # the four symbols simulate.py:10 imports that src/precise_mrd/performance.py lacks
    _write(
        pkg / "performance.py",
        "class CacheStrategy:\n    HYBRID = 'hybrid'\n\n"
        "class IntelligentCache:\n"
        "    def __init__(self, *a, **k): pass\n"
        "    def get(self, key): return None\n"
        "    def put(self, *a, **k): pass\n\n"
        "class ChunkedProcessor:\n"
        "    def __init__(self, *a, **k): pass\n\n"
        "def profile_function(f):\n    return f\n",
    )
    _write(
        pkg / "advanced_stats.py",
        "class BayesianErrorModel:\n    def __init__(self, *a, **k): raise NotImplementedError\n\n"
        "class MLVariantCaller:\n    def __init__(self, *a, **k): raise NotImplementedError\n\n"
        "class AdvancedConfidenceIntervals:\n    def __init__(self, *a, **k): raise NotImplementedError\n",
    )


def build_ga() -> None:
    src = REF / "precise-mrd-mini" / "src" / "mrd"
    pkg = OUT / "ga" / "mrd"
    for p in src.glob("*.py"):
        _link(p, pkg / p.name)


def build_gb() -> None:
    src = REF / "src" / "precise_mrd"
    pkg = OUT / "gb" / "precise_mrd_gb"
    for name in ("io", "umi", "context", "errors", "stats", "filters", "qc", "advanced_stats", "config"):
        _link(src / f"{name}.py", pkg / f"{name}.py")
    _write(pkg / "__init__.py", "")
    # umi.py:16 imports a validate_umi_output that exists nowhere in the tree
    _write(pkg / "validation.py", "def validate_umi_output(df):\n    return True\n")


def build_stubs() -> None:
    st = OUT / "stubs"
    _write(
        st / "matplotlib" / "__init__.py",
        "def use(*a, **k): pass\n",
    )
    _write(
        st / "matplotlib" / "pyplot.py",
        "class _Any:\n"
        "    def __getattr__(self, n): return _Any()\n"
        "    def __call__(self, *a, **k): return _Any()\n"
        "    def __iter__(self): return iter([_Any(), _Any()])\n"
        "    def __getitem__(self, i): return _Any()\n"
        "def __getattr__(n): return _Any()\n",
    )
    _write(st / "seaborn" / "__init__.py", "def __getattr__(n):\n    return lambda *a, **k: None\n")
    # statsmodels is not installed: stats.py:15 imports multipletests; the stand-in is the
    # restatement of its three methods in oracle/port.py
    _write(st / "statsmodels" / "__init__.py", "")
    _write(st / "statsmodels" / "stats" / "__init__.py", "")
    _write(
        st / "statsmodels" / "stats" / "multitest.py",
        "from oracle import port as _port\n\n"
        "def multipletests(pvals, alpha=0.05, method='hs', **k):\n"
        "    reject, corrected = _port.multipletests(pvals, alpha, method)\n"
        "    return reject, corrected, None, None\n",
    )


def main() -> int:
    if not REF.exists():
        print(f"[oracle/build_ref] {REF} not present (GPU box?) - nothing to do")
        return 0
    if OUT.exists():
        shutil.rmtree(OUT)
    build_gd_v1()
    build_gd_v2()
    build_ga()
    build_gb()
    build_stubs()
    print(f"[oracle/build_ref] assembled import shims under {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
