"""SHA-256 manifests (``determinism_utils/hash_artifacts.py:11-71``): ``"<sha256>  <path>\\n"``."""
from __future__ import annotations

import hashlib
import pathlib
from typing import Sequence


def hash_file(path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def hash_dir(directory, pattern: str = "*") -> dict[str, str]:
    d = pathlib.Path(directory)
    return {str(p.relative_to(d)): hash_file(p) for p in sorted(d.rglob(pattern)) if p.is_file()}


def write_manifest(paths: Sequence, out_manifest="reports/hash_manifest.txt") -> None:
    mp = pathlib.Path(out_manifest)
    mp.parent.mkdir(parents=True, exist_ok=True)
    with open(mp, "w") as f:
        for path in paths:
            p = pathlib.Path(path)
            f.write(f"{hash_file(p)}  {path}\n" if p.exists() else f"MISSING  {path}\n")


def verify_manifest(manifest_path) -> dict[str, bool]:
    out = {}
    for line in pathlib.Path(manifest_path).read_text().splitlines():
        parts = line.strip().split("  ", 1)
        if len(parts) == 2:
            p = pathlib.Path(parts[1])
            out[parts[1]] = p.exists() and hash_file(p) == parts[0]
    return out
