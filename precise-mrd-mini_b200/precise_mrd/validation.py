"""Artifact contract (``src/precise_mrd/validation.py:13-137``): required files, JSON-Schema
validation, CSV column checks, manifest comparison."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, Mapping

import jsonschema
import pandas as pd

SCHEMA_DIR = Path(__file__).resolve().parent / "assets" / "schemas"
BASE_REQUIRED_ARTIFACTS = ["metrics.json", "run_context.json", "hash_manifest.txt"]
JSON_SCHEMA_FILES: Mapping[str, str] = {
    "metrics.json": "metrics.schema.json",
    "run_context.json": "run_context.schema.json",
    "lob.json": "lob.schema.json",
    "contam_sensitivity.json": "contamination.schema.json",
    "power_by_stratum.json": "power.schema.json",
    "calibration_by_bin.json": "calibration.schema.json",
}
CSV_COLUMNS = {
    "lod_table.csv": ["depth", "lod_af", "lod_ci_lower", "lod_ci_upper"],
    "loq_table.csv": ["depth", "loq_af_cv", "loq_af_abs_error"],
}


def load_schema(name: str) -> Dict[str, Any]:
    return json.loads((SCHEMA_DIR / name).read_text(encoding="utf-8"))


def validate_json(path: Path, schema_name: str) -> None:
    jsonschema.validate(json.loads(Path(path).read_text(encoding="utf-8")), load_schema(schema_name))


def validate_artifacts(run_dir) -> None:
    run_dir = Path(run_dir)
    reports = run_dir if run_dir.name == "reports" else run_dir / "reports"
    if not reports.exists():
        raise AssertionError(f"Reports directory not found: {reports}")
    for required in BASE_REQUIRED_ARTIFACTS:
        cand = reports / required
        if not cand.exists():
            raise AssertionError(f"Required artifact missing: {cand}")
        if cand.suffix == ".txt" and not cand.read_text(encoding="utf-8").strip():
            raise AssertionError(f"Artifact is empty: {cand}")
    for filename, schema_name in JSON_SCHEMA_FILES.items():
        cand = reports / filename
        if cand.exists():
            validate_json(cand, schema_name)
    for filename, cols in CSV_COLUMNS.items():
        cand = reports / filename
        if cand.exists():
            frame = pd.read_csv(cand)
            missing = [c for c in cols if c not in frame.columns]
            if missing:
                raise ValueError(f"Missing expected columns in {cand}: {missing}")
    heatmap = reports / "contam_heatmap.png"
    if heatmap.exists() and heatmap.stat().st_size == 0:
        raise AssertionError(f"Heatmap artifact is empty: {heatmap}")


def _read_manifest(path: Path) -> Dict[str, str]:
    entries: Dict[str, str] = {}
    for raw in Path(path).read_text(encoding="utf-8").splitlines():
        parts = raw.strip().split("  ", 1)
        if len(parts) == 2:
            entries[parts[1]] = parts[0]
    return entries


def assert_hashes_stable(previous_manifest, current_manifest) -> None:
    prev, curr = _read_manifest(Path(previous_manifest)), _read_manifest(Path(current_manifest))
    if prev != curr:
        details = {"missing": sorted(set(prev) - set(curr)), "unexpected": sorted(set(curr) - set(prev)),
                   "changed": sorted(k for k in set(prev) & set(curr) if prev[k] != curr[k])}
        raise AssertionError(f"Hash manifest mismatch detected: {json.dumps(details, indent=2)}")
