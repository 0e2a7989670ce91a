"""Deterministic HTML report (role of ``src/precise_mrd/reporting.py:39-164``): a string template
whose bytes go into the hash manifest -- so no timestamps, no unordered iteration."""
from __future__ import annotations

import html
import json
from pathlib import Path
from typing import Any, Dict

import pandas as pd


def _fmt(v: Any) -> str:
    if isinstance(v, float):
        return f"{v:.6g}"
    return html.escape(str(v))


def render_report(calls_df: pd.DataFrame, metrics: Dict[str, Any], config: Dict[str, Any], run_context: Dict[str, Any],
                  output_path: str) -> str:
    rows = []
    cols = ["sample_id", "allele_fraction", "n_variants", "n_total", "p_value", "p_adjusted", "significant"]
    if len(calls_df):
        for _, r in calls_df[cols].head(50).iterrows():
            rows.append("<tr>" + "".join(f"<td>{_fmt(r[c])}</td>" for c in cols) + "</tr>")
    n_sig = int(calls_df["significant"].sum()) if len(calls_df) else 0
    body = f"""<!DOCTYPE html>
<html><head><meta charset="utf-8"><title>Precise MRD report</title></head><body>
<h1>Precise MRD &mdash; run {html.escape(str(config.get('run_id', '')))}</h1>
<p>seed {run_context.get('seed')} &middot; config {run_context.get('config_hash')} &middot; engine precise_mrd-b200 (sm_100a)</p>
<h2>Summary</h2>
<ul>
<li>samples: {len(calls_df)}</li>
<li>significant calls: {n_sig}</li>
<li>ROC AUC: {_fmt(metrics.get('roc_auc', 0.0))}</li>
<li>average precision: {_fmt(metrics.get('average_precision', 0.0))}</li>
<li>detected / total cases: {metrics.get('detected_cases', 0)} / {metrics.get('total_cases', 0)}</li>
</ul>
<h2>Calls (first 50)</h2>
<table border="1"><tr>{''.join(f'<th>{c}</th>' for c in cols)}</tr>
{chr(10).join(rows)}
</table>
<h2>Configuration</h2>
<pre>{html.escape(json.dumps(config, indent=2, sort_keys=True))}</pre>
</body></html>
"""
    Path(output_path).parent.mkdir(parents=True, exist_ok=True)
    Path(output_path).write_text(body, encoding="utf-8")
    return body


def write_png_heatmap(matrix, path, cell: int = 24) -> None:
    """Minimal dependency-free PNG writer (matplotlib is not a dependency of the hot path):
    one coloured block per matrix entry, blue (0) -> red (1)."""
    import struct
    import zlib

    import numpy as np

    m = np.atleast_2d(np.asarray(matrix, dtype=float))
    m = np.nan_to_num(np.clip(m, 0.0, 1.0))
    h, w = m.shape
    img = np.zeros((h * cell, w * cell, 3), np.uint8)
    for i in range(h):
        for j in range(w):
            v = m[i, j]
            img[i * cell:(i + 1) * cell - 1, j * cell:(j + 1) * cell - 1] = (int(255 * v), 64, int(255 * (1 - v)))
    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(img.shape[0]))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", img.shape[1], img.shape[0], 8, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b"")
    Path(path).write_bytes(png)
