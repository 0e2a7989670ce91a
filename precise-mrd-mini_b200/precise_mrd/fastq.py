"""FASTQ ingest -- drop-in for ``src/precise_mrd/fastq.py`` (``FASTQReader``,
``process_fastq_to_dataframe``, ``detect_umi_format``); SURVEY.md §8(f)-1.

The file's bytes are handed to the device once (gzip is inflated on the host first).  Kernel K13
(``csrc/fastq.cu``) finds the lines, applies the reference's three header patterns, sums the
Phred+33 scores and, after a stable radix sort on the packed UMI, builds the per-UMI family table
-- the work the reference does with ``readline``, ``re.search`` and a dict of lists.  The host
slices strings out of the buffer only for the rows it returns.  The table feeds
``collapse_umis(..., is_fastq_data=True)`` (``collapse.py:46-64``).

UMIs longer than 12 letters (only the unbounded ``UMI:([A-Z]+)`` pattern can produce them) do
not fit the packed 64-bit key; the host numbers those few strings and they join the same device
grouping under keys tagged with length nibble 15.  A lone ``\\r`` is not treated as a line end."""
from __future__ import annotations

import gzip
import re
from pathlib import Path
from typing import Any, Dict, Iterator, List, Optional

import numpy as np
import pandas as pd

from . import _native as nv

FQ_END, FQ_MALFORMED, FQ_NO_UMI, FQ_OK, FQ_LONG_UMI = 0, 1, 2, 3, 4


class FASTQReader:
    """FASTQ file reader with UMI extraction (device scan behind the reference's interface)."""

    def __init__(self, file_path):
        self.file_path = Path(file_path)
        self.is_gzipped = self.file_path.suffix == ".gz"
        self.umi_patterns = [r"UMI:([A-Z]+)", r"_([A-Z]{8,12})_", r":([A-Z]{8,12})$"]   # fastq.py:28-32
        self._bytes: Optional[bytes] = None
        self._rows: Optional[Dict[str, np.ndarray]] = None

    # ------------------------------------------------------------------ scalar helpers (API compatibility)
    def extract_umi(self, read_header: str) -> Optional[str]:
        for pattern in self.umi_patterns:
            m = re.search(pattern, read_header)
            if m:
                return m.group(1)
        return None

    def parse_quality_string(self, quality: str) -> List[int]:
        return [ord(c) - 33 for c in quality]

    # ------------------------------------------------------------------ device scan
    def _scan(self):
        if self._rows is None:
            opener = gzip.open if self.is_gzipped else open
            with opener(self.file_path, "rb") as f:
                self._bytes = f.read()
            self._rows = nv.default_context().fastq_scan(self._bytes)
        return self._bytes, self._rows

    def valid_rows(self, max_reads: Optional[int] = None) -> np.ndarray:
        """Row indices read_fastq would yield: records up to the first empty header, malformed ones
        skipped, at most ``max_reads`` of them (``max_reads`` of 0 / None: all -- fastq.py:79)."""
        _, rows = self._scan()
        st = rows["status"]
        end = np.flatnonzero(st == FQ_END)
        stop = int(end[0]) if len(end) else len(st)
        idx = np.flatnonzero(st[:stop] != FQ_MALFORMED)
        return idx[:max_reads] if max_reads else idx

    def umi_keys(self, rows_idx: np.ndarray) -> np.ndarray:
        """64-bit grouping key of each given row (all must carry a UMI): the packed letters, or for
        a UMI of more than 12 letters ``(id << 4) | 15`` with ids numbering the distinct long strings."""
        _, rows = self._scan()
        key = rows["umi_key"][rows_idx].copy()
        long_rows = np.flatnonzero(rows["status"][rows_idx] == FQ_LONG_UMI)
        if len(long_rows):
            ids, _ = pd.factorize(np.array([self._text(int(rows["umi_off"][r]), int(rows["umi_len"][r])) for r in rows_idx[long_rows]],
                                           dtype=object))
            key[long_rows] = (ids.astype(np.uint64) << np.uint64(4)) | np.uint64(15)
        return key

    def _text(self, off: int, length: int) -> str:
        return self._bytes[off:off + length].decode("utf-8", "replace")

    def read_fastq(self, max_reads: Optional[int] = None) -> Iterator[Dict[str, Any]]:
        """fastq.py:66-110."""
        buf, rows = self._scan()
        for r in self.valid_rows(max_reads):
            q = np.frombuffer(buf, np.uint8, int(rows["qual_len"][r]), int(rows["qual_off"][r])).astype(np.int64) - 33
            yield {
                "read_id": self._text(int(rows["hdr_off"][r]) + 1, int(rows["hdr_len"][r]) - 1),
                "sequence": self._text(int(rows["seq_off"][r]), int(rows["seq_len"][r])),
                "umi": self._text(int(rows["umi_off"][r]), int(rows["umi_len"][r])) if rows["status"][r] >= FQ_OK else None,
                "quality_scores": q.tolist(),
                "mean_quality": rows["qual_sum"][r] / rows["qual_len"][r],
                "sequence_length": int(rows["seq_len"][r]),
            }

    def validate_fastq(self) -> Dict[str, Any]:
        """fastq.py:112-155: statistics of the first 1000 reads (same running-average recurrences)."""
        stats = {"total_reads": 0, "valid_reads": 0, "umi_reads": 0, "mean_read_length": 0.0, "mean_quality": 0.0,
                 "is_valid": True, "errors": []}
        try:
            _, rows = self._scan()
            idx = self.valid_rows(1000)
            for i, r in enumerate(idx):
                stats["total_reads"] = i + 1
                stats["valid_reads"] += 1
                if rows["status"][r] >= FQ_OK:
                    stats["umi_reads"] += 1
                length, mq = int(rows["seq_len"][r]), rows["qual_sum"][r] / rows["qual_len"][r]
                if i == 0:
                    stats["mean_read_length"], stats["mean_quality"] = length, mq
                else:
                    stats["mean_read_length"] = (stats["mean_read_length"] * i + length) / (i + 1)
                    stats["mean_quality"] = (stats["mean_quality"] * i + mq) / (i + 1)
        except Exception as e:  # fastq.py:150-152
            stats["is_valid"] = False
            stats["errors"].append(str(e))
        return stats


def process_fastq_to_dataframe(fastq_path, config, rng: np.random.Generator, output_path: Optional[str] = None,
                               max_reads: Optional[int] = None) -> pd.DataFrame:
    """fastq.py:158-236: one row per UMI family, families in first-appearance order."""
    reader = FASTQReader(fastq_path)
    validation = reader.validate_fastq()
    if not validation["is_valid"]:
        raise ValueError(f"Invalid FASTQ file: {validation['errors']}")
    print(f"Processing FASTQ file: {fastq_path}")
    print(f"Validation: {validation['total_reads']} reads, {validation['umi_reads']} with UMIs, "
          f"mean length {validation['mean_read_length']:.1f}bp")
    buf, rows = reader._scan()
    idx = reader.valid_rows(max_reads)
    kept = idx[rows["status"][idx] >= FQ_OK]                       # reads without a UMI are dropped (:189-190)
    cols = ["sample_id", "umi", "family_size", "consensus_sequence", "quality_scores", "mean_quality", "consensus_agreement",
            "passes_quality", "passes_consensus", "config_hash"]
    if len(kept) == 0:
        df = pd.DataFrame()
    else:
        key_of_read = reader.umi_keys(kept)
        fam = nv.default_context().fastq_families(key_of_read, rows["qual_sum"][kept], rows["qual_len"][kept])
        first, best = kept[fam["first_read"]], kept[fam["best_read"]]
        mean_q = fam["qual_sum"] / fam["n_bases"]                  # np.mean over every base of the family (:222)
        # per-family list of all quality scores, reads in file order (:210)
        fam_of_key = {int(k): i for i, k in enumerate(fam["umi_key"])}
        qlists: List[List[int]] = [[] for _ in range(len(first))]
        for r, k in zip(kept, key_of_read):
            qlists[fam_of_key[int(k)]].extend(
                (np.frombuffer(buf, np.uint8, int(rows["qual_len"][r]), int(rows["qual_off"][r])).astype(np.int64) - 33).tolist())
        df = pd.DataFrame({
            "sample_id": np.arange(len(first)),
            "umi": [reader._text(int(rows["umi_off"][r]), int(rows["umi_len"][r])) for r in first],
            "family_size": fam["family_size"],
            "consensus_sequence": [reader._text(int(rows["seq_off"][r]), int(rows["seq_len"][r])) for r in best],
            "quality_scores": qlists,
            "mean_quality": mean_q,
            "consensus_agreement": 1.0,
            "passes_quality": mean_q >= 20,
            "passes_consensus": True,
            "config_hash": config.config_hash(),
        }, columns=cols)
    if output_path:
        df.to_parquet(output_path, index=False)
    return df


def detect_umi_format(fastq_path, sample_size: int = 100) -> str:
    """fastq.py:239-280: which header convention the first ``sample_size`` reads use."""
    reader = FASTQReader(fastq_path)
    formats = {"UMI:SEQUENCE": 0, "_UMI_": 0, ":UMI": 0, "none": 0}
    for read in reader.read_fastq(max_reads=sample_size):
        umi, header = read["umi"], read["read_id"]
        if not umi:
            formats["none"] += 1
        elif "UMI:" in header:
            formats["UMI:SEQUENCE"] += 1
        elif "_" in header and len(umi) >= 8:
            formats["_UMI_"] += 1
        elif header.endswith(umi):
            formats[":UMI"] += 1
    if max(formats.values()) == 0:
        return "No UMI detected in sample"
    return max(formats.items(), key=lambda x: x[1])[0]
