"""Read-level UMI processing -- the G-B surface (``src/precise_mrd/umi.py``: ``UMIProcessor``,
``UMIFamily``) and the G-A one (``precise-mrd-mini/src/mrd/umi.py``: ``collapse_reads``) over the
device collapse (kernels K4-K7, ``pmrd_collapse_reads``): the host only packs strings into the
12-byte read record and unpacks the family / locus tables."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import _native as nv

BASES = "ACGT"
_CODE = np.full(256, 255, np.uint8)
for _i, _b in enumerate(BASES):
    _CODE[ord(_b)] = _i


def pack_bases(seqs: Sequence[str]) -> np.ndarray:
    """2 bits per base, first base in the top bits (numeric order == lexicographic order)."""
    if len(seqs) == 0:
        return np.zeros(0, np.uint64)
    L = len(seqs[0])
    if L > 32 or any(len(s) != L for s in seqs):
        raise ValueError("sequences must share one length <= 32")
    arr = np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8).reshape(len(seqs), L)
    codes = _CODE[arr]
    if (codes == 255).any():
        raise ValueError("sequences must be over ACGT")
    out = np.zeros(len(seqs), np.uint64)
    for c in range(L):
        out = (out << np.uint64(2)) | codes[:, c].astype(np.uint64)
    return out


def unpack_bases(v: int, length: int) -> str:
    return "".join(BASES[(int(v) >> (2 * (length - 1 - i))) & 3] for i in range(length))


@dataclass
class UMIFamily:
    """One consensus family (``umi.py:19-47``), unpacked from the device family table."""
    umi: str
    site_key: str
    family_size: int
    consensus_allele: Optional[str] = None
    consensus_quality: Optional[float] = None
    n_alt_reads: int = 0


class UMIProcessor:
    """``src/precise_mrd/umi.py:50-285``.

    ``max_edit_distance == 0``: exact-UMI families (``umi.py:111-115``).  ``max_edit_distance > 0`` (the
    reference's default, ``configs/default.yaml:16``): greedy Hamming clustering.  The reference loop
    (``umi.py:117-144``) iterates ``list(set(...))`` of strings, so its result depends on the process's
    hash seed; the iteration order is DEFINED here as that of the reference's own
    deterministic form, ``umi_clustering.py:59-101`` (``cluster_umis_greedy``): unique UMIs by
    ``(-count, umi)``, the first unused one seeds a cluster and absorbs every unused UMI within the
    distance.  Families come back in the reference's order: sites by first appearance, then UMIs by
    first appearance (exact) / clusters in seed order (greedy)."""

    def __init__(self, min_family_size: int = 3, max_family_size: int = 1000, max_edit_distance: int = 1,
                 quality_threshold: int = 20, consensus_threshold: float = 0.6):
        if min_family_size < 1:
            raise ValueError("min_family_size must be at least 1")
        if max_family_size < min_family_size:
            raise ValueError("max_family_size must be >= min_family_size")
        if max_edit_distance < 0:
            raise ValueError("max_edit_distance must be non-negative")
        if quality_threshold < 0 or quality_threshold > 60:
            raise ValueError("quality_threshold must be between 0 and 60")
        if not 0.0 < consensus_threshold <= 1.0:
            raise ValueError("consensus_threshold must be between 0.0 and 1.0")
        self.min_family_size, self.max_family_size = min_family_size, max_family_size
        self.max_edit_distance, self.quality_threshold = max_edit_distance, quality_threshold
        self.consensus_threshold = consensus_threshold

    def edit_distance(self, umi1: str, umi2: str) -> int:
        """``umi.py:90-95`` / ``rust_ext/src/lib.rs:77-90`` (device: XOR + popcount on packed UMIs)."""
        if len(umi1) != len(umi2):
            return max(len(umi1), len(umi2))
        ctx = nv.default_context()
        return int(ctx.umi_hamming(pack_bases([umi1]).astype(np.uint32), pack_bases([umi2]).astype(np.uint32), len(umi1))[0])

    def pack_reads(self, reads) -> Tuple[np.ndarray, np.ndarray, List[str], int]:
        """reads: objects with chrom, pos, ref, alt (observed allele), umi, quality, strand,
        read_position -- ``src/precise_mrd/io.py:34-46``; or a DataFrame with those columns."""
        if isinstance(reads, pd.DataFrame):
            df = reads
        else:
            df = pd.DataFrame({k: [getattr(r, k) for r in reads] for k in
                               ("chrom", "pos", "ref", "alt", "umi", "quality", "strand", "read_position")})
        site = df["chrom"].astype(str) + ":" + df["pos"].astype(str) + ":" + df["ref"].astype(str)
        codes, sites = pd.factorize(site, sort=False)                 # dict insertion order (umi.py:97-103)
        umis = list(df["umi"].astype(str))
        umi_len = len(umis[0]) if umis else 12
        key = (codes.astype(np.uint64) << np.uint64(32)) | pack_bases(umis)
        allele = _CODE[np.frombuffer("".join(df["alt"].astype(str)).encode("ascii"), dtype=np.uint8)].astype(np.uint32)
        q = np.clip(df["quality"].values.astype(np.int64), 0, 63).astype(np.uint32)
        strand = (df["strand"].astype(str).values == "+").astype(np.uint32)
        rp = np.clip(df["read_position"].values.astype(np.int64), 0, 255).astype(np.uint32)
        payload = allele | (q << np.uint32(2)) | (strand << np.uint32(8)) | (rp << np.uint32(9))
        return key, payload.astype(np.uint32), list(sites), umi_len

    def collapse_packed(self, key: np.ndarray, payload: np.ndarray, umi_len: int = 12):
        """Device collapse of packed reads with this processor's thresholds; returns the family and group tables
        with the families in the reference's order (see the class docstring) and a ``first_read`` column."""
        greedy = self.max_edit_distance > 0
        params = nv.UmiParams(min_family_size=self.min_family_size, max_family_size=self.max_family_size,
                              quality_threshold=self.quality_threshold, consensus_threshold=self.consensus_threshold,
                              mode=nv.MODE_WEIGHTED, cluster=nv.CLUSTER_GREEDY if greedy else nv.CLUSTER_EXACT,
                              max_distance=self.max_edit_distance, umi_len=umi_len)
        fam, grp = nv.default_context().collapse_reads(key, payload, params)
        if not greedy and len(fam["key"]):
            # exact families arrive sorted by (site, packed UMI); the reference lists a site's UMIs in insertion
            # order (umi.py:111-115, dict order): reorder by each family's first read
            uniq, first = np.unique(key, return_index=True)
            assert len(uniq) == len(fam["key"])
            order = np.lexsort((first, fam["key"] >> np.uint64(32)))
            fam = {k: v[order] for k, v in fam.items()}
            fam["first_read"] = first[order]
        return fam, grp

    def process_reads(self, reads) -> List[UMIFamily]:
        """``umi.py:205-239``: group by site, group by UMI, call consensus, filter family size."""
        if reads is None or len(reads) == 0:
            return []
        key, payload, sites, umi_len = self.pack_reads(reads)
        fam, self._groups = self.collapse_packed(key, payload, umi_len)
        self._sites = sites
        out = []
        kept = (fam["flags"] & nv.FAM_KEPT) != 0
        for i in np.flatnonzero(kept):
            has = bool(fam["flags"][i] & nv.FAM_HAS_CONSENSUS)
            # filter_family_size (umi.py:190-199) caps an oversize family at max_family_size reads (which reads it
            # keeps is drawn from the global NumPy RNG there -- SURVEY.md App. A #10; the consensus was called before)
            out.append(UMIFamily(
                umi=unpack_bases(int(fam["key"][i]) & 0xFFFFFFFF, umi_len), site_key=sites[int(fam["key"][i]) >> 32],
                family_size=min(int(fam["size"][i]), self.max_family_size),
                consensus_allele=BASES[int(fam["consensus"][i])] if has else None,
                consensus_quality=(float(fam["qsum"][i]) / float(fam["n_best"][i])) if has else None,
                n_alt_reads=int(fam["n_alt"][i])))
        return out

    def get_consensus_counts(self, families: List[UMIFamily]) -> pd.DataFrame:
        """``umi.py:241-266``: consensus family counts per (site, allele)."""
        counts: Dict[Tuple[str, str], int] = {}
        for f in families:
            if f.consensus_allele is not None:
                counts[(f.site_key, f.consensus_allele)] = counts.get((f.site_key, f.consensus_allele), 0) + 1
        rows = []
        for (site_key, allele), c in counts.items():
            chrom, pos, ref = site_key.split(":")
            rows.append({"chrom": chrom, "pos": int(pos), "ref": ref, "allele": allele, "consensus_count": c, "site_key": site_key})
        return pd.DataFrame(rows)


def collapse_reads(reads: pd.DataFrame, min_family_size: int = 2, max_distance: int = 1) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """G-A ``mrd.umi.collapse_reads`` (``precise-mrd-mini/src/mrd/umi.py:57-115``): per
    (chrom, start, pos) group, first-fit Hamming clustering to the representative UMI, per-column
    majority consensus over the 10-mer read, per-locus alt / total / family counts."""
    if len(reads) == 0:
        return pd.DataFrame(), pd.DataFrame()
    gkeys = sorted(set(zip(reads["chrom"], reads["start"], reads["pos"])))          # pandas groupby order
    gid = {g: i for i, g in enumerate(gkeys)}
    g = np.array([gid[t] for t in zip(reads["chrom"], reads["start"], reads["pos"])], np.uint64)
    umis = list(reads["umi"].astype(str))
    seqs = list(reads["sequence"].astype(str))
    key = (g << np.uint64(32)) | pack_bases(umis)
    payload = pack_bases(seqs).astype(np.uint32) | (reads["true_alt"].values.astype(np.uint32) << np.uint32(31))
    params = nv.UmiParams(min_family_size=min_family_size, max_family_size=1 << 30, quality_threshold=0,
                          consensus_threshold=0.0, mode=nv.MODE_MAJORITY, cluster=nv.CLUSTER_FIRSTFIT,
                          max_distance=max_distance, umi_len=len(umis[0]))
    fam, grp = nv.default_context().collapse_reads(key, payload, params)
    kept = np.flatnonzero((fam["flags"] & nv.FAM_KEPT) != 0)
    patient = reads["patient_id"].iloc[0] if "patient_id" in reads else None
    fg = (fam["key"][kept] >> np.uint64(32)).astype(np.int64)
    family_df = pd.DataFrame({
        "patient_id": patient, "chrom": [gkeys[i][0] for i in fg], "pos": [gkeys[i][2] for i in fg],
        "umi_representative": [unpack_bases(int(k) & 0xFFFFFFFF, len(umis[0])) for k in fam["key"][kept]],
        "start": [gkeys[i][1] for i in fg], "consensus": [unpack_bases(int(c), len(seqs[0])) for c in fam["consensus"][kept]],
        "family_size": fam["size"][kept].astype(np.int64), "alt_count": fam["n_alt"][kept].astype(np.int64),
    })
    loci: Dict[Tuple[str, int], Dict] = {}
    for _, r in family_df.iterrows():                                                # first-appearance order (mrd/umi.py:86)
        e = loci.setdefault((r.chrom, r.pos), {"patient_id": patient, "chrom": r.chrom, "pos": r.pos, "alt_count": 0,
                                               "total_count": 0, "family_sizes": []})
        e["alt_count"] += int(r.alt_count)
        e["total_count"] += int(r.family_size)
        e["family_sizes"].append(int(r.family_size))
    locus_df = pd.DataFrame([{**e, "family_count": len(e["family_sizes"]),
                              "mean_family_size": float(np.mean(e["family_sizes"])) if e["family_sizes"] else 0.0}
                             for e in loci.values()])
    return family_df, locus_df
