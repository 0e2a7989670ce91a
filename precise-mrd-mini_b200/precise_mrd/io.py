"""G-B read-level I/O -- drop-in for ``src/precise_mrd/io.py:17-147`` (``TargetSite``, ``Read``,
``SyntheticReadGenerator``).

``generate_reads_for_site`` is the reference's per-read Python loop (4.8e4 reads/s); here the
draws are the device read generator (kernels ``reads_size_kernel`` / ``reads_gen_kernel`` behind
``pmrd_simulate_reads``): NegBinom(r=2)+1 family sizes, 12-mer UMIs, per-family variant flag with
contamination flip, per-read allele (0.9 / 0.95 concordance), ``int(N(30,5))`` quality clipped to
[10,40], strand and read position.  ``generate_read_arrays`` returns the packed 12-byte records
(what ``UMIProcessor`` and ``pmrd_collapse_reads`` consume) without materialising Python objects;
``generate_reads_for_site`` unpacks them into ``Read`` objects for API compatibility.  Streams are
Philox keyed by (seed, site, family, read), not the reference's ``RandomState`` -- parity is
distributional."""
from __future__ import annotations

import zlib
from collections import namedtuple
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _native as nv
from .umi import BASES, unpack_bases


@dataclass
class TargetSite:
    """Represents a genomic target site for MRD analysis."""
    chrom: str
    pos: int
    ref: str
    alt: str
    context: str  # trinucleotide context
    gene: Optional[str] = None

    @property
    def key(self) -> str:
        return f"{self.chrom}:{self.pos}:{self.ref}>{self.alt}"

    def __hash__(self) -> int:
        return hash(self.key)


@dataclass
class Read:
    """Represents a sequencing read with UMI."""
    chrom: str
    pos: int
    ref: str
    alt: str
    umi: str
    quality: int
    strand: str  # '+' or '-'
    read_position: int  # position within read (for end-repair artifacts)
    family_id: Optional[str] = None


ReadPair = namedtuple("ReadPair", ["read1", "read2"])


def site_cell_id(site: TargetSite) -> int:
    """Philox cell id of a site.  The device generator derives the allele codes from the id
    (ref = id & 3, alt = (ref + 1 + (id >> 2) % 3) & 3, ``csrc/reads.cuh``), so the id is built to
    reproduce the site's ref / alt and is otherwise a stable hash of the site key."""
    r, a = BASES.index(site.ref), BASES.index(site.alt)
    if r == a:
        raise ValueError("site.alt must differ from site.ref")
    k = (a - r - 1) & 3                     # in {0, 1, 2}
    h = zlib.crc32(site.key.encode()) & 0x3FFFFFF
    return r | ((k + 3 * h) << 2)


class SyntheticReadGenerator:
    """Generates synthetic reads for simulation purposes (on the device)."""

    def __init__(self, seed: int = 42):
        self.seed = int(seed)
        self.rng = np.random.default_rng(seed)          # host stream: target-site generation only
        self._calls = 0

    # ------------------------------------------------------------------ device generation
    def generate_read_arrays(self, site: TargetSite, n_umi_families: int, allele_fraction: float = 0.0,
                             contamination_rate: float = 0.0, mean_family_size: float = 5.0,
                             max_family_size: int = 65535) -> Tuple[np.ndarray, np.ndarray]:
        """(key u64 = 0:32 | umi:32, payload u32) of every read of the site, family-major -- the
        record layout of ``pmrd_collapse_reads`` (``include/pmrd.h``)."""
        ctx = nv.default_context()
        # successive calls of one generator draw fresh streams, like the reference's stateful RandomState
        ctx.set_seed((self.seed << 20) ^ self._calls)
        self._calls += 1
        sp = nv.ReadSimParams(mean_family_size=float(mean_family_size), family_model=0, max_family_size=int(max_family_size),
                              contamination_rate=float(contamination_rate), hop_rate=0.0, collision_rate=0.0, shuffle=0,
                              chunk_log2=0)
        return ctx.simulate_reads([site_cell_id(site)], [float(allele_fraction)], [int(n_umi_families)], sp)

    def generate_reads_for_site(self, site: TargetSite, n_umi_families: int, allele_fraction: float = 0.0,
                                contamination_rate: float = 0.0, mean_family_size: float = 5.0) -> List[Read]:
        """``io.py:79-123``."""
        keys, payload = self.generate_read_arrays(site, n_umi_families, allele_fraction, contamination_rate, mean_family_size)
        umi = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
        fam = np.concatenate([[0], np.cumsum(umi[1:] != umi[:-1])]) if len(umi) else np.zeros(0, np.int64)
        umi_str = {u: unpack_bases(u, 12) for u in np.unique(umi)}
        allele, q = payload & 3, (payload >> 2) & 63
        strand, rpos = (payload >> 8) & 1, (payload >> 9) & 0xFF
        return [Read(chrom=site.chrom, pos=site.pos, ref=site.ref, alt=BASES[int(allele[i])], umi=umi_str[int(umi[i])],
                     quality=int(q[i]), strand="+" if strand[i] else "-", read_position=int(rpos[i]),
                     family_id=f"fam_{int(fam[i]):06d}") for i in range(len(keys))]

    # ------------------------------------------------------------------ host helpers (io.py:59-77,125-147)
    def generate_umi(self, length: int = 12) -> str:
        return "".join(self.rng.choice(list(BASES), size=length))

    def generate_quality_score(self, mean_quality: int = 30) -> int:
        return max(10, min(40, int(self.rng.normal(mean_quality, 5))))

    def simulate_family_sizes(self, n_families: int, mean_size: float = 5.0) -> List[int]:
        r = 2.0
        return (self.rng.negative_binomial(r, r / (r + mean_size), size=n_families) + 1).tolist()

    def generate_target_sites(self, n_sites: int = 10) -> List[TargetSite]:
        contexts = ["ACG", "CCG", "GCG", "TCG", "ACA", "CCA", "GCA", "TCA"]
        sites = []
        for i in range(n_sites):
            ref = str(self.rng.choice(list(BASES)))
            sites.append(TargetSite(chrom=f"chr{int(self.rng.integers(1, 23))}", pos=int(self.rng.integers(1000000, 50000000)),
                                    ref=ref, alt=str(self.rng.choice([b for b in BASES if b != ref])),
                                    context=str(self.rng.choice(contexts)), gene=f"GENE_{i + 1}"))
        return sites
