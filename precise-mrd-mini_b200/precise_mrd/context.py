"""Trinucleotide contexts -- the part of ``src/precise_mrd/context.py`` on the read-level path
(SURVEY.md §8a-a16): pyrimidine-centred normalisation (``:36-54``), per-context mutation rates
from a consensus table (``:72-116``, vectorised) and the rate lookup with its 1e-6 / 1e-4
defaults (``:118-127``).  Host-side string plumbing: the rates it yields are what
``StatisticalTester.test_multiple_variants`` hands to the device tail kernel."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import pandas as pd

_COMPLEMENT = {"A": "T", "T": "A", "C": "G", "G": "C"}


class ContextAnalyzer:
    """Analyze trinucleotide contexts and mutation patterns."""

    def __init__(self):
        self.context_error_rates: Dict[str, Dict[str, float]] = {}

    @staticmethod
    def get_complement(base: str) -> str:
        return _COMPLEMENT.get(base, "N")

    @staticmethod
    def reverse_complement(sequence: str) -> str:
        return "".join(_COMPLEMENT.get(b, "N") for b in reversed(sequence))

    @staticmethod
    def normalize_context(context: str, ref: str, alt: str) -> Tuple[str, str, str]:
        """Canonical form: C or T as the central base; purine-centred contexts are reverse-
        complemented together with ref and alt."""
        if len(context) != 3:
            raise ValueError("Context must be exactly 3 bases")
        if context[1] in ("A", "G"):
            return (ContextAnalyzer.reverse_complement(context), ContextAnalyzer.get_complement(ref),
                    ContextAnalyzer.get_complement(alt))
        return context, ref, alt

    @staticmethod
    def get_mutation_type(ref: str, alt: str) -> str:
        return "transition" if (ref, alt) in {("A", "G"), ("G", "A"), ("C", "T"), ("T", "C")} else "transversion"

    def extract_context_from_sequence(self, sequence: str, position: int) -> Optional[str]:
        if position < 1 or position >= len(sequence) - 1:
            return None
        return sequence[position - 1:position + 2].upper()

    def estimate_context_error_rates(self, consensus_data: pd.DataFrame, negative_control_sites=None) -> Dict[str, Dict[str, float]]:
        """``context.py:72-116``.  Reference rows (allele == ref) add their count to the RAW
        context's total, mutated rows to the NORMALISED context's total and mutation count -- the
        reference's own asymmetry, kept; rows without a ``context`` column count as 'NNN'."""
        if len(consensus_data) == 0:
            self.context_error_rates = {}
            return {}
        ctx = consensus_data["context"].astype(str) if "context" in consensus_data else pd.Series("NNN", index=consensus_data.index)
        ref, alt = consensus_data["ref"].astype(str), consensus_data["allele"].astype(str)
        count = consensus_data["consensus_count"]
        totals: Dict[str, float] = {}
        muts: Dict[str, Dict[str, float]] = {}
        for c, r, a, n in zip(ctx, ref, alt, count):
            if a == r:
                totals[c] = totals.get(c, 0) + n
                continue
            nc, nr, na = self.normalize_context(c, r, a)
            muts.setdefault(nc, {})
            muts[nc][f"{nr}>{na}"] = muts[nc].get(f"{nr}>{na}", 0) + n
            totals[nc] = totals.get(nc, 0) + n
        rates = {c: {k: v / totals[c] for k, v in m.items()} for c, m in muts.items() if totals.get(c, 0) != 0}
        self.context_error_rates = rates
        return rates

    def get_expected_error_rate(self, context: str, ref: str, alt: str) -> float:
        """``context.py:118-127``: known context -> its mutation's rate (1e-6 if unseen); else 1e-4."""
        nc, nr, na = self.normalize_context(context, ref, alt)
        if nc in self.context_error_rates:
            return self.context_error_rates[nc].get(f"{nr}>{na}", 1e-6)
        return 1e-4
