"""``B200StatisticalVariantCaller`` -- the statistical caller of ``models/statistical.py:17-92``
behind the ``VariantCaller`` interface, with the per-sample loop replaced by ONE ``pmrd_call``:
K7 (segmented counts + fp64 means), K8 (two-sided Poisson / binomial tails) and K9
(Benjamini-Hochberg) on the device.  ``get_variant_caller`` (``precise_mrd/call.py:21-34``)
returns it for the statistical branch."""
from __future__ import annotations

import numpy as np
import pandas as pd

from .. import _native as nv
from ..config import section
from .base import VariantCaller


class B200StatisticalVariantCaller(VariantCaller):
    """Variant caller using statistical tests (Poisson, Binomial) on the GPU."""

    def predict(self, collapsed_df: pd.DataFrame, error_model_df: pd.DataFrame) -> pd.DataFrame:
        st = section(self.config, "stats")
        test_type, alpha, fdr_method = st["test_type"], float(st["alpha"]), st["fdr_method"]
        if test_type not in nv.TEST_TYPES:
            raise ValueError(f"Unknown test type: {test_type}")                     # call.py:187
        if len(collapsed_df) == 0:
            return pd.DataFrame()                                                   # call.py:214-215
        # call.py:161: samples in first-appearance order; rows of a sample need not be contiguous
        codes, uniques = pd.factorize(collapsed_df["sample_id"], sort=False)
        is_variant = collapsed_df["is_variant"].values.astype(np.uint8)
        quality = collapsed_df["quality_score"].values.astype(np.float64)
        agreement = collapsed_df["consensus_agreement"].values.astype(np.float64)
        af = collapsed_df["allele_fraction"].values
        if (np.diff(codes) < 0).any():
            order = np.argsort(codes, kind="stable")
            codes, is_variant, quality, agreement, af = codes[order], is_variant[order], quality[order], agreement[order], af[order]
        seg = np.concatenate([[0], np.flatnonzero(np.diff(codes)) + 1, [len(codes)]]).astype(np.int64)
        mean_error_rate = float(error_model_df["error_rate"].mean())                # call.py:178
        out = nv.default_context().call(seg, is_variant, quality, agreement, mean_error_rate, nv.TEST_TYPES[test_type], alpha)
        df = pd.DataFrame({
            "sample_id": np.asarray(uniques),
            "allele_fraction": af[seg[:-1]],
            "n_variants": out["n_variants"],
            "n_total": out["n_total"],
            "variant_fraction": out["variant_fraction"],
            "expected_variants": out["expected_variants"],
            "fold_change": out["fold_change"],
            "p_value": out["p_value"],
            "mean_quality": out["mean_quality"],
            "mean_consensus": out["mean_consensus"],
            "test_type": test_type,
            "config_hash": self.config.config_hash(),
            "p_adjusted": out["p_adjusted"],
            "significant": out["significant"],
            "alpha": alpha,
            "fdr_method": fdr_method,
        })
        # every eval-* consumer indexes this column; no reference producer emits it
        df["variant_call"] = df["significant"]
        return df


StatisticalVariantCaller = B200StatisticalVariantCaller   # the reference's class name
