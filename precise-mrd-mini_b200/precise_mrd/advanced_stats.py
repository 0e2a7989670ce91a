"""Feature table for the alternative (ML / Bayesian) callers.

Mirror of ``FeatureEngineer`` (reference ``src/precise_mrd/advanced_stats.py:21-108``).  The callers
themselves are out of scope; what the hot path owes them is their INPUT: the
feature matrix of the collapsed-family table, produced on the device (kernel K15,
``pmrd_feature_table``) with the reference's own IEEE operations, so the matrix is bit-equal to the
pandas expressions and a third-party caller behind ``models.VariantCaller`` trains on the same numbers."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import pandas as pd

from . import _native as nv

BASIC_FEATURES = ["family_size", "quality_score", "consensus_agreement", "allele_fraction"]


class FeatureEngineer:
    """advanced_stats.py:21-108 -- same constructor, methods, column names and quirks."""

    def __init__(self, config=None):
        self.config = config
        self.feature_names: List[str] = []

    def extract_features(self, collapsed_df: pd.DataFrame) -> Tuple[pd.DataFrame, List[str]]:
        present = [c for c in BASIC_FEATURES if c in collapsed_df.columns]
        n = len(collapsed_df)
        if len(present) == len(BASIC_FEATURES):
            # the full table: one device pass writes all six columns
            ctx = nv.default_context()
            X = ctx.feature_table(collapsed_df["family_size"].to_numpy(), collapsed_df["quality_score"].to_numpy(),
                                  collapsed_df["consensus_agreement"].to_numpy(), collapsed_df["allele_fraction"].to_numpy())
            self.feature_names = BASIC_FEATURES + ["family_quality_ratio", "confidence_af"]
            return pd.DataFrame(X, columns=self.feature_names), self.feature_names
        # partial tables (advanced_stats.py:44-64): only the columns present, the derived ones when both of their
        # inputs are; the NAMES are the first len(features) basic names, as in the reference (its quirk: a
        # missing column shifts the labels), plus the two derived names when any derived column exists
        cols = [collapsed_df[c].to_numpy().reshape(-1, 1) for c in present]
        n_basic = len(cols)
        if "family_size" in collapsed_df.columns and "quality_score" in collapsed_df.columns:
            cols.append((collapsed_df["family_size"] / (collapsed_df["quality_score"] + 1)).to_numpy().reshape(-1, 1))
        if "consensus_agreement" in collapsed_df.columns and "allele_fraction" in collapsed_df.columns:
            cols.append((collapsed_df["consensus_agreement"] * collapsed_df["allele_fraction"]).to_numpy().reshape(-1, 1))
        if cols:
            X = np.hstack(cols)
            self.feature_names = BASIC_FEATURES[: len(cols)]
            if len(cols) > len(BASIC_FEATURES):
                self.feature_names.extend(["family_quality_ratio", "confidence_af"])
        else:
            X = np.zeros((n, 1))
            self.feature_names = ["dummy_feature"]
        del n_basic
        return pd.DataFrame(X, columns=self.feature_names), self.feature_names

    def select_features(self, X: pd.DataFrame, y: np.ndarray, method: str = "mutual_info", k: int = 10):
        """advanced_stats.py:85-108 (scikit-learn's SelectKBest, as in the reference: host-side model code)."""
        from sklearn.feature_selection import SelectKBest, f_classif, mutual_info_classif

        if method == "f_classif":
            selector = SelectKBest(score_func=f_classif, k=min(k, X.shape[1]))
        elif method == "mutual_info":
            selector = SelectKBest(score_func=mutual_info_classif, k=min(k, X.shape[1]))
        else:
            raise ValueError(f"Unknown feature selection method: {method}")
        Xs = selector.fit_transform(X, y)
        names = X.columns[selector.get_support()].tolist()
        return pd.DataFrame(Xs, columns=names), names
