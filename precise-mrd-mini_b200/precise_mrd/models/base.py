"""``VariantCaller`` -- the abstract plugin interface of the reference (``models/base.py:15-50``):
``__init__(config)``, ``train(collapsed_df, rng) -> dict`` (a no-op for statistical callers) and
``predict(collapsed_df, error_model_df) -> DataFrame``."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict

import numpy as np
import pandas as pd


class VariantCaller(ABC):
    def __init__(self, config):
        self.config = config
        self.model = None

    def train(self, collapsed_df: pd.DataFrame, rng: np.random.Generator) -> Dict[str, Any]:
        return {}

    @abstractmethod
    def predict(self, collapsed_df: pd.DataFrame, error_model_df: pd.DataFrame) -> pd.DataFrame:
        raise NotImplementedError
