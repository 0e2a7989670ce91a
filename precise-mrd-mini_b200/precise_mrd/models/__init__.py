"""Pluggable variant callers -- the reference's plugin seam (``models/base.py:15-50``)."""
from .base import VariantCaller
from .statistical import B200StatisticalVariantCaller, StatisticalVariantCaller

__all__ = ["VariantCaller", "B200StatisticalVariantCaller", "StatisticalVariantCaller"]
