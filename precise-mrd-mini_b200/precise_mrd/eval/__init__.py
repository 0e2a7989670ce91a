from .lod import LODAnalyzer, estimate_lob, estimate_lod, estimate_loq  # noqa: F401
