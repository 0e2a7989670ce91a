"""precise_mrd -- B200-native drop-in for the simulate -> UMI-collapse -> call hot path of
altalanta/precise-mrd-mini.  Same entry points as the reference package
(``src/precise_mrd/__init__.py:8-22``); every stage runs in hand-written sm_100a CUDA
kernels behind ``libpmrd_b200.so`` (C ABI: ``include/pmrd.h``).  No CPU fallback.
"""
__version__ = "0.1.0"

from .config import PipelineConfig, dump_config, load_config  # noqa: E402,F401


def __getattr__(name):
    # stage functions are imported lazily so that `import precise_mrd` works on a box without
    # the built library (the C-ABI check and the config/CLI plumbing do not need it)
    import importlib

    table = {
        "simulate_reads": ".simulate", "collapse_umis": ".collapse", "fit_error_model": ".error_model",
        "call_mrd": ".call", "poisson_test": ".call", "binomial_test": ".call",
        "benjamini_hochberg_correction": ".call", "roc_auc_score": ".metrics", "average_precision": ".metrics",
        "calculate_metrics": ".metrics", "process_fastq_to_dataframe": ".fastq", "detect_umi_format": ".fastq",
        "get_variant_caller": ".call",
    }
    if name in table:
        return getattr(importlib.import_module(table[name], __name__), name)
    raise AttributeError(name)
