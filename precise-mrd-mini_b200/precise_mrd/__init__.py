"""precise_mrd -- B200-native drop-in for the simulate -> UMI-collapse -> call hot path of
altalanta/precise-mrd-mini.  Same entry points as the reference package
(``src/precise_mrd/__init__.py:8-22``); every stage runs in hand-written sm_100a CUDA
kernels behind ``libpmrd_b200.so`` (C ABI: ``include/pmrd.h``).  No CPU fallback.
"""
__version__ = "0.1.0"
