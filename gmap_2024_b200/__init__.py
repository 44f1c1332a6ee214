"""gmap_2024_b200 -- B200 (sm_100a) alignment-DP engine for GMAP 2024-02-22.

The product is the native library ``csrc/libgmapdp_b200.so`` (C ABI in ``include/gmapdp_b200.h`` and
``include/gmapdp_shim.h``); this package is the thin Python host binding used by the tests and
bench.py.  There is no CPU fallback: creating an :class:`Engine` without a CUDA device raises.
(The package directory is ``gmap_2024_b200`` because ``gmap-2024_b200`` is not importable.)
"""
from .engine import Engine, Batch, ChainBatch, Stream, EngineError, load_library  # noqa: F401
