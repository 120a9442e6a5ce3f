"""elbencho_b200 — Blackwell-native GPU I/O benchmark worker (drop-in for elbencho's LocalWorker).

The product is the native library ``libelbencho_b200.so`` (hand-written sm_100a kernels + the
C++ worker pipeline behind a plain C ABI, see include/elbencho_b200.h). This package is the
Python host-side mirror of that ABI, used by tests and bench.py:

* :mod:`elbencho_b200.kernels` — on-GPU block modifiers/checkers (fill pattern, verify, random fill)
* :mod:`elbencho_b200.worker`  — WorkerManager / phases / results (mirrors the reference's
  WorkerManager + Worker interface)
"""
from .worker import (BenchPhase, IOEngine, OffsetRandAlgo, PathType, WorkerConfig,  # noqa: F401
                     WorkerManager, WorkerError)
from . import kernels  # noqa: F401

__all__ = ["BenchPhase", "IOEngine", "OffsetRandAlgo", "PathType", "WorkerConfig", "WorkerManager", "WorkerError", "kernels"]
