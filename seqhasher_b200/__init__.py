"""seqhasher_b200 — B200 (sm_100a) batch hashing engine for the per-record hash path of vmikk/seqhasher.

Layout:
  csrc/        CUDA kernels + the C ABI (include/seqhash_b200.h) -> libseqhash_b200.so
  binding.py   ctypes binding of that ABI (the only route from Python to the kernels)
  host/        host-side mirror of the reference interface (getHashFunc, processSequences, parseFlags)
There is no CPU fallback: compute entry points raise when the CUDA library or an sm_100 GPU is missing.
"""
from . import binding  # noqa: F401

__all__ = ["binding"]
