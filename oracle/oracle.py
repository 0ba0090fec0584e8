"""ctypes binding of the CPU ORACLE (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by seqhasher_b200/.  The oracle
restates seqhasher.go:241-259 and seqhasher.go:290-349 of the reference; see oracle.h
for the pinning status of each algorithm.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

ALGO_NAMES = ["sha1", "sha3", "md5", "xxhash", "xxh3", "cityhash", "murmur3", "nthash", "blake3", "k12",
              "rapidhash", "rapidhash32"]
DIGEST_SIZES = [20, 64, 16, 8, 8, 16, 16, 8, 32, 128, 8, 4]
FLAG_UPPER = 1
FLAG_STRIP = 2


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc only, no reference sources)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.sho_algo_id.argtypes = [ctypes.c_char_p]
        L.sho_algo_id.restype = ctypes.c_int
        L.sho_digest_size.argtypes = [ctypes.c_int]
        L.sho_digest_size.restype = ctypes.c_size_t
        L.sho_hash.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        L.sho_hash.restype = ctypes.c_size_t
        L.sho_normalise.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p]
        L.sho_normalise.restype = ctypes.c_size_t
        L.sho_hash_hex.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        L.sho_hash_hex.restype = ctypes.c_size_t
        L.sho_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.sho_batch.restype = ctypes.c_int
        L.sho_accel_available.restype = ctypes.c_int
        for name in ("sho_cityhash64", "sho_rapidhash_upstream_v1"):
            f = getattr(L, name)
            f.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
            f.restype = ctypes.c_uint64
        _lib = L
    return _lib


def algo_id(name: str) -> int:
    return lib().sho_algo_id(name.encode())


def hash_raw(algo: str | int, data: bytes) -> bytes:
    """Raw digest bytes in hex order; b"" for empty data (seqhasher.go:292-295)."""
    a = algo if isinstance(algo, int) else algo_id(algo)
    out = ctypes.create_string_buffer(128)
    n = lib().sho_hash(a, data, len(data), out)
    return out.raw[:n]


def get_hash_func(hash_type: str):
    """getHashFunc (seqhasher.go:290): returns data -> lowercase hex string; unknown names hash as SHA-1."""
    def f(data: bytes) -> str:
        out = ctypes.create_string_buffer(260)
        n = lib().sho_hash_hex(hash_type.encode(), data, len(data), out)
        return out.raw[:n].decode()
    return f


def normalise(data: bytes, flags: int) -> bytes:
    out = ctypes.create_string_buffer(max(len(data), 1))
    n = lib().sho_normalise(data, len(data), flags, out)
    return out.raw[:n]


def batch(residues: np.ndarray, offsets: np.ndarray, algos: list[str | int], flags: int, threads: int = 1,
          accel: bool = False):
    """seqhasher.go:241-259 over a CSR batch.  Returns ([n x digest_size uint8 arrays], post-normalise lengths)."""
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    ids = [a if isinstance(a, int) else algo_id(a) for a in algos]
    outs = [np.zeros((n, DIGEST_SIZES[a]), dtype=np.uint8) for a in ids]
    lens = np.zeros(n, dtype=np.int64)
    ptrs = (ctypes.c_void_p * len(ids))(*[o.ctypes.data for o in outs])
    ida = (ctypes.c_int * len(ids))(*ids)
    rc = lib().sho_batch(residues.ctypes.data if residues.size else None, offsets.ctypes.data, n, ida, len(ids),
                         flags, ptrs, lens.ctypes.data, threads, 1 if accel else 0)
    if rc != 0:
        raise RuntimeError(f"sho_batch failed: {rc}")
    return outs, lens
