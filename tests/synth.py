"""Test-side inputs: the numpy twin of the device generator (csrc/kernels.cu synth_fill_kernel) and CSR helpers."""
from __future__ import annotations

import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + _G
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_bytes(total: int, seed: int, lower_permille: int = 0, n_permille: int = 0, start: int = 0) -> np.ndarray:
    """Bytes [start, start+total) of the synthetic residue stream for `seed` (same function as the CUDA fill)."""
    j = np.arange(start, start + total, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = splitmix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + j * _G)
    c = np.frombuffer(b"ACGT", dtype=np.uint8)[(z & np.uint64(3)).astype(np.int64)].copy()
    u_n = (((z >> np.uint64(32)) & np.uint64(0xFFFFF)) * np.uint64(1000)) >> np.uint64(20)
    u_l = (((z >> np.uint64(8)) & np.uint64(0xFFFFF)) * np.uint64(1000)) >> np.uint64(20)
    c[u_n < np.uint64(n_permille)] = ord("N")
    c[u_l < np.uint64(lower_permille)] |= 0x20
    return c


def csr_from_lengths(lengths, seed: int, lower_permille: int = 0, n_permille: int = 0):
    lengths = np.asarray(lengths, dtype=np.int64)
    offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    return synth_bytes(int(offsets[-1]), seed, lower_permille, n_permille), offsets


def csr_from_records(records):
    lengths = [len(r) for r in records]
    offsets = np.zeros(len(records) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    residues = np.frombuffer(b"".join(records), dtype=np.uint8).copy() if offsets[-1] else np.zeros(0, np.uint8)
    return residues, offsets


# every block / stripe / chunk edge of the 12 algorithms (SURVEY.md §4 "what is NOT tested")
BOUNDARY_LENGTHS = sorted(set(
    list(range(0, 260)) + [263, 264, 265, 271, 272, 287, 288, 319, 320, 335, 336, 337, 383, 384, 385, 503, 504, 505,
                           511, 512, 513, 575, 576, 577, 671, 672, 673, 1007, 1008, 1009, 1023, 1024, 1025, 1087, 1088,
                           1089, 2047, 2048, 2049, 3071, 3072, 3073, 4095, 4096, 4097, 5000, 8063, 8064, 8065, 8190,
                           8191, 8192, 8193, 8200, 8359, 8360, 8361, 10000, 16383, 16384, 16385, 24575, 24576, 24577,
                           33000, 50000, 70000]))
