"""Record-range sharding of a CSR batch across GPUs / ranks.

Records are independent (the reference loop seqhasher.go:227-283 carries no state between records), so a
batch is cut into contiguous record ranges balanced by cumulative residue bytes; every rank hashes its range
and the digests are concatenated in rank order, which IS input order.  No collective on the data path."""
from __future__ import annotations

import numpy as np


def shard_ranges(offsets: np.ndarray, world: int, per_record_cost: int = 0) -> list[tuple[int, int]]:
    """world contiguous [lo, hi) record ranges covering 0..n, balanced by bytes + per_record_cost * records
    (per_record_cost models the fixed per-record work, e.g. 73 for SHA-1's padding block)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    if world <= 0:
        raise ValueError("world must be positive")
    cost = offsets - offsets[0] + per_record_cost * np.arange(n + 1, dtype=np.int64)
    total = int(cost[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        c = int(np.searchsorted(cost, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def slice_csr(residues: np.ndarray, offsets: np.ndarray, lo: int, hi: int):
    """The CSR sub-batch of records [lo, hi): (residues view, rebased offsets)."""
    o = np.asarray(offsets[lo: hi + 1], dtype=np.int64)
    return residues[int(o[0]): int(o[-1])], o - o[0]


def hash_sharded(residues: np.ndarray, offsets: np.ndarray, algos, flags: int, n_gpus: int):
    """One process driving n_gpus devices through the C ABI: one shb_batch per device, all submitted before
    any is waited on, digests concatenated in device order (= input order)."""
    from . import binding as B
    B.init(0)
    ranges = shard_ranges(offsets, n_gpus)
    batches = []
    try:
        for d, (lo, hi) in enumerate(ranges):
            r, o = slice_csr(residues, offsets, lo, hi)
            b = B.Batch(max(int(o[-1]), 1), max(hi - lo, 1), device=d)
            b.load(r, o)
            b.submit(algos, flags)
            batches.append(b)
        for b in batches:
            b.wait()
        outs = [np.concatenate([b.digests(k) for b in batches]) for k in range(len(algos))]
        lens = np.concatenate([b.out_lengths() for b in batches])
        return outs, lens
    finally:
        for b in batches:
            b.close()
