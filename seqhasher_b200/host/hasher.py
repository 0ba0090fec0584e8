"""getHashFunc and the batched hash call on top of the C ABI (seqhasher.go:290-349, :246-259)."""
from __future__ import annotations

import sys

import numpy as np

from .. import binding as B

# seqhasher.go:41
SUPPORTED_HASH_TYPES = ["sha1", "sha3", "md5", "xxhash", "xxh3", "cityhash", "murmur3", "nthash", "blake3", "k12",
                        "rapidhash", "rapidhash32"]
EMPTY_MSG = "Error: Empty DNA sequence provided, resulting in an empty hash."   # seqhasher.go:293


def is_valid_hash_type(hash_type: str) -> bool:
    """isValidHashType (seqhasher.go:142-149): exact match against the 12 names."""
    return hash_type in SUPPORTED_HASH_TYPES


def _algo_ids(hash_types) -> list[int]:
    # getHashFunc's switch falls into `default:` = SHA-1 for any name it does not know, e.g. the untrimmed
    # " md5" that passes parseFlags' TrimSpace validation (seqhasher.go:132-137, :344-346)
    return [B.algo_id(t) if B.algo_id(t) >= 0 else B.algo_id("sha1") for t in hash_types]


class GpuHasher:
    """Batches records through one reusable shb_batch.  hash_records() is seqhasher.go:246-259 for a list
    of records: returns (hex strings per record per hash type, normalised sequences or None)."""

    def __init__(self, max_bytes: int = 64 << 20, max_records: int = 1 << 20, device: int = 0, log=sys.stderr):
        B.init(0)
        self.batch = B.Batch(max_bytes, max_records, device)
        self.log = log

    def close(self):
        self.batch.close()

    def hash_records(self, seqs: list[bytes], hash_types: list[str], case_sensitive: bool, want_seq: bool = False):
        ids = _algo_ids(hash_types)
        n = len(seqs)
        hexes: list[list[str]] = [[] for _ in range(n)]
        norm: list[bytes] | None = [b""] * n if want_seq else None
        flags = B.FLAG_STRIP | (0 if case_sensitive else B.FLAG_UPPER) | (B.FLAG_EMIT_SEQ if want_seq else 0)
        i = 0
        while i < n:
            # fill one batch
            j, used = i, 0
            while j < n and j - i < self.batch.max_records and used + len(seqs[j]) <= self.batch.max_bytes:
                used += len(seqs[j])
                j += 1
            if j == i:
                raise B.SeqhashError(f"record of {len(seqs[i])} bytes exceeds the batch capacity")
            buf = b"".join(seqs[i:j])
            self.batch.residues[:used] = np.frombuffer(buf, dtype=np.uint8)
            offs = np.zeros(j - i + 1, dtype=np.int64)
            np.cumsum([len(s) for s in seqs[i:j]], out=offs[1:])
            self.batch.offsets[: j - i + 1] = offs
            self.batch.submit(ids, flags, n=j - i)
            self.batch.wait()
            lens = self.batch.out_lengths()
            digs = [self.batch.digests(k) for k in range(len(ids))]
            if want_seq:
                nres, noffs = self.batch.out_sequences()
            for r in range(j - i):
                row = []
                for k in range(len(ids)):
                    if lens[r] == 0:   # seqhasher.go:292-295: one log line per hash type, "" as the hash
                        print(EMPTY_MSG, file=self.log)
                        row.append("")
                    else:
                        row.append(digs[k][r].tobytes().hex())
                hexes[i + r] = row
                if want_seq:
                    norm[i + r] = nres[noffs[r]:noffs[r + 1]].tobytes()
            i = j
        return hexes, norm


_shared: GpuHasher | None = None


def shared_hasher() -> GpuHasher:
    global _shared
    if _shared is None:
        _shared = GpuHasher()
    return _shared


def get_hash_func(hash_type: str):
    """getHashFunc(hashType) (seqhasher.go:290): returns f(data: bytes) -> lowercase hex str.  The data is
    hashed as given — strip/upper are processSequences' job (seqhasher.go:246-251), not getHashFunc's."""
    def f(data: bytes) -> str:
        h = shared_hasher()
        if len(data) == 0:
            print(EMPTY_MSG, file=h.log)
            return ""
        b = h.batch
        b.residues[: len(data)] = np.frombuffer(data, dtype=np.uint8)
        b.offsets[0:2] = [0, len(data)]
        b.submit(_algo_ids([hash_type]), 0, n=1)
        b.wait()
        return b.digests(0)[0].tobytes().hex()
    return f
