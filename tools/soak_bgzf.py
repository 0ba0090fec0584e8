"""Randomised soak of the BGZF path (device inflate + boundary search + splitter) against the plain-text path of the same
library: text kinds that stress different corners of DEFLATE (2-bit DNA, FASTQ with wide quality alphabets = long Huffman
codes, protein FASTA, long tandem repeats = overlapping copies and matches of 258, arbitrary header bytes), every zlib
level and strategy (stored, fixed-Huffman, RLE, Huffman-only blocks), member sizes from 64 B to 64 KiB, random chunking and
look-ahead.  usage: python tools/soak_bgzf.py [--iters 300] [--seed 1]"""
import argparse
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_text(rng):
    kind = int(rng.integers(0, 6))
    nrec = int(rng.integers(1, 400))
    parts = []
    fastq = kind in (1, 5)
    for i in range(nrec):
        L = int(rng.integers(0, 2000)) if rng.random() < 0.9 else int(rng.integers(2000, 150_000))
        if kind == 0 or fastq:
            seq = rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=L).tobytes()
        elif kind == 2:
            seq = rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZ*-", dtype=np.uint8), size=L).tobytes()
        elif kind == 3:
            unit = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(1, 40))).tobytes()
            seq = (unit * (L // len(unit) + 1))[:L]
        else:
            seq = (b"A" * L) if rng.random() < 0.5 else rng.choice(np.frombuffer(b"AC", dtype=np.uint8), size=L).tobytes()
        name = b"r%d " % i + bytes(rng.integers(33, 127, size=int(rng.integers(0, 30))).astype(np.uint8))
        if fastq:
            q = rng.integers(33, 74 if kind == 1 else 127, size=L).astype(np.uint8).tobytes() if L else b""
            parts.append(b"@" + name + b"\n" + seq + b"\n+\n" + q + b"\n")
        else:
            w = int(rng.choice([60, 70, 80, 10 ** 9]))
            body = b"\n".join(seq[k:k + w] for k in range(0, L, w)) if L else b""
            parts.append(b">" + name + b"\n" + body + (b"\n" if L else b""))
    return b"".join(parts), (2 if fastq else 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    from seqhasher_b200 import binding as B
    from seqhasher_b200.host import bgzf
    B.init(1)
    rng = np.random.default_rng(a.seed)
    tp = B.TextProc(48 << 20)
    bad = 0
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED]
    for it in range(a.iters):
        text, fmt = make_text(rng)
        if not text.strip():
            continue
        k = int(rng.integers(1, 4))
        ids = [int(x) for x in rng.integers(0, 12, size=k)]
        flags = B.FLAG_UPPER if rng.random() < 0.5 else 0
        ho = bool(rng.random() < 0.5)
        label = None if rng.random() < 0.3 else b"in.fa"
        # plain path (one chunk: the text is far below the capacity; index overflow handled by re-running the tail).  Like
        # the hosts, it ends the text at its last non-blank byte: a last FASTQ record with an empty sequence AND an empty quality
        # line is then an input that ends inside a record, on both paths (oracle/reader.py rule 14)
        want, pos = [], 0
        plain = text.rstrip(b" \t\r\n\v\f")
        while pos < len(plain):
            m = len(plain) - pos
            tp.input[:m] = np.frombuffer(plain, dtype=np.uint8)[pos:]
            o, consumed, nrec, _ = tp.run(m, fmt, True, ids, flags, ho, label)
            want.append(o.tobytes())
            if consumed == 0 or nrec == 0:
                break
            pos += consumed
        want = b"".join(want)
        level = int(rng.integers(0, 10))
        strat = strategies[int(rng.integers(0, len(strategies)))]
        block = int(rng.choice([64, 257, 1000, 4096, 20000, 65280, 65536])) if rng.random() < 0.7 else int(rng.integers(64, 65537))
        gz = bgzf.compress(text, level, block, strategy=strat)
        members = bgzf.scan_members(gz)
        own = int(rng.choice([1, 2, 3, 7, 50, 10 ** 6]))
        if len(members) / own > 400:
            own = len(members) // 400 + 1
        got = b"".join(o for o, _, _ in bgzf.run_chunks(tp, gz, members, fmt, 0, ids, flags, ho, label, own, int(rng.integers(1, 4))))
        if got != want:
            bad += 1
            print(f"MISMATCH it={it} fmt={fmt} level={level} strategy={strat} block={block} own={own} len={len(text)} ids={ids}", flush=True)
    tp.close()
    print(f"bgzf soak: {a.iters} random files, {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
