"""Randomised soak of the C ABI against the oracle: random record counts, length distributions, byte alphabets
(whitespace, lower case, non-ACGT), flags (upper / strip / emit, forced kernel families) and algorithm subsets.
usage: python tools/soak_gpu.py [--iters 200] [--seed 1]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O           # noqa: E402  (tools/ may use the oracle as a checker)
from seqhasher_b200 import binding as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--big", action="store_true", help="tens of thousands of records per batch (many CTAs in every sort / scan)")
    a = ap.parse_args()
    B.init(1)
    O.build()
    rng = np.random.default_rng(a.seed)
    alphabets = [b"ACGT", b"ACGTacgtN", b"ACGTacgtNn \n\t", b"ACGTURYKMSWBDHVN-.*acgtn 0123\r\n", bytes(range(1, 128))]
    dists = ["fixed", "narrow", "wide", "bimodal", "long", "tiny", "empties"]
    extra_flags = [0, B.FLAG_FORCE_DIRECT, B.FLAG_NO_LONG_PATH, B.FLAG_NO_LENGTH_SORT, B.FLAG_FORCE_DIRECT | B.FLAG_NO_LENGTH_SORT]
    bad = 0
    for it in range(a.iters):
        dist = dists[rng.integers(len(dists))]
        n = int(rng.choice([20_000, 50_000, 130_000])) if a.big else int(rng.choice([1, 2, 31, 127, 128, 129, 1000, 1500, 3000, 6000]))
        if dist == "fixed":
            lens = np.full(n, int(rng.choice([1, 63, 64, 150, 250, 1000, 5000])))
        elif dist == "narrow":
            base = int(rng.choice([100, 800, 3000]))
            lens = rng.integers(base, base + base // 4 + 1, size=n)
        elif dist == "wide":
            lens = rng.integers(1, int(rng.choice([600, 3000] if a.big else [600, 3000, 20000])), size=n)
        elif dist == "bimodal":
            lens = np.where(rng.random(n) < 0.5, rng.integers(20, 60, size=n), rng.integers(2000, 3000 if a.big else 9000, size=n))
        elif dist == "long":
            n = min(n, 300)
            lens = rng.integers(8000, int(rng.choice([30000, 120000])), size=n)
        elif dist == "tiny":
            lens = rng.integers(0, 20, size=n)
        else:
            lens = np.where(rng.random(n) < 0.3, 0, rng.integers(1, 400, size=n))
        lens = lens.astype(np.int64)[:n]
        offs = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        al = np.frombuffer(alphabets[rng.integers(len(alphabets))], dtype=np.uint8)
        res = al[rng.integers(0, len(al), size=int(offs[-1]))] if offs[-1] else np.zeros(0, np.uint8)
        k = int(rng.integers(1, 5))
        algos = [B.ALGO_NAMES[i] for i in rng.choice(12, size=k, replace=bool(rng.integers(2)) and k > 1)]
        pick = rng.random()
        if pick < 0.15:
            algos = ["xxhash", "xxh3", "murmur3"]          # the fused one-pass set of config C3 (hash_fused3.cuh)
        elif pick < 0.25:
            algos = ["blake3", "nthash"] + ([B.ALGO_NAMES[int(rng.integers(12))]] if rng.random() < 0.3 else [])   # C4: one pass on long records
        flags = int(rng.integers(0, 4)) | extra_flags[rng.integers(len(extra_flags))]
        got, glen = B.hash_batch(res, offs, algos, flags)
        want, wlen = O.batch(res, offs, algos, flags & 3, threads=8)
        ok = (glen == wlen).all() and all((got[j] == want[j]).all() for j in range(len(algos)))
        if not ok:
            bad += 1
            print("MISMATCH", it, dist, n, algos, hex(flags), "alphabet", al.tobytes()[:12], flush=True)
            if not (glen == wlen).all():
                j = int(np.nonzero(glen != wlen)[0][0])
                print("   lengths differ first at record", j, "raw", int(lens[j]), "got", int(glen[j]), "want", int(wlen[j]),
                      "count", int((glen != wlen).sum()), flush=True)
            for q, nm in enumerate(algos):
                d = (got[q] != want[q]).reshape(len(lens), -1).any(axis=1)
                if d.any():
                    j = int(np.nonzero(d)[0][0])
                    print("   ", nm, "differs at", int(d.sum()), "records, first", j, "raw len", int(lens[j]), "post", int(wlen[j]),
                          "tile pos", j % 128, "bytes", res[offs[j]:offs[j] + 24].tobytes(), flush=True)
    print(f"soak: {a.iters} random batches, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
