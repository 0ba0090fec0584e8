"""Memcheck substitute for the tile path (compute-sanitizer is closed on this GPU pool): run the randomised soak and the
tile / fused-strip shapes through a -DSHB_BOUNDS build of the library, in which every shared-memory access of the tile
kernels is checked against the CTA's dynamic shared-memory window, then report the violation count and the parity result.

    tools/build_variants.sh tile.cu bounds "-DSHB_BOUNDS"
    SHB_LIB=seqhasher_b200/_variants/lib_bounds.so python tools/bounds_check.py [--iters 300]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O           # noqa: E402
from seqhasher_b200 import binding as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    a = ap.parse_args()
    B.init(1)
    O.build()
    L = B.lib()
    L.shb_debug_oob.argtypes = []
    L.shb_debug_oob.restype = ctypes.c_longlong
    if L.shb_debug_oob() < 0:
        print(json.dumps({"error": "this library was not built with -DSHB_BOUNDS", "lib": B.LIB_PATH}))
        sys.exit(2)
    rng = np.random.default_rng(2025)
    alphabets = [b"ACGT", b"ACGTacgtN", b"ACGTacgtNn \n\t", bytes(range(1, 128))]
    bad = batches = records = 0
    for it in range(a.iters):
        n = int(rng.choice([1, 31, 127, 128, 129, 1000, 4000]))
        kind = it % 6
        if kind == 0:
            lens = np.full(n, int(rng.choice([1, 16, 17, 63, 64, 150, 240, 241, 250, 600])))
        elif kind == 1:
            lens = rng.integers(0, 700, size=n)
        elif kind == 2:
            lens = rng.integers(100, 600, size=n)
        elif kind == 3:
            lens = np.where(rng.random(n) < 0.3, 0, rng.integers(1, 300, size=n))
        elif kind == 4:
            lens = rng.integers(17, 241, size=n)                # the fused xxhash+xxh3+murmur3 pass
        else:
            lens = np.where(rng.random(n) < 0.02, rng.integers(3000, 9000, size=n), rng.integers(20, 200, size=n))   # oversize tiles -> per-CTA fallback
        lens = lens.astype(np.int64)
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        al = np.frombuffer(alphabets[rng.integers(len(alphabets))], dtype=np.uint8)
        res = al[rng.integers(0, len(al), size=int(offs[-1]))] if offs[-1] else np.zeros(0, np.uint8)
        if kind == 4 or rng.random() < 0.3:
            algos = ["xxhash", "xxh3", "murmur3"]
        else:
            algos = [B.ALGO_NAMES[i] for i in rng.choice(12, size=int(rng.integers(1, 5)), replace=False)]
        flags = int(rng.integers(0, 4))
        got, glen = B.hash_batch(res, offs, algos, flags)
        want, wlen = O.batch(res, offs, algos, flags & 3, threads=8)
        ok = (glen == wlen).all() and all((got[j] == want[j]).all() for j in range(len(algos)))
        bad += 0 if ok else 1
        batches += 1
        records += n
    out = {"lib": os.path.basename(B.LIB_PATH), "batches": batches, "records": records, "parity_mismatches": bad,
           "smem_out_of_window_accesses": int(L.shb_debug_oob()),
           "what": "every LDS/STS of hash_tile_kernel (SmemReader, GroupReader, read16_any, strip_in_smem, the in-place upper pass, "
                   "the tail-byte fix) checked against [dynamic smem base, base + %dynamic_smem_size)"}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bounds_check.json"), "w"), indent=1)
    sys.exit(1 if bad or out["smem_out_of_window_accesses"] else 0)


if __name__ == "__main__":
    main()
