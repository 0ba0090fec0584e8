"""Which pipe-balancing variant of the SHA-1 compression is fastest on this GPU? (hash_md.cuh)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqhasher_b200 import binding as B
B.init(1)
ref = None
for v in list(range(16)) + [24, 40, 56]:
    r = B.probe_int_peak(16 + v, 2000)
    c = B.probe_last_checksum()
    ref = c if ref is None else ref
    print(f"variant {v:2d} (rol30={v&1} sched={v>>1&1} rol5wide={v>>2&1} kw={v>>3&1}): {r/1e12:7.3f} T  checksum {'ok' if c == ref else 'MISMATCH'}", flush=True)
