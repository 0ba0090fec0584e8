"""A/B timing of library variants on the C3 shape (fixed-length reads, xxhash,xxh3,murmur3) and a few neighbours.
usage: python tools/ab_c3.py lib1.so lib2.so ...   (each variant in its own subprocess; digests are checksummed)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
from seqhasher_b200 import binding as B
B.init(1)
st = torch.cuda.current_stream().cuda_stream
out = {}
for name, n, L, algos, flags, lower in %s:
    total = n * L
    res = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
    B.synth_fill(res.data_ptr(), total, 0xC3, lower, 1, stream=st)
    offs = torch.arange(0, n + 1, dtype=torch.int64, device="cuda") * L
    outs = [torch.empty(n * B.DIGEST_SIZES[B.algo_id(a)], dtype=torch.uint8, device="cuda") for a in algos]
    scr = torch.empty(B.scratch_bytes(total, n), dtype=torch.uint8, device="cuda")
    run = lambda: B.hash_device(res.data_ptr(), offs.data_ptr(), n, total, algos, flags, [o.data_ptr() for o in outs], d_scratch=scr.data_ptr(), stream=st)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    chk = 0
    for o in outs:
        chk ^= int(o.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF
    out[name] = [round(e0.elapsed_time(e1) / 10, 4), "%%016x" %% chk]
print(json.dumps(out))
'''
CASES = [("c3_150", 20_000_000, 150, ["xxhash", "xxh3", "murmur3"], 1, 20),
         ("c3_150_uppercase", 20_000_000, 150, ["xxhash", "xxh3", "murmur3"], 1, 0),
         ("c3_150_strip", 20_000_000, 150, ["xxhash", "xxh3", "murmur3"], 3, 20),
         ("c3_100_nocase", 20_000_000, 100, ["xxhash", "xxh3", "murmur3"], 0, 20),
         ("c3_230", 10_000_000, 230, ["xxhash", "xxh3", "murmur3"], 1, 20),
         ("c2_250_3", 10_000_000, 250, ["xxhash", "xxh3", "murmur3"], 1, 20)]


def main():
    res = {}
    for lib in sys.argv[1:]:
        env = dict(os.environ, SHB_LIB=os.path.abspath(lib))
        p = subprocess.run([sys.executable, "-c", CHILD % (ROOT, repr(CASES))], capture_output=True, text=True, env=env)
        if p.returncode != 0:
            print(lib, "FAILED", p.stderr[-400:])
            continue
        res[os.path.basename(lib)] = json.loads(p.stdout.strip().splitlines()[-1])
        print(os.path.basename(lib), res[os.path.basename(lib)], flush=True)
    ref = None
    for lib, r in res.items():
        chks = {k: v[1] for k, v in r.items()}
        if ref is None:
            ref = chks
        elif chks != ref:
            print("CHECKSUM MISMATCH in", lib)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_c3.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
