"""A/B timing of library variants on the C4 shape (10-50 kb reads): nthash, xxh3, blake3 alone and blake3+nthash.
usage: python tools/ab_c4.py lib1.so lib2.so ..."""
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
n = 200_000
g = torch.Generator(device="cuda"); g.manual_seed(0xC4)
l = torch.randint(10_000, 50_001, (n,), generator=g, device="cuda", dtype=torch.int64)
offs = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); torch.cumsum(l, 0, out=offs[1:])
total = int(offs[-1].item())
res = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
B.synth_fill(res.data_ptr(), total, 0xC4, 0, 1, stream=st)
scr = torch.empty(B.scratch_bytes(total, n), dtype=torch.uint8, device="cuda")
out = {"bytes": total}
for algos, fl in ((["nthash"], 0), (["xxh3"], 0), (["blake3"], 0), (["blake3", "nthash"], 0), (["blake3", "nthash"], 0x800)):
    outs = [torch.empty(n * B.DIGEST_SIZES[B.algo_id(a)], dtype=torch.uint8, device="cuda") for a in algos]
    run = lambda: B.hash_device(res.data_ptr(), offs.data_ptr(), n, total, algos, fl, [o.data_ptr() for o in outs], d_scratch=scr.data_ptr(), stream=st)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    chk = 0
    for o in outs:
        chk ^= int(o.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFFFFFF
    ms = e0.elapsed_time(e1) / 5
    out["+".join(algos) + ("/2pass" if fl else "")] = [round(ms, 4), round(total / ms / 1e6, 1), "%%016x" %% chk]
print(json.dumps(out))
'''


def main():
    res = {}
    for lib in sys.argv[1:]:
        env = dict(os.environ, SHB_LIB=os.path.abspath(lib))
        p = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, env=env)
        if p.returncode != 0:
            print(lib, "FAILED", p.stderr[-400:])
            continue
        res[os.path.basename(lib)] = json.loads(p.stdout.strip().splitlines()[-1])
        print(os.path.basename(lib), res[os.path.basename(lib)], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_c4.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
