"""Kernel-only throughput of the other BASELINE.json configs (C3, C4, C5 of SURVEY.md §8d) on one B200.
These are parity-test shapes, not the headline bench line; the numbers go to profiles/ as evidence.
usage: python tools/bench_configs.py [--scale 1.0] [--only C3,C4,C5] [--steps 5]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from seqhasher_b200 import binding as B  # noqa: E402


def run(name, n, lens, algos, flags, seed, lower, n_perm, steps, hbm_peak):
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    if isinstance(lens, int):
        offs = torch.arange(0, n + 1, dtype=torch.int64, device=dev) * lens
    else:
        g = torch.Generator(device=dev); g.manual_seed(seed)
        l = torch.randint(lens[0], lens[1] + 1, (n,), generator=g, device=dev, dtype=torch.int64)
        offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(l, 0, out=offs[1:])
    total = int(offs[-1].item())
    res = torch.empty(total + 16, dtype=torch.uint8, device=dev)
    B.synth_fill(res.data_ptr(), total, seed, lower, n_perm, stream=st)
    outs = [torch.empty(n * B.DIGEST_SIZES[B.algo_id(a)], dtype=torch.uint8, device=dev) for a in algos]
    scratch = torch.empty(B.scratch_bytes(total, n), dtype=torch.uint8, device=dev)   # strip pre-pass / BLAKE3 CVs
    call = lambda: B.hash_device(res.data_ptr(), offs.data_ptr(), n, total, algos, flags, [o.data_ptr() for o in outs],
                                 d_scratch=scratch.data_ptr() if scratch is not None else None, stream=st)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    l0 = B.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    dig = sum(B.DIGEST_SIZES[B.algo_id(a)] for a in algos)
    alg_bytes = total + 8 * (n + 1) + n * dig
    r = {"config": name, "records": n, "residue_bytes": total, "algos": algos, "flags": flags, "ms_per_step": ms,
         "seqs_per_s": n / (ms * 1e-3), "residue_gbps": total / (ms * 1e-3) / 1e9,
         "algorithmic_gbps": alg_bytes / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": alg_bytes / (ms * 1e-3) / 1e9 / hbm_peak,
         "launches_per_step": (B.launch_count() - l0) // steps}
    print(json.dumps(r), flush=True)
    del res, outs, offs, scratch
    torch.cuda.empty_cache()
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="C3,C4,C5")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--algos", default="nthash,blake3,xxh3,sha1,k12,xxhash,murmur3,cityhash,rapidhash,md5,sha3", help="C4split: which algorithms")
    a = ap.parse_args()
    B.init(1)
    hbm = 6465.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = float(json.load(open(p))["hbm_gbs"])
    only = a.only.split(",")
    out = []
    if "C3" in only:
        out.append(run("C3: 150-bp Illumina reads, xxhash,xxh3,murmur3 fused, upper", int(100_000_000 * a.scale), 150,
                       ["xxhash", "xxh3", "murmur3"], B.FLAG_UPPER, 0xC3, 0, 0, a.steps, hbm))
    if "C4" in only:
        out.append(run("C4: 10-50 kb nanopore-like reads, blake3+nthash, case-sensitive", int(1_000_000 * a.scale),
                       (10_000, 50_000), ["blake3", "nthash"], 0, 0xC4, 0, 1, a.steps, hbm))
    if "C4split" in only:
        for algo in a.algos.split(","):
            out.append(run(f"C4 shape (x{a.scale}), {algo} alone", int(1_000_000 * a.scale), (10_000, 50_000), [algo], 0,
                           0xC4, 0, 1, a.steps, hbm))
    if "C5" in only:
        out.append(run("C5: 30-3000 bp variable reads, all 12, strip+upper", int(5_000_000 * a.scale), (30, 3000),
                       B.ALGO_NAMES, B.FLAG_UPPER | B.FLAG_STRIP, 0xC5, 50, 1, a.steps, hbm))
    if "README" in only:
        for algo in ["md5", "sha1", "sha3", "xxhash", "murmur3", "cityhash", "nthash", "blake3", "xxh3", "k12", "rapidhash"]:
            out.append(run(f"README bench shape: 500k x 30-3000 bp, {algo}, case-sensitive", 500_000, (30, 3000), [algo], 0,
                           0x5EED, 0, 0, a.steps, hbm))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "results": out},
              open(os.path.join(ROOT, "gpurun_out", "bench_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
