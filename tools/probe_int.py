"""Measure the B200 integer-pipe peaks the SHA-1 roofline is quoted against and write
profiles/INT_PEAKS.json.  Run on the GPU box: python tools/probe_int.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqhasher_b200 import binding as B  # noqa: E402

KINDS = {0: "lop3", 1: "iadd3", 2: "shf", 3: "imad", 4: "prmt", 5: "sha1_mix_synthetic", 6: "lop3_imad_1to1",
         7: "sha1_compress_regs"}


def main():
    B.init(1)
    out = {"unit": "lane-instr/s", "theoretical_issue_peak": 148 * 128 * 1.965e9, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    for k, name in KINDS.items():
        iters = 400 if k != 7 else 2000
        out[name] = B.probe_int_peak(k, iters)
        print(f"{name:22s} {out[name] / 1e12:8.3f} T lane-instr/s", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "INT_PEAKS.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
