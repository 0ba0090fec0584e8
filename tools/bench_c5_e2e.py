"""BASELINE config C5 end to end at reduced scale: variable-length multi-line FASTA (30-3000 bp, wrapped at 60
columns, soft-masked runs), gzip -6 input, all 12 algorithms, through the host CLI path (process_sequences_gpu:
host decompression -> shb_textproc).  Reports where the time goes; the host gunzip is the expected bound.
usage: python tools/bench_c5_e2e.py [--records 200000]"""
import argparse
import gzip
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_fasta(n, seed=0xC5):
    rng = np.random.default_rng(seed)
    lens = rng.integers(30, 3001, size=n)
    total = int(lens.sum())
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=total)
    # 5 % soft-masked: runs of 100 lower-case bases
    starts = rng.integers(0, max(total - 100, 1), size=total // 2000)
    for s in starts:
        seq[s:s + 100] |= 0x20
    out = io.BytesIO()
    pos = 0
    for i, L in enumerate(lens):
        out.write(b">seq_%d len=%d\n" % (i, L))
        rec = seq[pos:pos + L].tobytes()
        pos += L
        for k in range(0, L, 60):
            out.write(rec[k:k + 60] + b"\n")
    return out.getvalue(), total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=200_000)
    a = ap.parse_args()
    from seqhasher_b200 import binding as B
    from seqhasher_b200 import host as H
    B.init(1)
    text, residues = make_fasta(a.records)
    t0 = time.perf_counter()
    gz = gzip.compress(text, 6)
    t_gz = time.perf_counter() - t0
    t0 = time.perf_counter()
    assert gzip.decompress(gz) == text
    t_gunzip = time.perf_counter() - t0
    cfg = H.Config(hash_types=list(H.SUPPORTED_HASH_TYPES), headers_only=True, input_file_name="c5.fasta.gz")
    tp = B.TextProc(64 << 20)
    res = {}
    for label, data in (("plain text", text), ("gzip input", gz)):
        sink = io.BytesIO()
        H.process_sequences_gpu(io.BytesIO(data), sink, cfg, textproc=tp, log=io.StringIO())   # warm-up
        sink = io.BytesIO()
        t0 = time.perf_counter()
        H.process_sequences_gpu(io.BytesIO(data), sink, cfg, textproc=tp, log=io.StringIO())
        dt = time.perf_counter() - t0
        res[label] = {"seconds": dt, "seqs_per_s": a.records / dt, "text_MBps": len(text) / dt / 1e6, "out_bytes": sink.tell()}
    tp.close()
    print(json.dumps({"config": "C5 (scaled): multi-line FASTA 30-3000 bp, all 12 algorithms, --headersonly", "records": a.records,
                      "text_bytes": len(text), "gzip_bytes": len(gz), "residue_bytes": residues,
                      "host_gunzip_alone_s": t_gunzip, "host_gunzip_MBps": len(text) / t_gunzip / 1e6, "runs": res}))


if __name__ == "__main__":
    main()
