"""BASELINE config C5 (scaled) end to end through the COMPILED host: variable-length multi-line FASTA
(30-3000 bp, wrapped at 60 columns, 5 % soft-masked), compressed input, all 12 algorithms, --headersonly.
Reports (a) the decompression feeders alone (shb-zcat: single-stream library path vs the parallel readers) and
(b) whole-process wall time of `seqhasher-b200` per input format with one CUDA context kept resident.
usage: python tools/bench_c5_cli.py [--records 400000] [--small 130000]"""
import argparse
import json
import multiprocessing as mp
import os
import shutil
import struct
import subprocess
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "seqhasher_b200", "bin", "seqhasher-b200")
ZCAT = os.path.join(ROOT, "seqhasher_b200", "bin", "shb-zcat")
ALL12 = "sha1,sha3,md5,xxhash,xxh3,cityhash,murmur3,nthash,blake3,k12,rapidhash,rapidhash32"


def make_fasta(n, seed=0xC5):
    rng = np.random.default_rng(seed)
    lens = rng.integers(30, 3001, size=n)
    total = int(lens.sum())
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=total)
    for s in rng.integers(0, max(total - 100, 1), size=total // 2000):
        seq[s:s + 100] |= 0x20
    parts, pos = [], 0
    for i, L in enumerate(lens):
        rec = seq[pos:pos + L].tobytes()
        pos += L
        parts.append(b">seq_%d len=%d\n" % (i, L))
        parts.append(b"\n".join(rec[k:k + 60] for k in range(0, L, 60)))
        parts.append(b"\n")
    return b"".join(parts), total


def _deflate_piece(args):
    chunk, last = args
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    return c.compress(chunk) + c.flush(zlib.Z_FINISH if last else zlib.Z_SYNC_FLUSH)


def gzip_single_member(text, procs):
    """One standard gzip member, deflated in parallel pigz-style (independent 4 MiB pieces joined by sync flushes)."""
    step = 4 << 20
    pieces = [(text[i:i + step], i + step >= len(text)) for i in range(0, len(text), step)]
    with mp.Pool(procs) as pool:
        body = b"".join(pool.map(_deflate_piece, pieces))
    return b"\x1f\x8b\x08\0\0\0\0\0\0\x03" + body + struct.pack("<II", zlib.crc32(text) & 0xFFFFFFFF, len(text) & 0xFFFFFFFF)


def timed(cmd, env=None):
    t0 = time.perf_counter()
    with open(os.devnull, "wb") as dn:
        subprocess.run(cmd, cwd=ROOT, stdout=dn, check=True, env=env)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=400_000)
    ap.add_argument("--small", type=int, default=130_000, help="records of the bzip2/xz/zstd inputs")
    ap.add_argument("--no-gpu", action="store_true", help="feeders only")
    a = ap.parse_args()
    cores = os.cpu_count()
    tmp = "/tmp/c5cli"
    os.makedirs(tmp, exist_ok=True)
    text, residues = make_fasta(a.records)
    small, _ = make_fasta(a.small)
    files = {}
    open(f"{tmp}/c5.fa", "wb").write(text)
    files["plain"] = (f"{tmp}/c5.fa", len(text), a.records)
    t0 = time.perf_counter()
    open(f"{tmp}/c5.fa.gz", "wb").write(gzip_single_member(text, min(cores, 32)))
    t_gz = time.perf_counter() - t0
    files["gzip"] = (f"{tmp}/c5.fa.gz", len(text), a.records)
    open(f"{tmp}/small.fa", "wb").write(small)
    for ext, cmd in (("bz2", ["bzip2", "-k", "-9", "-f"]), ("xz", ["xz", "-k", "-f", "-T0", "-1"]), ("zst", ["zstd", "-q", "-f", "-3", "-T0"])):
        if shutil.which(cmd[0]):
            subprocess.run(cmd + [f"{tmp}/small.fa"], check=True)
            files[ext] = (f"{tmp}/small.fa.{ext}", len(small), a.small)
    out = {"config": "C5 (scaled): multi-line FASTA 30-3000 bp wrapped at 60, 5 % soft-masked, all 12 algorithms, --headersonly",
           "host_cores": cores, "records": a.records, "text_bytes": len(text), "residue_bytes": residues,
           "compressed_bytes": {k: os.path.getsize(v[0]) for k, v in files.items()}, "feeders": {}, "cli": {}}
    for name, (path, nbytes, _) in files.items():
        if name == "plain":
            continue
        row = {}
        for label, threads in (("single_stream_library", 1), ("parallel_reader", None)):
            env = dict(os.environ)
            if threads:
                env["SHB_DECOMP_THREADS"] = str(threads)
            best = None
            for _ in range(2):
                r = json.loads(subprocess.run([ZCAT, "-q", path], capture_output=True, check=True, env=env).stdout)
                best = r if best is None or r["seconds"] < best["seconds"] else best
            row[label] = {"seconds": best["seconds"], "text_MBps": best["MBps"]}
        row["speedup"] = row["single_stream_library"]["seconds"] / row["parallel_reader"]["seconds"]
        out["feeders"][name] = row
        print(json.dumps({name: row}), flush=True)
    if not a.no_gpu:
        sys.path.insert(0, ROOT)
        from seqhasher_b200 import binding as B
        B.init(1)
        keep = B.Batch(1 << 20, 16)   # resident CUDA context (what nvidia-persistenced provides)
        for name, (path, nbytes, nrec) in files.items():
            row = {}
            for label, threads in (("single_stream_library", 1), ("parallel_reader", None)):
                if name == "plain" and threads:
                    continue
                env = dict(os.environ)
                if threads:
                    env["SHB_DECOMP_THREADS"] = str(threads)
                cmd = [CLI, "--headersonly", "--hash", ALL12, path, "-"]
                ts = [timed(cmd, env) for _ in range(3)][1:]
                row[label] = {"min_s": min(ts), "seqs_per_s": nrec / min(ts), "text_MBps": nbytes / min(ts) / 1e6}
            out["cli"][name] = row
            print(json.dumps({"cli " + name: row}), flush=True)
        del keep
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c5_cli.json"), "w"), indent=1)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
