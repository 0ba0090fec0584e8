#!/usr/bin/env python
"""bench.py — benchmark of the per-record hash path on every configuration BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--only c2,c3,c4,c5]

The headline line (`metric`, `value`, `roofline`, `e2e`, `cpu_baseline`) is BASELINE.json configs[1] ("C2"):
10 M synthetic 250-bp amplicon reads per GPU, SHA-1 (value) and XXH3 (`xxh3` object) — weak scaling, every rank
hashes its own 10 M records.  The same JSON line carries one object per further configuration:
  c3  100 M x 150 bp reads, xxhash,xxh3,murmur3 fused            — STRONG scaling: 100 M / N records per rank
  c4  1 M x 10-50 kb reads, blake3 + nthash, case-sensitive      — strong scaling: 1 M / N records per rank
  c5  5 M x 30-3000 bp reads, all 12 algorithms, strip + upper   — strong scaling; + gzip-text end to end (N = 1)
each with ms_per_step, seqs/s, algorithmic GB/s, a roofline object, an end-to-end number through the C ABI with
host buffers, and a digest checksum compared with the CPU oracle (oracle/cpu_bench) on a sample of the same records.

One process per GPU (torchrun for N > 1); no data-path collective — records are independent; NCCL
carries the barrier and the max-over-ranks of the timings only.  A step = one pass of the hot path over the rank's
records.  Kernel-only numbers: inputs resident in HBM, CUDA events on the launching stream, max over ranks.  `e2e`
numbers go through shb_submit / shb_wait with HOST buffers (pinned H2D, kernels, D2H inside the timed region).
`--impl reference` times the CPU restatement of the reference (the oracle; the Go reference cannot be built in this
image) on all host cores on the C2 workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RECORDS = 10_000_000
READ_LEN = 250
SEED = 0xC2
LOWER_PERMILLE = 50            # 5 % soft-masked bases so upper-casing has work to do
SHA1_INSTR_PER_BLOCK = 613     # SURVEY.md §8d contract figure (LOP3/IADD3/SHF/PRMT fused-op count)
MD5_INSTR_PER_BLOCK = 324      # SURVEY.md §8d estimates for the other INT-bound hashes
B3_INSTR_PER_COMPRESS = 690
KECCAK_INSTR_PER_ROUND = 190
FAST_INSTR_PER_BYTE = 10.0     # xxhash 1.1 + xxh3 1 + cityhash 2 + murmur3 2.1 + nthash 1.5 + rapidhash 1 (x2) per byte
GOLDEN = 0x9E3779B97F4A7C15
LEN_SALT = 0x6c656e67746873    # oracle/cpu_bench.c uses the same length generator
ALL12 = ["sha1", "sha3", "md5", "xxhash", "xxh3", "cityhash", "murmur3", "nthash", "blake3", "k12", "rapidhash", "rapidhash32"]
DSIZE = dict(zip(ALL12, [20, 64, 16, 8, 8, 16, 16, 8, 32, 128, 8, 4]))
UNPINNED = {"cityhash", "rapidhash", "rapidhash32"}   # oracle pinned by one 4-byte reference vector only


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch from the committed ncu captures (profiles/traffic.json): measured once per kernel change,
    never in the run that prints the line."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None
        self.t_load = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.monotonic(), line.strip()))

    def wait_ready(self, timeout=5.0):
        """nvidia-smi needs 0.1-1 s to start (longer on an 8-GPU box): wait for its first line."""
        t0 = time.monotonic()
        while self.proc and not self.samples and time.monotonic() - t0 < timeout:
            time.sleep(0.01)

    def mark_load(self):
        self.t_load = time.monotonic()

    def under_load(self):
        return [s for t, s in self.samples if self.t_load is not None and t >= self.t_load]

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.under_load():
            parts = [x.strip() for x in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


CPU_BENCH = os.path.join(ROOT, "oracle", "_build", "cpu_bench")


def run_cpu_bench(n, read_len, seed, lower, flags, threads, steps, algos="sha1", n_permille=0):
    """The oracle port of seqhasher.go:241-259 as its own OpenMP process (oracle/cpu_bench.c).
    read_len: int, or (lo, hi) for the variable-length generator."""
    if not os.path.exists(CPU_BENCH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    ls = f"{read_len[0]}-{read_len[1]}" if isinstance(read_len, tuple) else str(read_len)
    p = subprocess.run([CPU_BENCH, str(n), ls, hex(seed), str(lower), str(flags), str(threads), "1", str(steps), algos,
                        str(n_permille)], capture_output=True, text=True, check=True)
    return json.loads(p.stdout.strip().splitlines()[-1])


def run_reference(args):
    """CPU arm: the oracle port of seqhasher.go:241-259 on all host cores, on the C2 workload (same records per step
    as the GPU arm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.records
    r = run_cpu_bench(n, READ_LEN, SEED, LOWER_PERMILLE, 3, cores, max(args.steps, 1))
    r1 = run_cpu_bench(400_000, READ_LEN, SEED, LOWER_PERMILLE, 3, 1, 1)
    val = r["seqs_per_s"]
    sample = (f"{n} x {READ_LEN} bp reads per step (the whole C2 workload of one GPU), sha1 + strip + upper-casing, "
              f"OpenMP record sharding over {cores} threads, "
              f"{'OpenSSL SHA-1 (SHA-NI)' if r['accel'] & 1 else 'portable C SHA-1'}; digest checksum {r['checksum']}")
    line = {"impl": "reference", "metric": "sha1 seqs/s hashed, 250-bp reads", "value": val, "unit": "seqs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "gbps": val * READ_LEN / 1e9,
            "config": {"workload": f"C2: {n} x {READ_LEN} bp synthetic amplicon reads per GPU, sha1, whitespace strip + upper-casing on",
                       "flags": "strip+upper",
                       "note": "CPU restatement of the reference (the Go reference cannot be built in this image); one host "
                               "hashes one GPU's records per step whatever --gpus says"},
            "cpu_baseline": {"value": val, "unit": "seqs/s", "cores": cores, "kind": "port", "sample": sample,
                             "single_thread_value": r1["seqs_per_s"]},
            "e2e": {"value": val, "unit": "seqs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath (NCCL prints its version banner with
    printf) write to fd 1 too, so fd 1 is pointed at stderr for the rest of the run and the JSON line goes to
    a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def np_splitmix64(z):
    import numpy as np
    with np.errstate(over="ignore"):
        z = z + np.uint64(GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_lengths(n, lo, hi, seed):
    """Record lengths of the variable-length configurations (identical generator in oracle/cpu_bench.c)."""
    import numpy as np
    with np.errstate(over="ignore"):
        z = np_splitmix64(np.uint64((seed ^ LEN_SALT) & 0xFFFFFFFFFFFFFFFF) + np.arange(n, dtype=np.uint64) * np.uint64(GOLDEN))
        return (np.uint64(lo) + (((z >> np.uint64(32)) * np.uint64(hi - lo + 1)) >> np.uint64(32))).astype(np.int64)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=N_RECORDS, help="C2 records per GPU (default: the C2 size)")
    ap.add_argument("--only", default="c2,c3,c4,c5", help="configurations to run (c2 always runs: it is the headline)")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the record counts of c3/c4/c5 (smoke runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    only = set(args.only.split(","))

    import numpy as np
    import torch

    from seqhasher_b200 import binding as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    B.init(0)
    dev = torch.device("cuda", local)
    st = torch.cuda.current_stream().cuda_stream
    hbm_peak, peak_src = load_peaks()
    traffic = load_traffic()
    cores = os.cpu_count() or 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def xor_fold(t: "torch.Tensor") -> int:
        """64-bit XOR of the tensor's bytes taken as little-endian words (oracle/cpu_bench.c's checksum)."""
        assert t.numel() % 8 == 0
        x = t.view(torch.int64)
        size = 1
        while size < x.numel():
            size *= 2
        if size != x.numel():
            x = torch.cat([x, torch.zeros(size - x.numel(), dtype=torch.int64, device=x.device)])
        while x.numel() > 1:
            h = x.numel() // 2
            x = x[:h] ^ x[h:]
        return int(x.item()) & 0xFFFFFFFFFFFFFFFF

    class Workload:
        """One rank's records of a synthetic configuration, resident in HBM (CSR: residues + int64 offsets)."""

        def __init__(self, n_total, length, seed, lower, n_perm, strong):
            # strong scaling: the job is n_total records, rank r owns the r-th contiguous range of it;
            # weak scaling: every rank owns n_total records of its own (seed + rank)
            self.length, self.seed, self.lower, self.n_perm = length, seed, lower, n_perm
            if strong:
                per = (n_total + world - 1) // world
                lo, hi = min(n_total, rank * per), min(n_total, (rank + 1) * per)
            else:
                lo, hi = 0, n_total
                self.seed = seed + rank
            self.n = hi - lo
            self.first = lo
            if isinstance(length, tuple):
                lens = synth_lengths(n_total if strong else hi, length[0], length[1], self.seed)
                cum = np.zeros(len(lens) + 1, dtype=np.int64)
                np.cumsum(lens, out=cum[1:])
                start = int(cum[lo])
                self.lens = lens[lo:hi]
                self.offs_host = cum[lo:hi + 1] - start
            else:
                start = lo * length
                self.lens = None
                self.offs_host = None
            self.total = int(self.offs_host[-1]) if self.offs_host is not None else self.n * length
            self.res = torch.empty(self.total + 16, dtype=torch.uint8, device=dev)
            B.synth_fill(self.res.data_ptr(), self.total, (self.seed + start * GOLDEN) & 0xFFFFFFFFFFFFFFFF, lower, n_perm,
                         stream=st, device=local)
            if self.offs_host is not None:
                self.offs = torch.from_numpy(self.offs_host).to(dev)
            else:
                self.offs = torch.arange(0, self.n + 1, dtype=torch.int64, device=dev) * length
            self.scratch = None

        def need_scratch(self):
            if self.scratch is None:
                self.scratch = torch.empty(B.scratch_bytes(self.total, self.n), dtype=torch.uint8, device=dev)
            return self.scratch.data_ptr()

        def outputs(self, algos):
            return [torch.empty(max(self.n * DSIZE[a], 8), dtype=torch.uint8, device=dev) for a in algos]

        def time(self, algos, flags, outs=None, scratch=True):
            """Kernel-only: ms per step (max over ranks), launches per step, clocks, outputs."""
            outs = outs or self.outputs(algos)
            sp = self.need_scratch() if scratch else None
            ptrs = [o.data_ptr() for o in outs]
            run = lambda: B.hash_device(self.res.data_ptr(), self.offs.data_ptr(), self.n, self.total, algos, flags, ptrs,
                                        d_scratch=sp, stream=st, device=local)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with ClockSampler(local) as cs:
                cs.wait_ready()
                for _ in range(args.warmup):
                    run()
                barrier()
                l0 = B.launch_count()
                cs.mark_load()
                e0.record()
                for _ in range(args.steps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                launches = B.launch_count() - l0
                # The timed region may last only a few tens of ms, nvidia-smi samples every 25 ms: keep the SAME
                # launches going (outside the event bracket) until at least 4 samples were taken under this load.
                t_tail = time.monotonic()
                while len(cs.under_load()) < 4 and time.monotonic() - t_tail < 1.0:
                    run()
                    torch.cuda.synchronize()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
            summ = cs.summary()
            summ["window"] = "timed region + the same launches continued until 4 samples (25 ms period) were taken"
            return ms, launches, summ, outs

        def checksum_vs_cpu(self, algos, flags, outs, sample_n):
            """XOR-fold of the GPU digests of the first sample_n records against oracle/cpu_bench on the same records
            (rank 0 owns the first records of the job)."""
            if rank != 0 or args.no_cpu:
                return None
            sample_n = min(sample_n, self.n) // 8 * 8
            if sample_n == 0:
                return None
            r = run_cpu_bench(sample_n, self.length, self.seed, self.lower, flags & 3, cores, 1, ",".join(algos), self.n_perm)
            g = 0
            for a, o in zip(algos, outs):
                g ^= xor_fold(o[: sample_n * DSIZE[a]])
            gpu = f"{g:016x}"
            out = {"records": sample_n, "cpu": r["checksum"], "gpu": gpu, "equal": r["checksum"] == gpu,
                   "cpu_seqs_per_s": r["seqs_per_s"], "cpu_threads": cores}
            unp = sorted(UNPINNED & set(algos))
            if unp:
                out["parity_unpinned"] = unp   # agreement of two restatements, not fidelity to the Go modules
            return out

        def e2e(self, algos, flags, records_per_batch, parts=8):
            """The same records through shb_submit / shb_wait from pinned HOST buffers, `parts` batches in flight."""
            if args.no_e2e:
                return None
            batches, lo = [], 0
            for _ in range(parts):
                hi = min(self.n, lo + records_per_batch)
                if hi <= lo:
                    break
                b0 = lo * self.length if self.offs_host is None else int(self.offs_host[lo])
                b1 = hi * self.length if self.offs_host is None else int(self.offs_host[hi])
                b = B.Batch(b1 - b0, hi - lo, device=local)
                torch.from_numpy(b.residues)[: b1 - b0].copy_(self.res[b0:b1])
                if self.offs_host is None:
                    b.offsets[: hi - lo + 1] = np.arange(hi - lo + 1, dtype=np.int64) * self.length
                else:
                    b.offsets[: hi - lo + 1] = self.offs_host[lo:hi + 1] - b0
                b.n = hi - lo
                batches.append(b)
                lo = hi
            torch.cuda.synchronize()
            nrec, nbytes = lo, (lo * self.length if self.offs_host is None else int(self.offs_host[lo]))

            def step():
                for b in batches:
                    b.submit(algos, flags)
                for b in batches:
                    b.wait()
            for _ in range(args.warmup):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            dt = time.perf_counter() - t0
            barrier()
            dt = max_over_ranks(dt) / args.steps
            got = [np.concatenate([b.digests(k) for b in batches]) for k in range(len(algos))]
            for b in batches:
                b.close()
            dig = sum(DSIZE[a] for a in algos)
            tot_rec, tot_bytes = sum_over_ranks(nrec), sum_over_ranks(nbytes)
            return {"value": tot_rec / dt, "unit": "seqs/s", "ms_per_step": dt * 1e3, "gbps": tot_bytes / dt / 1e9,
                    "records_per_step": int(tot_rec), "h2d_bytes_per_step": int(tot_bytes + 8 * (tot_rec + len(batches) * world)),
                    "d2h_bytes_per_step": int(dig * tot_rec),
                    "api": f"shb_submit/shb_wait on {len(batches)} pinned batches in flight per GPU"}, got, nrec

    # ================================================================ C2 (headline) ================================
    n, L = args.records, READ_LEN
    w2 = Workload(n, L, SEED, LOWER_PERMILLE, 0, strong=False)
    total = w2.total
    # the reference's default mode: whitespace strip (seqhasher.go:246) + upper-casing (:249-251)
    flags = B.FLAG_UPPER | B.FLAG_STRIP
    ms_sha1, launches, clocks, outs_sha1 = w2.time(["sha1"], flags)
    out_sha1 = outs_sha1[0]
    ms_xxh3, _, clocks_x, _ = w2.time(["xxh3"], flags)

    # ---- roofline of the two kernels (one launch per step each) ----
    blocks_per_rec = (L + 9 + 63) // 64
    theo = 148 * 128 * 1.965e9
    probes = {}
    for name, kind in (("alu_pipe_lop3", 0), ("fma_pipe_imad", 3), ("dual_pipe_lop3_imad", 6), ("sha1_compress_on_registers", 24)):
        try:
            probes[name] = B.probe_int_peak(kind, 400 if kind < 16 else 2000, device=local)   # measured live
        except Exception:
            probes[name] = None
    int_peak = probes.get("dual_pipe_lop3_imad") or theo
    int_src = ("measured live on this GPU: LOP3:IMAD 1:1 register-only probe = both integer pipes (ALU + FMA) issuing; "
               "the ALU pipe alone, the only home of SHA-1's LOP3/SHF, is `probes.alu_pipe_lop3`")
    sha1_ach = n * blocks_per_rec * SHA1_INSTR_PER_BLOCK / (ms_sha1 * 1e-3)
    alg_bytes_sha1 = total + 8 * (n + 1) + 20 * n
    alg_bytes_xxh3 = total + 8 * (n + 1) + 8 * n

    def traffic_of(key):
        t = traffic.get(key) if n == N_RECORDS else None
        return {"traffic": (t["dram_read"] + t["dram_write"]) if t else None, "traffic_measured_in_this_run": False,
                "traffic_source": (t or {}).get("source")}

    roof_sha1 = {"bound": "int", "achieved": sha1_ach / 1e12, "peak": int_peak / 1e12, "unit": "T lane-instr/s",
                 "frac": sha1_ach / int_peak, "peak_source": int_src, "frac_of_theoretical_issue": sha1_ach / theo,
                 "frac_of_alu_pipe": (sha1_ach / probes["alu_pipe_lop3"]) if probes.get("alu_pipe_lop3") else None,
                 "probes_T": {k: (v / 1e12 if v else None) for k, v in probes.items()}, "theoretical_issue_T": theo / 1e12,
                 "work": f"{SHA1_INSTR_PER_BLOCK} int ops x {blocks_per_rec} blocks x {n} records per launch",
                 "hbm_gbps": alg_bytes_sha1 / (ms_sha1 * 1e-3) / 1e9, "hbm_peak_gbps": hbm_peak,
                 "hbm_frac": alg_bytes_sha1 / (ms_sha1 * 1e-3) / 1e9 / hbm_peak,   # far from the HBM roof: the kernel is integer-bound
                 **traffic_of("c2_sha1_tile"), "algorithmic_bytes": alg_bytes_sha1,
                 "kernel": "hash_tile_kernel<sha1, upper, strip>"}
    xxh3_gbps = alg_bytes_xxh3 / (ms_xxh3 * 1e-3) / 1e9
    roof_xxh3 = {"bound": "hbm", "achieved": xxh3_gbps, "peak": hbm_peak, "unit": "GB/s", "frac": xxh3_gbps / hbm_peak,
                 "peak_source": peak_src, **traffic_of("c2_xxh3_tile"), "kernel": "hash_tile_kernel<xxh3, upper, strip>",
                 "bytes": f"{alg_bytes_xxh3} algorithmic bytes per launch (residues + 8 B offsets + 8 B digest per record)"}

    # ---- e2e through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e, got, nrec = w2.e2e(["sha1"], flags, (n + 7) // 8)
        same = bool((torch.from_numpy(got[0]).reshape(-1) == out_sha1[: nrec * 20].cpu()).all())   # parity of the two paths on this rank
        e2e["matches_device_path"] = same

    # ---- the whole record loop (split + strip + upper + hash + header rewrite + format) from TEXT, end to end ----
    e2e_text = None
    if not args.no_e2e:
        try:
            n_t = min(n, 4_000_000)
            base = 20_000
            seqs = w2.res[: base * L].cpu().numpy().reshape(base, L)
            block = b"".join(b">seq_%07d\n" % k + seqs[k].tobytes() + b"\n" for k in range(base))
            text = block * (n_t // base)
            n_t = base * (n_t // base)
            # two textproc objects fed from two host threads (each call blocks; the objects own their streams, so
            # one chunk's host-to-device copy overlaps the other's kernels and its copy back): the double buffering
            # a host loop over a large file would use
            from concurrent.futures import ThreadPoolExecutor
            tps = [B.TextProc(len(text) + 1024, device=local) for _ in range(2)]
            for t in tps:
                t.input[: len(text)] = np.frombuffer(text, dtype=np.uint8)
            tp = tps[0]
            ids = [B.algo_id("sha1")]
            run_tp = lambda t: t.run(len(text), B.FORMAT_FASTA, True, ids, B.FLAG_UPPER, True, b"big.fasta")
            for t in tps:
                for _ in range(2):
                    out_t, consumed, nrec, _ = run_tp(t)
            barrier()
            with ThreadPoolExecutor(2) as pool:
                t0 = time.perf_counter()
                results = list(pool.map(lambda k: run_tp(tps[k & 1]), range(2 * args.steps)))
                dt = time.perf_counter() - t0
            out_t, consumed, nrec, _ = results[-2] if tps[0] is tp else results[-1]
            barrier()
            dt = max_over_ranks(dt) / (2 * args.steps)
            e2e_text = {"value": n_t * world / dt, "unit": "seqs/s", "ms_per_step": dt * 1e3, "text_gbps": len(text) * world / dt / 1e9,
                        "h2d_bytes_per_step": len(text) * world, "d2h_bytes_per_step": int(out_t.size) * world,
                        "records_per_gpu": nrec, "device_ms": tp.last_ms(),
                        "api": "shb_textproc_run: single-line FASTA text in pinned host memory -> `file;sha1;name` lines in host "
                               "memory (--headersonly); replaces the whole loop body seqhasher.go:227-283",
                        "in_flight": 2, "first_line": out_t[:64].tobytes().split(b"\n")[0].decode()}
            for t in tps:
                t.close()
        except Exception as ex:   # keep the headline line even if this extra leg fails
            e2e_text = {"error": str(ex)}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_n = min(n, 4_000_000)
        r = run_cpu_bench(sample_n, L, SEED + rank, LOWER_PERMILLE, flags & 3, cores, 3)
        gpu_chk = f"{xor_fold(out_sha1[: sample_n * 20]):016x}"
        cpu = {"value": r["seqs_per_s"], "unit": "seqs/s", "cores": cores, "kind": "port",
               "sample": f"first {sample_n} of the {n} records, sha1 + strip + upper-casing, oracle/cpu_bench: OpenMP over {cores} threads, "
                         f"{'OpenSSL SHA-1 (SHA-NI)' if r['accel'] & 1 else 'portable C SHA-1'}, 3 steps; "
                         f"digest checksum cpu={r['checksum']} gpu={gpu_chk} equal={r['checksum'] == gpu_chk}"}
    del w2, outs_sha1, out_sha1
    torch.cuda.empty_cache()

    def guarded(fn):
        try:
            return fn()
        except Exception as ex:   # a failing extra configuration must not take the headline line with it
            import traceback
            traceback.print_exc()
            return {"error": f"{type(ex).__name__}: {ex}"}

    # ================================================================ C3 ============================================
    def bench_c3():
        n3 = int(100_000_000 * args.scale)
        algos = ["xxhash", "xxh3", "murmur3"]
        # C3 reads are 2 bits/base -> ACGT, i.e. upper case throughout, as sequencer FASTQ is.  (Round 1 timed
        # this configuration on reads with 2 % lower-case and 0.1 % N bytes sprinkled in; that variant is `softmasked_2pct`.)
        w = Workload(n3, 150, 0xC3, 0, 0, strong=True)
        ms, launches3, clk, outs = w.time(algos, B.FLAG_UPPER, scratch=False)
        nt = sum_over_ranks(w.n)
        alg = sum_over_ranks(w.total + 8 * (w.n + 1) + 32 * w.n)
        gbps = alg / (ms * 1e-3) / 1e9
        obj = {"workload": f"C3: {int(nt)} x 150 bp synthetic Illumina reads in total ({w.n} per GPU), xxhash,xxh3,murmur3 fused, upper-casing on",
               "scaling": "strong", "records": int(nt), "records_per_gpu": w.n, "algos": algos, "ms_per_step": ms,
               "value": nt / (ms * 1e-3), "unit": "seqs/s", "residue_gbps": sum_over_ranks(w.total) / (ms * 1e-3) / 1e9,
               "launches_per_step": launches3 // args.steps, "clocks": clk,
               "roofline": {"bound": "hbm", "achieved": gbps, "peak": hbm_peak * world, "unit": "GB/s", "frac": gbps / (hbm_peak * world),
                            "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
                            "algorithmic_bytes": int(alg), **({k: v for k, v in traffic_of("c3_tile").items()} if world == 1 and args.scale == 1.0 else
                                                              {"traffic": None, "traffic_measured_in_this_run": False}),
                            "kernel": "hash_tile_kernel<xxhash|xxh3|murmur3, upper>: one launch, one fused pass per record (hash_fused3.cuh)",
                            "bound_note": "HBM roofline as SURVEY.md §8d asks; the kernel is bound by the two integer pipes and the issue rate "
                                          "(three multiply-based hashes: ~9 instructions per byte), see DESIGN.md §3.1",
                            "bytes": "residues once + 8 B offsets + 32 B digests per record"},
               "checksum": w.checksum_vs_cpu(algos, B.FLAG_UPPER, outs, 2_000_000)}
        r = w.e2e(algos, B.FLAG_UPPER, 2_500_000)
        if r:
            obj["e2e"], got, nrec = r
            obj["e2e"]["matches_device_path"] = all(
                bool((torch.from_numpy(g).reshape(-1) == o[: nrec * DSIZE[a]].cpu()).all()) for g, o, a in zip(got, outs, algos))
        del w, outs
        torch.cuda.empty_cache()
        w = Workload(n3, 150, 0xC3, 20, 1, strong=True)
        ms_s, _, _, outs = w.time(algos, B.FLAG_UPPER, scratch=False)
        obj["softmasked_2pct"] = {"workload": "the same reads with 2 % lower-case and 0.1 % N bytes at random positions (round 1's C3 input): every "
                                              "16-byte group then takes the upper-casing path",
                                  "ms_per_step": ms_s, "value": nt / (ms_s * 1e-3), "frac_of_hbm": alg / (ms_s * 1e-3) / 1e9 / (hbm_peak * world),
                                  "checksum": w.checksum_vs_cpu(algos, B.FLAG_UPPER, outs, 1_000_000)}
        return obj

    # ================================================================ C4 ============================================
    def bench_c4():
        n4 = int(1_000_000 * args.scale)
        w = Workload(n4, (10_000, 50_000), 0xC4, 0, 1, strong=True)
        algos = ["blake3", "nthash"]
        ms, launches4, clk, outs = w.time(algos, 0)
        ms_b3, _, _, _ = w.time(["blake3"], 0)
        ms_nt, _, _, _ = w.time(["nthash"], 0)
        ms_x3, _, _, _ = w.time(["xxh3"], 0)
        nt, tb = sum_over_ranks(w.n), sum_over_ranks(w.total)
        compress = sum_over_ranks(float(((w.lens + 63) // 64).sum() + ((w.lens + 1023) // 1024 - 1).clip(min=0).sum()))
        alg = tb + 8 * (nt + world) + 40 * nt
        ops = compress * B3_INSTR_PER_COMPRESS
        obj = {"workload": f"C4: {int(nt)} x 10-50 kb synthetic nanopore-like reads in total ({w.n} per GPU), blake3 + nthash, case-sensitive",
               "scaling": "strong", "records": int(nt), "records_per_gpu": w.n, "residue_bytes": int(tb), "algos": algos,
               "ms_per_step": ms, "value": nt / (ms * 1e-3), "unit": "seqs/s", "residue_gbps": tb / (ms * 1e-3) / 1e9,
               "launches_per_step": launches4 // args.steps, "clocks": clk,
               "roofline": {"bound": "int", "achieved": ops / (ms * 1e-3) / 1e12, "peak": int_peak * world / 1e12, "unit": "T lane-instr/s",
                            "frac": ops / (ms * 1e-3) / (int_peak * world), "peak_source": int_src,
                            "work": f"{B3_INSTR_PER_COMPRESS} int ops x {int(compress)} BLAKE3 compressions (64-byte blocks + tree parents) per step; "
                                    "ntHash rides on the same pass over the bytes",
                            "hbm_gbps": alg / (ms * 1e-3) / 1e9, "hbm_frac": alg / (ms * 1e-3) / 1e9 / (hbm_peak * world),
                            "traffic": None, "traffic_measured_in_this_run": False,
                            "kernel": "b3_chunks_kernel<NT> (BLAKE3 chunk CVs + the chunk's ntHash share in one pass) + b3_merge_kernel"},
               "parts": {
                   "blake3": {"ms_per_step": ms_b3, "int_T": ops / (ms_b3 * 1e-3) / 1e12, "frac_of_int_peak": ops / (ms_b3 * 1e-3) / (int_peak * world),
                              "hbm_frac": (tb + 40 * nt) / (ms_b3 * 1e-3) / 1e9 / (hbm_peak * world)},
                   "nthash": {"ms_per_step": ms_nt, "gbps": (tb + 16 * nt) / (ms_nt * 1e-3) / 1e9,
                              "roofline": {"bound": "hbm", "achieved": (tb + 16 * nt) / (ms_nt * 1e-3) / 1e9, "peak": hbm_peak * world,
                                           "unit": "GB/s", "frac": (tb + 16 * nt) / (ms_nt * 1e-3) / 1e9 / (hbm_peak * world),
                                           "kernel": "nthash_warp_kernel (bit-plane path for units of pure ACGTU, look-ups for the rest)",
                                           "traffic": ((traffic.get("c4_nthash_warp") or {}).get("dram_read", 0) + (traffic.get("c4_nthash_warp") or {}).get("dram_write", 0)) or None
                                           if world == 1 and args.scale == 1.0 else None,
                                           "traffic_measured_in_this_run": False,
                                           "traffic_source": (traffic.get("c4_nthash_warp") or {}).get("source")}},
                   "xxh3": {"ms_per_step": ms_x3, "gbps": (tb + 16 * nt) / (ms_x3 * 1e-3) / 1e9,
                            "roofline": {"bound": "hbm", "achieved": (tb + 16 * nt) / (ms_x3 * 1e-3) / 1e9, "peak": hbm_peak * world,
                                         "unit": "GB/s", "frac": (tb + 16 * nt) / (ms_x3 * 1e-3) / 1e9 / (hbm_peak * world),
                                         "kernel": "xxh3_warp_kernel", "note": "not part of C4's hash list; the long-record XXH3 path on the C4 shape"}}},
               "checksum": w.checksum_vs_cpu(algos, 0, outs, 4_000)}
        r = w.e2e(algos, 0, 8_000)
        if r:
            obj["e2e"], got, nrec = r
            obj["e2e"]["matches_device_path"] = all(
                bool((torch.from_numpy(g).reshape(-1) == o[: nrec * DSIZE[a]].cpu()).all()) for g, o, a in zip(got, outs, algos))
        return obj

    # ================================================================ C5 ============================================
    def bench_c5():
        n5 = int(5_000_000 * args.scale)
        w = Workload(n5, (30, 3000), 0xC5, 50, 1, strong=True)
        f5 = B.FLAG_UPPER | B.FLAG_STRIP
        ms, launches5, clk, outs = w.time(ALL12, f5)
        nt, tb = sum_over_ranks(w.n), sum_over_ranks(w.total)
        ln = w.lens
        md_blocks = float(((ln + 9 + 63) // 64).sum())
        b3_comp = float(((ln + 63) // 64).clip(min=1).sum() + ((ln + 1023) // 1024 - 1).clip(min=0).sum())
        sha3_perms = float((ln // 72 + 1).sum())
        k12_perms = float(((ln + 1) // 168 + 1).sum())          # every C5 record is < 8192 B: one TurboSHAKE128 over M || 0x00... + squeeze
        ops = sum_over_ranks(md_blocks * (SHA1_INSTR_PER_BLOCK + MD5_INSTR_PER_BLOCK) + b3_comp * B3_INSTR_PER_COMPRESS +
                             sha3_perms * 24 * KECCAK_INSTR_PER_ROUND + k12_perms * 12 * KECCAK_INSTR_PER_ROUND + float(ln.sum()) * FAST_INSTR_PER_BYTE)
        alg = tb + 8 * (nt + world) + 328 * nt
        obj = {"workload": f"C5: {int(nt)} x 30-3000 bp variable-length synthetic reads in total ({w.n} per GPU), all 12 algorithms, strip + upper-casing on",
               "scaling": "strong", "records": int(nt), "records_per_gpu": w.n, "residue_bytes": int(tb), "algos": ALL12,
               "ms_per_step": ms, "value": nt / (ms * 1e-3), "unit": "seqs/s", "residue_gbps": tb / (ms * 1e-3) / 1e9,
               "launches_per_step": launches5 // args.steps, "clocks": clk,
               "roofline": {"bound": "int", "achieved": ops / (ms * 1e-3) / 1e12, "peak": int_peak * world / 1e12, "unit": "T lane-instr/s",
                            "frac": ops / (ms * 1e-3) / (int_peak * world), "peak_source": int_src,
                            "work": "SURVEY.md §8d op model: sha1 613 + md5 324 per 64-B block, blake3 690 per compression, Keccak 190 per round "
                                    "(sha3-512: 24 rounds per 72 B; k12: 12 per 168 B), the seven multiply-based hashes ~10 ops per byte together",
                            "int_ops_per_step": ops, "hbm_gbps": alg / (ms * 1e-3) / 1e9, "hbm_frac": alg / (ms * 1e-3) / 1e9 / (hbm_peak * world),
                            "traffic": None, "traffic_measured_in_this_run": False,
                            "kernel": "one fused launch for the 7 multiply-based hashes + one launch per crypto hash (sha3 dominates)"},
               "checksum": w.checksum_vs_cpu(ALL12, f5, outs, 200_000)}
        r = w.e2e(ALL12, f5, 125_000)
        if r:
            obj["e2e"], got, nrec = r
            obj["e2e"]["matches_device_path"] = all(
                bool((torch.from_numpy(g).reshape(-1) == o[: nrec * DSIZE[a]].cpu()).all()) for g, o, a in zip(got, outs, ALL12))
        del w, outs
        torch.cuda.empty_cache()
        # gzip text -> output text through the compiled host (rank 0 of a 1-GPU run): the configuration as BASELINE.json words it
        if world == 1 and not args.no_e2e:
            obj["e2e_gzip_cli"] = guarded(lambda: c5_gzip_cli(int(1_000_000 * min(1.0, args.scale))))   # 1.56 GB of text: start-up (~0.4 s) no longer dominates
        return obj

    def c5_gzip_cli(nrec):
        """Multi-line FASTA (wrapped at 60, 5 % soft-masked), gzip -6, through bin/seqhasher-b200 --headersonly with all 12
        hashes: whole-process wall time (CUDA context creation included), against the same command on the plain text."""
        import tempfile
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        lens = synth_lengths(nrec, 30, 3000, 0xC5)
        rng = np.random.default_rng(0xC5)
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(lens.sum()))
        for s in rng.integers(0, max(len(seq) - 100, 1), size=len(seq) // 2000):
            seq[s:s + 100] |= 0x20
        parts, pos = [], 0
        sb = seq.tobytes()
        for i, Lr in enumerate(lens):
            rec = sb[pos:pos + Lr]
            pos += int(Lr)
            parts.append(b">seq_%d len=%d\n" % (i, Lr))
            parts.append(b"\n".join(rec[k:k + 60] for k in range(0, int(Lr), 60)))
            parts.append(b"\n")
        text = b"".join(parts)
        step = 8 << 20

        def member(i):   # zlib releases the GIL: members are compressed in parallel, a multi-member gzip file like bgzip / pigz -i
            c = zlib.compressobj(6, zlib.DEFLATED, 31)
            return c.compress(text[i:i + step]) + c.flush()
        with ThreadPoolExecutor(min(cores, 32)) as pool:
            gz = b"".join(pool.map(member, range(0, len(text), step)))
        # the same text as blocked gzip (BGZF: what bgzip / htslib write): its members are inflated on the GPU (csrc/inflate.cuh)
        from seqhasher_b200.host import bgzf as bgzf_host
        bstep = 65280 * 64
        with ThreadPoolExecutor(min(cores, 32)) as pool:
            bz = b"".join(pool.map(lambda i: bgzf_host.compress(text[i:i + bstep], 6, 65280, eof_marker=False), range(0, len(text), bstep)))
        bz += bgzf_host.EOF_MEMBER
        cli = os.path.join(ROOT, "seqhasher_b200", "bin", "seqhasher-b200")
        out = {"records": nrec, "text_bytes": len(text), "gzip_bytes": len(gz), "bgzf_bytes": len(bz), "residue_bytes": int(lens.sum()),
               "command": "seqhasher-b200 --headersonly --name c5.fa --hash <all 12> c5.fa.gz out.txt", "host_cores": cores}
        with tempfile.TemporaryDirectory() as td:
            pg, pp, pb = os.path.join(td, "c5.fa.gz"), os.path.join(td, "c5.fa"), os.path.join(td, "c5.bgzf.fa.gz")
            open(pg, "wb").write(gz)
            open(pp, "wb").write(text)
            open(pb, "wb").write(bz)
            res = {}
            env = dict(os.environ, SHB_TIMING="1")
            for name, path, extra_env in (("gzip", pg, {}), ("plain", pp, {}), ("bgzf_device", pb, {}), ("bgzf_host_inflate", pb, {"SHB_NO_BGZF": "1"})):
                ts, outs_, ph = [], None, []
                for _ in range(3):
                    t0 = time.perf_counter()
                    # output to a file: ~700 B per record with all 12 hashes, and a pipe into this process would be the bottleneck
                    of = os.path.join(td, "out.txt")
                    p = subprocess.run([cli, "--headersonly", "--name", "c5.fa", "--hash", ",".join(ALL12), path, of], capture_output=True,
                                       check=True, env=dict(env, **extra_env))
                    ts.append(time.perf_counter() - t0)
                    outs_ = open(of, "rb").read()
                    ph.append(" | ".join(x.replace("[timing]", "").strip() for x in p.stderr.decode().splitlines() if "[timing]" in x))
                k = 1 + ts[1:].index(min(ts[1:]))
                res[name] = (ts[k], outs_)
                out[name + "_phases"] = ph[k]
            out["gzip_s"], out["plain_s"] = res["gzip"][0], res["plain"][0]
            out["value"], out["unit"] = nrec / res["gzip"][0], "seqs/s"
            out["text_gbps"] = len(text) / res["gzip"][0] / 1e9
            out["same_output_as_plain"] = res["gzip"][1] == res["plain"][1]
            out["bgzf"] = {"device_inflate_s": res["bgzf_device"][0], "host_inflate_s": res["bgzf_host_inflate"][0],
                           "value": nrec / res["bgzf_device"][0], "unit": "seqs/s", "text_gbps": len(text) / res["bgzf_device"][0] / 1e9,
                           "same_output_as_plain": res["bgzf_device"][1] == res["plain"][1] == res["bgzf_host_inflate"][1],
                           "how": "the same text as BGZF (64 KiB members, level 6): compressed members cross PCIe, one warp inflates one "
                                  "member (ISIZE + CRC-32 checked), shb_textproc_run_bgzf; host_inflate_s = the same file with SHB_NO_BGZF=1"}
            out["output_lines"] = res["gzip"][1].count(b"\n")
            out["note"] = "whole-process wall time, best of 2 after one warm-up run; includes CUDA context creation and teardown (~0.4 s)"
        out["bgzf_in_process"] = guarded(lambda: bgzf_in_process(text, bz, nrec))
        return out

    def bgzf_in_process(text, bz, nrec):
        """The BGZF path without process start-up: compressed members in pinned host memory -> output text in host memory through
        shb_textproc_run_bgzf, four chunks in flight (one thread + one shb_textproc + one stream each, as the compiled host runs
        them); all 12 hashes, --headersonly.  Output compared with the plain-text path of the same library."""
        import threading
        from seqhasher_b200.host import bgzf as bgzf_host
        members = bgzf_host.scan_members(bz)
        ids = [B.algo_id(a) for a in ALL12]
        own = (64 << 20) // 65280
        n = len(members)
        starts = list(range(0, n, own))
        mv = np.frombuffer(bz, dtype=np.uint8)
        nslots = 4
        tps = [B.TextProc(96 << 20) for _ in range(nslots)]
        stages = [t.bgzf_input(48 << 20) for t in tps]
        outs_, inf_ms, recs = {}, [0.0] * nslots, [0] * nslots

        def work(si, keep):
            t = tps[si]
            for ci in range(si, len(starts), nslots):
                k = starts[ci]
                n_own = min(own, n - k)
                extra = min(4, n - k - n_own)
                last = k + n_own + extra - 1
                lo, hi = int(members[k][3]), int(members[last][0] + members[last][1]) + 8
                stages[si][:hi - lo] = mv[lo:hi]
                tab = members[k:last + 1, :3].copy()
                tab[:, 0] -= np.uint64(lo)
                r = t.run_bgzf(hi - lo, tab, n_own, extra, 0 if k == 0 else -1, last == n - 1, B.FORMAT_FASTA, ids, B.FLAG_UPPER, True, b"c5.fa")
                recs[si] += r[2]
                inf_ms[si] += t.last_inflate_ms()
                if keep:
                    outs_[ci] = r[0].tobytes()
        best = None
        for rep in range(3):
            for i in range(nslots):
                recs[i], inf_ms[i] = 0, 0.0
            th = [threading.Thread(target=work, args=(i, rep == 0)) for i in range(nslots)]
            t0 = time.perf_counter()
            for x in th:
                x.start()
            for x in th:
                x.join()
            dt = time.perf_counter() - t0
            if rep and (best is None or dt < best[0]):
                best = (dt, sum(inf_ms))
        got = b"".join(outs_[ci] for ci in range(len(starts)))
        # the plain path of the same library on the same text, chunked at record starts by the host
        tp, want, pos = tps[0], [], 0
        tv = np.frombuffer(text, dtype=np.uint8)
        while pos < len(text):
            m = min(len(text) - pos, 64 << 20)
            if pos + m < len(text):
                m = text.rfind(b"\n>", pos, pos + m) + 1 - pos
            tp.input[:m] = tv[pos:pos + m]
            o, consumed, nr, _ = tp.run(m, B.FORMAT_FASTA, True, ids, B.FLAG_UPPER, True, b"c5.fa")
            want.append(o.tobytes())
            pos += m
        for t in tps:
            t.close()
        return {"value": nrec / best[0], "unit": "seqs/s", "text_gbps": len(text) / best[0] / 1e9, "wall_ms": best[0] * 1e3,
                "inflate_kernel_ms_sum": best[1], "inflate_kernel_text_gbps": len(text) / (best[1] * 1e-3) / 1e9 if best[1] else None,
                "chunks": len(starts), "in_flight": nslots, "records": int(sum(recs)),
                "h2d_bytes": len(bz), "text_bytes": len(text), "same_output_as_plain_path": got == b"".join(want),
                "api": "shb_textproc_run_bgzf: BGZF members in pinned host memory -> `file;h1;..;h12;name` lines in host memory; "
                       "inflate_kernel_* = bgzf_inflate_kernel alone (CUDA events), summed over chunks that overlap each other"}

    c3 = guarded(bench_c3) if "c3" in only else None
    torch.cuda.empty_cache()
    c4 = guarded(bench_c4) if "c4" in only else None
    torch.cuda.empty_cache()
    c5 = guarded(bench_c5) if "c5" in only else None

    if rank == 0:
        val = n * world / (ms_sha1 * 1e-3)
        line = {"metric": "sha1 seqs/s hashed, 250-bp reads", "value": val, "unit": "seqs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_sha1, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "gbps": val * L / 1e9,
                "config": {"workload": f"C2: {n} x {L} bp synthetic amplicon reads per GPU, sha1, whitespace strip + upper-casing on",
                           "flags": "strip+upper", "l2": f"inputs {total / 1e9:.2f} GB per step > 126 MB L2 (no reuse across steps)",
                           "parallelism": f"record-range sharding x{world}, no collective",
                           "other_configs": "c3 / c4 / c5 objects of this line: BASELINE.json configs[2..4], strong scaling over the ranks"},
                "roofline": roof_sha1,
                "xxh3": {"value": n * world / (ms_xxh3 * 1e-3), "unit": "seqs/s", "ms_per_step": ms_xxh3,
                         "gbps": n * world * L / (ms_xxh3 * 1e-3) / 1e9, "roofline": roof_xxh3, "clocks": clocks_x},
                "cpu_baseline": cpu, "e2e": e2e, "e2e_text": e2e_text, "gpu_launches": launches, "clocks": clocks,
                "c3": c3, "c4": c4, "c5": c5}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
