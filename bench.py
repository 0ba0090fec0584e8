#!/usr/bin/env python
"""bench.py — headline benchmark of the per-record hash path (BASELINE.json configs[1], "C2"):
10 M synthetic 250-bp amplicon reads, SHA-1 (value) and XXH3 (extra object), per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N > 1, weak scaling: every rank hashes its own 10 M records; no
data-path collective — records are independent).  A step = one pass of the hot path
(upper-case -> hash) over the rank's whole batch.  `value` is kernel-only (inputs resident in HBM, CUDA
events on the launching stream, max over ranks); `e2e` goes through the C ABI with HOST buffers
(pinned H2D, kernels, D2H inside the timed region).  `--impl reference` times the CPU restatement of the
reference (the oracle; the Go reference cannot be built in this image) on all host cores.
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


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def run_cpu_bench(n, read_len, seed, lower, flags, threads, steps, algos="sha1"):
    """The oracle port of seqhasher.go:241-259 as its own OpenMP process (oracle/cpu_bench.c)."""
    if not os.path.exists(CPU_BENCH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    p = subprocess.run([CPU_BENCH, str(n), str(read_len), hex(seed), str(lower), str(flags), str(threads), "1", str(steps), algos],
                       capture_output=True, text=True, check=True)
    return json.loads(p.stdout.strip().splitlines()[-1])


def run_reference(args):
    """CPU arm: the oracle port of seqhasher.go:241-259 on all host cores, same workload shape."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 4_000_000                                   # bounded sample of the 10 M-record workload per step
    r = run_cpu_bench(n, READ_LEN, SEED, LOWER_PERMILLE, 3, cores, max(args.steps, 1))
    r1 = run_cpu_bench(400_000, READ_LEN, SEED, LOWER_PERMILLE, 3, 1, 1)
    val = r["seqs_per_s"]
    sample = (f"{n} x {READ_LEN} bp reads per step (first 4 M records of the 10 M workload), sha1 + strip + upper-casing, "
              f"OpenMP record sharding over {cores} threads, "
              f"{'OpenSSL SHA-1 (SHA-NI)' if r['accel'] & 1 else 'portable C SHA-1'}; digest checksum {r['checksum']}")
    line = {"impl": "reference", "metric": "sha1 seqs/s hashed, 250-bp reads", "value": val, "unit": "seqs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "gbps": val * READ_LEN / 1e9,
            "config": {"workload": "C2: 250-bp amplicon reads, sha1, strip + upper-casing on (bounded 4 M-record sample per step)",
                       "note": "CPU restatement of the reference (the Go reference cannot be built in this image)"},
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


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=N_RECORDS, help="records per GPU (default: the C2 size)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

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

    n, L = args.records, READ_LEN
    total = n * L
    d_res = torch.empty(total + 16, dtype=torch.uint8, device=dev)
    B.synth_fill(d_res.data_ptr(), total, SEED + rank, LOWER_PERMILLE, 0, stream=st, device=local)
    d_offs = torch.arange(0, n + 1, dtype=torch.int64, device=dev) * L
    # the reference's default mode: whitespace strip (seqhasher.go:246) + upper-casing (:249-251)
    flags = B.FLAG_UPPER | B.FLAG_STRIP
    d_scratch = torch.empty(B.scratch_bytes(total, n), dtype=torch.uint8, device=dev)

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

    def time_kernel(algo: str):
        out = torch.empty(n * B.DIGEST_SIZES[B.algo_id(algo)], dtype=torch.uint8, device=dev)
        run = lambda: B.hash_device(d_res.data_ptr(), d_offs.data_ptr(), n, total, [algo], flags, [out.data_ptr()],
                                    d_scratch=d_scratch.data_ptr(), stream=st, device=local)
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
            # The timed region lasts a few tens of ms, nvidia-smi samples every 25 ms: keep the SAME launches
            # going (outside the event bracket) until at least 4 samples were taken under this load.
            t_tail = time.monotonic()
            while len(cs.under_load()) < 4 and time.monotonic() - t_tail < 1.0:
                run()
                torch.cuda.synchronize()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        summ = cs.summary()
        summ["window"] = "timed region + the same launches continued until 4 samples (25 ms period) were taken"
        return ms, launches, summ, out

    ms_sha1, launches, clocks, out_sha1 = time_kernel("sha1")
    ms_xxh3, _, clocks_x, _ = time_kernel("xxh3")

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
    roof_sha1 = {"bound": "int", "achieved": sha1_ach / 1e12, "peak": int_peak / 1e12, "unit": "T lane-instr/s",
                 "frac": sha1_ach / int_peak, "peak_source": int_src, "frac_of_theoretical_issue": sha1_ach / theo,
                 "frac_of_alu_pipe": (sha1_ach / probes["alu_pipe_lop3"]) if probes.get("alu_pipe_lop3") else None,
                 "probes_T": {k: (v / 1e12 if v else None) for k, v in probes.items()}, "theoretical_issue_T": theo / 1e12,
                 "work": f"{SHA1_INSTR_PER_BLOCK} int ops x {blocks_per_rec} blocks x {n} records per launch",
                 "hbm_gbps": alg_bytes_sha1 / (ms_sha1 * 1e-3) / 1e9, "hbm_peak_gbps": hbm_peak,
                 "hbm_frac": alg_bytes_sha1 / (ms_sha1 * 1e-3) / 1e9 / hbm_peak,   # far from the HBM roof: the kernel is integer-bound
                 "traffic": (2_580_202_000 + 197_736_960) if n == N_RECORDS else None,
                 "traffic_source": "ncu --set full, profiles/r01_sha1_tile.summary.txt: dram read 2.580 GB + write 0.198 GB per launch "
                                   "(algorithmic: %.3f GB)" % (alg_bytes_sha1 / 1e9),
                 "kernel": "hash_tile_kernel<sha1, upper, strip>"}
    xxh3_gbps = alg_bytes_xxh3 / (ms_xxh3 * 1e-3) / 1e9
    roof_xxh3 = {"bound": "hbm", "achieved": xxh3_gbps, "peak": hbm_peak, "unit": "GB/s", "frac": xxh3_gbps / hbm_peak,
                 "peak_source": peak_src, "traffic": (2_580_072_000 + 82_846_976) if n == N_RECORDS else None,
                 "traffic_source": "ncu --set full, profiles/r01_xxh3_tile.summary.txt", "kernel": "hash_tile_kernel<xxh3, upper, strip>",
                 "bytes": f"{alg_bytes_xxh3} algorithmic bytes per launch (residues + 8 B offsets + 8 B digest per record)"}

    # ---- e2e through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        parts = 8
        per = (n + parts - 1) // parts
        batches = []
        for p in range(parts):
            lo, hi = p * per, min(n, (p + 1) * per)
            if hi <= lo:
                break
            b = B.Batch((hi - lo) * L, hi - lo, device=local)
            torch.from_numpy(b.residues)[: (hi - lo) * L].copy_(d_res[lo * L: hi * L])
            b.offsets[: hi - lo + 1] = np.arange(hi - lo + 1, dtype=np.int64) * L
            b.n = hi - lo
            batches.append(b)
        torch.cuda.synchronize()

        def step():
            for b in batches:
                b.submit(["sha1"], flags)
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
        # parity of the two paths on this rank
        got = np.concatenate([b.digests(0) for b in batches])
        same = bool((torch.from_numpy(got).reshape(-1) == out_sha1.cpu()).all())
        e2e = {"value": n * world / dt, "unit": "seqs/s", "h2d_bytes_per_step": (total + 8 * (n + parts)) * world,
               "d2h_bytes_per_step": 20 * n * world, "ms_per_step": dt * 1e3, "gbps": n * world * L / dt / 1e9,
               "api": f"shb_submit/shb_wait on {len(batches)} pinned batches in flight per GPU", "matches_device_path": same}
        for b in batches:
            b.close()

    # ---- the whole record loop (split + strip + upper + hash + header rewrite + format) from TEXT, end to end ----
    e2e_text = None
    if not args.no_e2e:
        try:
            n_t = min(n, 4_000_000)
            base = 20_000
            seqs = d_res[: base * L].cpu().numpy().reshape(base, L)
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
        cores = os.cpu_count() or 1
        sample_n = min(n, 4_000_000)
        r = run_cpu_bench(sample_n, L, SEED + rank, LOWER_PERMILLE, flags & 3, cores, 3)
        # XOR-fold checksum of the GPU digests of the same records
        x = out_sha1[: sample_n * 20].view(torch.int64)
        size = 1
        while size < x.numel():
            size *= 2
        x = torch.cat([x, torch.zeros(size - x.numel(), dtype=torch.int64, device=dev)])
        while x.numel() > 1:
            h = x.numel() // 2
            x = x[:h] ^ x[h:]
        gpu_chk = f"{int(x.item()) & 0xFFFFFFFFFFFFFFFF:016x}"
        cpu = {"value": r["seqs_per_s"], "unit": "seqs/s", "cores": cores, "kind": "port",
               "sample": f"first {sample_n} of the {n} records, sha1 + strip + upper-casing, oracle/cpu_bench: OpenMP over {cores} threads, "
                         f"{'OpenSSL SHA-1 (SHA-NI)' if r['accel'] & 1 else 'portable C SHA-1'}, 3 steps; "
                         f"digest checksum cpu={r['checksum']} gpu={gpu_chk} equal={r['checksum'] == gpu_chk}"}

    if rank == 0:
        val = n * world / (ms_sha1 * 1e-3)
        line = {"metric": "sha1 seqs/s hashed, 250-bp reads", "value": val, "unit": "seqs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_sha1, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "gbps": val * L / 1e9,
                "config": {"workload": f"C2: {n} x {L} bp synthetic amplicon reads per GPU, sha1, whitespace strip + upper-casing on",
                           "flags": "strip+upper", "l2": f"inputs {total / 1e9:.2f} GB per step > 126 MB L2 (no reuse across steps)",
                           "parallelism": f"record-range sharding x{world}, no collective"},
                "roofline": roof_sha1,
                "xxh3": {"value": n * world / (ms_xxh3 * 1e-3), "unit": "seqs/s", "ms_per_step": ms_xxh3,
                         "gbps": n * world * L / (ms_xxh3 * 1e-3) / 1e9, "roofline": roof_xxh3, "clocks": clocks_x},
                "cpu_baseline": cpu, "e2e": e2e, "e2e_text": e2e_text, "gpu_launches": launches, "clocks": clocks}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
