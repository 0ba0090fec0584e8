"""BASELINE config C3 at 1/2/4/8 GPUs: 100 M synthetic 150-bp reads in total (strong scaling: the records are
sharded by contiguous range over the ranks), xxhash + xxh3 + murmur3 in one fused launch,
upper-casing + strip on.  Kernel-only (CUDA events, max over ranks) and end to end through the batch C ABI with
pinned host buffers.  Launch: python tools/bench_c3_multi.py   or   torchrun --nproc-per-node N ... (same flags
as bench.py)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from seqhasher_b200 import binding as B  # noqa: E402

TOTAL, L, ALGOS = 100_000_000, 150, ["xxhash", "xxh3", "murmur3"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--total", type=int, default=TOTAL)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    B.init(0)
    dev = torch.device("cuda", local)
    st = torch.cuda.current_stream().cuda_stream
    lo, hi = a.total * rank // world, a.total * (rank + 1) // world      # this rank's record range
    n = hi - lo
    total = n * L
    d_res = torch.empty(total + 16, dtype=torch.uint8, device=dev)
    B.synth_fill(d_res.data_ptr(), total, 0xC3 + rank, 20, 1, stream=st, device=local)
    d_offs = torch.arange(0, n + 1, dtype=torch.int64, device=dev) * L
    scratch = torch.empty(B.scratch_bytes(total, n), dtype=torch.uint8, device=dev)
    outs = [torch.empty(n * B.DIGEST_SIZES[B.algo_id(x)], dtype=torch.uint8, device=dev) for x in ALGOS]
    flags = B.FLAG_UPPER | B.FLAG_STRIP
    run = lambda: B.hash_device(d_res.data_ptr(), d_offs.data_ptr(), n, total, ALGOS, flags, [o.data_ptr() for o in outs],
                                d_scratch=scratch.data_ptr(), stream=st, device=local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    def mx(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(a.warmup):
        run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record(); torch.cuda.synchronize()
    barrier()
    ms = mx(e0.elapsed_time(e1)) / a.steps
    # end to end: 8 pinned batches per rank in flight
    parts, per = 8, (n + 7) // 8
    batches = []
    for p in range(parts):
        blo, bhi = p * per, min(n, (p + 1) * per)
        if bhi <= blo:
            break
        b = B.Batch((bhi - blo) * L, bhi - blo, device=local)
        torch.from_numpy(b.residues)[: (bhi - blo) * L].copy_(d_res[blo * L: bhi * L])
        b.offsets[: bhi - blo + 1] = np.arange(bhi - blo + 1, dtype=np.int64) * L
        b.n = bhi - blo
        batches.append(b)
    torch.cuda.synchronize()

    def step():
        for b in batches:
            b.submit(ALGOS, flags)
        for b in batches:
            b.wait()
    for _ in range(2):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    barrier()
    dt = mx(dt) / a.steps
    same = all(bool((torch.from_numpy(np.concatenate([b.digests(k) for b in batches])).reshape(-1) == outs[k].cpu()).all())
               for k in range(len(ALGOS)))
    if rank == 0:
        alg = a.total * (L + 8 + 32)
        print(json.dumps({"config": "C3: 100 M x 150 bp reads total, xxhash+xxh3+murmur3 fused, strip+upper", "n_gpus": world,
                          "scaling": "strong", "records_total": a.total, "kernel_ms_per_step": ms,
                          "kernel_reads_per_s": a.total / (ms * 1e-3), "kernel_algorithmic_gbps": alg / (ms * 1e-3) / 1e9,
                          "e2e_ms_per_step": dt * 1e3, "e2e_reads_per_s": a.total / dt, "e2e_gbps": a.total * L / dt / 1e9,
                          "e2e_matches_kernel_path": same}), flush=True)
    for b in batches:
        b.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
