"""Aggregate pinned host -> device copy bandwidth of the box with N GPUs copying at once: the roof of every end-to-end
number.  One process per GPU under torchrun (or a single process for N = 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_roof.py

Per rank and variant: shb_probe_h2d — 1 GiB of cudaHostAlloc memory (optionally write-combined; optionally with the process
bound to the GPU's NUMA node before the allocation and its first touch), plain cudaMemcpyAsync in 64 MiB pieces on one stream,
2 s, all ranks between barriers.  One JSON line per variant: aggregate GB/s, per-GPU min / max."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from seqhasher_b200 import binding as B  # noqa: E402
from seqhasher_b200 import numa  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B.init(0)
    pci = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                         capture_output=True, text=True).stdout.strip()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def gather(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    results = []
    bind_info = None
    for variant in ("default", "write_combined", "numa_bound", "numa_bound_write_combined"):
        if variant.startswith("numa_bound") and bind_info is None:
            bind_info = numa.bind_to_gpu_node(pci)
        barrier()
        g = B.probe_h2d(local, 1 << 30, 64 << 20, 2.0, "write_combined" in variant)
        barrier()
        per = gather(g)
        binds = gather(float(bind_info["numa_node"] if bind_info and bind_info["numa_node"] is not None else -1))
        if rank == 0:
            r = {"variant": variant, "n_gpus": world, "aggregate_gbps": sum(per), "per_gpu_min": min(per), "per_gpu_max": max(per),
                 "per_gpu": [round(x, 2) for x in per], "numa_nodes_on_host": numa.n_numa_nodes(),
                 "gpu_numa_nodes": [int(b) for b in binds] if variant.startswith("numa") else None, "host_cpus": os.cpu_count()}
            results.append(r)
            print(json.dumps(r), flush=True)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(results, open(os.path.join(ROOT, "gpurun_out", f"h2d_roof_n{world}.json"), "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
