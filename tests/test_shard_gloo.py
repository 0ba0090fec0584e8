"""Multi-rank host logic on CPU: world_size-2 gloo run of the record-range sharding — each
rank hashes its byte-balanced range (with the oracle standing in for its GPU), rank order = input order."""
import os
import socket
import subprocess
import sys

import numpy as np

from seqhasher_b200.shard import shard_ranges, slice_csr
from synth import csr_from_lengths

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["SHB_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SHB_ROOT"], "tests"))
from oracle import oracle as O
from seqhasher_b200.shard import shard_ranges, slice_csr
from synth import csr_from_lengths
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(99)
lens = rng.integers(0, 4000, size=3001)
res, offs = csr_from_lengths(lens, seed=5, lower_permille=100)
lo, hi = shard_ranges(offs, world)[rank]
r, o = slice_csr(res, offs, lo, hi)
(d,), _ = O.batch(r, o, ["sha1"], 1)
# "gather back over PCIe": here an all_gather of (count, padded digests); order = rank order
n_max = len(lens)
pad = np.zeros((n_max, 20), dtype=np.uint8); pad[: hi - lo] = d
bufs = [torch.zeros((n_max, 20), dtype=torch.uint8) for _ in range(world)]
cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(bufs, torch.from_numpy(pad)); dist.all_gather(cnts, torch.tensor([hi - lo]))
if rank == 0:
    got = np.concatenate([b.numpy()[: int(c)] for b, c in zip(bufs, cnts)])
    (want,), _ = O.batch(res, offs, ["sha1"], 1)
    assert got.shape == want.shape and (got == want).all()
    t = torch.tensor([float(rank + 1)]); 
    print("SHARD-OK", got.shape[0])
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)      # the max-over-ranks reduction bench.py uses for timings
assert t.item() == world
dist.barrier(); dist.destroy_process_group()
'''


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(1)
    lens = rng.integers(0, 5000, size=10_000)
    _, offs = csr_from_lengths([0], 0)
    offs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    for world in (1, 2, 3, 4, 8):
        rs = shard_ranges(offs, world)
        assert rs[0][0] == 0 and rs[-1][1] == len(lens)
        assert all(rs[k][1] == rs[k + 1][0] for k in range(world - 1))
        sizes = [int(offs[hi] - offs[lo]) for lo, hi in rs]
        assert max(sizes) - min(sizes) <= 5000 * 2
    # degenerate inputs: fewer records than ranks, all-empty records
    assert shard_ranges(np.array([0, 10]), 4)[-1] == (1, 1) or sum(hi - lo for lo, hi in shard_ranges(np.array([0, 10]), 4)) == 1
    rs = shard_ranges(np.zeros(6, dtype=np.int64), 2)
    assert rs[0][0] == 0 and rs[-1][1] == 5 and rs[0][1] == rs[1][0]


def test_slice_csr_rebases_offsets():
    res, offs = csr_from_lengths([3, 0, 5, 2], seed=1)
    r, o = slice_csr(res, offs, 1, 3)
    assert o.tolist() == [0, 0, 5] and r.tobytes() == res[3:8].tobytes()


def test_two_rank_gloo_run(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "shard_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SHB_ROOT=ROOT, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    assert "SHARD-OK 3001" in p.stdout
