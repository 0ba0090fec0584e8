"""Parity at BASELINE.json's full sizes through size-independent properties: (a) the shared-memory tile kernels
and the direct global-memory kernels are two independent data paths and must agree on every digest of the
whole workload; (b) a seeded sample of the records is checked against the CPU oracle; (c) a second run is
bit-identical.  Sizes: C2 = 10 M x 250 bp (sha1, xxh3), C3 = 100 M x 150 bp scaled to 20 M for test time
(xxhash, xxh3, murmur3 fused), C4 = 10-50 kb reads scaled to 20 k records (blake3, nthash, xxh3: long-record
kernels vs thread-per-record kernels)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shb():
    from seqhasher_b200 import binding
    binding.init(1)
    return binding


def _run(shb, torch, d_res, d_offs, n, total, algos, flags, scratch):
    outs = [torch.empty(n * shb.DIGEST_SIZES[shb.algo_id(a)], dtype=torch.uint8, device="cuda") for a in algos]
    shb.hash_device(d_res.data_ptr(), d_offs.data_ptr(), n, total, algos, flags, [o.data_ptr() for o in outs],
                    d_scratch=scratch.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return outs


def _check(shb, oracle, n, lens, algos, flags, seed, alt_flag, sample=20000):
    import torch
    from synth import synth_bytes
    st = torch.cuda.current_stream().cuda_stream
    if isinstance(lens, int):
        d_offs = torch.arange(0, n + 1, dtype=torch.int64, device="cuda") * lens
    else:
        g = torch.Generator(device="cuda"); g.manual_seed(seed)
        l = torch.randint(lens[0], lens[1] + 1, (n,), generator=g, device="cuda", dtype=torch.int64)
        d_offs = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        torch.cumsum(l, 0, out=d_offs[1:])
    total = int(d_offs[-1].item())
    d_res = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
    shb.synth_fill(d_res.data_ptr(), total, seed, 40, 2, stream=st)
    scratch = torch.empty(shb.scratch_bytes(total, n), dtype=torch.uint8, device="cuda")
    a = _run(shb, torch, d_res, d_offs, n, total, algos, flags, scratch)
    b = _run(shb, torch, d_res, d_offs, n, total, algos, flags | alt_flag, scratch)      # the other kernel family
    c = _run(shb, torch, d_res, d_offs, n, total, algos, flags, scratch)                 # determinism
    for k, name in enumerate(algos):
        assert torch.equal(a[k], b[k]), f"{name}: the two kernel families disagree"
        assert torch.equal(a[k], c[k]), f"{name}: second run differs"
    # seeded sample against the oracle (records regenerated on the host from the same generator)
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(n, size=min(sample, n), replace=False))
    offs = d_offs.cpu().numpy()
    recs = [synth_bytes(int(offs[i + 1] - offs[i]), seed, 40, 2, start=int(offs[i])) for i in idx[:2000 if not isinstance(lens, int) else sample]]
    idx = idx[: len(recs)]
    so = np.zeros(len(recs) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in recs], out=so[1:])
    want, _ = oracle.batch(np.concatenate(recs), so, algos, flags & 3, threads=8, accel=True)
    for k, name in enumerate(algos):
        ds = shb.DIGEST_SIZES[shb.algo_id(name)]
        got = a[k].view(n, ds)[torch.from_numpy(idx).cuda()].cpu().numpy()
        assert (got == want[k]).all(), f"{name}: sample differs from the oracle"


def test_c2_full_size(shb, oracle):
    _check(shb, oracle, 10_000_000, 250, ["sha1", "xxh3"], 3, 0xC2, shb.FLAG_FORCE_DIRECT)


def test_c3_20m_reads_fused(shb, oracle):
    _check(shb, oracle, 20_000_000, 150, ["xxhash", "xxh3", "murmur3"], 1, 0xC3, shb.FLAG_FORCE_DIRECT)


def test_c4_long_reads(shb, oracle):
    _check(shb, oracle, 20_000, (10_000, 50_000), ["blake3", "nthash", "xxh3", "k12"], 0, 0xC4, shb.FLAG_NO_LONG_PATH)


def test_very_long_reads_leaf_and_chunk_parallel_kernels(shb, oracle):
    # mean 85 kB: above both crossovers (BLAKE3 chunk-parallel from 10 kB, K12 leaf-parallel from 48 kB mean)
    _check(shb, oracle, 3_000, (50_000, 120_000), ["k12", "blake3", "xxh3", "nthash"], 1, 0xC6, shb.FLAG_NO_LONG_PATH, sample=300)


def test_window_local_order_on_long_reads(shb, oracle):
    # mean 25 kB, n >= 1024: sha1 / md5 / murmur3 run hash_direct_local_kernel (every CTA sorts its own 128 records);
    # SHB_FLAG_NO_LENGTH_SORT is the plain input-order kernel
    _check(shb, oracle, 4_000, (10_000, 40_000), ["sha1", "md5", "murmur3", "xxhash"], 1, 0xC7, shb.FLAG_NO_LENGTH_SORT, sample=400)


def test_c5_variable_all_algorithms(shb, oracle):
    from seqhasher_b200.binding import ALGO_NAMES
    _check(shb, oracle, 300_000, (30, 3000), ALGO_NAMES, 3, 0xC5, shb.FLAG_NO_LENGTH_SORT, sample=3000)


@pytest.mark.parametrize("lens", [(800, 1200), 1000, (2000, 8000)], ids=["narrow", "uniform", "2k-8k"])
def test_visiting_order_policy_never_changes_a_digest(shb, oracle, lens):
    """The direct kernels visit the records in length-sorted or input order depending on a divergence estimate made
    on the device (narrow spread: sorted for sha1/sha3/blake3/k12 only; uniform: identity permutation; 2-8 kB: the
    multiply-based hashes keep input order).  Whatever the order, digests land at the record's own index."""
    from seqhasher_b200.binding import ALGO_NAMES
    _check(shb, oracle, 120_000, lens, ALGO_NAMES, 1, 0x50 + (lens if isinstance(lens, int) else lens[0]) % 7, shb.FLAG_NO_LENGTH_SORT,
           sample=1500)
