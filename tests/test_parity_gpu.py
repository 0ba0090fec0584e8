"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact, on the reference's golden
inputs, on every block/stripe/chunk boundary length, on ragged and empty records, and in both case modes."""
import numpy as np
import pytest

import golden_vectors as G
from synth import BOUNDARY_LENGTHS, csr_from_lengths, csr_from_records

pytestmark = pytest.mark.gpu

ALL = ["sha1", "sha3", "md5", "xxhash", "xxh3", "cityhash", "murmur3", "nthash", "blake3", "k12", "rapidhash",
       "rapidhash32"]


@pytest.fixture(scope="module")
def shb():
    from seqhasher_b200 import binding
    binding.init(1)
    return binding


def _hex(d, lens):
    return [row.tobytes().hex() if n else "" for row, n in zip(d, lens)]


def test_reference_goldens_through_cabi(shb):
    recs = [b"ACTG", b"TGCA", b"AAAA", b"aaaa", b"ACTGINVALID", b"actg"]
    res, offs = csr_from_records(recs)
    outs, lens = shb.hash_batch(res, offs, ALL, flags=0)
    for k, a in enumerate(ALL):
        assert _hex(outs[k], lens)[0] == G.ACTG[a], a
    for a, data, want, cite in G.EXTRA:
        if data in recs:
            got = _hex(outs[ALL.index(a)], lens)[recs.index(data)]
            assert got.startswith(want), (a, cite)
    # default mode upper-cases (seqhasher.go:249-251): "actg" and "aaaa" hash like their upper-case forms
    outs_u, lens_u = shb.hash_batch(res, offs, ALL, flags=shb.FLAG_UPPER)
    for k, a in enumerate(ALL):
        h = _hex(outs_u[k], lens_u)
        assert h[5] == G.ACTG[a] and h[3] == h[2], a


def test_long_sequence_goldens(shb):
    # test/test2_binary.sh:148-169
    res, offs = csr_from_records([b"ACTG" * 25000, b"A" * 100000])
    outs, lens = shb.hash_batch(res, offs, ["sha1"], flags=shb.FLAG_UPPER)
    assert _hex(outs[0], lens) == ["3200abf5034212fb5d90fa242fc1cea5e0f3dd00", "e27b41eb7f7370b36f1b2fa63f05652f52423277"]


@pytest.mark.parametrize("flags", [0, 1], ids=["casesensitive", "upper"])
def test_boundary_lengths_all_algorithms(shb, oracle, flags):
    res, offs = csr_from_lengths(BOUNDARY_LENGTHS, seed=0xB200, lower_permille=250, n_permille=20)
    want, wl = oracle.batch(res, offs, ALL, flags=flags, threads=8)
    got, gl = shb.hash_batch(res, offs, ALL, flags=flags)
    assert (gl == wl).all()
    for k, a in enumerate(ALL):
        bad = np.nonzero((got[k] != want[k]).any(axis=1))[0]
        assert bad.size == 0, f"{a}: mismatch at lengths {[BOUNDARY_LENGTHS[i] for i in bad[:8]]}"


def test_arbitrary_bytes_and_alignments(shb, oracle):
    # non-DNA bytes are hashed as-is (seqhasher_test.go:817-834); every start alignment 0..15 occurs
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 700, size=4000)
    offs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    res = rng.integers(0, 256, size=int(offs[-1]), dtype=np.uint8)
    for flags in (0, 1):
        want, _ = oracle.batch(res, offs, ALL, flags=flags, threads=8)
        got, _ = shb.hash_batch(res, offs, ALL, flags=flags)
        for k, a in enumerate(ALL):
            assert (got[k] == want[k]).all(), (a, flags)


def test_strip_and_emit(shb, oracle):
    # whitespace strip (seqhasher.go:246) incl. records that become empty, plus the normalised sequences
    recs = [b"AC TG ", b"AC\tTG", b"AC\nTG", b"actg", b" \t\n", b"", b"A" * 5000 + b" " + b"c" * 4000, b"\x0b\x0cN\r"]
    rng = np.random.default_rng(9)
    for _ in range(300):
        n = int(rng.integers(0, 9000))
        r = rng.choice(np.frombuffer(b"ACGTacgtN \t\n\r", dtype=np.uint8), size=n)
        recs.append(r.tobytes())
    res, offs = csr_from_records(recs)
    for flags in (2, 3):
        want, wl = oracle.batch(res, offs, ALL, flags=flags, threads=8)
        got, gl, (nres, noffs) = shb.hash_batch(res, offs, ALL, flags=flags, emit=True)
        assert (gl == wl).all()
        for k, a in enumerate(ALL):
            assert (got[k] == want[k]).all(), (a, flags)
        for i in (0, 1, 2, 3, 4, 5, 6, 7, 50, 307):
            assert nres[noffs[i]:noffs[i + 1]].tobytes() == oracle.normalise(recs[i], flags)
    outs, lens = shb.hash_batch(*csr_from_records(recs[:6]), ["sha1"], flags=3)
    assert _hex(outs[0], lens) == [G.ACTG["sha1"]] * 4 + ["", ""]


def test_duplicate_algos_and_empty_batch(shb):
    res, offs = csr_from_records([b"ACTG", b"TGCA"])
    outs, lens = shb.hash_batch(res, offs, ["sha1", "md5", "sha1"], flags=0)
    assert (outs[0] == outs[2]).all() and _hex(outs[1], lens)[1] == "5c15f97a88433c48f8bf76745d9da437"
    outs, lens = shb.hash_batch(np.zeros(0, np.uint8), np.zeros(1, np.int64), ["sha1"], flags=3)
    assert outs[0].shape == (0, 20)


def test_synth_generator_matches_numpy_twin(shb):
    import torch
    from synth import synth_bytes
    total = 100003
    t = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
    shb.synth_fill(t.data_ptr(), total, seed=0xC2, lower_permille=50, n_permille=1,
                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (t[:total].cpu().numpy() == synth_bytes(total, 0xC2, 50, 1)).all()


def test_device_resident_path_fixed_length_reads(shb, oracle):
    # C2 / C3 shapes (250 bp and 150 bp fixed-length reads) at a size the oracle finishes in seconds
    import torch
    from synth import synth_bytes
    st = torch.cuda.current_stream().cuda_stream
    for L, algos in ((250, ["sha1", "xxh3"]), (150, ["xxhash", "xxh3", "murmur3"])):
        n = 200_000
        total = n * L
        d_res = torch.empty(total + 16, dtype=torch.uint8, device="cuda")
        shb.synth_fill(d_res.data_ptr(), total, seed=0xC2, lower_permille=30, stream=st)
        d_offs = torch.arange(0, n + 1, dtype=torch.int64, device="cuda") * L
        outs = [torch.empty(n * shb.DIGEST_SIZES[shb.algo_id(a)], dtype=torch.uint8, device="cuda") for a in algos]
        shb.hash_device(d_res.data_ptr(), d_offs.data_ptr(), n, total, algos, shb.FLAG_UPPER,
                        [o.data_ptr() for o in outs], stream=st)
        torch.cuda.synchronize()
        res = synth_bytes(total, 0xC2, 30)
        want, _ = oracle.batch(res, d_offs.cpu().numpy(), algos, flags=1, threads=8, accel=True)
        for k, a in enumerate(algos):
            assert (outs[k].cpu().numpy().reshape(n, -1) == want[k]).all(), (L, a)


def test_tile_path_ragged_tiles_and_fallback(shb, oracle):
    # mean length is short (shared-memory tile path) but some 128-record tiles hold a 60 kB record and do
    # not fit the tile budget: those CTAs must fall back to direct global reads; also empty records, a
    # tile made only of empty records, and a last partial tile
    rng = np.random.default_rng(77)
    lens = rng.integers(0, 260, size=128 * 40 + 17)
    lens[300] = 60_000
    lens[2000] = 75_001
    lens[128 * 5: 128 * 6] = 0
    res, offs = csr_from_lengths(lens, seed=0x7153, lower_permille=400, n_permille=30)
    for flags in (0, 1):
        want, _ = oracle.batch(res, offs, ALL, flags=flags, threads=8)
        got, _ = shb.hash_batch(res, offs, ALL, flags=flags)
        for k, a in enumerate(ALL):
            bad = np.nonzero((got[k] != want[k]).any(axis=1))[0]
            assert bad.size == 0, f"{a} flags={flags}: mismatch at records {bad[:8]} lens {lens[bad[:8]]}"


@pytest.mark.parametrize("L", [1, 31, 150, 250, 251, 640])
def test_tile_path_fixed_lengths_fused_sets(shb, oracle, L):
    n = 128 * 30 + 5
    res, offs = csr_from_lengths([L] * n, seed=L, lower_permille=100, n_permille=5)
    sets = [["xxhash", "xxh3", "murmur3"], ["nthash", "cityhash"], ["rapidhash", "rapidhash32", "xxh3", "sha1", "md5"],
            ["blake3", "k12", "sha3"]]
    for algos in sets:
        want, _ = oracle.batch(res, offs, algos, flags=1, threads=8)
        got, _ = shb.hash_batch(res, offs, algos, flags=1)
        for k, a in enumerate(algos):
            assert (got[k] == want[k]).all(), (L, a)


def test_long_record_path(shb, oracle):
    # mean length >= 2 KiB switches ntHash to warp-per-record and BLAKE3 to chunk-parallel + tree merge
    rng = np.random.default_rng(4)
    lens = [0, 1, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 3073, 4096, 5000, 31744, 32768, 32769, 33000, 65536, 65537,
            70000, 100000, 131072, 0, 64, 7 * 1024 + 1] + list(rng.integers(2000, 60000, size=300))
    res, offs = csr_from_lengths(lens, seed=0xC4, lower_permille=100, n_permille=10)
    algos = ["blake3", "nthash", "xxh3", "sha1", "k12"]
    for flags in (0, 1, 3):
        want, wl = oracle.batch(res, offs, algos, flags=flags, threads=8)
        got, gl = shb.hash_batch(res, offs, algos, flags=flags)
        assert (gl == wl).all()
        for k, a in enumerate(algos):
            bad = np.nonzero((got[k] != want[k]).any(axis=1))[0]
            assert bad.size == 0, f"{a} flags={flags}: mismatch at lengths {[lens[i] for i in bad[:8]]}"
    # arbitrary bytes (ntHash's non-ACGT -> 0 rule, every alignment) on the long path
    lens2 = rng.integers(1500, 9000, size=400)
    offs2 = np.zeros(len(lens2) + 1, dtype=np.int64)
    np.cumsum(lens2, out=offs2[1:])
    res2 = rng.integers(0, 256, size=int(offs2[-1]), dtype=np.uint8)
    want, _ = oracle.batch(res2, offs2, ["nthash", "blake3"], flags=1, threads=8)
    got, _ = shb.hash_batch(res2, offs2, ["nthash", "blake3"], flags=1)
    assert (got[0] == want[0]).all() and (got[1] == want[1]).all()


def test_fused_strip_in_tile_path(shb, oracle):
    # short records (tile path): strip fused into the staged tile; tiles without whitespace take the fast
    # path, tiles with whitespace compact in shared memory, an oversize tile compacts through scratch
    rng = np.random.default_rng(12)
    recs = []
    for k in range(128 * 12 + 3):
        n = int(rng.integers(0, 300))
        alphabet = b"ACGTacgtN" if (k // 128) % 2 == 0 else b"ACGTacgtN \t\n\r\x0b\x0c"
        recs.append(rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=n).tobytes())
    recs[700] = (b"AC GT\n" * 12000)          # 72 kB record in a short-mean batch: oversize tile
    recs[5] = b" \t \n"                        # strips to empty
    res, offs = csr_from_records(recs)
    assert int(offs[-1]) // len(recs) <= 640
    for flags in (2, 3):
        want, wl = oracle.batch(res, offs, ALL, flags=flags, threads=8)
        got, gl = shb.hash_batch(res, offs, ALL, flags=flags)
        assert (gl == wl).all()
        for k, a in enumerate(ALL):
            bad = np.nonzero((got[k] != want[k]).any(axis=1))[0]
            assert bad.size == 0, f"{a} flags={flags}: mismatch at records {bad[:8]}"


@pytest.mark.parametrize("extra", [0, "direct"], ids=["tile", "direct"])
def test_single_whitespace_byte_at_every_position(shb, oracle, extra):
    """The speculative strip only looks again at records in which the readers SAW a byte < 0x21: one whitespace byte
    of each kind, alone, at every offset (i.e. every byte lane of the 32-bit words, at every record alignment) must be
    seen.  (A wrong SWAR constant once hid a lone 0x20 in byte lane 1.)"""
    rng = np.random.default_rng(3)
    recs = []
    for ws in b" \t\n\x0b\x0c\r":
        for L in (5, 8, 16, 17, 33, 64, 65, 130, 250):
            for pos in range(min(L, 24)):
                body = bytearray(rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=L).tobytes())
                body[pos] = ws
                recs.append(bytes(body))
                body[L - 1 - pos % L] = ws          # a second one counted from the end
                recs.append(bytes(body))
    res, offs = csr_from_records(recs)
    for flags in (2, 3):
        f = flags | (shb.FLAG_FORCE_DIRECT if extra else 0)
        want, wl = oracle.batch(res, offs, ALL, flags=flags, threads=8)
        got, gl = shb.hash_batch(res, offs, ALL, flags=f)
        assert (gl == wl).all(), f"lengths differ at records {np.nonzero(gl != wl)[0][:8]}"
        for k, a in enumerate(ALL):
            bad = np.nonzero((got[k] != want[k]).any(axis=1))[0]
            assert bad.size == 0, f"{a} flags={flags}: mismatch at records {bad[:8]}"


def test_hash_sharded_single_process(shb, oracle):
    # the in-process multi-device driver (one shb_batch per device); on a 1-GPU box this is one shard
    import torch
    from seqhasher_b200.shard import hash_sharded
    n_gpus = min(torch.cuda.device_count(), 2)
    res, offs = csr_from_lengths(np.random.default_rng(3).integers(0, 600, size=5000), seed=21, lower_permille=50)
    want, wl = oracle.batch(res, offs, ["sha1", "xxh3"], flags=3, threads=4)
    got, gl = hash_sharded(res, offs, ["sha1", "xxh3"], 3, n_gpus)
    assert (gl == wl).all() and (got[0] == want[0]).all() and (got[1] == want[1]).all()


def test_blake3_nthash_one_pass_on_long_records(shb, oracle):
    """`--hash blake3,nthash` on long records (config C4): ntHash is folded into the BLAKE3 chunk pass.  Against the oracle
    and against the two-pass path (SHB_FLAG_NO_FUSED_B3NT), both case modes; chunk-boundary lengths, empty and
    single-chunk records mixed in (mean stays above the 10 KiB threshold of the chunk-parallel path)."""
    import numpy as np
    rng = np.random.default_rng(77)
    lens = np.concatenate([[0, 1, 63, 64, 65, 1023, 1024, 1025, 2047, 2048, 2049, 65536, 65537, 40000 * 3 + 1],
                           rng.integers(10_000, 50_000, size=60)])
    offs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    res = rng.choice(np.frombuffer(b"ACGTNacgtnU*-", dtype=np.uint8), size=int(offs[-1]))
    for flags in (0, shb.FLAG_UPPER):
        want, _ = oracle.batch(res, offs, ["blake3", "nthash"], flags, threads=4)
        got, _ = shb.hash_batch(res, offs, ["blake3", "nthash"], flags)
        two, _ = shb.hash_batch(res, offs, ["blake3", "nthash"], flags | shb.FLAG_NO_FUSED_B3NT)
        for k in range(2):
            assert (got[k] == want[k]).all(), (flags, k)
            assert (two[k] == want[k]).all(), (flags, k)
        # order of the hash list must not matter, nor other algorithms in between
        got3, _ = shb.hash_batch(res, offs, ["nthash", "xxh3", "blake3"], flags)
        assert (got3[0] == want[1]).all() and (got3[2] == want[0]).all()
