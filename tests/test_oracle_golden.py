"""Pins the CPU oracle (oracle/) against every golden vector the reference's own tests hold for the hash
path (seqhasher_test.go:318-329, :265, :278, :290, :834; test/test2_binary.sh), against independent
libraries on every block/stripe/chunk boundary, and against the published KATs of the un-vendored modules."""
import hashlib
import random

import pytest

import golden_vectors as G
from synth import BOUNDARY_LENGTHS


@pytest.mark.parametrize("algo", sorted(G.ACTG))
def test_reference_goldens_actg(oracle, algo):
    assert oracle.get_hash_func(algo)(b"ACTG") == G.ACTG[algo]


@pytest.mark.parametrize("algo,data,want,cite", G.EXTRA, ids=[f"{e[0]}-{e[3]}" for e in G.EXTRA])
def test_reference_goldens_extra(oracle, algo, data, want, cite):
    got = oracle.get_hash_func(algo)(data)
    assert got.startswith(want) if len(want) < 16 else got == want


def test_empty_sequence_rule_and_default_arm(oracle):
    # seqhasher.go:292-295: "" for empty data; seqhasher.go:344-346: unknown names hash as SHA-1
    for a in oracle.ALGO_NAMES:
        assert oracle.get_hash_func(a)(b"") == ""
    assert oracle.get_hash_func(" md5")(b"ACTG") == G.ACTG["sha1"]
    assert oracle.get_hash_func("nonsense")(b"ACTG") == G.ACTG["sha1"]


def test_against_independent_libraries(oracle):
    xxhash = pytest.importorskip("xxhash")
    blake3 = pytest.importorskip("blake3")
    rnd = random.Random(20261019)
    for n in BOUNDARY_LENGTHS:
        if n == 0:
            continue
        d = rnd.randbytes(n)
        assert oracle.get_hash_func("sha1")(d) == hashlib.sha1(d).hexdigest(), n
        assert oracle.get_hash_func("md5")(d) == hashlib.md5(d).hexdigest(), n
        assert oracle.get_hash_func("sha3")(d) == hashlib.sha3_512(d).hexdigest(), n
        assert oracle.get_hash_func("xxhash")(d) == xxhash.xxh64(d, seed=0).hexdigest(), n
        assert oracle.get_hash_func("xxh3")(d) == xxhash.xxh3_64(d, seed=0).hexdigest(), n
        assert oracle.get_hash_func("blake3")(d) == blake3.blake3(d).hexdigest(), n


def test_published_kats(oracle):
    for d, want in G.MURMUR3_KATS.items():
        assert oracle.get_hash_func("murmur3")(d) == want
    for d, want in G.KT128_KATS:
        assert oracle.hash_raw("k12", d)[:32].hex() == want
    assert oracle.get_hash_func("cityhash")(b"abc") == G.CITY128_ABC
    L = oracle.lib()
    assert L.sho_cityhash64(b"abc", 3) == G.CITY64_ABC
    assert L.sho_rapidhash_upstream_v1(b"hello world", 11) == G.RAPIDHASH_V1_HELLO_WORLD


def test_nthash_closed_form(oracle):
    # App. A.8: iterative h = rol(h,1)^seed[c] == XOR_i rol(seed[s_i], (L-1-i) mod 64); non-ACGT contribute 0
    seeds = {65: 0x3c8bfbb395c60474, 67: 0x3193c18562a02b4c, 71: 0x20323ed082572324, 84: 0x295549f54be24456}
    rnd = random.Random(7)
    for n in (1, 5, 63, 64, 65, 129, 5000):
        s = bytes(rnd.choice(b"ACGTN") for _ in range(n))
        h = 0
        for i, c in enumerate(s):
            r = (n - 1 - i) % 64
            v = seeds.get(c, 0)
            h ^= ((v << r) | (v >> (64 - r))) & 0xFFFFFFFFFFFFFFFF if r else v
        assert oracle.get_hash_func("nthash")(s) == f"{h:016x}"


def test_rapidhash32_is_fold_of_rapidhash(oracle):
    rnd = random.Random(3)
    for n in (1, 3, 4, 16, 17, 48, 49, 150, 250, 1000):
        d = rnd.randbytes(n)
        h = int(oracle.get_hash_func("rapidhash")(d), 16)
        assert oracle.get_hash_func("rapidhash32")(d) == f"{(h & 0xFFFFFFFF) ^ (h >> 32):08x}"


def test_normalise_matches_go_ascii_semantics(oracle):
    # bytes.Fields strips \t \n \v \f \r and space; bytes.ToUpper maps a-z only (seqhasher.go:246-251)
    raw = b"Ac \tT\ng\x0b\x0c\r N-*z{`@[1"
    assert oracle.normalise(raw, 3) == b"ACTGN-*Z{`@[1"
    assert oracle.normalise(raw, 2) == b"AcTgN-*z{`@[1"
    assert oracle.normalise(raw, 1) == raw.upper()
    assert oracle.normalise(b" \t\n", 3) == b""


def test_batch_matches_one_shot_and_threads_agree(oracle):
    import numpy as np
    from synth import csr_from_lengths
    lens = [0, 1, 4, 150, 250, 0, 1025, 8192, 77]
    res, offs = csr_from_lengths(lens, seed=11, lower_permille=300, n_permille=10)
    algos = oracle.ALGO_NAMES
    d1, l1 = oracle.batch(res, offs, algos, flags=3, threads=1)
    d4, l4 = oracle.batch(res, offs, algos, flags=3, threads=4, accel=True)
    assert (l1 == l4).all() and (l1 == np.array(lens)).all()
    for k, a in enumerate(algos):
        assert (d1[k] == d4[k]).all(), a
        for i, n in enumerate(lens):
            rec = oracle.normalise(res[offs[i]:offs[i + 1]].tobytes(), 3)
            want = oracle.get_hash_func(a)(rec)
            got = d1[k][i].tobytes().hex() if n else ""
            assert got == want, (a, i)


def test_unpinned_paths_three_way_agreement(oracle):
    # cityhash >= 16 B and rapidhash > 16 B have no reference vector: an independent pure-Python restatement
    # (tests/pyref.py) must agree with the C oracle on every length class (the CUDA code is the third, on GPU)
    import pyref
    rnd = random.Random(11)
    assert pyref.cityhash128_hex(b"ACTG") == G.ACTG["cityhash"] and pyref.rapidhash_hex(b"ACTG") == G.ACTG["rapidhash"]
    assert pyref.cityhash128_hex(b"abc") == G.CITY128_ABC
    assert int(pyref.rapidhash_hex(b"hello world", final_secret=1), 16) == G.RAPIDHASH_V1_HELLO_WORLD
    for n in list(range(0, 200)) + [239, 240, 241, 255, 256, 257, 383, 384, 385, 1000, 4097]:
        d = rnd.randbytes(n)
        if n:
            assert oracle.get_hash_func("cityhash")(d) == pyref.cityhash128_hex(d), n
            assert oracle.get_hash_func("rapidhash")(d) == pyref.rapidhash_hex(d), n
