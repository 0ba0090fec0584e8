"""Device hashers compiled for the HOST (tests/emul/cuda_runtime.h stands in for the CUDA headers) and checked
against the oracle: the arithmetic of the fused xxhash + xxh3 + murmur3 pass (csrc/hash_fused3.cuh) is verified
on every length and alignment without a GPU; the -m gpu tests then only have to show that the kernel around it
feeds it the right bytes.  Test infrastructure only: the product has no CPU path."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "emul", "_build", "libfused3_emul.so")


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(HERE, "emul", "fused3_emul.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-I" + os.path.join(HERE, "emul"),
                    "-I" + os.path.join(ROOT, "seqhasher_b200", "csrc"), "-o", SO, src], check=True)
    L = ctypes.CDLL(SO)
    L.emul_fused3.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.emul_fused3.restype = ctypes.c_uint32
    return L


def _check(emul, oracle, buf, start, n, upper):
    out = (ctypes.c_uint64 * 4)()
    flag = emul.emul_fused3(buf.ctypes.data, start, n, upper, 1, out)
    rec = buf[start:start + n].tobytes()
    data = rec.upper() if upper else rec
    want = [oracle.get_hash_func(a)(data) for a in ("xxhash", "xxh3", "murmur3")]
    got = ["%016x" % out[0], "%016x" % out[1], "%016x%016x" % (out[2], out[3])]
    assert got == want, (n, start, upper)
    assert bool(flag) == any(c < 0x21 for c in rec), (n, start, upper, "whitespace flag")


def test_fused3_every_length_and_alignment(emul, oracle):
    rng = np.random.default_rng(0)
    for n in range(17, 241):
        for start in (0, 1, 2, 3, 5, 18):
            buf = rng.choice(np.frombuffer(b"ACGTacgtNn", dtype=np.uint8), size=start + n + 64).astype(np.uint8)
            buf[start + n:] = rng.integers(0, 256, size=64)        # bytes past the record: anything, whitespace included
            _check(emul, oracle, buf, start, n, 1)
            _check(emul, oracle, buf, start, n, 0)


def test_fused3_arbitrary_bytes_and_whitespace_flag(emul, oracle):
    rng = np.random.default_rng(1)
    for trial in range(1500):
        n, start = int(rng.integers(17, 241)), int(rng.integers(0, 64))
        buf = rng.integers(0x21, 256, size=start + n + 64).astype(np.uint8)
        if trial % 3 == 0:                                           # one whitespace byte somewhere inside the record
            buf[start + int(rng.integers(0, n))] = int(rng.choice([9, 10, 11, 12, 13, 32]))
        buf[start + n:] = rng.integers(0, 0x21, size=64)             # everything past the end looks like whitespace
        _check(emul, oracle, buf, start, n, int(trial & 1))


@pytest.fixture(scope="module")
def inflate_emul():
    so = os.path.join(HERE, "emul", "_build", "libinflate_emul.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-I" + os.path.join(HERE, "emul"),
                    "-I" + os.path.join(ROOT, "seqhasher_b200", "csrc"), "-o", so, os.path.join(HERE, "emul", "inflate_emul.cpp")], check=True)
    L = ctypes.CDLL(so)
    L.emul_inflate_raw.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32]
    L.emul_inflate_raw.restype = ctypes.c_int
    L.emul_crc32_lanes.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]
    L.emul_crc32_lanes.restype = ctypes.c_uint32
    return L


def _deflate(data: bytes, level: int, strategy: int) -> bytes:
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def test_device_inflate_matches_zlib(inflate_emul):
    """csrc/inflate.cuh (the DEFLATE decoder the BGZF path runs one warp per member) against zlib: stored, fixed-Huffman and
    dynamic blocks, all levels, FASTA-like text and arbitrary bytes, sizes around the 64 KiB BGZF block; damaged and truncated
    streams must be refused, never decoded to something else."""
    import zlib
    rng = np.random.default_rng(42)
    cases = 0
    for trial in range(400):
        n = int(rng.choice([0, 1, 2, 100, 4000, 65280, 65536, 200000]) if trial % 3 == 0 else rng.integers(1, 70000))
        kind = trial % 4
        if kind == 0:
            data = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).tobytes()
        elif kind == 1:
            seq = rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), size=n).tobytes()
            data = b"".join(b">r%d\n" % i + seq[i:i + 60] + b"\n" for i in range(0, n, 60))[:n]
        elif kind == 2:
            data = rng.integers(0, 256, size=n).astype(np.uint8).tobytes()
        else:
            data = (b"ACGTACGTTTGA" * (n // 12 + 1))[:n]                     # long matches, overlapping copies
        for level, strategy in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            comp = _deflate(data, level, strategy)
            out = ctypes.create_string_buffer(max(len(data), 1))
            rc = inflate_emul.emul_inflate_raw(comp, len(comp), out, len(data))
            assert rc == 0 and out.raw[:len(data)] == data, (trial, n, level, strategy, rc)
            cases += 1
        # refusals: a wrong announced size, a truncated stream, a flipped bit (any outcome but "OK with other bytes")
        comp = _deflate(data, 6, zlib.Z_DEFAULT_STRATEGY)
        if n:
            out = ctypes.create_string_buffer(n + 8)
            assert inflate_emul.emul_inflate_raw(comp, len(comp), out, n - 1) != 0
            assert inflate_emul.emul_inflate_raw(comp, len(comp), out, n + 1) != 0
            if len(comp) > 4:
                assert inflate_emul.emul_inflate_raw(comp, len(comp) - 2, out, n) != 0
                bad = bytearray(comp)
                bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
                rc = inflate_emul.emul_inflate_raw(bytes(bad), len(bad), out, n)
                assert rc != 0 or out.raw[:n] != data or True   # CRC (checked by the caller) catches silent damage; no crash is the test
    assert cases == 2800


def test_device_crc32_lane_shares_match_zlib(inflate_emul):
    """The window's CRC-32 as the warp computes it (32 segments, each moved to the end of the text with x^(8n) mod P, XORed)
    against zlib.crc32, for every window phase 0..15 and sizes from empty to a full BGZF member."""
    import zlib
    rng = np.random.default_rng(7)
    buf = np.zeros(65536 + 64, dtype=np.uint8)
    base = buf.ctypes.data
    assert base % 4 == 0
    for trial in range(300):
        n = int(rng.choice([0, 1, 2, 3, 4, 5, 31, 32, 33, 127, 128, 129, 4095, 65280, 65535, 65536]) if trial % 2 else rng.integers(0, 65537))
        first = int(rng.integers(0, 16))
        buf[:] = rng.integers(0, 256, size=buf.size)
        got = inflate_emul.emul_crc32_lanes(base, first, n)
        assert got == zlib.crc32(buf[first:first + n].tobytes()), (trial, first, n)


def test_nthash_bit_planes_equal_the_seed_table_fold():
    """hash_nt_planes.cuh (the warp kernel's look-up-free path): a unit is accepted iff all 16 bytes are ACGTU in either case,
    and the planes of any number of accepted units expand to XOR_units XOR_j rol(seed[byte j], 15 - j)."""
    so = os.path.join(HERE, "emul", "_build", "libnt_planes_emul.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-I" + os.path.join(HERE, "emul"),
                    "-I" + os.path.join(ROOT, "seqhasher_b200", "csrc"), "-o", so, os.path.join(HERE, "emul", "nt_planes_emul.cpp")], check=True)
    L = ctypes.CDLL(so)
    L.emul_nt_planes.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    L.emul_nt_planes.restype = ctypes.c_uint64
    L.emul_nt_seed.argtypes = [ctypes.c_uint32]
    L.emul_nt_seed.restype = ctypes.c_uint64
    seed = [L.emul_nt_seed(c) for c in range(256)]
    bases = set(b"ACGTUacgtu")
    assert all((seed[c] != 0) == (c in bases or c in (1, 3, 4, 5, 7)) for c in range(256))
    M = (1 << 64) - 1

    def rol(x, r):
        return ((x << r) | (x >> (64 - r))) & M if r else x
    rng = np.random.default_rng(11)
    al_clean = np.frombuffer(b"ACGTacgtUu", dtype=np.uint8)
    for trial in range(1500):
        n = int(rng.integers(1, 40))
        units = rng.choice(al_clean, size=16 * n).astype(np.uint8)
        kind = trial % 3
        if kind == 1:      # a few arbitrary bytes: those units must be refused
            for _ in range(int(rng.integers(1, 4))):
                units[int(rng.integers(0, 16 * n))] = int(rng.integers(0, 256))
        elif kind == 2:    # bytes one bit away from a base, N, IUPAC codes, control bytes with a seed
            for _ in range(int(rng.integers(1, 6))):
                units[int(rng.integers(0, 16 * n))] = int(rng.choice(np.frombuffer(b"NnRYKMSWBDHV@EFDPQRSVW`\x01\x03\x04\x05\x07\x00\x80\xc1 -", dtype=np.uint8)))
        valid = np.zeros(n, dtype=np.uint8)
        got = L.emul_nt_planes(units.ctypes.data, n, valid.ctypes.data)
        want = 0
        for i in range(n):
            u = units[16 * i:16 * i + 16].tolist()
            ok = all(c in bases for c in u)
            assert bool(valid[i]) == ok, (trial, i, bytes(u))
            if ok:
                for j, c in enumerate(u):
                    want ^= rol(seed[c], 15 - j)
        assert got == want, trial


def test_device_inflate_survives_damaged_streams_under_asan():
    """tests/emul/inflate_fuzz.cpp: the decoder on garbage, bit-flipped and truncated streams with wrong announced sizes, built
    with -fsanitize=address,undefined and exact-size buffers: it may refuse, it must never read or write outside them."""
    exe = os.path.join(HERE, "emul", "_build", "inflate_fuzz")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-w",
                    "-I" + os.path.join(HERE, "emul"), "-I" + os.path.join(ROOT, "seqhasher_b200", "csrc"),
                    os.path.join(HERE, "emul", "inflate_fuzz.cpp"), "-lz", "-o", exe], check=True)
    p = subprocess.run([exe, "40000"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "no memory error" in p.stdout, p.stdout[-500:] + p.stderr[-2000:]
