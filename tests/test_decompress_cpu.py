"""Decompression feeders of the compiled host (seqhasher_b200/csrc/host/decompress.hpp), no GPU needed.

`shb-zcat` runs exactly the input stack the compiled CLI puts in front of the GPU record loop.  The parallel
gzip reader (pgzip.hpp + inflate_markers.hpp) must reproduce zlib's output bit for bit on every kind of
deflate stream, verify CRC-32/ISIZE like the reference's gzip reader, and fail on the same inputs
(seqhasher_test.go:355-360 reads the .gz/.zst/.xz/.bz2 twins of test/test.fasta through this layer)."""
import bz2
import lzma
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZCAT = os.path.join(ROOT, "seqhasher_b200", "bin", "shb-zcat")
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")


def fasta(nbytes, seed, wrap=60):
    r = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    out, tot, i = [], 0, 0
    while tot < nbytes:
        L = int(r.integers(30, 3000))
        seq = alpha[r.integers(0, 4 if i % 7 else 9, size=L)].tobytes()
        rec = b">s%d\n" % i + b"\n".join(seq[j:j + wrap] for j in range(0, L, wrap)) + b"\n"
        out.append(rec)
        tot += len(rec)
        i += 1
    return b"".join(out)


def gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flags=0, extra=b"", name=b"", comment=b""):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    body = c.compress(data) + c.flush()
    flg = (4 if extra else 0) | (8 if name else 0) | (16 if comment else 0) | flags
    hdr = b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\0\x03"
    if extra:
        hdr += struct.pack("<H", len(extra)) + extra
    if name:
        hdr += name + b"\0"
    if comment:
        hdr += comment + b"\0"
    if flg & 2:
        hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
    return hdr + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF)


def zcat(blob, threads=4, piece_kb=None):
    env = dict(os.environ, SHB_DECOMP_THREADS=str(threads))
    if piece_kb:
        env["SHB_PGZIP_PIECE_KB"] = str(piece_kb)
    p = subprocess.run([ZCAT, "-"], input=blob, capture_output=True, env=env, timeout=300)
    return p.returncode, p.stdout, p.stderr.decode()


DATA = fasta(6_000_000, seed=1)
CASES = {
    "level1": lambda: (gz(DATA, 1), DATA),
    "level6": lambda: (gz(DATA, 6), DATA),
    "level9": lambda: (gz(DATA, 9), DATA),
    "fixed_huffman": lambda: (gz(DATA[:2_000_000], 6, zlib.Z_FIXED), DATA[:2_000_000]),
    "huffman_only": lambda: (gz(DATA[:2_000_000], 6, zlib.Z_HUFFMAN_ONLY), DATA[:2_000_000]),
    "rle": lambda: (gz(DATA[:2_000_000], 6, zlib.Z_RLE), DATA[:2_000_000]),
    "stored": lambda: (gz(DATA[:1_500_000], 0), DATA[:1_500_000]),
    "incompressible": lambda: (lambda r: (gz(r, 6), r))(np.random.default_rng(3).bytes(1_500_000)),
    "header_fields": lambda: (gz(DATA[:900_000], 6, flags=2, extra=b"AB\x02\x00xy", name=b"f.fa", comment=b"c"), DATA[:900_000]),
    "multi_member": lambda: (b"".join(gz(DATA[i:i + 300_001], 6, name=b"x") for i in range(0, 3_000_000, 300_001)),
                             DATA[:3_000_010]),
    "bgzf_like": lambda: (b"".join(gz(DATA[i:i + 65280], 6, extra=b"BC\x02\x00\x00\x00") for i in range(0, 2_000_000, 65280)) + gz(b""),
                          DATA[:(2_000_000 + 65279) // 65280 * 65280]),
    "empty_member": lambda: (gz(b""), b""),
    "tiny": lambda: (gz(b">a\nACGT\n"), b">a\nACGT\n"),
    "members_with_empties": lambda: (gz(b"") + gz(b">a\nAC\n") + gz(b"") + gz(b"GT\n"), b">a\nAC\nGT\n"),
    "mixed": lambda: (lambda m: (gz(m, 6), m))(DATA[:1_000_000] + np.random.default_rng(4).bytes(700_000) + DATA[1_000_000:2_500_000]
                                               + b"A" * 2_000_000 + DATA[:1000]),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("threads,piece_kb", [(1, None), (2, None), (4, 64), (8, 16)])
def test_gzip_streams_match_zlib(name, threads, piece_kb):
    blob, want = CASES[name]()
    rc, out, err = zcat(blob, threads, piece_kb)
    assert rc == 0, err
    assert out == want


def test_gzip_budget_path_on_a_1000_to_1_stream():
    # one deflate piece inflating 1000:1 is handed out over several batches (PIECE_BUDGET)
    want = b">polyA\n" + b"A" * 120_000_000 + b"\n"
    rc, out, err = zcat(gz(want, 6), 4)
    assert rc == 0, err
    assert out == want


def test_gzip_errors_like_the_reference_reader():
    blob = gz(DATA[:2_000_000], 6)
    rc, out, err = zcat(blob[:-50_000], 4, 64)
    assert rc == 1 and "unexpected EOF" in err
    assert DATA.startswith(out) and len(out) > 1_000_000   # what was complete is still delivered first
    bad = bytearray(blob)
    bad[-6] ^= 0x55
    rc, _, err = zcat(bytes(bad), 4, 64)
    assert rc == 1 and "invalid checksum" in err
    rc, _, err = zcat(blob + b"\0" * 512, 4, 64)
    assert rc == 1 and "invalid header" in err
    hdr = bytearray(gz(DATA[:100_000], 6, flags=2, name=b"f.fa"))
    hdr[12] ^= 0x01                                            # inside FNAME: the header CRC-16 no longer matches
    rc, _, err = zcat(bytes(hdr), 4, 64)
    assert rc == 1 and "invalid header" in err
    bad = bytearray(blob)
    bad[300_000] ^= 0xFF
    rc, _, err = zcat(bytes(bad), 4, 64)
    assert rc == 1 and ("invalid" in err or "EOF" in err)


@pytest.mark.parametrize("ext", ["", ".gz", ".zst", ".xz", ".bz2"])
def test_reference_fixture_twins_decompress_identically(ext):
    plain = open(os.path.join(FIX, "test.fasta"), "rb").read()
    path = os.path.join(FIX, "test.fasta" + ext)
    if not os.path.exists(path):
        pytest.skip("fixture twin not present")
    p = subprocess.run([ZCAT, path], capture_output=True, timeout=60)
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout == plain


@pytest.mark.parametrize("threads", [1, 4])
def test_bzip2_and_xz_streams(threads):
    data = DATA[:3_000_000]
    for blob in (bz2.compress(data, 1), bz2.compress(data[:1_000_000], 9) + bz2.compress(data[1_000_000:], 5),
                 lzma.compress(data, preset=1), lzma.compress(data[:1_500_000], preset=0) + lzma.compress(data[1_500_000:], preset=0)):
        rc, out, err = zcat(blob, threads)
        assert rc == 0, err
        assert out == data


def test_bzip2_edge_cases_parallel():
    poly = b">polyA\n" + b"A" * 30_000_000 + b"\n"          # run-length expansion far beyond the block size
    small_blocks = b"".join(bz2.compress(DATA[i:i + 150_000], 1) for i in range(0, 1_500_000, 150_000))   # 10 streams
    for blob, want in ((bz2.compress(b""), b""), (bz2.compress(b">a\nACGT\n"), b">a\nACGT\n"), (bz2.compress(poly, 9), poly),
                       (small_blocks, DATA[:1_500_000]), (bz2.compress(b"") + bz2.compress(b"xyz"), b"xyz")):
        rc, out, err = zcat(blob, 4)
        assert rc == 0, err
        assert out == want
    blob = bz2.compress(DATA[:3_000_000], 5)
    rc, out, err = zcat(blob[:-40_000], 4)
    assert rc == 1 and "EOF" in err and DATA.startswith(out) and len(out) > 1_000_000
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0x10
    rc, out, err = zcat(bytes(bad), 4)
    assert rc == 1 and "bzip2" in err and DATA.startswith(out)
    bad = bytearray(blob)
    bad[-2] ^= 0x01                                           # stream (combined) CRC
    rc, out, err = zcat(bytes(bad), 4)
    assert rc == 1 and "checksum" in err


def test_xz_multiblock_threaded():
    import shutil
    if not shutil.which("xz"):
        pytest.skip("xz CLI not installed")
    data = DATA[:5_000_000]
    blob = subprocess.run(["xz", "-c", "-T2", "-1", "--block-size=512KiB"], input=data, capture_output=True, check=True).stdout
    for threads in (1, 4):
        rc, out, err = zcat(blob, threads)
        assert rc == 0, err
        assert out == data
    rc, out, err = zcat(blob[:-30_000], 4)
    assert rc == 1


def test_truncated_inputs_fail_on_the_single_thread_paths_too():
    import shutil
    data = DATA[:1_000_000]
    blobs = [gz(data, 6), bz2.compress(data, 3), lzma.compress(data, preset=0)]
    if shutil.which("zstd"):
        blobs.append(subprocess.run(["zstd", "-c", "-3"], input=data, capture_output=True, check=True).stdout)
    for blob in blobs:
        rc, out, err = zcat(blob, 1)
        assert rc == 0 and out == data, err
        rc, out, err = zcat(blob[:len(blob) * 2 // 3], 1)
        assert rc == 1 and "EOF" in err, (blob[:4], err)
        assert data.startswith(out)


@pytest.mark.parametrize("seed", [1, 2])
def test_gzip_random_streams_soak(seed):
    """Random deflate streams: mixed content, random levels / strategies, random sync / full flushes in the middle
    (empty stored blocks, window resets), several members, random worker counts and piece sizes."""
    rng = np.random.default_rng(seed)
    for it in range(12):
        parts, blob = [], b""
        for _member in range(int(rng.integers(1, 4))):
            c = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, -15, int(rng.integers(5, 10)),
                                 int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])))
            body, member = b"", b""
            for _piece in range(int(rng.integers(1, 6))):
                kind = rng.integers(4)
                n = int(rng.integers(0, 400_000))
                if kind == 0:
                    chunk = DATA[int(rng.integers(0, 1_000_000)):][:n]
                elif kind == 1:
                    chunk = rng.bytes(n // 4)
                elif kind == 2:
                    chunk = bytes([int(rng.integers(32, 127))]) * n
                else:
                    chunk = (b"ACGT" * 50 + b"\n") * (n // 201)
                member += chunk
                body += c.compress(chunk)
                if rng.random() < 0.5:
                    body += c.flush(int(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))
            body += c.flush()
            blob += (b"\x1f\x8b\x08\0\0\0\0\0\0\x03" + body +
                     struct.pack("<II", zlib.crc32(member) & 0xFFFFFFFF, len(member) & 0xFFFFFFFF))
            parts.append(member)
        want = b"".join(parts)
        rc, out, err = zcat(blob, int(rng.choice([2, 3, 8])), int(rng.choice([8, 32, 128, 1024])))
        assert rc == 0, (seed, it, err)
        assert out == want, (seed, it, len(out), len(want))


def test_bzip2_random_streams_soak():
    """Random bzip2 inputs: several concatenated streams of random levels and content (runs that expand far beyond the
    block size, incompressible bytes, text), random worker counts."""
    rng = np.random.default_rng(5)
    for it in range(8):
        parts, blob = [], b""
        for _stream in range(int(rng.integers(1, 5))):
            kind, n = rng.integers(4), int(rng.integers(0, 1_500_000))
            if kind == 0:
                chunk = DATA[int(rng.integers(0, 1_000_000)):][:n]
            elif kind == 1:
                chunk = rng.bytes(n // 3)
            elif kind == 2:
                chunk = bytes([int(rng.integers(32, 127))]) * (n * 4)
            else:
                chunk = (b"ACGT" * 50 + b"\n") * (n // 201)
            parts.append(chunk)
            blob += bz2.compress(chunk, int(rng.integers(1, 10)))
        rc, out, err = zcat(blob, int(rng.choice([2, 4, 8])))
        assert rc == 0, (it, err)
        assert out == b"".join(parts), (it, len(out))


# ---- zstd: frames decoded in parallel (pzstd.hpp) ----

def _zstd_lib():
    import ctypes
    try:
        L = ctypes.CDLL("libzstd.so.1")
    except OSError:
        pytest.skip("libzstd.so.1 not available")
    L.ZSTD_compressBound.argtypes = [ctypes.c_size_t]; L.ZSTD_compressBound.restype = ctypes.c_size_t
    L.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    L.ZSTD_compress.restype = ctypes.c_size_t
    L.ZSTD_isError.argtypes = [ctypes.c_size_t]; L.ZSTD_isError.restype = ctypes.c_uint
    return L


def zst_frame(L, data, level=3):
    import ctypes
    cap = L.ZSTD_compressBound(len(data))
    buf = ctypes.create_string_buffer(cap)
    n = L.ZSTD_compress(buf, cap, data, len(data), level)
    assert not L.ZSTD_isError(n)
    return buf.raw[:n]


def zcat_zst(blob, threads=4, batch_kb=None):
    env = dict(os.environ, SHB_DECOMP_THREADS=str(threads))
    if batch_kb:
        env["SHB_PZSTD_BATCH_KB"] = str(batch_kb)
    p = subprocess.run([ZCAT, "-"], input=blob, capture_output=True, env=env, timeout=300)
    return p.returncode, p.stdout, p.stderr.decode()


def test_zstd_many_frames_in_parallel_equal_the_single_stream():
    """A .zst file of many frames (pzstd, concatenated `zstd` runs): frames are delimited by their headers and decoded on
    several threads; batch seams anywhere, frames larger than a batch and the single-frame file fall back to the streaming
    decoder; every variant must give the same bytes as one thread."""
    L = _zstd_lib()
    rng = np.random.default_rng(3)
    data = DATA[:4_000_000]
    cuts = sorted(set([0, len(data)] + [int(x) for x in rng.integers(0, len(data), size=40)]))
    many = b"".join(zst_frame(L, data[a:b], int(rng.integers(1, 10))) for a, b in zip(cuts, cuts[1:]))
    one = zst_frame(L, data, 3)
    skippable = struct.pack("<II", 0x184D2A50, 5) + b"hello"                       # a skippable frame between data frames
    mixed = zst_frame(L, data[:1000]) + skippable + zst_frame(L, b"") + zst_frame(L, data[1000:])
    for blob in (many, one, mixed):
        for threads, batch_kb in ((1, None), (4, None), (4, 64), (3, 300), (8, 1)):
            rc, out, err = zcat_zst(blob, threads, batch_kb)
            assert rc == 0, (threads, batch_kb, err)
            assert out == data, (threads, batch_kb, len(out))


def test_zstd_damaged_and_truncated_multi_frame_input_fails_after_the_good_frames():
    L = _zstd_lib()
    data = DATA[:2_000_000]
    frames = [zst_frame(L, data[i:i + 100_000]) for i in range(0, len(data), 100_000)]
    blob = b"".join(frames)
    for threads, batch_kb in ((4, None), (4, 128), (1, None)):
        rc, out, err = zcat_zst(blob[:len(blob) - len(frames[-1]) // 2], threads, batch_kb)
        assert rc == 1 and "EOF" in err, err
        assert data.startswith(out) and len(out) >= 1_800_000
        bad = bytearray(blob)
        k = sum(len(f) for f in frames[:10]) + len(frames[10]) // 2                 # inside frame 10
        bad[k] ^= 0xff
        rc, out, err = zcat_zst(bytes(bad), threads, batch_kb)
        assert rc == 1 and "zstd" in err, err
        assert data.startswith(out)
        rc, out, err = zcat_zst(blob + b"garbage at the end", threads, batch_kb)
        assert rc == 1 and data.startswith(out)
