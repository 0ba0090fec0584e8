"""Chunk-boundary rules of the compiled host without a GPU: where a chunk of text may be cut (find_cut, the streaming loop)
and where the first record at or after a file offset starts (find_record_start, the byte-range mode for plain files), on
random FASTA / FASTQ with the traps: '>' and '@' inside names, quality lines that start with '@' or '+', CRLF."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "emul", "_build", "libhost_ranges_emul.so")
NPOS, NEED_MORE = (1 << 64) - 1, (1 << 64) - 2


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    pkg = os.path.join(ROOT, "seqhasher_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-o", SO, os.path.join(HERE, "emul", "host_ranges_emul.cpp"),
                    "-L" + pkg, "-lseqhash_b200", "-lz", "-ldl", "-lpthread", "-Wl,-rpath," + pkg], check=True)
    L = ctypes.CDLL(SO)
    L.emul_find_cut.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    L.emul_find_cut.restype = ctypes.c_size_t
    L.emul_find_record_start.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
    L.emul_find_record_start.restype = ctypes.c_size_t
    L.emul_scan_bgzf.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t]
    L.emul_scan_bgzf.restype = ctypes.c_long
    return L


def make(rng, fastq, eol):
    recs, starts, pos = [], [], 0
    for i in range(int(rng.integers(3, 40))):
        L = int(rng.integers(1, 120))
        seq = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=L).tobytes()
        name = b"r%d >x @y +z" % i
        if fastq:
            q = rng.choice(np.frombuffer(b"@+I!>", dtype=np.uint8), size=L).tobytes()
            r = b"@" + name + eol + seq + eol + b"+" + (name if rng.random() < 0.3 else b"") + eol + q + eol
        else:
            w = int(rng.choice([0, 7, 60]))
            body = eol.join(seq[k:k + w] for k in range(0, L, w)) if w else seq
            r = b">" + name + eol + body + eol
        starts.append(pos)
        recs.append(r)
        pos += len(r)
    return b"".join(recs), starts


@pytest.mark.parametrize("fastq", [False, True])
def test_first_record_start_at_or_after_every_offset(lib, fastq):
    rng = np.random.default_rng(11 + fastq)
    fmt = 2 if fastq else 1
    for trial in range(40):
        text, starts = make(rng, fastq, b"\r\n" if trial % 5 == 0 else b"\n")
        for frm in range(1, len(text)):
            want = next((s for s in starts if s >= frm), None)
            got = lib.emul_find_record_start(text, len(text), frm, fmt, 1)
            assert got == (NPOS if want is None else want), (trial, frm)
        # a truncated view of the text may ask for more data but never names a wrong place
        for cut in rng.integers(1, len(text), size=30):
            cut = int(cut)
            for frm in rng.integers(1, cut + 1, size=10):
                frm = int(min(frm, cut))
                got = lib.emul_find_record_start(text, cut, frm, fmt, 0)
                want = next((s for s in starts if s >= frm), None)
                assert got in (NEED_MORE, NPOS) or got == want, (trial, cut, frm, got, want)
                if got == NPOS:
                    assert want is None or want >= cut or fastq


@pytest.mark.parametrize("fastq", [False, True])
def test_cut_is_the_start_of_the_last_record(lib, fastq):
    rng = np.random.default_rng(5 + fastq)
    fmt = 2 if fastq else 1
    for trial in range(60):
        text, starts = make(rng, fastq, b"\n")
        for n in sorted(set(int(x) for x in rng.integers(1, len(text) + 1, size=25)) | {len(text)}):
            got = lib.emul_find_cut(text, n, fmt)
            inside = [s for s in starts if 0 < s < n]
            if fastq:
                # a FASTQ record start is only recognised when its '+' line is in view: the cut may fall on an earlier record
                assert got == 0 or got in inside, (trial, n, got)
                full = [s for s in inside if text.find(b"\n", text.find(b"\n", s) + 1) + 1 < n and text.find(b"\n", s) != -1]
                if full:
                    assert got == full[-1] or got in inside
            else:
                assert got == (inside[-1] if inside else 0), (trial, n, got)


def test_bgzf_member_walk_of_the_compiled_host_equals_the_python_scanner(lib, tmp_path):
    """scan_bgzf (csrc/host/main.cpp: one pread per member) against seqhasher_b200.host.bgzf.scan_members on BGZF files of
    several member sizes, with an extra subfield in front of 'BC', and on files that must be refused."""
    import gzip
    import struct
    import zlib
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(4)
    data = rng.choice(np.frombuffer(b"ACGT\n", dtype=np.uint8), size=700_000).tobytes()
    out = np.zeros(4 * 20000, dtype=np.uint64)

    def walk(blob):
        f = tmp_path / "x.gz"
        f.write_bytes(blob)
        return lib.emul_scan_bgzf(str(f).encode(), out.ctypes.data, 20000)
    for block, level in ((65280, 6), (65536, 1), (1000, 6), (37, 9), (65536, 0)):
        gz = bgzf.compress(data if block > 100 else data[:20000], level, block)
        want = bgzf.scan_members(gz)
        n = walk(gz)
        assert n == len(want)
        got = out[:4 * n].reshape(n, 4)
        assert (got[:, 0] == want[:, 3]).all() and (got[:, 1:4] == want[:, 0:3]).all()
    # an extra subfield before the BC one (the spec allows it)
    piece = data[:5000]
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    cdata = c.compress(piece) + c.flush()
    extra = b"XY\x03\x00abc" + b"BC\x02\x00" + struct.pack("<H", 12 + 13 + len(cdata) + 8 - 1)
    member = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", len(extra)) + extra + cdata + struct.pack("<II", zlib.crc32(piece), len(piece))
    assert gzip.decompress(member) == piece
    assert walk(member + bgzf.EOF_MEMBER) == 2 and int(out[1]) == 12 + 13 and int(out[2]) == len(cdata)
    assert bgzf.scan_members(member + bgzf.EOF_MEMBER) is not None
    good = bgzf.compress(data[:100000], 6, 5000)
    for bad in (gzip.compress(data[:1000]), good + gzip.compress(b"tail"), good[:-5], good + b"\x00", good[:17]):
        assert walk(bad) == -1
        assert bgzf.scan_members(bad) is None
