"""BGZF input inflated on the device (csrc/inflate.cuh, shb_textproc_run_bgzf) against the same text given to
shb_textproc_run uncompressed: the output must be byte-identical whatever the member size, the compression level, the
chunking and the look-ahead.  The plain path itself is checked against the oracle in test_textproc_gpu.py."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from seqhasher_b200 import binding
    binding.init(1)
    return binding


def _fasta(rng, nrec, lo, hi, width=60, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    parts = []
    for i in range(nrec):
        L = int(rng.integers(lo, hi + 1))
        seq = rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=L).tobytes()
        parts.append(b">seq_%d len=%d" % (i, L) + nl)
        for k in range(0, L, width):
            parts.append(seq[k:k + width] + nl)
        if i % 97 == 0:
            parts.append(nl)                        # a blank line between records
    return b"".join(parts)


def _fastq(rng, nrec, lo, hi):
    parts = []
    for i in range(nrec):
        L = int(rng.integers(lo, hi + 1))
        seq = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=L).tobytes()
        qual = rng.choice(np.frombuffer(b"@+IIIIFF:,#5>", dtype=np.uint8), size=L).tobytes()   # '@' and '+' may start a quality line
        parts.append(b"@read_%d\n%s\n+\n%s\n" % (i, seq, qual))
    return b"".join(parts)


def _plain(B, tp, text, fmt, ids, flags, headers_only, label):
    out = []
    pos = 0
    n = len(text)
    while pos < n:
        m = n - pos
        assert m <= tp.capacity
        tp.input[:m] = np.frombuffer(text, dtype=np.uint8)[pos:]
        o, consumed, nrec, nempty = tp.run(m, fmt, True, ids, flags, headers_only, label)
        out.append(o.tobytes())
        if consumed == 0 or nrec == 0:
            break
        pos += consumed
    return b"".join(out)


def _bgzf(B, tp, gz, fmt, first_start, ids, flags, headers_only, label, own, extra0=2):
    from seqhasher_b200.host import bgzf
    members = bgzf.scan_members(gz)
    assert members is not None
    return b"".join(o for o, _, _ in bgzf.run_chunks(tp, gz, members, fmt, first_start, ids, flags, headers_only, label, own, extra0))


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("block", [65280, 65536, 4000, 333])
def test_fasta_levels_and_member_sizes(B, level, block):
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(level * 100 + block)
    text = _fasta(rng, 3000 if block > 4000 else 600, 30, 900)
    gz = bgzf.compress(text, level, block)
    ids = [B.algo_id(a) for a in ("sha1", "xxh3", "nthash")]
    tp = B.TextProc(8 << 20)
    want = _plain(B, tp, text, 1, ids, B.FLAG_UPPER, False, b"in.fa")
    for own in (1, 7, 10 ** 6):
        if block == 333 and own == 1:
            continue   # a chunk per 333 bytes: covered by the FASTQ test below on a smaller file
        got = _bgzf(B, tp, gz, 1, 0, ids, B.FLAG_UPPER, False, b"in.fa", own)
        assert got == want, (level, block, own)
    tp.close()


def test_fastq_quality_lines_that_look_like_markers(B):
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(5)
    text = _fastq(rng, 2500, 20, 400)
    ids = [B.algo_id(a) for a in ("md5", "xxhash")]
    tp = B.TextProc(4 << 20)
    want = _plain(B, tp, text, 2, ids, 0, True, None)
    for block, own in ((65280, 1), (65280, 3), (1500, 1), (1500, 5), (257, 2)):
        gz = bgzf.compress(text, 6, block)
        got = _bgzf(B, tp, gz, 2, 0, ids, 0, True, None, own)
        assert got == want, (block, own)
    tp.close()


def test_long_records_need_more_lookahead(B):
    """Records of 100-400 kB span several 64 KiB members: the first call of a chunk answers SHB_NEED_MORE until the
    look-ahead covers the record that crosses the chunk's end; chunks that lie inside one record own nothing."""
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(11)
    text = _fasta(rng, 12, 100_000, 400_000, width=80)
    gz = bgzf.compress(text, 6)
    ids = [B.algo_id(a) for a in ("blake3", "k12")]
    tp = B.TextProc(8 << 20)
    want = _plain(B, tp, text, 1, ids, 0, False, b"x")
    for own in (1, 2, 5):
        got = _bgzf(B, tp, gz, 1, 0, ids, 0, False, b"x", own, extra0=1)
        assert got == want, own
    tp.close()


def test_leading_and_trailing_blank_lines_crlf_and_missing_last_newline(B):
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(3)
    body = _fasta(rng, 400, 10, 300, crlf=True)
    ids = [B.algo_id("sha1")]
    tp = B.TextProc(2 << 20)
    for lead, tail in ((b"", b""), (b"\n\n  \n", b"\n\n\n"), (b"\r\n", b"   ")):
        for cut_last_nl in (False, True):
            text = lead + (body[:-2] if cut_last_nl else body) + tail
            want = _plain(B, tp, text[len(lead):].rstrip(b" \t\r\n\v\f"), 1, ids, B.FLAG_UPPER, False, b"f")
            gz = bgzf.compress(text, 6, 5000)
            got = _bgzf(B, tp, gz, 1, len(lead), ids, B.FLAG_UPPER, False, b"f", 3)
            assert got == want, (lead, tail, cut_last_nl)
    tp.close()


def test_more_records_than_the_index_holds_resume(B):
    from seqhasher_b200.host import bgzf
    n = 90_000                                           # index of a 1 MiB textproc: 1 MiB / 16 + 1024 = 66 560 records
    text = b"".join(b">%d\nAC\n" % i for i in range(n))
    assert len(text) < (1 << 20)
    gz = bgzf.compress(text, 6)
    ids = [B.algo_id("murmur3")]
    tp = B.TextProc(1 << 20)
    want = _plain(B, tp, text, 1, ids, B.FLAG_UPPER, True, None)
    got = _bgzf(B, tp, gz, 1, 0, ids, B.FLAG_UPPER, True, None, 10 ** 6)
    assert got == want and got.count(b"\n") == n
    tp.close()


def test_damaged_members_are_refused(B):
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(9)
    text = _fasta(rng, 500, 30, 600)
    gz = bytearray(bgzf.compress(text, 6, 20000))
    members = bgzf.scan_members(bytes(gz))
    ids = [B.algo_id("sha1")]
    tp = B.TextProc(2 << 20)
    # a wrong CRC-32 in the trailer of member 2
    bad = bytearray(gz)
    t = int(members[2][0] + members[2][1])
    bad[t] ^= 0x40
    with pytest.raises(B.SeqhashError, match="invalid checksum"):
        _bgzf(B, tp, bytes(bad), 1, 0, ids, 0, False, None, 10 ** 6)
    # flipped bits inside the DEFLATE data: either the stream no longer decodes or the checksum catches it
    for trial in range(20):
        bad = bytearray(gz)
        m = members[int(rng.integers(0, len(members) - 1))]
        bad[int(m[0]) + int(rng.integers(0, int(m[1])))] ^= 1 << int(rng.integers(0, 8))
        with pytest.raises(B.SeqhashError, match="invalid checksum|corrupt input"):
            _bgzf(B, tp, bytes(bad), 1, 0, ids, 0, False, None, 10 ** 6)
    # and the undamaged file still goes through afterwards
    assert _bgzf(B, tp, bytes(gz), 1, 0, ids, 0, False, None, 4) == _plain(B, tp, text, 1, ids, 0, False, None)
    tp.close()


# ---- the compiled host on BGZF files ----

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "seqhasher_b200", "bin", "seqhasher-b200")


def _cli(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([BIN] + args, capture_output=True, timeout=300, env=e)


@pytest.mark.parametrize("kind", ["fasta", "fastq"])
def test_cli_bgzf_equals_plain_and_host_inflate(tmp_path, kind):
    """seqhasher-b200 on x.gz written as BGZF: device inflate (default), the host-side inflate (SHB_NO_BGZF=1) and the plain
    file give the same bytes; small chunks (SHB_CHUNK_KB) put dozens of chunk seams and look-ahead extensions in play."""
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(21)
    text = (b"\n\n" + _fasta(rng, 4000, 30, 3000) + b"\n\n") if kind == "fasta" else _fastq(rng, 6000, 50, 600)
    plain, gz = tmp_path / f"in.{kind}", tmp_path / f"in.{kind}.gz"
    plain.write_bytes(text)
    gz.write_bytes(bgzf.compress(text, 6))
    args = ["--name", "in", "--hash", "sha1,xxh3,k12"]
    want = _cli(args + [str(plain), "-"])
    assert want.returncode == 0 and want.stdout.count(b"\n") > 4000
    for env in ({}, {"SHB_CHUNK_KB": "300"}, {"SHB_CHUNK_KB": "70", "SHB_SLOTS_PER_GPU": "3"}, {"SHB_NO_BGZF": "1"}):
        got = _cli(args + [str(gz), "-"], env)
        assert got.returncode == 0, got.stderr
        assert got.stdout == want.stdout, env
    ho = _cli(["--headersonly", "--casesensitive", "--hash", "md5"] + [str(gz), "-"], {"SHB_CHUNK_KB": "500"})
    hw = _cli(["--headersonly", "--casesensitive", "--hash", "md5", "--name", str(gz)] + [str(plain), "-"])
    assert ho.returncode == 0 and ho.stdout == hw.stdout


def test_cli_bgzf_damaged_member_reports_and_keeps_earlier_records(tmp_path):
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(22)
    text = _fasta(rng, 3000, 100, 400)
    gz = bytearray(bgzf.compress(text, 6))
    members = bgzf.scan_members(bytes(gz))
    assert len(members) > 8
    k = len(members) - 3
    gz[int(members[k][0] + members[k][1])] ^= 1                     # CRC-32 of a late member
    f = tmp_path / "bad.fa.gz"
    f.write_bytes(bytes(gz))
    p = _cli(["--headersonly", str(f), "-"], {"SHB_CHUNK_KB": "200"})
    assert p.returncode == 1 and b"Error reading record: gzip: invalid checksum" in p.stderr
    assert p.stdout.count(b"\n") > 100                               # the chunks before the damaged one were written


def test_cli_gzip_that_is_not_bgzf_still_takes_the_host_inflate(tmp_path):
    import gzip
    text = b">a\nACGT\n>b\nacgtn\n"
    f = tmp_path / "x.fa.gz"
    f.write_bytes(gzip.compress(text))
    p = _cli(["--headersonly", "--name", "x", str(f), "-"])
    assert p.returncode == 0 and p.stdout.count(b"\n") == 2


def test_cli_bgzf_two_gpus_match_one(tmp_path):
    """SHB_GPUS=2 on a BGZF file: chunks (runs of members) alternate between two devices, each inflating its own members;
    output order = chunk order, byte-identical to one GPU and to the plain file."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    from seqhasher_b200.host import bgzf
    rng = np.random.default_rng(31)
    text = _fasta(rng, 20000, 30, 3000)
    plain, gz = tmp_path / "in.fa", tmp_path / "in.fa.gz"
    plain.write_bytes(text)
    gz.write_bytes(bgzf.compress(text, 6))
    args = ["--name", "in", "--hash", "sha1,blake3,nthash"]
    want = _cli(args + [str(plain), "-"], {"SHB_GPUS": "1"})
    assert want.returncode == 0
    for env in ({"SHB_GPUS": "1", "SHB_CHUNK_KB": "2048"}, {"SHB_GPUS": "2", "SHB_CHUNK_KB": "2048"}, {"SHB_GPUS": "2", "SHB_CHUNK_KB": "700", "SHB_SLOTS_PER_GPU": "3"}):
        got = _cli(args + [str(gz), "-"], env)
        assert got.returncode == 0, got.stderr
        assert got.stdout == want.stdout, env


def test_textproc_run_without_staged_text_is_refused(B):
    """The pinned text buffer is allocated by shb_textproc_input(); running a chunk that was never staged is a caller error."""
    import ctypes
    L = B.lib()
    h = L.shb_textproc_create(0, 1 << 20)
    assert h
    ids = (ctypes.c_int * 1)(0)
    rc = L.shb_textproc_run(h, 100, 1, 1, ids, 1, 0, 1, None)
    assert rc == -2 and b"shb_textproc_input" in L.shb_last_error()
    L.shb_textproc_destroy(h)
