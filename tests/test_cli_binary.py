"""The compiled host (seqhasher_b200/bin/seqhasher-b200): CLI contract without a GPU (version, usage, flag and
file errors: seqhasher_test.go:548-660, :112) and, on the GPU, the reference's binary test script
(test/test2_binary.sh:10-186) case by case."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "seqhasher_b200", "bin", "seqhasher-b200")
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
S1 = {"ACTG": "65c89f59d38cdbf90dfaf0b0a6884829df8396b0", "AAAA": "e2512172abf8cc9f67fdd49eb6cacf2df71bbad3",
      "TGCA": "e3da52abc8fbdb38b113a187ed0ac763fa86d1d4", "aaaa": "70c881d4a26984ddce795f6f71817c9cf4480e79"}


def run(args, stdin=None, cwd=FIX):
    p = subprocess.run([BIN] + args, input=stdin, capture_output=True, cwd=cwd, timeout=120)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


def test_binary_exists():
    assert os.access(BIN, os.X_OK), "build with __graft_entry__.build() (make -C seqhasher_b200/csrc)"


def test_version_usage_and_errors_without_gpu():
    assert run(["-version"]) == (0, "SeqHasher 1.2.0\n", "")
    assert run(["--version"])[1] == "SeqHasher 1.2.0\n"
    rc, out, _ = run([])
    assert rc == 0 and "Usage:" in out and "Supported hash types: sha1, sha3, md5" in out
    rc, _, err = run(["--hash", "invalid", "test.fasta"])
    assert rc == 1 and err.strip() == ("Invalid hash type: invalid. Supported types are: sha1, sha3, md5, xxhash, xxh3, "
                                       "cityhash, murmur3, nthash, blake3, k12, rapidhash, rapidhash32")
    rc, _, err = run(["/nonexistent/in.fa"])
    assert rc == 1 and err.startswith("Error opening input: open /nonexistent/in.fa: no such file or directory")
    rc, _, err = run(["--bogus"])
    assert rc == 2 and "flag provided but not defined: -bogus" in err
    rc, _, err = run(["-h"])
    assert rc == 0 and "Usage:" in err


@pytest.mark.gpu
def test_binary_script_cases():
    a, t, l = S1["AAAA"], S1["ACTG"], S1["aaaa"]
    assert run(["test2.fasta", "-"])[1] == f">test2.fasta;{a};seq1\nAAAA\n>test2.fasta;{t};seq2\nACTG\n>test2.fasta;{a};seq3\nAAAA\n"
    assert run(["--name", "custom_name", "test2.fasta", "-"])[1].startswith(f">custom_name;{a};seq1\nAAAA\n")
    assert run(["--headersonly", "test2.fasta", "-"])[1] == f"test2.fasta;{a};seq1\ntest2.fasta;{t};seq2\ntest2.fasta;{a};seq3\n"
    assert run(["--headersonly", "--nofilename", "test2.fasta", "-"])[1] == f"{a};seq1\n{t};seq2\n{a};seq3\n"
    assert run(["--headersonly", "--nofilename", "--hash", "xxhash", "--casesensitive", "test2.fasta", "-"])[1] == \
        "cf40b5b72bc43e77;seq1\n704b34bf20faedf2;seq2\n42a70d1abf84bf32;seq3\n"
    assert run(["--headersonly", "--nofilename", "--hash", "sha1,xxhash", "--casesensitive", "test2.fasta", "-"])[1] == \
        f"{a};cf40b5b72bc43e77;seq1\n{t};704b34bf20faedf2;seq2\n{l};42a70d1abf84bf32;seq3\n"
    for ext in ("bz2", "gz", "xz", "zst"):
        assert run(["--headersonly", "--nofilename", f"test2.fasta.{ext}", "-"])[1] == f"{a};seq1\n{t};seq2\n{a};seq3\n", ext
    names = ["seq1_uppercase", "seq2_lowercase", "seq3_mixed_case", "seq4_with_spaces", "seq5_with_tabs", "seq6_with_newline"]
    assert run(["--hash", "sha1", "--nofilename", "seqs_with_spaces.fasta", "-"])[1] == "".join(f">{t};{n}\nACTG\n" for n in names)
    fq = open(os.path.join(FIX, "test3.fastq"), "rb").read()
    assert run(["-", "-"], stdin=fq)[1] == (f"@{t};seq1\nACTG\n+\nDFGH\n@{a};seq2\nAAAA\n+\nBBBB\n@{S1['TGCA']};seq3\nTGCA\n+\nEEEE\n")


@pytest.mark.gpu
def test_binary_long_sequences_output_file_and_empty_record(tmp_path):
    data = b">seq1\n" + b"ACTG" * 25000 + b"\n>seq2\n" + b"A" * 100000 + b"\n"
    p = tmp_path / "test_long.fasta"
    p.write_bytes(data)
    rc, out, _ = run(["--headersonly", "--nofilename", str(p), "-"])
    assert out == "3200abf5034212fb5d90fa242fc1cea5e0f3dd00;seq1\ne27b41eb7f7370b36f1b2fa63f05652f52423277;seq2\n"
    o = tmp_path / "out.fa"
    rc, out, _ = run(["--hash", "md5", "test.fasta", str(o)])
    assert rc == 0 and out == "" and o.read_bytes().startswith(b">test.fasta;86bfb9f78dd8b6cd35962bb7324fdbf8;seq1\nACTG\n")
    rc, out, err = run(["--headersonly", "--hash", "sha1,md5", "problematic_sequences.fasta", "-"])
    assert out.splitlines()[-1] == "problematic_sequences.fasta;;;seq_empty" and err.count("Empty DNA sequence") == 2
    rc, _, err = run(["-"], stdin=b"not a sequence file\n")
    assert rc == 1 and err.strip() == "Failed to create reader: fastx: invalid FASTA/Q format"


def _records(n, seed, lo=30, hi=3000, wrap=0):
    import numpy as np
    r = np.random.default_rng(seed)
    lens = r.integers(lo, hi + 1, size=n)
    seq = r.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), size=int(lens.sum())).tobytes()
    recs, pos = [], 0
    for i, L in enumerate(lens):
        recs.append((b"r%d" % i, seq[pos:pos + L]))
        pos += L
    text = b"".join(b">" + nm + b" d\n" + (b"\n".join(s[j:j + wrap] for j in range(0, len(s), wrap)) if wrap else s) + b"\n"
                    for nm, s in recs)
    return recs, text


def _expected_headers(recs, label=None):
    import hashlib
    pre = (label + b";") if label else b""
    return b"".join(pre + hashlib.sha1(s.upper()).hexdigest().encode() + b";" + nm + b" d\n" for nm, s in recs)


@pytest.mark.gpu
def test_binary_many_chunks_plain_stdin_and_gzip(tmp_path):
    """> 64 MB of text = several chunks through the record loop: output order and carried partial records; file
    (parallel pread) vs stdin vs gzip input (parallel feeder on a read-ahead thread) must all give the same bytes."""
    import gzip
    recs, text = _records(150_000, seed=21, wrap=70)          # ~230 MB, 4 chunks
    want = _expected_headers(recs)
    p = tmp_path / "many.fa"
    p.write_bytes(text)
    rc, out, err = run(["--headersonly", "--nofilename", str(p), "-"])
    assert rc == 0, err
    assert out.encode() == want
    rc, out, err = run(["--headersonly", "--nofilename", "-"], stdin=text)
    assert rc == 0 and out.encode() == want
    g = tmp_path / "many.fa.gz"
    g.write_bytes(gzip.compress(text, 1))
    rc, out, err = run(["--headersonly", "--nofilename", str(g), "-"])
    assert rc == 0 and out.encode() == want
    # full-record output: sequences unwrapped and upper-cased
    rc, out, err = run(["--nofilename", str(p), "-"])
    lines = out.encode().split(b"\n")
    assert rc == 0 and len(lines) == 2 * len(recs) + 1
    assert lines[1] == recs[0][1].upper() and lines[-2] == recs[-1][1].upper()
    assert lines[2 * 77_777] == b">" + _expected_headers([recs[77_777]])[:-1]


@pytest.mark.gpu
def test_binary_grows_its_buffer_for_a_giant_record(tmp_path):
    """A record larger than the 64 MB chunk buffer (chromosome-sized FASTA entries) in the middle of a multi-chunk file."""
    import hashlib
    import numpy as np
    recs_a, text_a = _records(60_000, seed=5)                  # ~90 MB: the giant record starts in chunk 2
    big = np.random.default_rng(9).choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=150_000_000).tobytes()
    recs_b, text_b = _records(2_000, seed=6)
    recs_b = [(b"t" + nm, s) for nm, s in recs_b]
    text_b = text_b.replace(b">r", b">tr")
    text = text_a + b">giant d\n" + big + b"\n" + text_b
    p = tmp_path / "giant.fa"
    p.write_bytes(text)
    rc, out, err = run(["--headersonly", "--nofilename", "--casesensitive", str(p), "-"])
    assert rc == 0, err
    import hashlib as H
    want = b"".join(H.sha1(s).hexdigest().encode() + b";" + nm + b" d\n" for nm, s in recs_a + [(b"giant", big)] + recs_b)
    assert out.encode() == want


@pytest.mark.gpu
def test_binary_truncated_gzip_writes_complete_records_then_fails(tmp_path):
    import gzip
    recs, text = _records(40_000, seed=8)
    blob = gzip.compress(text, 1)
    g = tmp_path / "cut.fa.gz"
    g.write_bytes(blob[:len(blob) * 3 // 4])
    rc, out, err = run(["--headersonly", "--nofilename", str(g), "-"])
    assert rc == 1 and "unexpected EOF" in err
    got = out.encode()
    want = _expected_headers(recs)
    assert len(got) > len(want) // 2 and want.startswith(got)


def run_env(args, env, stdin=None, cwd=FIX):
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([BIN] + args, input=stdin, capture_output=True, cwd=cwd, timeout=300, env=e)
    return p.returncode, p.stdout, p.stderr.decode()


@pytest.mark.gpu
def test_binary_more_records_than_the_index_holds(tmp_path):
    """5 M five-byte records in one 64 MB chunk: more than the record index of a chunk (capacity / 16 + 1024) holds.
    The chunk is consumed in several passes (this used to fail with 'use smaller chunks')."""
    import hashlib
    n = 5_000_000
    bases = b"ACGT"
    text = b"".join(b">%d\n%c\n" % (i % 10, bases[i & 3]) for i in range(n))
    p = tmp_path / "tiny.fa"
    p.write_bytes(text)
    rc, out, err = run_env(["--headersonly", "--nofilename", str(p), "-"], {})
    assert rc == 0, err
    h = [hashlib.sha1(bytes([b])).hexdigest().encode() for b in bases]
    lines = out.split(b"\n")
    assert len(lines) == n + 1 and lines[-1] == b""
    assert lines[0] == h[0] + b";0" and lines[1234567] == h[1234567 & 3] + b";7" and lines[n - 1] == h[(n - 1) & 3] + b";9"
    import collections
    c = collections.Counter(lines[:-1])
    assert len(c) == 20 and sum(c.values()) == n   # ids 0-9 x bases 0-3 pair up as (i % 10, i & 3): 20 combinations


@pytest.mark.gpu
def test_binary_truncated_fastq_fails_after_the_complete_records(tmp_path):
    """The input ends inside a FASTQ record: the records before it are written, then the reader's error, rc 1
    (the tail used to vanish with rc 0)."""
    t = S1["ACTG"]
    for tail in (b"@r2\nAC", b"@r2\nACGT\n", b"@r2\nACGT\n+\n", b"@r2\nACGT\n+"):
        rc, out, err = run_env(["--nofilename", "-", "-"], {}, stdin=b"@r1\nACTG\n+\nIIII\n" + tail)
        assert rc == 1, tail
        assert out == f"@{t};r1\nACTG\n+\nIIII\n".encode(), tail
        assert err.strip() == "Error reading record: fastx: invalid FASTA/Q format", tail
    # a bad '+' line in the middle of the text: everything before that record is kept
    rc, out, err = run_env(["--nofilename", "--headersonly", "-", "-"], {},
                           stdin=b"@r1\nACTG\n+\nIIII\n@r2\nACTG\nX\nIIII\n@r3\nACTG\n+\nIIII\n")
    assert rc == 1 and out == f"{t};r1\n".encode() and "invalid FASTA/Q format" in err


@pytest.mark.gpu
def test_binary_small_chunks_pipeline_matches_one_chunk(tmp_path):
    """The chunk pipeline (reader -> 2 slots per GPU -> ordered writer) with 256 KiB chunks — hundreds of seams,
    records carried across them — against the same file in one chunk; FASTA (wrapped) and FASTQ (quality lines that
    start with '@' and '+')."""
    import numpy as np
    recs, text = _records(20_000, seed=33, wrap=60)
    p = tmp_path / "a.fa"
    p.write_bytes(text)
    rc, one, err = run_env(["--hash", "sha1,xxh3", str(p), "-"], {})
    assert rc == 0, err
    rc, many, err = run_env(["--hash", "sha1,xxh3", str(p), "-"], {"SHB_CHUNK_KB": "256"})
    assert rc == 0, err
    assert many == one
    r = np.random.default_rng(4)
    fq = []
    for i in range(30_000):
        L = int(r.integers(20, 300))
        s = r.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=L).tobytes()
        q = r.choice(np.frombuffer(b"@+!I5", dtype=np.uint8), size=L).tobytes()
        fq.append(b"@q%d x\n%s\n+\n%s\n" % (i, s, q))
    fq = b"".join(fq)
    rc, one, err = run_env(["--hash", "md5", "-", "-"], {}, stdin=fq)
    assert rc == 0, err
    rc, many, err = run_env(["--hash", "md5", "-", "-"], {"SHB_CHUNK_KB": "64"}, stdin=fq)
    assert rc == 0, err
    assert many == one and one.count(b"\n") == 4 * 30_000


@pytest.mark.gpu
def test_binary_two_gpus_match_one(tmp_path):
    """SHB_GPUS=2: chunks alternate between two devices, output order = chunk order (byte-identical to one GPU)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    recs, text = _records(60_000, seed=12, wrap=80)
    p = tmp_path / "b.fa"
    p.write_bytes(text)
    rc, one, err = run_env(["--hash", "sha1,blake3,nthash", str(p), "-"], {"SHB_GPUS": "1", "SHB_CHUNK_KB": "4096"})
    assert rc == 0, err
    rc, two, err = run_env(["--hash", "sha1,blake3,nthash", str(p), "-"], {"SHB_GPUS": "2", "SHB_CHUNK_KB": "4096"})
    assert rc == 0, err
    assert two == one


def test_binary_open_errors_use_the_errno_text(tmp_path):
    d = tmp_path / "dir.fa"
    d.mkdir()
    rc, _, err = run([str(tmp_path / "missing.fa")])
    assert rc == 1 and err.strip().endswith("no such file or directory")
    rc, _, err = run(["test.fasta", str(tmp_path / "nodir" / "out.fa")])
    assert rc == 1 and err.startswith("Error opening output: open ") and err.strip().endswith("no such file or directory")
    rc, _, err = run(["test.fasta", str(d)])
    assert rc == 1 and err.strip().endswith("is a directory")


@pytest.mark.gpu
def test_binary_refuses_nonascii_sequence_bytes():
    data = b">ok\nACGT\n>bad\nAC\xc3\xa9GT\n"
    rc, out, err = run_env(["--nofilename", "--headersonly", "-", "-"], {}, stdin=data)
    assert rc == 1 and "non-ASCII bytes in sequence data are not supported" in err
    rc, out, err = run_env(["--nofilename", "--headersonly", "-", "-"], {"SHB_ALLOW_NONASCII": "1"}, stdin=data)
    import hashlib
    assert rc == 0 and out == (hashlib.sha1(b"ACGT").hexdigest() + ";ok\n" + hashlib.sha1(b"AC\xc3\xa9GT").hexdigest() + ";bad\n").encode()
