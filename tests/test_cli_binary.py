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
