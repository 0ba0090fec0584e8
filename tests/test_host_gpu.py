"""The reference's own processSequences / getHashFunc / binary test cases, restated against the host mirror
running on the GPU (seqhasher_test.go:248-343, :346-409, :801-936; test/test2_binary.sh:10-186)."""
import io
import os

import pytest

import golden_vectors as G

pytestmark = pytest.mark.gpu

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")
TEST_SEQUENCES = b">seq1\nACTG\n>seq1_lowercase\nactg\n>seq2\nTGCA\n"   # seqhasher_test.go:18
S1 = {"ACTG": "65c89f59d38cdbf90dfaf0b0a6884829df8396b0", "AAAA": "e2512172abf8cc9f67fdd49eb6cacf2df71bbad3",
      "TGCA": "e3da52abc8fbdb38b113a187ed0ac763fa86d1d4", "aaaa": "70c881d4a26984ddce795f6f71817c9cf4480e79"}


@pytest.fixture(scope="module")
def host():
    from seqhasher_b200 import host as H
    return H


def _run(host, data, **kw):
    cfg = host.Config(**kw)
    out = io.BytesIO()
    host.process_sequences(io.BytesIO(data), out, cfg)
    return out.getvalue().decode()


def _cli(host, argv, stdin=b""):
    class W(io.StringIO):
        pass
    w = W()
    w.buffer = io.BytesIO()
    cwd = os.getcwd()
    os.chdir(FIX)
    try:
        host.run(argv, w, stdin=io.BytesIO(stdin))
    finally:
        os.chdir(cwd)
    return w.getvalue() + w.buffer.getvalue().decode()


@pytest.mark.parametrize("algo", sorted(G.ACTG))
def test_get_hash_func(host, algo):
    # TestGetHashFunc, seqhasher_test.go:311-343
    assert host.get_hash_func(algo)(b"ACTG") == G.ACTG[algo]
    assert host.get_hash_func(algo)(b"") == ""


def test_get_hash_func_default_arm(host):
    assert host.get_hash_func(" md5")(b"ACTG") == G.ACTG["sha1"]   # seqhasher.go:344-346


def test_process_sequences_cases(host):
    # TestProcessSequences, seqhasher_test.go:248-308
    assert _run(host, TEST_SEQUENCES, hash_types=["sha1"], input_file_name="test.fasta") == (
        f">test.fasta;{S1['ACTG']};seq1\nACTG\n>test.fasta;{S1['ACTG']};seq1_lowercase\nACTG\n"
        f">test.fasta;{S1['TGCA']};seq2\nTGCA\n")
    assert _run(host, TEST_SEQUENCES, headers_only=True, hash_types=["md5"], no_file_name=True,
                input_file_name="test.fasta") == (
        "86bfb9f78dd8b6cd35962bb7324fdbf8;seq1\n86bfb9f78dd8b6cd35962bb7324fdbf8;seq1_lowercase\n"
        "5c15f97a88433c48f8bf76745d9da437;seq2\n")
    assert _run(host, TEST_SEQUENCES, hash_types=["nthash"], input_file_name="test.fasta") == (
        ">test.fasta;508876b331232519;seq1\nACTG\n>test.fasta;508876b331232519;seq1_lowercase\nACTG\n"
        ">test.fasta;95cecc5106c8fccd;seq2\nTGCA\n")


def test_invalid_sequence_hashed_as_is(host):
    # seqhasher_test.go:801-839
    out = _run(host, b">seq1\nACTGINVALID\n", hash_types=["sha1"], input_file_name="x.fa", headers_only=True)
    assert out.startswith("x.fa;3e06752b") and out.endswith(";seq1\n")


def test_fastq_and_fasta_after_fastq(host):
    # seqhasher_test.go:841-936
    fq = open(os.path.join(FIX, "test3.fastq"), "rb").read()
    assert _run(host, fq, hash_types=["sha1"], input_file_name="-") == (
        f"@{S1['ACTG']};seq1\nACTG\n+\nDFGH\n@{S1['AAAA']};seq2\nAAAA\n+\nBBBB\n@{S1['TGCA']};seq3\nTGCA\n+\nEEEE\n")
    assert _run(host, b">s\nACTG\n", hash_types=["sha1"], no_file_name=True) == f">{S1['ACTG']};s\nACTG\n"
    with pytest.raises(RuntimeError) as e:
        _run(host, b"garbage\n", hash_types=["sha1"])
    assert "fastx: invalid FASTA/Q format" in str(e.value)


def test_binary_script_cases(host):
    # test/test2_binary.sh:10-109, run from the fixture directory like the script runs from test/
    a, t, l = S1["AAAA"], S1["ACTG"], S1["aaaa"]
    assert _cli(host, ["test2.fasta", "-"]) == (
        f">test2.fasta;{a};seq1\nAAAA\n>test2.fasta;{t};seq2\nACTG\n>test2.fasta;{a};seq3\nAAAA\n")
    assert _cli(host, ["--name", "custom_name", "test2.fasta", "-"]).startswith(f">custom_name;{a};seq1\nAAAA\n")
    assert _cli(host, ["--headersonly", "test2.fasta", "-"]) == (
        f"test2.fasta;{a};seq1\ntest2.fasta;{t};seq2\ntest2.fasta;{a};seq3\n")
    assert _cli(host, ["--headersonly", "--nofilename", "test2.fasta", "-"]) == f"{a};seq1\n{t};seq2\n{a};seq3\n"
    assert _cli(host, ["--headersonly", "--nofilename", "--hash", "xxhash", "--casesensitive", "test2.fasta", "-"]) == (
        "cf40b5b72bc43e77;seq1\n704b34bf20faedf2;seq2\n42a70d1abf84bf32;seq3\n")
    assert _cli(host, ["--headersonly", "--nofilename", "--hash", "sha1,xxhash", "--casesensitive", "test2.fasta", "-"]) == (
        f"{a};cf40b5b72bc43e77;seq1\n{t};704b34bf20faedf2;seq2\n{l};42a70d1abf84bf32;seq3\n")


@pytest.mark.parametrize("ext", ["bz2", "gz", "xz", "zst"])
def test_compressed_inputs(host, ext):
    # test/test2_binary.sh:112-128, seqhasher_test.go:346-409
    a, t = S1["AAAA"], S1["ACTG"]
    assert _cli(host, ["--headersonly", "--nofilename", f"test2.fasta.{ext}", "-"]) == f"{a};seq1\n{t};seq2\n{a};seq3\n"


def test_whitespace_stripping(host):
    # test/test2_binary.sh:131-145
    t = S1["ACTG"]
    names = ["seq1_uppercase", "seq2_lowercase", "seq3_mixed_case", "seq4_with_spaces", "seq5_with_tabs", "seq6_with_newline"]
    assert _cli(host, ["--hash", "sha1", "--nofilename", "seqs_with_spaces.fasta", "-"]) == "".join(
        f">{t};{n}\nACTG\n" for n in names)


def test_long_sequences_and_stdin_fastq(host):
    # test/test2_binary.sh:148-186
    data = b">seq1\n" + b"ACTG" * 25000 + b"\n>seq2\n" + b"A" * 100000 + b"\n"
    assert _run(host, data, headers_only=True, no_file_name=True, hash_types=["sha1"]) == (
        "3200abf5034212fb5d90fa242fc1cea5e0f3dd00;seq1\ne27b41eb7f7370b36f1b2fa63f05652f52423277;seq2\n")
    fq = open(os.path.join(FIX, "test3.fastq"), "rb").read()
    assert _cli(host, ["-", "-"], stdin=fq).startswith(f"@{S1['ACTG']};seq1\nACTG\n+\nDFGH\n")


def test_problematic_sequences_and_empty_record(host, oracle):
    # test/problematic_sequences.fasta is unused by the reference's tests; the empty record must give an
    # empty hash field (seqhasher.go:292-295) and the others are hashed as-is after strip + upper
    out = _cli(host, ["--headersonly", "--hash", "sha1,xxh3", "problematic_sequences.fasta", "-"]).splitlines()
    want_seqs = [b"ACTG123A", b"ACTGXYZ", b"ACGTRYSWKMBDHVN", b"AC!@#$%^&*()TG"]
    for line, s in zip(out, want_seqs):
        f = line.split(";")
        assert f[1] == oracle.get_hash_func("sha1")(s) and f[2] == oracle.get_hash_func("xxh3")(s)
    assert out[4] == "problematic_sequences.fasta;;;seq_empty"
