"""Host-side mirror, CPU-only parts: parseFlags / isValidHashType (seqhasher_test.go:65-173) and the reader
semantics the GPU path relies on (seqhasher_test.go:346-409 compressed twins)."""
import io
import os

import pytest

from seqhasher_b200.host import cli
from seqhasher_b200.host.fastx import ERR_INVALID, FastxError, read_records
from seqhasher_b200.host.hasher import SUPPORTED_HASH_TYPES, is_valid_hash_type

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")


def test_parse_flags_defaults_and_multi_hash():
    # seqhasher_test.go:65-140
    c = cli.parse_flags(["input.fasta"])
    assert c.hash_types == ["sha1"] and not c.headers_only and not c.no_file_name and not c.case_sensitive
    assert c.input_file_name == "input.fasta" and c.output_file_name == ""
    c = cli.parse_flags(["--hash", "sha1,md5,xxhash", "input.fasta", "out.fa"])
    assert c.hash_types == ["sha1", "md5", "xxhash"] and c.output_file_name == "out.fa"
    c = cli.parse_flags(["-o", "-n", "-c", "-f", "lbl", "-H=nthash", "-"])
    assert (c.headers_only, c.no_file_name, c.case_sensitive, c.name_override, c.hash_types, c.input_file_name) == \
           (True, True, True, "lbl", ["nthash"], "-")


def test_parse_flags_invalid_hash_error_text():
    # exact text pinned at seqhasher_test.go:112
    with pytest.raises(ValueError) as e:
        cli.parse_flags(["--hash", "invalid", "input.fasta"])
    assert str(e.value) == ("Invalid hash type: invalid. Supported types are: sha1, sha3, md5, xxhash, xxh3, cityhash, "
                            "murmur3, nthash, blake3, k12, rapidhash, rapidhash32")


def test_parse_flags_go_flag_quirks():
    # flags after the first positional are positionals; "sha1, md5" validates (TrimSpace) but is stored untrimmed
    c = cli.parse_flags(["in.fa", "--headersonly"])
    assert c.output_file_name == "--headersonly" and not c.headers_only
    c = cli.parse_flags(["--hash", "sha1, md5", "in.fa"])
    assert c.hash_types == ["sha1", " md5"]
    with pytest.raises(cli.FlagError):
        cli.parse_flags(["--bogus", "in.fa"])
    with pytest.raises(cli.HelpRequested):
        cli.parse_flags(["-h"])


def test_is_valid_hash_type():
    # seqhasher_test.go:143-173
    for t in SUPPORTED_HASH_TYPES:
        assert is_valid_hash_type(t)
    assert not is_valid_hash_type("invalid") and not is_valid_hash_type("")


def test_version_and_usage_short_circuit():
    # seqhasher_test.go:548-660: -version prints and returns; no args prints the usage and returns
    w = io.StringIO()
    cli.run(["-version"], w)
    assert w.getvalue() == "SeqHasher 1.2.0\n"
    w = io.StringIO()
    cli.run([], w)
    assert "Usage:" in w.getvalue() and "Supported hash types: sha1, sha3" in w.getvalue()
    with pytest.raises(RuntimeError) as e:
        cli.run(["/nonexistent/in.fa"], io.StringIO())
    assert str(e.value).startswith("Error opening input: open /nonexistent/in.fa: no such file or directory")


def _recs(name):
    with open(os.path.join(FIX, name), "rb") as f:
        return [(r.name, r.seq, r.qual) for r in read_records(f)]


def test_reader_fixtures():
    assert _recs("test.fasta") == [(b"seq1", b"ACTG", None), (b"seq1_lowercase", b"actg", None), (b"seq2", b"TGCA", None)]
    assert _recs("test3.fastq") == [(b"seq1", b"ACTG", b"DFGH"), (b"seq2", b"AAAA", b"BBBB"), (b"seq3", b"TGCA", b"EEEE")]
    sp = _recs("seqs_with_spaces.fasta")
    assert sp[3][1] == b"AC TG " and sp[4][1] == b"AC\tTG" and sp[5][1] == b"ACTG"   # line breaks joined, blanks kept
    assert _recs("problematic_sequences.fasta")[-1] == (b"seq_empty", b"", None)


@pytest.mark.parametrize("ext", ["gz", "bz2", "xz", "zst"])
def test_reader_compressed_twins(ext):
    # seqhasher_test.go:355-360: the four compressed fixtures decode to the plain records
    assert _recs(f"test.fasta.{ext}") == _recs("test.fasta")
    assert _recs(f"test2.fasta.{ext}") == _recs("test2.fasta")


def test_reader_rejects_non_fastx():
    # seqhasher_test.go:770-799
    with pytest.raises(FastxError) as e:
        list(read_records(io.BytesIO(b"this is not a sequence file\n")))
    assert str(e.value) == ERR_INVALID
    assert list(read_records(io.BytesIO(b""))) == []
