"""oracle/reader.py (the restated bio fastx reader + processSequences formatting) against every reader-level golden
the reference holds, and against the product's host-side reader mirror on quirky text (the two are written
independently; the GPU splitter is tested against the oracle one in test_textproc_gpu.py)."""
import io
import os

import numpy as np
import pytest

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")
S1 = {"ACTG": "65c89f59d38cdbf90dfaf0b0a6884829df8396b0", "AAAA": "e2512172abf8cc9f67fdd49eb6cacf2df71bbad3",
      "TGCA": "e3da52abc8fbdb38b113a187ed0ac763fa86d1d4"}


def fx(name):
    return open(os.path.join(FIX, name), "rb").read()


def test_reference_goldens(oracle):
    from oracle import reader as R
    t, a, g = S1["ACTG"], S1["AAAA"], S1["TGCA"]
    # seqhasher_test.go:263-265
    out, err = R.process_sequences(fx("test.fasta"), ["sha1"], input_file_name="test.fasta")
    assert err is None and out == (f">test.fasta;{t};seq1\nACTG\n>test.fasta;{t};seq1_lowercase\nACTG\n>test.fasta;{g};seq2\nTGCA\n").encode()
    # test2_binary.sh: headers only, no file name
    out, _ = R.process_sequences(fx("test2.fasta"), ["sha1"], headers_only=True, no_file_name=True)
    assert out == f"{a};seq1\n{t};seq2\n{a};seq3\n".encode()
    # test2_binary.sh:131-145 — blanks, tabs, a wrapped line, mixed case all hash as ACTG
    out, _ = R.process_sequences(fx("seqs_with_spaces.fasta"), ["sha1"], no_file_name=True)
    names = ["seq1_uppercase", "seq2_lowercase", "seq3_mixed_case", "seq4_with_spaces", "seq5_with_tabs", "seq6_with_newline"]
    assert out == "".join(f">{t};{n}\nACTG\n" for n in names).encode()
    # test3.fastq (last line without a line break), test2_binary.sh:172-186; stdin => no file name field
    out, _ = R.process_sequences(fx("test3.fastq"), ["sha1"])
    assert out == f"@{t};seq1\nACTG\n+\nDFGH\n@{a};seq2\nAAAA\n+\nBBBB\n@{g};seq3\nTGCA\n+\nEEEE\n".encode()
    # compressed twins decode to the same records
    for ext in ("gz", "bz2", "xz"):
        assert R.process_sequences(fx("test2.fasta." + ext), ["sha1"], headers_only=True, no_file_name=True)[0] == f"{a};seq1\n{t};seq2\n{a};seq3\n".encode()
    # seqhasher_test.go:780-799
    assert R.process_sequences(b"this is not a sequence file\n", ["sha1"]) == (b"", "Failed to create reader: fastx: invalid FASTA/Q format")
    assert R.process_sequences(b"", ["sha1"]) == (b"", None)
    # header-only record: empty hash fields
    out, _ = R.process_sequences(fx("problematic_sequences.fasta"), ["sha1", "md5"], headers_only=True, input_file_name="p")
    assert out.splitlines()[-1] == b"p;;;seq_empty"


def test_truncated_fastq(oracle):
    from oracle import reader as R
    for tail in (b"@r2\nAC", b"@r2\nACGT\n", b"@r2\nACGT\n+\n", b"@r2\nACGT\n+", b"@r2\nACGT\nX\nIIII\n"):
        out, err = R.process_sequences(b"@r1\nACTG\n+\nIIII\n" + tail, ["sha1"], no_file_name=True)
        assert out == f"@{S1['ACTG']};r1\nACTG\n+\nIIII\n".encode() and err == "Error reading record: fastx: invalid FASTA/Q format"


def test_agrees_with_the_host_reader_mirror(oracle):
    """Two independent restatements of the same reader (oracle/reader.py and seqhasher_b200/host/fastx.py) on quirky
    text: CRLF, blank lines, blanks inside sequence lines, '>' / '@' inside names, quality lines starting with '@' / '+'."""
    from oracle import reader as R
    from seqhasher_b200.host.fastx import read_records
    rng = np.random.default_rng(3)
    for trial in range(200):
        eol = b"\r\n" if trial % 4 == 0 else b"\n"
        recs = []
        for i in range(int(rng.integers(1, 30))):
            L = int(rng.integers(0, 200))
            seq = rng.choice(np.frombuffer(b"ACGTacgtN-*", dtype=np.uint8), size=L).tobytes()
            name = b"r%d d=1 >x @y" % i
            if trial & 1:
                L = max(L, 1)
                seq = seq or b"A"
                q = rng.choice(np.frombuffer(b"@+I!5", dtype=np.uint8), size=len(seq)).tobytes()
                recs.append(b"@" + name + eol + seq + eol + b"+" + eol + q + eol)
            else:
                w = int(rng.choice([0, 7, 60]))
                lines = [seq[k:k + w] for k in range(0, L, w)] if w and L else [seq]
                recs.append(b">" + name + eol + eol.join(ln[:3] + b" \t" + ln[3:] for ln in lines) + eol + (eol if rng.random() < 0.2 else b""))
        text = (eol if trial % 5 == 0 else b"") + b"".join(recs)
        if trial % 3 == 0:
            text = text.rstrip(b"\r\n")
        mine = [(r.name, r.seq, r.qual or None) for r in R.records(text)]
        theirs = [(r.name, r.seq, r.qual or None) for r in read_records(io.BytesIO(text))]
        assert mine == theirs, trial
