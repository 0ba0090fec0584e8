"""The whole record loop on the GPU (shb_textproc: split -> strip -> upper -> hash -> header rewrite -> format)
against a CPU expectation built from the oracle's reader restatement (oracle/reader.py) + the oracle hashes, on the reference's fixtures
and on synthetic multi-line FASTA / 4-line FASTQ that span many 4 KiB scan blocks and several chunks."""
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")


@pytest.fixture(scope="module")
def H():
    from seqhasher_b200 import host
    return host


def expected(oracle, data: bytes, cfg) -> bytes:
    """seqhasher.go:210-286 on the CPU, entirely from oracle/: the reader restatement (oracle/reader.py, with its table of
    which reference test pins which rule) + the oracle hashes + the reference's formatting.  No product code."""
    from oracle import reader
    out, err = reader.process_sequences(data, cfg.hash_types, cfg.headers_only, cfg.no_file_name, cfg.case_sensitive,
                                        cfg.input_file_name, cfg.name_override)
    assert err is None, err
    return out


def run_gpu(H, data: bytes, cfg, chunk=1 << 20) -> bytes:
    from seqhasher_b200 import binding as B
    tp = B.TextProc(chunk)
    try:
        out = io.BytesIO()
        H.process_sequences_gpu(io.BytesIO(data), out, cfg, textproc=tp, log=io.StringIO())
        return out.getvalue()
    finally:
        tp.close()


CFGS = [
    dict(hash_types=["sha1"], input_file_name="test.fasta"),
    dict(hash_types=["md5"], headers_only=True, no_file_name=True, input_file_name="x"),
    dict(hash_types=["sha1", "xxhash", "nthash"], case_sensitive=True, input_file_name="-"),
    dict(hash_types=["xxh3", "blake3", "k12", "rapidhash32"], name_override="lbl", input_file_name="in.fa", headers_only=True),
]


@pytest.mark.parametrize("fixture", ["test.fasta", "test2.fasta", "test3.fastq", "seqs_with_spaces.fasta",
                                     "problematic_sequences.fasta", "test.fasta.gz", "test2.fasta.zst", "test.fasta.xz",
                                     "test2.fasta.bz2"])
def test_fixtures(H, oracle, fixture):
    data = open(os.path.join(FIX, fixture), "rb").read()
    import gzip, bz2, lzma
    for kw in CFGS:
        cfg = H.Config(**kw)
        got = run_gpu(H, data, cfg)
        plain = b"".join(_plain(data))
        assert got == expected(oracle, plain, cfg), (fixture, kw)


def _plain(data):
    from seqhasher_b200.host.fastx import open_maybe_compressed
    yield open_maybe_compressed(io.BytesIO(data)).read()


def synth_fasta(n, seed, crlf=False, wrap=60):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTacgtNRY", dtype=np.uint8)
    eol = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n):
        L = int(rng.integers(0, 3000)) if i % 37 else 0
        seq = rng.choice(alphabet, size=L).tobytes()
        name = f"seq_{i} desc with spaces len={L}".encode() if i % 5 else b"s%d" % i
        out.append(b">" + name + eol)
        w = wrap if i % 3 else 10 ** 9
        for k in range(0, L, w):
            line = seq[k:k + w]
            if i % 11 == 0 and len(line) > 4:
                line = line[:2] + b" \t" + line[2:]     # interior blanks: stripped by seqhasher.go:246
            out.append(line + eol)
        if i % 13 == 0:
            out.append(eol)                              # blank line inside the file
    return b"".join(out)


def synth_fastq(n, seed, crlf=False):
    rng = np.random.default_rng(seed)
    eol = b"\r\n" if crlf else b"\n"
    qa = np.frombuffer(bytes(range(33, 74)), dtype=np.uint8)    # includes '@' (64) and '+' (43)
    out = []
    for i in range(n):
        L = int(rng.integers(1, 400))
        seq = rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), size=L).tobytes()
        qual = rng.choice(qa, size=L).tobytes()
        if i % 7 == 0:
            qual = b"@" + qual[1:]
        if i % 9 == 0:
            qual = b"+" + qual[1:]
        out.append(b"@read_%d some/1" % i + eol + seq + eol + b"+" + eol + qual + eol)
    return b"".join(out)


@pytest.mark.parametrize("crlf", [False, True])
def test_multiline_fasta_chunked(H, oracle, crlf):
    data = b"\n\n" + synth_fasta(900, 3, crlf) + b"\n\n"
    for kw in (CFGS[0], CFGS[1], dict(hash_types=["xxh3", "murmur3"], input_file_name="f.fa")):
        cfg = H.Config(**kw)
        want = expected(oracle, data, cfg)
        assert run_gpu(H, data, cfg, chunk=1 << 22) == want           # one chunk
        assert run_gpu(H, data, cfg, chunk=40_000) == want            # many chunks, records carried over
    # no trailing line break at EOF, header-only last record
    data2 = synth_fasta(50, 4) + b">last_empty"
    cfg = H.Config(hash_types=["sha1"], input_file_name="z")
    assert run_gpu(H, data2, cfg, chunk=9000) == expected(oracle, data2, cfg)


@pytest.mark.parametrize("crlf", [False, True])
def test_fastq_chunked(H, oracle, crlf):
    data = synth_fastq(3000, 8, crlf)
    for kw in (dict(hash_types=["sha1"], input_file_name="-"), dict(hash_types=["xxhash", "xxh3", "murmur3"], headers_only=True,
                                                                       input_file_name="r.fq")):
        cfg = H.Config(**kw)
        want = expected(oracle, data, cfg)
        assert run_gpu(H, data, cfg, chunk=1 << 22) == want
        assert run_gpu(H, data, cfg, chunk=30_000) == want
    nolf = data.rstrip()                                               # last quality line without a line break
    cfg = H.Config(hash_types=["md5"], input_file_name="q")
    assert run_gpu(H, nolf, cfg, chunk=50_000) == expected(oracle, nolf, cfg)


def test_errors_and_empty(H):
    cfg = H.Config(hash_types=["sha1"], input_file_name="x")
    with pytest.raises(RuntimeError) as e:
        run_gpu(H, b"this is not a sequence file\n", cfg)
    assert str(e.value) == "Failed to create reader: fastx: invalid FASTA/Q format"   # seqhasher_test.go:795
    assert run_gpu(H, b"", cfg) == b"" and run_gpu(H, b"\n \n", cfg) == b""
    with pytest.raises(RuntimeError):
        run_gpu(H, b"@r1\nACGT\nX\nIIII\n", cfg)                       # third line must start with '+'


def test_empty_sequence_log_lines(H):
    from seqhasher_b200 import binding as B
    tp = B.TextProc(1 << 16)
    log = io.StringIO()
    out = io.BytesIO()
    H.process_sequences_gpu(io.BytesIO(b">a\n\n>b\nAC\n>c\n"), out, H.Config(hash_types=["sha1", "md5"], input_file_name="f"),
                            textproc=tp, log=log)
    tp.close()
    assert out.getvalue().startswith(b">f;;;a\n\n>f;") and out.getvalue().endswith(b">f;;;c\n\n")
    assert log.getvalue().count("Empty DNA sequence") == 4            # 2 empty records x 2 hash types


def test_large_text_uses_tiled_scans(H, oracle):
    # > 4096 scan blocks (16 MiB of text): exercises the tiled (max, sum) block scan and the tiled count scans
    data = synth_fasta(14000, 21)
    assert len(data) > 17 << 20
    cfg = H.Config(hash_types=["xxh3"], headers_only=True, input_file_name="big")
    got = run_gpu(H, data, cfg, chunk=32 << 20)
    assert got == expected(oracle, data, cfg)
    fq = synth_fastq(90000, 22)
    assert len(fq) > 17 << 20
    got = run_gpu(H, fq, cfg, chunk=32 << 20)
    assert got == expected(oracle, fq, cfg)


def test_record_larger_than_chunk_grows_buffer(H, oracle):
    # a 300 kB record with a 64 kB text buffer: the host loop must grow the buffer instead of failing
    rng = np.random.default_rng(5)
    big = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=300_000).tobytes()
    data = b">small\nACGT\n>chrBig some description\n" + b"\n".join(big[i:i + 70] for i in range(0, len(big), 70)) + b"\n>tail\nTTGA\n"
    cfg = H.Config(hash_types=["sha1", "blake3"], headers_only=True, input_file_name="g.fa")
    out = io.BytesIO()
    H.process_sequences_gpu(io.BytesIO(data), out, cfg, chunk_bytes=1 << 16, log=io.StringIO())
    assert out.getvalue() == expected(oracle, data, cfg)


def _random_text(rng, fastq: bool) -> bytes:
    eol = b"\r\n" if rng.random() < 0.25 else b"\n"
    n = int(rng.integers(1, 400))
    alpha = np.frombuffer(b"ACGTacgtNn" if rng.random() < 0.7 else b"ACGTacgtNnRYKM-*.", dtype=np.uint8)
    out = [eol * int(rng.integers(0, 3))] if rng.random() < 0.3 else []
    for i in range(n):
        L = int(rng.choice([0, 1, 5, 59, 60, 61, 150, 700, 4100])) if rng.random() < 0.5 else int(rng.integers(0, 900))
        if fastq:
            L = max(L, 1)      # an empty FASTQ record is two blank lines: not a case the reference defines
        seq = alpha[rng.integers(0, len(alpha), size=L)].tobytes()
        name = b"r%d" % i + (b" some description; x=1" if rng.random() < 0.5 else b"") + (b" >not a record @either" if rng.random() < 0.1 else b"")
        if fastq:
            qual = bytes(rng.integers(33, 74, size=L).astype(np.uint8))
            out.append(b"@" + name + eol + seq + eol + b"+" + (name if rng.random() < 0.2 else b"") + eol + qual + eol)
        else:
            wrap = int(rng.choice([0, 60, 70, 7]))
            lines = [seq[k:k + wrap] for k in range(0, L, wrap)] if wrap and L else [seq]
            if rng.random() < 0.1 and L > 4:
                lines = [ln[:2] + b" " + ln[2:] for ln in lines]       # blanks inside sequence lines are stripped
            body = eol.join(lines)
            out.append(b">" + name + eol + body + eol + (eol if rng.random() < 0.05 else b""))
    text = b"".join(out)
    if rng.random() < 0.5:
        text = text.rstrip(b"\r\n")                                      # last line without a line break
    return text


@pytest.mark.parametrize("seed", [1, 2, 3] + [int(x) for x in os.environ.get("SHB_SOAK_SEEDS", "").split(",") if x])
def test_random_text_soak(H, oracle, seed):
    """Random FASTA / FASTQ text with the quirks bio's reader accepts (CRLF, wrapped and blank lines, blanks inside
    sequence lines, '>' and '@' inside headers, empty sequences, missing final line break), random chunk sizes that
    cut records anywhere, random configurations."""
    rng = np.random.default_rng(seed)
    algos = ["sha1", "md5", "xxh3", "xxhash", "murmur3", "nthash", "blake3", "rapidhash", "cityhash", "k12", "sha3", "rapidhash32"]
    for it in range(14):
        fastq = bool(rng.integers(2))
        data = _random_text(rng, fastq)
        kw = dict(hash_types=[algos[j] for j in rng.choice(12, size=int(rng.integers(1, 4)), replace=False)],
                  headers_only=bool(rng.integers(2)), case_sensitive=bool(rng.integers(2)), no_file_name=bool(rng.integers(2)),
                  input_file_name="in.fa")
        cfg = H.Config(**kw)
        chunk = int(rng.choice([12000, 20000, 70000, 1 << 20]))   # every record (<= 8.5 kB of text) fits
        got = run_gpu(H, data, cfg, chunk=chunk)
        want = expected(oracle, data, cfg)
        assert got == want, (seed, it, fastq, kw, chunk, len(data))


def test_more_records_than_the_chunk_index(H, oracle):
    """A 64 KiB text buffer indexes 64 Ki / 16 + 1024 = 5120 records; 30 000 six-byte records per chunk make every
    chunk overflow it, so each is consumed over several runs (used to fail with SHB_ERR_CAPACITY)."""
    data = b"".join(b">%d\n%s\n" % (i % 7, (b"A", b"CG", b"T", b"")[i & 3]) for i in range(30_000))
    cfg = H.Config(hash_types=["md5"], headers_only=True, no_file_name=True, input_file_name="t")
    assert run_gpu(H, data, cfg, chunk=1 << 16) == expected(oracle, data, cfg)
    fq = b"".join(b"@%d\n%s\n+\n%s\n" % (i % 7, b"ACGT"[: 1 + (i & 3)], b"IIII"[: 1 + (i & 3)]) for i in range(20_000))
    assert run_gpu(H, fq, cfg, chunk=1 << 16) == expected(oracle, fq, cfg)


@pytest.mark.parametrize("tail", [b"@r2\nAC", b"@r2\nACGT\n", b"@r2\nACGT\n+\n", b"@r2\nACGT\n+"])
def test_truncated_fastq_is_an_error_after_the_complete_records(H, tail):
    """a FASTQ input that ends inside a record lost that record silently (rc 0)."""
    from seqhasher_b200 import binding as B
    cfg = H.Config(hash_types=["sha1"], no_file_name=True, input_file_name="q")
    out = io.BytesIO()
    tp = B.TextProc(1 << 16)
    try:
        with pytest.raises(RuntimeError) as e:
            H.process_sequences_gpu(io.BytesIO(b"@r1\nACTG\n+\nIIII\n" + tail), out, cfg, textproc=tp, log=io.StringIO())
    finally:
        tp.close()
    assert str(e.value) == "Error reading record: fastx: invalid FASTA/Q format"
    assert out.getvalue() == b"@65c89f59d38cdbf90dfaf0b0a6884829df8396b0;r1\nACTG\n+\nIIII\n"
    # the host reader mirror agrees
    from seqhasher_b200.host.fastx import FastxError, read_records
    with pytest.raises(FastxError):
        list(read_records(io.BytesIO(b"@r1\nACTG\n+\nIIII\n" + tail)))


def test_nonascii_sequence_bytes_are_refused(H, monkeypatch):
    """seqhasher.go:243-251: the reference normalises a record that holds a byte >= 0x80 with Unicode rules; this engine
    is byte-wise ASCII, so both text paths refuse such input (records before it are still written) unless
    SHB_ALLOW_NONASCII is set — never a silently different digest."""
    data = b">ok\nACGT\n>bad\nAC\xc3\xa9GT\n>after\nTTTT\n"
    cfg = H.Config(hash_types=["sha1"], no_file_name=True, headers_only=True, input_file_name="x")
    monkeypatch.delenv("SHB_ALLOW_NONASCII", raising=False)
    with pytest.raises(RuntimeError) as e:
        run_gpu(H, data, cfg)
    assert "non-ASCII bytes in sequence data" in str(e.value)
    out = io.BytesIO()
    with pytest.raises(RuntimeError) as e:
        H.process_sequences(io.BytesIO(data), out, cfg)              # the host-side record loop
    assert "non-ASCII bytes in sequence data" in str(e.value)
    assert out.getvalue().startswith(b"2108994e17f6cca9f1e3f1c8a4e7a47c2c6d8c9b"[:0])   # records before it are flushed
    # non-ASCII in NAMES is fine (names are copied, never normalised)
    named = ">caf\xc3\xa9 1\nACTG\n".encode("latin-1")
    assert run_gpu(H, named, cfg) == b"65c89f59d38cdbf90dfaf0b0a6884829df8396b0;caf\xc3\xa9 1\n"
    monkeypatch.setenv("SHB_ALLOW_NONASCII", "1")
    import hashlib
    want = b"".join(hashlib.sha1(s).hexdigest().encode() + b";" + n + b"\n" for n, s in ((b"ok", b"ACGT"), (b"bad", b"AC\xc3\xa9GT"), (b"after", b"TTTT")))
    assert run_gpu(H, data, cfg) == want                              # opaque bytes: ASCII-only upper-casing


def test_batch_check_ascii_flag():
    from seqhasher_b200 import binding as B
    res = np.frombuffer(b"ACGT" * 1000 + b"AC\xffT" + b"ACGT" * 10, dtype=np.uint8)
    offs = np.array([0, 4000, 4004, 4044], dtype=np.int64)
    b = B.Batch(len(res), 3)
    try:
        b.load(res, offs)
        b.submit(["xxh3"], B.FLAG_UPPER | B.FLAG_CHECK_ASCII)
        b.wait()
        assert b.nonascii() == 1
        b.residues[4002] = ord("G")
        b.submit(["xxh3"], B.FLAG_UPPER | B.FLAG_CHECK_ASCII)
        b.wait()
        assert b.nonascii() == 0
        b.residues[4043] = 0x80                                      # the very last byte of the batch
        b.submit(["xxh3"], B.FLAG_CHECK_ASCII)
        b.wait()
        assert b.nonascii() == 1
    finally:
        b.close()
