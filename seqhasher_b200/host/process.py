"""processSequences (seqhasher.go:210-286) with the per-record hash statements replaced by batched GPU calls.
Header layout, --headersonly and FASTA/FASTQ formatting follow seqhasher.go:261-282."""
from __future__ import annotations

import io
import os

from .fastx import FastxError, read_records
from .hasher import GpuHasher, shared_hasher

# seqhasher.go:243-251: a record with a byte >= 0x80 is normalised with Unicode rules by the reference (UTF-8 decoding, Unicode
# spaces, rune-wise upper-casing with U+FFFD for invalid bytes); this engine is byte-wise ASCII, so it refuses such input rather
# than print digests that may differ (the reference's README.md:259 declares non-ASCII input unsupported)
ERR_NONASCII = "non-ASCII bytes in sequence data are not supported (set SHB_ALLOW_NONASCII=1 to hash them as plain bytes)"
BATCH_RECORDS = 1 << 18
BATCH_BYTES = 48 << 20


def process_sequences(input_stream, output_stream, cfg, hasher: GpuHasher | None = None) -> None:
    """Raises RuntimeError("Failed to create reader: ...") / ("Error reading record: ...") like the reference."""
    hasher = hasher or shared_hasher()
    input_file_name = cfg.input_file_name
    no_file_name = cfg.no_file_name
    if cfg.name_override != "":
        input_file_name = cfg.name_override
    elif input_file_name == "-":
        no_file_name = True   # seqhasher.go:217-219
    out = io.BufferedWriter(output_stream) if isinstance(output_stream, io.RawIOBase) else output_stream
    file_label = input_file_name.encode()

    def flush(batch):
        if not batch:
            return
        hexes, norm = hasher.hash_records([r.seq for r in batch], cfg.hash_types, cfg.case_sensitive,
                                          want_seq=not cfg.headers_only)
        for idx, rec in enumerate(batch):
            hashes = ";".join(hexes[idx]).encode()
            if no_file_name:
                name = hashes + b";" + rec.name if cfg.hash_types else rec.name
            else:
                name = (file_label + b";" + hashes + b";" + rec.name) if cfg.hash_types else (file_label + b";" + rec.name)
            if cfg.headers_only:
                out.write(name + b"\n")
            elif rec.qual is not None and len(rec.qual) > 0:   # record.Format(0): FASTQ iff quality present
                out.write(b"@" + name + b"\n" + norm[idx] + b"\n+\n" + rec.qual + b"\n")
            else:
                out.write(b">" + name + b"\n" + norm[idx] + b"\n")

    try:
        reader = read_records(input_stream)
        batch, nbytes = [], 0
        first = True
        while True:
            try:
                rec = next(reader)
            except StopIteration:
                break
            except FastxError as e:
                if first:
                    raise RuntimeError(f"Failed to create reader: {e}") from None
                raise RuntimeError(f"Error reading record: {e}") from None
            first = False
            if not rec.seq.isascii() and not os.environ.get("SHB_ALLOW_NONASCII"):
                flush(batch)
                raise RuntimeError(f"Error reading record: {ERR_NONASCII}")
            batch.append(rec)
            nbytes += len(rec.seq)
            if len(batch) >= BATCH_RECORDS or nbytes >= BATCH_BYTES:
                flush(batch)
                batch, nbytes = [], 0
        flush(batch)
    finally:
        out.flush()


def process_sequences_gpu(input_stream, output_stream, cfg, textproc=None, chunk_bytes: int = 64 << 20, log=None) -> None:
    """processSequences with the WHOLE record loop on the GPU (shb_textproc): the host only decompresses and
    moves bytes.  Same observable behaviour as process_sequences (output text, error texts, one stderr line
    per empty record and hash type)."""
    import lzma
    import sys
    import zlib

    from .. import binding as B
    from .fastx import ERR_INVALID, open_maybe_compressed
    from .hasher import EMPTY_MSG, _algo_ids
    log = log or sys.stderr
    label = cfg.input_file_name
    no_file_name = cfg.no_file_name
    if cfg.name_override != "":
        label = cfg.name_override
    elif label == "-":
        no_file_name = True   # seqhasher.go:217-219
    own = textproc is None
    if len(cfg.hash_types) > 48:
        # the GPU formatter carries at most 48 hash fields; longer lists take the host-side loop
        return process_sequences(input_stream, output_stream, cfg)
    tp = textproc or B.TextProc(chunk_bytes)
    ids = _algo_ids(cfg.hash_types)
    flags = 0 if cfg.case_sensitive else B.FLAG_UPPER
    blanks = b" \t\n\v\f\r"

    def fail_read(e):
        return RuntimeError(f"Error reading record: {e}")

    try:
        f = open_maybe_compressed(input_stream)
        fill = 0          # bytes currently in tp.input
        fmt = 0
        eof = False
        while True:
            while not eof and fill < tp.capacity:
                try:
                    if hasattr(f, "readinto"):      # straight into the pinned buffer: no intermediate bytes object
                        got = f.readinto(memoryview(tp.input)[fill: tp.capacity])
                        if not got:
                            eof = True
                            break
                        fill += got
                    else:
                        chunk = f.read(tp.capacity - fill)
                        if not chunk:
                            eof = True
                            break
                        tp.input[fill: fill + len(chunk)] = memoryview(chunk)
                        fill += len(chunk)
                except (EOFError, OSError, ValueError, zlib.error, lzma.LZMAError) as e:   # damaged / truncated compressed input
                    raise fail_read(e) from None
            if fmt == 0:
                # sniff: skip leading blank space, first byte decides (fastx.NewReaderFromIO, seqhasher.go:221)
                skip = 0
                while skip < fill:
                    head = bytes(tp.input[skip: min(fill, skip + 4096)])
                    k = len(head) - len(head.lstrip(blanks))
                    skip += k
                    if k < len(head):
                        break
                if skip == fill:
                    if eof:
                        return          # empty input: no records
                    fill = 0
                    continue
                first = bytes(tp.input[skip: skip + 1])
                if first == b">":
                    fmt = B.FORMAT_FASTA
                elif first == b"@":
                    fmt = B.FORMAT_FASTQ
                else:
                    raise RuntimeError(f"Failed to create reader: {ERR_INVALID}")
                if skip:
                    tp.input[: fill - skip] = tp.input[skip:fill].copy()
                    fill -= skip
            if eof:
                # trailing blank lines are not records
                while fill:
                    tail = bytes(tp.input[max(0, fill - 4096): fill])
                    k = len(tail) - len(tail.rstrip(blanks))
                    fill -= k
                    if k < len(tail):
                        break
                if fill == 0:
                    return
            try:
                out, consumed, nrec, nempty = tp.run(fill, fmt, eof, ids, flags, cfg.headers_only,
                                                     None if no_file_name else label.encode())
            except B.SeqhashError as e:
                raise fail_read(ERR_INVALID if ERR_INVALID in str(e) else e) from None
            if tp.nonascii and not os.environ.get("SHB_ALLOW_NONASCII"):
                raise fail_read(ERR_NONASCII)
            for _ in range(nempty * len(ids)):
                print(EMPTY_MSG, file=log)
            output_stream.write(memoryview(out))
            if tp.partial_format:       # the records before the malformed one have been written
                raise fail_read(ERR_INVALID)
            if eof and consumed >= fill:
                return
            if consumed == 0:
                # one record does not fit the text buffer (e.g. a chromosome): grow the buffer and retry
                if eof or not own or tp.capacity >= (1 << 31):
                    raise RuntimeError("Error reading record: record larger than the GPU text buffer")
                bigger = B.TextProc(min(tp.capacity * 4, 1 << 31))
                bigger.input[:fill] = tp.input[:fill]
                tp.close()
                tp = bigger
                continue
            rest = fill - consumed
            if rest:
                tp.input[:rest] = tp.input[consumed:fill].copy()
            fill = rest
    finally:
        if own:
            tp.close()
