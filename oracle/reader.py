"""ORACLE (test infrastructure only) — the record reader the reference gets from `fastx.NewReaderFromIO` /
`reader.Read()` (seqhasher.go:221-239; github.com/shenwei356/bio v0.14.0, go.mod:10) restated in plain Python, plus
`processSequences`' output formatting (seqhasher.go:261-282), so that the expectation of the text-path parity tests
(tests/test_textproc_gpu.py, tests/test_cli_binary.py) does not come from product code.  Nothing under
seqhasher_b200/ imports this file.

bio's source is not under /root/reference (un-vendored module) and there is no Go toolchain here, so this restates
the reader's documented behaviour and what the reference's own tests pin.  Every rule says which:

| # | rule                                                                                   | pinned by                                                        |
|---|----------------------------------------------------------------------------------------|------------------------------------------------------------------|
| 1 | format sniffed from the first non-blank byte: '>' FASTA, '@' FASTQ, else the error     | seqhasher_test.go:780-799 (error text); test.fasta / test3.fastq |
|   | "fastx: invalid FASTA/Q format"; empty input yields no record and no error             | seqhasher_test.go:(empty-input case of TestProcessSequences)     |
| 2 | FASTA: a record starts at '>' at the beginning of a line; Name = the rest of that line | seqhasher_test.go:263-265 (test.fasta), test2_binary.sh:10-60    |
| 3 | FASTA: sequence = the following lines up to the next record start, line breaks         | test/seqs_with_spaces.fasta seq6_with_newline,                   |
|   | dropped, lines concatenated                                                            | test2_binary.sh:131-145                                          |
| 4 | blanks and tabs INSIDE a sequence line survive the reader (seqhasher.go:246 strips)    | seqs_with_spaces.fasta seq4_with_spaces / seq5_with_tabs         |
| 5 | FASTQ: four lines '@Name' / sequence / '+...' / quality; the last line may lack '\\n'  | test/test3.fastq (no final line break), seqhasher_test.go:883-925|
| 6 | output is FASTQ iff the record has a quality string (record.Format; the reference      | seqhasher_test.go:841-880 (FASTA after FASTQ), CHANGELOG.md:7    |
|   | clears stale quality on FASTA, seqhasher.go:235-239)                                   |                                                                  |
| 7 | compressed input (gzip / bzip2 / xz / zstd) is detected by magic bytes                 | seqhasher_test.go:349-400, test2_binary.sh:100-128               |
| 8 | a header-only record (no sequence lines) is a record with an empty sequence            | UNPINNED (problematic_sequences.fasta has one; no reference test |
|   |                                                                                        | reads that file) — seqhasher.go:292-295 gives it empty hashes    |
| 9 | '\\r' before '\\n' is dropped from names and sequence lines (CRLF files)               | UNPINNED (bio trims line ends; no reference fixture has CRLF)    |
| 10| blank lines between FASTA records / inside a record contribute nothing                 | UNPINNED (follows from rules 3 + seqhasher.go:246)               |
| 11| '>' or '@' that is not at the beginning of a line is data (names may contain them)     | UNPINNED (follows from rule 2)                                   |
| 12| FASTQ quality lines may start with '@' or '+': line position decides, not the byte     | UNPINNED (strict 4-line reading; bio's multi-line FASTQ support  |
|   |                                                                                        | is not relied upon)                            |
| 13| a FASTQ record whose '+' line does not start with '+', or an input that ends inside a  | UNPINNED (the error text is rule 1's; which records are written  |
|   | record, is "fastx: invalid FASTA/Q format" after the complete records before it        | before the error follows from the streaming loop seqhasher.go:227)|
| 14| leading blank lines before the first record are skipped; trailing ones are ignored     | UNPINNED                                                         |
"""
from __future__ import annotations

import bz2
import gzip
import io
import lzma
from typing import Iterator, NamedTuple, Optional

ERR_INVALID = "fastx: invalid FASTA/Q format"
_BLANK = b" \t\n\v\f\r"


class ReaderError(Exception):
    pass


class Rec(NamedTuple):
    name: bytes
    seq: bytes                  # as the reader hands it over: line breaks removed, interior blanks kept
    qual: Optional[bytes]       # None for FASTA


def _plain(data: bytes) -> bytes:
    """rule 7: magic-byte sniffing (zstd: tests decompress with the zstd CLI themselves; not needed by the oracle)"""
    if data[:2] == b"\x1f\x8b":
        return gzip.decompress(data)
    if data[:3] == b"BZh":
        return bz2.decompress(data)
    if data[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(data)
    return data


def _lines(text: bytes):
    """(line without its line break and without a trailing '\\r')"""
    for ln in text.split(b"\n"):
        yield ln[:-1] if ln.endswith(b"\r") else ln


def records(data: bytes) -> Iterator[Rec]:
    """Yields the records in order; raises ReaderError(ERR_INVALID) where the reference's reader would fail
    (after having yielded the records before the failure, like the streaming loop of seqhasher.go:227-239)."""
    text = _plain(data)
    body = text.lstrip(_BLANK)                       # rule 14
    if not body:
        return                                       # rule 1: empty input, no records
    if body[:1] == b">":
        body = body.rstrip(b"\r\n")                  # rule 14 (blanks inside the last line belong to that line)
        name, parts = None, []
        for ln in _lines(body):
            if ln[:1] == b">":                       # rule 2 (rule 11: only at the beginning of a line)
                if name is not None:
                    yield Rec(name, b"".join(parts), None)
                name, parts = ln[1:], []
            else:
                parts.append(ln)                     # rules 3, 4, 10
        yield Rec(name, b"".join(parts), None)       # rule 8: possibly with an empty sequence
    elif body[:1] == b"@":
        body = body.rstrip(b"\r\n")
        lines = list(_lines(body))
        i = 0
        while i < len(lines):
            if i + 4 > len(lines):                   # rule 13: the input ends inside a record
                raise ReaderError(ERR_INVALID)
            head, seq, plus, qual = lines[i:i + 4]   # rule 12: position, not content, decides
            if head[:1] != b"@" or plus[:1] != b"+":
                raise ReaderError(ERR_INVALID)       # rule 13
            yield Rec(head[1:], seq, qual)
            i += 4
    else:
        raise ReaderError(ERR_INVALID)               # rule 1


def process_sequences(data: bytes, hash_types, headers_only=False, no_file_name=False, case_sensitive=False,
                      input_file_name="-", name_override="", on_empty=None):
    """processSequences (seqhasher.go:210-286) over `data` with the oracle hashes.  Returns (output bytes, error text or
    None): like the reference, everything written before a reader error stays written."""
    from . import oracle as O
    label, nofile = input_file_name, no_file_name
    if name_override != "":
        label = name_override
    elif label == "-":
        nofile = True                                # seqhasher.go:217-219
    flags = 2 | (0 if case_sensitive else 1)
    out = io.BytesIO()
    first = True
    try:
        for rec in records(data):
            first = False
            seq = O.normalise(rec.seq, flags)        # seqhasher.go:246-251
            hashes = []
            for t in hash_types:                     # seqhasher.go:255-259
                h = O.get_hash_func(t)(seq)
                if h == "" and on_empty:
                    on_empty()                       # the reference logs one line per empty hash (seqhasher.go:293)
                hashes.append(h)
            parts = ([] if nofile else [label.encode()]) + ([";".join(hashes).encode()] if hash_types else []) + [rec.name]
            name = b";".join(parts)                  # seqhasher.go:261-272
            if headers_only:
                out.write(name + b"\n")
            elif rec.qual:                           # rule 6
                out.write(b"@" + name + b"\n" + seq + b"\n+\n" + rec.qual + b"\n")
            else:
                out.write(b">" + name + b"\n" + seq + b"\n")
    except ReaderError as e:
        return out.getvalue(), ("Failed to create reader: " if first else "Error reading record: ") + str(e)
    return out.getvalue(), None
