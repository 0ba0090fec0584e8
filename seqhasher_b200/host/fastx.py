"""FASTA/FASTQ record reader with the semantics the reference gets from shenwei356/bio v0.14.0
(`fastx.NewReaderFromIO(seq.DNA, bufio.NewReader(input), fastx.DefaultIDRegexp)`, seqhasher.go:221), as
summarised in SURVEY.md App. B: format sniffed from the first non-blank byte ('>' FASTA, '@' FASTQ, else the
error "fastx: invalid FASTA/Q format"); transparent gzip / zstd / xz / bzip2 by magic bytes; FASTA sequence
= the following lines with line breaks removed (interior blanks survive and are stripped later by the hash
path, seqhasher.go:246); FASTQ = 4-line records; Name = the whole header line.

This stays on the host in the new design (north_star: "the Go host keeps the bio reader").
"""
from __future__ import annotations

import bz2
import ctypes
import ctypes.util
import gzip
import io
import lzma
from dataclasses import dataclass


class FastxError(Exception):
    pass


ERR_INVALID = "fastx: invalid FASTA/Q format"


@dataclass
class Record:
    name: bytes
    seq: bytes
    qual: bytes | None = None   # None for FASTA (seqhasher.go:236-239 nils stale quality)


def _zstd_decompress_stream(raw: io.BufferedIOBase) -> bytes:
    """libzstd.so.1 via ctypes (no python zstd module in this image)."""
    path = ctypes.util.find_library("zstd") or "libzstd.so.1"
    z = ctypes.CDLL(path)
    z.ZSTD_createDStream.restype = ctypes.c_void_p
    z.ZSTD_freeDStream.argtypes = [ctypes.c_void_p]
    z.ZSTD_initDStream.argtypes = [ctypes.c_void_p]
    z.ZSTD_initDStream.restype = ctypes.c_size_t
    z.ZSTD_isError.argtypes = [ctypes.c_size_t]

    class Buf(ctypes.Structure):
        _fields_ = [("p", ctypes.c_void_p), ("size", ctypes.c_size_t), ("pos", ctypes.c_size_t)]

    z.ZSTD_decompressStream.argtypes = [ctypes.c_void_p, ctypes.POINTER(Buf), ctypes.POINTER(Buf)]
    z.ZSTD_decompressStream.restype = ctypes.c_size_t
    ds = z.ZSTD_createDStream()
    z.ZSTD_initDStream(ds)
    out = bytearray()
    obuf = ctypes.create_string_buffer(1 << 17)
    try:
        while True:
            chunk = raw.read(1 << 17)
            if not chunk:
                break
            ibuf = ctypes.create_string_buffer(chunk, len(chunk))
            src = Buf(ctypes.cast(ibuf, ctypes.c_void_p), len(chunk), 0)
            while src.pos < src.size:
                dst = Buf(ctypes.cast(obuf, ctypes.c_void_p), len(obuf), 0)
                rc = z.ZSTD_decompressStream(ds, ctypes.byref(dst), ctypes.byref(src))
                if z.ZSTD_isError(rc):
                    raise FastxError("zstd: corrupt input")
                out += obuf.raw[: dst.pos]
    finally:
        z.ZSTD_freeDStream(ds)
    return bytes(out)


def open_maybe_compressed(stream: io.BufferedIOBase) -> io.BufferedIOBase:
    """Sniff the magic bytes (what xopen does underneath bio) and return a decompressing reader."""
    if not hasattr(stream, "peek"):
        stream = io.BufferedReader(stream)  # type: ignore[arg-type]
    head = stream.peek(6)[:6]
    if head[:2] == b"\x1f\x8b":
        return io.BufferedReader(gzip.GzipFile(fileobj=stream))  # multi-member aware, like pgzip
    if head[:3] == b"BZh":
        return io.BufferedReader(bz2.BZ2File(stream))
    if head[:6] == b"\xfd7zXZ\x00":
        return io.BufferedReader(lzma.LZMAFile(stream))
    if head[:4] == b"\x28\xb5\x2f\xfd":
        return io.BytesIO(_zstd_decompress_stream(stream))
    return stream


def _strip_eol(line: bytes) -> bytes:
    if line.endswith(b"\n"):
        line = line[:-1]
    if line.endswith(b"\r"):
        line = line[:-1]
    return line


def read_records(stream):
    """Generator of Record.  Raises FastxError(ERR_INVALID) when the first non-blank byte is not '>' / '@'."""
    f = open_maybe_compressed(stream)
    line = f.readline()
    while line and not line.strip():
        line = f.readline()
    if not line:
        return  # empty input: no records (bio returns io.EOF on the first Read)
    first = line.lstrip()[:1]
    if first == b">":
        name = _strip_eol(line.lstrip())[1:]
        parts: list[bytes] = []
        for line in f:
            if line[:1] == b">":
                yield Record(name, b"".join(parts), None)
                name, parts = _strip_eol(line)[1:], []
            else:
                parts.append(_strip_eol(line))
        yield Record(name, b"".join(parts), None)
    elif first == b"@":
        while line:
            if not line.strip():
                line = f.readline()
                continue
            if line[:1] != b"@":
                raise FastxError(ERR_INVALID)
            name = _strip_eol(line)[1:]
            seq = _strip_eol(f.readline())
            plus = f.readline()
            if plus[:1] != b"+":
                raise FastxError(ERR_INVALID)
            qline = f.readline()
            if not qline:            # the input ends inside the record (no quality line at all)
                raise FastxError(ERR_INVALID)
            qual = _strip_eol(qline)
            yield Record(name, seq, qual)
            line = f.readline()
    else:
        raise FastxError(ERR_INVALID)
