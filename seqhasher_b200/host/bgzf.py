"""BGZF (blocked gzip: bgzip, htslib) on the host side: finding the members, nothing more.

A BGZF file is a series of complete gzip members of at most 64 KiB of text each; every member header carries an extra
subfield 'BC' whose value is the member's total size - 1, so the members can be delimited without inflating anything.  That
is all the host does for such input: the members' bytes go to the GPU as they are and `shb_textproc_run_bgzf` inflates
them there (csrc/inflate.cuh).  The reference reads the same files through xopen -> pgzip on one goroutine (go.mod:20-31).

`scan_members` mirrors `scan_bgzf` of csrc/host/main.cpp (same acceptance rules); `compress` writes BGZF with zlib for
tests and benchmarks (the image has no bgzip)."""
from __future__ import annotations

import struct
import zlib

import numpy as np

MAX_ISIZE = 65536
EOF_MEMBER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def compress(data: bytes, level: int = 6, block: int = 65280, eof_marker: bool = True, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    """`data` as a BGZF file: one member per `block` bytes of text (bgzip uses 65280), raw DEFLATE at `level` / `strategy`."""
    assert 1 <= block <= MAX_ISIZE
    out = []
    for i in range(0, len(data), block):
        piece = data[i:i + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        cdata = c.compress(piece) + c.flush()
        bsize = 12 + 6 + len(cdata) + 8 - 1
        if bsize > 0xffff:   # incompressible input at level 0 can outgrow the 16-bit size field: halve the block
            out.append(compress(piece, level, max(block // 2, 1), eof_marker=False, strategy=strategy))
            continue
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + cdata +
                   struct.pack("<II", zlib.crc32(piece), len(piece)))
    if eof_marker:
        out.append(EOF_MEMBER)
    return b"".join(out)


def scan_members(buf) -> np.ndarray | None:
    """Member table of a whole BGZF file image: uint64 array (n, 4) of [offset of the raw DEFLATE data, its length,
    ISIZE | CRC32 << 32, offset of the member's first byte], or None when the bytes are not BGZF from the first byte to the
    last (any other gzip stream: the host-side inflate handles it)."""
    b = memoryview(buf).cast("B") if not isinstance(buf, (bytes, bytearray)) else buf
    n = len(b)
    rows = []
    p = 0
    while p < n:
        if n - p < 18 + 8 or b[p] != 0x1f or b[p + 1] != 0x8b or b[p + 2] != 8 or b[p + 3] != 4:
            return None
        xlen = b[p + 10] | (b[p + 11] << 8)
        q, xend, bsize = p + 12, p + 12 + xlen, -1
        if xend > n:
            return None
        while q + 4 <= xend:
            slen = b[q + 2] | (b[q + 3] << 8)
            if b[q] == 66 and b[q + 1] == 67 and slen == 2 and q + 6 <= xend:
                bsize = b[q + 4] | (b[q + 5] << 8)
            q += 4 + slen
        if bsize < 0:
            return None
        total = bsize + 1
        if total < 12 + xlen + 8 or p + total > n:
            return None
        crc, isize = struct.unpack_from("<II", b, p + total - 8)
        if isize > MAX_ISIZE:
            return None
        rows.append((p + 12 + xlen, total - 12 - xlen - 8, isize | (crc << 32), p))
        p += total
    if not rows:
        return None
    return np.array(rows, dtype=np.uint64)


def run_chunks(tp, data, members: np.ndarray, fmt: int, first_start: int, algos, flags: int, headers_only: bool,
               label: bytes | None, own_per_chunk: int, extra0: int = 2):
    """The host protocol of shb_textproc_run_bgzf over a whole file image, one chunk at a time (the compiled host runs several
    in flight): yields (output bytes, n_records, n_empty) per run.  `tp` is a binding.TextProc."""
    n = len(members)
    k = 0
    mv = np.frombuffer(data, dtype=np.uint8)
    while k < n:
        n_own = min(own_per_chunk, n - k)
        extra = min(extra0, n - k - n_own)
        while True:
            last = k + n_own + extra - 1
            lo = int(members[k][3])
            hi = int(members[last][0] + members[last][1]) + 8
            stage = tp.bgzf_input(hi - lo + 64)
            stage[:hi - lo] = mv[lo:hi]
            tab = members[k:last + 1, :3].copy()
            tab[:, 0] -= np.uint64(lo)
            r = tp.run_bgzf(hi - lo, tab, n_own, extra, first_start if k == 0 else -1, last == n - 1, fmt, algos, flags,
                            headers_only, label)
            if r is not None:
                break
            if last == n - 1:
                raise RuntimeError("SHB_NEED_MORE although the chunk reaches the end of the input")
            extra = min(max(extra * 4, 4), n - k - n_own)
        yield r[0].tobytes(), r[2], r[3]
        while tp.pending():
            r = tp.resume(fmt, algos, flags, headers_only, label)
            yield r[0].tobytes(), r[2], r[3]
        k += n_own
