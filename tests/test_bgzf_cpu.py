"""Host side of the BGZF path: the member scanner and the BGZF writer used by tests and benchmarks (no GPU)."""
import gzip
import zlib

import numpy as np

from seqhasher_b200.host import bgzf


def test_compress_is_valid_multi_member_gzip_and_scans_back():
    rng = np.random.default_rng(1)
    for n, level, block in ((0, 6, 65280), (1, 6, 65280), (200_000, 6, 65280), (200_000, 1, 65536), (70_000, 0, 65536), (5000, 9, 100)):
        data = rng.choice(np.frombuffer(b"ACGT\n", dtype=np.uint8), size=n).tobytes()
        gz = bgzf.compress(data, level, block)
        assert gzip.decompress(gz) == data
        m = bgzf.scan_members(gz)
        assert m is not None and int((m[:, 2] & np.uint64(0xffffffff)).sum()) == n
        pos = 0
        for off, clen, packed, start in m.tolist():
            piece = zlib.decompress(gz[off:off + clen], -15)
            assert len(piece) == packed & 0xffffffff and zlib.crc32(piece) == packed >> 32
            assert data[pos:pos + len(piece)] == piece
            pos += len(piece)
        assert gz.endswith(bgzf.EOF_MEMBER)


def test_scan_refuses_anything_that_is_not_bgzf_throughout():
    data = b">a\nACGT\n" * 1000
    assert bgzf.scan_members(gzip.compress(data)) is None                    # plain gzip: no 'BC' field
    good = bgzf.compress(data, 6, 1000)
    assert bgzf.scan_members(good) is not None
    assert bgzf.scan_members(good + gzip.compress(b"tail")) is None          # a foreign member behind BGZF members
    assert bgzf.scan_members(good[:-5]) is None                              # truncated
    assert bgzf.scan_members(good + b"\x00") is None                         # trailing garbage
    assert bgzf.scan_members(b"") is None
