// g++ build of csrc/inflate.cuh (see cuda_runtime.h in this directory): the device DEFLATE decoder and its lane-parallel
// CRC-32 as C entry points for tests/test_emul_cpu.py, which checks them against zlib.  One lane (nl = 1) runs the decoder;
// the CRC is evaluated lane by lane for 32 lanes, as the warp does.
#include <algorithm>
#include <vector>
#include "cuda_runtime.h"
#include "inflate.cuh"

extern "C" int emul_inflate_raw(const uint8_t *src, uint32_t src_len, uint8_t *dst, uint32_t dst_len)
{
    static shb::InflateScratch s;
    // the reader fetches aligned 32-bit words: give it an aligned copy with the stream at an odd offset and slack behind it
    std::vector<uint32_t> buf((src_len + 3 + 16) / 4 + 2, 0xa5a5a5a5u);
    uint8_t *p = reinterpret_cast<uint8_t *>(buf.data()) + 3;
    std::copy(src, src + src_len, p);
    uint32_t produced = 0;
    const int rc = shb::inflate_raw_warp(p, src_len, dst, dst_len, &produced, s, 0, 1);
    if (rc) return rc;
    return produced == dst_len ? 0 : (int)shb::INF_ERR_SIZE;
}

// crc32 of win[first, first + n) the way the warp computes it; win must be 4-byte aligned
extern "C" uint32_t emul_crc32_lanes(const uint8_t *win, uint32_t first, uint32_t n)
{
    static uint32_t tab[4][256];
    shb::crc_build_tables(tab, 0, 1);
    uint32_t c = 0;
    for (uint32_t lane = 0; lane < 32; lane++) c ^= shb::crc_lane_share(win, first, n, tab, lane, 32);
    return c;
}
