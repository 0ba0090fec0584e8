// AddressSanitizer / UBSan fuzz of the device DEFLATE decoder (csrc/inflate.cuh, host build through cuda_runtime.h of this
// directory): random garbage, valid streams, bit-flipped and truncated streams with wrong announced sizes, in exact-size heap
// buffers.  A decoder that can be steered out of its buffers by a damaged file would corrupt GPU memory; compute-sanitizer is not
// available on the GPU pool, so this is where that is checked.  usage: inflate_fuzz [trials]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <zlib.h>
#include "cuda_runtime.h"
#include "inflate.cuh"

static shb::InflateScratch scr;
static int run(const std::vector<uint8_t> &src, uint32_t dst_len, std::vector<uint8_t> &dst)
{
    // exact-size heap buffers: the reader may touch the aligned words that hold the stream, nothing else
    const size_t words = (src.size() + 3) / 4 + 1;
    uint32_t *buf = new uint32_t[words];
    memset(buf, 0x5a, words * 4);
    if (!src.empty()) memcpy(buf, src.data(), src.size());
    uint8_t *out = new uint8_t[dst_len ? dst_len : 1];
    uint32_t produced = 0;
    int rc = shb::inflate_raw_warp(reinterpret_cast<uint8_t *>(buf), (uint32_t)src.size(), out, dst_len, &produced, scr, 0, 1);
    if (produced > dst_len) { printf("produced > cap!\n"); abort(); }
    dst.assign(out, out + produced);
    delete[] buf;
    delete[] out;
    return rc;
}

int main(int argc, char **argv)
{
    const int trials = argc > 1 ? atoi(argv[1]) : 200000;
    std::mt19937_64 rng(123);
    long ok = 0, bad = 0;
    for (int trial = 0; trial < trials; trial++) {
        std::vector<uint8_t> src, dst;
        uint32_t dst_len = (uint32_t)(rng() % 70000);
        int kind = trial % 4;
        if (kind == 0) {                               // pure garbage
            src.resize(rng() % 300);
            for (auto &c : src) c = (uint8_t)rng();
        } else {                                       // a valid stream, then damaged
            std::vector<uint8_t> text(rng() % 5000);
            for (auto &c : text) c = "ACGT\n>"[rng() % (kind == 1 ? 4 : 6)];
            uLongf cl = compressBound(text.size()) + 16;
            std::vector<uint8_t> comp(cl);
            z_stream z{};
            deflateInit2(&z, (int)(rng() % 10), Z_DEFLATED, -15, 8, (int)(rng() % 4 == 0 ? Z_FIXED : Z_DEFAULT_STRATEGY));
            z.next_in = text.data(); z.avail_in = (uInt)text.size(); z.next_out = comp.data(); z.avail_out = (uInt)cl;
            deflate(&z, Z_FINISH);
            comp.resize(cl - z.avail_out);
            deflateEnd(&z);
            src = comp;
            dst_len = (uint32_t)text.size();
            if (kind == 2 && !src.empty()) { for (int k = 0; k < 1 + (int)(rng() % 3); k++) src[rng() % src.size()] ^= (uint8_t)(1u << (rng() % 8)); }
            if (kind == 3 && !src.empty()) { src.resize(rng() % src.size()); if (rng() % 2) dst_len = (uint32_t)(rng() % (text.size() + 10)); }
            if (kind == 1) {
                int rc = run(src, dst_len, dst);
                if (rc != 0 || dst != text) { printf("valid stream failed rc=%d trial=%d\n", rc, trial); return 1; }
                ok++;
                continue;
            }
        }
        int rc = run(src, dst_len, dst);
        (rc == 0 ? ok : bad)++;
    }
    printf("fuzz done: %ld accepted, %ld refused, no memory error\n", ok, bad);
    return 0;
}
