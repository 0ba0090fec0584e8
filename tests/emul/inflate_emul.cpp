// g++ build of csrc/inflate.cuh (see cuda_runtime.h in this directory): the device DEFLATE decoder as a C entry point for
// tests/test_emul_cpu.py, which checks it against zlib.
#include <algorithm>
#include "cuda_runtime.h"
#include "inflate.cuh"

extern "C" int emul_inflate_raw(const uint8_t *src, uint32_t src_len, uint8_t *dst, uint32_t dst_len)
{
    static shb::InflateScratch s;
    return shb::inflate_raw(src, src_len, dst, dst_len, s);
}
