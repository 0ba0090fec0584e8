// g++ build of hash_fused3.cuh (see cuda_runtime.h in this directory): C entry points for tests/test_emul_cpu.py.
#include <algorithm>
#include "cuda_runtime.h"
#include "hash_fused3.cuh"

extern "C" {

// record = tile[start, start+len); returns the whitespace flag; out = xxh64, xxh3, murmur h1, h2
uint32_t emul_fused3(const uint8_t *tile, uint32_t start, uint32_t len, int upper, int detect, uint64_t *out)
{
    shb::Fused3Out o;
    uint32_t f;
    if (upper) f = detect ? shb::fused3_record<true, true>(tile, start, len, o) : shb::fused3_record<true, false>(tile, start, len, o);
    else f = detect ? shb::fused3_record<false, true>(tile, start, len, o) : shb::fused3_record<false, false>(tile, start, len, o);
    out[0] = o.xxh64; out[1] = o.xxh3; out[2] = o.m1; out[3] = o.m2;
    return f & 0x80808080u;
}

}
