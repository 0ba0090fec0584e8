// g++ build of hash_nt_planes.cuh (see cuda_runtime.h in this directory): the bit-plane fold of ntHash units as a C entry point
// for tests/test_emul_cpu.py, which checks it against the seed-table definition (nthash.go ntf64 via the oracle's table).
#include <algorithm>
#include "cuda_runtime.h"
#include "hash_nt_planes.cuh"

extern "C" {
// units[0, 16 n): n 16-byte units of ONE lane (same rotation class).  valid[i] = 1 when unit i went into the planes.
// Returns the 64-bit value of the planes (before the lane's residual rotation): XOR over the valid units of
// XOR_j rol(seed[byte j], 15 - j).
uint64_t emul_nt_planes(const uint8_t *units, uint32_t n, uint8_t *valid)
{
    uint32_t aq = 0, and_ = 0, n_fast = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint4 v;
        memcpy(&v, units + 16 * i, 16);
        valid[i] = shb::nt_unit_planes(v, aq, and_, n_fast) ? 1 : 0;
    }
    return shb::nt_planes_to_hash(aq, and_, n_fast);
}
uint64_t emul_nt_seed(uint32_t c) { return shb::nt_seed_const(c & 0xffu); }
}
