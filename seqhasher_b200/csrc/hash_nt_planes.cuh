// hash_nt_planes.cuh — ntHash of 16-byte units that hold bases only, without table look-ups (used by nthash_warp_kernel in
// long.cu; a header of its own so that tests/emul can compile it for the host and check it against the seed-table definition).
#pragma once
#include "hash_fast.cuh"

namespace shb {

// ---- bit-plane path for units that hold nothing but A, C, G, T, U (either case) ----
//
// The look-up path above is bound by what shared memory returns (8 bytes per input byte).  For the bases the seed is a
// function of two bits of the ASCII code — b1 and b2: A 00, C 01 (b1), T and U 10 (b2), G 11 — and every such function is
// S = S_A ^ b1 D1 ^ b2 D2 ^ (b1 & b2) D3, so the hash of the record is
//     XOR_j rol(S_A, r_j)  ^  XOR_j b1_j rol(D1, r_j)  ^  XOR_j b2_j rol(D2, r_j)  ^  XOR_j (b1 b2)_j rol(D3, r_j)
// and all a lane has to keep per plane is ONE BIT PER BYTE POSITION OF ITS UNIT, XORed over its units (they lie 512 bytes
// apart, 512 = 0 mod 64, so the same position always carries the same rotation).  Per unit: the two code bits of all 16 bytes
// are packed into one register (four mask-merges; the shifts are IMADs on the other pipe), the b1 & b2 plane is two more
// instructions, and each is XORed into the lane's accumulator — no shared-memory traffic at all.  The planes are expanded
// into the 64-bit hash once per lane and record (nt_planes_to_hash).  A unit with any other byte (N, IUPAC codes, control
// bytes) is not folded here: the lane notes its index (a few per record for reads with scattered N) and folds those units
// with the look-up path after the main loop, so the result is the same whatever the input and a rare N does not make the
// whole warp take the slow path.
constexpr uint64_t NT_D1 = NT_SEED_A ^ NT_SEED_C, NT_D2 = NT_SEED_A ^ NT_SEED_T, NT_D3 = NT_SEED_A ^ NT_SEED_C ^ NT_SEED_T ^ NT_SEED_G;
__host__ __device__ constexpr uint64_t rotl64c(uint64_t x, int r) { return r ? (x << r) | (x >> (64 - r)) : x; }
__host__ __device__ constexpr uint64_t nt_k0()   // the S_A share of one whole unit: XOR_j rol(S_A, 15 - j)
{
    uint64_t k = 0;
    for (int j = 0; j < 16; j++) k ^= rotl64c(NT_SEED_A, 15 - j);
    return k;
}

// true when every byte of the word is one of ACGTU / acgtu.  (b4, b2 b1 b0): A 0,001  C 0,011  G 0,111  T 1,100  U 1,101;
// bits 7, 6, 3 are 0, 1, 0; bit 5 is the case bit and is ignored.  Everything is lined up on bit 4 with left shifts (IMADs).
__device__ __forceinline__ uint32_t nt_word_ok_bits(uint32_t w)
{
    const uint32_t y0 = w << 4, y1 = w << 3, y2 = w << 2;
    const uint32_t acg = y0 & (y1 | ~y2), tu = y2 & ~y1;
    return (~w & acg) | (w & tu);          // bit 4 of every byte: 1 = the low bits fit the base that bit 4 announces
}

// Folds the unit into the lane's planes if it holds bases only and returns true; returns false and changes nothing otherwise.
// aq: bits 8 k + 2 g, 8 k + 2 g + 1 = b1, b2 of unit byte 4 g + k;  and: bit 8 k + 2 g = b1 & b2 of that byte.
__device__ __forceinline__ bool nt_unit_planes(uint4 v, uint32_t &aq, uint32_t &and_, uint32_t &n_fast)
{
    const uint32_t okb = nt_word_ok_bits(v.x) & nt_word_ok_bits(v.y) & nt_word_ok_bits(v.z) & nt_word_ok_bits(v.w);
    const uint32_t any1 = v.x | v.y | v.z | v.w, all1 = v.x & v.y & v.z & v.w;
    const bool ok = ((okb & 0x10101010u) == 0x10101010u) && ((any1 & 0x88888888u) == 0u) && ((all1 & 0x40404040u) == 0x40404040u);
    const uint32_t q = ((v.x >> 1) & 0x03030303u) | ((v.y << 1) & 0x0c0c0c0cu) | ((v.z << 3) & 0x30303030u) | ((v.w << 5) & 0xc0c0c0c0u);
    const uint32_t m = ok ? 0xffffffffu : 0u;
    aq ^= q & m;
    and_ ^= q & (q >> 1) & 0x55555555u & m;
    n_fast += ok ? 1u : 0u;
    return ok;
}

// the 64-bit value of a lane's planes before the lane's residual rotation; n_units = units that went into them
__device__ __forceinline__ uint64_t nt_planes_to_hash(uint32_t aq, uint32_t and_, uint32_t n_units)
{
    constexpr uint64_t K0 = nt_k0();
    uint64_t h = (n_units & 1u) ? K0 : 0ull;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int bit = 8 * (j & 3) + 2 * (j >> 2);   // unit byte j = word j >> 2, byte j & 3
        const uint64_t m1 = 0ull - (uint64_t)((aq >> bit) & 1u), m2 = 0ull - (uint64_t)((aq >> (bit + 1)) & 1u),
                       m3 = 0ull - (uint64_t)((and_ >> bit) & 1u);
        const uint64_t c1 = rotl64c(NT_D1, 15 - j), c2 = rotl64c(NT_D2, 15 - j), c3 = rotl64c(NT_D3, 15 - j);
        h ^= (c1 & m1) ^ (c2 & m2) ^ (c3 & m3);
    }
    return h;
}
}  // namespace shb
