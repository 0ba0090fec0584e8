// hash_keccak.cuh — SHA3-512 and KangarooTwelve (KT128) device hashers (sponge family).
//
// Replaces sha3.Sum512 (seqhasher.go:302-304; FIPS 202, rate 72, suffix 0x06, 24 rounds) and
// k12.NewDraft10(nil)+Write+Read(128) (seqhasher.go:331-336; RFC 9861 KT128, empty customization,
// TurboSHAKE128 = Keccak-p[1600,12] rate 168, 128 output bytes).
// One thread owns one sponge: the 25 lanes stay in registers (every index is static after
// unrolling); the round loop is kept rolled so the body fits the instruction cache.
#pragma once
#include "common.cuh"

namespace shb {

__constant__ uint64_t c_keccak_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

template <int R>
__device__ __forceinline__ uint64_t rol64c(uint64_t x)
{
    if (R == 0) return x;
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (R == 32) return mk64(hi, lo);
    if (R < 32) return mk64(__funnelshift_l(hi, lo, R), __funnelshift_l(lo, hi, R));
    return mk64(__funnelshift_l(lo, hi, R - 32), __funnelshift_l(hi, lo, R - 32));
}

// a ^ b ^ c on 64 bits as two three-input LOP3 (ptxas does not merge the XORs of a value that has several uses)
__device__ __forceinline__ uint64_t xor3_64(uint64_t a, uint64_t b, uint64_t c)
{
#ifdef __CUDA_ARCH__
    uint32_t lo, hi;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(lo) : "r"((uint32_t)a), "r"((uint32_t)b), "r"((uint32_t)c));
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(hi) : "r"((uint32_t)(a >> 32)), "r"((uint32_t)(b >> 32)), "r"((uint32_t)(c >> 32)));
    return mk64(lo, hi);
#else
    return a ^ b ^ c;
#endif
}

// Keccak-p[1600, NR]: the last NR rounds of Keccak-f.  Lane A[x][y] is a[x + 5*y].
template <int NR>
__device__ __forceinline__ void keccak_p(uint64_t (&a)[25])
{
#pragma unroll 1
    for (int round = 24 - NR; round < 24; round++) {
        uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20];
        uint64_t c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21];
        uint64_t c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22];
        uint64_t c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23];
        uint64_t c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
        // theta's D[x] = C[x-1] ^ rol(C[x+1], 1) is never formed: a ^ C[x-1] ^ rol(C[x+1], 1) is ONE three-input LOP3 per
        // half, the same count as a ^ D, so the ten LOP3 of the D's are saved (5 % of the round's ALU-pipe instructions)
        const uint64_t r0 = rol64c<1>(c0), r1 = rol64c<1>(c1), r2 = rol64c<1>(c2), r3 = rol64c<1>(c3), r4 = rol64c<1>(c4);
#define KD0(x) xor3_64(x, c4, r1)
#define KD1(x) xor3_64(x, c0, r2)
#define KD2(x) xor3_64(x, c1, r3)
#define KD3(x) xor3_64(x, c2, r4)
#define KD4(x) xor3_64(x, c3, r0)
        // theta + rho + pi: b[y][(2x+3y)%5] = rol(a[x][y] ^ d[x], r[x][y])
        uint64_t b0 = KD0(a[0]);
        uint64_t b10 = rol64c<1>(KD1(a[1]));
        uint64_t b20 = rol64c<62>(KD2(a[2]));
        uint64_t b5 = rol64c<28>(KD3(a[3]));
        uint64_t b15 = rol64c<27>(KD4(a[4]));
        uint64_t b16 = rol64c<36>(KD0(a[5]));
        uint64_t b1 = rol64c<44>(KD1(a[6]));
        uint64_t b11 = rol64c<6>(KD2(a[7]));
        uint64_t b21 = rol64c<55>(KD3(a[8]));
        uint64_t b6 = rol64c<20>(KD4(a[9]));
        uint64_t b7 = rol64c<3>(KD0(a[10]));
        uint64_t b17 = rol64c<10>(KD1(a[11]));
        uint64_t b2 = rol64c<43>(KD2(a[12]));
        uint64_t b12 = rol64c<25>(KD3(a[13]));
        uint64_t b22 = rol64c<39>(KD4(a[14]));
        uint64_t b23 = rol64c<41>(KD0(a[15]));
        uint64_t b8 = rol64c<45>(KD1(a[16]));
        uint64_t b18 = rol64c<15>(KD2(a[17]));
        uint64_t b3 = rol64c<21>(KD3(a[18]));
        uint64_t b13 = rol64c<8>(KD4(a[19]));
        uint64_t b14 = rol64c<18>(KD0(a[20]));
        uint64_t b24 = rol64c<2>(KD1(a[21]));
        uint64_t b9 = rol64c<61>(KD2(a[22]));
        uint64_t b19 = rol64c<56>(KD3(a[23]));
        uint64_t b4 = rol64c<14>(KD4(a[24]));
#undef KD0
#undef KD1
#undef KD2
#undef KD3
#undef KD4
        // chi
        a[0] = b0 ^ (~b1 & b2);   a[1] = b1 ^ (~b2 & b3);   a[2] = b2 ^ (~b3 & b4);
        a[3] = b3 ^ (~b4 & b0);   a[4] = b4 ^ (~b0 & b1);
        a[5] = b5 ^ (~b6 & b7);   a[6] = b6 ^ (~b7 & b8);   a[7] = b7 ^ (~b8 & b9);
        a[8] = b8 ^ (~b9 & b5);   a[9] = b9 ^ (~b5 & b6);
        a[10] = b10 ^ (~b11 & b12); a[11] = b11 ^ (~b12 & b13); a[12] = b12 ^ (~b13 & b14);
        a[13] = b13 ^ (~b14 & b10); a[14] = b14 ^ (~b10 & b11);
        a[15] = b15 ^ (~b16 & b17); a[16] = b16 ^ (~b17 & b18); a[17] = b17 ^ (~b18 & b19);
        a[18] = b18 ^ (~b19 & b15); a[19] = b19 ^ (~b15 & b16);
        a[20] = b20 ^ (~b21 & b22); a[21] = b21 ^ (~b22 & b23); a[22] = b22 ^ (~b23 & b24);
        a[23] = b23 ^ (~b24 & b20); a[24] = b24 ^ (~b20 & b21);
        a[0] ^= c_keccak_rc[round];
    }
}

// 32-bit word at offset `off` (multiple of 4, relative to `start`) of the virtual padded block
// stream: record bytes [start, start+datalen) then zeros, with `pad` OR-ed in at byte `padpos`.
template <class R>
__device__ __forceinline__ uint32_t sponge_u32(const R &r, uint32_t start, uint32_t off, uint32_t datalen,
                                               uint32_t padpos, uint32_t pad)
{
    uint32_t v;
    if (off + 4 <= datalen) {
        return r.u32(start + off);
    } else if (off >= datalen) {
        v = 0;
    } else {
        uint32_t k = datalen - off;   // 1..3
        v = r.u32(start + off) & ((1u << (8 * k)) - 1u);
    }
    uint32_t rel = padpos - off;      // unsigned wrap: only 0..3 matter
    if (rel < 4) v |= pad << (8 * rel);
    return v;
}

// Absorb [start, start+datalen) + pad byte at padpos (>= datalen) + final 0x80 into `a`.
template <int RATE_LANES, int NR, class R>
__device__ __forceinline__ void sponge_absorb_padded(uint64_t (&a)[25], const R &r, uint32_t start,
                                                     uint32_t datalen, uint32_t padpos, uint32_t pad)
{
    constexpr uint32_t RATE = RATE_LANES * 8;
    const uint32_t nblocks = padpos / RATE + 1;
    for (uint32_t blk = 0; blk < nblocks; blk++) {
        const uint32_t base = blk * RATE;
        if (base + RATE <= datalen) {
#pragma unroll
            for (int i = 0; i < RATE_LANES; i++) a[i] ^= r.u64(start + base + 8 * i);   // start % 8 == 0
        } else {
#pragma unroll
            for (int i = 0; i < RATE_LANES; i++) {
                uint32_t lo = sponge_u32(r, start, base + 8 * i, datalen, padpos, pad);
                uint32_t hi = sponge_u32(r, start, base + 8 * i + 4, datalen, padpos, pad);
                a[i] ^= mk64(lo, hi);
            }
            if (blk == nblocks - 1) a[RATE_LANES - 1] ^= 0x8000000000000000ULL;
        }
        keccak_p<NR>(a);
    }
}

// out: 64 bytes, 8-byte aligned
template <class R>
__device__ __forceinline__ void sha3_512_record(const R &r, uint32_t len, uint8_t *out)
{
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    sponge_absorb_padded<9, 24>(a, r, 0, len, len, 0x06u);
    uint64_t *o = reinterpret_cast<uint64_t *>(out);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = a[i];
}

// out: 128 bytes, 8-byte aligned.  KT128(M, C="") with S = M || 0x00 (RFC 9861 §3).
template <class R>
__device__ __noinline__ void kt128_long(const R &r, uint32_t len, uint8_t *out);

template <class R>
__device__ __forceinline__ void kt128_record(const R &r, uint32_t len, uint8_t *out)
{
    if (len >= 8192) {
        kt128_long(r, len, out);
        return;
    }
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    // |S| = len + 1 <= 8192: TurboSHAKE128(S, 0x07); the 0x00 of S is an implicit zero byte
    sponge_absorb_padded<21, 12>(a, r, 0, len, len + 1, 0x07u);
    uint64_t *o = reinterpret_cast<uint64_t *>(out);
#pragma unroll
    for (int i = 0; i < 16; i++) o[i] = a[i];
}

// Tree mode: S split in 8192-byte chunks; CV_i = TurboSHAKE128(S_i, 0x0B, 32) for i >= 1;
// final node = S_0 || 03 00.. (8 B) || CV_1.. || length_encode(n-1) || FF FF, suffix 0x06.
// The final node's pending rate block is kept as 21 lanes in local memory because the lane
// position of each CV depends on the chunk index.

// chaining value of leaf chunk c (>= 1) of a record of `len` bytes
template <class R>
__device__ __forceinline__ void kt128_leaf_cv(const R &r, uint32_t len, uint32_t c, uint64_t (&cv)[4])
{
    const uint32_t slen = len + 1;
    const uint32_t start = c << 13;
    const uint32_t clen = min(8192u, slen - start);        // chunk length within S
    const uint32_t dlen = min(clen, len - start);          // message bytes in it
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    sponge_absorb_padded<21, 12>(a, r, start, dlen, clen, 0x0Bu);
#pragma unroll
    for (int j = 0; j < 4; j++) cv[j] = a[j];
}

// The final node in two steps.  Step 1, the 48 full rate blocks of S_0 (8064 of its 8192 bytes), depends on the
// record alone — it is 8 KiB of serial sponge per record and can run beside the leaves; step 2 takes that state, the
// last 128 bytes of S_0, the leaf CVs and the trailer.
template <class R>
__device__ __forceinline__ void kt128_s0_state(const R &r, uint64_t (&f)[25])
{
#pragma unroll
    for (int i = 0; i < 25; i++) f[i] = 0;
    for (uint32_t blk = 0; blk < 48; blk++) {
#pragma unroll
        for (int i = 0; i < 21; i++) f[i] ^= r.u64(blk * 168 + 8 * i);
        keccak_p<12>(f);
    }
}

// step 2; leaf CVs come from `leaf(c, cv)` (computed inline, or read back from the leaf kernel's output)
template <class R, class LeafFn>
__device__ __forceinline__ void kt128_final_from_state(const R &r, uint32_t len, uint8_t *out, uint64_t (&f)[25], LeafFn leaf)
{
    uint64_t pend[21];
#pragma unroll
    for (int i = 0; i < 16; i++) pend[i] = r.u64(48 * 168 + 8 * i);
    pend[16] = 0x03ULL;
    uint32_t pos = 17;
    const uint32_t slen = len + 1;
    const uint32_t n = (slen + 8191) >> 13;
    for (uint32_t c = 1; c < n; c++) {
        uint64_t cv[4];
        leaf(c, cv);
#pragma unroll 1
        for (int j = 0; j < 4; j++) {
            if (pos == 21) {
#pragma unroll
                for (int i = 0; i < 21; i++) f[i] ^= pend[i];
                keccak_p<12>(f);
                pos = 0;
            }
            pend[pos++] = cv[j];
        }
    }
    // trailer: length_encode(n-1) || FF FF, then the 0x06 suffix; n-1 < 2^24 here
    const uint32_t m = n - 1;
    const uint32_t k = m < 0x100u ? 1u : (m < 0x10000u ? 2u : 3u);
    uint64_t t = 0;
    for (uint32_t i = 0; i < k; i++) t |= (uint64_t)((m >> (8 * (k - 1 - i))) & 0xffu) << (8 * i);
    t |= (uint64_t)k << (8 * k);
    t |= 0xFFFFULL << (8 * (k + 1));
    t |= 0x06ULL << (8 * (k + 3));
    if (pos == 21) {
#pragma unroll
        for (int i = 0; i < 21; i++) f[i] ^= pend[i];
        keccak_p<12>(f);
        pos = 0;
    }
    pend[pos++] = t;
    while (pos < 21) pend[pos++] = 0;
    pend[20] ^= 0x8000000000000000ULL;
#pragma unroll
    for (int i = 0; i < 21; i++) f[i] ^= pend[i];
    keccak_p<12>(f);
    uint64_t *o = reinterpret_cast<uint64_t *>(out);
#pragma unroll
    for (int i = 0; i < 16; i++) o[i] = f[i];
}

template <class R, class LeafFn>
__device__ __forceinline__ void kt128_final_node(const R &r, uint32_t len, uint8_t *out, LeafFn leaf)
{
    uint64_t f[25];
    kt128_s0_state(r, f);
    kt128_final_from_state(r, len, out, f, leaf);
}

template <class R>
__device__ __noinline__ void kt128_long(const R &r, uint32_t len, uint8_t *out)
{
    kt128_final_node(r, len, out, [&](uint32_t c, uint64_t (&cv)[4]) { kt128_leaf_cv(r, len, c, cv); });
}

}  // namespace shb
