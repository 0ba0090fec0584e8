// hash_md.cuh — SHA-1 and MD5 device hashers (64-byte-block Merkle–Damgard family).
//
// Replaces crypto/sha1.Sum (seqhasher.go:299-301) and crypto/md5.Sum (seqhasher.go:305-307).
// One thread walks one record: the 80 (64) rounds are fully unrolled over a rolling 16-word
// schedule held in registers; data blocks and the padded tail go through the same single inlined
// compress so the instruction footprint stays one block body.  Integer-pipe bound: the work per
// block is fixed (SHA-1: 613 fused int ops).
#pragma once
#include "common.cuh"

namespace shb {

// Pipe balancing.  SHA-1 is bound by the ALU pipe (LOP3/SHF/IADD3/PRMT: 64 lanes/clk/SM on sm_100)
// while the FMA pipe (IMAD: another 64 lanes/clk/SM) idles.  A rotate by a constant is
// hi32(x * 2^k) + lo32(x * 2^k): one IMAD.WIDE plus one IMAD on the FMA pipe instead of one SHF on the
// ALU pipe.  The multipliers come from kernel parameters (values the compiler cannot see), otherwise
// ptxas folds the products back into shifts.  VARIANT bits choose what moves to the FMA pipe:
//   1: rol(b,30)   2: the schedule's rol(.,1)   4: rol(a,5) as two separate addends   8: W[t]+K
//   16 / 32: rol(b,30) / the schedule's rol as IMAD.HI + IMAD (no IMAD.WIDE) — also slower than SHF:
//   23.6 T (variant 8) against 21.0 / 21.3 / 18.2 T (variants 24 / 40 / 56), tools/probe_sha1_variants.py
struct PipeConsts {
    uint32_t one, p1, p5, p30;   // 1, 2, 32, 1<<30
};
__host__ __device__ inline PipeConsts pipe_consts() { return PipeConsts{1u, 2u, 32u, 1u << 30}; }

__device__ __forceinline__ uint32_t fma_rol(uint32_t x, uint32_t pow2, uint32_t one)
{
    uint64_t p = (uint64_t)x * (uint64_t)pow2;
    return (uint32_t)(p >> 32) * one + (uint32_t)p;
}
// the same rotate as IMAD.HI + IMAD (variant bits 16 / 32): no 64-bit register pair, no IMAD.WIDE
__device__ __forceinline__ uint32_t fma_rol_hi(uint32_t x, uint32_t pow2)
{
    return x * pow2 + __umulhi(x, pow2);
}

template <int VARIANT>
__device__ __forceinline__ void sha1_compress_v(uint32_t (&h)[5], uint32_t (&w)[16], const PipeConsts &pc)
{
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
#pragma unroll
    for (int t = 0; t < 80; t++) {
        uint32_t wt;
        if (t < 16) {
            wt = w[t];
        } else {
            uint32_t x = w[(t - 3) & 15] ^ w[(t - 8) & 15] ^ w[(t - 14) & 15] ^ w[t & 15];
            wt = (VARIANT & 32) ? fma_rol_hi(x, pc.p1) : (VARIANT & 2) ? fma_rol(x, pc.p1, pc.one) : rotl32(x, 1);
            w[t & 15] = wt;
        }
        uint32_t f, k;
        if (t < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
        else if (t < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
        else if (t < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
        else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
        const uint32_t kw = (VARIANT & 8) ? wt * pc.one + k : wt + k;
        uint32_t tmp;
        if (VARIANT & 4) {
            uint64_t p = (uint64_t)a * (uint64_t)pc.p5;
            tmp = ((uint32_t)p + (uint32_t)(p >> 32) + f) + (e + kw);
        } else {
            tmp = rotl32(a, 5) + f + e + kw;
        }
        e = d; d = c;
        c = (VARIANT & 16) ? fma_rol_hi(b, pc.p30) : (VARIANT & 1) ? fma_rol(b, pc.p30, pc.one) : rotl32(b, 30);
        b = a; a = tmp;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}

// Production form = variant 8 (measured fastest on B200, tools/probe_sha1_variants.py): only the
// W[t]+K adds are forced onto the FMA pipe, which also makes ptxas emit the remaining two-input adds as
// IMAD.IADD; IMAD.WIDE-based rotates are slower than SHF and stay off.
__device__ __forceinline__ void sha1_compress(uint32_t (&h)[5], uint32_t (&w)[16])
{
    sha1_compress_v<8>(h, w, PipeConsts{c_one, 2u, 32u, 1u << 30});
}

// Word `i` of the padded tail block that starts at `base` (base < len < base + 64): record bytes, then 0x80,
// then zeros.  Branch-free: the load address is clamped to the last data word, the word is masked to its k
// valid bytes and the pad marker dropped in with clamped funnel shifts.  BE = big-endian word (SHA-1).
template <bool BE, class R>
__device__ __forceinline__ uint32_t md_tail_word(const R &r, uint32_t off, uint32_t len, uint32_t last_word_off)
{
    const int k = (int)len - (int)off;                       // valid bytes from this word on (may be <= 0 or >= 4)
    const uint32_t raw = BE ? r.u32be(min(off, last_word_off)) : r.u32(min(off, last_word_off));
    const uint32_t sh = 8u * (uint32_t)max(min(k, 4), 0);     // 0, 8, 16, 24, 32
    uint32_t keep, mark;
    if (BE) {
        keep = ~__funnelshift_rc(0xffffffffu, 0u, sh);        // top k bytes
        mark = __funnelshift_rc(0x80000000u, 0u, sh);         // 0x80 right after them (0 when k >= 4)
    } else {
        keep = ~__funnelshift_lc(0u, 0xffffffffu, sh);        // low k bytes
        mark = __funnelshift_lc(0u, 0x80u, sh);
    }
    return (raw & keep) | (k >= 0 ? mark : 0u);
}

// out: 20 bytes, 4-byte aligned
// (The branch-free md_tail_word form measured 8 % slower here when the reader also carries the whitespace
// detector — the kernel runs at 20 warps/SM and the longer dependent chains of the tail stall it — so SHA-1
// keeps the predicated padded_u32 tail; MD5 below is 7 % faster with md_tail_word.  A/B: tools/ab_quick.py.
// Both keep 32-bit word loads on the direct path too: they are ALU-pipe-bound, and the select stages of the
// 128-bit Stream reader cost 2-18 % there while the multiply-based hashes gain 1.5-2.9x, tools/ab_direct.py.)
template <class R>
__device__ __forceinline__ void sha1_record(const R &r, uint32_t len, uint8_t *out)
{
    uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
    const uint32_t nblocks = (len + 9 + 63) >> 6;   // len < 2^31 guaranteed by the caller
    for (uint32_t blk = 0; blk < nblocks; blk++) {
        const uint32_t base = blk << 6;
        uint32_t w[16];
        if (base + 64 <= len) {
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = r.u32be(base + 4 * i);   // align + byte-swap in one PRMT
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = bswap32(padded_u32(r, base + 4 * i, len, 0x80u));
            if (blk == nblocks - 1) {
                w[14] = len >> 29;
                w[15] = len << 3;
            }
        }
        sha1_compress(h, w);
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(out);
#pragma unroll
    for (int i = 0; i < 5; i++) o[i] = bswap32(h[i]);
}

__device__ __forceinline__ void md5_compress(uint32_t (&h)[4], const uint32_t (&m)[16])
{
    constexpr uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
        0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
        0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
        0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
        0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
        0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
        0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
        0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    constexpr int S[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        uint32_t t = d;
        d = c; c = b;
        // (forcing these adds onto the FMA pipe, as in SHA-1, measured 15 % slower: MD5 is one serial chain and
        //  the cross-pipe hop adds latency to every step; moving only the off-chain m[g] + K[i] there was also
        //  slower, +1 % on C2 and +7-10 % on 100-600 bp reads)
        b = b + rotl32(a + f + K[i] + m[g], S[i >> 4][i & 3]);
        a = t;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}

// out: 16 bytes, 4-byte aligned
template <class R>
__device__ __forceinline__ void md5_record(const R &r, uint32_t len, uint8_t *out)
{
    uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    const uint32_t nblocks = (len + 9 + 63) >> 6;
    const uint32_t last_word_off = (len - 1) & ~3u;
    for (uint32_t blk = 0; blk < nblocks; blk++) {
        const uint32_t base = blk << 6;
        uint32_t m[16];
        if (base + 64 <= len) {
#pragma unroll
            for (int i = 0; i < 16; i++) m[i] = r.u32(base + 4 * i);
        } else if (base >= len) {
#pragma unroll
            for (int i = 0; i < 16; i++) m[i] = 0;
            if (base == len) m[0] = 0x80u;
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) m[i] = md_tail_word<false>(r, base + 4 * i, len, last_word_off);
        }
        if (blk == nblocks - 1) {
            m[14] = len << 3;
            m[15] = len >> 29;
        }
        md5_compress(h, m);
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(out);
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = h[i];
}

}  // namespace shb
