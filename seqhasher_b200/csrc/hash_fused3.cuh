// hash_fused3.cuh — XXH64 + XXH3-64 + MurmurHash3 x64_128 of one short record in ONE pass over its bytes
// (BASELINE config C3: `--hash xxhash,xxh3,murmur3` on 150-bp reads; seqhasher.go:255-259 calls the three
// hashers one after the other, each re-reading the record).
//
// The three algorithms of hash_fast.cuh restated as one loop over 32-byte stripes: a 16-byte group is loaded
// from the shared-memory tile once (4 aligned 32-bit loads + 4 funnel shifts: records sit at arbitrary byte
// offsets), upper-cased once (seqhasher.go:249-251) and fed to all three states while it is in registers.  The
// loop is unrolled over the 7 stripes a record of <= 240 bytes can hold, so every XXH3 secret word is an
// immediate.  Issue slots, not HBM, bound this kernel (three multiply-based hashes cost ~8 integer instructions
// per byte), which is why the bytes are touched once: the separate hashers spend a third of their instructions
// on loading, aligning and upper-casing the same words three times.
//
// Handles 17 <= len <= 240 (XXH3's "17-128" and "129-240" classes); other lengths use the separate hashers.
#pragma once
#include "hash_fast.cuh"

namespace shb {

// the XXH3 secret again, as compile-time values (c_xxh3_secret is the run-time copy for dynamic offsets)
struct XSecretBytes {
    uint8_t b[192];
};
constexpr XSecretBytes kXSecret = {{
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e}};

constexpr uint64_t ksec_at(uint32_t off)
{
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | kXSecret.b[off + i];
    return v;
}
template <uint32_t OFF>
struct KSec {
    static constexpr uint64_t v = ksec_at(OFF);
};

// The fused pass is bound by the FMA pipe (the 64-bit multiplies: IMAD.WIDE + 2 IMAD each), so its input preparation keeps
// its adds on the ALU pipe — the opposite choice from the ALU-bound crypto kernels (common.cuh upper4).
#ifndef F3_UPPER_MODE
#define F3_UPPER_MODE 1   // 0: ptxas decides, 1: FMA pipe, 2: ALU pipe (measured: 1 < 0 < 2, profiles/r02_c3_variants.txt)
#endif
__device__ __forceinline__ uint32_t f3_upper(uint32_t x)
{
    return F3_UPPER_MODE == 2 ? upper4_alu(x) : F3_UPPER_MODE == 1 ? upper4<true>(x) : upper4<false>(x);
}
#ifndef F3_UPPER_SKIP
#define F3_UPPER_SKIP 1   // skip the upper-casing of 16-byte groups that hold no byte with bit 5 set
#endif
#ifndef F3_MUL5_SHIFT
#define F3_MUL5_SHIFT 1
#endif
// h * 5 + c of the murmur3 block.  As a multiply it is an IMAD.WIDE + IMAD (3.6 FMA-pipe slots); as h + (h << 2) + c it is
// two shifts and a three-input 64-bit add on the ALU pipe.
__device__ __forceinline__ uint64_t mul5_add(uint64_t h, uint64_t c)
{
#if F3_MUL5_SHIFT
    uint64_t t = h << 2;
#ifdef __CUDACC__
    asm("" : "+l"(t));   // keeps the compiler from folding h + (h << 2) back into h * 5
#endif
    return h + t + c;
#else
    return h * 5 + c;
#endif
}

struct Fused3State {
    // murmur3
    uint64_t h1, h2;
    // xxh64
    uint64_t v1, v2, v3, v4;
    // xxh3: acc = len * P1 + the mixes of the first phase; acc_b = the mixes added after the mid avalanche (129-240)
    uint64_t acc, acc_b;
};

#ifndef F3_MAD64
#define F3_MAD64 1   // 64-bit multiplies by constants as mad64c (one IMAD.WIDE + two chained IMADs)
#endif
__device__ __forceinline__ uint64_t f3_mul(uint64_t a, uint64_t c) { return F3_MAD64 ? mul64c(a, c) : a * c; }
__device__ __forceinline__ uint64_t f3_mad(uint64_t a, uint64_t c, uint64_t acc) { return F3_MAD64 ? mad64c(a, c, acc) : a * c + acc; }

__device__ __forceinline__ void murmur3_block(uint64_t &h1, uint64_t &h2, uint64_t k1, uint64_t k2)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    k1 = f3_mul(k1, c1); k1 = rotl64_shf(k1, 31); k1 = f3_mul(k1, c2); h1 ^= k1;
    h1 = rotl64_shf(h1, 27); h1 += h2; h1 = mul5_add(h1, 0x52dce729ULL);
    k2 = f3_mul(k2, c2); k2 = rotl64_shf(k2, 33); k2 = f3_mul(k2, c1); h2 ^= k2;
    h2 = rotl64_shf(h2, 31); h2 += h1; h2 = mul5_add(h2, 0x38495ab5ULL);
}
__device__ __forceinline__ uint64_t xxh64_round_s(uint64_t acc, uint64_t x) { return f3_mul(rotl64_shf(f3_mad(x, XP64_2, acc), 31), XP64_1); }
// h ^= round(0, k); h = rotl(h, 27) * P1 + P4: the 8-byte step of the XXH64 tail and of its accumulator merge
__device__ __forceinline__ uint64_t xxh64_step8(uint64_t h, uint64_t k) { return f3_mad(rotl64_shf(h ^ xxh64_round_s(0, k), 27), XP64_1, XP64_4); }

__device__ __forceinline__ uint64_t xxh3_mix16_k(uint64_t lo, uint64_t hi, uint64_t s0, uint64_t s1)
{
    return mulfold64(lo ^ s0, hi ^ s1);
}

// The reader of the fused pass: sequential 16-byte groups of one record out of the tile.
template <bool UPPER, bool DETECT>
struct GroupReader {
    const uint32_t *w;   // next aligned word to load
    uint32_t prev, sh;
    uint32_t wsf;        // DETECT: bit 7 of a byte set <=> a byte < 0x21 was read (whitespace candidate)
    __device__ __forceinline__ GroupReader(const uint8_t *tile, uint32_t start)
    {
        w = reinterpret_cast<const uint32_t *>(tile + (start & ~3u));
        sh = (start & 3u) * 8u;
        smem_check(w, 4);
        prev = *w++;
        wsf = 0;
    }
    // the next 16 bytes as four little-endian words, not upper-cased yet; `valid` = how many of them belong to the record
    // (only the whitespace test looks at it: bytes past the record end must not raise the flag)
    __device__ __forceinline__ void next_raw(uint32_t (&d)[4], uint32_t valid = 16)
    {
        smem_check(w, 16);
        const uint32_t n0 = w[0], n1 = w[1], n2 = w[2], n3 = w[3];
        w += 4;
        d[0] = __funnelshift_r(prev, n0, sh);
        d[1] = __funnelshift_r(n0, n1, sh);
        d[2] = __funnelshift_r(n1, n2, sh);
        d[3] = __funnelshift_r(n2, n3, sh);
        prev = n3;
        if (DETECT) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t f = fma_add(d[i], 0xdedededfu) & ~d[i];   // bit 7 of a byte <=> byte < 0x21
                if (valid < 16) {
                    const int k = (int)valid - 4 * i;               // valid bytes of this word
                    f = k >= 4 ? f : k <= 0 ? 0u : f & (0xffffffffu >> (32 - 8 * k));
                }
                wsf |= f;
            }
        }
    }
    // Every lower-case letter has bit 5 set, upper-case letters (and most of what sequence files hold) do not: words with no
    // such byte need no upper-casing — 6 instructions saved per word against one test per 16 or 32 bytes.  Sequencer reads
    // are upper case throughout; soft-masked sequence comes in runs, so warps mostly agree on the branch.
    static __device__ __forceinline__ bool may_hold_lower(const uint32_t (&d)[4]) { return ((d[0] | d[1] | d[2] | d[3]) & 0x20202020u) != 0; }
    static __device__ __forceinline__ void upper_group(uint32_t (&d)[4])
    {
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = f3_upper(d[i]);
    }
    // one group, upper-cased, as two little-endian 64-bit words
    __device__ __forceinline__ void next(uint64_t &k1, uint64_t &k2, uint32_t valid = 16)
    {
        uint32_t d[4];
        next_raw(d, valid);
        if (UPPER && (!F3_UPPER_SKIP || may_hold_lower(d))) upper_group(d);
        k1 = mk64(d[0], d[1]);
        k2 = mk64(d[2], d[3]);
    }
    // two groups (one 32-byte stripe) with one lower-case test for both
    __device__ __forceinline__ void next2(uint64_t &a1, uint64_t &a2, uint64_t &b1, uint64_t &b2)
    {
        uint32_t d[4], e[4];
        next_raw(d);
        next_raw(e);
        if (UPPER && (!F3_UPPER_SKIP || ((d[0] | d[1] | d[2] | d[3] | e[0] | e[1] | e[2] | e[3]) & 0x20202020u) != 0)) {
            upper_group(d);
            upper_group(e);
        }
        a1 = mk64(d[0], d[1]); a2 = mk64(d[2], d[3]);
        b1 = mk64(e[0], e[1]); b2 = mk64(e[2], e[3]);
    }
};

// 16 bytes at any record offset (the "from the end" reads of XXH3), upper-cased; never raises the whitespace flag
// (every byte it touches has been seen by the sequential reader)
template <bool UPPER>
__device__ __forceinline__ void read16_any(const uint8_t *tile, uint32_t pos, uint64_t &k1, uint64_t &k2)
{
    const uint32_t *q = reinterpret_cast<const uint32_t *>(tile + (pos & ~3u));
    const uint32_t s = (pos & 3u) * 8u;
    smem_check(q, 20);
    const uint32_t a = q[0], b = q[1], c = q[2], d = q[3], e = q[4];
    uint32_t x0 = __funnelshift_r(a, b, s), x1 = __funnelshift_r(b, c, s), x2 = __funnelshift_r(c, d, s), x3 = __funnelshift_r(d, e, s);
    if (UPPER && (!F3_UPPER_SKIP || ((x0 | x1 | x2 | x3) & 0x20202020u) != 0)) { x0 = f3_upper(x0); x1 = f3_upper(x1); x2 = f3_upper(x2); x3 = f3_upper(x3); }
    k1 = mk64(x0, x1);
    k2 = mk64(x2, x3);
}

// one 32-byte stripe S (compile-time): groups 2S and 2S+1
template <int S, bool UPPER, bool DETECT>
__device__ __forceinline__ void fused3_stripe(GroupReader<UPPER, DETECT> &in, Fused3State &st, bool mid, uint32_t pairs)
{
    uint64_t a1, a2, b1, b2;
    in.next2(a1, a2, b1, b2);
    murmur3_block(st.h1, st.h2, a1, a2);
    st.v1 = xxh64_round_s(st.v1, a1);
    st.v2 = xxh64_round_s(st.v2, a2);
    st.v3 = xxh64_round_s(st.v3, b1);
    st.v4 = xxh64_round_s(st.v4, b2);
    murmur3_block(st.h1, st.h2, b1, b2);
    // XXH3.  129-240: groups 0..7 mix with the secret at 16 g, groups 8.. (added after the mid avalanche) at 16 (g - 8) + 3.
    // 17-128: group g < pairs = ceil(len / 32) (at most 4) mixes with the secret at 32 g.
    constexpr int G0 = 2 * S, G1 = 2 * S + 1;
    if (S >= 4) {   // only records > 128 bytes get this far
        st.acc_b += xxh3_mix16_k(a1, a2, KSec<16 * (G0 - 8 > 0 ? G0 - 8 : 0) + 3>::v, KSec<16 * (G0 - 8 > 0 ? G0 - 8 : 0) + 11>::v);
        st.acc_b += xxh3_mix16_k(b1, b2, KSec<16 * (G1 - 8 > 0 ? G1 - 8 : 0) + 3>::v, KSec<16 * (G1 - 8 > 0 ? G1 - 8 : 0) + 11>::v);
    } else if (mid) {
        st.acc += xxh3_mix16_k(a1, a2, KSec<16 * (G0 & 7)>::v, KSec<16 * (G0 & 7) + 8>::v);
        st.acc += xxh3_mix16_k(b1, b2, KSec<16 * (G1 & 7)>::v, KSec<16 * (G1 & 7) + 8>::v);
    } else if (S < 2) {
        if ((uint32_t)G0 < pairs) st.acc += xxh3_mix16_k(a1, a2, KSec<32 * (G0 & 3)>::v, KSec<32 * (G0 & 3) + 8>::v);
        if ((uint32_t)G1 < pairs) st.acc += xxh3_mix16_k(b1, b2, KSec<32 * (G1 & 3)>::v, KSec<32 * (G1 & 3) + 8>::v);
    }
}

// stripes S, S + 1, ... while they are complete: nested (not consecutive) conditionals, so the state joins the code after the
// stripes once instead of after every stripe (ptxas otherwise shuffles the 19 state registers at each join)
template <int S, bool UPPER, bool DETECT>
__device__ __forceinline__ void fused3_stripes(GroupReader<UPPER, DETECT> &in, Fused3State &st, bool mid, uint32_t pairs, uint32_t ns)
{
    if constexpr (S < 7) {
        if (ns > S) {
            fused3_stripe<S, UPPER, DETECT>(in, st, mid, pairs);
            fused3_stripes<S + 1, UPPER, DETECT>(in, st, mid, pairs, ns);
        }
    }
}

struct Fused3Out {
    uint64_t xxh64, xxh3, m1, m2;
};

// 17 <= len <= 240.  Returns the whitespace-candidate flag word (DETECT).
template <bool UPPER, bool DETECT>
__device__ __forceinline__ uint32_t fused3_record(const uint8_t *tile, uint32_t start, uint32_t len, Fused3Out &o)
{
    GroupReader<UPPER, DETECT> in(tile, start);
    Fused3State st;
    st.h1 = 0; st.h2 = 0;
    st.v1 = XP64_1 + XP64_2; st.v2 = XP64_2; st.v3 = 0; st.v4 = 0ULL - XP64_1;
    st.acc = (uint64_t)len * XP64_1;
    st.acc_b = 0;
    const bool mid = len > 128;
    const uint32_t ns = len >> 5;            // complete 32-byte stripes
    const uint32_t pairs = (len + 31) >> 5;  // 17-128: groups mixed from the front (and as many from the end)

    fused3_stripes<0, UPPER, DETECT>(in, st, mid, pairs, ns);

    // ---- the unpaired 16-byte group (when len >> 4 is odd) and the partial group ----
    const uint32_t ng = len >> 4, rem = len & 15u;
    uint64_t hx;   // XXH64 after the stripes
    if (ns) {
        hx = rotl64_shf(st.v1, 1) + rotl64_shf(st.v2, 7) + rotl64_shf(st.v3, 12) + rotl64_shf(st.v4, 18);
        hx = f3_mad(hx ^ xxh64_round_s(0, st.v1), XP64_1, XP64_4);
        hx = f3_mad(hx ^ xxh64_round_s(0, st.v2), XP64_1, XP64_4);
        hx = f3_mad(hx ^ xxh64_round_s(0, st.v3), XP64_1, XP64_4);
        hx = f3_mad(hx ^ xxh64_round_s(0, st.v4), XP64_1, XP64_4);
    } else {
        hx = XP64_5;
    }
    hx += (uint64_t)len;
    if (ng & 1u) {
        uint64_t k1, k2;
        in.next(k1, k2);
        const uint32_t g = ng - 1;           // = 2 * ns
        murmur3_block(st.h1, st.h2, k1, k2);
        hx = xxh64_step8(hx, k1);
        hx = xxh64_step8(hx, k2);
        if (mid) {
            if (g < 8) st.acc += xxh3_mix16_k(k1, k2, xsec(16 * g), xsec(16 * g + 8));     // g == 8 is the first possible here
            else st.acc_b += xxh3_mix16_k(k1, k2, xsec(16 * (g - 8) + 3), xsec(16 * (g - 8) + 11));
        } else if (g < pairs) {
            st.acc += xxh3_mix16_k(k1, k2, xsec(32 * g), xsec(32 * g + 8));
        }
    }
    if (rem) {
        uint64_t k1, k2;
        in.next(k1, k2, rem);
        // murmur3 tail: zero-extended
        {
            const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
            uint64_t t1 = k1, t2 = k2;
            if (rem < 8) t1 &= (1ULL << (8 * rem)) - 1ULL;
            if (rem > 8) {
                t2 &= (1ULL << (8 * (rem - 8))) - 1ULL;   // rem - 8 <= 7
                t2 = f3_mul(t2, c2); t2 = rotl64_shf(t2, 33); t2 = f3_mul(t2, c1); st.h2 ^= t2;
            }
            t1 = f3_mul(t1, c1); t1 = rotl64_shf(t1, 31); t1 = f3_mul(t1, c2); st.h1 ^= t1;
        }
        // XXH64 tail: 8-byte step, 4-byte step, single bytes
        uint64_t restw = k1;                 // the 8 bytes that hold what follows the 8-byte step
        if (rem >= 8) {
            hx = xxh64_step8(hx, k1);
            restw = k2;
        }
        uint32_t tw = (uint32_t)restw;
        if (rem & 4u) {
            hx ^= (uint64_t)tw * XP64_1;
            hx = rotl64_shf(hx, 23) * XP64_2 + XP64_3;
            tw = (uint32_t)(restw >> 32);
        }
        for (uint32_t k = 0; k < (rem & 3u); k++) {
            hx ^= (uint64_t)(tw & 0xffu) * XP64_5;
            hx = rotl64_shf(hx, 11) * XP64_1;
            tw >>= 8;
        }
    }
    o.xxh64 = xxh64_avalanche(hx);

    // ---- murmur3 finish ----
    {
        uint64_t h1 = st.h1 ^ (uint64_t)len, h2 = st.h2 ^ (uint64_t)len;
        h1 += h2; h2 += h1;
        h1 = fmix64(h1); h2 = fmix64(h2);
        h1 += h2; h2 += h1;
        o.m1 = h1; o.m2 = h2;
    }

    // ---- XXH3 finish: the reads from the end ----
    if (mid) {
        uint64_t k1, k2;
        read16_any<UPPER>(tile, start + len - 16, k1, k2);
        const uint64_t a = xxh3_avalanche(st.acc) + st.acc_b + xxh3_mix16_k(k1, k2, KSec<119>::v, KSec<127>::v);
        o.xxh3 = xxh3_avalanche(a);
    } else {
        uint64_t a = st.acc;
#define SHB_F3_BACK(I)                                                                        \
    if ((uint32_t)(I) < pairs) {                                                              \
        uint64_t k1, k2;                                                                      \
        read16_any<UPPER>(tile, start + len - 16 * ((I) + 1), k1, k2);                        \
        a += xxh3_mix16_k(k1, k2, KSec<32 * (I) + 16>::v, KSec<32 * (I) + 24>::v);            \
    }
        SHB_F3_BACK(0) SHB_F3_BACK(1) SHB_F3_BACK(2) SHB_F3_BACK(3)
#undef SHB_F3_BACK
        o.xxh3 = xxh3_avalanche(a);
    }
    return in.wsf;
}

}  // namespace shb
