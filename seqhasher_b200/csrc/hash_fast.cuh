// hash_fast.cuh — the non-cryptographic hashes: XXH64, XXH3-64, MurmurHash3 x64_128, CityHash128,
// ntHash-1 (forward), rapidhash (reference port variant) and its 32-bit fold.
//
// Replaces xxhash.Sum64 (seqhasher.go:308-310), xxh3.Hash (:311-313), city.Hash128 (:314-316),
// murmur3.Sum128 (:317-319), nthash NewHasher+Next(false) (:320-327), rapidhash.Hash (:337-339)
// and the rapidhash32 fold (:340-343).  These are HBM-bound (about 1-2 integer ops per byte), so
// every hasher takes a reader and touches each residue byte once; 64x64->128 products are
// mul.lo + mul.hi pairs on the FMA pipe.
#pragma once
#include "common.cuh"

namespace shb {

// ------------------------------------------------------------------------------------ XXH64

#define XP64_1 0x9E3779B185EBCA87ULL
#define XP64_2 0xC2B2AE3D27D4EB4FULL
#define XP64_3 0x165667B19E3779F9ULL
#define XP64_4 0x85EBCA77C2B2AE63ULL
#define XP64_5 0x27D4EB2F165667C5ULL
#define XP32_1 0x9E3779B1ULL
#define XP32_2 0x85EBCA77ULL
#define XP32_3 0xC2B2AE3DULL
#define XPMX1 0x165667919E3779F9ULL
#define XPMX2 0x9FB21C651E98DF25ULL

__device__ __forceinline__ uint64_t xxh64_round(uint64_t acc, uint64_t x)
{
    return rotl64(acc + x * XP64_2, 31) * XP64_1;
}
__device__ __forceinline__ uint64_t xxh64_avalanche(uint64_t h)
{
    h ^= h >> 33; h *= XP64_2; h ^= h >> 29; h *= XP64_3; h ^= h >> 32;
    return h;
}

template <class R>
__device__ __forceinline__ uint64_t xxh64_record(const R &r, uint32_t len)
{
    uint32_t p = 0;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = XP64_1 + XP64_2, v2 = XP64_2, v3 = 0, v4 = 0ULL - XP64_1;
        const uint32_t lim = len - 32;
        auto in = r.stream(0);
        do {
            uint32_t a[4], b[4];
            in.next(a);
            in.next(b);
            v1 = xxh64_round(v1, mk64(a[0], a[1]));
            v2 = xxh64_round(v2, mk64(a[2], a[3]));
            v3 = xxh64_round(v3, mk64(b[0], b[1]));
            v4 = xxh64_round(v4, mk64(b[2], b[3]));
            p += 32;
        } while (p <= lim);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = (h ^ xxh64_round(0, v1)) * XP64_1 + XP64_4;
        h = (h ^ xxh64_round(0, v2)) * XP64_1 + XP64_4;
        h = (h ^ xxh64_round(0, v3)) * XP64_1 + XP64_4;
        h = (h ^ xxh64_round(0, v4)) * XP64_1 + XP64_4;
    } else {
        h = XP64_5;
    }
    h += (uint64_t)len;
    while (p + 8 <= len) {
        h ^= xxh64_round(0, r.u64(p));
        h = rotl64(h, 27) * XP64_1 + XP64_4;
        p += 8;
    }
    if (p + 4 <= len) {
        h ^= (uint64_t)r.u32(p) * XP64_1;
        h = rotl64(h, 23) * XP64_2 + XP64_3;
        p += 4;
    }
    if (p < len) {
        uint32_t t = r.u32(p);   // 1..3 tail bytes, p % 4 == 0
        for (; p < len; p++) {
            h ^= (uint64_t)(t & 0xffu) * XP64_5;
            h = rotl64(h, 11) * XP64_1;
            t >>= 8;
        }
    }
    return xxh64_avalanche(h);
}

// ------------------------------------------------------------------------------------- XXH3

// default 192-byte secret of XXH3 (kSecret), as 64-bit LE words at byte offsets 0,8,..,184 plus the
// unaligned views the algorithm needs (offsets 3, 11, 19, .. and 119, 121, 127).
__constant__ __align__(16) uint8_t c_xxh3_secret[192 + 8] = {
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
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
    0, 0, 0, 0, 0, 0, 0, 0};

// 64-bit LE word of the secret at a compile-time or warp-uniform byte offset
__device__ __forceinline__ uint64_t xsec(uint32_t off)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(c_xxh3_secret);
    uint32_t i = off >> 2, s = (off & 3) * 8;
    uint32_t w0 = w[i], w1 = w[i + 1], w2 = w[i + 2];
    return mk64(__funnelshift_r(w0, w1, s), __funnelshift_r(w1, w2, s));
}

__device__ __forceinline__ uint64_t xxh3_avalanche(uint64_t h)
{
    h ^= h >> 37; h *= XPMX1; h ^= h >> 32;
    return h;
}
// mix16 with the input at an ALIGNED (multiple of 4) record offset
template <class R>
__device__ __forceinline__ uint64_t xxh3_mix16(const R &r, uint32_t off, uint32_t soff)
{
    return mulfold64(r.u64(off) ^ xsec(soff), r.u64(off + 8) ^ xsec(soff + 8));
}
// mix16 with the input at any record offset (the "from the end" reads)
template <class R>
__device__ __forceinline__ uint64_t xxh3_mix16x(const R &r, uint32_t off, uint32_t soff)
{
    return mulfold64(r.u64x(off) ^ xsec(soff), r.u64x(off + 8) ^ xsec(soff + 8));
}

template <bool ANY, class R>
__device__ __forceinline__ void xxh3_accumulate(uint64_t (&acc)[8], const R &r, uint32_t off, uint32_t soff)
{
    uint64_t d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = ANY ? r.u64x(off + 8 * i) : r.u64(off + 8 * i);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t k = d[i] ^ xsec(soff + 8 * i);
        acc[i ^ 1] += d[i];
        acc[i] += (uint64_t)(uint32_t)k * (uint64_t)(uint32_t)(k >> 32);
    }
}

// one 64-byte stripe taken from the reader's sequential stream (four 16-byte groups)
template <class S>
__device__ __forceinline__ void xxh3_accumulate_stream(uint64_t (&acc)[8], S &in, uint32_t soff)
{
    uint64_t d[8];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t w[4];
        in.next(w);
        d[2 * g] = mk64(w[0], w[1]);
        d[2 * g + 1] = mk64(w[2], w[3]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t k = d[i] ^ xsec(soff + 8 * i);
        acc[i ^ 1] += d[i];
        acc[i] += (uint64_t)(uint32_t)k * (uint64_t)(uint32_t)(k >> 32);
    }
}

template <class R>
__device__ __forceinline__ uint64_t xxh3_record(const R &r, uint32_t n)
{
    if (n <= 16) {
        if (n > 8) {
            uint64_t lo = r.u64(0) ^ (xsec(24) ^ xsec(32));
            uint64_t hi = r.u64x(n - 8) ^ (xsec(40) ^ xsec(48));
            return xxh3_avalanche((uint64_t)n + bswap64(lo) + hi + mulfold64(lo, hi));
        }
        if (n >= 4) {
            uint32_t in1 = r.u32(0), in2 = r.u32x(n - 4);
            uint64_t h = ((uint64_t)in2 + ((uint64_t)in1 << 32)) ^ (xsec(8) ^ xsec(16));
            h ^= rotl64(h, 49) ^ rotl64(h, 24);
            h *= XPMX2;
            h ^= (h >> 35) + (uint64_t)n;
            h *= XPMX2;
            return h ^ (h >> 28);
        }
        if (n > 0) {
            uint32_t w = r.u32(0);
            uint32_t c1 = w & 0xffu, c2 = (w >> (8 * (n >> 1))) & 0xffu, c3 = (w >> (8 * (n - 1))) & 0xffu;
            uint32_t combined = (c1 << 16) | (c2 << 24) | c3 | (n << 8);
            uint64_t flip = xsec(0);
            return xxh64_avalanche((uint64_t)(combined ^ ((uint32_t)flip ^ (uint32_t)(flip >> 32))));
        }
        return xxh64_avalanche(xsec(56) ^ xsec(64));
    }
    if (n <= 128) {
        uint64_t acc = (uint64_t)n * XP64_1;
        const uint32_t pairs = (n + 31) >> 5;
        for (uint32_t i = 0; i < pairs; i++) {
            acc += xxh3_mix16(r, 16 * i, 32 * i);
            acc += xxh3_mix16x(r, n - 16 * (i + 1), 32 * i + 16);
        }
        return xxh3_avalanche(acc);
    }
    if (n <= 240) {
        uint64_t acc = (uint64_t)n * XP64_1;
#pragma unroll
        for (int i = 0; i < 8; i++) acc += xxh3_mix16(r, 16 * i, 16 * i);
        acc = xxh3_avalanche(acc);
        const uint32_t rounds = n >> 4;
        for (uint32_t i = 8; i < rounds; i++) acc += xxh3_mix16(r, 16 * i, 16 * (i - 8) + 3);
        acc += xxh3_mix16x(r, n - 16, 119);
        return xxh3_avalanche(acc);
    }
    uint64_t acc[8] = {XP32_3, XP64_1, XP64_2, XP64_3, XP64_4, XP32_2, XP64_5, XP32_1};
    const uint32_t nblocks = (n - 1) >> 10;
    auto in = r.stream(0);   // the stripes are consecutive from offset 0: one 128-bit load per 16 bytes from global memory
    for (uint32_t b = 0; b < nblocks; b++) {
#pragma unroll 1
        for (uint32_t s = 0; s < 16; s++) xxh3_accumulate_stream(acc, in, 8 * s);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t a = acc[i];
            a ^= a >> 47;
            a ^= xsec(128 + 8 * i);
            acc[i] = a * XP32_1;
        }
    }
    const uint32_t nstripes = ((n - 1) - (nblocks << 10)) >> 6;
#pragma unroll 1
    for (uint32_t s = 0; s < nstripes; s++) xxh3_accumulate_stream(acc, in, 8 * s);
    xxh3_accumulate<true>(acc, r, n - 64, 121);
    uint64_t res = (uint64_t)n * XP64_1;
#pragma unroll
    for (int i = 0; i < 4; i++)
        res += mulfold64(acc[2 * i] ^ xsec(11 + 16 * i), acc[2 * i + 1] ^ xsec(19 + 16 * i));
    return xxh3_avalanche(res);
}

// ---------------------------------------------------------------------------------- murmur3

__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

template <class R>
__device__ __forceinline__ void murmur3_record(const R &r, uint32_t len, uint64_t &o1, uint64_t &o2)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 0, h2 = 0;
    const uint32_t nblocks = len >> 4;
    auto in = r.stream(0);
#pragma unroll 4
    for (uint32_t i = 0; i < nblocks; i++) {
        uint32_t kw[4];
        in.next(kw);
        uint64_t k1 = mk64(kw[0], kw[1]), k2 = mk64(kw[2], kw[3]);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
    }
    const uint32_t rem = len & 15u;
    if (rem) {
        const uint32_t t = nblocks << 4;
        // zero-extended tail: bytes past len masked off
        uint64_t k1 = r.u64(t), k2 = rem > 8 ? r.u64(t + 8) : 0;
        if (rem < 8) k1 &= (1ULL << (8 * rem)) - 1ULL;
        if (rem > 8) {
            k2 &= (1ULL << (8 * (rem - 8))) - 1ULL;
            k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    o1 = h1; o2 = h2;
}

// --------------------------------------------------------------------------------- CityHash

#define CK0 0xc3a5c85c97cb3127ULL
#define CK1 0xb492b66fbe98f273ULL
#define CK2 0x9ae16a3b2f90404fULL
#define CKMUL 0x9ddfea08eb382d69ULL

__device__ __forceinline__ uint64_t city_shift_mix(uint64_t v) { return v ^ (v >> 47); }
__device__ __forceinline__ uint64_t city_len16(uint64_t u, uint64_t v, uint64_t mul)
{
    uint64_t a = (u ^ v) * mul;
    a ^= a >> 47;
    uint64_t b = (v ^ a) * mul;
    b ^= b >> 47;
    return b * mul;
}
__device__ __forceinline__ uint64_t city_len16(uint64_t u, uint64_t v) { return city_len16(u, v, CKMUL); }

// HashLen0to16 over record bytes [s, s+len)
template <class R>
__device__ __forceinline__ uint64_t city_len0to16(const R &r, uint32_t s, uint32_t len)
{
    if (len >= 8) {
        uint64_t mul = CK2 + (uint64_t)len * 2;
        uint64_t a = r.u64x(s) + CK2;
        uint64_t b = r.u64x(s + len - 8);
        uint64_t c = rotr64(b, 37) * mul + a;
        uint64_t d = (rotr64(a, 25) + b) * mul;
        return city_len16(c, d, mul);
    }
    if (len >= 4) {
        uint64_t mul = CK2 + (uint64_t)len * 2;
        uint64_t a = r.u32x(s);
        return city_len16((uint64_t)len + (a << 3), r.u32x(s + len - 4), mul);
    }
    if (len > 0) {
        uint32_t a = r.u8(s), b = r.u8(s + (len >> 1)), c = r.u8(s + len - 1);
        uint32_t y = a + (b << 8);
        uint32_t z = len + (c << 2);
        return city_shift_mix((uint64_t)y * CK2 ^ (uint64_t)z * CK0) * CK2;
    }
    return CK2;
}

struct CityPair { uint64_t first, second; };

__device__ __forceinline__ CityPair city_weak32(uint64_t w, uint64_t x, uint64_t y, uint64_t z, uint64_t a, uint64_t b)
{
    a += w;
    b = rotr64(b + a + z, 21);
    uint64_t c = a;
    a += x;
    a += y;
    b += rotr64(a, 44);
    CityPair p = {a + z, b + c};
    return p;
}
template <class R>
__device__ __forceinline__ CityPair city_weak32_at(const R &r, uint32_t s, uint64_t a, uint64_t b)
{
    return city_weak32(r.u64x(s), r.u64x(s + 8), r.u64x(s + 16), r.u64x(s + 24), a, b);
}

// CityHash128WithSeed over record bytes [s0, s0+len)
template <class R>
__device__ __forceinline__ CityPair city128_with_seed(const R &r, uint32_t s0, uint32_t len, uint64_t seed_lo,
                                                      uint64_t seed_hi)
{
    uint32_t s = s0;
    if (len < 128) {   // CityMurmur
        uint64_t a = seed_lo, b = seed_hi, c, d;
        if (len <= 16) {
            a = city_shift_mix(a * CK1) * CK1;
            c = b * CK1 + city_len0to16(r, s, len);
            d = city_shift_mix(a + (len >= 8 ? r.u64x(s) : c));
        } else {
            c = city_len16(r.u64x(s + len - 8) + CK1, a);
            d = city_len16(b + (uint64_t)len, c + r.u64x(s + len - 16));
            a += d;
            int32_t l = (int32_t)len - 16;
            do {
                a ^= city_shift_mix(r.u64x(s) * CK1) * CK1;
                a *= CK1;
                b ^= a;
                c ^= city_shift_mix(r.u64x(s + 8) * CK1) * CK1;
                c *= CK1;
                d ^= c;
                s += 16;
                l -= 16;
            } while (l > 0);
        }
        a = city_len16(a, c);
        b = city_len16(d, b);
        CityPair p = {a ^ b, city_len16(b, a)};
        return p;
    }
    CityPair v, w;
    uint64_t x = seed_lo, y = seed_hi, z = (uint64_t)len * CK1;
    v.first = rotr64(y ^ CK1, 49) * CK1 + r.u64x(s);
    v.second = rotr64(v.first, 42) * CK1 + r.u64x(s + 8);
    w.first = rotr64(y + z, 35) * CK1 + x;
    w.second = rotr64(x + r.u64x(s + 88), 53) * CK1;
    auto in = r.stream(s);
    do {
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            uint64_t q[8];   // the 64-byte half-block, one 16-byte group at a time
#pragma unroll
            for (int g = 0; g < 4; g++) {
                uint32_t t4[4];
                in.next(t4);
                q[2 * g] = mk64(t4[0], t4[1]);
                q[2 * g + 1] = mk64(t4[2], t4[3]);
            }
            x = rotr64(x + y + v.first + q[1], 37) * CK1;
            y = rotr64(y + v.second + q[6], 42) * CK1;
            x ^= w.second;
            y += v.first + q[5];
            z = rotr64(z + w.first, 33) * CK1;
            v = city_weak32(q[0], q[1], q[2], q[3], v.second * CK1, x + w.first);
            w = city_weak32(q[4], q[5], q[6], q[7], z + w.second, y + q[2]);
            uint64_t t = z; z = x; x = t;
            s += 64;
        }
        len -= 128;
    } while (len >= 128);
    x += rotr64(v.first + z, 49) * CK0;
    y = y * CK0 + rotr64(w.second, 37);
    z = z * CK0 + rotr64(w.first, 27);
    w.first *= 9;
    v.first *= CK0;
    for (uint32_t tail_done = 0; tail_done < len;) {
        tail_done += 32;
        y = rotr64(x + y, 42) * CK0 + v.second;
        w.first += r.u64x(s + len - tail_done + 16);
        x = x * CK0 + w.first;
        z += w.second + r.u64x(s + len - tail_done);
        w.second += v.first;
        v = city_weak32_at(r, s + len - tail_done, v.first + z, v.second);
        v.first *= CK0;
    }
    x = city_len16(x, v.first);
    y = city_len16(y + z, w.first);
    CityPair p = {city_len16(x + v.second, w.second) + y, city_len16(x + w.second, y + v.second)};
    return p;
}

// low = .first, high = .second; printed High||Low (seqhasher.go:316)
template <class R>
__device__ __forceinline__ CityPair city128_record(const R &r, uint32_t len)
{
    if (len >= 16) return city128_with_seed(r, 16, len - 16, r.u64(0), r.u64(8) + CK0);
    return city128_with_seed(r, 0, len, CK0, CK1);
}

// ----------------------------------------------------------------------------------- ntHash

#define NT_SEED_A 0x3c8bfbb395c60474ULL
#define NT_SEED_C 0x3193c18562a02b4cULL
#define NT_SEED_G 0x20323ed082572324ULL
#define NT_SEED_T 0x295549f54be24456ULL

// ntHash v1 seedTab restricted to its non-zero entries: A/a, C/c, G/g, T/t, U/u and the
// complement look-up slots 1,3,4,5,7 of the upstream table.
__device__ __forceinline__ uint64_t nt_seed(uint32_t c)
{
    uint64_t s = 0;
    uint32_t u = c & 0xdfu;   // fold case
    if (u == 'A') s = NT_SEED_A;
    if (u == 'C') s = NT_SEED_C;
    if (u == 'G') s = NT_SEED_G;
    if (u == 'T' || u == 'U') s = NT_SEED_T;
    if (c < 8) {
        if (c == 4 || c == 5) s = NT_SEED_A;
        if (c == 7) s = NT_SEED_C;
        if (c == 3) s = NT_SEED_G;
        if (c == 1) s = NT_SEED_T;
    }
    return s;
}

// The same table for shared-memory look-ups: seeds pre-rotated by 0..3, so that a 4-byte word folds as
// rol(h,4) ^ L3[b0] ^ L2[b1] ^ L1[b2] ^ L0[b3] with one 64-bit load per byte instead of the compare chain above
// (about 20 ALU instructions per byte).  nt_seed_table() is evaluated at compile time into device memory.
constexpr uint64_t nt_seed_const(uint32_t c)
{
    const uint32_t u = c & 0xdfu;
    uint64_t s = 0;
    if (u == 'A') s = NT_SEED_A;
    if (u == 'C') s = NT_SEED_C;
    if (u == 'G') s = NT_SEED_G;
    if (u == 'T' || u == 'U') s = NT_SEED_T;
    if (c < 8) {
        if (c == 4 || c == 5) s = NT_SEED_A;
        if (c == 7) s = NT_SEED_C;
        if (c == 3) s = NT_SEED_G;
        if (c == 1) s = NT_SEED_T;
    }
    return s;
}
struct NtSeedTable {
    uint64_t v[256];
    constexpr NtSeedTable() : v{}
    {
        for (uint32_t c = 0; c < 256; c++) v[c] = nt_seed_const(c);
    }
};
__device__ const NtSeedTable g_nt_seed_table = NtSeedTable();

constexpr uint32_t NT_LUT_WORDS = 4 * 256;   // uint64 entries: [rotation k][byte]

// The entry of byte c sits at index c ^ (c >> 5): 64-bit entries 16 apart share their two banks, and 'A'/'a',
// 'C'/'c', ... are exactly 32 apart, so soft-masked (mixed-case) sequence would make every look-up a 2-way bank
// conflict.  With the swizzle the ten bytes ACGTN/acgtn land in ten different bank pairs.  SWZ is on when the kernel
// sees the bytes as they are (case-sensitive mode); after upper-casing there is nothing to separate and the two
// extra instructions per word would only cost (3-6 %).
template <bool SWZ>
__device__ __forceinline__ uint32_t nt_swz(uint32_t c) { return SWZ ? c ^ (c >> 5) : c; }

// cooperative fill by the whole CTA (coalesced reads of the 2 KiB seed table); the caller synchronises
template <bool SWZ>
__device__ __forceinline__ void nt_fill_lut(uint64_t *lut)
{
    for (uint32_t c = threadIdx.x; c < 256; c += blockDim.x) {
        const uint64_t s = g_nt_seed_table.v[c];
        const uint32_t k = nt_swz<SWZ>(c);
        lut[k] = s;
        lut[256 + k] = rotl64(s, 1);
        lut[512 + k] = rotl64(s, 2);
        lut[768 + k] = rotl64(s, 3);
    }
}

// fold 4 bytes (memory order b0..b3) into h
template <bool SWZ>
__device__ __forceinline__ uint64_t nt_fold4(uint64_t h, uint32_t w, const uint64_t *lut)
{
    if (SWZ) w ^= (w >> 5) & 0x07070707u;   // nt_swz on the four bytes at once
    return rotl64(h, 4) ^ lut[768 + (w & 0xffu)] ^ lut[512 + ((w >> 8) & 0xffu)] ^ lut[256 + ((w >> 16) & 0xffu)] ^ lut[w >> 24];
}

template <bool SWZ, class R>
__device__ __forceinline__ uint64_t nthash_record_lut(const R &r, uint32_t len, const uint64_t *lut)
{
    uint64_t h = 0;
    uint32_t p = 0;
    for (; p + 4 <= len; p += 4) h = nt_fold4<SWZ>(h, r.u32(p), lut);
    if (p < len) {
        uint32_t w = r.u32(p);
        for (; p < len; p++) {
            h = rotl64(h, 1) ^ lut[nt_swz<SWZ>(w & 0xffu)];
            w >>= 8;
        }
    }
    return h;
}

// the same from global memory: 16-byte groups through the reader's Stream (one 128-bit load each)
template <bool SWZ, class R>
__device__ __forceinline__ uint64_t nthash_record_lut_stream(const R &r, uint32_t len, const uint64_t *lut)
{
    uint64_t h = 0;
    uint32_t p = 0;
    if (len >= 16) {
        auto in = r.stream(0);
        do {
            uint32_t w[4];
            in.next(w);
#pragma unroll
            for (int k = 0; k < 4; k++) h = nt_fold4<SWZ>(h, w[k], lut);
            p += 16;
        } while (p + 16 <= len);
    }
    for (; p + 4 <= len; p += 4) h = nt_fold4<SWZ>(h, r.u32(p), lut);
    if (p < len) {
        uint32_t w = r.u32(p);
        for (; p < len; p++) {
            h = rotl64(h, 1) ^ lut[nt_swz<SWZ>(w & 0xffu)];
            w >>= 8;
        }
    }
    return h;
}

// forward hash of the whole record: h = rol(h,1) ^ seed[c] over all bytes
template <class R>
__device__ __forceinline__ uint64_t nthash_record(const R &r, uint32_t len)
{
    uint64_t h = 0;
    uint32_t p = 0;
    for (; p + 4 <= len; p += 4) {
        uint32_t w = r.u32(p);
        h = rotl64(h, 4) ^ rotl64(nt_seed(w & 0xffu), 3) ^ rotl64(nt_seed((w >> 8) & 0xffu), 2) ^
            rotl64(nt_seed((w >> 16) & 0xffu), 1) ^ nt_seed(w >> 24);
    }
    if (p < len) {
        uint32_t w = r.u32(p);
        for (; p < len; p++) {
            h = rotl64(h, 1) ^ nt_seed(w & 0xffu);
            w >>= 8;
        }
    }
    return h;
}

// -------------------------------------------------------------------------------- rapidhash

#define RAPID_SEED 0xbdd89aa982704029ULL
#define RAPID_S0 0x2d358dccaa6c78a5ULL
#define RAPID_S1 0x8bb84b93962eacc9ULL
#define RAPID_S2 0x4b33a62ed433d4a3ULL

// rapidhash v1 as restated by github.com/vkudryk/rapidhash-go (final mix keyed with secret[2])
template <class R>
__device__ __forceinline__ uint64_t rapidhash_record(const R &r, uint32_t len)
{
    uint64_t seed = RAPID_SEED, a, b;
    seed ^= mulfold64(seed ^ RAPID_S0, RAPID_S1) ^ (uint64_t)len;
    if (len <= 16) {
        if (len >= 4) {
            const uint32_t plast = len - 4;
            a = ((uint64_t)r.u32(0) << 32) | r.u32x(plast);
            const uint32_t delta = (len & 24u) >> (len >> 3);
            b = ((uint64_t)r.u32(delta) << 32) | r.u32x(plast - delta);
        } else if (len > 0) {
            uint32_t w = r.u32(0);
            a = ((uint64_t)(w & 0xffu) << 56) | ((uint64_t)((w >> (8 * (len >> 1))) & 0xffu) << 32) |
                ((w >> (8 * (len - 1))) & 0xffu);
            b = 0;
        } else {
            a = b = 0;
        }
    } else {
        uint32_t i = len, p = 0;
        if (i > 48) {
            uint64_t see1 = seed, see2 = seed;
            auto in = r.stream(0);
            do {
                uint32_t x[4], y[4], z[4];
                in.next(x);
                in.next(y);
                in.next(z);
                seed = mulfold64(mk64(x[0], x[1]) ^ RAPID_S0, mk64(x[2], x[3]) ^ seed);
                see1 = mulfold64(mk64(y[0], y[1]) ^ RAPID_S1, mk64(y[2], y[3]) ^ see1);
                see2 = mulfold64(mk64(z[0], z[1]) ^ RAPID_S2, mk64(z[2], z[3]) ^ see2);
                p += 48;
                i -= 48;
            } while (i >= 48);
            seed ^= see1 ^ see2;
        }
        if (i > 16) {
            seed = mulfold64(r.u64(p) ^ RAPID_S2, r.u64(p + 8) ^ seed ^ RAPID_S1);
            if (i > 32) seed = mulfold64(r.u64(p + 16) ^ RAPID_S2, r.u64(p + 24) ^ seed);
        }
        a = r.u64x(len - 16);
        b = r.u64x(len - 8);
    }
    a ^= RAPID_S1;
    b ^= seed;
    uint64_t lo = a * b, hi = __umul64hi(a, b);
    return mulfold64(lo ^ RAPID_S0 ^ (uint64_t)len, hi ^ RAPID_S2);
}

}  // namespace shb
