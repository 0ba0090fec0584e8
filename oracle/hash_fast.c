/*
 * hash_fast.c — ORACLE (test infrastructure): the non-cryptographic hashes.
 *
 * xxhash   : xxhash.Sum64 of github.com/cespare/xxhash/v2 v2.3.0 (go.mod:6),
 *            seqhasher.go:308-310 — XXH64, seed 0.
 * xxh3     : xxh3.Hash of github.com/zeebo/xxh3 v1.1.0 (go.mod:15),
 *            seqhasher.go:311-313 — XXH3_64bits, seed 0, default 192-byte secret.
 * murmur3  : murmur3.Sum128 of github.com/spaolacci/murmur3 v1.1.0 (go.mod:11),
 *            seqhasher.go:317-319 — MurmurHash3_x64_128, seed 0.
 * cityhash : city.Hash128 of github.com/go-faster/city v1.0.1 (go.mod:9),
 *            seqhasher.go:314-316 — CityHash128 v1.1 (not the ClickHouse 1.0.2 variant).
 * nthash   : nthash.NewHasher(&data, len).Next(false) of github.com/will-rowe/nthash
 *            v0.4.0 (go.mod:13), seqhasher.go:320-327 — ntHash-1 forward hash, k = len.
 * rapidhash: rapidhash.Hash of github.com/vkudryk/rapidhash-go
 *            v0.0.0-20250522125531-a03e5539e1bf (go.mod:12), seqhasher.go:337-343 —
 *            rapidhash v1 whose final mix uses secret[2] (the only variant that
 *            reproduces the reference golden seqhasher_test.go:328).
 */
#include "oracle_internal.h"

/* ================================================================= XXH64 */

#define P64_1 0x9E3779B185EBCA87ULL
#define P64_2 0xC2B2AE3D27D4EB4FULL
#define P64_3 0x165667B19E3779F9ULL
#define P64_4 0x85EBCA77C2B2AE63ULL
#define P64_5 0x27D4EB2F165667C5ULL
#define P32_1 0x9E3779B1U
#define P32_2 0x85EBCA77U
#define P32_3 0xC2B2AE3DU
#define PRIME_MX1 0x165667919E3779F9ULL
#define PRIME_MX2 0x9FB21C651E98DF25ULL

static uint64_t xxh64_round(uint64_t acc, uint64_t x)
{
    return rol64(acc + x * P64_2, 31) * P64_1;
}
static uint64_t xxh64_merge(uint64_t h, uint64_t v)
{
    return (h ^ xxh64_round(0, v)) * P64_1 + P64_4;
}
static uint64_t xxh64_avalanche(uint64_t h)
{
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32;
    return h;
}

uint64_t sho_xxh64(const uint8_t *p, size_t len)
{
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = P64_1 + P64_2, v2 = P64_2, v3 = 0, v4 = 0 - P64_1;
        do {
            v1 = xxh64_round(v1, load_le64(p));
            v2 = xxh64_round(v2, load_le64(p + 8));
            v3 = xxh64_round(v3, load_le64(p + 16));
            v4 = xxh64_round(v4, load_le64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = rol64(v1, 1) + rol64(v2, 7) + rol64(v3, 12) + rol64(v4, 18);
        h = xxh64_merge(h, v1); h = xxh64_merge(h, v2);
        h = xxh64_merge(h, v3); h = xxh64_merge(h, v4);
    } else {
        h = P64_5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) {
        h ^= xxh64_round(0, load_le64(p));
        h = rol64(h, 27) * P64_1 + P64_4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)load_le32(p) * P64_1;
        h = rol64(h, 23) * P64_2 + P64_3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * P64_5;
        h = rol64(h, 11) * P64_1;
        p++;
    }
    return xxh64_avalanche(h);
}

/* ================================================================== XXH3 */

static const uint8_t xxh3_secret[192] = {
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
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e};

static uint64_t fold64(uint64_t a, uint64_t b)
{
    uint64_t lo, hi;
    mul64x64(a, b, &lo, &hi);
    return lo ^ hi;
}
static uint64_t xxh3_avalanche(uint64_t h)
{
    h ^= h >> 37; h *= PRIME_MX1; h ^= h >> 32;
    return h;
}
static uint64_t xxh3_mix16(const uint8_t *p, const uint8_t *s)
{
    return fold64(load_le64(p) ^ load_le64(s), load_le64(p + 8) ^ load_le64(s + 8));
}
static void xxh3_accumulate(uint64_t acc[8], const uint8_t *p, const uint8_t *s)
{
    for (int i = 0; i < 8; i++) {
        uint64_t d = load_le64(p + 8 * i);
        uint64_t k = d ^ load_le64(s + 8 * i);
        acc[i ^ 1] += d;
        acc[i] += (uint64_t)(uint32_t)k * (uint64_t)(uint32_t)(k >> 32);
    }
}
static void xxh3_scramble(uint64_t acc[8], const uint8_t *s)
{
    for (int i = 0; i < 8; i++) {
        uint64_t a = acc[i];
        a ^= a >> 47;
        a ^= load_le64(s + 8 * i);
        a *= P32_1;
        acc[i] = a;
    }
}

uint64_t sho_xxh3_64(const uint8_t *p, size_t n)
{
    const uint8_t *K = xxh3_secret;
    if (n == 0) return xxh64_avalanche(load_le64(K + 56) ^ load_le64(K + 64));
    if (n <= 3) {
        uint32_t c1 = p[0], c2 = p[n >> 1], c3 = p[n - 1];
        uint32_t combined = (c1 << 16) | (c2 << 24) | c3 | ((uint32_t)n << 8);
        uint64_t bitflip = (uint64_t)(load_le32(K) ^ load_le32(K + 4));
        return xxh64_avalanche((uint64_t)combined ^ bitflip);
    }
    if (n <= 8) {
        uint32_t in1 = load_le32(p), in2 = load_le32(p + n - 4);
        uint64_t bitflip = load_le64(K + 8) ^ load_le64(K + 16);
        uint64_t h = ((uint64_t)in2 + ((uint64_t)in1 << 32)) ^ bitflip;
        h ^= rol64(h, 49) ^ rol64(h, 24);
        h *= PRIME_MX2;
        h ^= (h >> 35) + (uint64_t)n;
        h *= PRIME_MX2;
        return h ^ (h >> 28);
    }
    if (n <= 16) {
        uint64_t lo = load_le64(p) ^ (load_le64(K + 24) ^ load_le64(K + 32));
        uint64_t hi = load_le64(p + n - 8) ^ (load_le64(K + 40) ^ load_le64(K + 48));
        uint64_t acc = (uint64_t)n + __builtin_bswap64(lo) + hi + fold64(lo, hi);
        return xxh3_avalanche(acc);
    }
    if (n <= 128) {
        uint64_t acc = (uint64_t)n * P64_1;
        size_t pairs = (n + 31) / 32; /* 1..4 */
        for (size_t i = 0; i < pairs; i++) {
            acc += xxh3_mix16(p + 16 * i, K + 32 * i);
            acc += xxh3_mix16(p + n - 16 * (i + 1), K + 32 * i + 16);
        }
        return xxh3_avalanche(acc);
    }
    if (n <= 240) {
        uint64_t acc = (uint64_t)n * P64_1;
        size_t rounds = n / 16;
        for (size_t i = 0; i < 8; i++) acc += xxh3_mix16(p + 16 * i, K + 16 * i);
        acc = xxh3_avalanche(acc);
        for (size_t i = 8; i < rounds; i++) acc += xxh3_mix16(p + 16 * i, K + 16 * (i - 8) + 3);
        acc += xxh3_mix16(p + n - 16, K + 119);
        return xxh3_avalanche(acc);
    }
    uint64_t acc[8] = {P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1};
    size_t nblocks = (n - 1) / 1024;
    for (size_t b = 0; b < nblocks; b++) {
        for (int s = 0; s < 16; s++) xxh3_accumulate(acc, p + 1024 * b + 64 * s, K + 8 * s);
        xxh3_scramble(acc, K + 128);
    }
    size_t nstripes = ((n - 1) - 1024 * nblocks) / 64;
    for (size_t s = 0; s < nstripes; s++) xxh3_accumulate(acc, p + 1024 * nblocks + 64 * s, K + 8 * s);
    xxh3_accumulate(acc, p + n - 64, K + 121);
    uint64_t r = (uint64_t)n * P64_1;
    for (int i = 0; i < 4; i++)
        r += fold64(acc[2 * i] ^ load_le64(K + 11 + 16 * i), acc[2 * i + 1] ^ load_le64(K + 19 + 16 * i));
    return xxh3_avalanche(r);
}

/* =============================================================== murmur3 */

static uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

void sho_murmur3_x64_128(const uint8_t *p, size_t len, uint64_t *out1, uint64_t *out2)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 0, h2 = 0;
    size_t nblocks = len / 16;
    for (size_t i = 0; i < nblocks; i++) {
        uint64_t k1 = load_le64(p + 16 * i), k2 = load_le64(p + 16 * i + 8);
        k1 *= c1; k1 = rol64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rol64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rol64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rol64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *t = p + 16 * nblocks;
    size_t rem = len & 15;
    uint64_t k1 = 0, k2 = 0;
    for (size_t i = rem; i > 8; i--) k2 = (k2 << 8) | t[i - 1];
    if (rem > 8) { k2 *= c2; k2 = rol64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (size_t i = rem < 8 ? rem : 8; i > 0; i--) k1 = (k1 << 8) | t[i - 1];
    if (rem > 0) { k1 *= c1; k1 = rol64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    *out1 = h1; *out2 = h2;
}

/* ============================================================== CityHash */

#define CK0 0xc3a5c85c97cb3127ULL
#define CK1 0xb492b66fbe98f273ULL
#define CK2 0x9ae16a3b2f90404fULL
#define CKMUL 0x9ddfea08eb382d69ULL

typedef struct { uint64_t first, second; } u128pair;

static uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }
static uint64_t city_hash_len16_mul(uint64_t u, uint64_t v, uint64_t mul)
{
    uint64_t a = (u ^ v) * mul;
    a ^= a >> 47;
    uint64_t b = (v ^ a) * mul;
    b ^= b >> 47;
    return b * mul;
}
static uint64_t city_hash_len16(uint64_t u, uint64_t v) { return city_hash_len16_mul(u, v, CKMUL); }

static uint64_t city_hash_len0to16(const uint8_t *s, size_t len)
{
    if (len >= 8) {
        uint64_t mul = CK2 + len * 2;
        uint64_t a = load_le64(s) + CK2;
        uint64_t b = load_le64(s + len - 8);
        uint64_t c = ror64(b, 37) * mul + a;
        uint64_t d = (ror64(a, 25) + b) * mul;
        return city_hash_len16_mul(c, d, mul);
    }
    if (len >= 4) {
        uint64_t mul = CK2 + len * 2;
        uint64_t a = load_le32(s);
        return city_hash_len16_mul(len + (a << 3), load_le32(s + len - 4), mul);
    }
    if (len > 0) {
        uint8_t a = s[0], b = s[len >> 1], c = s[len - 1];
        uint32_t y = (uint32_t)a + ((uint32_t)b << 8);
        uint32_t z = (uint32_t)len + ((uint32_t)c << 2);
        return shift_mix(y * CK2 ^ z * CK0) * CK2;
    }
    return CK2;
}

static u128pair weak_hash_len32_seeds(uint64_t w, uint64_t x, uint64_t y, uint64_t z, uint64_t a,
                                      uint64_t b)
{
    a += w;
    b = ror64(b + a + z, 21);
    uint64_t c = a;
    a += x;
    a += y;
    b += ror64(a, 44);
    u128pair r = {a + z, b + c};
    return r;
}
static u128pair weak_hash_len32(const uint8_t *s, uint64_t a, uint64_t b)
{
    return weak_hash_len32_seeds(load_le64(s), load_le64(s + 8), load_le64(s + 16), load_le64(s + 24), a, b);
}

static u128pair city_murmur(const uint8_t *s, size_t len, uint64_t seed_lo, uint64_t seed_hi)
{
    uint64_t a = seed_lo, b = seed_hi, c, d;
    if (len <= 16) {
        a = shift_mix(a * CK1) * CK1;
        c = b * CK1 + city_hash_len0to16(s, len);
        d = shift_mix(a + (len >= 8 ? load_le64(s) : c));
    } else {
        c = city_hash_len16(load_le64(s + len - 8) + CK1, a);
        d = city_hash_len16(b + len, c + load_le64(s + len - 16));
        a += d;
        int64_t l = (int64_t)len - 16;
        do {
            a ^= shift_mix(load_le64(s) * CK1) * CK1;
            a *= CK1;
            b ^= a;
            c ^= shift_mix(load_le64(s + 8) * CK1) * CK1;
            c *= CK1;
            d ^= c;
            s += 16;
            l -= 16;
        } while (l > 0);
    }
    a = city_hash_len16(a, c);
    b = city_hash_len16(d, b);
    u128pair r = {a ^ b, city_hash_len16(b, a)};
    return r;
}

static u128pair city_hash128_with_seed(const uint8_t *s, size_t len, uint64_t seed_lo, uint64_t seed_hi)
{
    if (len < 128) return city_murmur(s, len, seed_lo, seed_hi);
    u128pair v, w;
    uint64_t x = seed_lo, y = seed_hi, z = len * CK1, t;
    v.first = ror64(y ^ CK1, 49) * CK1 + load_le64(s);
    v.second = ror64(v.first, 42) * CK1 + load_le64(s + 8);
    w.first = ror64(y + z, 35) * CK1 + x;
    w.second = ror64(x + load_le64(s + 88), 53) * CK1;
    do {
        for (int half = 0; half < 2; half++) {
            x = ror64(x + y + v.first + load_le64(s + 8), 37) * CK1;
            y = ror64(y + v.second + load_le64(s + 48), 42) * CK1;
            x ^= w.second;
            y += v.first + load_le64(s + 40);
            z = ror64(z + w.first, 33) * CK1;
            v = weak_hash_len32(s, v.second * CK1, x + w.first);
            w = weak_hash_len32(s + 32, z + w.second, y + load_le64(s + 16));
            t = z; z = x; x = t;
            s += 64;
        }
        len -= 128;
    } while (len >= 128);
    x += ror64(v.first + z, 49) * CK0;
    y = y * CK0 + ror64(w.second, 37);
    z = z * CK0 + ror64(w.first, 27);
    w.first *= 9;
    v.first *= CK0;
    for (size_t tail_done = 0; tail_done < len;) {
        tail_done += 32;
        y = ror64(x + y, 42) * CK0 + v.second;
        w.first += load_le64(s + len - tail_done + 16);
        x = x * CK0 + w.first;
        z += w.second + load_le64(s + len - tail_done);
        w.second += v.first;
        v = weak_hash_len32(s + len - tail_done, v.first + z, v.second);
        v.first *= CK0;
    }
    x = city_hash_len16(x, v.first);
    y = city_hash_len16(y + z, w.first);
    u128pair r = {city_hash_len16(x + v.second, w.second) + y, city_hash_len16(x + w.second, y + v.second)};
    return r;
}

void sho_cityhash128(const uint8_t *s, size_t len, uint64_t *low, uint64_t *high)
{
    u128pair r = len >= 16 ? city_hash128_with_seed(s + 16, len - 16, load_le64(s), load_le64(s + 8) + CK0)
                           : city_hash128_with_seed(s, len, CK0, CK1);
    *low = r.first;
    *high = r.second;
}

/* CityHash64 for len <= 16 only: KAT helper that exercises HashLen0to16 */
uint64_t sho_cityhash64(const uint8_t *s, size_t len)
{
    return len <= 16 ? city_hash_len0to16(s, len) : 0;
}

/* ================================================================ ntHash */

#define NT_A 0x3c8bfbb395c60474ULL
#define NT_C 0x3193c18562a02b4cULL
#define NT_G 0x20323ed082572324ULL
#define NT_T 0x295549f54be24456ULL

/* ntHash v1 seed table: A/a, C/c, G/g, T/t, U/u; the upstream table also keeps the
 * complement look-up slots 0..7 (cpOff = 0x07), which raw bytes 1,3,4,5,7 therefore hit. */
static uint64_t nt_seed(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': case 4: case 5: return NT_A;
    case 'C': case 'c': case 7: return NT_C;
    case 'G': case 'g': case 3: return NT_G;
    case 'T': case 't': case 'U': case 'u': case 1: return NT_T;
    default: return 0;
    }
}

uint64_t sho_nthash_fwd(const uint8_t *p, size_t len)
{
    uint64_t h = 0;
    for (size_t i = 0; i < len; i++) h = rol64(h, 1) ^ nt_seed(p[i]);
    return h;
}

/* ============================================================= rapidhash */

#define RAPID_SEED 0xbdd89aa982704029ULL
static const uint64_t rapid_secret[3] = {0x2d358dccaa6c78a5ULL, 0x8bb84b93962eacc9ULL, 0x4b33a62ed433d4a3ULL};

static uint64_t rapid_mix(uint64_t a, uint64_t b) { return fold64(a, b); }

static uint64_t rapidhash_core(const uint8_t *key, size_t len, int final_secret)
{
    const uint8_t *p = key;
    const uint64_t *S = rapid_secret;
    uint64_t seed = RAPID_SEED, a, b;
    seed ^= rapid_mix(seed ^ S[0], S[1]) ^ (uint64_t)len;
    if (len <= 16) {
        if (len >= 4) {
            const uint8_t *plast = p + len - 4;
            a = ((uint64_t)load_le32(p) << 32) | load_le32(plast);
            uint64_t delta = (len & 24) >> (len >> 3);
            b = ((uint64_t)load_le32(p + delta) << 32) | load_le32(plast - delta);
        } else if (len > 0) {
            a = ((uint64_t)p[0] << 56) | ((uint64_t)p[len >> 1] << 32) | p[len - 1];
            b = 0;
        } else {
            a = b = 0;
        }
    } else {
        size_t i = len;
        if (i > 48) {
            uint64_t see1 = seed, see2 = seed;
            do {
                seed = rapid_mix(load_le64(p) ^ S[0], load_le64(p + 8) ^ seed);
                see1 = rapid_mix(load_le64(p + 16) ^ S[1], load_le64(p + 24) ^ see1);
                see2 = rapid_mix(load_le64(p + 32) ^ S[2], load_le64(p + 40) ^ see2);
                p += 48;
                i -= 48;
            } while (i >= 48);
            seed ^= see1 ^ see2;
        }
        if (i > 16) {
            seed = rapid_mix(load_le64(p) ^ S[2], load_le64(p + 8) ^ seed ^ S[1]);
            if (i > 32) seed = rapid_mix(load_le64(p + 16) ^ S[2], load_le64(p + 24) ^ seed);
        }
        a = load_le64(key + len - 16);
        b = load_le64(key + len - 8);
    }
    a ^= S[1];
    b ^= seed;
    uint64_t lo, hi;
    mul64x64(a, b, &lo, &hi);
    a = lo; b = hi;
    return rapid_mix(a ^ S[0] ^ (uint64_t)len, b ^ S[final_secret]);
}

/* the reference's port: final mix with secret[2] */
uint64_t sho_rapidhash(const uint8_t *key, size_t len) { return rapidhash_core(key, len, 2); }
/* upstream rapidhash v1 (final mix with secret[1]); used only to check the shared structure
 * against the published upstream KAT rapidhash("hello world") */
uint64_t sho_rapidhash_upstream_v1(const uint8_t *key, size_t len) { return rapidhash_core(key, len, 1); }
