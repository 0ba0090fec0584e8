/* oracle_internal.h — ORACLE (test infrastructure): shared helpers. */
#ifndef SEQHASH_ORACLE_INTERNAL_H
#define SEQHASH_ORACLE_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "oracle.h"

static inline uint32_t rol32(uint32_t x, int r) { return (x << r) | (x >> ((32 - r) & 31)); }
static inline uint32_t ror32(uint32_t x, int r) { return (x >> r) | (x << ((32 - r) & 31)); }
static inline uint64_t rol64(uint64_t x, int r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
static inline uint64_t ror64(uint64_t x, int r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }

static inline uint32_t load_le32(const uint8_t *p)
{
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}
static inline uint64_t load_le64(const uint8_t *p)
{
    return (uint64_t)load_le32(p) | (uint64_t)load_le32(p + 4) << 32;
}
static inline uint32_t load_be32(const uint8_t *p)
{
    return (uint32_t)p[3] | (uint32_t)p[2] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[0] << 24;
}
static inline void store_le32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
static inline void store_le64(uint8_t *p, uint64_t v)
{
    store_le32(p, (uint32_t)v); store_le32(p + 4, (uint32_t)(v >> 32));
}
static inline void store_be32(uint8_t *p, uint32_t v)
{
    p[3] = (uint8_t)v; p[2] = (uint8_t)(v >> 8); p[1] = (uint8_t)(v >> 16); p[0] = (uint8_t)(v >> 24);
}
static inline void store_be64(uint8_t *p, uint64_t v)
{
    store_be32(p, (uint32_t)(v >> 32)); store_be32(p + 4, (uint32_t)v);
}
/* 64x64 -> 128 multiply: *hi:*lo */
static inline void mul64x64(uint64_t a, uint64_t b, uint64_t *lo, uint64_t *hi)
{
    unsigned __int128 p = (unsigned __int128)a * b;
    *lo = (uint64_t)p;
    *hi = (uint64_t)(p >> 64);
}

/* per-algorithm one-shots (raw native output) */
void sho_sha1(const uint8_t *data, size_t len, uint8_t out[20]);
void sho_md5(const uint8_t *data, size_t len, uint8_t out[16]);
void sho_sha3_512(const uint8_t *data, size_t len, uint8_t out[64]);
void sho_kt128(const uint8_t *data, size_t len, uint8_t *out, size_t outlen);
void sho_blake3(const uint8_t *data, size_t len, uint8_t out[32]);
uint64_t sho_xxh64(const uint8_t *data, size_t len);
uint64_t sho_xxh3_64(const uint8_t *data, size_t len);
void sho_murmur3_x64_128(const uint8_t *data, size_t len, uint64_t *h1, uint64_t *h2);
void sho_cityhash128(const uint8_t *data, size_t len, uint64_t *low, uint64_t *high);
uint64_t sho_cityhash64(const uint8_t *data, size_t len); /* KAT helper only */
uint64_t sho_nthash_fwd(const uint8_t *data, size_t len);
uint64_t sho_rapidhash(const uint8_t *data, size_t len);
uint64_t sho_rapidhash_upstream_v1(const uint8_t *data, size_t len); /* KAT helper only */

#endif
