/*
 * hash_blake3.c — ORACLE (test infrastructure): BLAKE3 default hash mode, 32-byte output.
 *
 * Restates blake3.Sum256 of github.com/zeebo/blake3 v0.2.4 (go.mod:14), called at
 * seqhasher.go:328-330.  Algorithm per the BLAKE3 specification: 1024-byte chunks of
 * 64-byte blocks (7 rounds), binary tree where the left subtree holds the largest power
 * of two of chunks strictly below the total, ROOT flag on the final compression.
 */
#include "oracle_internal.h"

enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };

static const uint32_t b3_iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                  0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t b3_perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};

static void b3_g(uint32_t *s, int a, int b, int c, int d, uint32_t mx, uint32_t my)
{
    s[a] = s[a] + s[b] + mx; s[d] = ror32(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = ror32(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = ror32(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = ror32(s[b] ^ s[c], 7);
}

/* out[8] = first 8 words of the compression output (the chaining value) */
static void b3_compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                        uint32_t block_len, uint32_t flags, uint32_t out[8])
{
    uint32_t s[16], m[16], t[16];
    memcpy(s, cv, 32);
    memcpy(s + 8, b3_iv, 16);
    s[12] = (uint32_t)counter;
    s[13] = (uint32_t)(counter >> 32);
    s[14] = block_len;
    s[15] = flags;
    memcpy(m, block, 64);
    for (int r = 0; r < 7; r++) {
        b3_g(s, 0, 4, 8, 12, m[0], m[1]);
        b3_g(s, 1, 5, 9, 13, m[2], m[3]);
        b3_g(s, 2, 6, 10, 14, m[4], m[5]);
        b3_g(s, 3, 7, 11, 15, m[6], m[7]);
        b3_g(s, 0, 5, 10, 15, m[8], m[9]);
        b3_g(s, 1, 6, 11, 12, m[10], m[11]);
        b3_g(s, 2, 7, 8, 13, m[12], m[13]);
        b3_g(s, 3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[b3_perm[i]];
        memcpy(m, t, 64);
    }
    for (int i = 0; i < 8; i++) out[i] = s[i] ^ s[i + 8];
}

/* chaining value of one chunk (<= 1024 bytes); root_flag = B3_ROOT when it is the only chunk */
static void b3_chunk_cv(const uint8_t *p, size_t len, uint64_t chunk_index, uint32_t root_flag,
                        uint32_t cv[8])
{
    memcpy(cv, b3_iv, 32);
    size_t nblocks = len ? (len + 63) / 64 : 1;
    for (size_t b = 0; b < nblocks; b++) {
        size_t off = 64 * b;
        size_t bl = len - off < 64 ? len - off : 64;
        uint8_t buf[64];
        uint32_t words[16];
        memset(buf, 0, 64);
        memcpy(buf, p + off, bl);
        for (int i = 0; i < 16; i++) words[i] = load_le32(buf + 4 * i);
        uint32_t flags = 0;
        if (b == 0) flags |= B3_CHUNK_START;
        if (b == nblocks - 1) flags |= B3_CHUNK_END | root_flag;
        b3_compress(cv, words, chunk_index, (uint32_t)bl, flags, cv);
    }
}

static void b3_parent_cv(const uint32_t l[8], const uint32_t r[8], uint32_t root_flag, uint32_t cv[8])
{
    uint32_t block[16];
    memcpy(block, l, 32);
    memcpy(block + 8, r, 32);
    b3_compress(b3_iv, block, 0, 64, B3_PARENT | root_flag, cv);
}

/* non-root subtree over len > 0 bytes starting at chunk index chunk0 */
static void b3_subtree_cv(const uint8_t *p, size_t len, uint64_t chunk0, uint32_t cv[8])
{
    if (len <= 1024) {
        b3_chunk_cv(p, len, chunk0, 0, cv);
        return;
    }
    size_t full_chunks = (len - 1) / 1024;
    size_t pow2 = 1;
    while (pow2 * 2 <= full_chunks) pow2 *= 2;
    size_t left_len = pow2 * 1024;
    uint32_t l[8], r[8];
    b3_subtree_cv(p, left_len, chunk0, l);
    b3_subtree_cv(p + left_len, len - left_len, chunk0 + pow2, r);
    b3_parent_cv(l, r, 0, cv);
}

void sho_blake3(const uint8_t *data, size_t len, uint8_t out[32])
{
    uint32_t cv[8];
    if (len <= 1024) {
        b3_chunk_cv(data, len, 0, B3_ROOT, cv);
    } else {
        size_t full_chunks = (len - 1) / 1024;
        size_t pow2 = 1;
        while (pow2 * 2 <= full_chunks) pow2 *= 2;
        size_t left_len = pow2 * 1024;
        uint32_t l[8], r[8];
        b3_subtree_cv(data, left_len, 0, l);
        b3_subtree_cv(data + left_len, len - left_len, pow2, r);
        b3_parent_cv(l, r, B3_ROOT, cv);
    }
    for (int i = 0; i < 8; i++) store_le32(out + 4 * i, cv[i]);
}
