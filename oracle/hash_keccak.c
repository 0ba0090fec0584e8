/*
 * hash_keccak.c — ORACLE (test infrastructure): SHA3-512 and KangarooTwelve (KT128).
 *
 * sha3: restates sha3.Sum512 of golang.org/x/crypto v0.52.0 (go.mod:16), called at
 *       seqhasher.go:302-304.  FIPS 202: Keccak-f[1600], 24 rounds, rate 72 B,
 *       domain suffix 0x06, 64-byte output.
 * k12 : restates k12.NewDraft10(nil) / Write / Read(128) of github.com/cloudflare/circl
 *       v1.6.3 (go.mod:7), called at seqhasher.go:331-336.  Draft-10 == RFC 9861
 *       KT128 with an empty customization string, 128 output bytes.
 */
#include "oracle_internal.h"

static const uint64_t keccak_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

/* rho offsets r[x][y] for lane A[x][y] stored at index x + 5*y */
static const int keccak_rho[5][5] = {
    {0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};

/* Keccak-p[1600, nrounds]: the LAST nrounds of the 24 round constants */
static void keccak_p(uint64_t a[25], int nrounds)
{
    for (int round = 24 - nrounds; round < 24; round++) {
        uint64_t c[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) {
            uint64_t d = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
            for (int y = 0; y < 5; y++) a[x + 5 * y] ^= d;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++)
                b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], keccak_rho[x][y]);
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++)
                a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= keccak_rc[round];
    }
}

/* generic sponge: absorb data (rate bytes per block), pad with suffix...0x80, squeeze */
typedef struct {
    uint64_t a[25];
    uint8_t buf[200];
    size_t fill, rate;
    int rounds;
} sponge;

static void sponge_init(sponge *s, size_t rate, int rounds)
{
    memset(s, 0, sizeof *s);
    s->rate = rate;
    s->rounds = rounds;
}
static void sponge_xor_block(sponge *s, const uint8_t *p)
{
    for (size_t i = 0; i < s->rate / 8; i++) s->a[i] ^= load_le64(p + 8 * i);
    keccak_p(s->a, s->rounds);
}
static void sponge_absorb(sponge *s, const uint8_t *p, size_t n)
{
    while (n) {
        if (s->fill == 0 && n >= s->rate) {
            sponge_xor_block(s, p);
            p += s->rate; n -= s->rate;
            continue;
        }
        size_t take = s->rate - s->fill;
        if (take > n) take = n;
        memcpy(s->buf + s->fill, p, take);
        s->fill += take; p += take; n -= take;
        if (s->fill == s->rate) { sponge_xor_block(s, s->buf); s->fill = 0; }
    }
}
static void sponge_finish(sponge *s, uint8_t suffix, uint8_t *out, size_t outlen)
{
    memset(s->buf + s->fill, 0, s->rate - s->fill);
    s->buf[s->fill] ^= suffix;
    s->buf[s->rate - 1] ^= 0x80;
    sponge_xor_block(s, s->buf);
    for (;;) {
        uint8_t blk[200];
        for (size_t i = 0; i < s->rate / 8; i++) store_le64(blk + 8 * i, s->a[i]);
        size_t take = outlen < s->rate ? outlen : s->rate;
        memcpy(out, blk, take);
        out += take; outlen -= take;
        if (!outlen) break;
        keccak_p(s->a, s->rounds);
    }
}

void sho_sha3_512(const uint8_t *data, size_t len, uint8_t out[64])
{
    sponge s;
    sponge_init(&s, 72, 24);
    sponge_absorb(&s, data, len);
    sponge_finish(&s, 0x06, out, 64);
}

/* TurboSHAKE128: Keccak-p[1600,12], rate 168, domain byte D */
static void turboshake128(const uint8_t *p1, size_t n1, const uint8_t *p2, size_t n2, uint8_t d,
                          uint8_t *out, size_t outlen)
{
    sponge s;
    sponge_init(&s, 168, 12);
    sponge_absorb(&s, p1, n1);
    if (n2) sponge_absorb(&s, p2, n2);
    sponge_finish(&s, d, out, outlen);
}

/* RFC 9861 length_encode: minimal big-endian bytes of x, then their count */
static size_t length_encode(uint64_t x, uint8_t *out)
{
    size_t n = 0;
    uint8_t tmp[8];
    while (x) { tmp[n++] = (uint8_t)x; x >>= 8; }
    for (size_t i = 0; i < n; i++) out[i] = tmp[n - 1 - i];
    out[n] = (uint8_t)n;
    return n + 1;
}

/* KT128(M, C = ""): S = M || C || length_encode(|C|) = M || 0x00 */
void sho_kt128(const uint8_t *data, size_t len, uint8_t *out, size_t outlen)
{
    static const uint8_t zero = 0x00;
    const size_t B = 8192;
    size_t slen = len + 1;
    if (slen <= B) {
        turboshake128(data, len, &zero, 1, 0x07, out, outlen);
        return;
    }
    /* chunks S_0 .. S_{n-1} of S; the trailing 0x00 belongs to the last chunk */
    size_t n = (slen + B - 1) / B;
    sponge fin;
    sponge_init(&fin, 168, 12);
    sponge_absorb(&fin, data, B); /* S_0 is entirely message bytes because slen > B */
    static const uint8_t marker[8] = {0x03, 0, 0, 0, 0, 0, 0, 0};
    sponge_absorb(&fin, marker, 8);
    for (size_t i = 1; i < n; i++) {
        size_t off = i * B;
        size_t clen = (i == n - 1) ? slen - off : B;
        uint8_t cv[32];
        if (i == n - 1) /* last chunk ends with the 0x00 of S */
            turboshake128(data + off, clen - 1, &zero, 1, 0x0B, cv, 32);
        else
            turboshake128(data + off, clen, NULL, 0, 0x0B, cv, 32);
        sponge_absorb(&fin, cv, 32);
    }
    uint8_t enc[16];
    size_t el = length_encode((uint64_t)(n - 1), enc);
    enc[el++] = 0xFF;
    enc[el++] = 0xFF;
    sponge_absorb(&fin, enc, el);
    sponge_finish(&fin, 0x06, out, outlen);
}
