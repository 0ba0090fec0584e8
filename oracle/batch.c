/*
 * batch.c — ORACLE (test infrastructure): dispatch, normalisation, hex and the batch loop.
 *
 * Restates seqhasher.go:241-259 (per-record strip -> upper -> hash loop) and
 * seqhasher.go:290-349 (getHashFunc: name dispatch, empty-sequence rule, output formats).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdlib.h>

#include "oracle_internal.h"

static const char *const algo_names[SHO_N_ALGOS] = {"sha1", "sha3", "md5", "xxhash", "xxh3", "cityhash",
                                                    "murmur3", "nthash", "blake3", "k12", "rapidhash",
                                                    "rapidhash32"};
static const size_t algo_sizes[SHO_N_ALGOS] = {20, 64, 16, 8, 8, 16, 16, 8, 32, 128, 8, 4};

int sho_algo_id(const char *name)
{
    for (int i = 0; i < SHO_N_ALGOS; i++)
        if (strcmp(name, algo_names[i]) == 0) return i;
    return -1;
}

size_t sho_digest_size(int algo) { return (algo >= 0 && algo < SHO_N_ALGOS) ? algo_sizes[algo] : 0; }

/* ---- optional tuned back-ends for the CPU-baseline timing (same results) ---- */

typedef unsigned char *(*ossl_oneshot_fn)(const unsigned char *, size_t, unsigned char *);
/* low-level (deprecated but present) init/update/final: unlike the one-shot SHA1()/MD5() of OpenSSL 3 they
 * do not fetch the algorithm under a global lock on every call, so they scale across threads */
typedef int (*ossl_init_fn)(void *);
typedef int (*ossl_update_fn)(void *, const void *, size_t);
typedef int (*ossl_final_fn)(unsigned char *, void *);
static ossl_init_fn sha1_init, md5_init;
static ossl_update_fn sha1_update, md5_update;
static ossl_final_fn sha1_final, md5_final;
typedef int (*evp_q_digest_fn)(void *, const char *, const char *, const void *, size_t, unsigned char *,
                               size_t *);
typedef uint64_t (*xxh64_fn)(const void *, size_t, uint64_t);
typedef uint64_t (*xxh3_fn)(const void *, size_t);

static ossl_oneshot_fn ossl_sha1, ossl_md5;
static evp_q_digest_fn ossl_q_digest;
static xxh64_fn lib_xxh64;
static xxh3_fn lib_xxh3;
static int accel_state = -1;

static void accel_load(void)
{
    accel_state = 0;
    void *c = dlopen("libcrypto.so.3", RTLD_NOW | RTLD_LOCAL);
    if (c) {
        ossl_sha1 = (ossl_oneshot_fn)dlsym(c, "SHA1");
        ossl_md5 = (ossl_oneshot_fn)dlsym(c, "MD5");
        ossl_q_digest = (evp_q_digest_fn)dlsym(c, "EVP_Q_digest");
        sha1_init = (ossl_init_fn)dlsym(c, "SHA1_Init");
        sha1_update = (ossl_update_fn)dlsym(c, "SHA1_Update");
        sha1_final = (ossl_final_fn)dlsym(c, "SHA1_Final");
        md5_init = (ossl_init_fn)dlsym(c, "MD5_Init");
        md5_update = (ossl_update_fn)dlsym(c, "MD5_Update");
        md5_final = (ossl_final_fn)dlsym(c, "MD5_Final");
        if (ossl_sha1 && ossl_md5 && ossl_q_digest && sha1_init && sha1_update && sha1_final && md5_init && md5_update &&
            md5_final)
            accel_state |= 1;
    }
    void *x = dlopen("libxxhash.so.0", RTLD_NOW | RTLD_LOCAL);
    if (x) {
        lib_xxh64 = (xxh64_fn)dlsym(x, "XXH64");
        lib_xxh3 = (xxh3_fn)dlsym(x, "XXH3_64bits");
        if (lib_xxh64 && lib_xxh3) accel_state |= 2;
    }
}

int sho_accel_available(void)
{
    if (accel_state < 0) accel_load();
    return accel_state;
}

/* ---- one-shot dispatch: seqhasher.go:297-347 ---- */

static size_t hash_dispatch(int algo, const uint8_t *data, size_t len, uint8_t *out, int accel)
{
    if (len == 0) return 0; /* seqhasher.go:292-295 */
    uint64_t a, b;
    switch (algo) {
    case SHO_SHA1:
        if (accel & 1) {
            uint64_t ctx[32];   /* SHA_CTX is 96 bytes */
            sha1_init(ctx); sha1_update(ctx, data, len); sha1_final(out, ctx);
        } else {
            sho_sha1(data, len, out);
        }
        return 20;
    case SHO_SHA3:
        if (accel & 1) { size_t ol = 64; ossl_q_digest(NULL, "SHA3-512", NULL, data, len, out, &ol); }
        else sho_sha3_512(data, len, out);
        return 64;
    case SHO_MD5:
        if (accel & 1) {
            uint64_t ctx[32];   /* MD5_CTX is 92 bytes */
            md5_init(ctx); md5_update(ctx, data, len); md5_final(out, ctx);
        } else {
            sho_md5(data, len, out);
        }
        return 16;
    case SHO_XXHASH: /* fmt "%016x" */
        store_be64(out, (accel & 2) ? lib_xxh64(data, len, 0) : sho_xxh64(data, len));
        return 8;
    case SHO_XXH3:
        store_be64(out, (accel & 2) ? lib_xxh3(data, len) : sho_xxh3_64(data, len));
        return 8;
    case SHO_CITYHASH: /* "%016x%016x", High then Low: seqhasher.go:316 */
        sho_cityhash128(data, len, &a, &b);
        store_be64(out, b);
        store_be64(out + 8, a);
        return 16;
    case SHO_MURMUR3: /* "%016x%016x" h1,h2: seqhasher.go:319 */
        sho_murmur3_x64_128(data, len, &a, &b);
        store_be64(out, a);
        store_be64(out + 8, b);
        return 16;
    case SHO_NTHASH:
        /* nthash.NewHasher errors when k > math.MaxUint32 => "" (seqhasher.go:322-325) */
        if (len > 0xFFFFFFFFULL) return 0;
        store_be64(out, sho_nthash_fwd(data, len));
        return 8;
    case SHO_BLAKE3:
        sho_blake3(data, len, out);
        return 32;
    case SHO_K12:
        sho_kt128(data, len, out, 128);
        return 128;
    case SHO_RAPIDHASH:
        store_be64(out, sho_rapidhash(data, len));
        return 8;
    case SHO_RAPIDHASH32: { /* seqhasher.go:341-343 */
        uint64_t h = sho_rapidhash(data, len);
        store_be32(out, (uint32_t)h ^ (uint32_t)(h >> 32));
        return 4;
    }
    default:
        return 0;
    }
}

size_t sho_hash(int algo, const uint8_t *data, size_t len, uint8_t *out)
{
    return hash_dispatch(algo, data, len, out, 0);
}

/* ---- seqhasher.go:246-251, ASCII semantics ---- */

size_t sho_normalise(const uint8_t *in, size_t len, int flags, uint8_t *out)
{
    size_t j = 0;
    for (size_t i = 0; i < len; i++) {
        uint8_t c = in[i];
        if ((flags & SHO_FLAG_STRIP) && (c == ' ' || (c >= '\t' && c <= '\r'))) continue;
        if ((flags & SHO_FLAG_UPPER) && c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
        out[j++] = c;
    }
    return j;
}

size_t sho_hash_hex(const char *hash_type, const uint8_t *data, size_t len, char *out_hex)
{
    static const char digits[] = "0123456789abcdef";
    int algo = sho_algo_id(hash_type);
    if (algo < 0) algo = SHO_SHA1; /* default: arm, seqhasher.go:344-346 */
    uint8_t raw[128];
    size_t n = sho_hash(algo, data, len, raw);
    for (size_t i = 0; i < n; i++) {
        out_hex[2 * i] = digits[raw[i] >> 4];
        out_hex[2 * i + 1] = digits[raw[i] & 15];
    }
    out_hex[2 * n] = 0;
    return 2 * n;
}

/* ---- seqhasher.go:241-259 over a CSR batch ---- */

int sho_batch(const uint8_t *residues, const int64_t *offsets, size_t n, const int *algos, int n_algos,
              int flags, uint8_t **digests, int64_t *out_lengths, int threads, int accel)
{
    for (int k = 0; k < n_algos; k++)
        if (algos[k] < 0 || algos[k] >= SHO_N_ALGOS) return -1;
    if (accel) accel = sho_accel_available();
    if (threads < 1) threads = 1;
    int bad = 0;
#pragma omp parallel num_threads(threads)
    {
        size_t cap = 1 << 16;
        uint8_t *scratch = (uint8_t *)malloc(cap);
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)n; i++) {
            const uint8_t *p = residues + offsets[i];
            size_t len = (size_t)(offsets[i + 1] - offsets[i]);
            if (flags & (SHO_FLAG_STRIP | SHO_FLAG_UPPER)) {
                if (len > cap) {
                    cap = len * 2;
                    free(scratch);
                    scratch = (uint8_t *)malloc(cap);
                }
                if (!scratch) { bad = 1; continue; }
                len = sho_normalise(p, len, flags, scratch);
                p = scratch;
            }
            if (out_lengths) out_lengths[i] = (int64_t)len;
            for (int k = 0; k < n_algos; k++) {
                size_t ds = algo_sizes[algos[k]];
                uint8_t *dst = digests[k] + (size_t)i * ds;
                if (hash_dispatch(algos[k], p, len, dst, accel) == 0) memset(dst, 0, ds);
            }
        }
        free(scratch);
    }
    return bad ? -2 : 0;
}
