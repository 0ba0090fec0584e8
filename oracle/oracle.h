/*
 * oracle.h — CPU ORACLE for the per-record hash path of vmikk/seqhasher.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm for the hot path (seqhasher.go:241-259 strip -> upper -> hashes, and
 * seqhasher.go:290-349 getHashFunc).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (seqhasher_b200/) never links, imports or calls anything in this directory.
 *
 * Why a restatement: the reference is Go, 10 of its 12 hash algorithms live in
 * un-vendored third-party Go modules (go.mod:5-17) and there is no Go toolchain in
 * this image, so the reference itself cannot be compiled here (oracle/_ref is
 * therefore absent).  Each algorithm below names the module + pinned version whose
 * published algorithm it restates and the reference call site.
 *
 * Pinning status (see tests/test_oracle_golden.py):
 *   sha1 md5 sha3 xxhash xxh3 blake3  : pinned — reference goldens
 *        (seqhasher_test.go:318-329, test/test2_binary.sh) + independent library
 *        (hashlib / python-xxhash / python-blake3) on every boundary length.
 *   k12      : pinned — reference golden + RFC 9861 KATs (single and multi chunk).
 *   murmur3  : pinned — reference golden + upstream MurmurHash3_x64_128 KATs.
 *   nthash   : pinned for A/C/G/T/other-uppercase (2 goldens + closed form);
 *              lower-case / U under --casesensitive: PARITY UNPINNED (table recalled
 *              from ntHash v1 upstream).
 *   cityhash : golden (4 B) + CityHash128("abc") KAT; the >=16 B paths follow
 *              upstream CityHash v1.1 — PARITY UNPINNED by any reference vector.
 *   rapidhash/rapidhash32 : golden (4 B) only; other lengths follow rapidhash v1
 *              with the port's final mix — PARITY UNPINNED.
 */
#ifndef SEQHASH_ORACLE_H
#define SEQHASH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* algorithm ids: order of supportedHashTypes, seqhasher.go:41-43 */
enum {
    SHO_SHA1 = 0, SHO_SHA3 = 1, SHO_MD5 = 2, SHO_XXHASH = 3, SHO_XXH3 = 4, SHO_CITYHASH = 5,
    SHO_MURMUR3 = 6, SHO_NTHASH = 7, SHO_BLAKE3 = 8, SHO_K12 = 9, SHO_RAPIDHASH = 10,
    SHO_RAPIDHASH32 = 11, SHO_N_ALGOS = 12
};

/* normalisation flags, same bit meaning as include/seqhash_b200.h */
#define SHO_FLAG_UPPER 1 /* !--casesensitive, seqhasher.go:249-251 */
#define SHO_FLAG_STRIP 2 /* bytes.Fields join, seqhasher.go:246 */

int sho_algo_id(const char *name);   /* exact name -> id, -1 if not one of the 12 */
size_t sho_digest_size(int algo);    /* 20,64,16,8,8,16,16,8,32,128,8,4 */

/* One-shot raw digest.  Output bytes are laid out so that lowercase hex of the bytes
 * equals the reference string (word hashes big-endian; cityhash High||Low; murmur3
 * h1||h2).  Returns the digest size, or 0 when len == 0 (empty-sequence rule,
 * seqhasher.go:292-295: the reference returns ""). */
size_t sho_hash(int algo, const uint8_t *data, size_t len, uint8_t *out);

/* seqhasher.go:246-251 for ASCII input: remove \t \n \v \f \r ' ' (if STRIP), map a-z
 * to A-Z (if UPPER).  Bytes >= 0x80 are copied through (the reference's README.md:259
 * declares non-ASCII input unsupported).  out may alias in.  Returns the new length. */
size_t sho_normalise(const uint8_t *in, size_t len, int flags, uint8_t *out);

/* getHashFunc(hashType)(data) as a C call (seqhasher.go:290-349): lowercase hex into
 * out_hex (NUL terminated, needs >= 257 bytes), "" for empty data, and ANY name that is
 * not one of the 12 (including untrimmed ones such as " md5") hashes as SHA-1, exactly
 * like the reference's `default:` arm.  Returns strlen(out_hex). */
size_t sho_hash_hex(const char *hash_type, const uint8_t *data, size_t len, char *out_hex);

/* Batch form of seqhasher.go:241-259 over a CSR batch: record i is
 * residues[offsets[i] .. offsets[i+1]).  For each record: normalise(flags), then every
 * algo in algos[].  digests[k] receives n * sho_digest_size(algos[k]) bytes (all-zero
 * slot for an empty record); out_lengths[i] (may be NULL) receives the post-normalise
 * length.  threads <= 1 runs the reference's sequential loop; > 1 shards records over
 * OpenMP threads.  accel != 0 lets sha1/md5/sha3 use OpenSSL EVP and xxhash/xxh3 use
 * libxxhash when those libraries can be dlopen'ed (CPU-baseline timing only; results
 * are identical).  Returns 0, or -1 on a bad algo id. */
int sho_batch(const uint8_t *residues, const int64_t *offsets, size_t n, const int *algos,
              int n_algos, int flags, uint8_t **digests, int64_t *out_lengths, int threads,
              int accel);

/* 1 if the accelerated back-ends could be loaded (bit0 OpenSSL, bit1 libxxhash). */
int sho_accel_available(void);

#ifdef __cplusplus
}
#endif
#endif
