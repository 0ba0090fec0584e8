// hash_all.cuh — one record -> one digest for each of the 12 algorithms (getHashFunc's switch,
// seqhasher.go:297-347), written in "hex order" (see include/seqhash_b200.h).
#pragma once
#include "hash_blake3.cuh"
#include "hash_fast.cuh"
#include "hash_keccak.cuh"
#include "hash_md.cuh"
#include "kernels.cuh"

namespace shb {

template <int ALGO, class R>
__device__ __forceinline__ void hash_one(const R &r, uint32_t len, uint8_t *out)
{
    if (ALGO == A_SHA1) sha1_record(r, len, out);
    else if (ALGO == A_MD5) md5_record(r, len, out);
    else if (ALGO == A_SHA3) sha3_512_record(r, len, out);
    else if (ALGO == A_K12) kt128_record(r, len, out);
    else if (ALGO == A_BLAKE3) blake3_record(r, len, out);
    else if (ALGO == A_XXHASH) store_be64(out, xxh64_record(r, len));
    else if (ALGO == A_XXH3) store_be64(out, xxh3_record(r, len));
    else if (ALGO == A_NTHASH) store_be64(out, nthash_record(r, len));
    else if (ALGO == A_RAPIDHASH) store_be64(out, rapidhash_record(r, len));
    else if (ALGO == A_RAPIDHASH32) {
        uint64_t h = rapidhash_record(r, len);
        *reinterpret_cast<uint32_t *>(out) = bswap32((uint32_t)h ^ (uint32_t)(h >> 32));
    } else if (ALGO == A_MURMUR3) {
        uint64_t h1, h2;
        murmur3_record(r, len, h1, h2);
        store_be64(out, h1);
        store_be64(out + 8, h2);
    } else if (ALGO == A_CITYHASH) {
        CityPair p = city128_record(r, len);
        store_be64(out, p.second);   // High
        store_be64(out + 8, p.first);   // Low
    }
}

template <int ALGO>
__device__ __forceinline__ void zero_digest(uint8_t *out)
{
    uint32_t *o = reinterpret_cast<uint32_t *>(out);
#pragma unroll
    for (int i = 0; i < digest_size(ALGO) / 4; i++) o[i] = 0;
}


}  // namespace shb
