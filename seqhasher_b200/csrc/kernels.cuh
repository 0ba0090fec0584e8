// kernels.cuh — host-callable launchers of the sm_100a kernels (internal to libseqhash_b200).
#pragma once
#include <cstddef>
#include <cstdint>
#include <atomic>
#include <cuda_runtime.h>

namespace shb {

constexpr int N_ALGOS = 12;
enum Algo : int {
    A_SHA1 = 0, A_SHA3 = 1, A_MD5 = 2, A_XXHASH = 3, A_XXH3 = 4, A_CITYHASH = 5, A_MURMUR3 = 6,
    A_NTHASH = 7, A_BLAKE3 = 8, A_K12 = 9, A_RAPIDHASH = 10, A_RAPIDHASH32 = 11
};
__host__ __device__ constexpr int digest_size(int a)
{
    constexpr int S[N_ALGOS] = {20, 64, 16, 8, 8, 16, 16, 8, 32, 128, 8, 4};
    return S[a];
}

extern std::atomic<unsigned long long> g_launches;   // kernels launched by this library (host counter)

// Byte blocks of the normalise pre-pass.
constexpr uint32_t NORM_BLK = 4096;
inline size_t norm_blocks(size_t total_bytes) { return total_bytes / NORM_BLK + 1; }

// Direct path: one thread per record straight from global memory.
// perm (may be null) = the order in which the records are visited, from launch_length_sort; DIRECT_LOCAL_ORDER =
// every CTA sorts windows of 512 consecutive records by length itself (long records: keeps a warp's reads close).
static const uint32_t *const DIRECT_LOCAL_ORDER = reinterpret_cast<const uint32_t *>(uintptr_t(1));
cudaError_t launch_hash_direct(int algo, bool upper, const uint8_t *res, const int64_t *offs, size_t n,
                               uint8_t *out, const uint32_t *perm, cudaStream_t st);
size_t length_sort_bytes(size_t n);
cudaError_t launch_length_sort(const int64_t *offs, size_t n, void *perm_and_tmp, uint32_t identity_below_q8, cudaStream_t st);

// Tile path (tile.cu): CTA-staged shared-memory tiles of 128 records, one thread per record.  `mask` has
// bit a set for every algorithm to compute; more than one bit is allowed only inside tile_fast_mask()
// (the HBM-bound hashes, fused so the residues are read once).  outs[a] = digest array of algo a.
// cap = shared-memory tile budget in bytes; tiles that exceed it read global memory directly.
// strip_scratch != nullptr fuses the whitespace strip into the tile pass (needs total_bytes + 32 bytes of
// scratch for the rare oversize tile); out_len (may be null) then receives the post-strip lengths.
cudaError_t launch_hash_tile(uint32_t mask, bool upper, const uint8_t *res, const int64_t *offs, size_t n,
                             uint8_t *const *outs, uint32_t cap, uint8_t *strip_scratch, int64_t *out_len,
                             cudaStream_t st);
uint32_t tile_fast_mask();
unsigned long long tile_debug_oob();   // -DSHB_BOUNDS builds: out-of-window shared-memory accesses of the tile path so far
bool tile_debug_bounds_build();

// Long-record path (long.cu): ntHash warp-per-record; BLAKE3 thread-per-chunk over the flattened chunk
// list + thread-per-record tree merge (scratch: b3_long_scratch_bytes()).
cudaError_t launch_nthash_long(bool upper, const uint8_t *res, const int64_t *offs, size_t n, uint8_t *out,
                               cudaStream_t st);
cudaError_t launch_xxh3_long(bool upper, const uint8_t *res, const int64_t *offs, size_t n, uint8_t *out, cudaStream_t st);
size_t b3_long_scratch_bytes(size_t total_bytes, size_t n);
// K12 leaf-parallel: thread per 8 KiB leaf over the flattened leaf list + thread per record final node
// (scratch: same layout/size as the BLAKE3 long path)
cudaError_t launch_k12_long(bool upper, const uint8_t *res, const int64_t *offs, size_t n, size_t total_bytes,
                            uint8_t *out, void *scratch, cudaStream_t st);
// nt_out != nullptr: the same pass also produces the ntHash digests (8 bytes per record) — `--hash blake3,nthash` reads the
// residues once
cudaError_t launch_blake3_long(bool upper, const uint8_t *res, const int64_t *offs, size_t n, size_t total_bytes,
                               uint8_t *out, void *scratch, uint8_t *nt_out, cudaStream_t st);

// Normalise pre-pass (seqhasher.go:246-251 materialised): whitespace strip (if strip) + upper-casing
// (if upper) of the whole CSR batch into out_res / out_offs (n+1) / out_len (n, may be null).
// block_counts: norm_blocks() uint32; block_prefix: norm_blocks()+1 int64.
cudaError_t launch_normalise(const uint8_t *res, const int64_t *offs, size_t n, size_t total_bytes, bool strip,
                             bool upper, uint32_t *block_counts, int64_t *block_prefix, uint8_t *out_res,
                             int64_t *out_offs, int64_t *out_len, void *scan_tmp, cudaStream_t st);

// Exclusive scan of n uint32 counts into n+1 int64 prefixes; tmp = scan_tmp_bytes(n) caller-owned bytes
// (nullptr: single-CTA scan, fine for small n).
size_t scan_tmp_bytes(size_t n);
cudaError_t scan_counts(const uint32_t *counts, size_t n, int64_t *prefix, void *tmp, cudaStream_t st);

// bytes >= 0x80 in res[0, total_bytes) into *count (device memory)
cudaError_t launch_count_nonascii(const uint8_t *res, size_t total_bytes, unsigned long long *count, cudaStream_t st);

// rapidhash32 digests (4 bytes each) from finished rapidhash digests (8 bytes each): the record is hashed once for both
cudaError_t launch_rapid_fold32(const uint8_t *h64, size_t n, uint8_t *h32, cudaStream_t st);

cudaError_t launch_lengths(const int64_t *offs, size_t n, int64_t *out_len, cudaStream_t st);

cudaError_t launch_synth_fill(uint8_t *res, size_t total_bytes, uint64_t seed, unsigned lower_permille,
                              unsigned n_permille, cudaStream_t st);

// returns lane-instructions executed; elapsed time measured by the caller
cudaError_t launch_probe(int kind, int iters, uint32_t *sink, double *lane_instr, cudaStream_t st);

}  // namespace shb
