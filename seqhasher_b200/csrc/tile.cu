// tile.cu — the short-record path: one CTA stages a contiguous run of records (a "tile") from the
// CSR residue buffer into shared memory with one 1-D TMA bulk copy (cp.async.bulk + mbarrier), then
// every thread hashes one record out of shared memory.
//
// Why: 150-250-byte records at arbitrary byte offsets waste most of every 32-byte sector when each
// thread reads its own record from global memory; the tile is read from HBM exactly once, fully
// coalesced, and the per-thread unaligned reads hit shared memory instead.  Several CTAs are resident
// per SM so one CTA's TMA load overlaps its neighbours' hashing.
//
// A kernel instance handles a compile-time SET of algorithms (MASK): with more than one, the tile is
// upper-cased in place once and every requested hash reads the same staged bytes (the reference
// re-reads the record per hash, seqhasher.go:255-259).  Tiles that do not fit the shared-memory
// budget fall back, per CTA, to direct global reads, so any input is still handled.
#include "hash_all.cuh"
#include "hash_fused3.cuh"

namespace shb {

struct OutPtrs {
    uint8_t *p[N_ALGOS];
};

template <uint32_t MASK, int ALGO, bool NT_SWZ, class R>
__device__ __forceinline__ void hash_if(const R &r, uint32_t len, size_t i, const OutPtrs &out, uint32_t rmask,
                                        const uint64_t *nt_lut)
{
    if ((MASK >> ALGO) & 1u) {
        if ((rmask >> ALGO) & 1u) {
            uint8_t *o = out.p[ALGO] + i * digest_size(ALGO);
            if (len == 0) zero_digest<ALGO>(o);   // empty-sequence rule, seqhasher.go:292-295
            else if (ALGO == A_NTHASH && nt_lut) store_be64(o, nthash_record_lut<NT_SWZ>(r, len, nt_lut));   // shared-memory seeds
            else hash_one<ALGO>(r, len, o);
        }
    }
}

// nt_lut: the CTA's pre-rotated ntHash seed table in shared memory (null on paths that did not build one)
template <uint32_t MASK, bool NT_SWZ = false, class R>
__device__ __forceinline__ void hash_set(const R &r, uint32_t len, size_t i, const OutPtrs &out, uint32_t rmask,
                                         const uint64_t *nt_lut = nullptr)
{
    hash_if<MASK, A_SHA1, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_SHA3, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_MD5, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_XXHASH, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_XXH3, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_CITYHASH, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_MURMUR3, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_NTHASH, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_BLAKE3, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    hash_if<MASK, A_K12, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    // rapidhash32 is a fold of rapidhash (seqhasher.go:340-343): when both are asked for the record is hashed once
    constexpr uint32_t RAPID_BOTH = (1u << A_RAPIDHASH) | (1u << A_RAPIDHASH32);
    if ((MASK & RAPID_BOTH) == RAPID_BOTH && (rmask & RAPID_BOTH) == RAPID_BOTH) {
        uint8_t *o64 = out.p[A_RAPIDHASH] + i * 8, *o32 = out.p[A_RAPIDHASH32] + i * 4;
        if (len == 0) {
            zero_digest<A_RAPIDHASH>(o64);
            zero_digest<A_RAPIDHASH32>(o32);
        } else {
            const uint64_t h = rapidhash_record(r, len);
            store_be64(o64, h);
            *reinterpret_cast<uint32_t *>(o32) = bswap32((uint32_t)h ^ (uint32_t)(h >> 32));
        }
    } else {
        hash_if<MASK, A_RAPIDHASH, NT_SWZ>(r, len, i, out, rmask, nt_lut);
        hash_if<MASK, A_RAPIDHASH32, NT_SWZ>(r, len, i, out, rmask, nt_lut);
    }
}

constexpr uint32_t TMA_CHUNK = 32768;   // one bulk copy moves at most this many bytes

// ---- whitespace strip fused into the tile pass (seqhasher.go:246) ----
// non-zero iff some byte of w is < 0x21 (every ASCII whitespace byte is): exact as an ANY test
__device__ __forceinline__ uint32_t has_byte_below_0x21(uint32_t w) { return (w - 0x21212121u) & ~w & 0x80808080u; }

// compact the record [start, start+len) of the tile in place; returns the new length
__device__ __noinline__ uint32_t strip_in_smem(uint8_t *tile, uint32_t start, uint32_t len)
{
    uint32_t j = start;
    for (uint32_t i = start; i < start + len; i++) {
        smem_check(tile + i, 1);
        const uint8_t c = tile[i];
        if (!is_ws(c)) tile[j++] = c;
    }
    return j - start;
}

// same for a record that is not staged: compacted copy into the scratch buffer at the same offset
__device__ __forceinline__ uint32_t strip_to_scratch(const uint8_t *__restrict__ res, uint8_t *__restrict__ scratch,
                                                     int64_t s, uint32_t len)
{
    uint32_t j = 0;
    for (uint32_t i = 0; i < len; i++) {
        const uint8_t c = res[s + i];
        if (!is_ws(c)) scratch[s + j++] = c;
    }
    return j;
}

constexpr uint32_t bit(int a) { return 1u << a; }
constexpr uint32_t FAST_MASK = bit(A_XXHASH) | bit(A_XXH3) | bit(A_CITYHASH) | bit(A_MURMUR3) | bit(A_NTHASH) |
                               bit(A_RAPIDHASH) | bit(A_RAPIDHASH32);
constexpr uint32_t C3_MASK = bit(A_XXHASH) | bit(A_XXH3) | bit(A_MURMUR3);

// STRIP fuses the whitespace strip (seqhasher.go:246) into the tile pass without a second look at the data:
// the readers OR a "byte < 0x21" test into a flag while the record is being hashed (for a fused set with
// upper-casing the flag comes from the in-place upper pass instead); only a record that really holds such a
// byte is compacted in shared memory and hashed again.  out_len (may be null) receives the post-strip lengths.
template <uint32_t MASK, bool UPPER, bool STRIP>
__global__ void __launch_bounds__(128) hash_tile_kernel(const uint8_t *__restrict__ res,
                                                        const int64_t *__restrict__ offs, size_t n, OutPtrs out,
                                                        uint32_t rmask, uint32_t cap, uint8_t *strip_scratch,
                                                        int64_t *__restrict__ out_len)
{
    constexpr bool MULTI = (MASK & (MASK - 1)) != 0;
    // xxhash + xxh3 + murmur3 (config C3): one fused pass per record that loads, aligns and upper-cases every 16 bytes
    // once in registers for the three hashes (hash_fused3.cuh), so no in-place upper-casing pass either
    constexpr bool FUSED3 = MASK == C3_MASK;
    constexpr bool INPLACE = MULTI && UPPER && !FUSED3;   // tile upper-cased in place by a cooperative pass
    constexpr bool SPECULATE = STRIP && !INPLACE;     // whitespace candidates detected by the readers
    // The CTA hands its 128 records to the threads longest first, so that a warp's 32 records finish together
    // (amplicon reads of 100-600 bp otherwise run every warp for its longest record: x1.3-1.6 for the INT-bound
    // hashes, x1.1-1.5 for the HBM-bound ones; fixed-length reads pay one shuffle and the barrier's vote, <= 2 %).
    constexpr bool REORDER = true;
    // input preparation (upper-casing, whitespace test) on the FMA pipe — everywhere except SHA-1 with the detector
    constexpr bool FMA_PREP = !(SPECULATE && ((MASK >> A_SHA1) & 1u)) && !((MASK >> A_K12) & 1u);   // k12: 168-216 registers, -4 %
    constexpr bool NT = (MASK >> A_NTHASH) & 1u;
    extern __shared__ __align__(128) uint8_t tile[];   // cap + 32 bytes
    __shared__ uint64_t mbar;
    __shared__ uint64_t s_nt[NT ? NT_LUT_WORDS : 1];   // ntHash seeds, filled before the barrier below
    __shared__ uint32_t s_bucket[REORDER ? 32 : 1];
    __shared__ uint32_t s_rec_len[REORDER ? 128 : 1];
    __shared__ int64_t s_rec_start[REORDER ? 128 : 1];
    __shared__ uint8_t s_slot[REORDER ? 128 : 1];

    const size_t r0 = (size_t)blockIdx.x * blockDim.x;
    const size_t r1 = (r0 + blockDim.x < n) ? r0 + blockDim.x : n;
    size_t i = r0 + threadIdx.x;
    // every thread reads the tile bounds (same address: one broadcast transaction) and its own record
    const int64_t t_start = offs[r0], t_end = offs[r1];
    int64_t s = 0, e = 0;
    if (i < n) { s = offs[i]; e = offs[i + 1]; }
    uint32_t len = (uint32_t)(e - s);
    const int64_t a0 = t_start & ~(int64_t)15;            // residue base is 16-byte aligned
    const int64_t span = (t_end - a0 + 15) & ~(int64_t)15;
    const bool staged = span <= (int64_t)cap && span > 0;
    int mixed = 0;

    if (NT && staged) nt_fill_lut<!UPPER>(s_nt);
    if (staged) {   // the copy flies while the records are being dealt out
        if (threadIdx.x == 0) mbar_init(&mbar, 1);
        // one barrier publishes the mbarrier and (REORDER) tells whether any warp holds records of different lengths
        if constexpr (REORDER) mixed = __syncthreads_or(len != __shfl_sync(0xffffffffu, len, 0));
        else __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)span;
            mbar_expect_tx(&mbar, bytes);
            for (uint32_t o = 0; o < bytes; o += TMA_CHUNK)
                tma_bulk_g2s(tile + o, res + a0 + o, (bytes - o < TMA_CHUNK) ? bytes - o : TMA_CHUNK, &mbar);
        }
    }
    if (REORDER) {
        // counting sort of the CTA's records by 64-byte blocks, descending (fixed-length reads skip it)
        if (mixed) {
            if (threadIdx.x < 32) s_bucket[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t blocks = (len + 63) >> 6;
            const uint32_t key = 31u - (blocks < 31u ? blocks : 31u);
            const uint32_t rank = atomicAdd(&s_bucket[key], 1u);
            s_rec_len[threadIdx.x] = len;
            s_rec_start[threadIdx.x] = s;
            __syncthreads();
            if (threadIdx.x < 32) {
                const uint32_t c = s_bucket[threadIdx.x];
                uint32_t incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (threadIdx.x >= (uint32_t)d) incl += v;
                }
                s_bucket[threadIdx.x] = incl - c;
            }
            __syncthreads();
            s_slot[s_bucket[key] + rank] = (uint8_t)threadIdx.x;
            __syncthreads();
            const uint32_t mine = s_slot[threadIdx.x];
            i = r0 + mine;
            len = s_rec_len[mine];
            s = s_rec_start[mine];
        }
    }

    if (span <= (int64_t)cap) {
        const uint32_t bytes = (uint32_t)span;
        const uint32_t start = (uint32_t)(s - a0);
        if (bytes) {
            mbar_wait(&mbar, 0);
            if (INPLACE) {
                // one cooperative pass: upper-case in place once for every hash of the set (seqhasher.go:249-251)
                // and, when stripping, look for whitespace candidates on the way
                uint4 *t4 = reinterpret_cast<uint4 *>(tile);
                uint32_t f = 0;
                for (uint32_t k = threadIdx.x; k < (bytes >> 4); k += blockDim.x) {
                    smem_check(t4 + k, 16);
                    uint4 v = t4[k];
                    if (STRIP) f |= has_byte_below_0x21(v.x) | has_byte_below_0x21(v.y) | has_byte_below_0x21(v.z) |
                                    has_byte_below_0x21(v.w);
                    v.x = upper4(v.x); v.y = upper4(v.y); v.z = upper4(v.z); v.w = upper4(v.w);
                    t4[k] = v;
                }
                const int any_ws = __syncthreads_or(STRIP && f != 0);
                if (STRIP && any_ws && i < n && len) len = strip_in_smem(tile, start, len);
            }
        }
        if (i < n) {
            // The readers fetch whole words, so the tail word of the tile's last record can reach up to 7 bytes past
            // the copied span.  Those bytes never enter a digest, but left as stale shared memory they can look like
            // whitespace to the speculative strip and send that record through the compaction + second hash while
            // its CTA waits (a 1.5x slowdown once the garbage happened to be small values).  Its owner, the only
            // thread that reads them, makes them letters first.
            if (SPECULATE && i + 1 == r1) {
                smem_check(tile + bytes, 8);
                *reinterpret_cast<uint2 *>(tile + bytes) = make_uint2(0x41414141u, 0x41414141u);
            }
            for (int pass = 0;; pass++) {
                bool cand;
                if (FUSED3 && len >= 17u && len <= 240u) {
                    Fused3Out fo;
                    const uint32_t wsf = fused3_record<UPPER, SPECULATE>(tile, start, len, fo);
                    store_be64(out.p[A_XXHASH] + i * 8, fo.xxh64);
                    store_be64(out.p[A_XXH3] + i * 8, fo.xxh3);
                    uint8_t *om = out.p[A_MURMUR3] + i * 16;
                    store_be64(om, fo.m1);
                    store_be64(om + 8, fo.m2);
                    cand = SPECULATE && (wsf & 0x80808080u) != 0;
                } else {
                    SmemReader<UPPER && !INPLACE, SPECULATE, FMA_PREP> r(tile, start);
                    hash_set<MASK, !UPPER>(r, len, i, out, rmask, NT ? s_nt : nullptr);
                    cand = r.saw_ws_candidate();
                }
                if (!SPECULATE || pass == 1 || !cand) break;
                len = strip_in_smem(tile, start, len);   // rare: compact this record, hash it again
            }
            if (STRIP && out_len) out_len[i] = len;
        }
    } else if (i < n) {
        if (STRIP) {
            len = strip_to_scratch(res, strip_scratch, s, len);
            if (out_len) out_len[i] = len;
            GlobalReader<UPPER, false, FMA_PREP> r(strip_scratch, s);   // written by this thread just above
            hash_set<MASK>(r, len, i, out, rmask);
        } else {
            GlobalReader<UPPER, true, FMA_PREP> r(res, s);
            hash_set<MASK>(r, len, i, out, rmask);
        }
    }
}

// largest dynamic shared memory a block may opt into on this kind of device (227 KiB on sm_100)
static int max_optin_smem()
{
    static int v = 0;
    if (v == 0) {
        int dev = 0, got = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&got, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || got <= 0) got = 232448;
        v = got;
    }
    return v;
}

template <uint32_t MASK, bool UPPER, bool STRIP>
static cudaError_t launch_tile_k(const uint8_t *res, const int64_t *offs, size_t n, const OutPtrs &out, uint32_t rmask,
                                 uint32_t cap, uint8_t *strip_scratch, int64_t *out_len, cudaStream_t st)
{
    const unsigned threads = 128;
    const size_t blocks = (n + threads - 1) / threads;
    const size_t smem = (size_t)cap + 32;
    // The limit is a property of the function, shared by every host thread that launches it (the chunk slots of the compiled
    // host do, with windows of different sizes): always ask for the device's maximum, so that concurrent callers set the same
    // value and no launch can meet a limit another thread has just lowered.
    static int dyn_max = 0;   // per instantiation: the opt-in maximum less the kernel's static shared memory (same value from every thread)
    if (dyn_max == 0) {
        cudaFuncAttributes fa;
        cudaError_t ea = cudaFuncGetAttributes(&fa, hash_tile_kernel<MASK, UPPER, STRIP>);
        if (ea != cudaSuccess) return ea;
        dyn_max = max_optin_smem() - (int)fa.sharedSizeBytes;
    }
    if (smem > (size_t)dyn_max) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(hash_tile_kernel<MASK, UPPER, STRIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max);
    if (e != cudaSuccess) return e;
    hash_tile_kernel<MASK, UPPER, STRIP><<<(unsigned)blocks, threads, smem, st>>>(res, offs, n, out, rmask, cap, strip_scratch, out_len);
    g_launches++;
    return cudaGetLastError();
}

template <uint32_t MASK>
static cudaError_t launch_tile_t(bool upper, const uint8_t *res, const int64_t *offs, size_t n, const OutPtrs &out,
                                 uint32_t rmask, uint32_t cap, uint8_t *strip_scratch, int64_t *out_len, cudaStream_t st)
{
    if (strip_scratch)
        return upper ? launch_tile_k<MASK, true, true>(res, offs, n, out, rmask, cap, strip_scratch, out_len, st)
                     : launch_tile_k<MASK, false, true>(res, offs, n, out, rmask, cap, strip_scratch, out_len, st);
    return upper ? launch_tile_k<MASK, true, false>(res, offs, n, out, rmask, cap, nullptr, out_len, st)
                 : launch_tile_k<MASK, false, false>(res, offs, n, out, rmask, cap, nullptr, out_len, st);
}

// Hash every algorithm of `mask` (bit a = algo a) over the batch; outs[a] is the digest array of algo a.
cudaError_t launch_hash_tile(uint32_t mask, bool upper, const uint8_t *res, const int64_t *offs, size_t n,
                             uint8_t *const *outs, uint32_t cap, uint8_t *strip_scratch, int64_t *out_len,
                             cudaStream_t st)
{
    if (n == 0 || mask == 0) return cudaSuccess;
    OutPtrs out;
    for (int a = 0; a < N_ALGOS; a++) out.p[a] = outs[a];
    if ((mask & (mask - 1)) == 0) {
        switch (mask) {
        case bit(A_SHA1): return launch_tile_t<bit(A_SHA1)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_SHA3): return launch_tile_t<bit(A_SHA3)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_MD5): return launch_tile_t<bit(A_MD5)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_XXHASH): return launch_tile_t<bit(A_XXHASH)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_XXH3): return launch_tile_t<bit(A_XXH3)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_CITYHASH): return launch_tile_t<bit(A_CITYHASH)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_MURMUR3): return launch_tile_t<bit(A_MURMUR3)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_NTHASH): return launch_tile_t<bit(A_NTHASH)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_BLAKE3): return launch_tile_t<bit(A_BLAKE3)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_K12): return launch_tile_t<bit(A_K12)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_RAPIDHASH): return launch_tile_t<bit(A_RAPIDHASH)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        case bit(A_RAPIDHASH32): return launch_tile_t<bit(A_RAPIDHASH32)>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
        }
        return cudaErrorInvalidValue;
    }
    if (mask == C3_MASK) return launch_tile_t<C3_MASK>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
    if ((mask & ~FAST_MASK) == 0) return launch_tile_t<FAST_MASK>(upper, res, offs, n, out, mask, cap, strip_scratch, out_len, st);
    return cudaErrorInvalidValue;   // callers split crypto algorithms into single-bit launches
}

uint32_t tile_fast_mask() { return FAST_MASK; }

// shared-memory accesses of the tile path outside the CTA's window, counted by a -DSHB_BOUNDS build (0 otherwise)
unsigned long long tile_debug_oob()
{
#ifdef SHB_BOUNDS
    unsigned long long v = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&v, g_smem_oob, sizeof v);
    return v;
#else
    return 0;
#endif
}
bool tile_debug_bounds_build()
{
#ifdef SHB_BOUNDS
    return true;
#else
    return false;
#endif
}

}  // namespace shb
