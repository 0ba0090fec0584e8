// kernels.cu — sm_100a kernels of the per-record hash path (seqhasher.go:246-259, :290-349).
//
//   hash_direct_kernel<ALGO, UPPER> : one thread per record, reading the CSR residue buffer directly
//                                     (any record length < 2^31); upper-casing fused into the loads.
//   nz_count / scan / compact       : the whitespace-strip pre-pass (flag, prefix sum, compaction,
//                                     offset rewrite) of seqhasher.go:246, optionally upper-casing.
//   synth_fill_kernel               : deterministic synthetic reads for the benchmarks.
//   probe kernels                   : register-only integer-pipe throughput probes (roofline peak).
#include "kernels.cuh"

#include "hash_all.cuh"

namespace shb {

std::atomic<unsigned long long> g_launches{0};

// perm (may be null): records in descending length order, so the 32 records of a warp need about the same
// number of blocks; digests are still written at the record's own index (input order).
constexpr uint32_t LEN_BUCKETS = 4096;   // 64-byte granularity up to 256 KiB, longer records share the first bucket
// layout of launch_length_sort's buffer: perm[n] | pad to 256 | hist[LEN_BUCKETS] | waste_q8
template <int ALGO, bool UPPER>
__global__ void __launch_bounds__(128) hash_direct_kernel(const uint8_t *__restrict__ res,
                                                          const int64_t *__restrict__ offs, size_t n,
                                                          uint8_t *__restrict__ out, const uint32_t *__restrict__ perm)
{
    // ntHash: pre-rotated seed table in shared memory (one 64-bit load per byte instead of a compare chain)
    __shared__ uint64_t s_nt[ALGO == A_NTHASH ? NT_LUT_WORDS : 1];
    if (ALGO == A_NTHASH) {
        nt_fill_lut<!UPPER>(s_nt);
        __syncthreads();
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const size_t i = perm ? (size_t)perm[t] : t;
        const int64_t s = offs[i], e = offs[i + 1];
        const uint32_t len = (uint32_t)(e - s);
        uint8_t *o = out + i * digest_size(ALGO);
        if (len == 0) {   // empty-sequence rule, seqhasher.go:292-295
            zero_digest<ALGO>(o);
            continue;
        }
        if (ALGO == A_NTHASH) {
            GlobalReader<UPPER> r(res, s);
            store_be64(o, nthash_record_lut_stream<!UPPER>(r, len, s_nt));
            continue;
        }
        // upper-casing adds on the FMA pipe, except where it measured slower (k12: 216 registers; sha1: FMA pipe busy)
        GlobalReader<UPPER, true, (ALGO != A_K12 && ALGO != A_SHA1)> r(res, s);
        hash_one<ALGO>(r, len, o);
    }
}

// Local visiting order for long records: the global length sort scatters a warp's 32 records over the whole buffer,
// which costs more (TLB reach, DRAM locality) than it saves once records are tens of kilobytes long.  Here a CTA
// takes a WINDOW of 128 consecutive records, sorts them by length in shared memory (bitonic, 28 stages, noise next
// to hashing multi-kilobyte records) and thread t takes rank t: each warp holds 32 records of similar length that
// lie within the same few megabytes.  (A window of 512 with four records per thread was tried: it quarters the
// number of records in flight and was 1.3-3x slower on 10-50 kB reads.)
constexpr uint32_t LOCAL_WINDOW = 128;   // = threads per CTA: one record per thread
template <int ALGO, bool UPPER>
__global__ void __launch_bounds__(128) hash_direct_local_kernel(const uint8_t *__restrict__ res,
                                                                const int64_t *__restrict__ offs, size_t n,
                                                                uint8_t *__restrict__ out)
{
    __shared__ uint64_t s_nt[ALGO == A_NTHASH ? NT_LUT_WORDS : 1];
    __shared__ uint64_t s_key[LOCAL_WINDOW];   // (length << 32) | index inside the window
    if (ALGO == A_NTHASH) {
        nt_fill_lut<!UPPER>(s_nt);
        __syncthreads();
    }
    const size_t windows = (n + LOCAL_WINDOW - 1) / LOCAL_WINDOW;
    for (size_t w = blockIdx.x; w < windows; w += gridDim.x) {
        const size_t base = w * LOCAL_WINDOW;
        size_t i = base + threadIdx.x;
        {
            const uint32_t len = i < n ? (uint32_t)(offs[i + 1] - offs[i]) : 0u;
            s_key[threadIdx.x] = ((uint64_t)len << 32) | threadIdx.x;
            // equal lengths inside every warp: input order is already as good as it gets
            if (__syncthreads_or(len != __shfl_sync(0xffffffffu, len, 0))) {
                for (uint32_t size = 2; size <= LOCAL_WINDOW; size <<= 1)
                    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                        if (threadIdx.x < LOCAL_WINDOW / 2) {
                            const uint32_t p = threadIdx.x;                           // pair number
                            const uint32_t lo = 2 * p - (p & (stride - 1));           // lower element of the pair
                            const uint32_t hi = lo + stride;
                            const bool descending = (lo & size) == 0;
                            const uint64_t a = s_key[lo], b = s_key[hi];
                            if ((a < b) == descending) { s_key[lo] = b; s_key[hi] = a; }
                        }
                        __syncthreads();
                    }
                i = base + (uint32_t)s_key[threadIdx.x];
            }
        }
        if (i < n) {
            const int64_t s = offs[i];
            const uint32_t len = (uint32_t)(offs[i + 1] - s);
            uint8_t *o = out + i * digest_size(ALGO);
            if (len == 0) {   // empty-sequence rule, seqhasher.go:292-295
                zero_digest<ALGO>(o);
            } else if (ALGO == A_NTHASH) {
                GlobalReader<UPPER> r(res, s);
                store_be64(o, nthash_record_lut_stream<!UPPER>(r, len, s_nt));
            } else {
                GlobalReader<UPPER, true, (ALGO != A_K12 && ALGO != A_SHA1)> r(res, s);
                hash_one<ALGO>(r, len, o);
            }
        }
        __syncthreads();   // the keys are rewritten for the next window
    }
}

template <int ALGO>
static cudaError_t launch_direct_t(bool upper, const uint8_t *res, const int64_t *offs, size_t n, uint8_t *out,
                                   const uint32_t *perm, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const int threads = 128;
    if (perm == DIRECT_LOCAL_ORDER) {
        size_t windows = (n + LOCAL_WINDOW - 1) / LOCAL_WINDOW;
        if (windows > 148u * 64u) windows = 148u * 64u;
        if (upper) hash_direct_local_kernel<ALGO, true><<<(unsigned)windows, threads, 0, st>>>(res, offs, n, out);
        else hash_direct_local_kernel<ALGO, false><<<(unsigned)windows, threads, 0, st>>>(res, offs, n, out);
        g_launches++;
        return cudaGetLastError();
    }
    size_t blocks = (n + threads - 1) / threads;
    if (blocks > 148u * 64u) blocks = 148u * 64u;
    if (upper) hash_direct_kernel<ALGO, true><<<(unsigned)blocks, threads, 0, st>>>(res, offs, n, out, perm);
    else hash_direct_kernel<ALGO, false><<<(unsigned)blocks, threads, 0, st>>>(res, offs, n, out, perm);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_hash_direct(int algo, bool upper, const uint8_t *res, const int64_t *offs, size_t n,
                               uint8_t *out, const uint32_t *perm, cudaStream_t st)
{
    switch (algo) {
    case A_SHA1: return launch_direct_t<A_SHA1>(upper, res, offs, n, out, perm, st);
    case A_SHA3: return launch_direct_t<A_SHA3>(upper, res, offs, n, out, perm, st);
    case A_MD5: return launch_direct_t<A_MD5>(upper, res, offs, n, out, perm, st);
    case A_XXHASH: return launch_direct_t<A_XXHASH>(upper, res, offs, n, out, perm, st);
    case A_XXH3: return launch_direct_t<A_XXH3>(upper, res, offs, n, out, perm, st);
    case A_CITYHASH: return launch_direct_t<A_CITYHASH>(upper, res, offs, n, out, perm, st);
    case A_MURMUR3: return launch_direct_t<A_MURMUR3>(upper, res, offs, n, out, perm, st);
    case A_NTHASH: return launch_direct_t<A_NTHASH>(upper, res, offs, n, out, perm, st);
    case A_BLAKE3: return launch_direct_t<A_BLAKE3>(upper, res, offs, n, out, perm, st);
    case A_K12: return launch_direct_t<A_K12>(upper, res, offs, n, out, perm, st);
    case A_RAPIDHASH: return launch_direct_t<A_RAPIDHASH>(upper, res, offs, n, out, perm, st);
    case A_RAPIDHASH32: return launch_direct_t<A_RAPIDHASH32>(upper, res, offs, n, out, perm, st);
    default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// Length binning for the thread-per-record kernels: a counting sort of the record indices by
// ceil(len / 64) (descending), so that a warp's 32 records finish together.  Thread-per-record over
// 30-3000-byte records otherwise runs every warp for its longest record (about half the lanes idle).
__device__ __forceinline__ uint32_t len_bucket(int64_t len)
{
    const int64_t b = (len + 63) >> 6;
    return (LEN_BUCKETS - 1) - (uint32_t)(b < (int64_t)LEN_BUCKETS - 1 ? b : (int64_t)LEN_BUCKETS - 1);   // long first
}

__global__ void __launch_bounds__(256) len_hist_kernel(const int64_t *__restrict__ offs, size_t n, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_h[LEN_BUCKETS];
    for (uint32_t k = threadIdx.x; k < LEN_BUCKETS; k += blockDim.x) s_h[k] = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&s_h[len_bucket(offs[i + 1] - offs[i])], 1u);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < LEN_BUCKETS; k += blockDim.x)
        if (s_h[k]) atomicAdd(&hist[k], s_h[k]);
}

// One CTA of 1024 threads, 4 buckets per thread: counts -> exclusive cursors (in place), plus the divergence
// estimate of INPUT order, hist[LEN_BUCKETS] = 256 * E[longest of 32 records] / E[record] in 64-byte blocks with
// the records taken as iid: P(max <= b) = F(b)^32.  (A single thread walking the 4096 buckets took 0.1-0.4 ms.)
__global__ void __launch_bounds__(1024) len_scan_kernel(uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_warp[32], s_total;
    __shared__ float s_emax, s_sumb;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (t == 0) { s_emax = 0.f; s_sumb = 0.f; }
    uint32_t c[4];
#pragma unroll
    for (int j = 0; j < 4; j++) c[j] = hist[4 * t + j];
    const uint32_t mine = c[0] + c[1] + c[2] + c[3];
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += v;
        }
        s_warp[lane] = wi - w;   // exclusive warp bases; lane 31's inclusive value is the total
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    const uint32_t total = s_total;
    uint32_t excl = s_warp[warp] + incl - mine;
    float emax = 0.f, sumb = 0.f;
    const float inv = total ? 1.f / (float)total : 0.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t k = 4 * t + j;
        hist[k] = excl;
        if (c[j]) {
            // bucket k holds the records of b = LEN_BUCKETS-1-k blocks; records no longer than that: total - excl
            float f = (float)(total - excl) * inv, fp = (float)(total - excl - c[j]) * inv;
            f *= f; f *= f; f *= f; f *= f; f *= f;
            fp *= fp; fp *= fp; fp *= fp; fp *= fp; fp *= fp;
            const float b = (float)(LEN_BUCKETS - 1 - k);
            emax += b * (f - fp);
            sumb += b * (float)c[j];
        }
        excl += c[j];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        emax += __shfl_xor_sync(0xffffffffu, emax, d);
        sumb += __shfl_xor_sync(0xffffffffu, sumb, d);
    }
    if (lane == 0 && (emax != 0.f || sumb != 0.f)) { atomicAdd(&s_emax, emax); atomicAdd(&s_sumb, sumb); }
    __syncthreads();
    if (t == 0) {
        const float mean_b = total ? s_sumb * inv : 0.f;
        hist[LEN_BUCKETS] = mean_b > 0.f ? (uint32_t)(s_emax / mean_b * 256.f) : 256u;
    }
}

// Each CTA ranks a contiguous run of SCATTER_PER_CTA records inside shared memory and reserves one range per
// bucket it met with a single global atomic.  (One global atomic per record serialises on the few hot buckets:
// 0.3-0.5 ms for a million records of nearly equal length.)
constexpr uint32_t SCATTER_PER_THREAD = 8, SCATTER_PER_CTA = 256 * SCATTER_PER_THREAD;
__global__ void __launch_bounds__(256) len_scatter_kernel(const int64_t *__restrict__ offs, size_t n,
                                                          uint32_t *__restrict__ cursor, uint32_t *__restrict__ perm,
                                                          uint32_t identity_below_q8)
{
    __shared__ uint32_t s_cnt[LEN_BUCKETS];    // records of this CTA per bucket, then the reserved base
    if (cursor[LEN_BUCKETS] < identity_below_q8) {   // input order wastes too little to pay for scattered reads
        for (uint32_t j = 0; j < SCATTER_PER_THREAD; j++) {
            const size_t i = (size_t)blockIdx.x * SCATTER_PER_CTA + j * 256 + threadIdx.x;
            if (i < n) perm[i] = (uint32_t)i;
        }
        return;
    }
    for (uint32_t k = threadIdx.x; k < LEN_BUCKETS; k += 256) s_cnt[k] = 0;
    __syncthreads();
    const size_t first = (size_t)blockIdx.x * SCATTER_PER_CTA;
    uint32_t bucket[SCATTER_PER_THREAD], rank[SCATTER_PER_THREAD];
#pragma unroll
    for (uint32_t j = 0; j < SCATTER_PER_THREAD; j++) {
        const size_t i = first + j * 256 + threadIdx.x;
        if (i < n) {
            bucket[j] = len_bucket(offs[i + 1] - offs[i]);
            rank[j] = atomicAdd(&s_cnt[bucket[j]], 1u);
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < LEN_BUCKETS; k += 256) {
        const uint32_t c = s_cnt[k];
        if (c) s_cnt[k] = atomicAdd(&cursor[k], c);
    }
    __syncthreads();
#pragma unroll
    for (uint32_t j = 0; j < SCATTER_PER_THREAD; j++) {
        const size_t i = first + j * 256 + threadIdx.x;
        if (i < n) perm[s_cnt[bucket[j]] + rank[j]] = (uint32_t)i;
    }
}

size_t length_sort_bytes(size_t n) { return (n * 4 + 255) / 256 * 256 + (LEN_BUCKETS + 64) * 4; }

// perm_and_tmp: length_sort_bytes(n) bytes; perm = first n uint32.  identity_below_q8: when the divergence
// estimate of input order (x256) is below it, perm is the identity (input order is kept).
cudaError_t launch_length_sort(const int64_t *offs, size_t n, void *perm_and_tmp, uint32_t identity_below_q8, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    uint32_t *perm = static_cast<uint32_t *>(perm_and_tmp);
    uint32_t *hist = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(perm_and_tmp) + (n * 4 + 255) / 256 * 256);
    cudaError_t e = cudaMemsetAsync(hist, 0, LEN_BUCKETS * 4, st);
    if (e != cudaSuccess) return e;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    len_hist_kernel<<<(unsigned)blocks, 256, 0, st>>>(offs, n, hist);
    len_scan_kernel<<<1, 1024, 0, st>>>(hist);
    len_scatter_kernel<<<(unsigned)((n + SCATTER_PER_CTA - 1) / SCATTER_PER_CTA), 256, 0, st>>>(offs, n, hist, perm, identity_below_q8);
    g_launches += 3;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Normalise pre-pass.  Byte block b covers [b*4096, (b+1)*4096); thread t of the CTA owns the 16
// bytes at b*4096 + 16*t.

// nonzero iff some byte of w is below 0x21 (the classic "has a byte less than n" test: exact as an any-test)
__device__ __forceinline__ uint32_t any_below_0x21(uint32_t w) { return (w - 0x21212121u) & ~w & 0x80808080u; }

// bit j of the result = byte j of the 16-byte chunk is kept (inside the buffer and not whitespace)
template <bool STRIP>
__device__ __forceinline__ uint32_t keep_mask16(const uint4 &v, size_t pos, size_t total)
{
    // sequence bytes are letters: unless some byte is below 0x21 nothing can be whitespace, and the byte-wise test (a hundred
    // instructions per chunk) is skipped — it made the pre-pass ALU-bound at a third of the memory rate
    if (pos + 16 <= total && (!STRIP || (any_below_0x21(v.x) | any_below_0x21(v.y) | any_below_0x21(v.z) | any_below_0x21(v.w)) == 0u))
        return 0xffffu;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t c = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
        bool keep = (pos + j < total) && !(STRIP && is_ws(c));
        m |= (keep ? 1u : 0u) << j;
    }
    return m;
}

__device__ __forceinline__ uint4 load16_guarded(const uint8_t *res, size_t pos, size_t total)
{
    // the residue buffer is readable up to the next 16-byte boundary past `total`
    if (pos < total) return __ldg(reinterpret_cast<const uint4 *>(res + pos));
    return make_uint4(0, 0, 0, 0);
}

template <bool STRIP>
__global__ void __launch_bounds__(256) nz_count_kernel(const uint8_t *__restrict__ res, size_t total,
                                                       uint32_t *__restrict__ block_counts, size_t nb)
{
    // one warp per 4 KiB block, eight 128-bit loads in flight per lane (one unit per thread and one block per CTA left the
    // memory pipe at 1.7 TB/s)
    const size_t blk = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= nb) return;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const size_t pos = blk * NORM_BLK + 16u * (lane + 32u * k);
        cnt += __popc(keep_mask16<STRIP>(load16_guarded(res, pos, total), pos, total));
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) block_counts[blk] = cnt;
}

// exclusive scan of nb block counts into nb+1 prefixes; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const uint32_t *__restrict__ counts, size_t nb,
                                                           int64_t *__restrict__ prefix)
{
    const size_t per = (nb + 1023) / 1024;
    const size_t lo = (size_t)threadIdx.x * per;
    const size_t hi = lo + per < nb ? lo + per : nb;
    int64_t sum = 0;
    for (size_t i = lo; i < hi; i++) sum += counts[i];
    __shared__ int64_t s_w[32];
    int64_t incl = sum;
    for (int d = 1; d < 32; d <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        int64_t v = s_w[threadIdx.x], iv = v;
        for (int d = 1; d < 32; d <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, iv, d);
            if (threadIdx.x >= d) iv += t;
        }
        s_w[threadIdx.x] = iv - v;
    }
    __syncthreads();
    int64_t run = s_w[threadIdx.x >> 5] + incl - sum;
    for (size_t i = lo; i < hi; i++) {
        prefix[i] = run;
        run += counts[i];
    }
    if (hi == nb && lo <= nb) prefix[nb] = run;   // several threads may write the same total
}

// ---- general exclusive scan: n uint32 counts -> n+1 int64 prefixes (strip pre-pass, BLAKE3 chunk bases,
// record splitter, formatter).  Three phases over tiles of 4096 counts: tile sums, scan of the tile sums by
// one CTA, per-tile scan seeded with its base.  tile_sums needs ceil(n/4096) uint32 + (that + 1) int64.
constexpr uint32_t SCAN_TILE = 4096;   // 256 threads x 16 counts

__global__ void __launch_bounds__(256) scan_tile_sum_kernel(const uint32_t *__restrict__ counts, size_t n,
                                                            uint32_t *__restrict__ tile_sums)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + 16u * threadIdx.x;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++)
        if (base + k < n) s += counts[base + k];
    __shared__ uint32_t s_w[8];
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) t += s_w[i];
        tile_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) scan_tile_apply_kernel(const uint32_t *__restrict__ counts, size_t n,
                                                              const int64_t *__restrict__ tile_base,
                                                              int64_t *__restrict__ prefix)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + 16u * threadIdx.x;
    uint32_t v[16];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) { v[k] = (base + k < n) ? counts[base + k] : 0u; s += v[k]; }
    __shared__ uint32_t s_w[8];
    uint32_t incl = s;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    int64_t run = tile_base[blockIdx.x] + (incl - s);
    for (uint32_t i = 0; i < (threadIdx.x >> 5); i++) run += s_w[i];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (base + k < n) prefix[base + k] = run;
        run += v[k];
    }
    // the closing element prefix[n] = grand total is written by scan_blocks_kernel's tile_base[ntiles]
}

__global__ void scan_close_kernel(const int64_t *__restrict__ tile_base, size_t ntiles, size_t n, int64_t *__restrict__ prefix)
{
    prefix[n] = tile_base[ntiles];
}

size_t scan_tmp_bytes(size_t n)
{
    const size_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return ((ntiles + 64) * 4 + 7) / 8 * 8 + (ntiles + 64) * 8;
}

// tmp: scan_tmp_bytes(n) bytes owned by the caller (so scans on different streams never share it)
cudaError_t scan_counts(const uint32_t *counts, size_t n, int64_t *prefix, void *tmp, cudaStream_t st)
{
    if (n <= 65536 || tmp == nullptr) {   // small inputs: one CTA does it all
        scan_blocks_kernel<<<1, 1024, 0, st>>>(counts, n, prefix);
        g_launches++;
        return cudaGetLastError();
    }
    const size_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *tile_sums = static_cast<uint32_t *>(tmp);
    int64_t *tile_base = reinterpret_cast<int64_t *>(static_cast<uint8_t *>(tmp) + ((ntiles + 64) * 4 + 7) / 8 * 8);
    scan_tile_sum_kernel<<<(unsigned)ntiles, 256, 0, st>>>(counts, n, tile_sums);
    scan_blocks_kernel<<<1, 1024, 0, st>>>(tile_sums, ntiles, tile_base);
    scan_tile_apply_kernel<<<(unsigned)ntiles, 256, 0, st>>>(counts, n, tile_base, prefix);
    scan_close_kernel<<<1, 1, 0, st>>>(tile_base, ntiles, n, prefix);
    g_launches += 4;
    return cudaGetLastError();
}

template <bool STRIP, bool UPPER>
__global__ void __launch_bounds__(256) compact_kernel(const uint8_t *__restrict__ res, size_t total,
                                                      const int64_t *__restrict__ offs, size_t n,
                                                      const int64_t *__restrict__ block_prefix,
                                                      uint8_t *__restrict__ out_res, int64_t *__restrict__ out_offs,
                                                      size_t nb, size_t per_cta)
{
    // A CTA takes per_cta CONSECUTIVE 4 KiB blocks: the record boundaries that fall into them are consecutive entries of offs,
    // so one binary search per CTA finds the first and the rest is a walk.  (One block per CTA with its own search was 23
    // dependent loads of latency per 4 KiB: 15 of the 78 ms of config C5, profiles/r02_launches_c5_before_prepass.csv.)
    __shared__ uint32_t s_excl[256];
    __shared__ uint32_t s_mask[256];
    __shared__ uint32_t s_w[8];
    __shared__ __align__(16) uint8_t s_out[NORM_BLK + 32];
    __shared__ uint4 s_last[8];                                    // every warp's last unit, for the next warp's first lane
    const size_t blk0 = (size_t)blockIdx.x * per_cta, blk1 = blk0 + per_cta < nb ? blk0 + per_cta : nb;
    if (block_prefix[nb] == (int64_t)total) {
        // Nothing is dropped anywhere in the batch (the usual case: a reader has removed the line breaks already, and
        // sequences hold no blanks): the output is the input, upper-cased, at the same place — a copy at memory speed instead
        // of the scan / realign machinery below (5.2 -> 2.6 ms for the 7.6 GB of config C5).
        const size_t units = (total + 15) >> 4, u0 = blk0 * (NORM_BLK / 16), u1 = blk1 * (NORM_BLK / 16) < units ? blk1 * (NORM_BLK / 16) : units;
        const uint4 *src = reinterpret_cast<const uint4 *>(res);
        uint4 *dst = reinterpret_cast<uint4 *>(out_res);
        size_t u = u0 + threadIdx.x;
        for (; u + 768 < u1; u += 1024) {                          // four loads in flight per thread
            uint4 a = __ldg(src + u), b = __ldg(src + u + 256), c = __ldg(src + u + 512), d = __ldg(src + u + 768);
            if (UPPER) {
                a = make_uint4(upper4(a.x), upper4(a.y), upper4(a.z), upper4(a.w));
                b = make_uint4(upper4(b.x), upper4(b.y), upper4(b.z), upper4(b.w));
                c = make_uint4(upper4(c.x), upper4(c.y), upper4(c.z), upper4(c.w));
                d = make_uint4(upper4(d.x), upper4(d.y), upper4(d.z), upper4(d.w));
            }
            dst[u] = a; dst[u + 256] = b; dst[u + 512] = c; dst[u + 768] = d;
        }
        for (; u < u1; u += 256) {
            uint4 a = __ldg(src + u);
            if (UPPER) a = make_uint4(upper4(a.x), upper4(a.y), upper4(a.z), upper4(a.w));
            dst[u] = a;                                            // the last unit may run past `total`: both buffers are padded to 16
        }
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (size_t)gridDim.x * blockDim.x) out_offs[i] = offs[i];
        return;
    }
    if (blk0 >= blk1) return;
    size_t cur = 0;                                                // first i with offs[i] >= start of the CTA's first block
    {
        size_t lo = 0, hi = n + 1;
        const size_t first = blk0 * NORM_BLK;
        while (lo < hi) {
            size_t mid = (lo + hi) >> 1;
            if ((size_t)offs[mid] < first) lo = mid + 1; else hi = mid;
        }
        cur = lo;
    }
    uint4 vnext = load16_guarded(res, blk0 * NORM_BLK + 16u * threadIdx.x, total);
  for (size_t blk = blk0; blk < blk1; blk++) {
    const size_t bstart = blk * NORM_BLK;
    const size_t pos = bstart + 16u * threadIdx.x;
    const uint4 v = vnext;
    if (blk + 1 < blk1) vnext = load16_guarded(res, pos + NORM_BLK, total);   // the next block's bytes are on their way
    const uint32_t mask = keep_mask16<STRIP>(v, pos, total);
    const uint32_t cnt = __popc(mask);
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) wbase += (i < (int)(threadIdx.x >> 5)) ? s_w[i] : 0u;
    const uint32_t excl = wbase + incl - cnt;
    s_excl[threadIdx.x] = excl;
    s_mask[threadIdx.x] = mask;
    const int64_t bp = block_prefix[blk];
    // The block's kept bytes form one contiguous run out_res[bp, bp + blk_total).  They are compacted in shared memory at the
    // same 16-byte phase as their destination and copied out with aligned 128-bit stores (byte stores straight to global
    // memory made this kernel 0.85 TB/s: 17.9 of the 83.7 ms of config C5, profiles/r02_launches_bench.csv).
    const uint32_t phase = (uint32_t)(bp & 15);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (UPPER) {
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = upper4(w[k]);
    }
    uint32_t blk_total = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) blk_total += s_w[i];
    const uint32_t blk_valid = (uint32_t)(total - bstart < NORM_BLK ? total - bstart : NORM_BLK);
    uint8_t *gbase = out_res + (bp - phase);                       // 16-byte aligned: out_res is, and bp - phase is a multiple of 16
    const uint32_t end = phase + blk_total;                        // the block's bytes are [phase, end) of the window at gbase
    if (blk_total == blk_valid) {
        // Nothing to drop in this block (the usual case: a reader has removed the line breaks already): the output is the input
        // moved by a constant, so every thread builds one ALIGNED 16-byte output unit from its own unit and its left
        // neighbour's (shuffle; the warp's first lane takes it from the previous warp through shared memory) and stores it
        // straight to global memory.  Staging through shared memory costs 16 byte stores per thread at a 16-byte lane stride,
        // a 4-way bank conflict each: 14 of the 78 ms of config C5.
        if ((threadIdx.x & 31) == 31) s_last[threadIdx.x >> 5] = make_uint4(w[0], w[1], w[2], w[3]);
        __syncthreads();
        uint32_t pw[4];
#pragma unroll
        for (int k = 0; k < 4; k++) pw[k] = __shfl_up_sync(0xffffffffu, w[k], 1);
        if ((threadIdx.x & 31) == 0 && threadIdx.x) {
            const uint4 q = s_last[(threadIdx.x >> 5) - 1];
            pw[0] = q.x; pw[1] = q.y; pw[2] = q.z; pw[3] = q.w;
        }
        // window = [left neighbour's 16 bytes | my 16 bytes]; my output unit starts `phase` bytes before my own bytes
        const uint32_t W[8] = {pw[0], pw[1], pw[2], pw[3], w[0], w[1], w[2], w[3]};
        const uint32_t ws = phase >> 2, sh = 32u - 8u * (phase & 3u);   // sh = 32: whole words, funnelshift_rc clamps to "take the high word"
        uint32_t o[4];
        switch (ws) {                                                // uniform over the block
        case 0:
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = __funnelshift_rc(W[3 + j], W[4 + j], sh);
            break;
        case 1:
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = __funnelshift_rc(W[2 + j], W[3 + j], sh);
            break;
        case 2:
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = __funnelshift_rc(W[1 + j], W[2 + j], sh);
            break;
        default:
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = __funnelshift_rc(W[0 + j], W[1 + j], sh);
            break;
        }
        const uint32_t lo16 = 16u * threadIdx.x, hi16 = lo16 + 16u;
        if (lo16 >= phase && hi16 <= end) {
            *reinterpret_cast<uint4 *>(gbase + lo16) = make_uint4(o[0], o[1], o[2], o[3]);
        } else if (lo16 < end) {                                     // the first unit (shared with the previous block) or the last valid one
            for (uint32_t k = (lo16 > phase ? lo16 : phase); k < (hi16 < end ? hi16 : end); k++)
                gbase[k] = (uint8_t)(o[(k - lo16) >> 2] >> (8 * ((k - lo16) & 3)));
        }
        // the spill-over unit: the last `phase` bytes of the block land in the unit after the last thread's
        if (phase && threadIdx.x == blockDim.x - 1 && hi16 < end) {
            for (uint32_t k = hi16; k < end; k++) {
                const uint32_t srcb = k - phase - lo16;             // byte of my own unit
                gbase[k] = (uint8_t)(w[srcb >> 2] >> (8 * (srcb & 3)));
            }
        }
    } else {
        uint8_t *dst = s_out + phase + excl;
        uint32_t m = mask;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (m & 1u) *dst++ = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
            m >>= 1;
        }
        __syncthreads();
        for (uint32_t u = threadIdx.x; u * 16u < end; u += blockDim.x) {
            const uint32_t lo16 = u * 16u, hi16 = lo16 + 16u;
            if (lo16 >= phase && hi16 <= end) {
                *reinterpret_cast<uint4 *>(gbase + lo16) = *reinterpret_cast<const uint4 *>(s_out + lo16);
            } else {                                                 // the run's first / last unit is shared with the neighbouring block
                for (uint32_t k = (lo16 > phase ? lo16 : phase); k < (hi16 < end ? hi16 : end); k++) gbase[k] = s_out[k];
            }
        }
    }
    __syncthreads();
    // record boundaries that fall inside this byte block: offs[i] in [bstart, bstart + 4096), i = cur, cur + 1, ...
    while (true) {
        const size_t i = cur + threadIdx.x;
        const size_t p = i <= n ? (size_t)offs[i] : ~(size_t)0;
        const bool inside = p < bstart + NORM_BLK;
        if (inside) {
            const uint32_t rel = (uint32_t)(p - bstart);
            const uint32_t t = rel >> 4, within = rel & 15u;
            out_offs[i] = bp + s_excl[t] + __popc(s_mask[t] & ((1u << within) - 1u));
        }
        const int got = __syncthreads_count(inside);               // also the barrier before the shared arrays are rewritten
        cur += (size_t)got;
        if (got < (int)blockDim.x) break;
    }
  }
}

__global__ void lengths_kernel(const int64_t *__restrict__ offs, size_t n, int64_t *__restrict__ out_len)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out_len[i] = offs[i + 1] - offs[i];
}

// rapidhash32 from finished rapidhash digests (both big-endian in memory): lo32 ^ hi32, seqhasher.go:340-343
__global__ void rapid_fold32_kernel(const uint2 *__restrict__ h64, size_t n, uint32_t *__restrict__ h32)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint2 v = h64[i];
        h32[i] = v.x ^ v.y;   // an XOR of the halves commutes with the byte order
    }
}

cudaError_t launch_rapid_fold32(const uint8_t *h64, size_t n, uint8_t *h32, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    rapid_fold32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint2 *>(h64), n, reinterpret_cast<uint32_t *>(h32));
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_lengths(const int64_t *offs, size_t n, int64_t *out_len, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    lengths_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(offs, n, out_len);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_normalise(const uint8_t *res, const int64_t *offs, size_t n, size_t total_bytes, bool strip,
                             bool upper, uint32_t *block_counts, int64_t *block_prefix, uint8_t *out_res,
                             int64_t *out_offs, int64_t *out_len, void *scan_tmp, cudaStream_t st)
{
    const size_t nb = norm_blocks(total_bytes);
    if (strip) nz_count_kernel<true><<<(unsigned)((nb + 7) / 8), 256, 0, st>>>(res, total_bytes, block_counts, nb);
    else nz_count_kernel<false><<<(unsigned)((nb + 7) / 8), 256, 0, st>>>(res, total_bytes, block_counts, nb);
    {
        cudaError_t es = scan_counts(block_counts, nb, block_prefix, scan_tmp, st);
        if (es != cudaSuccess) return es;
    }
    // persistent grid: exactly the CTAs that are resident at once (one more per SM would run as a second, short wave), every CTA
    // walking a run of consecutive blocks (at least 16, so that the one binary search per CTA is noise)
    static int resident = 0;
    if (resident == 0) {
        int occ = 0, sms = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, compact_kernel<true, true>, 256, 0);
        resident = (occ > 0 ? occ : 4) * (sms > 0 ? sms : 148);
    }
    size_t ctas = (size_t)resident;
    size_t per_cta = (nb + ctas - 1) / ctas;
    if (per_cta < 16) per_cta = 16;
    ctas = (nb + per_cta - 1) / per_cta;
    if (strip && upper)
        compact_kernel<true, true><<<(unsigned)ctas, 256, 0, st>>>(res, total_bytes, offs, n, block_prefix, out_res, out_offs, nb, per_cta);
    else if (strip)
        compact_kernel<true, false><<<(unsigned)ctas, 256, 0, st>>>(res, total_bytes, offs, n, block_prefix, out_res, out_offs, nb, per_cta);
    else if (upper)
        compact_kernel<false, true><<<(unsigned)ctas, 256, 0, st>>>(res, total_bytes, offs, n, block_prefix, out_res, out_offs, nb, per_cta);
    else
        compact_kernel<false, false><<<(unsigned)ctas, 256, 0, st>>>(res, total_bytes, offs, n, block_prefix, out_res, out_offs, nb, per_cta);
    g_launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (out_len) return launch_lengths(out_offs, n, out_len, st);
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------
// synthetic reads: byte j = f(seed, j), so any slice can be regenerated on the host

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void synth_fill_kernel(uint8_t *__restrict__ res, size_t total, uint64_t seed, uint32_t lower_permille,
                                  uint32_t n_permille)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t j0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; j0 < total; j0 += stride) {
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint64_t z = splitmix64(seed + (j0 + k) * 0x9E3779B97F4A7C15ULL);
            uint32_t c = (0x54474341u >> (8 * (uint32_t)(z & 3))) & 0xffu;   // "ACGT"
            const uint32_t u_n = (uint32_t)(((z >> 32) & 0xFFFFFu) * 1000u >> 20);
            const uint32_t u_l = (uint32_t)(((z >> 8) & 0xFFFFFu) * 1000u >> 20);
            if (u_n < n_permille) c = 'N';
            if (u_l < lower_permille) c |= 0x20u;
            packed |= c << (8 * k);
        }
        if (j0 + 4 <= total) {
            *reinterpret_cast<uint32_t *>(res + j0) = packed;
        } else {
            for (int k = 0; j0 + k < total; k++) res[j0 + k] = (uint8_t)(packed >> (8 * k));
        }
    }
}

// bytes >= 0x80 in res[0, total): SHB_FLAG_CHECK_ASCII (the caller cannot vouch for ASCII input; seqhasher.go:243-251)
__global__ void __launch_bounds__(256) count_nonascii_kernel(const uint8_t *__restrict__ res, size_t total, unsigned long long *__restrict__ count)
{
    const size_t units = (total + 15) >> 4;   // the buffer is readable (zero padded) to the next 16 bytes
    uint32_t c = 0;
    for (size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(res) + u);
        uint32_t m[4] = {v.x & 0x80808080u, v.y & 0x80808080u, v.z & 0x80808080u, v.w & 0x80808080u};
        if (16 * u + 16 > total) {            // the last unit: bytes past the end do not count
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long long valid = (long long)total - (long long)(16 * u + 4 * k);
                if (valid <= 0) m[k] = 0;
                else if (valid < 4) m[k] &= (1u << (8 * valid)) - 1u;
            }
        }
        c += __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

cudaError_t launch_count_nonascii(const uint8_t *res, size_t total_bytes, unsigned long long *count, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess || total_bytes == 0) return e;
    count_nonascii_kernel<<<148 * 8, 256, 0, st>>>(res, total_bytes, count);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_synth_fill(uint8_t *res, size_t total_bytes, uint64_t seed, unsigned lower_permille,
                              unsigned n_permille, cudaStream_t st)
{
    if (total_bytes == 0) return cudaSuccess;
    synth_fill_kernel<<<148 * 16, 256, 0, st>>>(res, total_bytes, seed, lower_permille, n_permille);
    g_launches++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// integer-pipe probes: 8 independent register chains per thread, 64 ops per chain per iteration

#define PROBE_OP_LOP3(x, y, z) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z))
#define PROBE_OP_IADD3(x, y, z) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(x) : "r"(y), "r"(z))
#define PROBE_OP_SHF(x, y) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(y))
#define PROBE_OP_IMAD(x, y, z) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z))
#define PROBE_OP_PRMT(x, y) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x) : "r"(y))
#define PROBE_OP_ADDI(x) asm volatile("add.u32 %0, %0, 0x1f1f1f1f;" : "+r"(x))
#define PROBE_OP_ADD2(x, y) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y))

template <int KIND>
__global__ void __launch_bounds__(256) probe_kernel(int iters, uint32_t *sink)
{
    uint32_t c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x;
    const uint32_t y = c[0] | 1u, z = c[1] | 3u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 64; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (KIND == 0) PROBE_OP_LOP3(c[k], y, z);
                else if (KIND == 1) PROBE_OP_IADD3(c[k], y, z);
                else if (KIND == 2) PROBE_OP_SHF(c[k], y);
                else if (KIND == 3) PROBE_OP_IMAD(c[k], y, z);
                else if (KIND == 4) PROBE_OP_PRMT(c[k], y);
                else if (KIND == 5) {
                    // SHA-1 mix LOP3:IADD3:SHF:PRMT = 208:165:224:16 (of 613) ~ 22:17:23:2 of 64
                    if (u < 22) PROBE_OP_LOP3(c[k], y, z);
                    else if (u < 39) PROBE_OP_IADD3(c[k], y, z);
                    else if (u < 62) PROBE_OP_SHF(c[k], y);
                    else PROBE_OP_PRMT(c[k], y);
                } else if (KIND == 6) {
                    if (u & 1) PROBE_OP_IMAD(c[k], y, z);
                    else PROBE_OP_LOP3(c[k], y, z);
                } else if (KIND == 32) {          // add with an immediate: ptxas emits VIADD — which pipe takes it?
                    PROBE_OP_ADDI(c[k]);
                } else if (KIND == 33) {          // VIADD : IMAD 1:1
                    if (u & 1) PROBE_OP_IMAD(c[k], y, z);
                    else PROBE_OP_ADDI(c[k]);
                } else if (KIND == 34) {          // VIADD : LOP3 1:1
                    if (u & 1) PROBE_OP_LOP3(c[k], y, z);
                    else PROBE_OP_ADDI(c[k]);
                } else if (KIND == 35) {          // two-input register add (IADD3 / IMAD.IADD: ptxas' choice) : IMAD 1:1
                    if (u & 1) PROBE_OP_IMAD(c[k], y, z);
                    else PROBE_OP_ADD2(c[k], y);
                } else if (KIND == 36) {          // two-input register add : LOP3 1:1
                    if (u & 1) PROBE_OP_LOP3(c[k], y, z);
                    else PROBE_OP_ADD2(c[k], y);
                } else if (KIND == 37) {          // LOP3 : IMAD 2:1
                    if (u % 3 == 2) PROBE_OP_IMAD(c[k], y, z);
                    else PROBE_OP_LOP3(c[k], y, z);
                } else if (KIND == 39) {          // LOP3 : add-imm : IMAD : add-imm
                    if ((u & 3) == 0) PROBE_OP_LOP3(c[k], y, z);
                    else if ((u & 3) == 2) PROBE_OP_IMAD(c[k], y, z);
                    else PROBE_OP_ADDI(c[k]);
                } else if (KIND == 38) {          // LOP3 : IMAD 1:2
                    if (u % 3 == 2) PROBE_OP_LOP3(c[k], y, z);
                    else PROBE_OP_IMAD(c[k], y, z);
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= c[k];
    if (acc == 0x12345u) sink[0] = acc;   // practically never true; keeps the chains alive
}

// kinds 8..13: the 64-bit building blocks of the multiply-based hashes, 8 independent 64-bit chains per thread, 64 ops per
// chain per iteration, counted as ONE op each:  8 mad.wide.u32 (IMAD.WIDE.U32)   9 mad.hi.u32 (IMAD.HI.U32)
// 10 c = c * y + z on 64 bits (mul64: what xxh64 / murmur3 rounds are made of)   11 c = fold(c * y) (128-bit product, xxh3 /
// rapidhash)   12 64-bit rotate   13 64-bit add
template <int KIND>
__global__ void __launch_bounds__(256) probe64_kernel(int iters, uint32_t *sink)
{
    uint64_t c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = ((uint64_t)(threadIdx.x * 2654435761u + k * 40503u + blockIdx.x) << 32) | (threadIdx.x * 40503u + k);
    const uint64_t y = c[0] | 1u, z = c[1] | 3u;
    const uint32_t y32 = (uint32_t)y;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 64; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (KIND == 8) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((uint32_t)c[k]), "r"(y32));
                else if (KIND == 9) {
                    uint32_t lo = (uint32_t)c[k];
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo) : "r"(y32), "r"((uint32_t)z));
                    c[k] = (c[k] & 0xffffffff00000000ull) | lo;
                } else if (KIND == 10) c[k] = c[k] * y + z;
                else if (KIND == 11) c[k] = (c[k] * y) ^ __umul64hi(c[k], y);
                else if (KIND == 12) c[k] = (c[k] << 31) | (c[k] >> 33);
                else if (KIND == 13) c[k] = c[k] + y;
            }
            if (KIND == 12 || KIND == 13) {   // keep the compiler from collapsing the chain
#pragma unroll
                for (int k = 0; k < 8; k++) asm volatile("" : "+l"(c[k]));
            }
        }
    }
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= c[k];
    if (acc == 0x12345u) sink[0] = (uint32_t)acc;
}

// kinds 16..31: the real SHA-1 compression on register data, pipe-balancing variant = kind - 16
// (613 fused ops per call by the survey's count); kind 7 = variant 0
template <int VARIANT>
__global__ void __launch_bounds__(128) probe_sha1_kernel(int iters, uint32_t *sink, PipeConsts pc)
{
    uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
    for (int it = 0; it < iters; it++) {
        sha1_compress_v<VARIANT>(h, w, pc);
        w[0] ^= h[0];
    }
    // digest of the chain; the host cross-checks every variant against variant 0
    uint32_t x = h[0] ^ h[1] ^ h[2] ^ h[3] ^ h[4];
    for (int d = 16; d; d >>= 1) x ^= __shfl_xor_sync(0xffffffffu, x, d);
    if ((threadIdx.x & 31) == 0) atomicXor(sink + 1, x);
}

template <int V>
static void launch_probe_sha1(int iters, uint32_t *sink, cudaStream_t st)
{
    probe_sha1_kernel<V><<<148 * 16, 128, 0, st>>>(iters, sink, pipe_consts());
}

cudaError_t launch_probe(int kind, int iters, uint32_t *sink, double *lane_instr, cudaStream_t st)
{
    const unsigned blocks = 148 * 8;
    switch (kind) {
    case 0: probe_kernel<0><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 1: probe_kernel<1><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 2: probe_kernel<2><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 3: probe_kernel<3><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 4: probe_kernel<4><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 5: probe_kernel<5><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 6: probe_kernel<6><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 32: probe_kernel<32><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 33: probe_kernel<33><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 34: probe_kernel<34><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 35: probe_kernel<35><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 36: probe_kernel<36><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 37: probe_kernel<37><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 38: probe_kernel<38><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 39: probe_kernel<39><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 8: probe64_kernel<8><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 9: probe64_kernel<9><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 10: probe64_kernel<10><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 11: probe64_kernel<11><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 12: probe64_kernel<12><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 13: probe64_kernel<13><<<blocks, 256, 0, st>>>(iters, sink); break;
    case 7: case 16: launch_probe_sha1<0>(iters, sink, st); break;
    case 17: launch_probe_sha1<1>(iters, sink, st); break;
    case 18: launch_probe_sha1<2>(iters, sink, st); break;
    case 19: launch_probe_sha1<3>(iters, sink, st); break;
    case 20: launch_probe_sha1<4>(iters, sink, st); break;
    case 21: launch_probe_sha1<5>(iters, sink, st); break;
    case 22: launch_probe_sha1<6>(iters, sink, st); break;
    case 23: launch_probe_sha1<7>(iters, sink, st); break;
    case 24: launch_probe_sha1<8>(iters, sink, st); break;
    case 25: launch_probe_sha1<9>(iters, sink, st); break;
    case 26: launch_probe_sha1<10>(iters, sink, st); break;
    case 27: launch_probe_sha1<11>(iters, sink, st); break;
    case 28: launch_probe_sha1<12>(iters, sink, st); break;
    case 29: launch_probe_sha1<13>(iters, sink, st); break;
    case 30: launch_probe_sha1<14>(iters, sink, st); break;
    case 31: launch_probe_sha1<15>(iters, sink, st); break;
    case 40: launch_probe_sha1<24>(iters, sink, st); break;   // 8 + rol30 as IMAD.HI
    case 56: launch_probe_sha1<40>(iters, sink, st); break;   // 8 + schedule rol1 as IMAD.HI
    case 72: launch_probe_sha1<56>(iters, sink, st); break;   // 8 + both
    default: return cudaErrorInvalidValue;
    }
    g_launches++;
    if (kind == 7 || (kind >= 16 && kind < 32) || kind == 40 || kind == 56 || kind == 72) *lane_instr = 148.0 * 16 * 128 * (double)iters * 613.0;
    else *lane_instr = (double)blocks * 256.0 * (double)iters * 64.0 * 8.0 * (kind == 1 ? 1.0 : 1.0);
    return cudaGetLastError();
}

}  // namespace shb
