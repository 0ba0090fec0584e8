// long.cu — the long-record path (10-50 kb nanopore-like reads, BASELINE config C4): the hashes that
// have parallelism INSIDE one record get it here instead of one serial thread per record.
//
//   ntHash  : warp per record.  The forward hash is a pure XOR of rotated seeds (nthash.go ntf64;
//             seqhasher.go:320-327), so lanes take 16-byte units with coalesced 128-bit loads, fold each
//             unit with pre-rotated seed look-ups and XOR-reduce across the warp.
//   BLAKE3  : thread per 1024-byte CHUNK over the flattened chunk list of the whole batch (every lane
//             always has a full chunk of work, whatever the record lengths), chaining values to a scratch
//             array; then one thread per record merges its CVs up the binary tree (n-1 parents against
//             16n chunk compressions).  Replaces blake3.Sum256 (seqhasher.go:328-330).
//
// Upper-casing (seqhasher.go:249-251) is fused into the loads of both.
#include "hash_all.cuh"
#include "hash_nt_planes.cuh"

namespace shb {

// Grid of a persistent (grid-stride) kernel: exactly the CTAs that are resident at once.  Registers and shared memory decide
// how many fit per SM; one CTA more than fit would run as a second, nearly empty wave after the others have done their share.
template <class K>
static int resident_ctas(K kernel, int threads)
{
    int occ = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0);
    return (occ > 0 ? occ : 4) * (sms > 0 ? sms : 148);
}

// ------------------------------------------------------------------------------------------- ntHash

__device__ __forceinline__ uint64_t rotl64v(uint64_t x, uint32_t r)   // r in 0..63, run-time
{
    r &= 63u;
    return r ? (x << r) | (x >> (64 - r)) : x;
}

// byte k of the result is 0xff where position (pos + k) lies inside [0, len), else 0
__device__ __forceinline__ uint32_t valid_mask4(int64_t pos, int64_t len)
{
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (pos + k >= 0 && pos + k < len) m |= 0xffu << (8 * k);
    return m;
}

// Warp per record.  A lane folds whole 16-byte units (one coalesced LDG.128 each): byte j of a unit looks up table
// 15 - j of sixteen pre-rotated copies of the seed table (32 KiB of shared memory per CTA), so a unit is 16 look-ups XORed
// together with no rotation at all; the units of one lane lie 512 bytes apart, i.e. they all carry the same rotation
// (512 = 0 mod 64), which is applied once per lane before the warp's XOR reduction.  Four units are in flight per lane.
// Bound: the 8 bytes every look-up returns per lane — shared memory delivers 128 B per clock and SM, two wavefronts per
// LDS.64 — not HBM and not the ALU.
__device__ __forceinline__ uint4 ldg16_issue(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

#ifndef NT_MIN_BLOCKS
#define NT_MIN_BLOCKS 4   // 64 registers: room for four 128-bit loads in flight per lane (A/B: profiles/r02_nthash_variants.txt)
#endif
constexpr uint32_t NT16_WORDS = 16 * 256;   // uint64 entries: [rotation 0..15][byte]
#ifndef NT_PLANES
#define NT_PLANES 1   // units of pure ACGT take the bit-plane path (no shared-memory look-ups); 0: every unit takes the look-ups
#endif

// One look-up = PRMT (byte out of the word, ALU pipe) + IMAD (byte * 8 + table base, FMA pipe: the multiplier comes from
// constant memory so that ptxas cannot turn it into an ALU-pipe LEA) + LDS.64 with the table's offset as an immediate + one
// LOP3 per half (three-input XORs).  Two ALU instructions per byte: what the fused BLAKE3 + ntHash pass pays on its
// bounding pipe.
static __constant__ uint32_t c_eight = 8u;
template <int I>   // byte I (0..15) of the unit: table 15 - I
__device__ __forceinline__ void nt_look(uint32_t x, uint32_t base, uint32_t &lo, uint32_t &hi)
{
    const uint32_t b = __byte_perm(x, 0, 0x4440 | (I & 3));               // byte I & 3 of the word, zero-extended
    const uint32_t a = b * c_eight + base;
    uint32_t l, h;
    asm("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(l), "=r"(h) : "r"(a), "n"((15 - I) * 2048));
    lo ^= l;
    hi ^= h;
}
template <bool UPPER>
__device__ __forceinline__ uint64_t nt_unit16(uint4 v, const uint64_t *lut)
{
    const uint32_t base = smem_u32(lut);
    const uint32_t x0 = UPPER ? upper4(v.x) : v.x, x1 = UPPER ? upper4(v.y) : v.y, x2 = UPPER ? upper4(v.z) : v.z,
                   x3 = UPPER ? upper4(v.w) : v.w;
    uint32_t lo = 0, hi = 0;
    nt_look<0>(x0, base, lo, hi); nt_look<1>(x0, base, lo, hi); nt_look<2>(x0, base, lo, hi); nt_look<3>(x0, base, lo, hi);
    nt_look<4>(x1, base, lo, hi); nt_look<5>(x1, base, lo, hi); nt_look<6>(x1, base, lo, hi); nt_look<7>(x1, base, lo, hi);
    nt_look<8>(x2, base, lo, hi); nt_look<9>(x2, base, lo, hi); nt_look<10>(x2, base, lo, hi); nt_look<11>(x2, base, lo, hi);
    nt_look<12>(x3, base, lo, hi); nt_look<13>(x3, base, lo, hi); nt_look<14>(x3, base, lo, hi); nt_look<15>(x3, base, lo, hi);
    return mk64(lo, hi);
}

constexpr uint32_t NT_PEND = 8;   // units per lane that may wait for the look-up path

template <bool UPPER>
__global__ void __launch_bounds__(256, NT_MIN_BLOCKS) nthash_warp_kernel(const uint8_t *__restrict__ res,
                                                          const int64_t *__restrict__ offs, size_t n,
                                                          uint8_t *__restrict__ out)
{
    __shared__ uint64_t lut[NT16_WORDS];
    __shared__ uint32_t pend[NT_PLANES ? NT_PEND : 1][256];              // per thread: units waiting for the look-up path
    for (uint32_t c = threadIdx.x; c < 256; c += blockDim.x) {
        const uint64_t s = g_nt_seed_table.v[c];
#pragma unroll
        for (int r = 0; r < 16; r++) lut[r * 256 + c] = r ? rotl64(s, r) : s;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const int64_t s = offs[i], e = offs[i + 1];
        const int64_t len = e - s;
        uint8_t *o = out + i * 8;
        if (len == 0) {   // empty-sequence rule, seqhasher.go:292-295
            if (lane == 0) *reinterpret_cast<uint2 *>(o) = make_uint2(0u, 0u);
            continue;
        }
        // 16-byte units of the aligned window that covers the record; unit u holds record bytes [16 u - head, 16 u - head + 16)
        const int64_t a0 = s & ~(int64_t)15;
        const uint32_t head = (uint32_t)(s - a0);
        const uint32_t nunits = (uint32_t)((e - a0 + 15) >> 4);           // record < 2^31 bytes
        const uint4 *base = reinterpret_cast<const uint4 *>(res + a0);
        uint64_t acc = 0;
        uint32_t aq = 0, and_ = 0, n_fast = 0, n_pend = 0;                // the bit-plane path's share (NT_PLANES)
        // interior units: 1 .. nunits - 2 (the first and the last may hold bytes of the neighbours)
        uint32_t u = lane;
        if (u == 0) u = 32;                                               // lane 0 takes unit 0 separately below
        const uint32_t last = nunits - 1;
        // a unit the plane path refuses: note it for later, or fold it now when this lane's list is full
        auto other = [&](uint4 v, uint32_t uu) {
            if (n_pend < NT_PEND) pend[n_pend++][threadIdx.x] = uu;
            else acc ^= nt_unit16<UPPER>(v, lut);
        };
        for (; u + 96 < last; u += 128) {
            // four loads issued back to back (volatile: ptxas otherwise sinks the third and fourth below the first units' work)
            const uint4 v0 = ldg16_issue(base + u), v1 = ldg16_issue(base + u + 32), v2 = ldg16_issue(base + u + 64), v3 = ldg16_issue(base + u + 96);
            if (NT_PLANES) {
                if (!nt_unit_planes(v0, aq, and_, n_fast)) other(v0, u);
                if (!nt_unit_planes(v1, aq, and_, n_fast)) other(v1, u + 32);
                if (!nt_unit_planes(v2, aq, and_, n_fast)) other(v2, u + 64);
                if (!nt_unit_planes(v3, aq, and_, n_fast)) other(v3, u + 96);
            } else {
                acc ^= nt_unit16<UPPER>(v0, lut) ^ nt_unit16<UPPER>(v1, lut);
                acc ^= nt_unit16<UPPER>(v2, lut) ^ nt_unit16<UPPER>(v3, lut);
            }
        }
        for (; u < last; u += 32) {
            const uint4 v = __ldg(base + u);
            if (!NT_PLANES) acc ^= nt_unit16<UPPER>(v, lut);
            else if (!nt_unit_planes(v, aq, and_, n_fast)) other(v, u);
        }
        if (NT_PLANES) {
            acc ^= nt_planes_to_hash(aq, and_, n_fast);
            const uint32_t rounds = __reduce_max_sync(0xffffffffu, n_pend);
            for (uint32_t k = 0; k < rounds; k++)
                if (k < n_pend) acc ^= nt_unit16<UPPER>(__ldg(base + pend[k][threadIdx.x]), lut);
        }
        // edge units: bytes outside the record become 0, whose seed is 0
        if (lane == 0 || (last % 32u == lane && last != 0)) {
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                const uint32_t eu = pass == 0 ? 0u : last;
                if (pass == 0 ? lane != 0 : (last == 0 || last % 32u != lane)) continue;
                uint4 v = __ldg(base + eu);
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
                const int64_t pos = (int64_t)16 * eu - head;
#pragma unroll
                for (int k = 0; k < 4; k++) w[k] &= valid_mask4(pos + 4 * k, len);
                acc ^= nt_unit16<UPPER>(make_uint4(w[0], w[1], w[2], w[3]), lut);
            }
        }
        // byte j of unit u sits at record position p = 16 u - head + j and carries rotation (len - 1 - p) mod 64; the table gave
        // it 15 - j, so the unit still owes (len - 16 - 16 u + head) mod 64 — the same for every unit of this lane
        acc = rotl64v(acc, (uint32_t)((len - 16 + head - 16 * (int64_t)lane) & 63));
#pragma unroll
        for (int d = 16; d; d >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) store_be64(o, acc);
    }
}

cudaError_t launch_nthash_long(bool upper, const uint8_t *res, const int64_t *offs, size_t n, uint8_t *out,
                               cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    static int per_sm[2] = {0, 0};
    if (!per_sm[upper]) per_sm[upper] = upper ? resident_ctas(nthash_warp_kernel<true>, 256) : resident_ctas(nthash_warp_kernel<false>, 256);
    size_t blocks = (n + 7) / 8;   // 8 warps per CTA
    if (blocks > (size_t)per_sm[upper]) blocks = (size_t)per_sm[upper];
    if (upper) nthash_warp_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(res, offs, n, out);
    else nthash_warp_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(res, offs, n, out);
    g_launches++;
    return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------- XXH3

// 16 bytes at any address as 4 LE words: two aligned 128-bit loads + funnel shift (m is warp-uniform here).  The loads and
// the assembly are separate calls so that a caller can issue the loads of its next step before it consumes these.
struct Raw16 {
    uint4 a, b;
};
__device__ __forceinline__ Raw16 load16_raw(const uint8_t *p)
{
    const uint4 *q = reinterpret_cast<const uint4 *>((uintptr_t)p & ~(uintptr_t)15);
    Raw16 r;
    r.a = __ldg(q);
    r.b = make_uint4(0, 0, 0, 0);
    if ((uintptr_t)p & 15u) r.b = __ldg(q + 1);
    return r;
}
template <bool UPPER>
__device__ __forceinline__ void assemble16(const Raw16 &r, uint32_t m, uint32_t (&w)[4])
{
    const uint4 a = r.a, b = r.b;
    const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const uint32_t sh = (m & 3u) * 8u;
    switch (m >> 2) {
    case 0:
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = __funnelshift_r(u[i], u[i + 1], sh);
        break;
    case 1:
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = __funnelshift_r(u[i + 1], u[i + 2], sh);
        break;
    case 2:
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = __funnelshift_r(u[i + 2], u[i + 3], sh);
        break;
    default:
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = __funnelshift_r(u[i + 3], u[i + 4], sh);
        break;
    }
    if (UPPER) {
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = upper4(w[i]);
    }
}
template <bool UPPER>
__device__ __forceinline__ void load16_any(const uint8_t *p, uint32_t (&w)[4])
{
    assemble16<UPPER>(load16_raw(p), (uint32_t)((uintptr_t)p & 15u), w);
}

// 64-bit LE word of the secret copy in shared memory at any byte offset
__device__ __forceinline__ uint64_t ssec(const uint32_t *sec, uint32_t off)
{
    const uint32_t i = off >> 2, s = (off & 3u) * 8u;
    const uint32_t w0 = sec[i], w1 = sec[i + 1], w2 = sec[i + 2];
    return mk64(__funnelshift_r(w0, w1, s), __funnelshift_r(w1, w2, s));
}

// Warp per record, records > 240 bytes (the XXH3 "long" path, A.12): lane = 4*sub + pair owns accumulators
// (2*pair, 2*pair+1) of stripes sub and sub+8 of every 1024-byte block; one warp load covers 8 consecutive
// stripes (512 contiguous bytes).  Stripes inside a block commute, so the 8 sub-groups are summed with
// shuffles before the per-block scramble.  Shorter records fall to lane 0 running the generic routine.
template <bool UPPER>
__global__ void __launch_bounds__(256) xxh3_warp_kernel(const uint8_t *__restrict__ res, const int64_t *__restrict__ offs,
                                                        size_t n, uint8_t *__restrict__ out)
{
    __shared__ uint32_t sec[52];
    if (threadIdx.x < 50) sec[threadIdx.x] = reinterpret_cast<const uint32_t *>(c_xxh3_secret)[threadIdx.x];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, sub = lane >> 2, pair = lane & 3u;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    constexpr uint64_t INIT[8] = {XP32_3, XP64_1, XP64_2, XP64_3, XP64_4, XP32_2, XP64_5, XP32_1};
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const int64_t s = offs[i];
        const uint64_t len = (uint64_t)(offs[i + 1] - s);
        uint8_t *o = out + i * 8;
        if (len <= 240) {
            if (lane == 0) {
                if (len == 0) *reinterpret_cast<uint2 *>(o) = make_uint2(0u, 0u);
                else { GlobalReader<UPPER> r(res, s); store_be64(o, xxh3_record(r, (uint32_t)len)); }
            }
            continue;
        }
        const uint8_t *p = res + s;
        const uint64_t init0 = pair == 0 ? INIT[0] : pair == 1 ? INIT[2] : pair == 2 ? INIT[4] : INIT[6];
        const uint64_t init1 = pair == 0 ? INIT[1] : pair == 1 ? INIT[3] : pair == 2 ? INIT[5] : INIT[7];
        uint64_t a0 = sub == 0 ? init0 : 0, a1 = sub == 0 ? init1 : 0;   // sub 0 carries the state, the others deltas
        const uint64_t nblocks = (len - 1) >> 10;
        const uint32_t tail_stripes = (uint32_t)(((len - 1) - (nblocks << 10)) >> 6);
        // the loads of block b + 1 are issued before block b is folded and reduced (10 shuffles + the scramble per block would
        // otherwise leave the memory pipe idle: long-scoreboard stalls were 10 per issue slot)
        const uint32_t mis = (uint32_t)((uintptr_t)(p + 16u * pair) & 15u);   // the same for every unit of this lane (64 | strides)
        Raw16 nxt[2];
#pragma unroll
        for (uint32_t half = 0; half < 2; half++) {
            const uint32_t st = sub + 8u * half;
            nxt[half].a = nxt[half].b = make_uint4(0, 0, 0, 0);
            if (st < (nblocks ? 16u : tail_stripes)) nxt[half] = load16_raw(p + 64u * st + 16u * pair);
        }
        for (uint64_t b = 0; b <= nblocks; b++) {
            const bool last = b == nblocks;
            const uint32_t nst = last ? tail_stripes : 16u;
            Raw16 cur[2] = {nxt[0], nxt[1]};
            if (!last) {
                const uint32_t nst_next = (b + 1 == nblocks) ? tail_stripes : 16u;
#pragma unroll
                for (uint32_t half = 0; half < 2; half++) {
                    const uint32_t st = sub + 8u * half;
                    if (st < nst_next) nxt[half] = load16_raw(p + ((b + 1) << 10) + 64u * st + 16u * pair);
                }
            }
#pragma unroll
            for (uint32_t half = 0; half < 2; half++) {
                const uint32_t st = sub + 8u * half;
                if (st < nst) {
                    uint32_t w[4];
                    assemble16<UPPER>(cur[half], mis, w);
                    const uint64_t d0 = mk64(w[0], w[1]), d1 = mk64(w[2], w[3]);
                    const uint64_t k0 = d0 ^ ssec(sec, 8u * st + 16u * pair), k1 = d1 ^ ssec(sec, 8u * st + 16u * pair + 8u);
                    a0 += d1 + (uint64_t)(uint32_t)k0 * (uint64_t)(uint32_t)(k0 >> 32);
                    a1 += d0 + (uint64_t)(uint32_t)k1 * (uint64_t)(uint32_t)(k1 >> 32);
                }
            }
            // sum the 8 sub-groups (lane bits 2..4)
#pragma unroll
            for (int d = 4; d < 32; d <<= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, d);
                a1 += __shfl_xor_sync(0xffffffffu, a1, d);
            }
            if (!last) {   // scramble (every lane holds the block sum), then only sub 0 keeps it
                a0 = (a0 ^ (a0 >> 47) ^ ssec(sec, 128u + 16u * pair)) * XP32_1;
                a1 = (a1 ^ (a1 >> 47) ^ ssec(sec, 136u + 16u * pair)) * XP32_1;
                if (sub) { a0 = 0; a1 = 0; }
            }
        }
        // last stripe at len-64 with the secret at 121, then the merge; lanes 0..3 (sub 0) each own one pair
        uint64_t term = 0;
        if (sub == 0) {
            uint32_t w[4];
            load16_any<UPPER>(p + len - 64 + 16u * pair, w);
            const uint64_t d0 = mk64(w[0], w[1]), d1 = mk64(w[2], w[3]);
            const uint64_t k0 = d0 ^ ssec(sec, 121u + 16u * pair), k1 = d1 ^ ssec(sec, 129u + 16u * pair);
            a0 += d1 + (uint64_t)(uint32_t)k0 * (uint64_t)(uint32_t)(k0 >> 32);
            a1 += d0 + (uint64_t)(uint32_t)k1 * (uint64_t)(uint32_t)(k1 >> 32);
            term = mulfold64(a0 ^ ssec(sec, 11u + 16u * pair), a1 ^ ssec(sec, 19u + 16u * pair));
        }
        term += __shfl_xor_sync(0xffffffffu, term, 1);
        term += __shfl_xor_sync(0xffffffffu, term, 2);
        if (lane == 0) store_be64(o, xxh3_avalanche(len * XP64_1 + term));
    }
}

cudaError_t launch_xxh3_long(bool upper, const uint8_t *res, const int64_t *offs, size_t n, uint8_t *out, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    static int per_sm[2] = {0, 0};
    if (!per_sm[upper]) per_sm[upper] = upper ? resident_ctas(xxh3_warp_kernel<true>, 256) : resident_ctas(xxh3_warp_kernel<false>, 256);
    size_t blocks = (n + 7) / 8;
    if (blocks > (size_t)per_sm[upper]) blocks = (size_t)per_sm[upper];
    if (upper) xxh3_warp_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(res, offs, n, out);
    else xxh3_warp_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(res, offs, n, out);
    g_launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- K12

// leaves of record i = chunks 1.. of S = M || 0x00 (8192 bytes each); records of <= 8191 bytes have none
__global__ void k12_count_leaves_kernel(const int64_t *__restrict__ offs, size_t n, uint32_t *__restrict__ counts)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t len = offs[i + 1] - offs[i];
        counts[i] = len >= 8192 ? (uint32_t)(((len + 1 + 8191) >> 13) - 1) : 0u;
    }
}

// One thread per leaf over the flattened leaf list of the batch: CV (32 bytes) to cvs[global leaf id].  Threads
// from s0_first on take one record each and run the record-only part of its final node (the 48 full rate blocks of
// S_0) into s0[record][25]: those 8 KiB per record would otherwise be serial work of the final kernel.
template <bool UPPER>
__global__ void __launch_bounds__(128) k12_leaves_kernel(const uint8_t *__restrict__ res, const int64_t *__restrict__ offs,
                                                         size_t n, const int64_t *__restrict__ leaf_base,
                                                         uint64_t *__restrict__ cvs, int64_t s0_first, uint64_t *__restrict__ s0)
{
    const int64_t total = leaf_base[n];
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= s0_first) {
        const size_t i = (size_t)(g - s0_first);
        if (i >= n) return;
        const int64_t s = offs[i];
        if (offs[i + 1] - s < 8192) return;
        GlobalReader<UPPER, true, false> r(res, s);
        uint64_t f[25];
        kt128_s0_state(r, f);
        uint64_t *dst = s0 + i * 25;
#pragma unroll
        for (int k = 0; k < 25; k++) dst[k] = f[k];
        return;
    }
    if (g >= total) return;
    size_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const size_t mid = (lo + hi) >> 1;
        if (leaf_base[mid] <= g) lo = mid; else hi = mid;
    }
    const size_t i = lo;
    const int64_t s = offs[i];
    const uint32_t len = (uint32_t)(offs[i + 1] - s);
    const uint32_t c = (uint32_t)(g - leaf_base[i]) + 1u;
    GlobalReader<UPPER, true, false> r(res, s);   // k12: upper-casing adds stay on the ALU pipe (measured)
    uint64_t cv[4];
    kt128_leaf_cv(r, len, c, cv);
    uint64_t *dst = cvs + g * 4;
    *reinterpret_cast<ulonglong2 *>(dst) = make_ulonglong2(cv[0], cv[1]);
    *reinterpret_cast<ulonglong2 *>(dst + 2) = make_ulonglong2(cv[2], cv[3]);
}

// One thread per record: the single-sponge form for short records, else the final node over S_0 and the
// leaf CVs computed by k12_leaves_kernel.
template <bool UPPER>
__global__ void __launch_bounds__(128) k12_final_kernel(const uint8_t *__restrict__ res, const int64_t *__restrict__ offs,
                                                        size_t n, const int64_t *__restrict__ leaf_base,
                                                        const uint64_t *__restrict__ cvs, const uint64_t *__restrict__ s0,
                                                        uint8_t *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = offs[i];
    const uint32_t len = (uint32_t)(offs[i + 1] - s);
    uint8_t *o = out + i * 128;
    if (len == 0) {   // empty-sequence rule
        zero_digest<A_K12>(o);
        return;
    }
    GlobalReader<UPPER, true, false> r(res, s);   // k12: upper-casing adds stay on the ALU pipe (measured)
    if (len < 8192) {
        kt128_record(r, len, o);
        return;
    }
    const uint64_t *my = cvs + leaf_base[i] * 4;
    uint64_t f[25];
#pragma unroll
    for (int k = 0; k < 25; k++) f[k] = s0[i * 25 + k];
    kt128_final_from_state(r, len, o, f, [&](uint32_t c, uint64_t (&cv)[4]) {
        const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(my + (size_t)(c - 1) * 4));
        const ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(my + (size_t)(c - 1) * 4 + 2));
        cv[0] = a.x; cv[1] = a.y; cv[2] = b.x; cv[3] = b.y;
    });
}

cudaError_t launch_k12_long(bool upper, const uint8_t *res, const int64_t *offs, size_t n, size_t total_bytes,
                            uint8_t *out, void *scratch, cudaStream_t st)
{
    // scratch layout = the BLAKE3 one (counts, bases, 32-byte CVs, scan tmp): b3_long_scratch_bytes() covers it
    if (n == 0) return cudaSuccess;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    uint8_t *base = static_cast<uint8_t *>(scratch);
    uint32_t *counts = reinterpret_cast<uint32_t *>(base);
    int64_t *leaf_base = reinterpret_cast<int64_t *>(base + up(n * sizeof(uint32_t)));
    uint64_t *cvs = reinterpret_cast<uint64_t *>(base + up(n * sizeof(uint32_t)) + up((n + 1) * sizeof(int64_t)));
    const size_t max_chunks0 = total_bytes / 1024 + n + 1;
    void *scan_tmp = base + up(n * sizeof(uint32_t)) + up((n + 1) * sizeof(int64_t)) + up(max_chunks0 * 32);
    k12_count_leaves_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(offs, n, counts);
    g_launches++;
    cudaError_t e = scan_counts(counts, n, leaf_base, scan_tmp, st);
    if (e != cudaSuccess) return e;
    const size_t max_leaves = total_bytes / 8192 + 1;
    const unsigned leaf_blocks = (unsigned)((max_leaves + 127) / 128);
    const int64_t s0_first = (int64_t)leaf_blocks * 128;
    const unsigned blocks = leaf_blocks + (unsigned)((n + 127) / 128);
    // S_0 states (25 lanes per record) behind the CVs inside the region sized for BLAKE3's 32 bytes per KiB
    uint64_t *s0 = cvs + up(max_leaves * 32) / 8;
    if (up(max_leaves * 32) + n * 200 > up(max_chunks0 * 32)) return cudaErrorInvalidValue;   // callers pass mean >= 8 KiB
    if (upper) {
        k12_leaves_kernel<true><<<blocks, 128, 0, st>>>(res, offs, n, leaf_base, cvs, s0_first, s0);
        k12_final_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(res, offs, n, leaf_base, cvs, s0, out);
    } else {
        k12_leaves_kernel<false><<<blocks, 128, 0, st>>>(res, offs, n, leaf_base, cvs, s0_first, s0);
        k12_final_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(res, offs, n, leaf_base, cvs, s0, out);
    }
    g_launches += 2;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------- BLAKE3

// chunks of record i = max(1, ceil(len/1024)); written as uint32 counts for the shared scan kernel
__global__ void b3_count_chunks_kernel(const int64_t *__restrict__ offs, size_t n, uint32_t *__restrict__ counts)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t len = offs[i + 1] - offs[i];
        counts[i] = len ? (uint32_t)((len + 1023) >> 10) : 0u;   // empty records own no chunk
    }
}

// 16 message words of the 64-byte block at p (any alignment): five aligned 128-bit loads + funnel shifts
template <bool UPPER>
__device__ __forceinline__ void load_block16(const uint8_t *p, uint32_t (&w)[16])
{
    const uint4 *q = reinterpret_cast<const uint4 *>((uintptr_t)p & ~(uintptr_t)15);
    const uint32_t m = (uint32_t)((uintptr_t)p & 15u);
    uint32_t u[20];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint4 v = __ldg(q + k);
        u[4 * k] = v.x; u[4 * k + 1] = v.y; u[4 * k + 2] = v.z; u[4 * k + 3] = v.w;
    }
    if (m) {
        uint4 v = __ldg(q + 4);
        u[16] = v.x; u[17] = v.y; u[18] = v.z; u[19] = v.w;
    } else {
        u[16] = u[17] = u[18] = u[19] = 0;
    }
    const uint32_t sh = (m & 3u) * 8u;
    switch (m >> 2) {
    case 0:
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = __funnelshift_r(u[i], u[i + 1], sh);
        break;
    case 1:
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = __funnelshift_r(u[i + 1], u[i + 2], sh);
        break;
    case 2:
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = __funnelshift_r(u[i + 2], u[i + 3], sh);
        break;
    default:
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = __funnelshift_r(u[i + 3], u[i + 4], sh);
        break;
    }
    if (UPPER) {
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = upper4(w[i]);
    }
}

// One thread per chunk of the flattened chunk list.  chunk_base[i] = number of chunks before record i
// (exclusive scan of the counts, chunk_base[n] = total).  Single-chunk records are finished here (ROOT);
// for the others the chunk CV goes to cvs[global chunk id].
// NT: the same pass also folds the chunk's bytes into its share of the record's ntHash (`--hash blake3,nthash`, config C4:
// the residues are read once for both).  Byte j of a chunk sits at record position 1024 c + j and carries the rotation
// (len - 1 - j) mod 64 — the chunk index drops out because 1024 = 0 mod 64 — so a chunk's share is T_c = XOR_j rol(seed[b_j],
// 63 - j), the same for every 64-byte block of it, and the record's hash is rol(XOR_c T_c, len mod 64).  BLAKE3 is bound by the
// ALU pipe; the look-ups ride on the LSU / shared-memory path it leaves idle and add one PRMT + one LOP3 per byte to the ALU.
template <bool UPPER, bool NT>
__global__ void __launch_bounds__(128) b3_chunks_kernel(const uint8_t *__restrict__ res,
                                                        const int64_t *__restrict__ offs, size_t n,
                                                        const int64_t *__restrict__ chunk_base,
                                                        uint32_t *__restrict__ cvs, uint8_t *__restrict__ out,
                                                        uint64_t *__restrict__ nt_part, uint8_t *__restrict__ nt_out)
{
    __shared__ uint64_t lut[NT ? NT16_WORDS : 1];
    if (NT) {
        for (uint32_t c = threadIdx.x; c < 256; c += blockDim.x) {
            const uint64_t s = g_nt_seed_table.v[c];
#pragma unroll
            for (int r = 0; r < 16; r++) lut[r * 256 + c] = r ? rotl64(s, r) : s;
        }
        __syncthreads();
    }
    uint64_t nt = 0;
    const int64_t total = chunk_base[n];
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    // record that owns chunk g: last i with chunk_base[i] <= g (records with zero chunks are skipped)
    size_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const size_t mid = (lo + hi) >> 1;
        if (chunk_base[mid] <= g) lo = mid; else hi = mid;
    }
    const size_t i = lo;
    const int64_t s = offs[i], len = offs[i + 1] - s;
    const uint32_t c = (uint32_t)(g - chunk_base[i]);
    const uint32_t nchunks = (uint32_t)((len + 1023) >> 10);
    const uint32_t clen = (c + 1 < nchunks) ? 1024u : (uint32_t)(len - ((int64_t)c << 10));
    const uint32_t root = (nchunks == 1) ? (uint32_t)B3_ROOT : 0u;
    // One loop for every lane of the warp: full 64-byte blocks take the aligned 128-bit loads, the single
    // partial block of a record's last chunk takes masked reads, and the compression (the expensive part)
    // is issued once per iteration for all lanes whatever their chunk length.
    uint32_t cv[8];
    b3_set_iv(cv);
    const int64_t cbase = s + ((int64_t)c << 10);
    const uint32_t nblocks = clen ? (clen + 63) >> 6 : 1;
#pragma unroll 1
    for (uint32_t b = 0; b < 16; b++) {
        if (b < nblocks) {
            uint32_t m[16];
            const uint32_t off = b << 6;
            uint32_t bl = 64;
            if (off + 64 <= clen) {
                load_block16<UPPER>(res + cbase + off, m);
            } else {
                bl = clen - off;
                GlobalReader<UPPER> r(res, cbase);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t o = off + 4 * k;
                    uint32_t v = 0;
                    if (o < clen) {
                        v = r.u32(o);
                        if (o + 4 > clen) v &= (1u << (8 * (clen - o))) - 1u;
                    }
                    m[k] = v;
                }
            }
            const uint32_t flags = (b == 0 ? (uint32_t)B3_CHUNK_START : 0u) |
                                   (b == nblocks - 1 ? ((uint32_t)B3_CHUNK_END | root) : 0u);
            if (NT) {
                // bytes past the chunk's end are zero in m[] (seed 0); group q of the block still owes the rotation 48 - 16 q
                const uint64_t g0 = nt_unit16<false>(make_uint4(m[0], m[1], m[2], m[3]), lut);
                const uint64_t g1 = nt_unit16<false>(make_uint4(m[4], m[5], m[6], m[7]), lut);
                const uint64_t g2 = nt_unit16<false>(make_uint4(m[8], m[9], m[10], m[11]), lut);
                const uint64_t g3 = nt_unit16<false>(make_uint4(m[12], m[13], m[14], m[15]), lut);
                nt ^= rotl64_shf(g0, 48) ^ rotl64_shf(g1, 32) ^ rotl64_shf(g2, 16) ^ g3;
            }
            b3_compress(cv, m, c, 0u, bl, flags);
        }
    }
    uint32_t *dst = root ? reinterpret_cast<uint32_t *>(out + i * 32) : cvs + g * 8;
    *reinterpret_cast<uint4 *>(dst) = make_uint4(cv[0], cv[1], cv[2], cv[3]);
    *reinterpret_cast<uint4 *>(dst + 4) = make_uint4(cv[4], cv[5], cv[6], cv[7]);
    if (NT) {
        if (root) store_be64(nt_out + i * 8, rotl64v(nt, (uint32_t)(len & 63)));
        else nt_part[g] = nt;
    }
}

// One thread per record with >= 2 chunks: lazy binary-counter merge of its chunk CVs (same tree as
// blake3_record in hash_blake3.cuh); empty records get the all-zero slot.
template <bool NT>
__global__ void __launch_bounds__(128) b3_merge_kernel(const int64_t *__restrict__ offs, size_t n,
                                                       const int64_t *__restrict__ chunk_base,
                                                       const uint32_t *__restrict__ cvs, uint8_t *__restrict__ out,
                                                       const uint64_t *__restrict__ nt_part, uint8_t *__restrict__ nt_out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t base = chunk_base[i];
    const uint32_t nchunks = (uint32_t)(chunk_base[i + 1] - base);
    uint32_t *o = reinterpret_cast<uint32_t *>(out + i * 32);
    if (nchunks == 0) {   // empty-sequence rule
        *reinterpret_cast<uint4 *>(o) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(o + 4) = make_uint4(0, 0, 0, 0);
        if (NT) *reinterpret_cast<uint2 *>(nt_out + i * 8) = make_uint2(0u, 0u);
        return;
    }
    if (nchunks == 1) return;   // finished by b3_chunks_kernel
    if (NT) {
        uint64_t t = 0;
        for (uint32_t c = 0; c < nchunks; c++) t ^= __ldg(nt_part + base + c);
        store_be64(nt_out + i * 8, rotl64v(t, (uint32_t)((offs[i + 1] - offs[i]) & 63)));
    }
    uint32_t stack[22][8];
    int depth = 0;
    uint32_t cv[8];
    for (uint32_t c = 0; c < nchunks; c++) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(cvs + (base + c) * 8));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(cvs + (base + c) * 8 + 4));
        cv[0] = a.x; cv[1] = a.y; cv[2] = a.z; cv[3] = a.w; cv[4] = b.x; cv[5] = b.y; cv[6] = b.z; cv[7] = b.w;
        if (c + 1 < nchunks) {
            for (uint32_t tot = c + 1; (tot & 1u) == 0; tot >>= 1) b3_parent(stack[--depth], cv, 0u);
#pragma unroll
            for (int k = 0; k < 8; k++) stack[depth][k] = cv[k];
            depth++;
        }
    }
    while (depth > 0) {
        depth--;
        b3_parent(stack[depth], cv, depth == 0 ? (uint32_t)B3_ROOT : 0u);
    }
    *reinterpret_cast<uint4 *>(o) = make_uint4(cv[0], cv[1], cv[2], cv[3]);
    *reinterpret_cast<uint4 *>(o + 4) = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

size_t b3_long_scratch_bytes(size_t total_bytes, size_t n)
{
    const size_t max_chunks = total_bytes / 1024 + n + 1;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    // counts | chunk bases | 32-byte CVs | scan scratch | 8-byte ntHash shares (blake3 + nthash in one pass)
    return up(n * sizeof(uint32_t)) + up((n + 1) * sizeof(int64_t)) + up(max_chunks * 32) + up(scan_tmp_bytes(n)) + up(max_chunks * 8);
}


cudaError_t launch_blake3_long(bool upper, const uint8_t *res, const int64_t *offs, size_t n, size_t total_bytes,
                               uint8_t *out, void *scratch, uint8_t *nt_out, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    uint8_t *base = static_cast<uint8_t *>(scratch);
    uint32_t *counts = reinterpret_cast<uint32_t *>(base);
    int64_t *chunk_base = reinterpret_cast<int64_t *>(base + up(n * sizeof(uint32_t)));
    uint32_t *cvs = reinterpret_cast<uint32_t *>(base + up(n * sizeof(uint32_t)) + up((n + 1) * sizeof(int64_t)));
    b3_count_chunks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(offs, n, counts);
    g_launches++;
    const size_t max_chunks0 = total_bytes / 1024 + n + 1;
    void *scan_tmp = base + up(n * sizeof(uint32_t)) + up((n + 1) * sizeof(int64_t)) + up(max_chunks0 * 32);
    cudaError_t e = scan_counts(counts, n, chunk_base, scan_tmp, st);
    if (e != cudaSuccess) return e;
    const size_t max_chunks = total_bytes / 1024 + n + 1;   // upper bound of chunk_base[n]; the kernel re-checks
    const unsigned blocks = (unsigned)((max_chunks + 127) / 128);
    uint64_t *nt_part = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(scan_tmp) + up(scan_tmp_bytes(n)));
    const unsigned mblocks = (unsigned)((n + 127) / 128);
    if (nt_out) {   // blake3 + nthash: one pass over the residues
        if (upper) b3_chunks_kernel<true, true><<<blocks, 128, 0, st>>>(res, offs, n, chunk_base, cvs, out, nt_part, nt_out);
        else b3_chunks_kernel<false, true><<<blocks, 128, 0, st>>>(res, offs, n, chunk_base, cvs, out, nt_part, nt_out);
        b3_merge_kernel<true><<<mblocks, 128, 0, st>>>(offs, n, chunk_base, cvs, out, nt_part, nt_out);
    } else {
        if (upper) b3_chunks_kernel<true, false><<<blocks, 128, 0, st>>>(res, offs, n, chunk_base, cvs, out, nullptr, nullptr);
        else b3_chunks_kernel<false, false><<<blocks, 128, 0, st>>>(res, offs, n, chunk_base, cvs, out, nullptr, nullptr);
        b3_merge_kernel<false><<<mblocks, 128, 0, st>>>(offs, n, chunk_base, cvs, out, nullptr, nullptr);
    }
    g_launches += 2;
    return cudaGetLastError();
}

}  // namespace shb
