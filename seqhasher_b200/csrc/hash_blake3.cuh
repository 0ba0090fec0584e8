// hash_blake3.cuh — BLAKE3 (default hash mode, 32-byte output) device hasher.
//
// Replaces blake3.Sum256 (seqhasher.go:328-330).  Thread-per-record form: chunks are compressed
// in order and their chaining values merged with the lazy binary-counter stack (a CV is merged
// when the count of finished chunks gains a trailing zero bit), which yields the specification's
// tree (left subtree = largest power of two of chunks strictly below the total); ROOT is set on
// the single last compression.  The long-record kernels reuse b3_compress / b3_chunk_cv with one
// chunk per lane.
#pragma once
#include "common.cuh"

namespace shb {

enum : uint32_t { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };

#define B3_IV0 0x6A09E667u
#define B3_IV1 0xBB67AE85u
#define B3_IV2 0x3C6EF372u
#define B3_IV3 0xA54FF53Au
#define B3_IV4 0x510E527Fu
#define B3_IV5 0x9B05688Cu
#define B3_IV6 0x1F83D9ABu
#define B3_IV7 0x5BE0CD19u

__device__ __forceinline__ void b3_g(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, uint32_t mx, uint32_t my)
{
    // the message-word adds are forced onto the FMA pipe (fma_add): BLAKE3 is ALU-pipe bound on sm_100
    a = fma_add(a + b, mx); d = rotr32(d ^ a, 16);
    c = c + d;              b = rotr32(b ^ c, 12);
    a = fma_add(a + b, my); d = rotr32(d ^ a, 8);
    c = c + d;              b = rotr32(b ^ c, 7);
}

// message word schedule: round r uses m[SCHED[r][i]] (the spec's permutation applied r times)
__device__ __forceinline__ constexpr int b3_sched(int r, int i)
{
    constexpr int P[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
    int idx = i;
    for (int k = 0; k < r; k++) idx = P[idx];
    return idx;
}

// cv <- first 8 words of compress(cv, m, counter, block_len, flags)
__device__ __forceinline__ void b3_compress(uint32_t (&cv)[8], const uint32_t (&m)[16], uint32_t counter_lo,
                                            uint32_t counter_hi, uint32_t block_len, uint32_t flags)
{
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = B3_IV0, s9 = B3_IV1, s10 = B3_IV2, s11 = B3_IV3;
    uint32_t s12 = counter_lo, s13 = counter_hi, s14 = block_len, s15 = flags;
#pragma unroll
    for (int r = 0; r < 7; r++) {
        b3_g(s0, s4, s8, s12, m[b3_sched(r, 0)], m[b3_sched(r, 1)]);
        b3_g(s1, s5, s9, s13, m[b3_sched(r, 2)], m[b3_sched(r, 3)]);
        b3_g(s2, s6, s10, s14, m[b3_sched(r, 4)], m[b3_sched(r, 5)]);
        b3_g(s3, s7, s11, s15, m[b3_sched(r, 6)], m[b3_sched(r, 7)]);
        b3_g(s0, s5, s10, s15, m[b3_sched(r, 8)], m[b3_sched(r, 9)]);
        b3_g(s1, s6, s11, s12, m[b3_sched(r, 10)], m[b3_sched(r, 11)]);
        b3_g(s2, s7, s8, s13, m[b3_sched(r, 12)], m[b3_sched(r, 13)]);
        b3_g(s3, s4, s9, s14, m[b3_sched(r, 14)], m[b3_sched(r, 15)]);
    }
    cv[0] = s0 ^ s8;  cv[1] = s1 ^ s9;  cv[2] = s2 ^ s10; cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14; cv[7] = s7 ^ s15;
}

__device__ __forceinline__ void b3_set_iv(uint32_t (&cv)[8])
{
    cv[0] = B3_IV0; cv[1] = B3_IV1; cv[2] = B3_IV2; cv[3] = B3_IV3;
    cv[4] = B3_IV4; cv[5] = B3_IV5; cv[6] = B3_IV6; cv[7] = B3_IV7;
}

// chaining value of the chunk [start, start+clen) of the record, clen in 0..1024 (0 only for an
// empty record); last_flags is OR-ed into the chunk's final block (B3_ROOT for a one-chunk input)
template <class R>
__device__ __forceinline__ void b3_chunk_cv(const R &r, uint32_t start, uint32_t clen, uint32_t chunk_index,
                                            uint32_t last_flags, uint32_t (&cv)[8])
{
    b3_set_iv(cv);
    const uint32_t nblocks = clen ? (clen + 63) >> 6 : 1;
    for (uint32_t b = 0; b < nblocks; b++) {
        const uint32_t off = b << 6;
        uint32_t m[16];
        uint32_t bl = 64;
        if (off + 64 <= clen) {
#pragma unroll
            for (int i = 0; i < 16; i++) m[i] = r.u32(start + off + 4 * i);
        } else {
            bl = clen - off;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint32_t o = off + 4 * i;
                uint32_t v = 0;
                if (o < clen) {
                    v = r.u32(start + o);
                    if (o + 4 > clen) v &= (1u << (8 * (clen - o))) - 1u;
                }
                m[i] = v;
            }
        }
        uint32_t flags = (b == 0 ? B3_CHUNK_START : 0u) | (b == nblocks - 1 ? (B3_CHUNK_END | last_flags) : 0u);
        b3_compress(cv, m, chunk_index, 0u, bl, flags);
    }
}

// cv <- parent(left, cv): one shared out-of-line copy keeps the tree walk small
static __device__ __noinline__ void b3_parent(const uint32_t *left, uint32_t (&cv)[8], uint32_t extra_flags)
{
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = left[i]; m[8 + i] = cv[i]; }
    b3_set_iv(cv);
    b3_compress(cv, m, 0u, 0u, 64u, B3_PARENT | extra_flags);
}

// out: 32 bytes, 4-byte aligned
template <class R>
__device__ __forceinline__ void blake3_record(const R &r, uint32_t len, uint8_t *out)
{
    uint32_t cv[8];
    const uint32_t nchunks = len ? (len + 1023) >> 10 : 1;
    uint32_t stack[22][8];   // 2^31 bytes / 1024 = 2^21 chunks
    int depth = 0;
    for (uint32_t c = 0; c < nchunks; c++) {
        const bool last = (c == nchunks - 1);
        const uint32_t clen = last ? len - (c << 10) : 1024u;
        b3_chunk_cv(r, c << 10, clen, c, (last && nchunks == 1) ? (uint32_t)B3_ROOT : 0u, cv);
        if (!last) {
            // lazy merge: finished-chunk count gained a trailing zero => a complete subtree pair
            for (uint32_t total = c + 1; (total & 1u) == 0; total >>= 1) b3_parent(stack[--depth], cv, 0u);
#pragma unroll
            for (int i = 0; i < 8; i++) stack[depth][i] = cv[i];
            depth++;
        }
    }
    while (depth > 0) {
        depth--;
        b3_parent(stack[depth], cv, depth == 0 ? (uint32_t)B3_ROOT : 0u);
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(out);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = cv[i];
}

}  // namespace shb
