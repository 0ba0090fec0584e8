// inflate.cuh — raw DEFLATE (RFC 1951) decoder as a device function: one WARP inflates one independent stream.
//
// Why it exists: blocked gzip (BGZF — what bgzip, htslib and the sequencers' own converters write: independent gzip members of
// at most 64 KiB of text, each carrying its compressed size in a 'BC' extra field) is the one compressed container whose
// members the host can delimit without decoding anything.  For such input only the compressed bytes cross PCIe (a third to a
// quarter of the text) and every member is inflated by its own warp into the text buffer the record splitter reads — the
// host never touches the text (the reference inflates on one goroutine: xopen -> pgzip, go.mod:20-31).
// Any other gzip stream keeps the host-side parallel feeder (csrc/host/pgzip.hpp), which needs the marker scheme to find
// block starts.
//
// How the warp is used.  Huffman decoding is one dependent chain (bits -> table entry -> shift), so all 32 lanes run it as
// identical replicas: same registers, same shared-memory reads (broadcasts), and only idempotent same-value stores.  That costs
// nothing (a warp instruction takes one issue slot whatever the lane count) and it leaves every lane with the decoder state,
// so the parts that ARE parallel need no hand-over:
//   * a match (length, distance) is copied by the lanes side by side, 32 bytes per step, straight in the text buffer (global
//     memory; a BGZF member never reaches behind its own first byte).  A short copy is carried into the next turn of the decode
//     loop — load at the top, store at the bottom — so the load's L2 round trip runs under the decoding of the next symbol;
//   * the literal/length table holds up to three literals per 10-bit look-up (DNA text: A/C/G/T get 2-3-bit codes), written by
//     lanes 0..2 at once; the table is built lane-parallel from the per-code table;
//   * the CRC-32 of the finished text is computed by 32 lanes on 32 segments, combined with the x^(8n) mod P operator (the
//     crc32_combine identity).
// Only the tables live in shared memory (10 KiB per warp), so ~20 members are resident per SM and the schedulers interleave
// their dependent chains (a first version kept the 64 KiB window in shared memory: 3 warps per SM, 12 % issue utilisation).
// Lanes are replicas, not lock-stepped by assumption: every cross-lane dependency sits behind a __syncwarp().
//
// Written as plain C++ (lane, nl = lane index and lane count) so that tests/emul compiles it for the host with nl = 1 and
// checks it against zlib on thousands of streams without a GPU; the CRC part is emulated lane by lane.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace shb {

enum : int {
    INF_OK = 0,
    INF_ERR_INPUT = 1,      // ran out of input
    INF_ERR_OUTPUT = 2,     // more output than the member announced (ISIZE)
    INF_ERR_BLOCK = 3,      // reserved block type
    INF_ERR_STORED = 4,     // stored block length check failed
    INF_ERR_TABLE = 5,      // over-subscribed or incomplete code lengths
    INF_ERR_SYMBOL = 6,     // invalid literal/length or distance symbol
    INF_ERR_DISTANCE = 7,   // distance reaches before the start of the member
    INF_ERR_SIZE = 8,       // the stream ended before ISIZE bytes were produced
    INF_ERR_CRC = 9         // CRC-32 of the text differs from the member's trailer
};

constexpr int INF_MAXBITS = 15, INF_MAXLCODES = 286, INF_MAXDCODES = 30, INF_FIXLCODES = 288;
constexpr int INF_LBITS = 10, INF_DBITS = 8, INF_CBITS = 7;
constexpr uint32_t INF_WINDOW = 65536;   // most text one BGZF member holds
#ifdef __CUDA_ARCH__
#define INF_OPAQUE(p) asm volatile("" : "+l"(p))        // keep the pointer in registers instead of re-deriving it from the kernel arguments
#else
#define INF_OPAQUE(p) (void)(p)
#endif
constexpr uint32_t INF_COPY_LANES = 32;  // lanes that copy a match side by side (the warp; the host emulation keeps the same case split)

// literal/length table entry: bits 0-3 = code bits to drop, bits 4-6 = kind, bits 8-31 = payload
//   kind 1..3: that many literals (payload bytes in output order)
//   kind 4: a length symbol: bits 8-16 = base length, bits 17-19 = number of extra bits
//   kind 5: end of block;  kind 6: a symbol that does not exist (286, 287)
//   kind 0: code longer than INF_LBITS bits, take the bit-by-bit path
// distance table entry: bits 0-3 = code bits, bits 4-7 = number of extra bits, bits 8-23 = base distance; 0 = bit-by-bit path
struct InflateScratch {
    uint32_t multi[1 << INF_LBITS];
    union {
        uint16_t single[1 << INF_LBITS];   // (symbol << 4) | code length; while `multi` is built, and for the code-length code of a header
        uint32_t crc[4][256];              // slice-by-4 tables; only after the last block
    } u;
    uint32_t fastd[1 << INF_DBITS];
    uint16_t lcount[INF_MAXBITS + 1], lsym[INF_FIXLCODES], dcount[INF_MAXBITS + 1], dsym[INF_MAXDCODES + 2];
    uint8_t lengths[INF_MAXLCODES + INF_MAXDCODES + 4];
};

// base | extra bits << 16, indexed by symbol - 256 (lengths) or by distance symbol; 0 = invalid symbol
static __constant__ uint32_t k_inf_len[32] = {
    0, 3, 4, 5, 6, 7, 8, 9, 10, 11 | 1 << 16, 13 | 1 << 16, 15 | 1 << 16, 17 | 1 << 16, 19 | 2 << 16, 23 | 2 << 16, 27 | 2 << 16,
    31 | 2 << 16, 35 | 3 << 16, 43 | 3 << 16, 51 | 3 << 16, 59 | 3 << 16, 67 | 4 << 16, 83 | 4 << 16, 99 | 4 << 16, 115 | 4 << 16,
    131 | 5 << 16, 163 | 5 << 16, 195 | 5 << 16, 227 | 5 << 16, 258, 0, 0};
static __constant__ uint32_t k_inf_dist[32] = {
    1, 2, 3, 4, 5 | 1 << 16, 7 | 1 << 16, 9 | 2 << 16, 13 | 2 << 16, 17 | 3 << 16, 25 | 3 << 16, 33 | 4 << 16, 49 | 4 << 16,
    65 | 5 << 16, 97 | 5 << 16, 129 | 6 << 16, 193 | 6 << 16, 257 | 7 << 16, 385 | 7 << 16, 513 | 8 << 16, 769 | 8 << 16,
    1025 | 9 << 16, 1537 | 9 << 16, 2049 | 10 << 16, 3073 | 10 << 16, 4097 | 11 << 16, 6145 | 11 << 16, 8193 | 12 << 16,
    12289 | 12 << 16, 16385 | 13 << 16, 24577 | 13 << 16, 0, 0};
// x^(2^k) mod P of the (reflected) CRC-32 polynomial, k = 0..31
static __constant__ uint32_t k_crc_x2n[32] = {
    0x40000000u, 0x20000000u, 0x08000000u, 0x00800000u, 0x00008000u, 0xedb88320u, 0xb1e6b092u, 0xa06a2517u, 0xed627daeu,
    0x88d14467u, 0xd7bbfe6au, 0xec447f11u, 0x8e7ea170u, 0x6427800eu, 0x4d47bae0u, 0x09fe548fu, 0x83852d0fu, 0x30362f1au,
    0x7b5a9cc3u, 0x31fec169u, 0x9fec022au, 0x6c8dedc4u, 0x15d6874du, 0x5fde7a4eu, 0xbad90e37u, 0x2e4e5eefu, 0x4eaba214u,
    0xa8a472c0u, 0x429a969eu, 0x148d302au, 0xc40ba6d0u, 0xc4e22c3cu};

// ---- bit reader: aligned 32-bit words of the compressed stream, one word fetched ahead of its use ----
struct InfBits {
    const uint32_t *w;       // 4-byte-aligned base at or before the first byte read
    uint32_t widx, nwords;   // index of `next`; words that hold stream bytes
    uint32_t next;           // w[widx], already loaded
    uint64_t buf;
    int cnt;                 // valid bits in buf
    uint32_t base_bits;      // stream position of bit 0 of w[0], plus 2^31 (it may lie up to 24 bits before the stream)
    uint32_t limit_bits;     // 8 * stream length
};

__device__ __forceinline__ uint32_t inf_word(const InfBits &b, uint32_t i) { return i < b.nwords ? __ldg(b.w + i) : 0u; }

// start reading at byte `at` of src[0, src_len)
__device__ __forceinline__ void inf_init(InfBits &b, const uint8_t *src, uint32_t src_len, uint32_t at)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(src + at);
    const uint32_t skip = (uint32_t)(a & 3u);
    b.w = reinterpret_cast<const uint32_t *>(a - skip);
    const uint32_t rest = src_len > at ? src_len - at : 0u;
    b.nwords = (skip + rest + 3u) >> 2;
    b.buf = (uint64_t)(inf_word(b, 0) >> (8u * skip));
    b.cnt = 32 - 8 * (int)skip;
    b.widx = 1;
    b.next = inf_word(b, 1);
    b.base_bits = 0x80000000u + 8u * at - 8u * skip;
    b.limit_bits = 8u * src_len;
}
// at least 32 valid bits afterwards (past the end of the stream: zeros, and inf_overrun() turns true once they are consumed)
__device__ __forceinline__ void inf_refill(InfBits &b)
{
    if (b.cnt <= 32) {
        b.buf |= (uint64_t)b.next << b.cnt;
        b.cnt += 32;
        b.widx++;
        b.next = inf_word(b, b.widx);
    }
}
__device__ __forceinline__ uint32_t inf_peek(const InfBits &b, int n) { return (uint32_t)b.buf & ((1u << n) - 1u); }
__device__ __forceinline__ void inf_drop(InfBits &b, int n) { b.buf >>= n; b.cnt -= n; }
__device__ __forceinline__ uint32_t inf_take(InfBits &b, int n)   // n <= 16, after a refill
{
    const uint32_t v = inf_peek(b, n);
    inf_drop(b, n);
    return v;
}
// stream bits consumed so far (bits in buf are not consumed yet)
__device__ __forceinline__ uint32_t inf_used_bits(const InfBits &b) { return b.base_bits + 32u * b.widx - (uint32_t)b.cnt - 0x80000000u; }
__device__ __forceinline__ bool inf_overrun(const InfBits &b) { return inf_used_bits(b) > b.limit_bits; }

// canonical Huffman code from code lengths: count[len], symbol[] sorted by (length, symbol).
// Returns < 0 over-subscribed, 0 complete, > 0 incomplete.  Replica-safe: counts are kept privately, stores are same-value.
__device__ __forceinline__ int inf_construct(uint16_t *count, uint16_t *symbol, const uint8_t *length, int n)
{
    uint16_t cnt[INF_MAXBITS + 1], offs[INF_MAXBITS + 1];
    for (int len = 0; len <= INF_MAXBITS; len++) cnt[len] = 0;
    for (int s = 0; s < n; s++) cnt[length[s]]++;
    for (int len = 0; len <= INF_MAXBITS; len++) count[len] = cnt[len];
    if (cnt[0] == n) return 0;   // no codes: complete, but decoding will fail
    int left = 1;
    for (int len = 1; len <= INF_MAXBITS; len++) {
        left <<= 1;
        left -= cnt[len];
        if (left < 0) return left;
    }
    offs[1] = 0;
    for (int len = 1; len < INF_MAXBITS; len++) offs[len + 1] = (uint16_t)(offs[len] + cnt[len]);
    for (int s = 0; s < n; s++)
        if (length[s] != 0) symbol[offs[length[s]]++] = (uint16_t)s;
    return left;
}

// bit-by-bit decode against count[] / symbol[] (codes longer than the direct tables; needs 15 readable bits)
__device__ __forceinline__ int inf_decode_slow(InfBits &b, const uint16_t *count, const uint16_t *symbol)
{
    int code = 0, first = 0, index = 0;
    uint32_t bits = inf_peek(b, INF_MAXBITS);
    for (int len = 1; len <= INF_MAXBITS; len++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[len];
        if (code - c < first) {
            inf_drop(b, len);
            return symbol[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// direct table of `bits` index bits: entry = (symbol << 4) | code length for codes of <= bits bits, else 0.
// The symbols are walked by every lane, the copies of one symbol's entry are spread over the lanes.
__device__ __forceinline__ void inf_build_direct(const uint16_t *count, const uint16_t *symbol, uint16_t *tab, int bits,
                                                 uint32_t lane, uint32_t nl)
{
    const uint32_t size = 1u << bits;
    for (uint32_t i = lane; i < size; i += nl) tab[i] = 0;
    __syncwarp();
    uint32_t code = 0, index = 0;
    for (int len = 1; len <= bits; len++) {
        const uint32_t c = count[len];
        for (uint32_t k = 0; k < c; k++) {
            const uint32_t sym = symbol[index + k];
            const uint32_t r = __brev(code + k) >> (32 - len);    // the code's bits arrive LSB first in the stream
            const uint16_t e = (uint16_t)((sym << 4) | (uint32_t)len);
            for (uint32_t fill = r + (lane << len); fill < size; fill += nl << len) tab[fill] = e;
        }
        index += c;
        code = (code + c) << 1;
    }
    __syncwarp();
}

// distance look-up: base and extra-bit count baked into the entry
__device__ __forceinline__ void inf_build_dist(const uint16_t *count, const uint16_t *symbol, uint32_t *tab, uint32_t lane, uint32_t nl)
{
    const uint32_t size = 1u << INF_DBITS;
    for (uint32_t i = lane; i < size; i += nl) tab[i] = 0;
    __syncwarp();
    uint32_t code = 0, index = 0;
    for (int len = 1; len <= INF_DBITS; len++) {
        const uint32_t c = count[len];
        for (uint32_t k = 0; k < c; k++) {
            const uint32_t sym = symbol[index + k];
            const uint32_t di = k_inf_dist[sym & 31u];
            const uint32_t r = __brev(code + k) >> (32 - len);
            // a symbol that does not exist (30, 31) keeps entry 0: the bit-by-bit path reports it
            const uint32_t e = di ? ((uint32_t)len | ((di >> 16) << 4) | ((di & 0xffffu) << 8)) : 0u;
            for (uint32_t fill = r + (lane << len); fill < size; fill += nl << len) tab[fill] = e;
        }
        index += c;
        code = (code + c) << 1;
    }
    __syncwarp();
}

// literal/length look-up with up to three literals per entry, from the per-code table u.single
__device__ __forceinline__ void inf_build_multi(InflateScratch &s, uint32_t lane, uint32_t nl)
{
    const uint32_t size = 1u << INF_LBITS;
    for (uint32_t i = lane; i < size; i += nl) {
        const uint32_t e1 = s.u.single[i];
        uint32_t out = 0;
        if (e1) {
            const uint32_t l1 = e1 & 15u, s1 = e1 >> 4;
            if (s1 == 256u) {
                out = l1 | (5u << 4);
            } else if (s1 > 256u) {
                const uint32_t li = k_inf_len[(s1 - 256u) & 31u];
                out = li ? (l1 | (4u << 4) | ((li & 0xffffu) << 8) | ((li >> 16) << 17)) : (l1 | (6u << 4));
            } else {
                uint32_t n = 1, total = l1, lits = s1 << 8, rem = i >> l1, avail = INF_LBITS - l1;
                const uint32_t e2 = s.u.single[rem];
                if (e2 && (e2 & 15u) <= avail && (e2 >> 4) < 256u) {
                    const uint32_t l2 = e2 & 15u;
                    n = 2; total += l2; lits |= (e2 >> 4) << 16; rem >>= l2; avail -= l2;
                    const uint32_t e3 = s.u.single[rem];
                    if (e3 && (e3 & 15u) <= avail && (e3 >> 4) < 256u) {
                        n = 3; total += e3 & 15u; lits |= (e3 >> 4) << 24;
                    }
                }
                out = total | (n << 4) | lits;
            }
        }
        s.multi[i] = out;
    }
    __syncwarp();
}

// the symbols of one block, tables in s; out[0, cap) is the member's text
__device__ __forceinline__ int inf_codes(InfBits &b, InflateScratch &s, uint8_t *out, uint32_t &pos, uint32_t cap, uint32_t lane, uint32_t nl)
{
    // A short copy (one step, source and destination apart) is split over two turns of the loop: its load is issued at the top
    // of the NEXT turn and its store sits at the bottom of that turn, so the L2 round trip of the load runs under the decoding
    // of the next symbol (~40 instructions) instead of stalling the warp.  have / p_len are the same in every lane.
    bool have = false;
    uint32_t p_len = 0, pv = 0;
    uint8_t *p_dst = nullptr;
    const uint8_t *p_src = nullptr;
    int rc = -1;
    uint8_t *lane0 = out + lane;                        // this lane's column of the text: one 64-bit add per access
    INF_OPAQUE(lane0);
    while (rc < 0) {
        inf_refill(b);
        if (have) {
            __syncwarp();                               // the bytes the copy reads were written by other lanes
            if (lane < p_len) pv = *p_src;
        }
        const uint32_t e = s.multi[inf_peek(b, INF_LBITS)];
        const uint32_t kind = (e >> 4) & 7u;
        uint32_t len = 0, dist = 0;                     // len != 0: a match was decoded
        if (kind == 4u || kind == 0u) {
            uint32_t extra;
            bool is_len = true;
            if (kind == 4u) {                           // a length symbol: the usual case in text that compresses
                inf_drop(b, (int)(e & 15u));
                len = (e >> 8) & 0x1ffu;
                extra = (e >> 17) & 7u;
            } else {                                    // a code longer than the table's index
                const int v = inf_decode_slow(b, s.lcount, s.lsym);
                extra = 0;
                if (v < 0) { rc = INF_ERR_SYMBOL; is_len = false; }
                else if (v < 256) {
                    is_len = false;
                    if (pos >= cap) rc = INF_ERR_OUTPUT;
                    else { if (lane == 0) out[pos] = (uint8_t)v; pos++; }
                } else if (v == 256) { rc = inf_overrun(b) ? INF_ERR_INPUT : INF_OK; is_len = false; }
                else {
                    const uint32_t li = k_inf_len[(uint32_t)(v - 256) & 31u];
                    if (li == 0) { rc = INF_ERR_SYMBOL; is_len = false; }
                    len = li & 0xffffu;
                    extra = li >> 16;
                }
            }
            if (is_len) {
                len += inf_take(b, (int)extra);
                inf_refill(b);
                uint32_t de = s.fastd[inf_peek(b, INF_DBITS)];
                if (de) {
                    inf_drop(b, (int)(de & 15u));
                } else {
                    const int ds = inf_decode_slow(b, s.dcount, s.dsym);
                    const uint32_t di = ds < 0 ? 0u : k_inf_dist[ds & 31];
                    if (di == 0) rc = INF_ERR_SYMBOL;
                    de = ((di >> 16) << 4) | ((di & 0xffffu) << 8);
                }
                dist = (de >> 8) + inf_take(b, (int)((de >> 4) & 15u));
                if (rc < 0 && (dist > pos || pos + len > cap)) rc = dist > pos ? INF_ERR_DISTANCE : INF_ERR_OUTPUT;
                if (rc >= 0) len = 0;
            } else {
                len = 0;
            }
        } else if (kind - 1u < 3u) {
            if (pos + kind > cap) rc = INF_ERR_OUTPUT;
            else {
                for (uint32_t l = lane; l < kind; l += nl) lane0[pos + l - lane] = (uint8_t)(e >> (8u + 8u * l));
                pos += kind;
                inf_drop(b, (int)(e & 15u));
            }
        } else if (kind == 5u) {
            inf_drop(b, (int)(e & 15u));
            rc = inf_overrun(b) ? INF_ERR_INPUT : INF_OK;
        } else {
            rc = INF_ERR_SYMBOL;
        }
        // bottom of the turn: the held-back copy lands, then this turn's match (if any) is copied or held back in its place
        if (have) {
            if (lane < p_len) *p_dst = (uint8_t)pv;
            have = false;
        }
        if (len) {
            if (len <= nl && dist >= len) {
                have = true;
                p_len = len;
                p_dst = lane0 + pos;
                p_src = p_dst - dist;
            } else {
                __syncwarp();                           // the bytes the copy reads were written by other lanes
                const uint8_t *from = out + pos - dist;
                if (dist >= len || dist >= INF_COPY_LANES) {
                    // a step of <= 32 bytes never reads what the same step writes
                    for (uint32_t k0 = 0; k0 < len; k0 += nl) {
                        const uint32_t k = k0 + lane;
                        if (k < len) out[pos + k] = from[k];
                        if (dist < len) __syncwarp();
                    }
                } else {
                    for (uint32_t k = lane; k < len; k += nl) out[pos + k] = from[k % dist];   // the period repeats, as the format wants
                }
            }
            pos += len;
        }
    }
    if (have) {                                         // the block ended right behind a short copy
        __syncwarp();
        if (lane < p_len) *p_dst = *p_src;
    }
    return rc;
}

// Inflate one raw deflate stream src[0, src_len) into out[0, cap) (the member's place in the text); *produced = bytes written.
// The caller compares *produced with the announced size.
__device__ __forceinline__ int inflate_raw_warp(const uint8_t *src, uint32_t src_len, uint8_t *out, uint32_t cap, uint32_t *produced,
                                                InflateScratch &s, uint32_t lane, uint32_t nl)
{
    InfBits b;
    inf_init(b, src, src_len, 0);
    uint32_t pos = 0;
    *produced = 0;
    int last;
    do {
        inf_refill(b);
        last = (int)inf_take(b, 1);
        const int type = (int)inf_take(b, 2);
        if (inf_overrun(b)) return INF_ERR_INPUT;
        if (type == 0) {
            inf_drop(b, b.cnt & 7);                     // to the next byte boundary
            inf_refill(b);
            const uint32_t len = inf_take(b, 16), nlen = inf_take(b, 16);
            if (inf_overrun(b)) return INF_ERR_INPUT;
            if ((len ^ 0xffffu) != nlen) return INF_ERR_STORED;
            const uint32_t at = inf_used_bits(b) >> 3;
            if (at + len > src_len) return INF_ERR_INPUT;
            if (pos + len > cap) return INF_ERR_OUTPUT;
            for (uint32_t k = lane; k < len; k += nl) out[pos + k] = src[at + k];
            pos += len;
            inf_init(b, src, src_len, at + len);
        } else if (type == 1 || type == 2) {
            __syncwarp();                               // nobody still decodes with the previous block's tables
            if (type == 1) {
                int i = 0;
                for (; i < 144; i++) s.lengths[i] = 8;
                for (; i < 256; i++) s.lengths[i] = 9;
                for (; i < 280; i++) s.lengths[i] = 7;
                for (; i < INF_FIXLCODES; i++) s.lengths[i] = 8;
                inf_construct(s.lcount, s.lsym, s.lengths, INF_FIXLCODES);
                inf_build_direct(s.lcount, s.lsym, s.u.single, INF_LBITS, lane, nl);
                for (i = 0; i < INF_MAXDCODES; i++) s.lengths[i] = 5;
                inf_construct(s.dcount, s.dsym, s.lengths, INF_MAXDCODES);
            } else {
                const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                inf_refill(b);
                const int nlen = (int)inf_take(b, 5) + 257, ndist = (int)inf_take(b, 5) + 1, ncode = (int)inf_take(b, 4) + 4;
                if (nlen > INF_MAXLCODES || ndist > INF_MAXDCODES) return INF_ERR_TABLE;
                int idx = 0;
                for (; idx < ncode; idx++) {
                    inf_refill(b);
                    s.lengths[order[idx]] = (uint8_t)inf_take(b, 3);
                }
                for (; idx < 19; idx++) s.lengths[order[idx]] = 0;
                if (inf_construct(s.lcount, s.lsym, s.lengths, 19) != 0) return INF_ERR_TABLE;   // the code-length code must be complete
                inf_build_direct(s.lcount, s.lsym, s.u.single, INF_CBITS, lane, nl);             // its codes have at most 7 bits
                idx = 0;
                while (idx < nlen + ndist) {
                    inf_refill(b);
                    const uint32_t ce = s.u.single[inf_peek(b, INF_CBITS)];
                    if (ce == 0) return INF_ERR_SYMBOL;
                    inf_drop(b, (int)(ce & 15u));
                    const int sym = (int)(ce >> 4);
                    if (sym < 16) {
                        s.lengths[idx++] = (uint8_t)sym;
                    } else {
                        int len = 0, rep;
                        if (sym == 16) {
                            if (idx == 0) return INF_ERR_TABLE;
                            len = s.lengths[idx - 1];
                            rep = 3 + (int)inf_take(b, 2);
                        } else if (sym == 17) {
                            rep = 3 + (int)inf_take(b, 3);
                        } else {
                            rep = 11 + (int)inf_take(b, 7);
                        }
                        if (idx + rep > nlen + ndist) return INF_ERR_TABLE;
                        while (rep--) s.lengths[idx++] = (uint8_t)len;
                    }
                    if (inf_overrun(b)) return INF_ERR_INPUT;
                }
                if (s.lengths[256] == 0) return INF_ERR_TABLE;   // no end-of-block code
                __syncwarp();                                     // every replica is done with the code-length table in u.single
                int err = inf_construct(s.lcount, s.lsym, s.lengths, nlen);
                if (err < 0 || (err > 0 && nlen - s.lcount[0] != 1)) return INF_ERR_TABLE;   // incomplete only allowed for a single code
                inf_build_direct(s.lcount, s.lsym, s.u.single, INF_LBITS, lane, nl);
                err = inf_construct(s.dcount, s.dsym, s.lengths + nlen, ndist);
                if (err < 0 || (err > 0 && ndist - s.dcount[0] != 1)) return INF_ERR_TABLE;
            }
            inf_build_multi(s, lane, nl);
            inf_build_dist(s.dcount, s.dsym, s.fastd, lane, nl);
            const int rc = inf_codes(b, s, out, pos, cap, lane, nl);
            *produced = pos;
            if (rc) return rc;
        } else {
            return INF_ERR_BLOCK;
        }
        *produced = pos;
    } while (!last);
    return INF_OK;
}

// ---- CRC-32 of the finished window, one segment per lane ----

__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b)   // a(x) * b(x) mod P, reflected representation
{
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; i--) {
        p ^= b & (0u - ((a >> i) & 1u));
        b = (b >> 1) ^ (0xedb88320u & (0u - (b & 1u)));
    }
    return p;
}
__device__ __forceinline__ uint32_t crc_x8n_modp(uint32_t n)   // x^(8n) mod P
{
    uint32_t p = 0x80000000u;   // x^0
    for (int k = 3; n; n >>= 1, k++)
        if (n & 1u) p = crc_multmodp(k_crc_x2n[k & 31], p);
    return p;
}
// slice-by-4 tables into tab (1024 words), lane-parallel
__device__ __forceinline__ void crc_build_tables(uint32_t (*tab)[256], uint32_t lane, uint32_t nl)
{
    for (uint32_t i = lane; i < 256; i += nl) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
        tab[0][i] = c;
    }
    __syncwarp();
    for (uint32_t i = lane; i < 256; i += nl) {
        uint32_t c = tab[0][i];
        for (int t = 1; t < 4; t++) {
            c = tab[0][c & 0xffu] ^ (c >> 8);
            tab[t][i] = c;
        }
    }
    __syncwarp();
}
// This lane's share of crc32(win[first, first + n)): the CRC of its segment moved to the end of the text.  win is 4-byte
// aligned; XOR over all lanes = the CRC-32 of the whole text (crc(A||B) = crc(A) * x^(8|B|) + crc(B) over GF(2)).
__device__ __forceinline__ uint32_t crc_lane_share(const uint8_t *win, uint32_t first, uint32_t n, const uint32_t (*tab)[256],
                                                   uint32_t lane, uint32_t nl)
{
    const uint32_t end = first + n;
    const uint32_t seg = (((end + nl - 1u) / nl) + 3u) & ~3u;           // bytes of window per lane, a multiple of 4
    uint32_t lo = lane * seg, hi = lo + seg;
    if (lo < first) lo = first;
    if (hi > end) hi = end;
    if (lo >= hi) return 0u;
    uint32_t c = 0xffffffffu, k = lo;
    for (; k < hi && (k & 3u); k++) c = tab[0][(c ^ win[k]) & 0xffu] ^ (c >> 8);
    for (; k + 4u <= hi; k += 4u) {
        c ^= *reinterpret_cast<const uint32_t *>(win + k);
        c = tab[3][c & 0xffu] ^ tab[2][(c >> 8) & 0xffu] ^ tab[1][(c >> 16) & 0xffu] ^ tab[0][c >> 24];
    }
    for (; k < hi; k++) c = tab[0][(c ^ win[k]) & 0xffu] ^ (c >> 8);
    c ^= 0xffffffffu;
    return crc_multmodp(crc_x8n_modp(end - hi), c);
}

}  // namespace shb
