// inflate.cuh — raw DEFLATE (RFC 1951) decoder as a device function: one thread inflates one independent stream.
//
// Why it exists: blocked gzip (BGZF — what bgzip, htslib and the sequencers' own converters write: independent gzip members of
// at most 64 KiB, each carrying its compressed size in a 'BC' extra field) is the one compressed container whose members the
// host can delimit without decoding anything.  For such input the compressed bytes cross PCIe (a third to a quarter of the
// text) and every member is inflated by its own warp straight into the text buffer the record splitter reads — the host never
// touches the text (the reference inflates on one goroutine: xopen -> pgzip, go.mod:20-31).  Any other
// gzip stream keeps the host-side parallel feeder (csrc/host/pgzip.hpp), which needs the marker scheme to find block starts.
//
// The decoder is the classic canonical-Huffman one: per block a count[] / symbol[] table pair (puff-style slow path) plus a
// 9-bit direct look-up table for the literal/length code built from it (one shared-memory access decodes the common short
// codes).  State lives in an InflateScratch the caller places in shared memory (one per warp).  Written as plain C++ so that
// tests/emul compiles it for the host and checks it against zlib on thousands of streams without a GPU.
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
    INF_ERR_SIZE = 8        // the stream ended before ISIZE bytes were produced
};

constexpr int INF_MAXBITS = 15, INF_MAXLCODES = 286, INF_MAXDCODES = 30, INF_FIXLCODES = 288;
constexpr int INF_FAST_BITS = 9;

struct InflateScratch {
    uint16_t lcount[INF_MAXBITS + 1], lsym[INF_FIXLCODES];
    uint16_t dcount[INF_MAXBITS + 1], dsym[INF_MAXDCODES];
    uint16_t fast[1 << INF_FAST_BITS];      // (symbol << 4) | code length for codes of <= 9 bits, 0 = take the slow path
    uint16_t lengths[INF_MAXLCODES + INF_MAXDCODES + 2];
};

struct InfBits {
    const uint8_t *p, *end;
    uint64_t buf;
    int cnt;
};

// at least 57 valid bits afterwards; past the end of the input the stream continues with virtual zero bytes and p keeps
// counting, so that inf_overrun() can tell when more bits were consumed than the input held
__device__ __forceinline__ void inf_refill(InfBits &b)
{
    while (b.cnt <= 56) {
        const uint64_t byte = b.p < b.end ? *b.p : 0;
        b.p++;
        b.buf |= byte << b.cnt;
        b.cnt += 8;
    }
}
__device__ __forceinline__ uint32_t inf_peek(const InfBits &b, int n) { return (uint32_t)(b.buf & ((1ull << n) - 1)); }
__device__ __forceinline__ void inf_drop(InfBits &b, int n) { b.buf >>= n; b.cnt -= n; }
__device__ __forceinline__ uint32_t inf_bits(InfBits &b, int n)
{
    if (b.cnt < n) inf_refill(b);
    const uint32_t v = inf_peek(b, n);
    inf_drop(b, n);
    return v;
}
// more bits consumed than the input holds?
__device__ __forceinline__ bool inf_overrun(const InfBits &b) { return (b.p - b.end) * 8 > (long long)b.cnt; }

// canonical Huffman tables from code lengths; returns < 0 over-subscribed, 0 complete, > 0 incomplete
__device__ __forceinline__ int inf_construct(uint16_t *count, uint16_t *symbol, const uint16_t *length, int n)
{
    for (int len = 0; len <= INF_MAXBITS; len++) count[len] = 0;
    for (int s = 0; s < n; s++) count[length[s]]++;
    if (count[0] == n) return 0;   // no codes: complete, but decoding will fail
    int left = 1;
    for (int len = 1; len <= INF_MAXBITS; len++) {
        left <<= 1;
        left -= count[len];
        if (left < 0) return left;
    }
    uint16_t offs[INF_MAXBITS + 1];
    offs[1] = 0;
    for (int len = 1; len < INF_MAXBITS; len++) offs[len + 1] = offs[len] + count[len];
    for (int s = 0; s < n; s++)
        if (length[s] != 0) symbol[offs[length[s]]++] = (uint16_t)s;
    return left;
}

// slow path: one bit at a time against count[] / symbol[]
__device__ __forceinline__ int inf_decode_slow(InfBits &b, const uint16_t *count, const uint16_t *symbol)
{
    int code = 0, first = 0, index = 0;
    if (b.cnt < INF_MAXBITS) inf_refill(b);
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

// 9-bit direct table for the literal/length code: entry = (symbol << 4) | length
__device__ __forceinline__ void inf_build_fast(InflateScratch &s)
{
    for (int i = 0; i < (1 << INF_FAST_BITS); i++) s.fast[i] = 0;
    int code = 0, index = 0;
    for (int len = 1; len <= INF_FAST_BITS; len++) {
        const int c = s.lcount[len];
        for (int k = 0; k < c; k++) {
            const int sym = s.lsym[index + k];
            // the code's bits arrive LSB first in the stream: reverse `len` bits of (code + k)
            uint32_t v = (uint32_t)(code + k), r = 0;
            for (int t = 0; t < len; t++) { r = (r << 1) | (v & 1u); v >>= 1; }
            for (uint32_t fill = r; fill < (1u << INF_FAST_BITS); fill += 1u << len) s.fast[fill] = (uint16_t)((sym << 4) | len);
        }
        index += c;
        code = (code + c) << 1;
    }
}

__device__ __forceinline__ int inf_codes(InfBits &b, InflateScratch &s, uint8_t *dst, uint32_t &pos, uint32_t dst_len)
{
    const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                8193, 12289, 16385, 24577};
    const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    while (true) {
        if (b.cnt < 32) inf_refill(b);
        int sym;
        const uint16_t e = s.fast[inf_peek(b, INF_FAST_BITS)];
        if (e) { inf_drop(b, e & 15); sym = e >> 4; }
        else sym = inf_decode_slow(b, s.lcount, s.lsym);
        if (sym < 0) return INF_ERR_SYMBOL;
        if (sym < 256) {
            if (pos >= dst_len) return INF_ERR_OUTPUT;
            dst[pos++] = (uint8_t)sym;
        } else if (sym == 256) {
            return inf_overrun(b) ? INF_ERR_INPUT : INF_OK;
        } else {
            sym -= 257;
            if (sym >= 29) return INF_ERR_SYMBOL;
            if (b.cnt < 32) inf_refill(b);
            const uint32_t len = lbase[sym] + inf_bits(b, lext[sym]);
            const int ds = inf_decode_slow(b, s.dcount, s.dsym);
            if (ds < 0 || ds >= 30) return INF_ERR_SYMBOL;
            if (b.cnt < 16) inf_refill(b);
            const uint32_t dist = dbase[ds] + inf_bits(b, dext[ds]);
            if (dist > pos) return INF_ERR_DISTANCE;
            if (pos + len > dst_len) return INF_ERR_OUTPUT;
            for (uint32_t k = 0; k < len; k++, pos++) dst[pos] = dst[pos - dist];   // overlapping copies repeat, as the format wants
        }
        if (inf_overrun(b)) return INF_ERR_INPUT;
    }
}

// Inflate one raw deflate stream src[0, src_len) into dst[0, dst_len) where dst_len is the announced size (gzip ISIZE).
__device__ __forceinline__ int inflate_raw(const uint8_t *src, uint32_t src_len, uint8_t *dst, uint32_t dst_len, InflateScratch &s)
{
    InfBits b;
    b.p = src; b.end = src + src_len; b.buf = 0; b.cnt = 0;
    uint32_t pos = 0;
    int last;
    do {
        last = (int)inf_bits(b, 1);
        const int type = (int)inf_bits(b, 2);
        if (inf_overrun(b)) return INF_ERR_INPUT;
        if (type == 0) {
            inf_drop(b, b.cnt & 7);                 // to the next byte boundary
            if (b.cnt < 32) inf_refill(b);
            const uint32_t len = inf_bits(b, 16), nlen = inf_bits(b, 16);
            if (inf_overrun(b)) return INF_ERR_INPUT;
            if ((len ^ 0xffffu) != nlen) return INF_ERR_STORED;
            if (pos + len > dst_len) return INF_ERR_OUTPUT;
            for (uint32_t k = 0; k < len; k++) {
                if (b.cnt < 8) inf_refill(b);
                dst[pos++] = (uint8_t)inf_peek(b, 8);
                inf_drop(b, 8);
            }
            if (inf_overrun(b)) return INF_ERR_INPUT;
        } else if (type == 1) {
            int i = 0;
            for (; i < 144; i++) s.lengths[i] = 8;
            for (; i < 256; i++) s.lengths[i] = 9;
            for (; i < 280; i++) s.lengths[i] = 7;
            for (; i < INF_FIXLCODES; i++) s.lengths[i] = 8;
            inf_construct(s.lcount, s.lsym, s.lengths, INF_FIXLCODES);
            for (i = 0; i < INF_MAXDCODES; i++) s.lengths[i] = 5;
            inf_construct(s.dcount, s.dsym, s.lengths, INF_MAXDCODES);
            inf_build_fast(s);
            const int rc = inf_codes(b, s, dst, pos, dst_len);
            if (rc) return rc;
        } else if (type == 2) {
            const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            if (b.cnt < 32) inf_refill(b);
            const int nlen = (int)inf_bits(b, 5) + 257, ndist = (int)inf_bits(b, 5) + 1, ncode = (int)inf_bits(b, 4) + 4;
            if (nlen > INF_MAXLCODES || ndist > INF_MAXDCODES) return INF_ERR_TABLE;
            int idx = 0;
            for (; idx < ncode; idx++) s.lengths[order[idx]] = (uint16_t)inf_bits(b, 3);
            for (; idx < 19; idx++) s.lengths[order[idx]] = 0;
            if (inf_construct(s.lcount, s.lsym, s.lengths, 19) != 0) return INF_ERR_TABLE;   // the code-length code must be complete
            idx = 0;
            while (idx < nlen + ndist) {
                int sym = inf_decode_slow(b, s.lcount, s.lsym);
                if (sym < 0) return INF_ERR_SYMBOL;
                if (sym < 16) {
                    s.lengths[idx++] = (uint16_t)sym;
                } else {
                    int len = 0, rep;
                    if (sym == 16) {
                        if (idx == 0) return INF_ERR_TABLE;
                        len = s.lengths[idx - 1];
                        rep = 3 + (int)inf_bits(b, 2);
                    } else if (sym == 17) {
                        rep = 3 + (int)inf_bits(b, 3);
                    } else {
                        rep = 11 + (int)inf_bits(b, 7);
                    }
                    if (idx + rep > nlen + ndist) return INF_ERR_TABLE;
                    while (rep--) s.lengths[idx++] = (uint16_t)len;
                }
                if (inf_overrun(b)) return INF_ERR_INPUT;
            }
            if (s.lengths[256] == 0) return INF_ERR_TABLE;   // no end-of-block code
            // the distance lengths are copied out before lsym / lcount are rebuilt over the code-length tables
            uint16_t dl[INF_MAXDCODES];
            for (int i = 0; i < ndist; i++) dl[i] = s.lengths[nlen + i];
            int err = inf_construct(s.lcount, s.lsym, s.lengths, nlen);
            if (err < 0 || (err > 0 && nlen - s.lcount[0] != 1)) return INF_ERR_TABLE;   // incomplete only allowed for a single code
            err = inf_construct(s.dcount, s.dsym, dl, ndist);
            if (err < 0 || (err > 0 && ndist - s.dcount[0] != 1)) return INF_ERR_TABLE;
            inf_build_fast(s);
            const int rc = inf_codes(b, s, dst, pos, dst_len);
            if (rc) return rc;
        } else {
            return INF_ERR_BLOCK;
        }
    } while (!last);
    return pos == dst_len ? INF_OK : INF_ERR_SIZE;
}

}  // namespace shb
