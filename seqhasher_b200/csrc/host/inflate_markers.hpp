// inflate_markers.hpp — a DEFLATE (RFC 1951) decoder that can start in the middle of a stream.
//
// A gzip stream is one long chain of back-references, so it cannot be cut and inflated in pieces by zlib:
// a piece that starts at a later block does not know the 32 KiB of output before it.  This decoder writes
// 16-bit symbols instead of bytes: 0..255 are literal bytes, 256+i means "byte i of the (still unknown) 32 KiB
// window that precedes this piece".  The output buffer is pre-filled with the 32768 window markers in front of
// the first real symbol, so a back-reference that reaches before the piece simply copies markers; when the
// preceding piece has been resolved, one table lookup per symbol turns the markers into bytes.
// (Published technique: Kerbiriou & Chikhi, "Parallel decompression of gzip-compressed files and random
// access to DNA sequences", 2019.)  It feeds pgzip.hpp, the parallel replacement for the pgzip/xopen reader
// the reference puts in front of its FASTA/FASTQ parser (go.mod:20-31).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>

namespace host {
namespace pgz {

constexpr uint32_t WINDOW = 32768;
constexpr size_t INPUT_PAD = 1024;   // zero bytes the caller keeps after the compressed data

struct BitReader {
    const uint8_t *p = nullptr;
    uint64_t bitpos = 0;
    uint64_t end_bits = 0;   // number of valid bits
    inline uint64_t peek() const
    {
        uint64_t v;
        memcpy(&v, p + (bitpos >> 3), 8);
        return v >> (bitpos & 7);   // at least 57 valid bits
    }
    inline uint32_t get(unsigned n)
    {
        uint32_t v = (uint32_t)(peek() & ((1ull << n) - 1ull));
        bitpos += n;
        return v;
    }
    inline bool overrun() const { return bitpos > end_bits; }
};

// Canonical Huffman decoding table: primary lookup of PB bits, second level for the longer codes.
// Entries are pre-digested so the decode loop needs no second table for lengths and distances:
//   bits 0-4  code length in bits (link entries: PB)      bits 8-12  extra-bit count (link entries: sub-table bits)
//   bits 13-15 kind                                       bits 16-31 literal / base value / sub-table offset
enum Kind : uint32_t { K_INVALID = 0, K_LITERAL = 1, K_BASE = 2, K_EOB = 3, K_LINK = 4 };
enum TableUse { USE_PLAIN = 0, USE_LITLEN = 1, USE_DIST = 2 };

inline const uint16_t *len_base()
{
    static const uint16_t t[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    return t;
}
inline const uint8_t *len_extra()
{
    static const uint8_t t[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    return t;
}
inline const uint16_t *dist_base()
{
    static const uint16_t t[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    return t;
}
inline const uint8_t *dist_extra()
{
    static const uint8_t t[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    return t;
}

struct Huff {
    static constexpr int MAX_PB = 11;
    uint32_t prim[1 << MAX_PB];
    uint32_t sub[1 << 14];   // generous: <= 288 long codes x 2^(15-PB) entries each
    int max_len = 0;
    int pb = 10;

    static inline uint32_t make_entry(int sym, int len, TableUse use)
    {
        if (use == USE_LITLEN) {
            if (sym < 256) return ((uint32_t)sym << 16) | (K_LITERAL << 13) | (uint32_t)len;
            if (sym == 256) return (K_EOB << 13) | (uint32_t)len;
            if (sym > 285) return (K_INVALID << 13) | (uint32_t)len;
            return ((uint32_t)len_base()[sym - 257] << 16) | (K_BASE << 13) | ((uint32_t)len_extra()[sym - 257] << 8) | (uint32_t)len;
        }
        if (use == USE_DIST) {
            if (sym > 29) return (K_INVALID << 13) | (uint32_t)len;
            return ((uint32_t)dist_base()[sym] << 16) | (K_BASE << 13) | ((uint32_t)dist_extra()[sym] << 8) | (uint32_t)len;
        }
        return ((uint32_t)sym << 16) | (K_LITERAL << 13) | (uint32_t)len;
    }

    // returns 0 ok, -1 over-subscribed, 1 incomplete (table still usable: holes decode as K_INVALID; with
    // need_complete the table is not even filled — the block-start search rejects most candidates here)
    int build(const uint8_t *lens, int n, TableUse use, int primary_bits, bool need_complete = false)
    {
        pb = primary_bits;
        int count[16] = {0};
        for (int i = 0; i < n; i++) count[lens[i]]++;
        count[0] = 0;
        max_len = 0;
        for (int l = 15; l >= 1; l--) if (count[l]) { max_len = l; break; }
        int left = 1;
        for (int l = 1; l <= 15; l++) {
            left = left * 2 - count[l];
            if (left < 0) return -1;
        }
        if (need_complete && (left > 0 || max_len == 0)) return 1;
        memset(prim, 0, sizeof(uint32_t) << pb);
        if (max_len == 0) return 1;
        uint32_t next_code[16];
        uint32_t code = 0;
        for (int l = 1; l <= 15; l++) {
            code = (code + (uint32_t)count[l - 1]) << 1;
            next_code[l] = code;
        }
        const uint32_t pmask = (1u << pb) - 1u;
        uint8_t sub_bits[1 << MAX_PB];
        const bool any_long = max_len > pb;
        if (any_long) memset(sub_bits, 0, (size_t)1 << pb);
        uint32_t codes[320];
        for (int s = 0; s < n; s++) {
            int l = lens[s];
            if (!l) continue;
            uint32_t c = next_code[l]++;
            uint32_t r = 0;
            for (int b = 0; b < l; b++) r |= ((c >> b) & 1u) << (l - 1 - b);
            codes[s] = r;
            if (l > pb) {
                uint32_t pre = r & pmask;
                if ((uint8_t)(l - pb) > sub_bits[pre]) sub_bits[pre] = (uint8_t)(l - pb);
            }
        }
        uint32_t sub_used = 0;
        if (any_long) {
            for (uint32_t pre = 0; pre <= pmask; pre++)
                if (sub_bits[pre]) {
                    prim[pre] = (sub_used << 16) | (K_LINK << 13) | ((uint32_t)sub_bits[pre] << 8) | (uint32_t)pb;
                    memset(sub + sub_used, 0, sizeof(uint32_t) << sub_bits[pre]);
                    sub_used += 1u << sub_bits[pre];
                }
        }
        for (int s = 0; s < n; s++) {
            int l = lens[s];
            if (!l) continue;
            uint32_t r = codes[s];
            if (l <= pb) {
                uint32_t entry = make_entry(s, l, use);
                for (uint32_t j = r; j <= pmask; j += 1u << l) prim[j] = entry;
            } else {
                uint32_t pre = r & pmask;
                uint32_t base = prim[pre] >> 16, sb = (prim[pre] >> 8) & 31u;
                uint32_t entry = make_entry(s, l - pb, use);   // the link already consumed pb bits
                for (uint32_t j = r >> pb; j < (1u << sb); j += 1u << (l - pb)) sub[base + j] = entry;
            }
        }
        return left > 0 ? 1 : 0;
    }
    // plain symbol decode through the BitReader (code-length code only); -1 on an unassigned code
    inline int decode(BitReader &br) const
    {
        uint64_t bits = br.peek();
        uint32_t e = prim[bits & ((1u << pb) - 1u)];
        if ((e >> 13 & 7u) == K_LINK) {
            br.bitpos += pb;
            bits >>= pb;
            e = sub[(e >> 16) + (bits & ((1u << ((e >> 8) & 31u)) - 1u))];
        }
        if ((e >> 13 & 7u) == K_INVALID) return -1;
        br.bitpos += e & 31u;
        return (int)(e >> 16);
    }
};

// Growable 16-bit symbol buffer; the first WINDOW entries are the window markers.
struct SymBuf {
    uint16_t *d = nullptr;
    size_t n = 0, cap = 0;
    SymBuf() {}
    SymBuf(const SymBuf &) = delete;
    SymBuf &operator=(const SymBuf &) = delete;
    ~SymBuf() { free(d); }
    void reset(size_t expect)
    {
        if (cap < expect + WINDOW + 1024) {
            free(d);
            cap = expect + WINDOW + 1024;
            d = (uint16_t *)malloc(cap * sizeof(uint16_t));
            if (!d) throw std::bad_alloc();
            for (uint32_t i = 0; i < WINDOW; i++) d[i] = (uint16_t)(256 + i);
        }
        n = WINDOW;
    }
    inline void reserve(size_t extra)
    {
        if (n + extra > cap) {
            size_t ncap = cap * 2 + extra;
            uint16_t *nd = (uint16_t *)realloc(d, ncap * sizeof(uint16_t));
            if (!nd) throw std::bad_alloc();
            d = nd;
            cap = ncap;
        }
    }
    size_t produced() const { return n - WINDOW; }
    const uint16_t *symbols() const { return d + WINDOW; }
};

enum Status { BLOCK_OK = 0, BLOCK_FINAL = 1, NEED_INPUT = 2, BAD_DATA = 3 };
// MODE_ANY: everything zlib accepts.  The two search modes only accept what can be told from noise: a
// dynamic block with strictly complete codes — non-final (a block start in mid-stream) or any (first block of
// a gzip member).
enum Mode { MODE_ANY = 0, MODE_BLOCK_START = 1, MODE_MEMBER_START = 2 };

struct BlockDecoder {
    Huff lit, dist;
    Huff pre;

    // Dynamic-block header (after the 3 type bits), with zlib's acceptance rules: complete code-length
    // code, an end-of-block symbol, literal/length and distance codes complete or a single 1-bit code
    // (distance: or none at all).  `strict` (block-start search) also refuses the single-code literal set.
    bool read_dynamic_tables(BitReader &br, bool strict)
    {
        uint32_t hlit = br.get(5) + 257, hdist = br.get(5) + 1, hclen = br.get(4) + 4;
        if (hlit > 286 || hdist > 30) return false;
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t pl[19] = {0};
        for (uint32_t i = 0; i < hclen; i++) pl[order[i]] = (uint8_t)br.get(3);
        if (pre.build(pl, 19, USE_PLAIN, 7, true) != 0) return false;
        uint8_t lens[320];
        uint32_t i = 0, total = hlit + hdist;
        while (i < total) {
            if (br.overrun()) return false;
            int s = pre.decode(br);
            if (s < 0) return false;
            if (s < 16) { lens[i++] = (uint8_t)s; continue; }
            uint32_t rep;
            uint8_t v = 0;
            if (s == 16) {
                if (i == 0) return false;
                v = lens[i - 1];
                rep = 3 + br.get(2);
            } else if (s == 17) rep = 3 + br.get(3);
            else rep = 11 + br.get(7);
            if (i + rep > total) return false;
            while (rep--) lens[i++] = v;
        }
        if (lens[256] == 0) return false;
        int lrc = lit.build(lens, (int)hlit, USE_LITLEN, 11, strict);
        if (lrc < 0 || (lrc > 0 && (strict || lit.max_len != 1))) return false;
        int rc = dist.build(lens + hlit, (int)hdist, USE_DIST, 8);
        if (rc < 0) return false;
        if (rc > 0 && dist.max_len > 1) return false;   // incomplete only as "no distances" or one 1-bit code
        return true;
    }
    void fixed_tables()
    {
        uint8_t lens[288];
        for (int i = 0; i < 144; i++) lens[i] = 8;
        for (int i = 144; i < 256; i++) lens[i] = 9;
        for (int i = 256; i < 280; i++) lens[i] = 7;
        for (int i = 280; i < 288; i++) lens[i] = 8;
        lit.build(lens, 288, USE_LITLEN, 9);
        uint8_t dl[32];
        for (int i = 0; i < 32; i++) dl[i] = 5;
        dist.build(dl, 32, USE_DIST, 5);   // codes 30 and 31 exist in the code space but are invalid
    }

    // An invalid code is only evidence of bad data when every bit it was decoded from is real input; within
    // 64 bits of the end of the buffer it may be the zero padding speaking, so the caller is asked for more.
    static Status bad(const BitReader &br) { return br.bitpos + 64 > br.end_bits ? NEED_INPUT : BAD_DATA; }

    // One whole block starting at br.bitpos (the 3-bit header included).  On NEED_INPUT / BAD_DATA the caller
    // restores br.bitpos and out.n itself.
    Status block(BitReader &br, SymBuf &out, Mode mode = MODE_ANY)
    {
        if (br.bitpos + 3 > br.end_bits) return NEED_INPUT;
        uint32_t hdr = br.get(3);
        const bool final = hdr & 1u;
        const uint32_t type = hdr >> 1;
        if (mode != MODE_ANY && (type != 2 || (final && mode == MODE_BLOCK_START))) return BAD_DATA;
        if (type == 3) return BAD_DATA;
        if (type == 0) {
            br.bitpos = (br.bitpos + 7) & ~7ull;
            if (br.bitpos + 32 > br.end_bits) return NEED_INPUT;
            uint32_t len = br.get(16), nlen = br.get(16);
            if ((len ^ nlen) != 0xffffu) return BAD_DATA;
            if (br.bitpos + 8ull * len > br.end_bits) return NEED_INPUT;
            out.reserve(len);
            const uint8_t *s = br.p + (br.bitpos >> 3);
            uint16_t *o = out.d + out.n;
            for (uint32_t i = 0; i < len; i++) o[i] = s[i];
            out.n += len;
            br.bitpos += 8ull * len;
            return final ? BLOCK_FINAL : BLOCK_OK;
        }
        if (type == 1) fixed_tables();
        else {
            if (!read_dynamic_tables(br, mode != MODE_ANY)) return bad(br);
        }
        // Huffman-coded data.  The bit buffer lives in registers: refilled to >= 56 bits once per iteration,
        // which covers a length + distance pair (<= 48 bits) or up to three literals.
        const uint8_t *in = br.p + (br.bitpos >> 3);
        const uint8_t *const in_limit = br.p + (br.end_bits >> 3) + 16;   // decoding zero padding: stop soon
        uint64_t bitbuf = 0;
        unsigned bitcnt = 0;
#define PGZ_REFILL()                                         \
    do {                                                     \
        uint64_t w_;                                         \
        memcpy(&w_, in, 8);                                  \
        bitbuf |= w_ << bitcnt;                              \
        in += (63 - bitcnt) >> 3;                            \
        bitcnt |= 56;                                        \
    } while (0)
#define PGZ_LOOKUP(e_)                                                                        \
    do {                                                                                      \
        e_ = lit.prim[bitbuf & lmask];                                                        \
        if (((e_ >> 13) & 7u) == K_LINK) {                                                    \
            bitbuf >>= lpb;                                                                   \
            bitcnt -= lpb;                                                                    \
            e_ = lit.sub[(e_ >> 16) + (bitbuf & ((1u << ((e_ >> 8) & 31u)) - 1u))];          \
        }                                                                                     \
    } while (0)
        PGZ_REFILL();
        {
            unsigned skip = (unsigned)(br.bitpos & 7);
            bitbuf >>= skip;
            bitcnt -= skip;
        }
        const unsigned lpb = (unsigned)lit.pb, dpb = (unsigned)dist.pb;
        const uint32_t lmask = (1u << lpb) - 1u, dmask = (1u << dpb) - 1u;
        Status result = BLOCK_OK;
        size_t n = out.n;
        for (;;) {
            if (in > in_limit) { result = NEED_INPUT; break; }
            if (n + 3 + 258 + 8 > out.cap) {
                out.n = n;
                out.reserve(3 + 258 + 8 + 65536);
            }
            uint16_t *o = out.d;
            PGZ_REFILL();
            uint32_t e;
            PGZ_LOOKUP(e);
            uint32_t kind = (e >> 13) & 7u;
            if (kind == K_LITERAL) {
                bitbuf >>= (e & 31u);
                bitcnt -= (e & 31u);
                o[n++] = (uint16_t)(e >> 16);
                // second and third literal on the same refill (each code <= 15 bits, >= 26 bits still valid);
                // anything else is looked up again after the next refill
                uint32_t e2 = lit.prim[bitbuf & lmask];
                if (((e2 >> 13) & 7u) != K_LITERAL) continue;
                bitbuf >>= (e2 & 31u);
                bitcnt -= (e2 & 31u);
                o[n++] = (uint16_t)(e2 >> 16);
                e2 = lit.prim[bitbuf & lmask];
                if (((e2 >> 13) & 7u) != K_LITERAL) continue;
                bitbuf >>= (e2 & 31u);
                bitcnt -= (e2 & 31u);
                o[n++] = (uint16_t)(e2 >> 16);
                continue;
            }
            if (kind == K_EOB) {
                bitbuf >>= (e & 31u);
                bitcnt -= (e & 31u);
                break;
            }
            if (kind != K_BASE) { result = BAD_DATA; break; }
            unsigned cl = e & 31u, xb = (e >> 8) & 31u;
            uint32_t len = (e >> 16) + (uint32_t)((bitbuf >> cl) & ((1u << xb) - 1u));
            bitbuf >>= cl + xb;
            bitcnt -= cl + xb;
            uint32_t de = dist.prim[bitbuf & dmask];
            if (((de >> 13) & 7u) == K_LINK) {
                bitbuf >>= dpb;
                bitcnt -= dpb;
                de = dist.sub[(de >> 16) + (bitbuf & ((1u << ((de >> 8) & 31u)) - 1u))];
            }
            if (((de >> 13) & 7u) != K_BASE) { result = BAD_DATA; break; }
            cl = de & 31u;
            xb = (de >> 8) & 31u;
            const uint32_t d = (de >> 16) + (uint32_t)((bitbuf >> cl) & ((1u << xb) - 1u));
            bitbuf >>= cl + xb;
            bitcnt -= cl + xb;
            // n >= WINDOW always and d <= 32768: the source never underflows the buffer
            const uint16_t *src = o + n - d;
            uint16_t *dst = o + n;
            if (d >= 8) {   // 8 symbols (16 bytes) at a time; may write up to 7 symbols past the match (reserved)
                for (uint32_t i = 0; i < len; i += 8) memcpy(dst + i, src + i, 16);
            } else {
                for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
            }
            n += len;
        }
#undef PGZ_REFILL
#undef PGZ_LOOKUP
        out.n = n;
        br.bitpos = (uint64_t)(in - br.p) * 8 - bitcnt;
        if (result != BLOCK_OK) return result == NEED_INPUT ? NEED_INPUT : bad(br);
        if (br.overrun()) return NEED_INPUT;
        return final ? BLOCK_FINAL : BLOCK_OK;
    }
};

}  // namespace pgz
}  // namespace host
