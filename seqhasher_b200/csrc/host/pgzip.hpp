// pgzip.hpp — parallel gzip reader for the compiled host (decompression feeders).
//
// The reference reads .gz input through xopen → klauspost/pgzip (go.mod:20-31), which inflates on ONE
// goroutine; config C5 (gzip'd multi-line FASTA, all 12 hashes) is bound by that single inflate long before
// the hashes matter.  Here the compressed stream is cut into T pieces per batch; every piece finds the next
// deflate block start in its region (or the next gzip member header — bgzip/pgzip multi-member files),
// inflates from there into 16-bit symbols with window markers (inflate_markers.hpp) and stops exactly on the
// start its successor found.  The pieces are then chained front to back (a piece is only trusted when its
// predecessor landed exactly on its first bit, so a mis-detected start costs time, never correctness), the
// 32 KiB windows are handed down the chain, markers are resolved and CRC-32s computed in parallel, and every
// member's CRC-32 / ISIZE trailer is verified like gzip.Reader does.
#pragma once
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "inflate_markers.hpp"
#include "source.hpp"

namespace host {
namespace pgz {

// Length of the gzip member header at p (RFC 1952), 0 = more bytes needed, -1 = not a gzip header.
inline long gzip_header_len(const uint8_t *p, size_t n)
{
    if (n < 10) return (n >= 1 && p[0] != 0x1f) || (n >= 2 && p[1] != 0x8b) || (n >= 3 && p[2] != 8) ? -1 : 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xe0)) return -1;
    const uint8_t flg = p[3];
    size_t q = 10;
    if (flg & 4) {
        if (q + 2 > n) return 0;
        size_t xlen = p[q] | (p[q + 1] << 8);
        q += 2 + xlen;
        if (q > n) return 0;
    }
    for (int field = 0; field < 2; field++)
        if (flg & (field == 0 ? 8 : 16)) {
            while (q < n && p[q]) q++;
            if (q >= n) return 0;
            q++;
        }
    if (flg & 2) {   // FHCRC: CRC-16 of the header so far (gzip.Reader checks it too)
        if (q + 2 > n) return 0;
        const uint32_t want = p[q] | (p[q + 1] << 8);
        if ((crc32(crc32(0L, Z_NULL, 0), p, (uInt)q) & 0xffffu) != want) return -1;
        q += 2;
    }
    return (long)q;
}

// Kraft sum of the code-length code of a dynamic block whose 3 header bits are at br.bitpos: three 3-bit
// lengths per table lookup.  A real header has a complete code (sum of 2^-len == 1).
inline bool precode_complete(const BitReader &br)
{
    static const struct Lut {
        uint8_t v[512];
        Lut()
        {
            for (int i = 0; i < 512; i++) {
                int a = i & 7, b = (i >> 3) & 7, c = i >> 6;
                v[i] = (uint8_t)((a ? 128 >> a : 0) + (b ? 128 >> b : 0) + (c ? 128 >> c : 0));   // <= 192
            }
        }
    } lut;
    BitReader r = br;
    const unsigned hclen = (unsigned)((r.peek() >> 13) & 15) + 4;
    r.bitpos += 17;
    uint64_t bits = r.peek() & ((1ull << (3 * hclen)) - 1ull);   // 57 valid bits = 19 lengths
    unsigned sum = 0;
    for (; bits; bits >>= 9) sum += lut.v[bits & 511];
    return sum == 128;
}

template <class F>
inline void parallel_for(unsigned n, F f)
{
    if (n <= 1) { if (n) f(0u); return; }
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> errs(n);
    for (unsigned i = 0; i < n; i++)
        th.emplace_back([&, i] { try { f(i); } catch (...) { errs[i] = std::current_exception(); } });
    for (auto &t : th) t.join();
    for (auto &e : errs) if (e) std::rethrow_exception(e);
}

struct MemberEnd { size_t out_n; uint32_t crc, isize; };

enum EndKind { END_LANDED = 0, END_INPUT = 1, END_STREAM = 2, END_BAD = 3, END_BAD_HEADER = 4, END_BUDGET = 5 };

// A piece stops at the next block boundary once it holds this many symbols (a poly-A FASTA inflates 1000:1);
// the rest of its input is carried into the next batch.
constexpr size_t PIECE_BUDGET = (size_t)48 << 20;

struct Piece {
    SymBuf sym;
    BlockDecoder dec;
    bool found = false;
    uint64_t start = 0;        // bit position of the first block
    uint64_t cur = 0;          // bit position after the last complete block
    size_t cur_n = 0;
    size_t cur_ends = 0;
    int end_kind = END_INPUT;
    int landed = -1;
    std::vector<MemberEnd> ends;
    std::vector<std::pair<uint32_t, size_t>> seg_crc;   // phase D: (crc, length) between member ends
};

}  // namespace pgz

struct ParallelGzipSource : Source {
    std::unique_ptr<Source> src;
    unsigned T;
    size_t S;                          // compressed bytes per piece of a full batch (the search for a block
                                       // start costs the same whatever the piece size)
    size_t S_min;                      // short inputs are still cut into pieces of at least this size
    std::vector<uint8_t> comp;         // [carry | fresh] + INPUT_PAD zero bytes
    size_t comp_len = 0;
    unsigned start_bit = 0;            // bit offset of the first block inside comp[0]
    bool src_eof = false, finished = false, started = false;
    std::vector<uint8_t> window;       // last 32 KiB of output
    uint32_t run_crc = 0;              // CRC-32 / length of the current member so far
    uint64_t run_len = 0;
    std::vector<uint8_t> out;
    size_t out_pos = 0, out_len = 0;
    std::vector<std::unique_ptr<pgz::Piece>> pieces;
    std::string pending_error;         // raised after the bytes decoded before it have been handed out
    double t_fill = 0, t_search = 0, t_run = 0, t_window = 0, t_resolve = 0, t_crc = 0;   // SHB_TIMING=1
    unsigned n_batches = 0, n_discarded = 0;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    ~ParallelGzipSource() override
    {
        if (getenv("SHB_TIMING"))
            fprintf(stderr, "[pgzip] threads=%u batches=%u discarded_pieces=%u fill=%.3f search=%.3f inflate=%.3f windows=%.3f "
                    "resolve+crc=%.3f verify=%.3f s\n", T, n_batches, n_discarded, t_fill, t_search, t_run, t_window, t_resolve, t_crc);
    }

    ParallelGzipSource(std::unique_ptr<Source> s, unsigned threads, size_t piece_bytes = (size_t)2 << 20,
                       size_t min_piece_bytes = (size_t)256 << 10)
        : src(std::move(s)), T(std::max(1u, threads)), S(piece_bytes), S_min(std::min(piece_bytes, min_piece_bytes)),
          window(pgz::WINDOW, 0)
    {
        for (unsigned i = 0; i < T; i++) pieces.emplace_back(new pgz::Piece());
    }

    size_t read(uint8_t *dst, size_t n) override
    {
        size_t produced = 0;
        while (produced < n) {
            if (out_pos == out_len) {
                if (!pending_error.empty()) {
                    if (produced) break;   // bytes decoded before the failure go out first
                    throw std::runtime_error(pending_error);
                }
                if (finished) break;
                // a whole batch is resolved straight into the caller's buffer when it fits (short read)
                size_t direct = next_batch(dst + produced, n - produced);
                produced += direct;
                if (direct) break;
                continue;
            }
            size_t k = std::min(n - produced, out_len - out_pos);
            memcpy(dst + produced, out.data() + out_pos, k);
            out_pos += k;
            produced += k;
        }
        return produced;
    }

  private:
    void fill(size_t want_total)
    {
        if (comp.size() < want_total + pgz::INPUT_PAD) comp.resize(want_total + pgz::INPUT_PAD);
        while (!src_eof && comp_len < want_total) {
            size_t got = src->read(comp.data() + comp_len, want_total - comp_len);
            if (got == 0) src_eof = true;
            comp_len += got;
        }
        memset(comp.data() + comp_len, 0, pgz::INPUT_PAD);
    }

    // Decode from p.cur until the piece lands on a later piece's start, runs out of input, or the stream ends.
    void run_piece(unsigned k, unsigned np)
    {
        using namespace pgz;
        Piece &p = *pieces[k];
        BitReader br;
        br.p = comp.data();
        br.end_bits = (uint64_t)comp_len * 8;
        unsigned ti = k + 1;
        for (;;) {
            while (ti < np && (!pieces[ti]->found || pieces[ti]->start < p.cur)) ti++;
            if (ti < np && pieces[ti]->start == p.cur) { p.end_kind = END_LANDED; p.landed = (int)ti; return; }
            if (p.sym.produced() > PIECE_BUDGET) { p.end_kind = END_BUDGET; return; }
            br.bitpos = p.cur;
            Status st = p.dec.block(br, p.sym);
            if (st == BLOCK_OK) {
                p.cur = br.bitpos;
                p.cur_n = p.sym.n;
                p.cur_ends = p.ends.size();
                continue;
            }
            if (st == BAD_DATA) { p.end_kind = END_BAD; return; }
            if (st == BLOCK_FINAL) {
                size_t q = (size_t)((br.bitpos + 7) >> 3);
                if (q + 8 <= comp_len) {
                    const uint8_t *t = comp.data() + q;
                    MemberEnd me;
                    me.out_n = p.sym.produced();
                    me.crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
                    me.isize = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
                    q += 8;
                    if (q == comp_len && src_eof) {
                        p.ends.push_back(me);
                        p.cur = (uint64_t)q * 8;
                        p.cur_n = p.sym.n;
                        p.cur_ends = p.ends.size();
                        p.end_kind = END_STREAM;
                        return;
                    }
                    long h = gzip_header_len(comp.data() + q, comp_len - q);
                    if (h > 0) {
                        p.ends.push_back(me);
                        p.cur = (uint64_t)(q + (size_t)h) * 8;
                        p.cur_n = p.sym.n;
                        p.cur_ends = p.ends.size();
                        continue;
                    }
                    if (h < 0 || src_eof) { p.end_kind = END_BAD_HEADER; return; }
                }
                st = NEED_INPUT;   // trailer or next header not complete in this batch: redo the final block later
            }
            // NEED_INPUT: roll back to the last block boundary
            p.sym.n = p.cur_n;
            p.ends.resize(p.cur_ends);
            p.end_kind = END_INPUT;
            return;
        }
    }

    // Find the first trustworthy start in [lo, hi) (byte range): a gzip member header followed by a valid
    // block, or a non-final dynamic block that decodes cleanly and is followed by another valid block.
    void search_piece(unsigned k, size_t lo, size_t hi)
    {
        using namespace pgz;
        Piece &p = *pieces[k];
        p.found = false;
        BitReader br;
        br.p = comp.data();
        br.end_bits = (uint64_t)comp_len * 8;
        const uint8_t *c = comp.data();
        for (size_t i = lo; i < hi; i++) {
            if (c[i] == 0x1f && c[i + 1] == 0x8b && c[i + 2] == 8) {
                long h = gzip_header_len(c + i, comp_len - i);
                // only trust it right after a member trailer would end, i.e. validated by decoding
                if (h > 0) {
                    br.bitpos = (uint64_t)(i + (size_t)h) * 8;
                    p.sym.n = WINDOW;
                    Status st = p.dec.block(br, p.sym, MODE_MEMBER_START);
                    if (st == BLOCK_OK || st == BLOCK_FINAL) {
                        p.found = true;
                        p.start = p.cur = (uint64_t)(i + (size_t)h) * 8;
                        p.sym.n = p.cur_n = WINDOW;   // the block is decoded again by run_piece
                        return;
                    }
                }
            }
            uint64_t v;
            memcpy(&v, c + i, 8);
            for (unsigned b = 0; b < 8; b++) {
                uint64_t t = v >> b;
                if ((t & 7) != 4 || ((t >> 3) & 31) > 29 || ((t >> 8) & 31) > 29) continue;
                br.bitpos = (uint64_t)i * 8 + b;
                if (!precode_complete(br)) continue;   // rejects ~99 % of what is left without building a table
                p.sym.n = WINDOW;
                if (p.dec.block(br, p.sym, MODE_BLOCK_START) != BLOCK_OK) continue;
                const uint64_t after = br.bitpos;
                const size_t n_after = p.sym.n;
                Status st2 = p.dec.block(br, p.sym);   // the next block must make sense too
                if (st2 == BAD_DATA) continue;
                p.found = true;
                p.start = (uint64_t)i * 8 + b;
                p.cur = after;                          // keep the first block's output
                p.sym.n = p.cur_n = n_after;
                return;
            }
        }
        p.sym.n = WINDOW;
    }

    // returns the bytes written straight to `direct` (0: the batch is in `out`, or the stream has ended)
    size_t next_batch(uint8_t *direct, size_t direct_cap)
    {
        using namespace pgz;
        if (!started) {
            fill(T * S);
            long h = gzip_header_len(comp.data(), comp_len);
            if (h <= 0) throw std::runtime_error(comp_len == 0 ? "EOF" : "gzip: invalid header");
            memmove(comp.data(), comp.data() + h, comp_len - (size_t)h + INPUT_PAD);
            comp_len -= (size_t)h;
            start_bit = 0;
            started = true;
        }
        size_t want = T * S;
        for (;;) {
            double t0 = now();
            fill(std::max(want, comp_len));
            t_fill += now() - t0;
            const unsigned np = (unsigned)std::max<size_t>(1, std::min<size_t>(T, comp_len / S_min));
            for (unsigned k = 0; k < T; k++) {
                Piece &p = *pieces[k];
                p.found = false;
                p.ends.clear();
                p.cur_ends = 0;
                p.landed = -1;
                p.end_kind = END_INPUT;
                if (k < np) p.sym.reset(comp_len / np * 5);
            }
            Piece &p0 = *pieces[0];
            p0.found = true;
            p0.start = p0.cur = start_bit;
            p0.cur_n = WINDOW;
            const size_t piece_len = comp_len / np;
            t0 = now();
            parallel_for(np, [&](unsigned k) {
                if (k > 0) search_piece(k, k * piece_len, std::min(comp_len, (k + 1) * piece_len));
            });
            t_search += now() - t0;
            t0 = now();
            parallel_for(np, [&](unsigned k) { if (pieces[k]->found) run_piece(k, np); });
            t_run += now() - t0;
            n_batches++;

            // chain the pieces front to back
            std::vector<unsigned> chain;
            for (unsigned k = 0;;) {
                chain.push_back(k);
                if (pieces[k]->end_kind != END_LANDED) break;
                k = (unsigned)pieces[k]->landed;
            }
            Piece &last = *pieces[chain.back()];
            if (last.end_kind == END_BAD) throw std::runtime_error("gzip: invalid compressed data");
            if (last.end_kind == END_BAD_HEADER) throw std::runtime_error("gzip: invalid header");
            size_t total = 0;
            for (unsigned k : chain) total += pieces[k]->sym.produced();
            if (last.end_kind == END_INPUT && total == 0 && chain.size() == 1) {
                // not one complete block in the whole batch
                if (src_eof) throw std::runtime_error("unexpected EOF");
                want = std::max(want, comp_len) * 2;
                continue;
            }
            if (last.end_kind == END_INPUT && src_eof) pending_error = "unexpected EOF";   // truncated stream

            n_discarded += np - (unsigned)chain.size();
            t0 = now();
            // windows down the chain (sequential, 32 KiB per piece), then resolve + CRC in parallel
            std::vector<std::vector<uint8_t>> win(chain.size() + 1);
            std::vector<size_t> off(chain.size() + 1, 0);
            win[0] = window;
            for (size_t ci = 0; ci < chain.size(); ci++) {
                const Piece &p = *pieces[chain[ci]];
                const size_t n = p.sym.produced();
                const uint16_t *s = p.sym.symbols();
                const std::vector<uint8_t> &w = win[ci];
                std::vector<uint8_t> nw(WINDOW);
                if (n >= WINDOW) {
                    for (size_t j = 0; j < WINDOW; j++) {
                        uint16_t v = s[n - WINDOW + j];
                        nw[j] = v < 256 ? (uint8_t)v : w[v - 256];
                    }
                } else {
                    memcpy(nw.data(), w.data() + n, WINDOW - n);
                    for (size_t j = 0; j < n; j++) {
                        uint16_t v = s[j];
                        nw[WINDOW - n + j] = v < 256 ? (uint8_t)v : w[v - 256];
                    }
                }
                win[ci + 1] = std::move(nw);
                off[ci + 1] = off[ci] + n;
            }
            t_window += now() - t0;
            t0 = now();
            const bool to_caller = total <= direct_cap;
            if (!to_caller && out.size() < total) out.resize(total);
            uint8_t *const dest = to_caller ? direct : out.data();
            parallel_for((unsigned)chain.size(), [&](unsigned ci) {
                Piece &p = *pieces[chain[ci]];
                const size_t n = p.sym.produced();
                const uint16_t *s = p.sym.symbols();
                std::vector<uint8_t> lut(256 + WINDOW);
                for (unsigned j = 0; j < 256; j++) lut[j] = (uint8_t)j;
                memcpy(lut.data() + 256, win[ci].data(), WINDOW);
                uint8_t *o = dest + off[ci];
                size_t j = 0;
#if defined(__SSE2__)
                const __m128i hi = _mm_set1_epi16((short)0xff00);
                for (; j + 16 <= n; j += 16) {   // markers are rare after the first 32 KiB: pack 16 literals at once
                    __m128i a = _mm_loadu_si128((const __m128i *)(s + j)), b = _mm_loadu_si128((const __m128i *)(s + j + 8));
                    if (_mm_movemask_epi8(_mm_cmpeq_epi16(_mm_and_si128(_mm_or_si128(a, b), hi), _mm_setzero_si128())) == 0xffff)
                        _mm_storeu_si128((__m128i *)(o + j), _mm_packus_epi16(a, b));
                    else
                        for (size_t q = j; q < j + 16; q++) o[q] = lut[s[q]];
                }
#endif
                for (; j < n; j++) o[j] = lut[s[j]];
                p.seg_crc.clear();
                size_t from = 0;
                for (size_t e = 0; e <= p.ends.size(); e++) {
                    size_t to = e < p.ends.size() ? p.ends[e].out_n : n;
                    uint32_t c = (uint32_t)crc32(0L, Z_NULL, 0);
                    for (size_t a = from; a < to;) {   // crc32() takes a 32-bit length
                        size_t step = std::min<size_t>(to - a, (size_t)1 << 30);
                        c = (uint32_t)crc32(c, o + a, (uInt)step);
                        a += step;
                    }
                    p.seg_crc.emplace_back(c, to - from);
                    from = to;
                }
            });
            t_resolve += now() - t0;
            t0 = now();
            for (unsigned k : chain) {
                Piece &p = *pieces[k];
                for (size_t e = 0; e < p.seg_crc.size(); e++) {
                    run_crc = (uint32_t)crc32_combine(run_crc, p.seg_crc[e].first, (z_off_t)p.seg_crc[e].second);
                    run_len += p.seg_crc[e].second;
                    if (e < p.ends.size()) {
                        if (run_crc != p.ends[e].crc || (uint32_t)run_len != p.ends[e].isize)
                            throw std::runtime_error("gzip: invalid checksum");
                        run_crc = 0;
                        run_len = 0;
                    }
                }
            }
            t_crc += now() - t0;
            window = std::move(win[chain.size()]);
            out_pos = 0;
            out_len = to_caller ? 0 : total;
            const size_t direct_n = to_caller ? total : 0;
            if (last.end_kind == END_STREAM) {
                finished = true;
                comp_len = 0;
            } else {
                const size_t from = (size_t)(last.cur >> 3);
                start_bit = (unsigned)(last.cur & 7);
                memmove(comp.data(), comp.data() + from, comp_len - from);
                comp_len -= from;
            }
            return direct_n;
        }
    }
};

}  // namespace host
