// pbzip2.hpp — parallel bzip2 reader for the compiled host.
//
// bzip2 blocks are independent (each carries its own Huffman tables and BWT origin) but start at arbitrary
// BIT positions behind a 48-bit magic.  The reference inflates them one after the other on one goroutine
// (dsnet/compress via xopen, go.mod:20-31) at 20-40 MB/s.  Here every batch of compressed bytes is scanned for
// the block and end-of-stream magics at all eight bit alignments, each block is re-packed as a tiny
// single-block stream ("BZh" + level, the block's bits shifted to a byte boundary, end-of-stream magic, the
// block's own CRC as stream CRC) and handed to libbz2 on a worker thread.  libbz2 verifies every block CRC;
// the stream's combined CRC is re-computed here from the block CRCs and checked against the trailer.
#pragma once
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "pgzip.hpp"   // parallel_for
#include "source.hpp"

namespace host {

struct ParallelBz2Source : Source {
    static constexpr uint64_t BLOCK_MAGIC = 0x314159265359ull, EOS_MAGIC = 0x177245385090ull;
    std::unique_ptr<Source> src;
    unsigned T;
    std::vector<uint8_t> comp;
    size_t comp_len = 0;
    bool src_eof = false, finished = false;
    // stream state at comp[0]: bit offset of the next marker, level of the current stream, running CRC
    unsigned start_bit = 0;
    int level = 0;            // 0 = a stream header ("BZh1".."BZh9") is expected at comp[0]
    uint32_t combined = 0;
    bool any_stream = false;
    struct Task { uint64_t from, to; int level; uint32_t crc; std::vector<uint8_t> out; size_t n = 0; std::string err; };
    std::vector<Task> tasks;
    size_t cur_task = 0, cur_pos = 0;
    std::string pending_error;
    void *lib = nullptr;
    int (*b2b)(char *, unsigned *, char *, unsigned, int, int) = nullptr;

    ParallelBz2Source(std::unique_ptr<Source> s, unsigned threads) : src(std::move(s)), T(std::max(1u, threads))
    {
        lib = dlopen("libbz2.so.1.0", RTLD_NOW | RTLD_LOCAL);
        if (!lib) lib = dlopen("libbz2.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib) throw std::runtime_error("bzip2: libbz2 not available");
        b2b = (int (*)(char *, unsigned *, char *, unsigned, int, int))dlsym(lib, "BZ2_bzBuffToBuffDecompress");
        if (!b2b) throw std::runtime_error("bzip2: symbols missing");
    }

    size_t read(uint8_t *dst, size_t n) override
    {
        size_t produced = 0;
        while (produced < n) {
            if (cur_task == tasks.size()) {
                if (!pending_error.empty()) {
                    if (produced) break;   // bytes decoded before the failure go out first
                    throw std::runtime_error(pending_error);
                }
                if (finished) break;
                next_batch();
                continue;
            }
            Task &t = tasks[cur_task];
            size_t k = std::min(n - produced, t.n - cur_pos);
            memcpy(dst + produced, t.out.data() + cur_pos, k);
            cur_pos += k;
            produced += k;
            if (cur_pos == t.n) { cur_task++; cur_pos = 0; }
        }
        return produced;
    }

  private:
    static inline uint64_t be64(const uint8_t *p)
    {
        uint64_t v;
        memcpy(&v, p, 8);
        return __builtin_bswap64(v);
    }
    // the 48 (or 32) bits at bit position `bit` (MSB-first bit order)
    inline uint64_t bits_at(uint64_t bit, unsigned nbits) const
    {
        const uint8_t *p = comp.data() + (bit >> 3);
        unsigned sh = (unsigned)(bit & 7);
        uint64_t hi = be64(p), lo = p[8];
        uint64_t v = sh ? (hi << sh) | (lo >> (8 - sh)) : hi;
        return v >> (64 - nbits);
    }
    void fill(size_t want)
    {
        if (comp.size() < want + 64) comp.resize(want + 64);
        while (!src_eof && comp_len < want) {
            size_t got = src->read(comp.data() + comp_len, want - comp_len);
            if (got == 0) src_eof = true;
            comp_len += got;
        }
        memset(comp.data() + comp_len, 0, 64);
    }
    void decode_task(Task &t)
    {
        const uint64_t nbits = t.to - t.from;
        const size_t body = (size_t)((nbits + 7) >> 3);
        std::vector<uint8_t> mini(4 + body + 16, 0);
        mini[0] = 'B'; mini[1] = 'Z'; mini[2] = 'h'; mini[3] = (uint8_t)('0' + t.level);
        const uint8_t *p = comp.data() + (t.from >> 3);
        const unsigned sh = (unsigned)(t.from & 7);
        if (sh == 0) memcpy(mini.data() + 4, p, body);
        else for (size_t i = 0; i < body; i++) mini[4 + i] = (uint8_t)((p[i] << sh) | (p[i + 1] >> (8 - sh)));
        if (nbits & 7) mini[4 + body - 1] &= (uint8_t)(0xff00u >> (nbits & 7));
        // end-of-stream magic + stream CRC (= the CRC of its only block), MSB first, at bit 32 + nbits
        uint64_t wbit = 32 + nbits;
        auto put = [&](uint64_t v, unsigned n) {
            for (unsigned i = 0; i < n; i++, wbit++)
                if ((v >> (n - 1 - i)) & 1) mini[wbit >> 3] |= (uint8_t)(0x80u >> (wbit & 7));
        };
        put(EOS_MAGIC, 48);
        put(t.crc, 32);
        const unsigned mini_len = (unsigned)((wbit + 7) >> 3);
        size_t cap = (size_t)t.level * 100000 + 4096;
        for (;;) {
            if (t.out.size() < cap) t.out.resize(cap);
            unsigned dlen = (unsigned)std::min<size_t>(t.out.size(), 0xfffffff0u);
            int rc = b2b((char *)t.out.data(), &dlen, (char *)mini.data(), mini_len, 0, 0);
            if (rc == 0) { t.n = dlen; return; }
            if (rc == -8 /* BZ_OUTBUFF_FULL: run-length expansion */ && cap < ((size_t)1 << 31)) { cap *= 4; continue; }
            t.err = "bzip2: invalid compressed data";
            return;
        }
    }

    void next_batch()
    {
        tasks.clear();
        cur_task = 0;
        cur_pos = 0;
        size_t want = std::max<size_t>((size_t)8 << 20, (size_t)T << 20);
        for (;;) {
            fill(std::max(want, comp_len));
            if (comp_len == 0 && src_eof) {
                if (!any_stream) throw std::runtime_error("EOF");
                finished = true;
                return;
            }
            // 1. markers at every bit alignment, in parallel over byte ranges
            const uint64_t total_bits = (uint64_t)comp_len * 8;
            const unsigned np = (unsigned)std::max<size_t>(1, std::min<size_t>(T, comp_len >> 16));
            std::vector<std::vector<std::pair<uint64_t, bool>>> found(np);
            pgz::parallel_for(np, [&](unsigned k) {
                size_t lo = comp_len * k / np, hi = comp_len * (k + 1) / np;
                const uint8_t *c = comp.data();
                for (size_t i = lo; i < hi; i++) {
                    // a marker starting inside byte i: its first full byte is one of 8 values per magic, cheap reject
                    uint64_t hi64 = be64(c + i);
                    uint64_t lo8 = c[i + 8];
                    for (unsigned sh = 0; sh < 8; sh++) {
                        uint64_t v = (sh ? (hi64 << sh) | (lo8 >> (8 - sh)) : hi64) >> 16;
                        if (v == BLOCK_MAGIC) found[k].emplace_back((uint64_t)i * 8 + sh, false);
                        else if (v == EOS_MAGIC) found[k].emplace_back((uint64_t)i * 8 + sh, true);
                    }
                }
            });
            std::vector<std::pair<uint64_t, bool>> marks;
            for (auto &f : found) marks.insert(marks.end(), f.begin(), f.end());
            // 2. walk the stream structure from the carried position
            uint64_t pos = start_bit;
            size_t mi = 0;
            uint64_t consumed_to = pos;       // everything before this bit is done
            int lvl = level;
            uint32_t comb = combined;
            bool stalled = false, at_end = false;
            std::string structural_error;
            while (true) {
                if (lvl == 0) {               // stream header at a byte boundary
                    size_t q = (size_t)((pos + 7) >> 3);
                    if (q == comp_len && src_eof) { at_end = true; break; }
                    if (q + 4 > comp_len) { stalled = true; break; }
                    const uint8_t *h = comp.data() + q;
                    if (h[0] != 'B' || h[1] != 'Z' || h[2] != 'h' || h[3] < '1' || h[3] > '9') {
                        structural_error = any_stream ? "bzip2: invalid stream header" : "bzip2: invalid header";
                        break;
                    }
                    lvl = h[3] - '0';
                    comb = 0;
                    any_stream = true;
                    pos = (uint64_t)(q + 4) * 8;
                    consumed_to = pos;
                    continue;
                }
                while (mi < marks.size() && marks[mi].first < pos) mi++;
                if (pos + 80 > total_bits) { stalled = true; break; }
                const uint64_t magic = bits_at(pos, 48);
                if (magic == EOS_MAGIC) {
                    uint32_t want_crc = (uint32_t)bits_at(pos + 48, 32);
                    // checked once the blocks before it have been decoded: remember it as a zero-length task
                    Task t;
                    t.from = t.to = pos;
                    t.level = -1;
                    t.crc = want_crc ^ comb;   // 0 when the combined CRC matches
                    tasks.push_back(std::move(t));
                    pos += 80;
                    lvl = 0;
                    consumed_to = (uint64_t)((pos + 7) >> 3) * 8;
                    continue;
                }
                if (magic != BLOCK_MAGIC) { structural_error = "bzip2: invalid compressed data"; break; }
                // the block runs to the next marker
                size_t mj = mi;
                while (mj < marks.size() && marks[mj].first <= pos) mj++;
                if (mj == marks.size()) { stalled = true; break; }   // its end is not in this batch
                Task t;
                t.from = pos;
                t.to = marks[mj].first;
                t.level = lvl;
                t.crc = (uint32_t)bits_at(pos + 48, 32);
                comb = ((comb << 1) | (comb >> 31)) ^ t.crc;
                tasks.push_back(std::move(t));
                pos = marks[mj].first;
                consumed_to = pos;
            }
            if (tasks.empty() && stalled) {
                if (src_eof) throw std::runtime_error("unexpected EOF");
                want = std::max(want, comp_len) * 2;   // one block larger than the batch
                continue;
            }
            // 3. blocks in parallel
            std::vector<unsigned> todo;
            for (unsigned i = 0; i < tasks.size(); i++) if (tasks[i].level > 0) todo.push_back(i);
            std::atomic<unsigned> next{0};
            pgz::parallel_for(std::min<unsigned>(T, (unsigned)todo.size()), [&](unsigned) {
                for (unsigned j; (j = next.fetch_add(1)) < todo.size();) decode_task(tasks[todo[j]]);
            });
            for (size_t i = 0; i < tasks.size(); i++) {
                Task &t = tasks[i];
                std::string e = t.level > 0 ? t.err : (t.crc != 0 ? std::string("bzip2: invalid checksum") : std::string());
                if (!e.empty()) {   // hand out what precedes the damage, then fail
                    tasks.resize(i);
                    pending_error = e;
                    return;
                }
            }
            if (!structural_error.empty()) { pending_error = structural_error; return; }
            if (stalled && src_eof) pending_error = "unexpected EOF";
            level = lvl;
            combined = comb;
            if (at_end) {
                finished = true;
                comp_len = 0;
                return;
            }
            const size_t from = (size_t)(consumed_to >> 3);
            start_bit = (unsigned)(consumed_to & 7);
            memmove(comp.data(), comp.data() + from, comp_len - from);
            comp_len -= from;
            return;
        }
    }
};

}  // namespace host
