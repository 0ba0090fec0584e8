// pzstd.hpp — zstd input whose frames are decoded in parallel.
//
// The reference reads .zst through xopen -> klauspost/compress/zstd (go.mod:20-31).  A zstd FRAME is one sequential stream,
// but a file may hold many (pzstd, `zstd` runs concatenated, per-chunk writers): frames are independent, every frame header
// says how long the frame is once its blocks are walked (ZSTD_findFrameCompressedSize: block headers only, nothing is
// decoded) and usually how much it holds (ZSTD_getFrameContentSize).  So: read a batch of the file, delimit the complete
// frames in it, decode them on T threads straight into their places of one output buffer, hand that out in order.  A frame that
// does not fit the batch, a frame without a content size, damaged or truncated input all go to the streaming decoder
// (ZstdSource) from that byte on — the single-stream path with its own error texts — so the worst case is what it was before.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstring>
#include <exception>
#include <memory>
#include <stdexcept>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include "source.hpp"

namespace host {

// bytes already read from `src`, then `src` itself
struct PrefixSource : Source {
    std::vector<uint8_t> pre;
    size_t pos = 0;
    std::unique_ptr<Source> src;
    PrefixSource(std::vector<uint8_t> p, std::unique_ptr<Source> s) : pre(std::move(p)), src(std::move(s)) {}
    size_t read(uint8_t *dst, size_t n) override
    {
        if (pos < pre.size()) {
            const size_t k = std::min(n, pre.size() - pos);
            memcpy(dst, pre.data() + pos, k);
            pos += k;
            return k;
        }
        return src->read(dst, n);
    }
};

template <class StreamingZstd>   // ZstdSource (decompress.hpp): the fallback for everything that is not a run of small frames
struct ParallelZstdSourceT : Source {
    std::unique_ptr<Source> src;
    unsigned threads;
    size_t batch;
    void *lib = nullptr;
    size_t (*find_size)(const void *, size_t) = nullptr;
    unsigned long long (*content_size)(const void *, size_t) = nullptr;
    size_t (*decompress)(void *, size_t, const void *, size_t) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    std::unique_ptr<uint8_t[]> in, out;      // new[]: never zero-filled
    size_t in_cap = 0, out_cap = 0, out_len = 0;
    size_t in_len = 0, in_pos = 0, out_pos = 0;
    bool src_eof = false;
    std::unique_ptr<Source> tail;   // set once the streaming decoder has taken over
    std::exception_ptr pending;

    ParallelZstdSourceT(std::unique_ptr<Source> s, unsigned t, size_t batch_bytes = (size_t)64 << 20)
        : src(std::move(s)), threads(t < 1 ? 1 : t), batch(batch_bytes)
    {
        lib = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib) throw std::runtime_error("zstd: libzstd.so.1 not available");
        find_size = (size_t(*)(const void *, size_t))dlsym(lib, "ZSTD_findFrameCompressedSize");
        content_size = (unsigned long long (*)(const void *, size_t))dlsym(lib, "ZSTD_getFrameContentSize");
        decompress = (size_t(*)(void *, size_t, const void *, size_t))dlsym(lib, "ZSTD_decompress");
        is_error = (unsigned (*)(size_t))dlsym(lib, "ZSTD_isError");
        if (!find_size || !content_size || !decompress || !is_error) throw std::runtime_error("zstd: symbols missing");
        in.reset(new uint8_t[batch]);
        in_cap = batch;
    }

    size_t read(uint8_t *dst, size_t n) override
    {
        if (pending) {   // a failure met after good bytes had been handed out: report it now
            std::exception_ptr e = pending;
            pending = nullptr;
            std::rethrow_exception(e);
        }
        size_t produced = 0;
        try {
            while (produced < n) {
                if (tail) {
                    const size_t k = tail->read(dst + produced, n - produced);
                    produced += k;
                    if (k == 0) break;
                    continue;
                }
                if (out_pos < out_len) {
                    const size_t k = std::min(n - produced, out_len - out_pos);
                    memcpy(dst + produced, out.get() + out_pos, k);
                    out_pos += k;
                    produced += k;
                    continue;
                }
                size_t direct = 0;
                if (!next_batch(dst + produced, n - produced, &direct)) break;
                produced += direct;
            }
        } catch (...) {
            // the bytes decoded before the damage are delivered first, like the other readers (and the reference's) do
            if (produced == 0) throw;
            pending = std::current_exception();
        }
        return produced;
    }

  private:
    struct Frame { size_t off, size, content, out_off; };

    // Decodes the complete frames at the front of the input: straight into dst when their text fits in `room` (*direct = bytes
    // written there; the usual case, the consumer is the read-ahead thread with buffers of 192 MB), else into `out`.  False at
    // the end of the input.
    bool next_batch(uint8_t *dst, size_t room, size_t *direct)
    {
        // keep the unconsumed tail of the previous batch, fill up
        if (in_pos) {
            memmove(in.get(), in.get() + in_pos, in_len - in_pos);
            in_len -= in_pos;
            in_pos = 0;
        }
        while (!src_eof && in_len < in_cap) {
            const size_t k = src->read(in.get() + in_len, in_cap - in_len);
            if (k == 0) src_eof = true;
            in_len += k;
        }
        if (in_len == 0) return false;
        std::vector<Frame> frames;
        size_t p = 0, total = 0;
        const size_t cap_out = std::max<size_t>(room, (size_t)32 << 20);   // decoded bytes per batch
        while (p < in_len) {
            const size_t sz = find_size(in.get() + p, in_len - p);
            if (is_error(sz)) break;                                   // incomplete in this batch, or not a frame
            const unsigned long long cs = content_size(in.get() + p, in_len - p);
            if (cs == ~0ull || cs == ~0ull - 1) break;                  // unknown / error: the streaming decoder handles it
            if (!frames.empty() && total + cs > cap_out) break;
            if (cs > ((size_t)1 << 31)) break;
            frames.push_back({p, sz, (size_t)cs, total});
            total += (size_t)cs;
            p += sz;
        }
        if (frames.empty()) {
            // one long frame, a frame without a content size, a truncated or damaged stream: stream it from here
            std::vector<uint8_t> pre(in.get(), in.get() + in_len);
            in_len = 0;
            tail.reset(new StreamingZstd(std::unique_ptr<Source>(new PrefixSource(std::move(pre), std::move(src)))));
            return true;
        }
        uint8_t *target = dst;
        if (total <= room) {
            *direct = total;
        } else {
            if (total > out_cap) {
                out.reset(new uint8_t[total]);
                out_cap = total;
            }
            target = out.get();
            out_len = total;
            out_pos = 0;
        }
        std::atomic<size_t> next{0};
        std::atomic<bool> bad{false};
        auto work = [&] {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= frames.size() || bad) return;
                const Frame &f = frames[i];
                const size_t rc = decompress(target + f.out_off, f.content, in.get() + f.off, f.size);
                if (is_error(rc) || rc != f.content) bad = true;
            }
        };
        const unsigned nt = (unsigned)std::min<size_t>(threads, frames.size());
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto &th : pool) th.join();
        if (bad) {
            *direct = 0;
            out_len = out_pos = 0;
            throw std::runtime_error("zstd: invalid compressed data");
        }
        in_pos = p;
        return true;
    }
};

}  // namespace host
