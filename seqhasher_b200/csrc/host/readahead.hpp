// readahead.hpp — runs a Source on its own thread, a few large buffers ahead of the consumer, so that
// decompression overlaps the GPU record loop and the output writes (the role of the read-ahead goroutine in
// the reference's pgzip reader, go.mod:20-31).
#pragma once
#include <condition_variable>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "source.hpp"

namespace host {

struct ReadAhead : Source {
    static constexpr int NB = 3;
    std::unique_ptr<Source> src;
    struct Buf { std::unique_ptr<uint8_t[]> d; size_t len = 0; };   // uninitialised: pages are touched when used
    size_t cap;
    Buf bufs[NB];
    std::mutex m;
    std::condition_variable cv;
    int head = 0, tail = 0, count = 0;
    size_t pos = 0;
    bool eof = false, stop = false;
    std::exception_ptr err;
    std::thread th;

    explicit ReadAhead(std::unique_ptr<Source> s, size_t buf_bytes = (size_t)192 << 20) : src(std::move(s)), cap(buf_bytes)
    {
        for (auto &b : bufs) b.d.reset(new uint8_t[buf_bytes]);
        th = std::thread([this] { feeder(); });
    }
    ~ReadAhead() override
    {
        {
            std::lock_guard<std::mutex> g(m);
            stop = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
    size_t read(uint8_t *dst, size_t n) override
    {
        size_t produced = 0;
        while (produced < n) {
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [this] { return count > 0 || eof; });
                if (count == 0) {
                    if (err && produced == 0) std::rethrow_exception(err);
                    break;   // bytes read before a failure are handed out first; the next call raises it
                }
            }
            Buf &b = bufs[head];
            size_t k = std::min(n - produced, b.len - pos);
            memcpy(dst + produced, b.d.get() + pos, k);
            pos += k;
            produced += k;
            if (pos == b.len) {
                std::lock_guard<std::mutex> g(m);
                head = (head + 1) % NB;
                count--;
                pos = 0;
                cv.notify_all();
            }
        }
        return produced;
    }

  private:
    void feeder()
    {
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m);
                cv.wait(g, [this] { return count < NB || stop; });
                if (stop) return;
            }
            Buf &b = bufs[tail];
            b.len = 0;
            std::exception_ptr failed;
            try {
                // one read per buffer: a source may hand back less than asked (the parallel gzip reader
                // resolves one whole batch straight into this buffer); only 0 means the end
                b.len = src->read(b.d.get(), cap);
            } catch (...) {
                failed = std::current_exception();
            }
            std::lock_guard<std::mutex> g(m);
            if (b.len > 0) {
                tail = (tail + 1) % NB;
                count++;
            }
            if (failed) err = failed;
            if (failed || b.len == 0) eof = true;
            cv.notify_all();
            if (eof) return;
        }
    }
};

}  // namespace host
