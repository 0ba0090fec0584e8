// decompress.hpp — transparent input decompression for the compiled host (gzip, zstd, xz, bzip2 by magic
// bytes, like the xopen layer under bio's fastx reader; go.mod:20-31).  zlib is linked; libzstd, liblzma and
// libbz2 ship without headers in this image, so their few entry points are declared here and bound with dlopen.
#pragma once
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <algorithm>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "pbzip2.hpp"
#include "pgzip.hpp"
#include "readahead.hpp"
#include "pzstd.hpp"
#include "source.hpp"

namespace host {

struct FileSource : Source {
    FILE *f;
    bool own;
    std::vector<uint8_t> pushed;   // bytes handed back after sniffing
    size_t ppos = 0;
    explicit FileSource(FILE *fp, bool owned) : f(fp), own(owned) {}
    ~FileSource() override { if (own && f) fclose(f); }
    size_t read(uint8_t *dst, size_t n) override
    {
        if (ppos < pushed.size()) {
            size_t k = std::min(n, pushed.size() - ppos);
            memcpy(dst, pushed.data() + ppos, k);
            ppos += k;
            return k;
        }
        // A large read from a regular file is split over a few threads: one thread copies out of the page
        // cache at 5-8 GB/s, which is slower than the GPU record loop consumes plain FASTA.
        constexpr size_t kParallelMin = (size_t)16 << 20;
        if (n >= kParallelMin) {
            struct stat st;
            const int fd = fileno(f);
            off_t off = ftello(f);
            if (off >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > off) {
                const size_t want = std::min<size_t>(n, (size_t)(st.st_size - off));
                if (want >= kParallelMin) {
                    const unsigned parts = 4;
                    const size_t step = (want / parts + 4095) & ~(size_t)4095;
                    size_t got[parts] = {0, 0, 0, 0};
                    pgz::parallel_for(parts, [&](unsigned i) {
                        const size_t lo = std::min(want, (size_t)i * step), hi = std::min(want, lo + step);
                        size_t done = 0;
                        while (lo + done < hi) {
                            ssize_t r = pread(fd, dst + lo + done, hi - lo - done, off + (off_t)(lo + done));
                            if (r <= 0) break;
                            done += (size_t)r;
                        }
                        got[i] = done;
                    });
                    size_t total = 0;
                    for (unsigned i = 0; i < parts; i++) {
                        const size_t lo = std::min(want, (size_t)i * step), hi = std::min(want, lo + step);
                        total += got[i];
                        if (got[i] < hi - lo) break;   // short part (file shrank): stop at the first hole
                    }
                    fseeko(f, off + (off_t)total, SEEK_SET);
                    if (total) return total;
                }
            }
        }
        return fread(dst, 1, n, f);
    }
    void unread(const uint8_t *p, size_t n) { pushed.assign(p, p + n); ppos = 0; }
};

// Generic "compressed chunk in, plain bytes out" adapter
struct Inflater : Source {
    std::unique_ptr<Source> src;
    std::vector<uint8_t> in;
    size_t in_pos = 0, in_len = 0;
    bool src_eof = false, done = false;
    explicit Inflater(std::unique_ptr<Source> s) : src(std::move(s)), in(1 << 18) {}
    void refill()
    {
        if (in_pos == in_len && !src_eof) {
            in_len = src->read(in.data(), in.size());
            in_pos = 0;
            if (in_len == 0) src_eof = true;
        }
    }
};

struct GzipSource : Inflater {
    z_stream z;
    explicit GzipSource(std::unique_ptr<Source> s) : Inflater(std::move(s))
    {
        memset(&z, 0, sizeof z);
        if (inflateInit2(&z, 15 + 32) != Z_OK) throw std::runtime_error("gzip: init failed");
    }
    ~GzipSource() override { inflateEnd(&z); }
    size_t read(uint8_t *dst, size_t n) override
    {
        size_t produced = 0;
        while (produced < n && !done) {
            refill();
            if (in_pos == in_len && src_eof) {   // input ended inside a member
                if (produced) break;
                throw std::runtime_error("unexpected EOF");
            }
            z.next_in = in.data() + in_pos;
            z.avail_in = (uInt)(in_len - in_pos);
            z.next_out = dst + produced;
            z.avail_out = (uInt)std::min<size_t>(n - produced, 1u << 30);
            int rc = inflate(&z, Z_NO_FLUSH);
            in_pos = in_len - z.avail_in;
            produced = (size_t)(z.next_out - dst);
            if (rc == Z_STREAM_END) {   // next member of a multi-member file (pgzip / bgzip output)
                refill();
                if (in_pos == in_len && src_eof) done = true;
                else inflateReset(&z);
            } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                throw std::runtime_error("gzip: invalid compressed data");
            }
        }
        return produced;
    }
};

struct ZstdSource : Inflater {
    struct Buf { void *p; size_t size, pos; };
    void *lib = nullptr, *ds = nullptr;
    bool mid_frame = false;
    void *(*create)() = nullptr;
    size_t (*freeds)(void *) = nullptr;
    size_t (*decompress)(void *, Buf *, Buf *) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    explicit ZstdSource(std::unique_ptr<Source> s) : Inflater(std::move(s))
    {
        lib = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib) throw std::runtime_error("zstd: libzstd.so.1 not available");
        create = (void *(*)())dlsym(lib, "ZSTD_createDStream");
        freeds = (size_t(*)(void *))dlsym(lib, "ZSTD_freeDStream");
        decompress = (size_t(*)(void *, Buf *, Buf *))dlsym(lib, "ZSTD_decompressStream");
        is_error = (unsigned (*)(size_t))dlsym(lib, "ZSTD_isError");
        if (!create || !freeds || !decompress || !is_error) throw std::runtime_error("zstd: symbols missing");
        ds = create();
    }
    ~ZstdSource() override { if (ds) freeds(ds); }
    size_t read(uint8_t *dst, size_t n) override
    {
        Buf out{dst, n, 0};
        while (out.pos < n && !done) {
            refill();
            if (in_pos == in_len && src_eof) {
                if (!mid_frame) { done = true; break; }
                if (out.pos) break;
                throw std::runtime_error("unexpected EOF");   // input ended inside a frame
            }
            Buf ib{in.data(), in_len, in_pos};
            size_t rc = decompress(ds, &out, &ib);
            in_pos = ib.pos;
            if (is_error(rc)) throw std::runtime_error("zstd: invalid compressed data");
            mid_frame = rc != 0;
        }
        return out.pos;
    }
};

struct Bz2Source : Inflater {
    struct bz_stream {
        char *next_in; unsigned avail_in, total_in_lo32, total_in_hi32;
        char *next_out; unsigned avail_out, total_out_lo32, total_out_hi32;
        void *state; void *(*bzalloc)(void *, int, int); void (*bzfree)(void *, void *); void *opaque;
    } z;
    void *lib = nullptr;
    int (*init)(bz_stream *, int, int) = nullptr;
    int (*dec)(bz_stream *) = nullptr;
    int (*end)(bz_stream *) = nullptr;
    explicit Bz2Source(std::unique_ptr<Source> s) : Inflater(std::move(s))
    {
        lib = dlopen("libbz2.so.1.0", RTLD_NOW | RTLD_LOCAL);
        if (!lib) lib = dlopen("libbz2.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib) throw std::runtime_error("bzip2: libbz2 not available");
        init = (int (*)(bz_stream *, int, int))dlsym(lib, "BZ2_bzDecompressInit");
        dec = (int (*)(bz_stream *))dlsym(lib, "BZ2_bzDecompress");
        end = (int (*)(bz_stream *))dlsym(lib, "BZ2_bzDecompressEnd");
        memset(&z, 0, sizeof z);
        if (!init || !dec || !end || init(&z, 0, 0) != 0) throw std::runtime_error("bzip2: init failed");
    }
    ~Bz2Source() override { if (end) end(&z); }
    size_t read(uint8_t *dst, size_t n) override
    {
        size_t produced = 0;
        while (produced < n && !done) {
            refill();
            if (in_pos == in_len && src_eof) {   // input ended inside a stream
                if (produced) break;
                throw std::runtime_error("unexpected EOF");
            }
            z.next_in = (char *)in.data() + in_pos;
            z.avail_in = (unsigned)(in_len - in_pos);
            z.next_out = (char *)dst + produced;
            z.avail_out = (unsigned)std::min<size_t>(n - produced, 1u << 30);
            int rc = dec(&z);
            in_pos = in_len - z.avail_in;
            produced = (size_t)((uint8_t *)z.next_out - dst);
            if (rc == 4 /* BZ_STREAM_END */) {
                refill();
                if (in_pos == in_len && src_eof) done = true;
                else { end(&z); memset(&z, 0, sizeof z); init(&z, 0, 0); }   // concatenated streams
            } else if (rc != 0) {
                throw std::runtime_error("bzip2: invalid compressed data");
            }
        }
        return produced;
    }
};

struct XzSource : Inflater {
    struct lzma_stream {
        const uint8_t *next_in; size_t avail_in; uint64_t total_in;
        uint8_t *next_out; size_t avail_out; uint64_t total_out;
        const void *allocator; void *internal;
        void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3, *reserved_ptr4;
        uint64_t reserved_int1, reserved_int2; size_t reserved_int3, reserved_int4;
        int reserved_enum1, reserved_enum2;
    } z;
    void *lib = nullptr;
    bool pending_eof = false;
    int (*decoder)(lzma_stream *, uint64_t, uint32_t) = nullptr;
    int (*code)(lzma_stream *, int) = nullptr;
    void (*end)(lzma_stream *) = nullptr;
    // lzma_mt of liblzma >= 5.4 (the threaded decoder splits multi-block files, i.e. what `xz -T` writes;
    // single-block files run on one thread inside the same call)
    struct lzma_mt {
        uint32_t flags, threads;
        uint64_t block_size;
        uint32_t timeout, preset;
        const void *filters;
        int check, reserved_enum1, reserved_enum2, reserved_enum3;
        uint32_t reserved_int1, reserved_int2, reserved_int3, reserved_int4;
        uint64_t memlimit_threading, memlimit_stop, reserved_int7, reserved_int8;
        void *reserved_ptr1, *reserved_ptr2, *reserved_ptr3, *reserved_ptr4;
    };
    explicit XzSource(std::unique_ptr<Source> s, unsigned threads = 1) : Inflater(std::move(s))
    {
        lib = dlopen("liblzma.so.5", RTLD_NOW | RTLD_LOCAL);
        if (!lib) throw std::runtime_error("xz: liblzma.so.5 not available");
        decoder = (int (*)(lzma_stream *, uint64_t, uint32_t))dlsym(lib, "lzma_stream_decoder");
        code = (int (*)(lzma_stream *, int))dlsym(lib, "lzma_code");
        end = (void (*)(lzma_stream *))dlsym(lib, "lzma_end");
        memset(&z, 0, sizeof z);
        if (!decoder || !code || !end) throw std::runtime_error("xz: symbols missing");
        auto decoder_mt = (int (*)(lzma_stream *, const lzma_mt *))dlsym(lib, "lzma_stream_decoder_mt");
        bool ok = false;
        if (threads >= 2 && decoder_mt) {
            lzma_mt mt;
            memset(&mt, 0, sizeof mt);
            mt.flags = 0x08;   // LZMA_CONCATENATED
            mt.threads = threads;
            mt.memlimit_threading = (uint64_t)4 << 30;
            mt.memlimit_stop = UINT64_MAX;
            ok = decoder_mt(&z, &mt) == 0;
            if (!ok) memset(&z, 0, sizeof z);
        }
        if (!ok && decoder(&z, UINT64_MAX, 0x08 /* LZMA_CONCATENATED */) != 0) throw std::runtime_error("xz: init failed");
    }
    ~XzSource() override { if (end) end(&z); }
    size_t read(uint8_t *dst, size_t n) override
    {
        if (pending_eof) throw std::runtime_error("unexpected EOF");
        z.next_out = dst;
        z.avail_out = n;
        while (z.avail_out && !done) {
            refill();
            z.next_in = in.data() + in_pos;
            z.avail_in = in_len - in_pos;
            int rc = code(&z, (in_pos == in_len && src_eof) ? 3 /* LZMA_FINISH */ : 0 /* LZMA_RUN */);
            in_pos = in_len - z.avail_in;
            if (rc == 1 /* LZMA_STREAM_END */) done = true;
            else if (rc != 0 && rc != 10 /* LZMA_BUF_ERROR */) throw std::runtime_error("xz: invalid compressed data");
            else if (rc == 10 && in_pos == in_len && src_eof) {   // input ended inside a stream
                if (z.avail_out != n) { pending_eof = true; break; }   // hand out what was decoded first
                throw std::runtime_error("unexpected EOF");
            }
        }
        return n - z.avail_out;
    }
};

// Worker threads for the parallel decompressors: SHB_DECOMP_THREADS, else the machine's hardware threads
// (capped at 32); 1 selects the plain single-stream library paths.
inline unsigned decomp_threads()
{
    if (const char *e = getenv("SHB_DECOMP_THREADS")) {
        int v = atoi(e);
        if (v >= 1) return (unsigned)std::min(v, 64);
    }
    unsigned hc = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hc ? hc : 1u, 32u));
}

// Sniff the magic bytes and wrap the file in the matching decompressor; compressed inputs are decoded on a
// read-ahead thread so that they overlap the consumer.
inline std::unique_ptr<Source> open_maybe_compressed(FILE *f, bool own, bool *compressed = nullptr)
{
    auto fs = std::make_unique<FileSource>(f, own);
    uint8_t head[6];
    size_t k = fread(head, 1, sizeof head, f);
    fs->unread(head, k);
    if (compressed)
        *compressed = (k >= 2 && head[0] == 0x1f && head[1] == 0x8b) || (k >= 3 && head[0] == 'B' && head[1] == 'Z' && head[2] == 'h') ||
                      (k >= 6 && memcmp(head, "\xfd" "7zXZ\0", 6) == 0) ||
                      (k >= 4 && head[0] == 0x28 && head[1] == 0xb5 && head[2] == 0x2f && head[3] == 0xfd);
    const unsigned threads = decomp_threads();
    if (k >= 2 && head[0] == 0x1f && head[1] == 0x8b) {
        if (threads >= 2) {
            size_t piece = (size_t)2 << 20, min_piece = (size_t)256 << 10;
            if (const char *e = getenv("SHB_PGZIP_PIECE_KB")) {   // tests: tiny pieces stress the chaining
                long kb = atol(e);
                if (kb >= 1) piece = min_piece = (size_t)kb << 10;
            }
            return std::make_unique<ReadAhead>(std::make_unique<ParallelGzipSource>(std::move(fs), threads, piece, min_piece));
        }
        return std::make_unique<GzipSource>(std::move(fs));
    }
    if (k >= 3 && head[0] == 'B' && head[1] == 'Z' && head[2] == 'h') {
        if (threads >= 2) return std::make_unique<ReadAhead>(std::make_unique<ParallelBz2Source>(std::move(fs), threads));
        return std::make_unique<Bz2Source>(std::move(fs));
    }
    if (k >= 6 && memcmp(head, "\xfd" "7zXZ\0", 6) == 0) {
        if (threads >= 2) return std::make_unique<ReadAhead>(std::make_unique<XzSource>(std::move(fs), threads));
        return std::make_unique<XzSource>(std::move(fs));
    }
    if (k >= 4 && head[0] == 0x28 && head[1] == 0xb5 && head[2] == 0x2f && head[3] == 0xfd) {
        if (threads >= 2) {
            size_t batch = (size_t)64 << 20;
            if (const char *e = getenv("SHB_PZSTD_BATCH_KB")) {   // tests: small batches exercise the batch seams and the fallback
                long kb = atol(e);
                if (kb >= 1) batch = (size_t)kb << 10;
            }
            return std::make_unique<ReadAhead>(std::make_unique<ParallelZstdSourceT<ZstdSource>>(std::move(fs), threads, batch));
        }
        return std::make_unique<ZstdSource>(std::move(fs));
    }
    return fs;
}

}  // namespace host
