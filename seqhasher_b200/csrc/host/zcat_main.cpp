// shb-zcat — writes the decompressed bytes of a file (or stdin) to stdout through exactly the input stack the
// compiled host puts in front of the GPU record loop (decompress.hpp).  It exists to test and time the
// decompression feeders without a GPU:  shb-zcat [-t threads] [-q] file|-   (-q: count bytes, write nothing)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "decompress.hpp"

int main(int argc, char **argv)
{
    const char *path = nullptr;
    bool quiet = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-t") && i + 1 < argc) setenv("SHB_DECOMP_THREADS", argv[++i], 1);
        else if (!strcmp(argv[i], "-q")) quiet = true;
        else path = argv[i];
    }
    if (!path) { fprintf(stderr, "usage: shb-zcat [-t threads] [-q] file|-\n"); return 2; }
    FILE *f = strcmp(path, "-") ? fopen(path, "rb") : stdin;
    if (!f) { fprintf(stderr, "shb-zcat: cannot open %s\n", path); return 1; }
    try {
        auto t0 = std::chrono::steady_clock::now();
        auto src = host::open_maybe_compressed(f, f != stdin);
        std::vector<uint8_t> buf((size_t)32 << 20);
        size_t total = 0;
        for (;;) {
            size_t got = src->read(buf.data(), buf.size());
            if (got == 0) break;
            total += got;
            if (!quiet && fwrite(buf.data(), 1, got, stdout) != got) { fprintf(stderr, "shb-zcat: short write\n"); return 1; }
        }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (quiet) printf("{\"bytes\": %zu, \"seconds\": %.4f, \"MBps\": %.1f}\n", total, s, total / s / 1e6);
    } catch (const std::exception &e) {
        fflush(stdout);
        fprintf(stderr, "shb-zcat: %s\n", e.what());
        return 1;
    }
    fflush(stdout);
    return 0;
}
