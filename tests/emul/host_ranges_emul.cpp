// The chunk-boundary rules of the compiled host (find_cut, find_record_start in csrc/host/main.cpp) as C entry points for
// tests/test_host_ranges_cpu.py.  main.cpp is included whole with its main() renamed; the shb_* symbols it references come
// from libseqhash_b200.so, none of them is called here (no GPU needed).
#define main shb_cli_main
#include "../../seqhasher_b200/csrc/host/main.cpp"
#undef main

extern "C" {
// BGZF member walk of the file at `path`: n members written as 4 words each (start, data, clen, packed) into out (cap
// entries); returns the member count, -1 when the file is not BGZF throughout, -2 when out is too small
long emul_scan_bgzf(const char *path, unsigned long long *out, size_t cap)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -3;
    struct stat st;
    fstat(fileno(f), &st);
    std::vector<BgzfMember> m;
    const bool ok = scan_bgzf(fileno(f), (size_t)st.st_size, m);
    fclose(f);
    if (!ok) return -1;
    if (m.size() > cap) return -2;
    for (size_t i = 0; i < m.size(); i++) { out[4 * i] = m[i].start; out[4 * i + 1] = m[i].data; out[4 * i + 2] = m[i].clen; out[4 * i + 3] = m[i].packed; }
    return (long)m.size();
}
size_t emul_find_cut(const uint8_t *buf, size_t n, int format) { return find_cut(buf, n, format); }
size_t emul_find_record_start(const uint8_t *buf, size_t n, size_t from, int format, int at_eof)
{
    return find_record_start(buf, n, from, format, at_eof != 0);
}
}
