// The chunk-boundary rules of the compiled host (find_cut, find_record_start in csrc/host/main.cpp) as C entry points for
// tests/test_host_ranges_cpu.py.  main.cpp is included whole with its main() renamed; the shb_* symbols it references come
// from libseqhash_b200.so, none of them is called here (no GPU needed).
#define main shb_cli_main
#include "../../seqhasher_b200/csrc/host/main.cpp"
#undef main

extern "C" {
size_t emul_find_cut(const uint8_t *buf, size_t n, int format) { return find_cut(buf, n, format); }
size_t emul_find_record_start(const uint8_t *buf, size_t n, size_t from, int format, int at_eof)
{
    return find_record_start(buf, n, from, format, at_eof != 0);
}
}
