// main.cpp — compiled host of seqhasher_b200: the CLI contract of seqhasher (seqhasher.go:56-149: flags,
// positionals, version / usage short-circuits, error texts) in front of the C ABI.  The host keeps file I/O
// and decompression; everything inside processSequences' record loop (seqhasher.go:227-283) runs on the GPU
// through shb_textproc.  Go's `flag` rules are reproduced: -x and --x alike, `-x=v` or `-x v`, parsing stops
// at the first positional ("-" included) or after "--".
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/seqhash_b200.h"
#include "decompress.hpp"

namespace {

// SHB_TIMING=1: phase times on stderr (where does the wall time of one invocation go?)
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
const bool kTiming = getenv("SHB_TIMING") != nullptr;
const double kT0 = now_s();
void phase(const char *what) { if (kTiming) fprintf(stderr, "[timing] %-18s %.3f s\n", what, now_s() - kT0); }

const char *kVersion = "1.2.0";   // seqhasher.go:38
const char *kTypes[12] = {"sha1", "sha3", "md5", "xxhash", "xxh3", "cityhash", "murmur3", "nthash", "blake3", "k12",
                          "rapidhash", "rapidhash32"};
const char *kEmptyMsg = "Error: Empty DNA sequence provided, resulting in an empty hash.";   // seqhasher.go:293

struct Config {   // seqhasher.go:45-54
    bool headers_only = false, no_file_name = false, case_sensitive = false, show_version = false;
    std::vector<std::string> hash_types;
    std::string input, output, name_override;
};

std::string supported_list()
{
    std::string s;
    for (int i = 0; i < 12; i++) { if (i) s += ", "; s += kTypes[i]; }
    return s;
}

bool is_valid_hash_type(const std::string &t)   // seqhasher.go:142-149
{
    for (const char *k : kTypes) if (t == k) return true;
    return false;
}

std::string trim_space(const std::string &s)
{
    size_t a = 0, b = s.size();
    auto ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };
    while (a < b && ws(s[a])) a++;
    while (b > a && ws(s[b - 1])) b--;
    return s.substr(a, b - a);
}

void print_usage(FILE *w, const char *prog)
{
    fprintf(w, "SeqHasher v%s\n", kVersion);
    fprintf(w, "Usage: %s [options] <input_file> [output_file]\n", prog);
    fprintf(w, "Options:\n");
    fprintf(w, "  -o, --headersonly    Output only headers\n");
    fprintf(w, "  -H, --hash string    Hash type(s) (comma-separated: %s) (default \"sha1\")\n", supported_list().c_str());
    fprintf(w, "  -n, --nofilename     Do not include file name in output\n");
    fprintf(w, "  -c, --casesensitive  Case-sensitive hashing\n");
    fprintf(w, "  -f, --name string    Override input file name in output\n");
    fprintf(w, "  -v, --version        Show version information\n");
    fprintf(w, "\nSupported hash types: %s\n", supported_list().c_str());
    fprintf(w, "If input_file is '-' or omitted, reads from stdin.\n");
    fprintf(w, "If output_file is '-' or omitted, writes to stdout.\n");
    fprintf(w, "\nFor more detailed help, use -h or --help.\n");
}

enum ParseResult { PARSE_OK, PARSE_HELP, PARSE_FLAG_ERROR, PARSE_HASH_ERROR };

// parseFlags, seqhasher.go:101-140
ParseResult parse_flags(int argc, char **argv, Config &cfg, std::string &err)
{
    std::string hash_string = "sha1";
    int i = 1;
    for (; i < argc; i++) {
        std::string a = argv[i];
        if (a.size() < 2 || a[0] != '-') break;
        if (a == "--") { i++; break; }
        std::string body = a.substr(a[1] == '-' ? 2 : 1);
        if (body.empty() || body[0] == '-' || body[0] == '=') { err = "bad flag syntax: " + a; return PARSE_FLAG_ERROR; }
        std::string name = body, val;
        bool has_val = false;
        size_t eq = body.find('=');
        if (eq != std::string::npos) { name = body.substr(0, eq); val = body.substr(eq + 1); has_val = true; }
        if (name == "h" || name == "help") return PARSE_HELP;
        bool *bflag = nullptr;
        if (name == "headersonly" || name == "o") bflag = &cfg.headers_only;
        else if (name == "nofilename" || name == "n") bflag = &cfg.no_file_name;
        else if (name == "casesensitive" || name == "c") bflag = &cfg.case_sensitive;
        else if (name == "version" || name == "v") bflag = &cfg.show_version;
        if (bflag) {
            bool v = true;
            if (has_val) {
                if (val == "1" || val == "t" || val == "T" || val == "true" || val == "TRUE" || val == "True") v = true;
                else if (val == "0" || val == "f" || val == "F" || val == "false" || val == "FALSE" || val == "False") v = false;
                else { err = "invalid boolean value \"" + val + "\" for -" + name + ": parse error"; return PARSE_FLAG_ERROR; }
            }
            *bflag = v;
            continue;
        }
        if (name == "hash" || name == "H" || name == "name" || name == "f") {
            if (!has_val) {
                if (++i >= argc) { err = "flag needs an argument: -" + name; return PARSE_FLAG_ERROR; }
                val = argv[i];
            }
            if (name == "hash" || name == "H") hash_string = val; else cfg.name_override = val;
            continue;
        }
        err = "flag provided but not defined: -" + name;
        return PARSE_FLAG_ERROR;
    }
    if (i < argc) cfg.input = argv[i++];
    if (i < argc) cfg.output = argv[i++];
    // split on ','; validated after TrimSpace but stored untrimmed (seqhasher.go:132-137)
    size_t start = 0;
    while (true) {
        size_t comma = hash_string.find(',', start);
        cfg.hash_types.push_back(hash_string.substr(start, comma == std::string::npos ? std::string::npos : comma - start));
        if (comma == std::string::npos) break;
        start = comma + 1;
    }
    for (const std::string &ht : cfg.hash_types)
        if (!is_valid_hash_type(trim_space(ht))) {
            err = "Invalid hash type: " + ht + ". Supported types are: " + supported_list();
            return PARSE_HASH_ERROR;
        }
    return PARSE_OK;
}

bool is_blank(uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// processSequences (seqhasher.go:210-286) with the loop body on the GPU
int process_sequences(host::Source &in, FILE *out, const Config &cfg, std::string &err)
{
    std::string label = cfg.input;
    bool no_file_name = cfg.no_file_name;
    if (!cfg.name_override.empty()) label = cfg.name_override;
    else if (label == "-") no_file_name = true;   // seqhasher.go:217-219
    std::vector<int> ids;
    for (const std::string &t : cfg.hash_types) {
        int id = shb_algo_id(t.c_str());
        ids.push_back(id < 0 ? SHB_SHA1 : id);    // getHashFunc's default: arm (seqhasher.go:344-346)
    }
    if (ids.size() > 12) { err = "at most 12 hash types per run are supported by the GPU formatter"; return 1; }
    size_t cap = (size_t)64 << 20;
    phase("before create");
    shb_textproc *tp = shb_textproc_create(0, cap);
    if (!tp) { err = std::string("seqhash_b200: ") + shb_last_error(); return 1; }
    phase("textproc created");
    uint8_t *buf = shb_textproc_input(tp);
    const unsigned flags = cfg.case_sensitive ? 0u : SHB_FLAG_UPPER;
    size_t fill = 0;
    int format = 0;
    bool eof = false;
    int rc = 0;
    std::string read_err;
    while (true) {
        while (!eof && fill < cap) {
            size_t got = 0;
            try {
                got = in.read(buf + fill, cap - fill);
            } catch (const std::exception &e) {
                // A damaged or truncated input: the records that are complete in what was decoded before the
                // failure are still written (as a non-final chunk), then the error is reported — the reference's
                // reader also fails on the record it cannot finish, not on the ones before it.
                read_err = std::string("Error reading record: ") + e.what();
            }
            if (got == 0) eof = true;
            fill += got;
        }
        if (format == 0) {   // fastx.NewReaderFromIO sniffs the first non-blank byte (seqhasher.go:221)
            size_t skip = 0;
            while (skip < fill && is_blank(buf[skip])) skip++;
            if (skip == fill) {
                if (eof) break;   // empty input: no records
                fill = 0;
                continue;
            }
            if (buf[skip] == '>') format = SHB_FORMAT_FASTA;
            else if (buf[skip] == '@') format = SHB_FORMAT_FASTQ;
            else { err = "Failed to create reader: fastx: invalid FASTA/Q format"; rc = 1; break; }
            memmove(buf, buf + skip, fill - skip);
            fill -= skip;
        }
        if (eof) {
            while (fill > 0 && is_blank(buf[fill - 1])) fill--;   // trailing blank lines are not records
            if (fill == 0) break;
        }
        if (shb_textproc_run(tp, fill, format, (eof && read_err.empty()) ? 1 : 0, ids.empty() ? nullptr : ids.data(), (int)ids.size(), flags,
                             cfg.headers_only ? 1 : 0, no_file_name ? nullptr : label.c_str()) < 0) {
            err = std::string("Error reading record: ") + shb_last_error();
            rc = 1;
            break;
        }
        size_t nout = 0;
        const uint8_t *o = shb_textproc_output(tp, &nout);
        if (nout && fwrite(o, 1, nout, out) != nout) { err = "Error writing record: short write"; rc = 1; break; }
        for (size_t k = 0; k < shb_textproc_empty_records(tp) * ids.size(); k++) fprintf(stderr, "%s\n", kEmptyMsg);
        if (eof) break;
        const size_t used = shb_textproc_consumed(tp);
        if (used == 0) {   // one record larger than the buffer: grow it
            if (cap >= ((size_t)1 << 31)) { err = "Error reading record: record larger than the GPU text buffer"; rc = 1; break; }
            size_t ncap = cap * 4 > ((size_t)1 << 31) ? ((size_t)1 << 31) : cap * 4;
            shb_textproc *bigger = shb_textproc_create(0, ncap);
            if (!bigger) { err = std::string("seqhash_b200: ") + shb_last_error(); rc = 1; break; }
            memcpy(shb_textproc_input(bigger), buf, fill);
            shb_textproc_destroy(tp);
            tp = bigger;
            buf = shb_textproc_input(tp);
            cap = ncap;
            continue;
        }
        memmove(buf, buf + used, fill - used);
        fill -= used;
    }
    if (rc == 0 && !read_err.empty()) { err = read_err; rc = 1; }
    fflush(out);
    phase("records done");
    // no shb_textproc_destroy: unpinning the staging buffers costs about a second and the process is exiting
    return rc;
}

}  // namespace

int main(int argc, char **argv)
{
    // This process uses one GPU.  Unless the caller chose, hide the others: creating the CUDA context otherwise
    // initialises every device of the node (about 0.1 s each on an 8-GPU box) for a run that lasts half a second.
    setenv("CUDA_VISIBLE_DEVICES", "0", 0);
    Config cfg;
    std::string err;
    switch (parse_flags(argc, argv, cfg, err)) {
    case PARSE_HELP:
        print_usage(stderr, argv[0]);
        return 0;
    case PARSE_FLAG_ERROR:
        fprintf(stderr, "%s\n", err.c_str());
        print_usage(stderr, argv[0]);
        return 2;
    case PARSE_HASH_ERROR:
        fprintf(stderr, "%s\n", err.c_str());   // log.Fatalf("%v", err), seqhasher.go:57-59
        return 1;
    case PARSE_OK:
        break;
    }
    if (cfg.show_version) {   // seqhasher.go:72-75
        printf("SeqHasher %s\n", kVersion);
        return 0;
    }
    if (cfg.input.empty()) {  // seqhasher.go:77-80
        print_usage(stdout, argv[0]);
        return 0;
    }
    FILE *fin = cfg.input == "-" ? stdin : fopen(cfg.input.c_str(), "rb");
    if (!fin) {
        fprintf(stderr, "Error opening input: open %s: no such file or directory\n", cfg.input.c_str());
        return 1;
    }
    FILE *fout = stdout;
    if (!cfg.output.empty() && cfg.output != "-") {
        fout = fopen(cfg.output.c_str(), "wb");
        if (!fout) {
            fprintf(stderr, "Error opening output: open %s: no such file or directory\n", cfg.output.c_str());
            return 1;
        }
    }
    static char obuf[1 << 20];
    setvbuf(fout, obuf, _IOFBF, sizeof obuf);
    int rc = 1;
    try {
        auto src = host::open_maybe_compressed(fin, fin != stdin);
        rc = process_sequences(*src, fout, cfg, err);
    } catch (const std::exception &e) {
        err = std::string("Error reading record: ") + e.what();
        rc = 1;
    }
    if (rc != 0) fprintf(stderr, "%s\n", err.c_str());
    fflush(fout);
    fflush(stderr);
    _Exit(rc);   // skip CUDA teardown (see process_sequences)
}
