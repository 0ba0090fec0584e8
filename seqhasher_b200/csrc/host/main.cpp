// main.cpp — compiled host of seqhasher_b200: the CLI contract of seqhasher (seqhasher.go:56-149: flags,
// positionals, version / usage short-circuits, error texts) in front of the C ABI.  The host keeps file I/O
// and decompression; everything inside processSequences' record loop (seqhasher.go:227-283) runs on the GPU
// through shb_textproc.  Go's `flag` rules are reproduced: -x and --x alike, `-x=v` or `-x v`, parsing stops
// at the first positional ("-" included) or after "--".
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>
#include <unistd.h>

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

// Where the last record of buf[0, n) starts, so that buf[0, cut) holds complete records only (0: no such place).
// FASTA: the last line-initial '>' (the splitter's own definition of a record start).  FASTQ: the last line that
// starts with '@' and whose next-but-one line starts with '+' — a quality line may start with '@' too, but then
// the line two below it is a sequence line, which never starts with '+'.
size_t find_cut(const uint8_t *buf, size_t n, int format)
{
    if (format == SHB_FORMAT_FASTA) {
        size_t hi = n;
        while (hi > 1) {
            const void *q = memrchr(buf + 1, '>', hi - 1);
            if (!q) return 0;
            const size_t p = (size_t)((const uint8_t *)q - buf);
            if (buf[p - 1] == '\n') return p;
            hi = p;
        }
        return 0;
    }
    size_t l1 = 0, l2 = 0;   // starts of the two lines below the candidate (0 = not seen yet)
    size_t hi = n;
    while (hi > 0) {
        const void *q = hi > 1 ? memrchr(buf, '\n', hi - 1) : nullptr;   // the line break that ends the previous line
        const size_t ls = q ? (size_t)((const uint8_t *)q - buf) + 1 : 0;
        if (ls == 0) return 0;
        if (ls < n && l2 && buf[ls] == '@' && buf[l2] == '+') return ls;
        l2 = l1;
        l1 = ls;
        hi = ls;   // continue before this line's start (hi - 1 is the '\n' just found)
    }
    return 0;
}

// One chunk slot = one shb_textproc (pinned text buffer, device buffers, stream) + the host thread that runs it.
// Chunks are dealt to the slots round-robin (slot = seq % K, K = 2 per GPU) and written in chunk order, so the
// copy to the device, the kernels and the copy back of one chunk overlap the reading of the next and the writing
// of the previous one, on as many GPUs as were selected (seqhasher.go:227-283 is one sequential loop).
struct Slot {
    int device = 0;
    shb_textproc *tp = nullptr;
    size_t cap = 0;
    uint8_t *buf = nullptr;
    // the job (written by the reader before `ready` is raised)
    size_t n_bytes = 0;
    bool is_final = false;
    bool whole = false;      // the chunk holds complete records only: run it to the end, however many passes
    // the result (written by the worker before `done` is raised)
    int rc = 0;
    std::string err;
    std::vector<uint8_t> acc;   // output of all passes but the last
    size_t consumed = 0, n_empty = 0, n_nonascii = 0;
    long long ready = -1, done = -1, freed = -1;   // chunk sequence numbers
    bool quit = false;
    std::mutex mu;
    std::condition_variable cv;
    std::thread worker;
};

struct RunArgs {
    int format = 0;
    std::vector<int> ids;
    unsigned flags = 0;
    int headers_only = 0;
    const char *label = nullptr;
};

void slot_worker(Slot *s, const RunArgs *a)
{
    long long seq = -1;
    while (true) {
        {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->quit || s->ready > seq; });
            if (s->ready <= seq) return;
            seq = s->ready;
        }
        s->acc.clear();
        s->n_empty = 0;
        s->n_nonascii = 0;
        s->err.clear();
        size_t off = 0, n = s->n_bytes;
        int rc = 0;
        while (true) {
            rc = shb_textproc_run(s->tp, n, a->format, s->is_final ? 1 : 0, a->ids.empty() ? nullptr : a->ids.data(),
                                  (int)a->ids.size(), a->flags, a->headers_only, a->label);
            if (rc < 0) { s->err = shb_last_error(); break; }
            s->n_empty += shb_textproc_empty_records(s->tp);
            s->n_nonascii += shb_textproc_nonascii(s->tp);
            const size_t used = shb_textproc_consumed(s->tp);
            off += used;
            // more records than the index holds: the rest of a whole chunk goes through again
            if (rc != 0 || !s->whole || used == 0 || used >= n) break;
            size_t nout = 0;
            const uint8_t *o = shb_textproc_output(s->tp, &nout);
            s->acc.insert(s->acc.end(), o, o + nout);
            memmove(s->buf, s->buf + used, n - used);
            n -= used;
        }
        s->rc = rc;
        s->consumed = off;
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->done = seq;
        }
        s->cv.notify_all();
    }
}

// processSequences (seqhasher.go:210-286) with the loop body on the GPU
int process_sequences(host::Source &in, FILE *out, const Config &cfg, int n_gpus, std::string &err)
{
    std::string label = cfg.input;
    bool no_file_name = cfg.no_file_name;
    if (!cfg.name_override.empty()) label = cfg.name_override;
    else if (label == "-") no_file_name = true;   // seqhasher.go:217-219
    RunArgs args;
    for (const std::string &t : cfg.hash_types) {
        int id = shb_algo_id(t.c_str());
        args.ids.push_back(id < 0 ? SHB_SHA1 : id);    // getHashFunc's default: arm (seqhasher.go:344-346)
    }
    if (args.ids.size() > 48) { err = "at most 48 hash fields per record are supported by the GPU formatter"; return 1; }
    args.flags = cfg.case_sensitive ? 0u : SHB_FLAG_UPPER;
    args.headers_only = cfg.headers_only ? 1 : 0;
    args.label = no_file_name ? nullptr : label.c_str();
    size_t base_cap = (size_t)64 << 20;
    if (const char *e = getenv("SHB_CHUNK_KB")) {   // tests: small chunks exercise the chunk seams and the slot rotation
        const long kb = atol(e);
        if (kb >= 1) base_cap = (size_t)kb << 10;
    }
    phase("before create");
    const int K = 2 * n_gpus;
    std::vector<std::unique_ptr<Slot>> slots;
    for (int k = 0; k < K; k++) {
        slots.emplace_back(new Slot());
        slots[k]->device = k % n_gpus;   // consecutive chunks go to different GPUs
    }
    // slots are created on first use: a small input pays for one set of pinned buffers only
    auto ensure = [&](Slot &s, size_t need) -> bool {
        if (s.tp && s.cap > need) return true;
        size_t ncap = s.tp ? s.cap : base_cap;
        while (ncap <= need) ncap *= 4;
        if (ncap > ((size_t)1 << 31)) ncap = (size_t)1 << 31;
        if (ncap <= need) { err = "Error reading record: record larger than the GPU text buffer"; return false; }
        shb_textproc *t = shb_textproc_create(s.device, ncap);
        if (!t) { err = std::string("seqhash_b200: ") + shb_last_error(); return false; }
        if (s.tp) shb_textproc_destroy(s.tp);
        s.tp = t;
        s.cap = ncap;
        s.buf = shb_textproc_input(t);
        if (!s.worker.joinable()) s.worker = std::thread(slot_worker, &s, &args);
        return true;
    };

    // writer: chunk outputs in chunk order
    std::mutex wmu;
    std::condition_variable wcv;
    long long submitted = -1;        // last chunk sequence handed to a slot
    bool reader_done = false;
    int rc = 0;                      // first failure, in chunk order
    const bool allow_nonascii = getenv("SHB_ALLOW_NONASCII") != nullptr;
    std::atomic<bool> failed{false};
    std::thread writer([&] {
        for (long long seq = 0;; seq++) {
            {
                std::unique_lock<std::mutex> lk(wmu);
                wcv.wait(lk, [&] { return submitted >= seq || reader_done; });
                if (submitted < seq) return;
            }
            Slot &s = *slots[seq % K];
            {
                std::unique_lock<std::mutex> lk(s.mu);
                s.cv.wait(lk, [&] { return s.done >= seq; });
            }
            if (rc == 0) {
                if (s.rc < 0) {
                    err = "Error reading record: " + s.err;
                    rc = 1;
                } else if (s.n_nonascii && !allow_nonascii) {
                    // seqhasher.go:243-251: for such records the reference's strip / upper-casing follow Unicode rules; this
                    // engine is byte-wise ASCII, so its digests could differ: refuse instead (README.md:259 of the reference
                    // declares non-ASCII input unsupported)
                    err = "Error reading record: non-ASCII bytes in sequence data are not supported "
                          "(set SHB_ALLOW_NONASCII=1 to hash them as plain bytes)";
                    rc = 1;
                } else {
                    size_t nout = 0;
                    const uint8_t *o = shb_textproc_output(s.tp, &nout);
                    const bool any = shb_textproc_records(s.tp) > 0;
                    if ((!s.acc.empty() && fwrite(s.acc.data(), 1, s.acc.size(), out) != s.acc.size()) ||
                        (any && nout && fwrite(o, 1, nout, out) != nout)) {
                        err = std::string("Error writing record: ") + strerror(errno);
                        rc = 1;
                    }
                    for (size_t k = 0; k < s.n_empty * args.ids.size(); k++) fprintf(stderr, "%s\n", kEmptyMsg);
                    if (rc == 0 && s.rc == SHB_PARTIAL_FORMAT) {
                        err = "Error reading record: fastx: invalid FASTA/Q format";
                        rc = 1;
                    }
                }
                if (rc) failed = true;
            }
            {
                std::lock_guard<std::mutex> lk(s.mu);
                s.freed = seq;
            }
            s.cv.notify_all();
        }
    });

    std::vector<uint8_t> carry;      // the unconsumed tail of the previous chunk
    bool eof = false;
    std::string read_err;
    long long seq = 0;
    int setup_rc = 0;
    while (!failed) {
        Slot &s = *slots[seq % K];
        if (seq >= K) {   // wait until the slot's previous chunk has been written
            std::unique_lock<std::mutex> lk(s.mu);
            s.cv.wait(lk, [&] { return s.freed >= seq - K; });
        }
        if (!ensure(s, carry.size())) { setup_rc = 1; break; }
        if (seq == 0) phase("textproc created");
        size_t fill = carry.size();
        if (fill) memcpy(s.buf, carry.data(), fill);
        carry.clear();
        while (!eof && fill < s.cap) {
            size_t got = 0;
            try {
                got = in.read(s.buf + fill, s.cap - fill);
            } catch (const std::exception &e) {
                // A damaged or truncated input: the records that are complete in what was decoded before the
                // failure are still written (as a non-final chunk), then the error is reported — the reference's
                // reader also fails on the record it cannot finish, not on the ones before it.
                read_err = std::string("Error reading record: ") + e.what();
            }
            if (got == 0) eof = true;
            fill += got;
        }
        if (args.format == 0) {   // fastx.NewReaderFromIO sniffs the first non-blank byte (seqhasher.go:221)
            size_t skip = 0;
            while (skip < fill && is_blank(s.buf[skip])) skip++;
            if (skip == fill) {
                if (eof) break;   // empty input: no records
                continue;         // nothing but blank space so far: read on (same slot, same sequence number)
            }
            if (s.buf[skip] == '>') args.format = SHB_FORMAT_FASTA;
            else if (s.buf[skip] == '@') args.format = SHB_FORMAT_FASTQ;
            else { err = "Failed to create reader: fastx: invalid FASTA/Q format"; setup_rc = 1; break; }
            memmove(s.buf, s.buf + skip, fill - skip);
            fill -= skip;
        }
        if (eof) {
            while (fill > 0 && is_blank(s.buf[fill - 1])) fill--;   // trailing blank lines are not records
            if (fill == 0) break;
        }
        const bool last = eof;
        size_t cut = fill;
        s.whole = true;
        s.is_final = true;
        if (!(eof && read_err.empty())) {
            cut = find_cut(s.buf, fill, args.format);
            if (cut == 0) {       // one record fills the buffer (or FASTQ line structure unclear): let the GPU say how far it got
                cut = fill;
                s.whole = false;
                s.is_final = false;
            } else {
                carry.assign(s.buf + cut, s.buf + fill);
            }
        }
        s.n_bytes = cut;
        {
            std::lock_guard<std::mutex> lk(s.mu);
            s.ready = seq;
        }
        s.cv.notify_all();
        {
            std::lock_guard<std::mutex> lk(wmu);
            submitted = seq;
        }
        wcv.notify_all();
        if (!s.whole) {   // serial step: the next chunk starts where this one stopped
            std::unique_lock<std::mutex> lk(s.mu);
            s.cv.wait(lk, [&] { return s.done >= seq; });
            if (s.rc >= 0) {
                const size_t used = s.consumed;
                carry.assign(s.buf + used, s.buf + fill);   // used == 0: the next slot is grown to hold it
                if (used == 0 && last) { seq++; break; }      // nothing more will come: an unfinished record at a damaged end
            }
        }
        seq++;
        if (last) break;
    }
    {
        std::lock_guard<std::mutex> lk(wmu);
        reader_done = true;
    }
    wcv.notify_all();
    writer.join();
    for (auto &sp : slots) {
        {
            std::lock_guard<std::mutex> lk(sp->mu);
            sp->quit = true;
        }
        sp->cv.notify_all();
        if (sp->worker.joinable()) sp->worker.join();
    }
    if (rc == 0 && setup_rc) rc = setup_rc;
    if (rc == 0 && !read_err.empty()) { err = read_err; rc = 1; }
    if (fflush(out) != 0 || ferror(out)) {
        if (rc == 0) { err = std::string("Error writing record: ") + strerror(errno); rc = 1; }
    }
    phase("records done");
    // no shb_textproc_destroy: unpinning the staging buffers costs about a second and the process is exiting
    return rc;
}

// ---- plain regular files: every chunk slot reads its own byte range ----
//
// A sequential reader tops out at one thread's page-cache copy rate (~10 GB/s with four pread helpers), below what a single
// B200 takes (~13 GB/s of text), so more GPUs cannot help a stream.  A seekable plain file needs no reader at all: chunk k
// is the records that START in [first + k C, first + (k + 1) C); its slot preads that range (plus the tail of the record that
// crosses the upper bound) straight into its pinned buffer and finds both boundaries itself with the splitter's own rule — a
// FASTA record starts at a line-initial '>', a FASTQ record at a line that starts with '@' and whose next-but-one line starts
// with '+'.  Neighbouring slots compute the shared boundary from the same bytes with the same rule, so no record is lost or
// taken twice; the writer emits the chunks in order (seqhasher.go:227-283 is one sequential loop).

constexpr size_t kNpos = (size_t)-1;
constexpr size_t kNeedMore = (size_t)-2;

// first record start at buffer index >= from (from > 0: buf[from - 1] is readable); kNpos: none in buf[0, n); kNeedMore:
// a FASTQ candidate whose next-but-one line lies beyond the buffer (only when !at_eof)
size_t find_record_start(const uint8_t *buf, size_t n, size_t from, int format, bool at_eof)
{
    if (format == SHB_FORMAT_FASTA) {
        size_t p = from;
        while (p < n) {
            const void *q = memchr(buf + p, '>', n - p);
            if (!q) return kNpos;
            p = (size_t)((const uint8_t *)q - buf);
            if (p == 0 || buf[p - 1] == '\n') return p;
            p++;
        }
        return kNpos;
    }
    // FASTQ: walk the line starts from `from` on
    size_t ls = from;
    if (from > 0 && buf[from - 1] != '\n') {
        const void *q = memchr(buf + from, '\n', n - from);
        if (!q) return at_eof ? kNpos : kNeedMore;
        ls = (size_t)((const uint8_t *)q - buf) + 1;
    }
    while (ls < n) {
        const void *q1 = memchr(buf + ls, '\n', n - ls);
        if (!q1) return at_eof ? kNpos : kNeedMore;
        const size_t l1 = (size_t)((const uint8_t *)q1 - buf) + 1;
        if (buf[ls] == '@') {
            const void *q2 = l1 < n ? memchr(buf + l1, '\n', n - l1) : nullptr;
            if (!q2) { if (at_eof) return kNpos; return kNeedMore; }
            const size_t l2 = (size_t)((const uint8_t *)q2 - buf) + 1;
            if (l2 >= n) { if (at_eof) return kNpos; return kNeedMore; }
            if (buf[l2] == '+') return ls;
        }
        ls = l1;
    }
    return kNpos;
}

struct RangeJob {
    int fd = -1;
    size_t file_size = 0, first = 0, C = 0;
};

struct RangeSlot {
    int device = 0;
    shb_textproc *tp = nullptr;
    size_t cap = 0;
    uint8_t *buf = nullptr;
    int rc = 0;
    std::string err;
    std::vector<uint8_t> acc;
    size_t n_empty = 0, n_nonascii = 0;
    long long ready = -1, done = -1, freed = -1;
    bool quit = false;
    std::mutex mu;
    std::condition_variable cv;
    std::thread worker;
};

size_t pread_full(int fd, uint8_t *dst, size_t n, size_t off)
{
    size_t got = 0;
    while (got < n) {
        const ssize_t r = pread(fd, dst + got, n - got, (off_t)(off + got));
        if (r <= 0) break;
        got += (size_t)r;
    }
    return got;
}

// what a slot does with chunk `seq`: returns the rc of the last run; fills s->acc (output of all runs but the last, which stays
// in s->tp), s->err and the counters
using ChunkFn = std::function<int(RangeSlot *, long long)>;

void slot_thread(RangeSlot *s, size_t cap0, const ChunkFn *fn, bool text_staging)
{
    // the slot's textproc is created here, on its own thread: the CUDA contexts of several devices come up side by side
    s->tp = shb_textproc_create(s->device, cap0);
    if (s->tp) { s->cap = cap0; s->buf = text_staging ? shb_textproc_input(s->tp) : nullptr; }   // BGZF slots never pin a text buffer
    else s->err = std::string("seqhash_b200: ") + shb_last_error();
    if (s->tp && text_staging && !s->buf) { s->err = std::string("seqhash_b200: ") + shb_last_error(); shb_textproc_destroy(s->tp); s->tp = nullptr; }
    long long seq = -1;
    while (true) {
        {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->quit || s->ready > seq; });
            if (s->ready <= seq) return;
            seq = s->ready;
        }
        s->acc.clear();
        s->n_empty = s->n_nonascii = 0;
        int rc = -1;
        if (s->tp) {
            s->err.clear();
            rc = (*fn)(s, seq);
        }
        s->rc = rc;
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->done = seq;
        }
        s->cv.notify_all();
    }
}

int range_chunk(RangeSlot *s, long long seq, const RunArgs *a, const RangeJob *job)
{
    int rc = 0;
    {
        {
            s->err.clear();
            const size_t lo_nom = job->first + (size_t)seq * job->C;
            const size_t hi_nom = std::min(job->file_size, lo_nom + job->C);
            const size_t read_lo = seq == 0 ? lo_nom : lo_nom - 1;
            size_t n = pread_full(job->fd, s->buf, std::min(s->cap, job->file_size - read_lo), read_lo);
            bool at_eof = read_lo + n >= job->file_size;
            auto grow = [&]() -> bool {   // more room (a record longer than the slack): a bigger textproc, then read on
                if (s->cap >= ((size_t)1 << 31)) return false;
                const size_t ncap = std::min((size_t)1 << 31, s->cap * 4);
                shb_textproc *t = shb_textproc_create(s->device, ncap);
                if (!t) return false;
                memcpy(shb_textproc_input(t), s->buf, n);
                shb_textproc_destroy(s->tp);
                s->tp = t; s->cap = ncap; s->buf = shb_textproc_input(t);
                n += pread_full(job->fd, s->buf + n, std::min(s->cap - n, job->file_size - (read_lo + n)), read_lo + n);
                at_eof = read_lo + n >= job->file_size;
                return true;
            };
            size_t start = 0, end = 0;
            bool ok = true;
            if (seq > 0) {
                while (true) {
                    start = find_record_start(s->buf, n, 1, a->format, at_eof);
                    if (start != kNeedMore) break;
                    if (!grow()) { ok = false; break; }
                }
                if (ok && (start == kNpos || read_lo + start >= hi_nom)) start = kNpos;   // no record starts in this range
            }
            if (ok && start != kNpos) {
                if (hi_nom >= job->file_size) {
                    while (!at_eof) { if (!grow()) { ok = false; break; } }
                    end = n;
                    while (end > start && is_blank(s->buf[end - 1])) end--;   // trailing blank lines are not records
                } else {
                    while (true) {
                        end = find_record_start(s->buf, n, hi_nom - read_lo, a->format, at_eof);
                        if (end != kNeedMore && !(end == kNpos && !at_eof)) break;
                        if (!grow()) { ok = false; break; }
                    }
                    if (ok && end == kNpos) {   // the file ends inside this chunk's last record
                        end = n;
                        while (end > start && is_blank(s->buf[end - 1])) end--;
                    }
                }
            }
            if (!ok) {
                rc = -1;
                s->err = "record larger than the GPU text buffer";
            } else if (start != kNpos && end > start) {
                if (start) memmove(s->buf, s->buf + start, end - start);
                size_t len = end - start;
                while (true) {
                    rc = shb_textproc_run(s->tp, len, a->format, 1, a->ids.empty() ? nullptr : a->ids.data(), (int)a->ids.size(),
                                          a->flags, a->headers_only, a->label);
                    if (rc < 0) { s->err = shb_last_error(); break; }
                    s->n_empty += shb_textproc_empty_records(s->tp);
                    s->n_nonascii += shb_textproc_nonascii(s->tp);
                    const size_t used = shb_textproc_consumed(s->tp);
                    if (rc != 0 || used == 0 || used >= len) break;   // else: record index full, run the rest
                    size_t nout = 0;
                    const uint8_t *o = shb_textproc_output(s->tp, &nout);
                    s->acc.insert(s->acc.end(), o, o + nout);
                    memmove(s->buf, s->buf + used, len - used);
                    len -= used;
                }
            } else {
                shb_textproc_run(s->tp, 0, a->format, 1, nullptr, 0, a->flags, a->headers_only, a->label);   // clears the last output
            }
        }
    }
    return rc;
}

int run_chunk_slots(long long nchunks, int n_gpus, size_t cap0, const ChunkFn &fn, FILE *out, const RunArgs &args, std::string &err, bool text_staging = true);

int process_file_ranges(int fd, size_t file_size, FILE *out, const Config &cfg, int n_gpus, std::string &err)
{
    std::string label = cfg.input;
    bool no_file_name = cfg.no_file_name;
    if (!cfg.name_override.empty()) label = cfg.name_override;
    RunArgs args;
    for (const std::string &t : cfg.hash_types) {
        int id = shb_algo_id(t.c_str());
        args.ids.push_back(id < 0 ? SHB_SHA1 : id);
    }
    if (args.ids.size() > 48) { err = "at most 48 hash fields per record are supported by the GPU formatter"; return 1; }
    args.flags = cfg.case_sensitive ? 0u : SHB_FLAG_UPPER;
    args.headers_only = cfg.headers_only ? 1 : 0;
    args.label = no_file_name ? nullptr : label.c_str();
    RangeJob job;
    job.fd = fd;
    job.file_size = file_size;
    job.C = (size_t)64 << 20;
    if (const char *e = getenv("SHB_CHUNK_KB")) {
        const long kb = atol(e);
        if (kb >= 1) job.C = (size_t)kb << 10;
    }
    // sniff: the first non-blank byte decides (fastx.NewReaderFromIO, seqhasher.go:221)
    {
        uint8_t head[4096];
        size_t pos = 0;
        bool found = false;
        while (pos < file_size && !found) {
            const size_t got = pread_full(fd, head, std::min(sizeof head, file_size - pos), pos);
            if (got == 0) break;
            for (size_t i = 0; i < got; i++)
                if (!is_blank(head[i])) { job.first = pos + i; found = true; break; }
            pos += got;
        }
        if (!found) return 0;   // empty input: no records
        uint8_t c = 0;
        pread_full(fd, &c, 1, job.first);
        if (c == '>') args.format = SHB_FORMAT_FASTA;
        else if (c == '@') args.format = SHB_FORMAT_FASTQ;
        else { err = "Failed to create reader: fastx: invalid FASTA/Q format"; return 1; }
    }
    const long long nchunks = (long long)((file_size - job.first + job.C - 1) / job.C);
    const ChunkFn fn = [&](RangeSlot *s, long long seq) { return range_chunk(s, seq, &args, &job); };
    return run_chunk_slots(nchunks, n_gpus, job.C + job.C / 4 + 4096, fn, out, args, err);
}

// Chunks 0..nchunks-1 dealt round-robin to K slots (each a thread + a textproc), their outputs written in chunk order.
int run_chunk_slots(long long nchunks, int n_gpus, size_t cap0, const ChunkFn &fn, FILE *out, const RunArgs &args, std::string &err, bool text_staging)
{
    // Slots: a slot's pread is one thread copying out of the page cache (5-6 GB/s) and a B200 takes ~13 GB/s of text, so one
    // GPU gets four; with more GPUs two each are enough readers, and every slot costs ~0.1 s to set up (80 MB of pinned
    // memory, ~0.5 GB of device buffers) — 32 slots on an 8-GPU box delayed the first chunk by 2.6 s.
    int per_gpu = n_gpus == 1 ? 4 : 2;
    if (const char *e = getenv("SHB_SLOTS_PER_GPU")) { const int v = atoi(e); if (v >= 1 && v <= 16) per_gpu = v; }
    const int K = (int)std::max<long long>(1, std::min<long long>((long long)per_gpu * n_gpus, nchunks));
    phase("before create");
    std::vector<std::unique_ptr<RangeSlot>> slots;
    for (int k = 0; k < K; k++) {
        slots.emplace_back(new RangeSlot());
        slots[k]->device = k % n_gpus;
        slots[k]->worker = std::thread(slot_thread, slots[k].get(), cap0, &fn, text_staging);
    }
    std::mutex wmu;
    std::condition_variable wcv;
    long long submitted = -1;
    bool dispatch_done = false;
    int rc = 0;
    const bool allow_nonascii = getenv("SHB_ALLOW_NONASCII") != nullptr;
    std::atomic<bool> failed{false};
    std::thread writer([&] {
        for (long long seq = 0;; seq++) {
            {
                std::unique_lock<std::mutex> lk(wmu);
                wcv.wait(lk, [&] { return submitted >= seq || dispatch_done; });
                if (submitted < seq) return;
            }
            RangeSlot &s = *slots[seq % K];
            {
                std::unique_lock<std::mutex> lk(s.mu);
                s.cv.wait(lk, [&] { return s.done >= seq; });
            }
            if (seq == 0) phase("first chunk done");
            if (rc == 0) {
                if (s.rc < 0) {
                    err = (s.tp ? "Error reading record: " : "") + s.err;
                    rc = 1;
                } else if (s.n_nonascii && !allow_nonascii) {
                    err = "Error reading record: non-ASCII bytes in sequence data are not supported "
                          "(set SHB_ALLOW_NONASCII=1 to hash them as plain bytes)";
                    rc = 1;
                } else {
                    size_t nout = 0;
                    const uint8_t *o = shb_textproc_output(s.tp, &nout);
                    if ((!s.acc.empty() && fwrite(s.acc.data(), 1, s.acc.size(), out) != s.acc.size()) ||
                        (nout && fwrite(o, 1, nout, out) != nout)) {
                        err = std::string("Error writing record: ") + strerror(errno);
                        rc = 1;
                    }
                    for (size_t k = 0; k < s.n_empty * args.ids.size(); k++) fprintf(stderr, "%s\n", kEmptyMsg);
                    if (rc == 0 && s.rc == SHB_PARTIAL_FORMAT) {
                        err = "Error reading record: fastx: invalid FASTA/Q format";
                        rc = 1;
                    }
                }
                if (rc) failed = true;
            }
            {
                std::lock_guard<std::mutex> lk(s.mu);
                s.freed = seq;
            }
            s.cv.notify_all();
        }
    });
    for (long long seq = 0; seq < nchunks && !failed; seq++) {
        RangeSlot &s = *slots[seq % K];
        if (seq >= K) {
            std::unique_lock<std::mutex> lk(s.mu);
            s.cv.wait(lk, [&] { return s.freed >= seq - K; });
        }
        {
            std::lock_guard<std::mutex> lk(s.mu);
            s.ready = seq;
        }
        s.cv.notify_all();
        {
            std::lock_guard<std::mutex> lk(wmu);
            submitted = seq;
        }
        wcv.notify_all();
    }
    {
        std::lock_guard<std::mutex> lk(wmu);
        dispatch_done = true;
    }
    wcv.notify_all();
    writer.join();
    for (auto &sp : slots) {
        {
            std::lock_guard<std::mutex> lk(sp->mu);
            sp->quit = true;
        }
        sp->cv.notify_all();
        if (sp->worker.joinable()) sp->worker.join();
    }
    if (fflush(out) != 0 || ferror(out)) {
        if (rc == 0) { err = std::string("Error writing record: ") + strerror(errno); rc = 1; }
    }
    phase("records done");
    return rc;
}

// ---- BGZF files (bgzip, htslib): the members go to the GPU compressed and are inflated there ----
//
// The reference reads .gz through xopen -> pgzip on one goroutine (go.mod:20-31).  Blocked gzip is the one gzip container the
// host can cut without decoding: every member announces its size in a 'BC' extra field and holds <= 64 KiB of text.  The host
// walks the member headers once (one 26-byte pread per member), groups the members into chunks of ~64 MiB of text, and
// every chunk slot preads its members' bytes into pinned memory and calls shb_textproc_run_bgzf: one warp inflates one member
// (csrc/inflate.cuh), the splitter runs on the inflated text, and a third to a quarter of the bytes cross PCIe.  A file that
// is not BGZF from its first byte to its last takes the host-side inflate (pgzip.hpp) as before.

struct BgzfMember {
    uint64_t start, data, clen, packed;   // first byte of the member, first byte and length of its DEFLATE data, ISIZE | CRC32 << 32
};

// Every member of the file, or false when the bytes are not BGZF throughout (same rules as host/bgzf.py scan_members).
// One small pread per member: the 8-byte trailer of member k and the header of member k + 1 are adjacent.  (An mmap of the
// file did the same walk in 0.05 s on one box and 0.7 s on another, whose file system served the page faults slowly.)
bool scan_bgzf(int fd, size_t n, std::vector<BgzfMember> &out)
{
    size_t p = 0;
    uint8_t hb[8 + 512];
    if (pread_full(fd, hb + 8, std::min<size_t>(512, n), 0) < 18) return false;
    size_t have = std::min<size_t>(512, n);   // valid bytes at hb + 8 = file bytes [p, p + have)
    while (p < n) {
        const uint8_t *b = hb + 8;
        if (n - p < 18 + 8 || have < 18 || b[0] != 0x1f || b[1] != 0x8b || b[2] != 8 || b[3] != 4) return false;
        const size_t xlen = b[10] | ((size_t)b[11] << 8);
        if (12 + xlen > have) return false;                   // BGZF headers are 18 bytes; 500 covers any sane extra field
        size_t q = 12;
        const size_t xend = q + xlen;
        long bsize = -1;
        while (q + 4 <= xend) {
            const size_t slen = b[q + 2] | ((size_t)b[q + 3] << 8);
            if (b[q] == 'B' && b[q + 1] == 'C' && slen == 2 && q + 6 <= xend) bsize = b[q + 4] | ((long)b[q + 5] << 8);
            q += 4 + slen;
        }
        if (bsize < 0) return false;
        const size_t total = (size_t)bsize + 1;
        if (total < 12 + xlen + 8 || p + total > n) return false;
        // the member's trailer and the next member's header in one read
        const size_t next = p + total;
        const size_t want = std::min<size_t>(8 + 512, n - (next - 8));
        if (pread_full(fd, hb, want, next - 8) != want) return false;
        uint32_t crc, isize;
        memcpy(&crc, hb, 4);
        memcpy(&isize, hb + 4, 4);
        if (isize > 65536) return false;
        out.push_back({p, p + 12 + xlen, total - 12 - xlen - 8, (uint64_t)isize | ((uint64_t)crc << 32)});
        p = next;
        have = want - 8;
    }
    return !out.empty();
}

struct BgzfJob {
    int fd = -1;
    std::vector<BgzfMember> m;
    std::vector<size_t> first;        // first member of chunk k; first[nchunks] = m.size()
    long long first_start = 0;        // text offset of the first non-blank byte inside member first[0]
    size_t lookahead = 256 << 10;     // text bytes of look-ahead members a chunk starts with
    size_t file_size = 0;
};

int bgzf_chunk(RangeSlot *s, long long seq, const RunArgs *a, const BgzfJob *job)
{
    const size_t lo_m = job->first[(size_t)seq], hi_m = job->first[(size_t)seq + 1], nm = job->m.size();
    size_t want_extra = job->lookahead;
    std::vector<uint64_t> tab;
    int rc = 0;
    while (true) {
        size_t e = hi_m, extra_text = 0;
        while (e < nm && (extra_text < want_extra || e == hi_m)) extra_text += (uint32_t)job->m[e++].packed;
        size_t text = extra_text;
        for (size_t i = lo_m; i < hi_m; i++) text += (uint32_t)job->m[i].packed;
        if (text + 64 > s->cap) {   // a record longer than the slack: a bigger textproc
            if (s->cap >= ((size_t)1 << 31)) { s->err = "record larger than the GPU text buffer"; return -1; }
            size_t ncap = s->cap;
            while (ncap < text + 64 && ncap < ((size_t)1 << 31)) ncap *= 2;
            if (ncap < text + 64) { s->err = "record larger than the GPU text buffer"; return -1; }
            shb_textproc *t = shb_textproc_create(s->device, ncap);
            if (!t) { s->err = shb_last_error(); return -1; }
            shb_textproc_destroy(s->tp);
            s->tp = t; s->cap = ncap; s->buf = nullptr;
        }
        const size_t lo = job->m[lo_m].start, hi = job->m[e - 1].data + job->m[e - 1].clen + 8;
        uint8_t *stage = shb_textproc_bgzf_input(s->tp, hi - lo + 64);
        if (!stage) { s->err = shb_last_error(); return -1; }
        if (pread_full(job->fd, stage, hi - lo, lo) != hi - lo) { s->err = "unexpected EOF"; return -1; }
        tab.resize(3 * (e - lo_m));
        for (size_t i = lo_m; i < e; i++) {
            tab[3 * (i - lo_m)] = job->m[i].data - lo;
            tab[3 * (i - lo_m) + 1] = job->m[i].clen;
            tab[3 * (i - lo_m) + 2] = job->m[i].packed;
        }
        rc = shb_textproc_run_bgzf(s->tp, hi - lo, tab.data(), hi_m - lo_m, e - hi_m, seq == 0 ? job->first_start : -1, e == nm ? 1 : 0,
                                   a->format, a->ids.empty() ? nullptr : a->ids.data(), (int)a->ids.size(), a->flags, a->headers_only, a->label);
        if (rc != SHB_NEED_MORE) break;
        if (e == nm) { s->err = "BGZF look-ahead exhausted"; return -1; }   // cannot happen: at_eof was set
        want_extra *= 4;
    }
    while (true) {
        if (rc < 0) { s->err = shb_last_error(); return rc; }
        s->n_empty += shb_textproc_empty_records(s->tp);
        s->n_nonascii += shb_textproc_nonascii(s->tp);
        if (rc != 0 || !shb_textproc_pending(s->tp)) return rc;
        size_t nout = 0;                                     // record index full: keep this output, run the rest
        const uint8_t *o = shb_textproc_output(s->tp, &nout);
        s->acc.insert(s->acc.end(), o, o + nout);
        rc = shb_textproc_resume(s->tp, a->format, a->ids.empty() ? nullptr : a->ids.data(), (int)a->ids.size(), a->flags,
                                 a->headers_only, a->label);
    }
}

// the member table of a BGZF file, or null: not a regular file, not BGZF throughout, or the device path is switched off
// (SHB_NO_BGZF) — the caller then takes the host-side inflate
std::unique_ptr<BgzfJob> open_bgzf(int fd)
{
    struct stat st;
    if (getenv("SHB_NO_BGZF") || fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 28) return nullptr;
    const size_t file_size = (size_t)st.st_size;
    uint8_t head[18];
    if (pread_full(fd, head, 18, 0) != 18 || head[0] != 0x1f || head[1] != 0x8b || head[2] != 8 || head[3] != 4) return nullptr;
    std::unique_ptr<BgzfJob> job(new BgzfJob());
    job->fd = fd;
    job->file_size = file_size;
    if (!scan_bgzf(fd, file_size, job->m)) return nullptr;
    return job;
}

int process_bgzf(BgzfJob &job, FILE *out, const Config &cfg, int n_gpus, std::string &err)
{
    phase("bgzf scanned");
    RunArgs args;
    std::string label = cfg.input;
    bool no_file_name = cfg.no_file_name;
    if (!cfg.name_override.empty()) label = cfg.name_override;
    for (const std::string &t : cfg.hash_types) {
        int id = shb_algo_id(t.c_str());
        args.ids.push_back(id < 0 ? SHB_SHA1 : id);
    }
    if (args.ids.size() > 48) { err = "at most 48 hash fields per record are supported by the GPU formatter"; return 1; }
    args.flags = cfg.case_sensitive ? 0u : SHB_FLAG_UPPER;
    args.headers_only = cfg.headers_only ? 1 : 0;
    args.label = no_file_name ? nullptr : label.c_str();
    // sniff (fastx.NewReaderFromIO, seqhasher.go:221): the first non-blank byte of the text, found by inflating members on the
    // host until one holds it; members of nothing but blank space before it belong to no chunk
    size_t m0 = 0;
    bool found = false;
    {
        std::vector<uint8_t> txt(65536 + 16);
        for (; m0 < job.m.size() && !found; m0++) {
            const BgzfMember &mb = job.m[m0];
            const uint32_t isize = (uint32_t)mb.packed;
            if (isize == 0) continue;
            std::vector<uint8_t> cbuf(mb.clen + 8);
            if (pread_full(job.fd, cbuf.data(), mb.clen, mb.data) != mb.clen) { err = "Error reading record: unexpected EOF"; return 1; }
            z_stream z{};
            if (inflateInit2(&z, -15) != Z_OK) { err = "Error reading record: gzip: init failed"; return 1; }
            z.next_in = cbuf.data();
            z.avail_in = (uInt)mb.clen;
            z.next_out = txt.data();
            z.avail_out = (uInt)txt.size();
            const int zr = inflate(&z, Z_FINISH);
            const size_t got = txt.size() - z.avail_out;
            inflateEnd(&z);
            if (zr != Z_STREAM_END || got != isize) { err = "Error reading record: flate: corrupt input"; return 1; }
            for (size_t i = 0; i < got; i++) {
                if (!is_blank(txt[i])) {
                    found = true;
                    job.first_start = (long long)i;
                    if (txt[i] == '>') args.format = SHB_FORMAT_FASTA;
                    else if (txt[i] == '@') args.format = SHB_FORMAT_FASTQ;
                    else { err = "Failed to create reader: fastx: invalid FASTA/Q format"; return 1; }
                    break;
                }
            }
            if (found) break;
        }
    }
   
    if (!found) return 0;   // empty input: no records
    size_t C = (size_t)64 << 20;
    if (const char *e = getenv("SHB_CHUNK_KB")) { const long kb = atol(e); if (kb >= 1) C = (size_t)kb << 10; }
    if (C < job.lookahead) job.lookahead = std::max<size_t>(C, 1024);
    size_t acc = 0;
    job.first.push_back(m0);
    for (size_t i = m0; i < job.m.size(); i++) {
        const uint32_t isize = (uint32_t)job.m[i].packed;
        if (acc && acc + isize > C) { job.first.push_back(i); acc = 0; }
        acc += isize;
    }
    job.first.push_back(job.m.size());
    const long long nchunks = (long long)job.first.size() - 1;
    const ChunkFn fn = [&](RangeSlot *s, long long seq) { return bgzf_chunk(s, seq, &args, &job); };
    return run_chunk_slots(nchunks, n_gpus, C + 65536 + 4 * job.lookahead + 4096, fn, out, args, err, false);
}

// How many GPUs this run uses: SHB_GPUS if set; otherwise one per 12 GiB of a plain regular file — bringing up one more
// device costs 0.3-0.6 s of CUDA initialisation on a cold driver, the time one GPU needs for 4-8 GB of text — and one for
// anything else (a stream is bound by its reader or its host-side inflate, not by the GPU).
int pick_gpus(FILE *fin, bool regular_plain)
{
    if (const char *e = getenv("SHB_GPUS")) {
        const int v = atoi(e);
        if (v > 0) return v;
    }
    if (!regular_plain) return 1;
    struct stat st;
    if (fstat(fileno(fin), &st) != 0 || !S_ISREG(st.st_mode)) return 1;
    const long long per = 12LL << 30;
    long long n = (st.st_size + per - 1) / per;
    return (int)(n < 1 ? 1 : n > 8 ? 8 : n);
}

std::string go_errno_text(int e)
{
    std::string t = strerror(e);
    if (!t.empty() && t[0] >= 'A' && t[0] <= 'Z') t[0] = (char)(t[0] - 'A' + 'a');   // Go's syscall error texts are lower-case
    return t;
}

}  // namespace

int main(int argc, char **argv)
{
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
        fprintf(stderr, "Error opening input: open %s: %s\n", cfg.input.c_str(), go_errno_text(errno).c_str());
        return 1;
    }
    FILE *fout = stdout;
    if (!cfg.output.empty() && cfg.output != "-") {
        fout = fopen(cfg.output.c_str(), "wb");
        if (!fout) {
            fprintf(stderr, "Error opening output: open %s: %s\n", cfg.output.c_str(), go_errno_text(errno).c_str());
            return 1;
        }
    }
    static char obuf[1 << 20];
    setvbuf(fout, obuf, _IOFBF, sizeof obuf);
    int rc = 1;
    try {
        bool compressed = false;
        std::unique_ptr<host::Source> src;
        std::unique_ptr<BgzfJob> bgzf = fin != stdin ? open_bgzf(fileno(fin)) : nullptr;
        if (bgzf) compressed = true;   // no host-side reader at all: the members are inflated on the GPU
        else src = host::open_maybe_compressed(fin, fin != stdin, &compressed);
        int n_gpus = pick_gpus(fin, fin != stdin && !compressed);
        // Unless the caller chose, hide the GPUs this run does not use: creating the CUDA context otherwise initialises
        // every device of the node (about 0.1 s each on an 8-GPU box) for a run that may last half a second.
        if (!getenv("CUDA_VISIBLE_DEVICES")) {
            std::string vis;
            for (int d = 0; d < n_gpus; d++) vis += (d ? "," : "") + std::to_string(d);
            setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
        }
        const int have = shb_init(0, 0);
        if (have <= 0) {
            err = std::string("seqhash_b200: ") + shb_last_error();
        } else {
            if (n_gpus > have) n_gpus = have;
            struct stat st;
            const bool plain_file = fin != stdin && !compressed && fstat(fileno(fin), &st) == 0 && S_ISREG(st.st_mode);
            size_t chunk = (size_t)64 << 20;
            if (const char *e = getenv("SHB_CHUNK_KB")) { const long kb = atol(e); if (kb >= 1) chunk = (size_t)kb << 10; }
            // a seekable plain file of several chunks: every slot reads its own byte range (no sequential reader)
            if (bgzf)
                rc = process_bgzf(*bgzf, fout, cfg, n_gpus, err);
            else if (plain_file && !getenv("SHB_NO_RANGES") && (n_gpus > 1 || (size_t)st.st_size >= 4 * chunk))
                rc = process_file_ranges(fileno(fin), (size_t)st.st_size, fout, cfg, n_gpus, err);
            else
                rc = process_sequences(*src, fout, cfg, n_gpus, err);
        }
    } catch (const std::exception &e) {
        err = std::string("Error reading record: ") + e.what();
        rc = 1;
    }
    if (rc != 0) fprintf(stderr, "%s\n", err.c_str());
    fflush(fout);
    fflush(stderr);
    _Exit(rc);   // skip CUDA teardown (see process_sequences)
}
