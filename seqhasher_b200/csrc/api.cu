// api.cu — the C ABI of include/seqhash_b200.h: device selection, batch objects (pinned staging,
// device buffers, one stream each), the submit/wait pipeline and the device-resident entry point.
// There is no CPU fallback anywhere in this file: without a usable sm_100 device every call fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/seqhash_b200.h"
#include "kernels.cuh"
#include "parse.cuh"

namespace {

thread_local std::string t_err;
std::mutex g_mu;
int g_ndev = 0;   // devices 0..g_ndev-1 in use; 0 = not initialised

int fail(int code, const std::string &msg)
{
    t_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what)
{
    return fail(SHB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                          \
    do {                                                  \
        cudaError_t e_ = (call);                          \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

const char *const kNames[SHB_N_ALGOS] = {"sha1", "sha3", "md5", "xxhash", "xxh3", "cityhash",
                                         "murmur3", "nthash", "blake3", "k12", "rapidhash", "rapidhash32"};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// scratch layout of the normalise pre-pass inside one allocation
struct ScratchLayout {
    size_t res_off, offs_off, counts_off, prefix_off, scan_off, perm_off, b3_off, total;
    ScratchLayout(size_t total_bytes, size_t n)
    {
        const size_t nb = shb::norm_blocks(total_bytes);
        res_off = 0;
        offs_off = align_up(total_bytes + 32, 256);
        counts_off = offs_off + align_up((n + 1) * sizeof(int64_t), 256);
        prefix_off = counts_off + align_up(nb * sizeof(uint32_t), 256);
        scan_off = prefix_off + align_up((nb + 1) * sizeof(int64_t), 256);
        perm_off = scan_off + align_up(shb::scan_tmp_bytes(nb), 256);        // length-sorted record order
        b3_off = perm_off + align_up(shb::length_sort_bytes(n), 256);        // BLAKE3 / K12 long path: counts, bases, CVs
        total = b3_off + shb::b3_long_scratch_bytes(total_bytes, n);
    }
};

// The hashing pipeline on device-resident buffers; shared by shb_submit and shb_hash_device.
int run_pipeline(const uint8_t *d_res, const int64_t *d_offs, size_t n, size_t total_bytes, const int *algos,
                 int n_algos, unsigned flags, uint8_t *const *d_digests, int64_t *d_out_len, void *d_scratch,
                 cudaStream_t st, const uint8_t **norm_res, const int64_t **norm_offs)
{
    const bool upper = flags & SHB_FLAG_UPPER, strip = flags & SHB_FLAG_STRIP, emit = flags & SHB_FLAG_EMIT_SEQ;
    const uint8_t *res = d_res;
    const int64_t *offs = d_offs;
    bool fused_upper = upper;
    // Path choice.  Short records (mean <= ~740 B): shared-memory tiles of 128 records, sized so that at
    // least two CTAs fit per SM; the HBM-bound hashes requested together are fused into one launch so the
    // residues are read once, and the whitespace strip is fused into the tile pass.  Longer records: one
    // thread per record straight from global memory, or the long-record kernels (mean >= 2 KiB).
    const size_t mean = n ? total_bytes / n : 0;
    const size_t tile_need = (mean * 128 * 17 / 16 + 1024 + 1023) / 1024 * 1024;   // 6 % slack + 1 KiB: oversize tiles fall back per CTA
    const bool tile_mode = n && tile_need <= 100 * 1024 && !(flags & SHB_FLAG_FORCE_DIRECT);
    const bool prepass = emit || (strip && !tile_mode);   // materialise the normalised batch first
    uint8_t *fused_strip_scratch = nullptr;
    int64_t *fused_out_len = nullptr;
    if ((strip || emit) && !d_scratch) return fail(SHB_ERR_BAD_ARG, "SHB_FLAG_STRIP/EMIT_SEQ need a scratch buffer");
    if (prepass) {
        ScratchLayout L(total_bytes, n);
        uint8_t *base = static_cast<uint8_t *>(d_scratch);
        uint8_t *o_res = base + L.res_off;
        int64_t *o_offs = reinterpret_cast<int64_t *>(base + L.offs_off);
        CU(shb::launch_normalise(d_res, d_offs, n, total_bytes, strip, upper,
                                 reinterpret_cast<uint32_t *>(base + L.counts_off),
                                 reinterpret_cast<int64_t *>(base + L.prefix_off), o_res, o_offs, d_out_len,
                                 base + L.scan_off, st));
        res = o_res;
        offs = o_offs;
        fused_upper = false;   // already applied while compacting
    } else if (strip) {
        fused_strip_scratch = static_cast<uint8_t *>(d_scratch) + ScratchLayout(total_bytes, n).res_off;
        fused_out_len = d_out_len;
    } else if (d_out_len) {
        CU(shb::launch_lengths(d_offs, n, d_out_len, st));
    }
    if (norm_res) *norm_res = res;
    if (norm_offs) *norm_offs = offs;
    uint8_t *slot_of[SHB_N_ALGOS] = {};
    uint32_t want = 0;
    for (int k = 0; k < n_algos; k++)
        if (!slot_of[algos[k]]) { slot_of[algos[k]] = d_digests[k]; want |= 1u << algos[k]; }
    if (tile_mode) {
        const uint32_t cap = (uint32_t)(tile_need < 8192 ? 8192 : tile_need);
        const uint32_t fast = want & shb::tile_fast_mask();
        if (fast) {
            CU(shb::launch_hash_tile(fast, fused_upper, res, offs, n, slot_of, cap, fused_strip_scratch, fused_out_len, st));
            fused_out_len = nullptr;   // lengths are written by the first launch only
        }
        for (int a = 0; a < SHB_N_ALGOS; a++)
            if ((want & ~fast) >> a & 1u) {
                CU(shb::launch_hash_tile(1u << a, fused_upper, res, offs, n, slot_of, cap, fused_strip_scratch, fused_out_len, st));
                fused_out_len = nullptr;
            }
        if (fused_out_len)   // strip requested with no algorithm at all: lengths still have to be produced
            CU(shb::launch_normalise(d_res, d_offs, n, total_bytes, true, false,
                                     reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(d_scratch) + ScratchLayout(total_bytes, n).counts_off),
                                     reinterpret_cast<int64_t *>(static_cast<uint8_t *>(d_scratch) + ScratchLayout(total_bytes, n).prefix_off),
                                     static_cast<uint8_t *>(d_scratch) + ScratchLayout(total_bytes, n).res_off,
                                     reinterpret_cast<int64_t *>(static_cast<uint8_t *>(d_scratch) + ScratchLayout(total_bytes, n).offs_off),
                                     d_out_len, static_cast<uint8_t *>(d_scratch) + ScratchLayout(total_bytes, n).scan_off, st));
    } else {
        // Kernels with parallelism inside a record pay off for multi-kilobyte records only: below the thresholds the
        // thread-per-record kernels win (length-sorted visiting order, 128-bit stream loads, shared-memory seed table).
        const bool long_ok = n && !(flags & (SHB_FLAG_FORCE_DIRECT | SHB_FLAG_NO_LONG_PATH));
        // crossovers measured against the length-sorted direct kernels (tools/ab_direct.py, flags 0 vs 0x200):
        // BLAKE3 chunk-parallel from ~10 kB mean (2-8 kB: direct x1.28; 8-24 kB: long x1.09; 10-50 kB: long x1.7),
        // K12 leaf-parallel from ~48 kB (10-50 kB: direct x1.22; 50-100 kB: long x1.13; 100-200 kB: long x1.28)
        const bool b3_long = long_ok && mean >= 10240;
        const bool k12_long = long_ok && mean >= 49152;
        const size_t xxh3_warp_min = 12288;   // below: thread per record, stripes through 128-bit stream loads (x1.2-1.8)
        const size_t nt_warp_min = 2560;   // below: thread per record with the shared-memory seed table (x1.4-1.9)
        // Thread-per-record kernels can walk the records in descending length order (counting sort of the indices by
        // 64-byte blocks) so that a warp's 32 records finish together.  Whether that pays is decided on the device
        // (one order per batch: the lowest threshold among the requested algorithms; below it the permutation is the
        // identity) from the sort's own histogram: w = E[longest of 32 records] / E[record] is what input order wastes, the
        // scattered reads of sorted order cost more the longer the records and the lighter the hash per byte
        // (tools/ab_direct.py, sorted vs input order on one B200):
        //   sha3 / k12 / blake3 (direct)   sorted whenever lengths differ at all: x1.15 (w 1.19) ... x2.2 (w 1.9)
        //   sha1                           same below 16 kB mean (x1.17 ... x2.0; 10-50 kB reads: x0.9-1.0)
        //   md5                            w >= 1.15 and mean < 16 kB (x1.2 ... x1.9; 10-50 kB reads: x0.6-0.7)
        //   xxhash / cityhash / murmur3 / rapidhash(32)   w >= 1.4 and mean < 2.5 kB (x1.1 ... x1.6 on 30-3000 bp;
        //                                  x0.6-0.8 at 2-8 kB, x0.3-0.5 at 10-50 kB)
        const uint32_t kHeavy = (1u << SHB_SHA3) | (1u << SHB_BLAKE3) | (1u << SHB_K12);
        const uint32_t kStream = (1u << SHB_XXHASH) | (1u << SHB_CITYHASH) | (1u << SHB_MURMUR3) | (1u << SHB_RAPIDHASH) |
                                 (1u << SHB_RAPIDHASH32) | (1u << SHB_NTHASH) | (1u << SHB_XXH3);
        const uint32_t kSortMask = kHeavy | (mean < 16384 ? (1u << SHB_SHA1) | (1u << SHB_MD5) : 0u) | (mean < 2560 ? kStream : 0u);
        auto min_waste_q8 = [&](int a) -> uint32_t {   // x256
            if (kStream >> a & 1u) return 358u;        // 1.40
            if (a == SHB_MD5) return 294u;             // 1.15
            return 269u;                               // 1.05
        };
        const uint32_t *perm = nullptr;
        if (d_scratch && n >= 1024 && n < ((size_t)1 << 32) && (want & kSortMask) && !(flags & SHB_FLAG_NO_LENGTH_SORT)) {   // 32-bit indices
            uint32_t lowest = 0xffffffffu;
            for (int a = 0; a < SHB_N_ALGOS; a++)
                if ((want & kSortMask) >> a & 1u) lowest = std::min(lowest, min_waste_q8(a));
            void *pbuf = static_cast<uint8_t *>(d_scratch) + ScratchLayout(total_bytes, n).perm_off;
            CU(shb::launch_length_sort(offs, n, pbuf, lowest, st));
            perm = static_cast<const uint32_t *>(pbuf);
        }
        // blake3 and nthash together on long records: ntHash is folded into the BLAKE3 chunk pass (one read of the residues)
        const bool b3_nt_fused = b3_long && d_scratch && (want >> SHB_BLAKE3 & 1u) && (want >> SHB_NTHASH & 1u) && !(flags & SHB_FLAG_NO_FUSED_B3NT);
        for (int a = 0; a < SHB_N_ALGOS; a++) {
            if (!(want >> a & 1u)) continue;
            if (b3_nt_fused && a == SHB_NTHASH) continue;
            if (a == SHB_RAPIDHASH32 && (want >> SHB_RAPIDHASH & 1u)) {   // a fold of rapidhash, which ran just before (a = 10)
                CU(shb::launch_rapid_fold32(slot_of[SHB_RAPIDHASH], n, slot_of[a], st));
                continue;
            }
            if (long_ok && a == SHB_NTHASH && mean >= nt_warp_min) {
                CU(shb::launch_nthash_long(fused_upper, res, offs, n, slot_of[a], st));
            } else if (long_ok && a == SHB_XXH3 && mean >= xxh3_warp_min) {
                CU(shb::launch_xxh3_long(fused_upper, res, offs, n, slot_of[a], st));
            } else if (k12_long && a == SHB_K12 && d_scratch) {
                ScratchLayout L(total_bytes, n);
                CU(shb::launch_k12_long(fused_upper, res, offs, n, total_bytes, slot_of[a],
                                        static_cast<uint8_t *>(d_scratch) + L.b3_off, st));
            } else if (b3_long && a == SHB_BLAKE3 && d_scratch) {
                ScratchLayout L(total_bytes, n);
                CU(shb::launch_blake3_long(fused_upper, res, offs, n, total_bytes, slot_of[a],
                                           static_cast<uint8_t *>(d_scratch) + L.b3_off, b3_nt_fused ? slot_of[SHB_NTHASH] : nullptr, st));
            } else {
                const uint32_t *order = (kSortMask >> a & 1u) ? perm : nullptr;
                // Long records, where the global sort's scattered reads cost more than they save: every CTA sorts its own
                // 128 consecutive records instead (kernels.cu: hash_direct_local_kernel).  Measured on 10-50 kB reads:
                // sha1 x1.24, md5 x1.10, murmur3 x1.13 (x1.06-1.10 from 2 kB); xxhash, rapidhash, cityhash and xxh3 lose
                // (x0.94 ... x0.6) and keep input order.
                const bool local_ok = n >= 1024 && !(flags & SHB_FLAG_NO_LENGTH_SORT);
                if (!order && local_ok &&
                    (((a == SHB_SHA1 || a == SHB_MD5) && mean >= 16384) || (a == SHB_MURMUR3 && mean >= 4096)))
                    order = shb::DIRECT_LOCAL_ORDER;
                CU(shb::launch_hash_direct(a, fused_upper, res, offs, n, slot_of[a], order, st));
            }
        }
    }
    // an algorithm requested twice ("--hash sha1,sha1"): computed once, the slot is copied
    for (int k = 0; k < n_algos; k++)
        if (slot_of[algos[k]] != d_digests[k])
            CU(cudaMemcpyAsync(d_digests[k], slot_of[algos[k]], n * shb::digest_size(algos[k]), cudaMemcpyDeviceToDevice, st));
    return SHB_OK;
}

bool offsets_ok_range(const int64_t *o, size_t lo, size_t hi)   // checks the steps o[i] - o[i-1], lo < i <= hi
{
    uint64_t bad = 0;
    for (size_t i = lo + 1; i <= hi; i++) bad |= (uint64_t)(o[i] - o[i - 1]) > 0x7FFFFFFFULL;   // negative steps wrap to huge values
    return bad == 0;
}
bool offsets_ok(const int64_t *o, size_t n)
{
    if (n < ((size_t)1 << 18)) return offsets_ok_range(o, 0, n);
    unsigned t = std::thread::hardware_concurrency();
    t = t < 2 ? 2 : t > 8 ? 8 : t;
    std::vector<std::thread> th;
    std::vector<char> ok(t, 1);
    for (unsigned k = 0; k < t; k++)
        th.emplace_back([&, k] { ok[k] = offsets_ok_range(o, n * k / t, n * (k + 1) / t); });
    bool all = true;
    for (unsigned k = 0; k < t; k++) { th[k].join(); all = all && ok[k]; }
    return all;
}

int check_algos(const int *algos, int n_algos)
{
    if (n_algos < 0 || (n_algos > 0 && !algos)) return fail(SHB_ERR_BAD_ARG, "bad algo list");
    for (int k = 0; k < n_algos; k++)
        if (algos[k] < 0 || algos[k] >= SHB_N_ALGOS) return fail(SHB_ERR_BAD_ARG, "algo id out of range");
    return SHB_OK;
}

}  // namespace

struct shb_batch {
    int device = 0;
    size_t max_bytes = 0, max_records = 0;
    cudaStream_t stream = nullptr;
    uint8_t *h_res = nullptr, *d_res = nullptr;
    int64_t *h_offs = nullptr, *d_offs = nullptr;
    // outputs, grown on demand
    uint8_t *h_dig = nullptr, *d_dig = nullptr;
    size_t dig_cap = 0;
    int64_t *h_len = nullptr, *d_len = nullptr;   // max_records
    void *d_scratch = nullptr;
    size_t scratch_cap = 0;
    uint8_t *h_out_res = nullptr;
    int64_t *h_out_offs = nullptr;
    // state of the last submit
    size_t n = 0;
    unsigned flags = 0;
    std::vector<int> algos;
    std::vector<size_t> slot_off;
    bool lengths_on_device = false, lengths_valid = false;
    unsigned long long *d_nonascii = nullptr, *h_nonascii = nullptr;   // SHB_FLAG_CHECK_ASCII
};

extern "C" {

int shb_init(int n_gpus, int flags)
{
    (void)flags;
    std::lock_guard<std::mutex> lk(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(SHB_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0"));
    int want = n_gpus <= 0 ? count : n_gpus;
    if (want > count) return fail(SHB_ERR_NO_DEVICE, "fewer CUDA devices than requested");
    for (int d = 0; d < want; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) != cudaSuccess || p.major != 10)
            return fail(SHB_ERR_NO_DEVICE, "device is not sm_100 (Blackwell B200); this library has no other code path");
    }
    g_ndev = want;
    return want;
}

void shb_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_ndev = 0;
}

const char *shb_last_error(void) { return t_err.c_str(); }
const char *shb_version(void) { return "seqhash_b200 0.1.0 (sm_100a)"; }

int shb_algo_id(const char *name)
{
    if (!name) return -1;
    for (int i = 0; i < SHB_N_ALGOS; i++)
        if (std::strcmp(name, kNames[i]) == 0) return i;
    return -1;
}
const char *shb_algo_name(int algo) { return (algo >= 0 && algo < SHB_N_ALGOS) ? kNames[algo] : nullptr; }
size_t shb_digest_size(int algo) { return (algo >= 0 && algo < SHB_N_ALGOS) ? (size_t)shb::digest_size(algo) : 0; }

unsigned long long shb_launch_count(void) { return shb::g_launches.load(); }
long long shb_debug_oob(void) { return shb::tile_debug_bounds_build() ? (long long)shb::tile_debug_oob() : -1; }

size_t shb_scratch_bytes(size_t total_bytes, size_t n_records) { return ScratchLayout(total_bytes, n_records).total; }

shb_batch *shb_batch_create_on(int device, size_t max_bytes, size_t max_records)
{
    if (g_ndev == 0 && shb_init(0, 0) < 0) return nullptr;
    if (device < 0 || device >= g_ndev) {
        fail(SHB_ERR_BAD_ARG, "device not selected by shb_init");
        return nullptr;
    }
    if (max_bytes >= ((size_t)1 << 40) || max_records >= ((size_t)1 << 32)) {
        fail(SHB_ERR_BAD_ARG, "batch capacity too large");
        return nullptr;
    }
    shb_batch *b = new shb_batch();
    b->device = device;
    b->max_bytes = max_bytes;
    b->max_records = max_records;
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaHostAlloc(&b->h_res, max_bytes + 32, cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc(&b->h_offs, (max_records + 1) * sizeof(int64_t), cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc(&b->h_len, (max_records + 1) * sizeof(int64_t), cudaHostAllocDefault) == cudaSuccess &&
              cudaMalloc(&b->d_res, max_bytes + 32) == cudaSuccess &&
              cudaMalloc(&b->d_offs, (max_records + 1) * sizeof(int64_t)) == cudaSuccess &&
              cudaMalloc(&b->d_len, (max_records + 1) * sizeof(int64_t)) == cudaSuccess &&
              cudaMalloc(&b->d_nonascii, 8) == cudaSuccess && cudaHostAlloc(&b->h_nonascii, 8, cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        fail(SHB_ERR_CUDA, std::string("batch allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        shb_batch_destroy(b);
        return nullptr;
    }
    b->h_offs[0] = 0;
    *b->h_nonascii = 0;
    return b;
}

shb_batch *shb_batch_create(size_t max_bytes, size_t max_records) { return shb_batch_create_on(0, max_bytes, max_records); }

void shb_batch_destroy(shb_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) { cudaStreamSynchronize(b->stream); cudaStreamDestroy(b->stream); }
    cudaFreeHost(b->h_res); cudaFreeHost(b->h_offs); cudaFreeHost(b->h_len); cudaFreeHost(b->h_dig);
    cudaFreeHost(b->h_out_res); cudaFreeHost(b->h_out_offs); cudaFreeHost(b->h_nonascii); cudaFree(b->d_nonascii);
    cudaFree(b->d_res); cudaFree(b->d_offs); cudaFree(b->d_len); cudaFree(b->d_dig); cudaFree(b->d_scratch);
    delete b;
}

uint8_t *shb_batch_residues(shb_batch *b) { return b ? b->h_res : nullptr; }
int64_t *shb_batch_offsets(shb_batch *b) { return b ? b->h_offs : nullptr; }

int shb_submit(shb_batch *b, size_t n, const int *algos, int n_algos, unsigned flags)
{
    if (!b) return fail(SHB_ERR_BAD_ARG, "null batch");
    if (int rc = check_algos(algos, n_algos)) return rc;
    if (n > b->max_records) return fail(SHB_ERR_CAPACITY, "more records than the batch capacity");
    if (b->h_offs[0] != 0) return fail(SHB_ERR_OFFSETS, "offsets[0] must be 0");
    // offsets must be monotone and each record shorter than 2^31 bytes: one pass over n+1 words, branch-free so that it
    // vectorises, and split over a few threads for large batches (it sits on the submitting thread, inside every end-to-end
    // measurement: 8 B per record at one word per ns is 20 ms for the 20 M records of a C3 step)
    if (!offsets_ok(b->h_offs, n)) return fail(SHB_ERR_OFFSETS, "offsets not monotone or record >= 2^31 bytes");
    const size_t total = (size_t)b->h_offs[n];
    if (total > b->max_bytes) return fail(SHB_ERR_CAPACITY, "more residue bytes than the batch capacity");
    CU(cudaSetDevice(b->device));
    // outputs
    b->algos.assign(algos, algos + n_algos);
    b->slot_off.assign(n_algos + 1, 0);
    for (int k = 0; k < n_algos; k++) b->slot_off[k + 1] = b->slot_off[k] + align_up(n * shb::digest_size(algos[k]), 256);
    const size_t need = b->slot_off[n_algos];
    if (need > b->dig_cap) {
        CU(cudaStreamSynchronize(b->stream));
        cudaFreeHost(b->h_dig); cudaFree(b->d_dig);
        b->h_dig = nullptr; b->d_dig = nullptr; b->dig_cap = 0;
        CU(cudaHostAlloc(&b->h_dig, need, cudaHostAllocDefault));
        CU(cudaMalloc(&b->d_dig, need));
        b->dig_cap = need;
    }
    const bool norm = flags & (SHB_FLAG_STRIP | SHB_FLAG_EMIT_SEQ);
    bool wants_b3_long = false;
    for (int k = 0; k < n_algos; k++) wants_b3_long |= ((algos[k] == SHB_BLAKE3 || algos[k] == SHB_K12) && n && total / n >= 4096);
    const bool direct_path = n && (total / n) * 128 * 17 / 16 + 2047 > 100 * 1024;   // mirrors run_pipeline's tile_mode
    if (norm || wants_b3_long || direct_path) {
        const size_t sneed = ScratchLayout(b->max_bytes, b->max_records).total;
        if (sneed > b->scratch_cap) {
            CU(cudaStreamSynchronize(b->stream));
            cudaFree(b->d_scratch);
            b->d_scratch = nullptr; b->scratch_cap = 0;
            CU(cudaMalloc(&b->d_scratch, sneed));
            b->scratch_cap = sneed;
        }
        if ((flags & SHB_FLAG_EMIT_SEQ) && !b->h_out_res) {
            CU(cudaHostAlloc(&b->h_out_res, b->max_bytes + 32, cudaHostAllocDefault));
            CU(cudaHostAlloc(&b->h_out_offs, (b->max_records + 1) * sizeof(int64_t), cudaHostAllocDefault));
        }
    }
    b->n = n;
    b->flags = flags;
    b->lengths_on_device = norm;
    b->lengths_valid = false;
    // H2D (the 16 pad bytes past `total` travel too so device readers never see stale bytes)
    std::memset(b->h_res + total, 0, 32);
    CU(cudaMemcpyAsync(b->d_res, b->h_res, align_up(total + 16, 16), cudaMemcpyHostToDevice, b->stream));
    CU(cudaMemcpyAsync(b->d_offs, b->h_offs, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, b->stream));
    *b->h_nonascii = 0;
    if (flags & SHB_FLAG_CHECK_ASCII) {
        CU(shb::launch_count_nonascii(b->d_res, total, b->d_nonascii, b->stream));
        CU(cudaMemcpyAsync(b->h_nonascii, b->d_nonascii, 8, cudaMemcpyDeviceToHost, b->stream));
    }
    std::vector<uint8_t *> d_slots(n_algos);
    for (int k = 0; k < n_algos; k++) d_slots[k] = b->d_dig + b->slot_off[k];
    const uint8_t *norm_res = nullptr;
    const int64_t *norm_offs = nullptr;
    if (int rc = run_pipeline(b->d_res, b->d_offs, n, total, algos, n_algos, flags, d_slots.data(),
                              norm ? b->d_len : nullptr, b->d_scratch, b->stream, &norm_res, &norm_offs))
        return rc;
    // D2H
    if (need) CU(cudaMemcpyAsync(b->h_dig, b->d_dig, need, cudaMemcpyDeviceToHost, b->stream));
    if (norm) CU(cudaMemcpyAsync(b->h_len, b->d_len, n * sizeof(int64_t), cudaMemcpyDeviceToHost, b->stream));
    if (flags & SHB_FLAG_EMIT_SEQ) {
        CU(cudaMemcpyAsync(b->h_out_offs, norm_offs, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, b->stream));
        // normalised residues are never longer than the input
        CU(cudaMemcpyAsync(b->h_out_res, norm_res, align_up(total, 16), cudaMemcpyDeviceToHost, b->stream));
    }
    return SHB_OK;
}

int shb_wait(shb_batch *b)
{
    if (!b) return fail(SHB_ERR_BAD_ARG, "null batch");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    return SHB_OK;
}

const uint8_t *shb_digests(shb_batch *b, int slot)
{
    if (!b || slot < 0 || slot >= (int)b->algos.size()) return nullptr;
    return b->h_dig + b->slot_off[slot];
}

const int64_t *shb_out_lengths(shb_batch *b)
{
    if (!b) return nullptr;
    if (!b->lengths_on_device && !b->lengths_valid) {   // nothing was stripped: lengths are the input's
        for (size_t i = 0; i < b->n; i++) b->h_len[i] = b->h_offs[i + 1] - b->h_offs[i];
        b->lengths_valid = true;
    }
    return b->h_len;
}

size_t shb_batch_nonascii(shb_batch *b) { return b ? (size_t)*b->h_nonascii : 0; }

const int64_t *shb_out_offsets(shb_batch *b) { return (b && (b->flags & SHB_FLAG_EMIT_SEQ)) ? b->h_out_offs : nullptr; }
const uint8_t *shb_out_residues(shb_batch *b) { return (b && (b->flags & SHB_FLAG_EMIT_SEQ)) ? b->h_out_res : nullptr; }

int shb_hash_device(int device, const uint8_t *d_residues, const int64_t *d_offsets, size_t n_records,
                    size_t total_bytes, const int *algos, int n_algos, unsigned flags, uint8_t *const *d_digests,
                    int64_t *d_out_lengths, void *d_scratch, void *stream)
{
    if (g_ndev == 0 && shb_init(0, 0) < 0) return SHB_ERR_NO_DEVICE;
    if (device < 0 || device >= g_ndev) return fail(SHB_ERR_BAD_ARG, "device not selected by shb_init");
    if (int rc = check_algos(algos, n_algos)) return rc;
    if ((reinterpret_cast<uintptr_t>(d_residues) & 15u) != 0) return fail(SHB_ERR_BAD_ARG, "d_residues must be 16-byte aligned");
    if (n_records >= ((size_t)1 << 32)) return fail(SHB_ERR_BAD_ARG, "too many records");
    CU(cudaSetDevice(device));
    return run_pipeline(d_residues, d_offsets, n_records, total_bytes, algos, n_algos, flags, d_digests,
                        d_out_lengths, d_scratch, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

int shb_synth_fill(int device, uint8_t *d_residues, size_t total_bytes, uint64_t seed, unsigned lower_permille,
                   unsigned n_permille, void *stream)
{
    if (g_ndev == 0 && shb_init(0, 0) < 0) return SHB_ERR_NO_DEVICE;
    CU(cudaSetDevice(device));
    CU(shb::launch_synth_fill(d_residues, total_bytes, seed, lower_permille, n_permille, static_cast<cudaStream_t>(stream)));
    return SHB_OK;
}

static uint32_t g_probe_checksum = 0;
unsigned shb_probe_last_checksum(void) { return g_probe_checksum; }

double shb_probe_int_peak(int device, int kind, int iters)
{
    if (g_ndev == 0 && shb_init(0, 0) < 0) return -1.0;
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    uint32_t *sink = nullptr;
    cudaEvent_t e0, e1;
    if (cudaMalloc(&sink, 64) != cudaSuccess) return -1.0;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double lane_instr = 0;
    double best = -1.0;
    for (int rep = 0; rep < 4; rep++) {   // rep 0 warms up
        cudaMemsetAsync(sink, 0, 64, 0);
        cudaEventRecord(e0, 0);
        if (shb::launch_probe(kind, iters, sink, &lane_instr, 0) != cudaSuccess) { best = -1.0; break; }
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double rate = lane_instr / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    uint32_t host_sink[2] = {0, 0};
    cudaMemcpy(host_sink, sink, sizeof host_sink, cudaMemcpyDeviceToHost);
    g_probe_checksum = host_sink[1];
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

// Pinned host -> device copy rate of this process: cudaHostAlloc (optionally write-combined), first touch on the calling
// thread, plain cudaMemcpyAsync in `piece`-byte copies on one stream for `seconds`.  The roof of every end-to-end number;
// run from one process per GPU at once for the aggregate of the box (tools/h2d_roof.py).  Returns GB/s, < 0 on failure.
double shb_probe_h2d(int device, size_t bytes, size_t piece, double seconds, int write_combined)
{
    if (g_ndev == 0 && shb_init(0, 0) < 0) return -1.0;
    if (cudaSetDevice(device) != cudaSuccess || bytes == 0 || piece == 0) return -1.0;
    uint8_t *h = nullptr, *d = nullptr;
    cudaStream_t st = nullptr;
    double gbps = -1.0;
    if (cudaHostAlloc(&h, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault) == cudaSuccess &&
        cudaMalloc(&d, bytes) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess) {
        std::memset(h, 1, bytes);
        auto pass = [&]() {
            for (size_t off = 0; off < bytes; off += piece)
                cudaMemcpyAsync(d + off, h + off, std::min(piece, bytes - off), cudaMemcpyHostToDevice, st);
            return cudaStreamSynchronize(st) == cudaSuccess;
        };
        bool ok = pass();   // warm-up
        const auto t0 = std::chrono::steady_clock::now();
        double dt = 0;
        size_t moved = 0;
        while (ok && dt < seconds) {
            ok = pass();
            moved += bytes;
            dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        if (ok && dt > 0) gbps = (double)moved / dt / 1e9;
    }
    if (st) cudaStreamDestroy(st);
    cudaFree(d);
    cudaFreeHost(h);
    return gbps;
}

// ---- whole-record-loop path: raw FASTA/FASTQ text in, formatted output text out ----

struct shb_textproc {
    int device = 0;
    size_t cap = 0;
    long long max_records = 0;
    cudaStream_t stream = nullptr;
    uint8_t *h_text = nullptr, *d_text = nullptr, *d_res = nullptr;
    void *d_parse_scratch = nullptr;
    shb::FxOut rec{};
    shb::FxInfo *d_info = nullptr, *h_info = nullptr;
    uint32_t *d_out_len = nullptr;
    long long *d_out_off = nullptr;
    uint8_t *d_label = nullptr;
    void *d_fmt_scan = nullptr;
    uint8_t *d_dig = nullptr;
    size_t dig_cap = 0;
    void *d_scratch = nullptr;
    size_t scratch_cap = 0;
    uint8_t *d_out = nullptr, *h_out = nullptr;
    size_t out_cap = 0;
    // BGZF input (allocated on first use): pinned + device staging of the compressed members, their table, the inflated text
    uint8_t *h_comp = nullptr, *d_comp = nullptr, *d_stage = nullptr;
    size_t comp_cap = 0, tab_cap = 0;
    unsigned long long *h_tab = nullptr, *d_tab = nullptr;
    long long *d_bounds = nullptr, *h_bounds = nullptr;   // 4 results of fx_bounds_kernel + the inflate status
    size_t dev_text_bytes = 0;                            // bytes of d_text a device-resident run worked on (shb_textproc_resume)
    cudaEvent_t ev_inf[2] = {};                           // around the inflate kernel
    float inflate_ms = 0.f;
    // results of the last run
    size_t out_bytes = 0, consumed = 0, n_records = 0, n_empty = 0, n_nonascii = 0;
    cudaEvent_t ev[6] = {};     // start, after H2D, after split, after hash, after format, after D2H
    float ms[5] = {};
};

shb_textproc *shb_textproc_create(int device, size_t max_text_bytes)
{
    if (g_ndev == 0 && shb_init(0, 0) < 0) return nullptr;
    if (device < 0 || device >= g_ndev || max_text_bytes == 0 || max_text_bytes >= ((size_t)1 << 40)) {
        fail(SHB_ERR_BAD_ARG, "bad device or capacity");
        return nullptr;
    }
    shb_textproc *t = new shb_textproc();
    t->device = device;
    t->cap = max_text_bytes;
    t->max_records = (long long)(max_text_bytes / 16 + 1024);
    const size_t nr = (size_t)t->max_records;
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaHostAlloc(&t->h_info, sizeof(shb::FxInfo) + 64, cudaHostAllocDefault) == cudaSuccess &&
              cudaMalloc(&t->d_text, max_text_bytes + 64) == cudaSuccess &&
              cudaMalloc(&t->d_res, max_text_bytes + 64) == cudaSuccess &&
              cudaMalloc(&t->d_parse_scratch, shb::fx_scratch_bytes(max_text_bytes)) == cudaSuccess &&
              cudaMalloc(&t->d_info, sizeof(shb::FxInfo)) == cudaSuccess &&
              cudaMalloc(&t->rec.offsets, (nr + 1) * 8) == cudaSuccess && cudaMalloc(&t->rec.name_start, nr * 8) == cudaSuccess &&
              cudaMalloc(&t->rec.name_len, nr * 4) == cudaSuccess && cudaMalloc(&t->rec.qual_start, nr * 8) == cudaSuccess &&
              cudaMalloc(&t->rec.qual_len, nr * 4) == cudaSuccess && cudaMalloc(&t->rec.rec_start, (nr + 1) * 8) == cudaSuccess &&
              cudaMalloc(&t->d_out_len, (nr + 1) * 4) == cudaSuccess && cudaMalloc(&t->d_out_off, (nr + 1) * 8) == cudaSuccess &&
              cudaMalloc(&t->d_label, 4096) == cudaSuccess &&
              cudaMalloc(&t->d_fmt_scan, shb::scan_tmp_bytes(nr + 1)) == cudaSuccess;
    for (int k = 0; ok && k < 6; k++) ok = cudaEventCreate(&t->ev[k]) == cudaSuccess;
    if (!ok) {
        fail(SHB_ERR_CUDA, std::string("textproc allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        shb_textproc_destroy(t);
        return nullptr;
    }
    t->rec.residues = t->d_res;
    t->rec.max_records = t->max_records;
    return t;
}

void shb_textproc_destroy(shb_textproc *t)
{
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->stream) { cudaStreamSynchronize(t->stream); cudaStreamDestroy(t->stream); }
    for (int k = 0; k < 6; k++) if (t->ev[k]) cudaEventDestroy(t->ev[k]);
    for (int k = 0; k < 2; k++) if (t->ev_inf[k]) cudaEventDestroy(t->ev_inf[k]);
    cudaFreeHost(t->h_text); cudaFreeHost(t->h_info); cudaFreeHost(t->h_out);
    cudaFree(t->d_text); cudaFree(t->d_res); cudaFree(t->d_parse_scratch); cudaFree(t->d_info);
    cudaFree(t->rec.offsets); cudaFree(t->rec.name_start); cudaFree(t->rec.name_len); cudaFree(t->rec.qual_start);
    cudaFree(t->rec.qual_len); cudaFree(t->rec.rec_start); cudaFree(t->d_out_len); cudaFree(t->d_out_off);
    cudaFree(t->d_label); cudaFree(t->d_fmt_scan); cudaFree(t->d_dig); cudaFree(t->d_scratch); cudaFree(t->d_out);
    cudaFreeHost(t->h_comp); cudaFreeHost(t->h_tab); cudaFreeHost(t->h_bounds);
    cudaFree(t->d_comp); cudaFree(t->d_stage); cudaFree(t->d_tab); cudaFree(t->d_bounds);
    delete t;
}

// the pinned text buffer exists from the first call on: a textproc that only ever sees BGZF members never pins it
uint8_t *shb_textproc_input(shb_textproc *t)
{
    if (!t) return nullptr;
    if (!t->h_text) {
        if (cudaSetDevice(t->device) != cudaSuccess || cudaHostAlloc(&t->h_text, t->cap + 64, cudaHostAllocDefault) != cudaSuccess) {
            t->h_text = nullptr;
            fail(SHB_ERR_CUDA, std::string("textproc staging allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
            return nullptr;
        }
    }
    return t->h_text;
}
size_t shb_textproc_capacity(shb_textproc *t) { return t ? t->cap : 0; }

// The record loop on a text that is already in t->d_text (16 zero bytes behind it): split, hash, format, copy the output back.
// ev[0] / ev[1] have been recorded by the caller around whatever brought the text there.
static int textproc_process(shb_textproc *t, size_t n_bytes, int format, int is_final, const int *algos, int n_algos,
                            unsigned flags, int headers_only, const char *label, size_t label_len)
{
    cudaStream_t st = t->stream;
    if (label_len) CU(cudaMemcpyAsync(t->d_label, label, label_len, cudaMemcpyHostToDevice, st));
    // 1. split into records (CSR residues + name / quality spans)
    CU(shb::launch_fx_parse(t->d_text, n_bytes, format == SHB_FORMAT_FASTQ, is_final != 0, t->rec, t->d_info, t->d_parse_scratch, st));
    CU(cudaMemcpyAsync(t->h_info, t->d_info, sizeof(shb::FxInfo), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(t->ev[2], st));
    CU(cudaStreamSynchronize(st));
    // FX_BAD_MARKER: the records before the malformed one are still hashed and formatted (the reference's reader yields
    // them before it fails); FX_MORE: the record index was full, the caller resumes at `consumed`.
    const bool malformed = t->h_info->status == shb::FX_BAD_MARKER;
    const int done_rc = malformed ? SHB_PARTIAL_FORMAT : SHB_OK;
    if (malformed) t_err = "fastx: invalid FASTA/Q format";
    const size_t n = (size_t)t->h_info->n_records, total = (size_t)t->h_info->n_residues;
    t->n_records = n;
    t->consumed = (size_t)t->h_info->consumed;
    t->n_nonascii = (size_t)t->h_info->n_nonascii;
    if (n == 0) return done_rc;
    // 2. hash (the residues are already whitespace-free; upper-casing is fused into the hash kernels)
    std::vector<size_t> slot_off(n_algos + 1, 0);
    for (int k = 0; k < n_algos; k++) slot_off[k + 1] = slot_off[k] + align_up(n * shb::digest_size(algos[k]), 256);
    if (slot_off[n_algos] > t->dig_cap) {
        cudaFree(t->d_dig);
        t->d_dig = nullptr; t->dig_cap = 0;
        CU(cudaMalloc(&t->d_dig, slot_off[n_algos]));
        t->dig_cap = slot_off[n_algos];
    }
    const size_t sneed = ScratchLayout(t->cap, (size_t)t->max_records).total;
    if (sneed > t->scratch_cap) {
        cudaFree(t->d_scratch);
        t->d_scratch = nullptr; t->scratch_cap = 0;
        CU(cudaMalloc(&t->d_scratch, sneed));
        t->scratch_cap = sneed;
    }
    std::vector<uint8_t *> d_slots(n_algos);
    for (int k = 0; k < n_algos; k++) d_slots[k] = t->d_dig + slot_off[k];
    if (int rc = run_pipeline(t->d_res, reinterpret_cast<const int64_t *>(t->rec.offsets), n, total, algos, n_algos,
                              flags & SHB_FLAG_UPPER, d_slots.data(), nullptr, t->d_scratch, st, nullptr, nullptr))
        return rc;
    CU(cudaEventRecord(t->ev[3], st));
    // 3. format: `file;h1;..;hn;Name` (+ record body), seqhasher.go:261-282
    shb::FmtCfg cfg{};
    cfg.label = t->d_label;
    cfg.label_len = label ? (int)label_len : -1;
    cfg.n_algos = n_algos;
    for (int k = 0; k < n_algos; k++) { cfg.dsize[k] = shb::digest_size(algos[k]); cfg.digests[k] = d_slots[k]; }
    cfg.headers_only = headers_only;
    cfg.fastq = format == SHB_FORMAT_FASTQ;
    cfg.upper = (flags & SHB_FLAG_UPPER) ? 1 : 0;
    CU(shb::launch_fx_format(cfg, t->rec, n, t->d_text, t->d_out_len, t->d_out_off, nullptr, t->d_info, true, t->d_fmt_scan, st));
    CU(cudaMemcpyAsync(&t->h_info->reserved, t->d_out_off + n, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&t->h_info->n_empty, &t->d_info->n_empty, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const size_t out_bytes = (size_t)t->h_info->reserved;
    t->n_empty = (size_t)t->h_info->n_empty;
    if (out_bytes > t->out_cap) {
        cudaFree(t->d_out); cudaFreeHost(t->h_out);
        t->d_out = nullptr; t->h_out = nullptr; t->out_cap = 0;
        const size_t want = out_bytes + out_bytes / 4 + 4096;
        CU(cudaMalloc(&t->d_out, want));
        CU(cudaHostAlloc(&t->h_out, want, cudaHostAllocDefault));
        t->out_cap = want;
    }
    CU(shb::launch_fx_format(cfg, t->rec, n, t->d_text, t->d_out_len, t->d_out_off, t->d_out, t->d_info, false, t->d_fmt_scan, st));
    CU(cudaEventRecord(t->ev[4], st));
    CU(cudaMemcpyAsync(t->h_out, t->d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(t->ev[5], st));
    CU(cudaStreamSynchronize(st));
    for (int k = 0; k < 5; k++) cudaEventElapsedTime(&t->ms[k], t->ev[k], t->ev[k + 1]);
    t->out_bytes = out_bytes;
    return done_rc;
}

static int textproc_check_args(shb_textproc *t, int format, const int *algos, int n_algos, const char *label, size_t *label_len)
{
    if (!t) return fail(SHB_ERR_BAD_ARG, "null textproc");
    if (int rc = check_algos(algos, n_algos)) return rc;
    if (n_algos > shb::FMT_MAX_FIELDS) return fail(SHB_ERR_BAD_ARG, "too many hash types (at most 48 fields per record)");
    if (format != SHB_FORMAT_FASTA && format != SHB_FORMAT_FASTQ) return fail(SHB_ERR_BAD_ARG, "format must be FASTA or FASTQ");
    *label_len = label ? std::strlen(label) : 0;
    if (*label_len > 4096) return fail(SHB_ERR_BAD_ARG, "label too long");
    return SHB_OK;
}

int shb_textproc_run(shb_textproc *t, size_t n_bytes, int format, int is_final, const int *algos, int n_algos,
                     unsigned flags, int headers_only, const char *label)
{
    size_t label_len = 0;
    if (int rc = textproc_check_args(t, format, algos, n_algos, label, &label_len)) return rc;
    if (n_bytes > t->cap) return fail(SHB_ERR_CAPACITY, "text larger than the textproc capacity");
    CU(cudaSetDevice(t->device));
    t->out_bytes = t->consumed = t->n_records = t->n_empty = t->n_nonascii = 0;
    t->dev_text_bytes = 0;
    if (n_bytes == 0) return SHB_OK;
    if (!t->h_text) return fail(SHB_ERR_BAD_ARG, "no text staged: write the chunk into shb_textproc_input() first");
    if (is_final && t->h_text[n_bytes - 1] != '\n') t->h_text[n_bytes++] = '\n';   // the pinned buffer has 64 spare bytes
    std::memset(t->h_text + n_bytes, 0, 32);
    cudaStream_t st = t->stream;
    for (float &m : t->ms) m = 0.f;
    CU(cudaEventRecord(t->ev[0], st));
    CU(cudaMemcpyAsync(t->d_text, t->h_text, align_up(n_bytes + 16, 16), cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(t->ev[1], st));
    return textproc_process(t, n_bytes, format, is_final, algos, n_algos, flags, headers_only, label, label_len);
}

// ---- BGZF input: compressed members in, the same output out; the text exists on the device only ----

uint8_t *shb_textproc_bgzf_input(shb_textproc *t, size_t comp_capacity)
{
    if (!t) { fail(SHB_ERR_BAD_ARG, "null textproc"); return nullptr; }
    if (comp_capacity <= t->comp_cap) return t->h_comp;
    if (cudaSetDevice(t->device) != cudaSuccess) { fail(SHB_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    cudaStreamSynchronize(t->stream);
    uint8_t *nh = nullptr, *nd = nullptr;
    if (cudaHostAlloc(&nh, comp_capacity + 64, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&nd, comp_capacity + 64) != cudaSuccess) {
        cudaFreeHost(nh);
        fail(SHB_ERR_CUDA, std::string("BGZF staging allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    if (t->h_comp && t->comp_cap) std::memcpy(nh, t->h_comp, t->comp_cap);   // growing keeps what the caller has staged
    cudaFreeHost(t->h_comp); cudaFree(t->d_comp);
    t->h_comp = nh; t->d_comp = nd; t->comp_cap = comp_capacity;
    return t->h_comp;
}

int shb_textproc_run_bgzf(shb_textproc *t, size_t n_comp_bytes, const uint64_t *members, size_t n_own, size_t n_extra,
                          long long first_start, int at_eof, int format, const int *algos, int n_algos,
                          unsigned flags, int headers_only, const char *label)
{
    size_t label_len = 0;
    if (int rc = textproc_check_args(t, format, algos, n_algos, label, &label_len)) return rc;
    if (n_comp_bytes > t->comp_cap || !t->h_comp) return fail(SHB_ERR_CAPACITY, "more compressed bytes than shb_textproc_bgzf_input was asked for");
    const size_t nm = n_own + n_extra;
    if (nm && !members) return fail(SHB_ERR_BAD_ARG, "null member table");
    CU(cudaSetDevice(t->device));
    t->out_bytes = t->consumed = t->n_records = t->n_empty = t->n_nonascii = 0;
    t->dev_text_bytes = 0;
    if (nm == 0) return SHB_OK;
    if (!t->d_stage) {
        CU(cudaMalloc(&t->d_stage, t->cap + 64));
        CU(cudaMalloc(&t->d_bounds, 8 * sizeof(long long)));
        CU(cudaHostAlloc(&t->h_bounds, 8 * sizeof(long long), cudaHostAllocDefault));
        CU(cudaEventCreate(&t->ev_inf[0]));
        CU(cudaEventCreate(&t->ev_inf[1]));
    }
    if (nm > t->tab_cap) {
        cudaStreamSynchronize(t->stream);
        cudaFreeHost(t->h_tab); cudaFree(t->d_tab);
        t->h_tab = nullptr; t->d_tab = nullptr; t->tab_cap = 0;
        const size_t want = nm + nm / 2 + 1024;
        CU(cudaHostAlloc(&t->h_tab, want * 32, cudaHostAllocDefault));
        CU(cudaMalloc(&t->d_tab, want * 32));
        t->tab_cap = want;
    }
    // the library's table: where every member's text goes (prefix sum of ISIZE)
    size_t text_bytes = 0, own_bytes = 0;
    for (size_t i = 0; i < nm; i++) {
        const uint64_t off = members[3 * i], clen = members[3 * i + 1], packed = members[3 * i + 2];
        const uint32_t isize = (uint32_t)packed;
        if (off > n_comp_bytes || clen > n_comp_bytes - off) return fail(SHB_ERR_BAD_ARG, "BGZF member outside the staged bytes");
        if (isize > 65536u) return fail(SHB_ERR_BAD_ARG, "BGZF member announces more than 64 KiB of text");
        t->h_tab[4 * i] = off; t->h_tab[4 * i + 1] = clen; t->h_tab[4 * i + 2] = packed; t->h_tab[4 * i + 3] = text_bytes;
        text_bytes += isize;
        if (i + 1 == n_own) own_bytes = text_bytes;
    }
    if (text_bytes + 1 > t->cap) return fail(SHB_ERR_CAPACITY, "text larger than the textproc capacity");
    cudaStream_t st = t->stream;
    for (float &m : t->ms) m = 0.f;
    CU(cudaEventRecord(t->ev[0], st));
    CU(cudaMemcpyAsync(t->d_comp, t->h_comp, align_up(n_comp_bytes + 16, 16), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(t->d_tab, t->h_tab, nm * 32, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(t->ev_inf[0], st));
    CU(shb::launch_bgzf_inflate(t->d_comp, t->d_tab, nm, t->d_stage, reinterpret_cast<unsigned long long *>(t->d_bounds + 4), st));
    CU(cudaEventRecord(t->ev_inf[1], st));
    // Which records are this chunk's: the record starts p with 0 < p <= own_bytes, p counted from the chunk's first text byte
    // (the first chunk of the input also owns p = first_start).  The next chunk looks for its first record start from ITS
    // byte 1 on, which is this chunk's own_bytes + 1: both derive the shared boundary from the same bytes with the same rule,
    // so no record is lost or taken twice.
    const long long from0 = first_start >= 0 ? first_start : 1;
    const long long from1 = (long long)own_bytes + 1;
    CU(shb::launch_fx_bounds(t->d_stage, text_bytes, first_start >= 0 ? -1 : from0, from1, format == SHB_FORMAT_FASTQ, t->d_bounds, st));
    CU(cudaMemcpyAsync(t->h_bounds, t->d_bounds, 5 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&t->inflate_ms, t->ev_inf[0], t->ev_inf[1]);
    const unsigned long long inf_status = (unsigned long long)t->h_bounds[4];
    if (inf_status != ~0ull) {
        const unsigned code = (unsigned)(inf_status & 0xffu);
        return fail(SHB_ERR_INFLATE, code == 9u ? "gzip: invalid checksum" : "flate: corrupt input");
    }
    long long start = first_start >= 0 ? first_start : t->h_bounds[0];
    long long end = t->h_bounds[1];
    const long long trimmed = t->h_bounds[2];
    int is_final = 0;
    if (end < 0) {
        if (!at_eof) return SHB_NEED_MORE;          // the record that crosses the chunk's end is longer than the extra members
        end = trimmed;                               // the input ends inside (or right behind) this chunk's last record
        is_final = 1;
    }
    if (start < 0 || start >= end) {                 // no record starts in this chunk (one long record runs through it)
        t->consumed = 0;
        for (int k = 0; k < 5; k++) t->ms[k] = 0.f;
        return SHB_OK;
    }
    size_t n_bytes = (size_t)(end - start);
    CU(cudaMemcpyAsync(t->d_text, t->d_stage + start, n_bytes, cudaMemcpyDeviceToDevice, st));
    bool add_nl = false;
    if (is_final) {
        // a missing last line break is added (the reader's rule for the end of the input)
        if ((size_t)end < text_bytes) add_nl = true;                       // blanks were trimmed: whatever followed, the line ends here
        else add_nl = (uint8_t)t->h_bounds[3] != '\n';
    }
    if (add_nl) { CU(cudaMemsetAsync(t->d_text + n_bytes, '\n', 1, st)); n_bytes++; }
    CU(cudaMemsetAsync(t->d_text + n_bytes, 0, 32, st));
    CU(cudaEventRecord(t->ev[1], st));
    t->dev_text_bytes = n_bytes;
    // every chunk holds whole records only, so each is "final" for the splitter
    return textproc_process(t, n_bytes, format, 1, algos, n_algos, flags, headers_only, label, label_len);
}

int shb_textproc_resume(shb_textproc *t, int format, const int *algos, int n_algos, unsigned flags, int headers_only, const char *label)
{
    size_t label_len = 0;
    if (int rc = textproc_check_args(t, format, algos, n_algos, label, &label_len)) return rc;
    if (!t->d_stage || t->dev_text_bytes == 0 || t->consumed == 0 || t->consumed >= t->dev_text_bytes)
        return fail(SHB_ERR_BAD_ARG, "nothing to resume: the last run was not a partly consumed device-resident chunk");
    CU(cudaSetDevice(t->device));
    cudaStream_t st = t->stream;
    const size_t rest = t->dev_text_bytes - t->consumed;
    CU(cudaEventRecord(t->ev[0], st));
    CU(cudaMemcpyAsync(t->d_stage, t->d_text + t->consumed, rest, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(t->d_text, t->d_stage, rest, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(t->d_text + rest, 0, 32, st));
    CU(cudaEventRecord(t->ev[1], st));
    t->out_bytes = t->consumed = t->n_records = t->n_empty = t->n_nonascii = 0;
    t->dev_text_bytes = rest;
    return textproc_process(t, rest, format, 1, algos, n_algos, flags, headers_only, label, label_len);
}

float shb_textproc_last_inflate_ms(shb_textproc *t) { return t ? t->inflate_ms : 0.f; }

int shb_textproc_pending(shb_textproc *t)
{
    return t && t->dev_text_bytes && t->n_records && t->consumed && t->consumed < t->dev_text_bytes ? 1 : 0;
}

const uint8_t *shb_textproc_output(shb_textproc *t, size_t *out_bytes)
{
    if (!t) return nullptr;
    if (out_bytes) *out_bytes = t->out_bytes;
    return t->h_out;
}
size_t shb_textproc_consumed(shb_textproc *t) { return t ? t->consumed : 0; }
void shb_textproc_last_ms(shb_textproc *t, float *ms5)
{
    for (int k = 0; k < 5; k++) ms5[k] = t ? t->ms[k] : 0.f;
}
size_t shb_textproc_records(shb_textproc *t) { return t ? t->n_records : 0; }
size_t shb_textproc_empty_records(shb_textproc *t) { return t ? t->n_empty : 0; }
size_t shb_textproc_nonascii(shb_textproc *t) { return t ? t->n_nonascii : 0; }

}  // extern "C"
