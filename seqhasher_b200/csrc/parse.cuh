// parse.cuh — record splitter / formatter interface (internal to libseqhash_b200).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace shb {

constexpr uint32_t FX_BLK = 4096;
// FX_BAD_MARKER: the text at `consumed` is not a valid record (FASTQ '@' / '+' markers, a truncated last record);
// the n_records before it are complete and usable.  FX_MORE: the record index is full — the chunk was cut after
// max_records records and the caller runs the rest (from `consumed`) as another chunk.
enum : long long { FX_OK = 0, FX_BAD_MARKER = 1, FX_MORE = 2 };

struct FxInfo {            // device-side result block
    long long n_records;   // complete records
    long long n_residues;  // residue bytes of those records (= offsets[n_records])
    long long status;      // FX_*
    long long consumed;    // bytes of the text covered by the complete records
    long long n_empty;     // records whose sequence is empty (formatter)
    long long reserved;
    long long first_bad;   // index of the first malformed record (LLONG_MAX: none); set by the splitter
    long long n_nonascii;  // sequence bytes >= 0x80 in the chunk (the reference normalises those with Unicode rules, seqhasher.go:243-251)
};

struct FxOut {
    uint8_t *residues;       // compacted (whitespace-free) sequence bytes
    long long *offsets;      // max_records + 1
    long long *name_start;   // first byte of the record name in the text
    int *name_len;
    long long *qual_start;   // FASTQ: quality line
    int *qual_len;
    long long *rec_start;    // line start of the record's name line
    long long max_records;
};

constexpr int FMT_MAX_FIELDS = 48;   // hash fields per output record ("--hash" may repeat names, seqhasher.go:132-137)
struct FmtCfg {
    const uint8_t *label;    // device copy of the file label
    int label_len;           // -1: no file name field
    int n_algos;
    int dsize[FMT_MAX_FIELDS];
    const uint8_t *digests[FMT_MAX_FIELDS];
    int headers_only, fastq, upper;
};

// BGZF members inflated on the device (inflate.cuh): member i = deflate data comp[tab[4i], +tab[4i+1]) -> text[tab[4i+3], +isize),
// tab[4i+2] = isize | crc32 << 32; text is 16-byte aligned.  *status receives all ones, or ((1 + index) << 8 | INF_ERR_*) of the
// first member that failed to decode, did not produce isize bytes or whose CRC-32 differs.
cudaError_t launch_bgzf_inflate(const uint8_t *comp, const unsigned long long *tab, size_t n_members, uint8_t *text,
                                unsigned long long *status, cudaStream_t st);
// first record start at index >= from0 -> res[0], >= from1 -> res[1] (-1: none before n; a negative `from` is not searched);
// res[2] = one past the last non-blank byte, res[3] = the last byte.  See fx_bounds_kernel.
cudaError_t launch_fx_bounds(const uint8_t *text, size_t n, long long from0, long long from1, bool fastq, long long *res, cudaStream_t st);

size_t fx_scratch_bytes(size_t n_bytes);
cudaError_t launch_fx_parse(const uint8_t *text, size_t n, bool fastq, bool is_final, FxOut out, FxInfo *info,
                            void *scratch, cudaStream_t st);
// lengths_only: per-record output lengths into out_len + exclusive scan into out_off (n+1); else write the text
cudaError_t launch_fx_format(const FmtCfg &cfg, const FxOut &rec, size_t n, const uint8_t *text, uint32_t *out_len,
                             long long *out_off, uint8_t *out, FxInfo *info, bool lengths_only, void *scan_tmp,
                             cudaStream_t st);

}  // namespace shb
