/*
 * seqhash_b200.h — C ABI of libseqhash_b200.so, the B200 (sm_100a) batch hashing engine that
 * replaces the per-record hash path of vmikk/seqhasher.
 *
 * What it replaces.  The reference has no FFI/plugin seam; its hot path is the body of the
 * per-record loop in processSequences:
 *     seqhasher.go:246      seq = bytes.Join(bytes.Fields(seq), nil)     (whitespace strip)
 *     seqhasher.go:249-251  seq = bytes.ToUpper(seq) unless --casesensitive
 *     seqhasher.go:255-259  for each hash type: getHashFunc(t)(seq)
 *     seqhasher.go:290-349  getHashFunc: 12 algorithms, empty-sequence rule, output format
 * A per-record C call is useless for a GPU, so this ABI lifts exactly those statements to
 * BATCH granularity: the host packs records into a CSR batch (one concatenated residue
 * buffer + int64 offsets), submits it, and receives raw digests in input order.  Each entry
 * point below names the reference statement it stands in for.  INTEGRATION.md shows the cgo
 * shim a seqhasher maintainer would add.
 *
 * Conventions: plain C types only; the library owns every buffer it hands out; return codes
 * (0 ok, <0 error, text via shb_last_error()); nothing is written to stderr/stdout; there is
 * NO CPU fallback — every call fails with SHB_ERR_NO_DEVICE when no sm_100 GPU is usable.
 * Raw digests are laid out so that lowercase hex of the bytes equals the reference string:
 * byte digests verbatim; word hashes big-endian (fmt "%016x"/"%08x"); cityhash High||Low
 * (seqhasher.go:316); murmur3 h1||h2 (seqhasher.go:319).
 */
#ifndef SEQHASH_B200_H
#define SEQHASH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Algorithm ids = index into supportedHashTypes (seqhasher.go:41). */
enum shb_algo {
    SHB_SHA1 = 0, SHB_SHA3 = 1, SHB_MD5 = 2, SHB_XXHASH = 3, SHB_XXH3 = 4, SHB_CITYHASH = 5,
    SHB_MURMUR3 = 6, SHB_NTHASH = 7, SHB_BLAKE3 = 8, SHB_K12 = 9, SHB_RAPIDHASH = 10,
    SHB_RAPIDHASH32 = 11, SHB_N_ALGOS = 12
};

/* Normalisation flags of shb_submit / shb_hash_device. */
#define SHB_FLAG_UPPER 1u     /* ASCII a-z -> A-Z: seqhasher.go:249-251 (set unless --casesensitive) */
#define SHB_FLAG_STRIP 2u     /* drop \t \n \v \f \r ' ': seqhasher.go:246 */
#define SHB_FLAG_EMIT_SEQ 4u  /* also return the normalised sequences (record.Seq.Seq = seq, :252) */
/* Non-ASCII input.  For a record that holds a byte >= 0x80 the reference's bytes.Fields / bytes.ToUpper switch to Unicode
 * rules (UTF-8 decoding, Unicode spaces, rune-wise upper-casing that re-encodes invalid bytes as U+FFFD: seqhasher.go:243-251;
 * its README declares such input unsupported).  This engine is byte-wise ASCII: such bytes are hashed as they are.  A host
 * that cannot vouch for ASCII input sets SHB_FLAG_CHECK_ASCII (one more pass over the residues) and refuses the batch when
 * shb_batch_nonascii() is not 0; shb_textproc counts them for free (shb_textproc_nonascii). */
#define SHB_FLAG_CHECK_ASCII 8u
/* Path override for cross-checking (tests): the library normally picks the kernel family from the mean record
 * length; these force the thread-per-record direct kernels / forbid the long-record kernels. */
#define SHB_FLAG_FORCE_DIRECT 0x100u
#define SHB_FLAG_NO_LONG_PATH 0x200u
#define SHB_FLAG_NO_LENGTH_SORT 0x400u   /* thread-per-record kernels in input order instead of length order */
#define SHB_FLAG_NO_FUSED_B3NT 0x800u    /* blake3 + nthash on long records as two passes instead of one */

/* Error codes. */
#define SHB_OK 0
#define SHB_ERR_NO_DEVICE (-1)   /* no CUDA device / not sm_100 / CUDA runtime failure at init */
#define SHB_ERR_BAD_ARG (-2)
#define SHB_ERR_CAPACITY (-3)    /* batch larger than the capacity given to shb_batch_create */
#define SHB_ERR_CUDA (-4)        /* a CUDA call failed; see shb_last_error() */
#define SHB_ERR_OFFSETS (-5)     /* offsets not monotone / record longer than 2^31-1 bytes */
#define SHB_ERR_FORMAT (-6)      /* text is not FASTA/FASTQ ("fastx: invalid FASTA/Q format") */
/* shb_textproc_run only (> 0, not a failure of the call): the text at shb_textproc_consumed() is not a valid record
 * (bad FASTQ '@' / '+' marker, input ending inside a FASTQ record).  The output holds the records before it — the
 * reference's reader yields those before it fails (seqhasher.go:228-239) — and the host then reports
 * "Error reading record: fastx: invalid FASTA/Q format". */
#define SHB_PARTIAL_FORMAT 1
#define SHB_ERR_INFLATE (-7)     /* a BGZF member does not decode, or its ISIZE / CRC-32 differs ("flate: corrupt input", "gzip: invalid checksum") */
/* shb_textproc_run_bgzf only (> 0): the record that crosses the end of the chunk's own members does not end inside the extra
 * members; nothing was produced, the host stages more extra members and calls again. */
#define SHB_NEED_MORE 2

/* Library lifetime.  n_gpus <= 0 uses every visible device; devices 0..n-1 otherwise.  Returns the
 * number of devices in use (>0) or an error.  Idempotent. */
int shb_init(int n_gpus, int flags);
void shb_shutdown(void);
const char *shb_last_error(void);   /* thread-local text of the last failure */
const char *shb_version(void);

/* Name <-> id.  shb_algo_id is isValidHashType + lookup (seqhasher.go:142-149): exact match only,
 * -1 otherwise.  NOTE the reference hashes any name it does not recognise inside getHashFunc
 * (e.g. the untrimmed " md5" that passes validation) as SHA-1 (seqhasher.go:344-346); the host
 * maps such names to SHB_SHA1 before calling — see seqhasher_b200/host. */
int shb_algo_id(const char *name);
const char *shb_algo_name(int algo);
size_t shb_digest_size(int algo);   /* 20,64,16,8,8,16,16,8,32,128,8,4 bytes */

/* ---- host-buffer path: one shb_batch = pinned staging + device buffers + a CUDA stream ---- */

typedef struct shb_batch shb_batch;

/* Capacity in residue bytes and records.  Two or more batches in flight give the double
 * buffering (H2D of batch k+1 overlaps the kernels of batch k).  shb_batch_create uses device 0;
 * shb_batch_create_on binds the batch to one of the devices selected by shb_init. */
shb_batch *shb_batch_create(size_t max_bytes, size_t max_records);
shb_batch *shb_batch_create_on(int device, size_t max_bytes, size_t max_records);
void shb_batch_destroy(shb_batch *b);

/* Pinned host staging the caller fills: record i is residues[offsets[i] .. offsets[i+1]),
 * offsets[0] == 0, n+1 monotone entries (the CSR form of `record.Seq.Seq`, seqhasher.go:241). */
uint8_t *shb_batch_residues(shb_batch *b);
int64_t *shb_batch_offsets(shb_batch *b);

/* Asynchronous: H2D copy, [strip], [upper], one digest set per algo, D2H copy, all on the batch's
 * stream.  Replaces seqhasher.go:246-259 for n_records records at once.  algos may repeat (the
 * reference allows "--hash sha1,sha1"); slot k of the outputs belongs to algos[k]. */
int shb_submit(shb_batch *b, size_t n_records, const int *algos, int n_algos, unsigned flags);
/* Blocks until the batch's outputs are valid in host memory. */
int shb_wait(shb_batch *b);

/* Outputs, valid after shb_wait until the next shb_submit on the same batch.
 * shb_digests: n_records x shb_digest_size(algos[slot]) raw bytes in INPUT ORDER.  A record whose
 * normalised length is 0 has an all-zero slot; the host prints "" for it (empty-sequence rule,
 * seqhasher.go:292-295) by testing shb_out_lengths()[i] == 0.
 * shb_out_lengths: post-normalisation length of each record.
 * shb_out_offsets / shb_out_residues: with SHB_FLAG_EMIT_SEQ, the normalised sequences in CSR form
 * (what the reference writes back to record.Seq.Seq at seqhasher.go:252 for full-record output). */
const uint8_t *shb_digests(shb_batch *b, int slot);
const int64_t *shb_out_lengths(shb_batch *b);
size_t shb_batch_nonascii(shb_batch *b);   /* residue bytes >= 0x80 in the batch (SHB_FLAG_CHECK_ASCII), else 0 */
const int64_t *shb_out_offsets(shb_batch *b);
const uint8_t *shb_out_residues(shb_batch *b);

/* ---- device-resident path (kernel-only timing, chaining after a GPU-side producer) ---- */

/* Same computation with every buffer already in the memory of `device`, enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = the legacy default stream).  d_residues must be 16-byte
 * aligned and readable for 16 bytes past offsets[n]; d_digests[k] receives
 * n x shb_digest_size(algos[k]) bytes; d_out_lengths may be NULL.  d_scratch (may be NULL unless
 * SHB_FLAG_STRIP is set) must hold shb_scratch_bytes(total_bytes, n) bytes.  No host sync. */
size_t shb_scratch_bytes(size_t total_bytes, size_t n_records);
int shb_hash_device(int device, const uint8_t *d_residues, const int64_t *d_offsets, size_t n_records,
                    size_t total_bytes, const int *algos, int n_algos, unsigned flags,
                    uint8_t *const *d_digests, int64_t *d_out_lengths, void *d_scratch, void *stream);

/* Number of kernels this library has launched in this process (monotone; for bench accounting). */
unsigned long long shb_launch_count(void);

/* ---- whole record loop on the GPU: raw FASTA/FASTQ text in, output text out ----
 *
 * Replaces the entire body of processSequences' loop (seqhasher.go:227-283) for one chunk of DECOMPRESSED
 * text: record splitting with the reader semantics of bio's fastx (seqhasher.go:221-239; FASTA records start
 * at a line-initial '>', FASTQ is strict 4-line), whitespace strip (:246), upper-casing (:249-251), the
 * hashes (:255-259), the header rewrite `file;h1;..;hn;Name` (:261-272) and --headersonly / record.Format(0)
 * output (:274-282).  The host keeps file I/O and decompression only.
 *
 * Protocol: write up to shb_textproc_capacity() bytes into shb_textproc_input(), starting at a record marker
 * (the host skips leading blank lines and sniffs the format from the first byte: '>' FASTA, '@' FASTQ, else
 * the reference's "fastx: invalid FASTA/Q format" error); call shb_textproc_run (blocking); write
 * shb_textproc_output() to the sink; move the unconsumed tail (input + shb_textproc_consumed()) to the front
 * and refill; a chunk with more records than the index holds (capacity / 16 + 1024) is consumed only up to that many,
 * so also after the final chunk the host runs the unconsumed tail again until nothing is left.  is_final = this chunk
 * ends the input (a missing last line break is added).  label = the file
 * name field (cfg.inputFileName or --name), NULL for --nofilename / stdin.  flags: SHB_FLAG_UPPER only (the
 * splitter always strips).  shb_textproc_empty_records() = records with an empty sequence, for which the
 * reference logs one line per hash type (seqhasher.go:293). */
#define SHB_FORMAT_FASTA 1
#define SHB_FORMAT_FASTQ 2
typedef struct shb_textproc shb_textproc;
shb_textproc *shb_textproc_create(int device, size_t max_text_bytes);
void shb_textproc_destroy(shb_textproc *t);
uint8_t *shb_textproc_input(shb_textproc *t);   /* the pinned text buffer (capacity + 64 bytes), allocated by the first call; NULL on failure */
size_t shb_textproc_capacity(shb_textproc *t);
int shb_textproc_run(shb_textproc *t, size_t n_bytes, int format, int is_final, const int *algos, int n_algos,
                     unsigned flags, int headers_only, const char *label);
const uint8_t *shb_textproc_output(shb_textproc *t, size_t *out_bytes);
size_t shb_textproc_consumed(shb_textproc *t);
size_t shb_textproc_records(shb_textproc *t);
size_t shb_textproc_empty_records(shb_textproc *t);
/* sequence bytes >= 0x80 seen in the last chunk (whole chunk, also its unconsumed tail): not 0 means the output may differ
 * from the reference's, which normalises such records with Unicode rules (seqhasher.go:243-251); the hosts refuse the input */
size_t shb_textproc_nonascii(shb_textproc *t);
/* device time of the last run, ms: H2D copy (BGZF: + inflate and boundary search), record split, hashes, formatting (incl. its
 * host syncs), D2H copy */
void shb_textproc_last_ms(shb_textproc *t, float *ms5);

/* ---- the same loop for BGZF input (bgzip / htslib blocked gzip): the members are inflated ON THE DEVICE ----
 *
 * Replaces xopen's gzip reader in front of the record loop (seqhasher.go:210-226 -> pgzip, go.mod:20-31) for the one gzip
 * container whose members the host can delimit without decoding: every BGZF member carries its compressed size in a 'BC' extra
 * field and holds at most 64 KiB of text.  Only compressed bytes cross PCIe; one warp inflates one member (ISIZE and CRC-32
 * verified) into the text buffer the splitter reads.  Any other gzip stream stays with the host feeder.
 *
 * The host cuts the member sequence into chunks wherever it likes.  A chunk = n_own members it owns + n_extra members that
 * follow them (look-ahead for the record that crosses the chunk's end).  The chunk's records are those that START after its
 * first text byte and not after the first byte behind its own members; the chunk that begins the input passes first_start =
 * text offset of the first non-blank byte (the host sniffs the format there by inflating the first member itself), every other
 * chunk passes -1.  at_eof = the last extra (or own) member is the last of the input.  Chunks are independent: any order, any
 * number in flight, any GPU.
 *   shb_textproc_bgzf_input: pinned staging for the members' bytes (grown on demand, contents kept); NULL on failure.
 *   members: 3 words per member: offset of the raw DEFLATE data in the staging buffer (behind the member's header), its
 *            length, ISIZE | CRC32 << 32 (the member's trailer).
 * Returns SHB_OK (output / consumed / records / ... as after shb_textproc_run), SHB_PARTIAL_FORMAT, SHB_NEED_MORE, or < 0
 * (SHB_ERR_INFLATE for damaged members, SHB_ERR_CAPACITY when the members' text exceeds shb_textproc_capacity()).
 * When the record index fills up (consumed < the chunk's text, SHB_OK), shb_textproc_resume runs the rest of the
 * device-resident text; repeat until shb_textproc_records() == 0 or everything is consumed. */
uint8_t *shb_textproc_bgzf_input(shb_textproc *t, size_t comp_capacity);
int shb_textproc_run_bgzf(shb_textproc *t, size_t n_comp_bytes, const uint64_t *members, size_t n_own, size_t n_extra,
                          long long first_start, int at_eof, int format, const int *algos, int n_algos,
                          unsigned flags, int headers_only, const char *label);
int shb_textproc_resume(shb_textproc *t, int format, const int *algos, int n_algos, unsigned flags, int headers_only,
                        const char *label);
/* device time of the inflate kernel of the last shb_textproc_run_bgzf, ms (part of the first entry of shb_textproc_last_ms) */
float shb_textproc_last_inflate_ms(shb_textproc *t);
/* 1 when the last device-resident chunk has text left for shb_textproc_resume */
int shb_textproc_pending(shb_textproc *t);

/* ---- measurement helpers (not part of the hashing path) ---- */

/* Deterministic synthetic residues written on the device: record i of fixed length `len` (or of
 * length lens[i] when d_offsets != NULL) gets bases "ACGT"[2 bits] from splitmix64(seed, byte
 * index); lower_permille of the bytes are lower-cased; n_permille become 'N'.  tests/ carries the
 * identical numpy generator. */
int shb_synth_fill(int device, uint8_t *d_residues, size_t total_bytes, uint64_t seed, unsigned lower_permille,
                   unsigned n_permille, void *stream);

/* Register-only integer throughput probes: kind 0 LOP3, 1 IADD3, 2 SHF, 3 IMAD, 4 PRMT,
 * 5 the SHA-1 instruction mix (LOP3:IADD3:SHF:PRMT = 208:165:224:16),
 * 6 LOP3+IMAD interleaved 1:1;
 * 16+v: the SHA-1 compression loop on register data with pipe-balancing variant v (hash_md.cuh);
 * 7 = 16.  Returns lane-instructions per second (613 per compression), or <0. */
double shb_probe_int_peak(int device, int kind, int iters);
/* Pinned host -> device copy rate of the calling process, GB/s (< 0 on failure): `bytes` of cudaHostAlloc memory (write-combined
 * if asked), first touched by the calling thread, copied with plain cudaMemcpyAsync in `piece`-byte pieces for `seconds`.  The
 * roof of the end-to-end numbers; tools/h2d_roof.py runs it from one process per GPU at once. */
double shb_probe_h2d(int device, size_t bytes, size_t piece, double seconds, int write_combined);
/* Debug builds of the library (-DSHB_BOUNDS): shared-memory accesses of the tile path that fell outside the CTA's window since
 * the library was loaded; -1 in a normal build (no checks compiled in).  tools/bounds_check.py. */
long long shb_debug_oob(void);
/* XOR checksum of the final states of the last SHA-1 probe (every variant must give the same value). */
unsigned shb_probe_last_checksum(void);

#ifdef __cplusplus
}
#endif
#endif
