// parse.cu — FASTA/FASTQ record splitter and output formatter on the GPU: the callers either side of the
// hash path (north_star "a FASTA/FASTQ newline scan with a prefix sum that finds
// record and line boundaries and strips line breaks").
//
// Splitter.  Reader semantics follow what the reference gets from bio's fastx reader (seqhasher.go:221-239;
// SURVEY.md App. B): FASTA — a record starts at '>' at the beginning of a line, its name is the rest of that
// line, its sequence every following byte up to the next line-initial '>'; FASTQ — strict 4-line records.
// Whitespace is dropped while compacting (line breaks by the reader, blanks by seqhasher.go:246), so the CSR
// batch that comes out is already stripped.  The text must start at the first record marker and, when it is
// the last chunk of the input, end with '\n' (the host wrapper guarantees both).
//   pass 1  fx_lines_kernel     : per 4 KiB block, position of its last '\n' and its '\n' count
//   scan    fx_scan_kernel      : exclusive prefix (max position, sum of counts) over the blocks
//   pass 2  fx_emit_kernel<..,true>  : classify every byte with its line state -> kept-byte / record counts
//   scan    scan_counts x2
//   pass 3  fx_emit_kernel<..,false> : same classification, then scatter the residues and the per-record
//           fields (sequence offset, name span, quality span)
//   fx_finish_kernel            : record count, closing offset, bytes consumed (chunked input)
// Formatter.  fmt_len_kernel + scan_counts + fmt_write_kernel build the output of seqhasher.go:261-282
// (`file;h1;..;hn;Name`, --headersonly, record.Format(0) for FASTA/FASTQ), hex encoding included.
#include "hash_all.cuh"
#include "inflate.cuh"
#include "parse.cuh"

namespace shb {

constexpr int FX_THREADS = 256;   // 16 bytes per thread, FX_BLK = 4096

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[4], int j) { return (w[j >> 2] >> (8 * (j & 3))) & 0xffu; }

// bit k of the result = byte k of w equals c (exact; SWAR zero-byte test on w ^ c)
__device__ __forceinline__ uint32_t eq_mask4(uint32_t w, uint32_t c4)
{
    const uint32_t z = w ^ c4;
    const uint32_t m = ~(((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z | 0x7f7f7f7fu);   // 0x80 where the byte is zero
    return ((m >> 7) * 0x00204081u) >> 21 & 0xfu;                                  // gather bits 0,8,16,24 -> 0..3
}
// 16-bit masks over the thread's 16 bytes: newlines, whitespace (\t \n \v \f \r ' '), bytes inside the text
__device__ __forceinline__ uint32_t nl_mask16(const uint32_t (&w)[4])
{
    return eq_mask4(w[0], 0x0a0a0a0au) | eq_mask4(w[1], 0x0a0a0a0au) << 4 | eq_mask4(w[2], 0x0a0a0a0au) << 8 |
           eq_mask4(w[3], 0x0a0a0a0au) << 12;
}
__device__ __forceinline__ uint32_t ws_mask16(const uint32_t (&w)[4])
{
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        // c == 0x20 or 9 <= c <= 13: bytes below 0x0e with bit pattern >= 9, tested per byte without carries
        const uint32_t x = w[k];
        const uint32_t lo7 = x & 0x7f7f7f7fu;
        const uint32_t ge9 = lo7 + 0x77777777u;     // bit7 set <=> (c & 0x7f) >= 9
        const uint32_t ge14 = lo7 + 0x72727272u;    // bit7 set <=> (c & 0x7f) >= 14
        const uint32_t ctl = ge9 & ~ge14 & ~x & 0x80808080u;
        m |= ((((ctl >> 7) * 0x00204081u) >> 21 & 0xfu) | eq_mask4(x, 0x20202020u)) << (4 * k);
    }
    return m;
}
__device__ __forceinline__ uint32_t valid_mask16(size_t pos, size_t n)
{
    return pos + 16 <= n ? 0xffffu : (pos < n ? (1u << (uint32_t)(n - pos)) - 1u : 0u);
}

__global__ void __launch_bounds__(FX_THREADS) fx_lines_kernel(const uint8_t *__restrict__ text, size_t n,
                                                              long long *__restrict__ blk_last_nl,
                                                              uint32_t *__restrict__ blk_nl_count)
{
    const size_t pos = (size_t)blockIdx.x * FX_BLK + 16u * threadIdx.x;
    long long last = -1;
    uint32_t cnt = 0;
    if (pos < n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + pos));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const uint32_t m = nl_mask16(w) & valid_mask16(pos, n);
        cnt = __popc(m);
        if (m) last = (long long)pos + (31 - __clz(m));
    }
    __shared__ long long s_last[FX_THREADS / 32];
    __shared__ uint32_t s_cnt[FX_THREADS / 32];
    for (int d = 16; d; d >>= 1) {
        last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0) { s_last[threadIdx.x >> 5] = last; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < FX_THREADS / 32; k++) { last = max(last, s_last[k]); cnt += s_cnt[k]; }
        blk_last_nl[blockIdx.x] = last;
        blk_nl_count[blockIdx.x] = cnt;
    }
}

// Exclusive prefixes over the nb blocks (one CTA; thread t owns a contiguous run of blocks):
// prev_nl[b] = max last_nl over blocks < b (-1 if none), nl_before[b] = newline count of blocks < b;
// element nb holds the totals.
__global__ void __launch_bounds__(1024) fx_scan_kernel(const long long *__restrict__ blk_last_nl,
                                                       const uint32_t *__restrict__ blk_nl_count, size_t nb,
                                                       long long *__restrict__ prev_nl, long long *__restrict__ nl_before)
{
    __shared__ long long s_mx[1024], s_sum[1024];
    const size_t per = (nb + 1023) / 1024;
    const size_t lo = (size_t)threadIdx.x * per;
    const size_t hi = lo + per < nb ? lo + per : nb;
    long long mx = -1, sum = 0;
    for (size_t i = lo; i < hi; i++) { mx = max(mx, blk_last_nl[i]); sum += blk_nl_count[i]; }
    s_mx[threadIdx.x] = mx;
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 per-thread summaries
    for (int d = 1; d < 1024; d <<= 1) {
        long long a = -1, b = 0;
        if ((int)threadIdx.x >= d) { a = s_mx[threadIdx.x - d]; b = s_sum[threadIdx.x - d]; }
        __syncthreads();
        if ((int)threadIdx.x >= d) { s_mx[threadIdx.x] = max(s_mx[threadIdx.x], a); s_sum[threadIdx.x] += b; }
        __syncthreads();
    }
    long long run_mx = threadIdx.x ? s_mx[threadIdx.x - 1] : -1;
    long long run_sum = threadIdx.x ? s_sum[threadIdx.x - 1] : 0;
    for (size_t i = lo; i < hi; i++) {
        prev_nl[i] = run_mx;
        nl_before[i] = run_sum;
        run_mx = max(run_mx, blk_last_nl[i]);
        run_sum += blk_nl_count[i];
    }
    if (threadIdx.x == 1023) { prev_nl[nb] = s_mx[1023]; nl_before[nb] = s_sum[1023]; }
}

// Large inputs: the same scan in three phases over tiles of 1024 blocks (one block summary per thread).
__global__ void __launch_bounds__(1024) fx_tile_reduce_kernel(const long long *__restrict__ blk_last_nl,
                                                              const uint32_t *__restrict__ blk_nl_count, size_t nb,
                                                              long long *__restrict__ tile_last, uint32_t *__restrict__ tile_cnt)
{
    const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
    long long mx = i < nb ? blk_last_nl[i] : -1;
    uint32_t cnt = i < nb ? blk_nl_count[i] : 0;
    __shared__ long long s_mx[32];
    __shared__ uint32_t s_cnt[32];
    for (int d = 16; d; d >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0) { s_mx[threadIdx.x >> 5] = mx; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; k++) { mx = max(mx, s_mx[k]); cnt += s_cnt[k]; }
        tile_last[blockIdx.x] = mx;
        tile_cnt[blockIdx.x] = cnt;
    }
}

__global__ void __launch_bounds__(1024) fx_tile_apply_kernel(const long long *__restrict__ blk_last_nl,
                                                             const uint32_t *__restrict__ blk_nl_count, size_t nb,
                                                             const long long *__restrict__ tile_prev, const long long *__restrict__ tile_before,
                                                             size_t ntiles, long long *__restrict__ prev_nl, long long *__restrict__ nl_before)
{
    __shared__ long long s_mx[1024], s_sum[1024];
    const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
    const long long my_mx = i < nb ? blk_last_nl[i] : -1;
    const long long my_cnt = i < nb ? (long long)blk_nl_count[i] : 0;
    s_mx[threadIdx.x] = my_mx;
    s_sum[threadIdx.x] = my_cnt;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        long long a = -1, b = 0;
        if ((int)threadIdx.x >= d) { a = s_mx[threadIdx.x - d]; b = s_sum[threadIdx.x - d]; }
        __syncthreads();
        if ((int)threadIdx.x >= d) { s_mx[threadIdx.x] = max(s_mx[threadIdx.x], a); s_sum[threadIdx.x] += b; }
        __syncthreads();
    }
    const long long base_mx = tile_prev[blockIdx.x], base_sum = tile_before[blockIdx.x];
    if (i < nb) {
        prev_nl[i] = max(base_mx, threadIdx.x ? s_mx[threadIdx.x - 1] : -1);
        nl_before[i] = base_sum + (threadIdx.x ? s_sum[threadIdx.x - 1] : 0);
    }
    if (blockIdx.x == ntiles - 1 && threadIdx.x == 0) {   // totals
        prev_nl[nb] = tile_prev[ntiles];
        nl_before[nb] = tile_before[ntiles];
    }
}

// Classify the block's bytes.  Thread t owns bytes [bstart + 16t, +16).  The line state entering the
// thread's chunk (start of the current line, its index) comes from a block-wide exclusive scan of
// (last newline position, newline count) seeded with the per-block prefixes.  A record "event" is the
// newline that ends a name line: there the name span and the sequence offset of the record are known.
template <bool FASTQ, bool COUNT>
__global__ void __launch_bounds__(FX_THREADS) fx_emit_kernel(const uint8_t *__restrict__ text, size_t n,
                                                             const long long *__restrict__ prev_nl_blk,
                                                             const long long *__restrict__ nl_before_blk,
                                                             uint32_t *__restrict__ blk_kept, uint32_t *__restrict__ blk_recs,
                                                             const long long *__restrict__ kept_before_blk,
                                                             const long long *__restrict__ recs_before_blk, FxOut out,
                                                             FxInfo *__restrict__ info)
{
    __shared__ long long s_a[FX_THREADS / 32];
    __shared__ uint32_t s_b[FX_THREADS / 32], s_c[FX_THREADS / 32];
    const size_t pos = (size_t)blockIdx.x * FX_BLK + 16u * threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t w[4] = {0, 0, 0, 0};
    if (pos < n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + pos));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    // --- phase A: newline summary of my chunk, block-exclusive scan (max position, count) ---
    const uint32_t valid = valid_mask16(pos, n);
    const uint32_t nl_mask = nl_mask16(w) & valid;
    const long long my_last = nl_mask ? (long long)pos + (31 - __clz(nl_mask)) : -1;
    const uint32_t my_cnt = __popc(nl_mask);
    long long inc_last = my_last;
    uint32_t inc_cnt = my_cnt;
    for (int d = 1; d < 32; d <<= 1) {
        const long long a = __shfl_up_sync(0xffffffffu, inc_last, d);
        const uint32_t b = __shfl_up_sync(0xffffffffu, inc_cnt, d);
        if (lane >= (uint32_t)d) { inc_last = max(inc_last, a); inc_cnt += b; }
    }
    if (lane == 31) { s_a[wid] = inc_last; s_b[wid] = inc_cnt; }
    __syncthreads();
    long long prev_nl = prev_nl_blk[blockIdx.x];
    long long nl_before = nl_before_blk[blockIdx.x];
    for (uint32_t k = 0; k < wid; k++) { prev_nl = max(prev_nl, s_a[k]); nl_before += s_b[k]; }
    {
        const long long a = __shfl_up_sync(0xffffffffu, inc_last, 1);
        const uint32_t b = __shfl_up_sync(0xffffffffu, inc_cnt, 1);
        if (lane) { prev_nl = max(prev_nl, a); nl_before += b; }
    }
    __syncthreads();   // s_b is reused below
    // --- phase B: my 16 bytes are cut into segments by their newlines; a segment shares one line state ---
    long long line_start = prev_nl + 1;
    long long line_no = nl_before;
    bool header_line = !FASTQ && pos < n && text[line_start] == '>';   // FASTA: line starts with '>'
    const uint32_t cand = ~ws_mask16(w) & valid;   // non-blank bytes inside the text
    uint32_t kept_mask = 0, ev_mask = 0;
    {
        uint32_t m = nl_mask, seg_lo = 0;
        while (true) {
            const uint32_t j = m ? (uint32_t)__ffs(m) - 1u : 16u;        // next newline (16 = end of chunk)
            const bool seq_line = FASTQ ? ((line_no & 3) == 1) : !header_line;
            if (seq_line) kept_mask |= cand & ((1u << j) - 1u) & ~((1u << seg_lo) - 1u);
            if (j == 16u) break;
            const bool name_line = FASTQ ? ((line_no & 3) == 0) : header_line;
            if (name_line) ev_mask |= 1u << j;
            line_start = (long long)(pos + j) + 1;
            line_no++;
            if (!FASTQ) header_line = (size_t)line_start < n && text[line_start] == '>';
            seg_lo = j + 1;
            m &= m - 1;
        }
    }
    if (!COUNT) {
        // Sequence bytes >= 0x80: Go's bytes.Fields / bytes.ToUpper switch to Unicode rules for a record that holds one
        // (seqhasher.go:243-251), this engine is byte-wise ASCII.  They are counted so that the host can refuse the input
        // instead of printing digests that differ from the reference's.
        uint32_t hi = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) hi |= ((((w[k] & 0x80808080u) >> 7) * 0x00204081u) >> 21 & 0xfu) << (4 * k);
        hi &= kept_mask;
        if (__any_sync(0xffffffffu, hi != 0)) {
            const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(hi));
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&info->n_nonascii), (unsigned long long)c);
        }
    }
    const uint32_t my_kept = __popc(kept_mask), my_ev = __popc(ev_mask);
    uint32_t inc_k = my_kept, inc_e = my_ev;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, inc_k, d), b = __shfl_up_sync(0xffffffffu, inc_e, d);
        if (lane >= (uint32_t)d) { inc_k += a; inc_e += b; }
    }
    if (lane == 31) { s_b[wid] = inc_k; s_c[wid] = inc_e; }
    __syncthreads();
    if (COUNT) {
        if (threadIdx.x == 0) {
            uint32_t tk = 0, te = 0;
            for (int k = 0; k < FX_THREADS / 32; k++) { tk += s_b[k]; te += s_c[k]; }
            blk_kept[blockIdx.x] = tk;
            blk_recs[blockIdx.x] = te;
        }
        return;
    }
    const long long blk_base = kept_before_blk[blockIdx.x];
    uint32_t local_excl = inc_k - my_kept;
    long long rec_run = recs_before_blk[blockIdx.x] + inc_e - my_ev;
    uint32_t blk_total = 0;
    for (uint32_t k = 0; k < FX_THREADS / 32; k++) {
        if (k < wid) { local_excl += s_b[k]; rec_run += s_c[k]; }
        blk_total += s_b[k];
    }
    long long kept_run = blk_base + local_excl;
    // The block's residues form one contiguous run of the output.  They are compacted in shared memory at
    // the same 16-byte phase as their destination, then copied out with aligned 128-bit stores.
    __shared__ __align__(16) uint8_t s_out[FX_BLK + 32];
    const uint32_t phase = (uint32_t)(blk_base & 15);
    // --- scatter ---
    line_start = prev_nl + 1;
    line_no = nl_before;
    {
        uint8_t *dst = s_out + phase + local_excl;
        if (kept_mask == 0xffffu) {
#pragma unroll
            for (int j = 0; j < 16; j++) dst[j] = (uint8_t)byte_of(w, j);
        } else {
            uint32_t km = kept_mask;
            while (km) {
                const int j = __ffs(km) - 1;
                *dst++ = (uint8_t)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
                km &= km - 1;
            }
        }
    }
    for (uint32_t m = nl_mask; m; m &= m - 1) {
        const uint32_t j = (uint32_t)__ffs(m) - 1u;
        const long long p = (long long)(pos + j);
        const long long kept_here = kept_run + __popc(kept_mask & ((1u << j) - 1u));   // residues before this newline
        long long le = p;                                   // line end, '\r' of a CRLF break dropped
        if (le > line_start && text[le - 1] == '\r') le--;
        if ((ev_mask >> j) & 1u) {
            const long long r = rec_run++;
            if (r <= out.max_records) {                     // offsets / rec_start hold max_records + 1 entries: the
                out.offsets[r] = kept_here;                 // record after the last indexed one marks where to resume
                out.rec_start[r] = line_start;              // (nothing of a name line is ever kept)
            }
            if (r < out.max_records) {
                out.name_start[r] = line_start + 1;         // skip '>' / '@'
                out.name_len[r] = (int)(le - line_start - 1 > 0 ? le - line_start - 1 : 0);
                if (text[line_start] != (FASTQ ? '@' : '>')) atomicMin(reinterpret_cast<unsigned long long *>(&info->first_bad), (unsigned long long)r);
            }
        } else if (FASTQ) {
            const long long kind = line_no & 3;
            const long long r = rec_run - 1;                // the record this line belongs to
            if (kind == 2 && text[line_start] != '+' && r >= 0) atomicMin(reinterpret_cast<unsigned long long *>(&info->first_bad), (unsigned long long)r);
            if (kind == 3 && r >= 0 && r < out.max_records) {
                out.qual_start[r] = line_start;
                out.qual_len[r] = (int)(le - line_start);
            }
        }
        line_start = p + 1;
        line_no++;
    }
    __syncthreads();
    // copy-out: 16-byte unit u of s_out <-> global (blk_base & ~15) + 16u; edge units byte by byte
    {
        uint8_t *gbase = out.residues + (blk_base - phase);
        const uint32_t end = phase + blk_total;                 // valid bytes are s_out[phase, end)
        for (uint32_t u = threadIdx.x; u * 16u < end; u += FX_THREADS) {
            const uint32_t lo = u * 16u, hi = lo + 16u;
            if (lo >= phase && hi <= end) {
                *reinterpret_cast<uint4 *>(gbase + lo) = *reinterpret_cast<const uint4 *>(s_out + lo);
            } else {
                for (uint32_t k = (lo > phase ? lo : phase); k < (hi < end ? hi : end); k++) gbase[k] = s_out[k];
            }
        }
    }
}

// Record count, closing offset and consumed bytes.  FASTA: a record is complete when the next record's
// name line has been seen, or at the end of the final chunk.  FASTQ: when its 4th line has ended.
__global__ void fx_finish_kernel(const long long *__restrict__ kept_prefix, const long long *__restrict__ rec_prefix,
                                 const long long *__restrict__ nl_before, const long long *__restrict__ prev_nl, size_t nb,
                                 size_t n, int fastq, int is_final, FxOut out, FxInfo *__restrict__ info)
{
    const long long n_ev = rec_prefix[nb];
    const long long kept = kept_prefix[nb];
    long long nrec, consumed, status = FX_OK;
    if (n_ev > out.max_records) {
        // The record index is full: hand back the first max_records records (all complete: the name line of the next
        // one was seen, and a FASTQ record's four lines precede it) and let the caller resume at the next record.
        nrec = out.max_records;
        consumed = out.rec_start[nrec];
        status = FX_MORE;
    } else if (fastq) {
        const long long lines = nl_before[nb];
        nrec = lines / 4 < n_ev ? lines / 4 : n_ev;
        if (nrec == n_ev) {          // every record whose name line ended is complete
            out.offsets[nrec] = kept;
            consumed = is_final ? (long long)n : prev_nl[nb] + 1;   // up to the last line break
            // the input ends inside a record (1-3 stray lines after the last complete one): the reader of the reference
            // fails on it after yielding the records before
            if (is_final && (lines & 3) != 0) { status = FX_BAD_MARKER; consumed = prev_nl[nb] + 1; }
        } else {
            consumed = out.rec_start[nrec];   // offsets[nrec] was written by the incomplete record's event
            if (is_final) status = FX_BAD_MARKER;   // truncated last record
        }
    } else if (is_final) {
        nrec = n_ev;
        out.offsets[nrec] = kept;
        consumed = (long long)n;
    } else {
        nrec = n_ev > 0 ? n_ev - 1 : 0;       // the last record may continue in the next chunk
        consumed = n_ev > 0 ? out.rec_start[nrec] : 0;
        if (n_ev == 0) out.offsets[0] = 0;
    }
    const long long bad = info->first_bad;
    if (bad >= 0 && bad < nrec) {             // a malformed record inside the chunk: keep what precedes it
        nrec = bad;
        consumed = out.rec_start[bad];
        status = FX_BAD_MARKER;
    } else if (bad == nrec && bad < n_ev) {   // ... or right after the complete ones
        status = FX_BAD_MARKER;
    }
    if (nrec == 0 && n_ev == 0) out.offsets[0] = 0;
    info->n_records = nrec;
    info->n_residues = out.offsets[nrec];
    info->consumed = consumed;
    info->status = status;
}

// ------------------------------------------------------------------------------------------- formatter

__device__ __forceinline__ uint32_t fmt_name_field_len(const FmtCfg &c, bool empty, int name_len)
{
    uint32_t len = (c.label_len >= 0 ? (uint32_t)c.label_len + 1u : 0u) + (uint32_t)name_len;
    for (int k = 0; k < c.n_algos; k++) len += (empty ? 0u : 2u * (uint32_t)c.dsize[k]) + 1u;
    return len;
}

__global__ void fmt_len_kernel(FmtCfg c, FxOut rec, size_t n, uint32_t *__restrict__ out_len,
                               unsigned long long *__restrict__ n_empty)
{
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const long long slen = rec.offsets[r + 1] - rec.offsets[r];
    const bool empty = slen == 0;
    uint32_t len = fmt_name_field_len(c, empty, rec.name_len[r]) + 1u;   // name + '\n'
    if (!c.headers_only) {
        len += 1u + (uint32_t)slen + 1u;                                  // marker, sequence, '\n'
        if (c.fastq && rec.qual_len[r] > 0) len += 2u + (uint32_t)rec.qual_len[r] + 1u;   // "+\n" quality '\n'
    }
    out_len[r] = len;
    if (empty) atomicAdd(n_empty, 1ULL);
}

// One warp per record; lane j writes output bytes j, j+32, ...  Segments in order:
// [marker] [label ';'] [hex(h_k) ';']* name '\n' [sequence '\n' ['+' '\n' quality '\n']]
__global__ void __launch_bounds__(256) fmt_write_kernel(FmtCfg c, FxOut rec, size_t n, const uint8_t *__restrict__ text,
                                                        const long long *__restrict__ out_off, uint8_t *__restrict__ out)
{
    const uint32_t lane = threadIdx.x & 31u;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        const long long s0 = rec.offsets[r], slen = rec.offsets[r + 1] - s0;
        const bool empty = slen == 0;
        const int name_len = rec.name_len[r];
        const bool fq = c.fastq && rec.qual_len[r] > 0;   // record.Format: FASTQ iff quality present
        const uint32_t b_marker = c.headers_only ? 0u : 1u;
        const uint32_t b_label = b_marker + (c.label_len >= 0 ? (uint32_t)c.label_len + 1u : 0u);
        uint32_t b_hash = b_label;
        for (int k = 0; k < c.n_algos; k++) b_hash += (empty ? 0u : 2u * (uint32_t)c.dsize[k]) + 1u;
        const uint32_t b_name = b_hash + (uint32_t)name_len + 1u;
        const uint32_t b_seq = c.headers_only ? b_name : b_name + (uint32_t)slen + 1u;
        const uint32_t b_end = (uint32_t)(out_off[r + 1] - out_off[r]);
        uint8_t *o = out + out_off[r];
        for (uint32_t j = lane; j < b_end; j += 32) {
            uint8_t ch;
            if (j < b_marker) {
                ch = fq ? '@' : '>';
            } else if (j < b_label) {
                const uint32_t q = j - b_marker;
                ch = q < (uint32_t)c.label_len ? c.label[q] : ';';
            } else if (j < b_hash) {
                uint32_t q = j - b_label;
                ch = ';';
                for (int k = 0; k < c.n_algos; k++) {
                    const uint32_t fl = empty ? 0u : 2u * (uint32_t)c.dsize[k];
                    if (q < fl) {
                        const uint8_t byte = c.digests[k][r * (size_t)c.dsize[k] + (q >> 1)];
                        const uint32_t nib = (q & 1u) ? (byte & 15u) : (byte >> 4);
                        ch = (uint8_t)(nib < 10 ? '0' + nib : 'a' + nib - 10);
                        break;
                    }
                    if (q == fl) break;   // the ';' after this hash
                    q -= fl + 1u;
                }
            } else if (j < b_name) {
                const uint32_t q = j - b_hash;
                ch = q < (uint32_t)name_len ? text[rec.name_start[r] + q] : '\n';
            } else if (j < b_seq) {
                const uint32_t q = j - b_name;
                if (q < (uint32_t)slen) {
                    uint32_t b = rec.residues[s0 + q];
                    if (c.upper && b - 'a' < 26u) b -= 32u;      // seqhasher.go:249-252
                    ch = (uint8_t)b;
                } else {
                    ch = '\n';
                }
            } else {
                const uint32_t q = j - b_seq;   // "+\n" quality "\n"
                if (q == 0) ch = '+';
                else if (q == 1) ch = '\n';
                else if (q - 2 < (uint32_t)rec.qual_len[r]) ch = text[rec.qual_start[r] + (q - 2)];
                else ch = '\n';
            }
            o[j] = ch;
        }
    }
}

// ------------------------------------------------------------------------------------------- BGZF inflate

// One warp (= one CTA) per BGZF member: inflate.cuh decodes the member straight into its place in the text, then the warp
// checks ISIZE and the CRC-32 (32 segments, combined).  Members are independent by construction of the format (each is a
// complete gzip member of <= 64 KiB of text), so a 64 MiB text chunk is ~1000 warps of work, and the compressed bytes — a
// third to a quarter of the text — are all that crosses PCIe.  Only the decoder's tables are in shared memory (10 KiB), so
// ~20 members are resident per SM and several chunks' kernels (one stream per chunk slot) share the GPU.
__global__ void __launch_bounds__(32) bgzf_inflate_kernel(const uint8_t *__restrict__ comp, const unsigned long long *__restrict__ tab,
                                                          size_t n_members, uint8_t *text, unsigned long long *__restrict__ status)
{
    __shared__ InflateScratch scr;
    const size_t mi = blockIdx.x;
    if (mi >= n_members) return;
    const uint32_t lane = threadIdx.x;
    const unsigned long long off = tab[4 * mi], clen = tab[4 * mi + 1], packed = tab[4 * mi + 2], out = tab[4 * mi + 3];
    const uint32_t isize = (uint32_t)packed, want_crc = (uint32_t)(packed >> 32);
    int rc = INF_ERR_OUTPUT;
    if (isize <= INF_WINDOW) {
        uint32_t produced = 0;
        rc = inflate_raw_warp(comp + off, (uint32_t)clen, text + out, isize, &produced, scr, lane, 32u);
        if (rc == INF_OK && produced != isize) rc = INF_ERR_SIZE;
        __syncwarp();
        if (rc == INF_OK) {
            crc_build_tables(scr.u.crc, lane, 32u);
            const uint32_t phase = (uint32_t)(out & 15ull);           // the CRC reads aligned words: start from the 16-byte line
            const uint32_t c = __reduce_xor_sync(0xffffffffu, crc_lane_share(text + (out - phase), phase, isize, scr.u.crc, lane, 32u));
            if (c != want_crc) rc = INF_ERR_CRC;
        }
    }
    if (rc != INF_OK && lane == 0) atomicMin(status, ((unsigned long long)(mi + 1) << 8) | (unsigned)rc);
}

cudaError_t launch_bgzf_inflate(const uint8_t *comp, const unsigned long long *tab, size_t n_members, uint8_t *text,
                                unsigned long long *status, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(status, 0xff, sizeof(unsigned long long), st);   // "no failure" = the largest value (atomicMin)
    if (e != cudaSuccess || n_members == 0) return e;
    bgzf_inflate_kernel<<<(unsigned)n_members, 32, 0, st>>>(comp, tab, n_members, text, status);
    g_launches++;
    return cudaGetLastError();
}

// Record boundaries of a device-resident text whose ends are arbitrary (a run of inflated BGZF members): CTA 0 finds the first
// record start at index >= from0, CTA 1 the first one at index >= from1; res[0], res[1] = the positions, or -1 when there is
// none before n.  A record start is what the splitter itself calls one: FASTA, a line-initial '>'; FASTQ, a line that starts
// with '@' and whose next-but-one line starts with '+' (a quality line may start with '@', but then the line two below it is
// a sequence line, which never starts with '+') — the rule of find_record_start in csrc/host/main.cpp.  from >= 1 (text[from-1]
// decides "line-initial"); from == 0 means "position 0 is a line start" (the very beginning of the input).
// res[2] = index one past the last non-blank byte of the text (trailing blank lines are not records), res[3] = text[n - 1].
__global__ void __launch_bounds__(256) fx_bounds_kernel(const uint8_t *__restrict__ text, size_t n, long long from0, long long from1,
                                                        int fastq, long long *__restrict__ res)
{
    __shared__ long long s_ls[64];    // first line starts found (a FASTQ record start is among any four consecutive ones, plus two lines of look-ahead)
    __shared__ int s_n;
    __shared__ unsigned long long s_min;
    __shared__ uint32_t s_scan[256];
    const long long from = blockIdx.x == 0 ? from0 : from1;
    const uint32_t t = threadIdx.x;
    if (blockIdx.x == 0 && t == 0) {
        long long e = (long long)n;
        while (e > 0 && (text[e - 1] == ' ' || (text[e - 1] >= '\t' && text[e - 1] <= '\r'))) e--;
        res[2] = e;
        res[3] = n ? text[n - 1] : '\n';
    }
    if (from < 0 || (size_t)from >= n) {
        if (t == 0) res[blockIdx.x] = -1;
        return;
    }
    if (t == 0) { s_n = 0; s_min = ~0ull; }
    __syncthreads();
    // line starts p >= from in windows of 4 KiB: p == 0 when from == 0, else text[p - 1] == '\n'
    for (size_t base = (size_t)from; base < n; base += 4096) {
        const size_t p0 = base + 16u * t;
        uint32_t mask = 0;
#pragma unroll 4
        for (int k = 0; k < 16; k++) {
            const size_t p = p0 + k;
            if (p < n) {
                const bool ls = p == 0 ? true : text[p - 1] == '\n';
                if (ls && (fastq || text[p] == '>')) mask |= 1u << k;
            }
        }
        if (!fastq) {
            if (mask) atomicMin(&s_min, (unsigned long long)(p0 + (__ffs(mask) - 1)));
            __syncthreads();
            if (s_min != ~0ull) break;
        } else {
            s_scan[t] = __popc(mask);
            __syncthreads();
            const int had = s_n;
            __syncthreads();
            if (t == 0) {                           // 256 counts: a serial scan by one thread is a few hundred cycles
                uint32_t acc = (uint32_t)had;
                for (int i = 0; i < 256; i++) { const uint32_t v = s_scan[i]; s_scan[i] = acc; acc += v; }
                s_n = (int)(acc < 64u ? acc : 64u);
            }
            __syncthreads();
            uint32_t rank = s_scan[t];
            for (uint32_t mm = mask; mm && rank < 64u; mm &= mm - 1, rank++) s_ls[rank] = (long long)(p0 + (__ffs(mm) - 1));
            __syncthreads();
            if (t == 0) {
                for (int i = (had >= 2 ? had - 2 : 0); i + 2 < s_n; i++)
                    if (text[s_ls[i]] == '@' && text[s_ls[i + 2]] == '+') { s_min = (unsigned long long)s_ls[i]; break; }
            }
            __syncthreads();
            if (s_min != ~0ull || s_n >= 64) break;   // 62 lines without a record start: not FASTQ; the splitter will say so
        }
    }
    __syncthreads();
    if (t == 0) res[blockIdx.x] = s_min != ~0ull ? (long long)s_min : -1;
}

cudaError_t launch_fx_bounds(const uint8_t *text, size_t n, long long from0, long long from1, bool fastq, long long *res, cudaStream_t st)
{
    fx_bounds_kernel<<<2, 256, 0, st>>>(text, n, from0, from1, fastq ? 1 : 0, res);
    g_launches++;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------- launchers


size_t fx_scratch_bytes(size_t n_bytes)
{
    const size_t nb = n_bytes / FX_BLK + 1;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    // last_nl, nl_count, prev_nl, nl_before, blk_kept, blk_recs, kept_prefix, rec_prefix
    const size_t nt = nb / 1024 + 2;
    return up(nb * 8) + up(nb * 4) + 2 * up((nb + 1) * 8) + 2 * up(nb * 4) + 2 * up((nb + 1) * 8) + up(scan_tmp_bytes(nb)) +
           up(nt * 8) + up(nt * 4) + 2 * up((nt + 1) * 8);
}

cudaError_t launch_fx_parse(const uint8_t *text, size_t n, bool fastq, bool is_final, FxOut out, FxInfo *info,
                            void *scratch, cudaStream_t st)
{
    const size_t nb = n / FX_BLK + 1;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    uint8_t *p = static_cast<uint8_t *>(scratch);
    long long *last_nl = reinterpret_cast<long long *>(p); p += up(nb * 8);
    uint32_t *nl_count = reinterpret_cast<uint32_t *>(p); p += up(nb * 4);
    long long *prev_nl = reinterpret_cast<long long *>(p); p += up((nb + 1) * 8);
    long long *nl_before = reinterpret_cast<long long *>(p); p += up((nb + 1) * 8);
    uint32_t *blk_kept = reinterpret_cast<uint32_t *>(p); p += up(nb * 4);
    uint32_t *blk_recs = reinterpret_cast<uint32_t *>(p); p += up(nb * 4);
    long long *kept_prefix = reinterpret_cast<long long *>(p); p += up((nb + 1) * 8);
    long long *rec_prefix = reinterpret_cast<long long *>(p); p += up((nb + 1) * 8);
    void *scan_tmp = p; p += up(scan_tmp_bytes(nb));
    const size_t nt = nb / 1024 + 2;
    long long *tile_last = reinterpret_cast<long long *>(p); p += up(nt * 8);
    uint32_t *tile_cnt = reinterpret_cast<uint32_t *>(p); p += up(nt * 4);
    long long *tile_prev = reinterpret_cast<long long *>(p); p += up((nt + 1) * 8);
    long long *tile_before = reinterpret_cast<long long *>(p);
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(FxInfo), st);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(&info->first_bad, 0x7f, sizeof(long long), st)) != cudaSuccess) return e;   // "none"
    if (fastq && (e = cudaMemsetAsync(out.qual_len, 0, (size_t)out.max_records * sizeof(int), st)) != cudaSuccess) return e;
    fx_lines_kernel<<<(unsigned)nb, FX_THREADS, 0, st>>>(text, n, last_nl, nl_count);
    if (nb <= 4096) {
        fx_scan_kernel<<<1, 1024, 0, st>>>(last_nl, nl_count, nb, prev_nl, nl_before);
    } else {
        const size_t ntiles = (nb + 1023) / 1024;
        fx_tile_reduce_kernel<<<(unsigned)ntiles, 1024, 0, st>>>(last_nl, nl_count, nb, tile_last, tile_cnt);
        fx_scan_kernel<<<1, 1024, 0, st>>>(tile_last, tile_cnt, ntiles, tile_prev, tile_before);
        fx_tile_apply_kernel<<<(unsigned)ntiles, 1024, 0, st>>>(last_nl, nl_count, nb, tile_prev, tile_before, ntiles, prev_nl, nl_before);
        g_launches += 2;
    }
    if (fastq) fx_emit_kernel<true, true><<<(unsigned)nb, FX_THREADS, 0, st>>>(text, n, prev_nl, nl_before, blk_kept, blk_recs, nullptr, nullptr, out, info);
    else fx_emit_kernel<false, true><<<(unsigned)nb, FX_THREADS, 0, st>>>(text, n, prev_nl, nl_before, blk_kept, blk_recs, nullptr, nullptr, out, info);
    g_launches += 3;
    if ((e = scan_counts(blk_kept, nb, reinterpret_cast<int64_t *>(kept_prefix), scan_tmp, st)) != cudaSuccess) return e;
    if ((e = scan_counts(blk_recs, nb, reinterpret_cast<int64_t *>(rec_prefix), scan_tmp, st)) != cudaSuccess) return e;
    if (fastq) fx_emit_kernel<true, false><<<(unsigned)nb, FX_THREADS, 0, st>>>(text, n, prev_nl, nl_before, nullptr, nullptr, kept_prefix, rec_prefix, out, info);
    else fx_emit_kernel<false, false><<<(unsigned)nb, FX_THREADS, 0, st>>>(text, n, prev_nl, nl_before, nullptr, nullptr, kept_prefix, rec_prefix, out, info);
    fx_finish_kernel<<<1, 1, 0, st>>>(kept_prefix, rec_prefix, nl_before, prev_nl, nb, n, fastq ? 1 : 0, is_final ? 1 : 0, out, info);
    g_launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_fx_format(const FmtCfg &cfg, const FxOut &rec, size_t n, const uint8_t *text, uint32_t *out_len,
                             long long *out_off, uint8_t *out, FxInfo *info, bool lengths_only, void *scan_tmp,
                             cudaStream_t st)
{
    if (lengths_only) {
        if (n) fmt_len_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cfg, rec, n, out_len,
                                                                         reinterpret_cast<unsigned long long *>(&info->n_empty));
        g_launches++;
        return scan_counts(out_len, n, reinterpret_cast<int64_t *>(out_off), scan_tmp, st);
    }
    if (n == 0) return cudaSuccess;
    size_t blocks = (n + 7) / 8;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    fmt_write_kernel<<<(unsigned)blocks, 256, 0, st>>>(cfg, rec, n, text, out_off, out);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace shb
