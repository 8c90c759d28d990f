// fastq_kernels.cuh -- the __global__ kernels of the FASTQ-in -> trimmed-FASTQ-out path (sm_100a).
// All of them are HBM-bound byte work: 16-byte vector loads where the layout allows, one pass per stage,
// nothing re-read from PCIe. Per-record logic lives in fastq_core.cuh (shared with the host simulator).
#pragma once
#include <cuda_runtime.h>
#include "fastq_core.cuh"

#define FQ_THREADS 256
#define FQ_TILE (FQ_THREADS * 16)        // bytes of text per CTA of the newline kernels

// device-side bookkeeping of one chunk
struct FqInfo {
    unsigned long long err_key;      // min over errors of (line index << 8 | kind); ~0 = none
    unsigned long long out_bytes;    // total size of the formatted chunk
    long long n_nl;                  // '\n' count of the chunk
    long long n_rec;                 // complete records
    long long consumed;              // bytes up to and including the last complete record
    int nl_overflow;                 // the newline index did not fit its buffer
    int lines_left;                  // lines of a trailing partial record
    int bare_cr;
    int pad;
};

// statistics block shared by all chunks of a call (atomics)
struct FqCounters {
    unsigned long long records, with_adapters, bp_in, bp_out, overflow, invalid;
};

// one atomicAdd per warp: every lane of the warp must call this (convergent)
__device__ __forceinline__ void fq_warp_add(unsigned long long* counter, unsigned v) {
    const unsigned t = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(counter, (unsigned long long)t);
}

// 16 text bytes -> bit i set iff byte i == c. __vcmpeq4 gives 0xff per equal byte; the multiply gathers the four
// low bits into bits 24..27 (distinct powers, no carries).
__device__ __forceinline__ unsigned fq_eq_mask16(const uint4& v, unsigned c4) {
    const unsigned a = ((__vcmpeq4(v.x, c4) & 0x01010101u) * 0x01020408u) >> 24;
    const unsigned b = ((__vcmpeq4(v.y, c4) & 0x01010101u) * 0x01020408u) >> 24;
    const unsigned c = ((__vcmpeq4(v.z, c4) & 0x01010101u) * 0x01020408u) >> 24;
    const unsigned d = ((__vcmpeq4(v.w, c4) & 0x01010101u) * 0x01020408u) >> 24;
    return a | (b << 4) | (c << 8) | (d << 12);
}

__device__ __forceinline__ uint4 fq_load16(const unsigned char* __restrict__ text, long long nbytes, long long off) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (off + 16 <= nbytes) {
        v = *reinterpret_cast<const uint4*>(text + off);
    } else if (off < nbytes) {
        unsigned w[4] = {0u, 0u, 0u, 0u};
        for (int i = 0; off + i < nbytes; i++) w[i >> 2] |= (unsigned)text[off + i] << (8 * (i & 3));
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return v;
}

// pass 1: newlines per tile; bare carriage returns
__global__ void __launch_bounds__(FQ_THREADS) k_fq_nl_count(const unsigned char* __restrict__ text, long long nbytes, int final_text,
                                                            unsigned* __restrict__ tile_counts, FqInfo* __restrict__ info) {
    __shared__ unsigned s_warp[FQ_THREADS / 32];
    const long long off = ((long long)blockIdx.x * FQ_THREADS + threadIdx.x) * 16;
    const uint4 v = fq_load16(text, nbytes, off);
    const unsigned nlm = fq_eq_mask16(v, 0x0a0a0a0au);
    const unsigned crm = fq_eq_mask16(v, 0x0d0d0d0du);
    if (crm) {
        unsigned ok = nlm >> 1;                                   // '\r' at byte i is fine if byte i+1 is '\n'
        if ((crm & 0x8000u) && off + 16 < nbytes && text[off + 16] == '\n') ok |= 0x8000u;
        unsigned bad = crm & ~ok;
        // the text's very last byte has no successor yet: acceptable only when more text follows in a later call
        const long long last = nbytes - 1 - off;
        if (!final_text && last >= 0 && last < 16) bad &= ~(1u << last);
        if (bad) info->bare_cr = 1;
    }
    unsigned c = __popc(nlm);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int w = 0; w < FQ_THREADS / 32; w++) t += s_warp[w];
        tile_counts[blockIdx.x] = t;
    }
}

// pass 2: positions of the newlines, in order (tile_offs = exclusive scan of tile_counts)
__global__ void __launch_bounds__(FQ_THREADS) k_fq_nl_fill(const unsigned char* __restrict__ text, long long nbytes,
                                                           const unsigned* __restrict__ tile_offs, uint32_t* __restrict__ nl,
                                                           long long nl_cap, FqInfo* __restrict__ info) {
    __shared__ unsigned s_warp[FQ_THREADS / 32];
    const long long off = ((long long)blockIdx.x * FQ_THREADS + threadIdx.x) * 16;
    const uint4 v = fq_load16(text, nbytes, off);
    unsigned nlm = fq_eq_mask16(v, 0x0a0a0a0au);
    const unsigned cnt = __popc(nlm);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < FQ_THREADS / 32; w++) wbase += (w < wid) ? s_warp[w] : 0u;
    long long pos = (long long)tile_offs[blockIdx.x] + wbase + (inc - cnt);
    while (nlm) {
        const int b = __ffs((int)nlm) - 1;
        nlm &= nlm - 1;
        if (pos < nl_cap) nl[pos] = (uint32_t)(off + b);
        else info->nl_overflow = 1;
        pos++;
    }
}

// one thread: lines -> complete records, consumed bytes (tile_offs[n_tiles] = total newline count)
// The result is also stored straight into mapped pinned host memory (h_info): a D2H copy of these 64 bytes would
// queue in the copy engine behind the previous chunk's formatted text and stall the pipeline by a whole chunk.
__global__ void k_fq_info(const unsigned* __restrict__ tile_offs, int n_tiles, const uint32_t* __restrict__ nl, long long nl_cap,
                          long long nbytes, int unterminated_last_line, FqInfo* __restrict__ info, FqInfo* h_info) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const long long n_nl = tile_offs[n_tiles];
    const long long lines = n_nl + (unterminated_last_line ? 1 : 0);
    const long long n_rec = lines / 4;
    info->n_nl = n_nl;
    info->n_rec = n_rec;
    info->lines_left = (int)(lines % 4);
    long long consumed = 0;
    if (n_rec > 0) {
        const long long last = 4 * n_rec - 1;                     // line index of the last quality line
        if (last < n_nl) consumed = (last < nl_cap) ? (long long)nl[last] + 1 : 0;
        else consumed = nbytes;                                   // it is the unterminated last line
    }
    info->consumed = consumed;
    *h_info = *info;
}

// end of a chunk's kernels: error key and output size -> mapped pinned host memory
__global__ void k_fq_publish(const FqInfo* __restrict__ info, FqInfo* h_info) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *h_info = *info;
}

__device__ __forceinline__ void fq_report(FqInfo* info, long long line, int kind) {
    atomicMin(&info->err_key, ((unsigned long long)line << 8) | (unsigned long long)kind);
}

// frame + validate: one thread per record (plus one for a trailing partial record when the text is final)
// Error key = (r * key_mul + key_add + line) << 8 | kind: single-end 4 / 0 (the line index), paired-end 16 / 0 for
// file 1 and 16 / 4 for file 2 (pairs are read in the order read 1, read 2, names: io/seqio.py:431-452).
__global__ void __launch_bounds__(256) k_fq_frame(const unsigned char* __restrict__ text, const uint32_t* __restrict__ nl,
                                                  long long n_nl, long long nbytes, long long n_rec, int lines_left,
                                                  FqRec* __restrict__ recs, long long* __restrict__ seq_len64,
                                                  FqInfo* __restrict__ info, int key_mul, int key_add) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec || (r == n_rec && lines_left == 0)) return;
    FqRec R;
    int bad = 0;
    const int avail = r < n_rec ? 4 : lines_left;
    const int kind = fq_frame(text, nl, n_nl, nbytes, r, avail, R, bad);
    if (kind != ATR_FQ_OK) fq_report(info, r * key_mul + key_add + bad, kind);
    if (r < n_rec) {
        recs[r] = R;
        seq_len64[r] = R.seq_len;
    }
}

// 4-bit packing straight from the text (the record table says where each read is): one warp per record, one lane per
// output word, like k_pack. The contiguous ASCII batch is no longer needed for this.
__global__ void __launch_bounds__(256) k_fq_pack(const unsigned char* __restrict__ text, const FqRec* __restrict__ recs, long long n_rec,
                                                 int fold_case, const AtrTables* __restrict__ tables, const uint32_t* __restrict__ woff,
                                                 uint32_t* __restrict__ codes, uint16_t* __restrict__ len_out) {
    __shared__ unsigned char s_iupac[256];
    __shared__ uint32_t s_seq[ATR_PK_READS], s_woff[ATR_PK_READS + 1];
    __shared__ int s_len[ATR_PK_READS], s_esc[ATR_PK_READS];
    const int tid = threadIdx.x;
    s_iupac[tid] = tables->iupac[tid];
    // a sequence line is followed by "\n+...\n" and a quality line as long as itself inside the chunk: the bytes the
    // fast path reads behind a word (up to 3 behind a full word, up to 10 behind the read's last base) exist
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 3) == 0;
    for (long long r0 = (long long)blockIdx.x * ATR_PK_READS; r0 < n_rec; r0 += (long long)gridDim.x * ATR_PK_READS) {
        const int cnt = (int)(n_rec - r0 < ATR_PK_READS ? n_rec - r0 : ATR_PK_READS);
        __syncthreads();
        if (tid <= cnt) s_woff[tid] = woff[r0 + tid];
        if (tid < cnt) { const FqRec R = recs[r0 + tid]; s_seq[tid] = R.seq_b; s_len[tid] = R.seq_len; s_esc[tid] = 0; }
        __syncthreads();
        const uint32_t wbeg = s_woff[0], wcnt = s_woff[cnt] - wbeg;
        for (uint32_t j = tid; j < wcnt; j += 256) {                   // one thread per output word, like k_pack
            const int i = pk_find(s_woff, cnt, wbeg + j);
            const int w = (int)(wbeg + j - s_woff[i]);
            const int len = s_len[i];
            const unsigned char* __restrict__ seq = text + s_seq[i];
            uint32_t v;
            int esc = 0;
            if (!(aligned && pack8_acgt(seq + 8 * w, atr_min(8, len - 8 * w), v))) v = atr::pack_word(seq, len, w, fold_case, s_iupac, &esc);
            codes[wbeg + j] = v;
            if (esc) s_esc[i] = 1;
        }
        __syncthreads();
        if (tid < cnt) len_out[r0 + tid] = (uint16_t)(s_len[tid] | (s_esc[tid] ? ATR_ESC_BIT : 0));
    }
}

__global__ void __launch_bounds__(256) k_fq_word_counts(const FqRec* __restrict__ recs, long long n, uint32_t* __restrict__ counts) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) counts[r] = ((uint32_t)recs[r].seq_len + 7u) >> 3;
}

// The byte-exact kernels (k_locate_gen, k_insert_bytes) address reads through a contiguous ASCII batch + offsets; they
// single-end they only ever touch ESCAPED reads (a byte outside the packed alphabet) and reads longer than the register
// kernels take, so only those are copied there (all = 0). all = 1 copies every read: adapter sets with an adapter that
// always takes the byte-exact kernel, and the paired-end path (a pair goes to k_insert_bytes as a whole when either
// mate is escaped, or holds an X). Warp per record.
__global__ void __launch_bounds__(256) k_fq_gather(const unsigned char* __restrict__ text, const FqRec* __restrict__ recs,
                                                   const long long* __restrict__ offsets, const uint16_t* __restrict__ len16,
                                                   long long n_rec, int all, unsigned char* __restrict__ ascii) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rec) return;
    if (!all && !(len16[r] & ATR_ESC_BIT) && (len16[r] & ATR_LEN_MASK) <= ATR_K1A_MAXN) return;
    const int lane = threadIdx.x & 31;
    const FqRec R = recs[r];
    const unsigned char* src = text + R.seq_b;
    unsigned char* dst = ascii + offsets[r];
    for (int i = lane; i < (int)R.seq_len; i += 32) dst[i] = src[i];
}

// UnconditionalCutter + QualityTrimmer before the adapters (fq_pre_ops): narrows the record table entries; also the
// place where the input bases are counted
__global__ void __launch_bounds__(256) k_fq_pre(const unsigned char* __restrict__ text, FqRec* __restrict__ recs, long long n_rec,
                                                const __grid_constant__ atr_read_ops ops, int side, long long* __restrict__ seq_len64,
                                                unsigned long long* __restrict__ bp_in, FqOpsCounters* __restrict__ oc) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bp = 0, cut = 0, qual = 0, nxs = 0;
    if (r < n_rec) {
        FqRec R = recs[r];
        bp = R.seq_len;
        unsigned c, q, g;
        fq_pre_ops(ops, side, text, R, c, q, g);
        cut = c; qual = q; nxs = g;
        if (c || q || g) { recs[r] = R; seq_len64[r] = R.seq_len; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        bp += __shfl_down_sync(0xffffffffu, bp, d);
        cut += __shfl_down_sync(0xffffffffu, cut, d);
        qual += __shfl_down_sync(0xffffffffu, qual, d);
        nxs += __shfl_down_sync(0xffffffffu, nxs, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bp) atomicAdd(bp_in, bp);
        if (cut) atomicAdd(&oc->bp_cut[side], cut);
        if (qual) atomicAdd(&oc->bp_quality[side], qual);
        if (nxs) atomicAdd(&oc->bp_nextseq[side], nxs);
    }
}

// initial windows
__global__ void __launch_bounds__(256) k_fq_init_win(const FqRec* __restrict__ recs, long long n_rec, uint16_t* __restrict__ fwin,
                                                     unsigned long long* __restrict__ records) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rec) { fwin[2 * r] = 0; fwin[2 * r + 1] = recs[r].seq_len; }
    if (r == 0 && records != nullptr) atomicAdd(records, (unsigned long long)n_rec);
}

// one round of the AdapterCutter loop: shrink the windows, count. hist_front/back: [a][max_len+1][max_errors+1]
__global__ void __launch_bounds__(256) k_fq_apply(const unsigned char* __restrict__ text, const FqRec* __restrict__ recs,
                                                  const atr_match* __restrict__ matches, long long n_rec, int round, int more_rounds,
                                                  const signed char* __restrict__ front_flags, int max_len, int max_errors,
                                                  uint16_t* __restrict__ fwin, uint16_t* __restrict__ rwin, unsigned char* __restrict__ flags,
                                                  unsigned long long* __restrict__ hist_front, unsigned long long* __restrict__ hist_back,
                                                  unsigned long long* __restrict__ adjacent, FqCounters* __restrict__ ctr,
                                                  int adapter_base) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    unsigned long long *h_addr = nullptr, *a_addr = nullptr;       // this read's histogram / adjacent-base counters
    if (r < n_rec) {
        atr_match m = matches[r];
        if (m.adapter >= 0) m.adapter = (int16_t)(m.adapter + adapter_base);   // linked adapters: the back adapter is index 1
        const int lo = fwin[2 * r], hi = fwin[2 * r + 1];
        FqApply a;
        if (m.status == ATR_ST_INVALID) atomicAdd(&ctr->invalid, 1ull);
        if (fq_apply(m, m.adapter >= 0 ? front_flags[m.adapter] : 0, lo, hi, text + recs[r].seq_b, a)) {
            hit = true;
            if (round == 0) flags[r] = 1;                            // read.match is not None (filters.py:170-180)
            if (a.length <= max_len && a.errors <= max_errors) {
                unsigned long long* h = a.front ? hist_front : hist_back;
                h_addr = &h[((size_t)m.adapter * (size_t)(max_len + 1) + (size_t)a.length) * (size_t)(max_errors + 1) + (size_t)a.errors];
            } else {
                atomicAdd(&ctr->overflow, 1ull);
            }
            if (!a.front) a_addr = &adjacent[(size_t)m.adapter * 5 + (size_t)a.adjacent];
            fwin[2 * r] = (uint16_t)a.new_lo; fwin[2 * r + 1] = (uint16_t)a.new_hi;
            if (more_rounds) { rwin[2 * r] = (uint16_t)a.new_lo; rwin[2 * r + 1] = (uint16_t)a.new_hi; }
        } else if (more_rounds) {
            rwin[2 * r] = 0; rwin[2 * r + 1] = 0;                // no match: the loop ends for this read (:145-147)
        }
    }
    // The removed lengths pile up on a few bins (3-mers at the read end, the full adapter length): lanes that hit the
    // same counter add once per warp.
    {
        const int lane = threadIdx.x & 31;
        unsigned peers = __match_any_sync(0xffffffffu, (unsigned long long)h_addr);
        if (h_addr != nullptr && lane == __ffs((int)peers) - 1) atomicAdd(h_addr, (unsigned long long)__popc(peers));
        peers = __match_any_sync(0xffffffffu, (unsigned long long)a_addr);
        if (a_addr != nullptr && lane == __ffs((int)peers) - 1) atomicAdd(a_addr, (unsigned long long)__popc(peers));
    }
    if (round == 0) {
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&ctr->with_adapters, (unsigned long long)__popc(b));
    }
}

__global__ void __launch_bounds__(256) k_fq_outlen(const FqRec* __restrict__ recs, const uint16_t* __restrict__ fwin, long long n_rec,
                                                   long long* __restrict__ out_len, unsigned long long* __restrict__ bp_out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bp = 0;
    if (r < n_rec) {
        const int lo = fwin[2 * r], hi = fwin[2 * r + 1];
        if (lo <= hi) {                                            // lo > hi marks a read the filters discarded
            out_len[r] = fq_out_len(recs[r], lo, hi);
            bp = (unsigned long long)(hi - lo);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bp += __shfl_down_sync(0xffffffffu, bp, d);
    if ((threadIdx.x & 31) == 0 && bp) atomicAdd(bp_out, bp);
}

// Formatted records. One CTA = FQF_RECS consecutive records; their text is one contiguous span of the chunk, staged in
// shared memory by a single TMA bulk copy (cp.async.bulk + mbarrier) while the warps fetch the records' table entries,
// so that the piece-wise copies read shared memory instead of waiting on scattered global loads. A warp formats one
// record at a time: the record is produced as aligned 32-bit words of the output; a word inside one of the four copied
// pieces ('@name', sequence, name after '+', qualities) comes from two aligned loads and a funnel shift; the few words
// that touch a piece boundary or a literal ('\n', '+'), the bytes before the first aligned word and the last bytes
// are written bytewise by one lane each in a single pass. Spans that do not fit the tile are read from global memory.
#define FQF_RECS 32
#define FQF_TILE 32768
__device__ __forceinline__ void fq_format_record(const unsigned char* __restrict__ src /* src[pos] = text[pos] */, const FqRec& R,
                                                 int lo, int hi, int total, unsigned char* __restrict__ dst, int lane) {
    const int w = hi - lo, H = R.hdr_len, P = R.name2 ? H : 1;
    // output coordinates: [0,H) header | H '\n' | [s2,e2) sequence | e2 '\n' | e2+1 '+' | [s3,e3) name | e3 '\n' | [s4,e4) qualities | e4 '\n'
    const int s2 = H + 1, e2 = s2 + w, s3 = e2 + 2, e3 = s3 + P - 1, s4 = e3 + 1, e4 = s4 + w;
    // source offset such that src[off + i] is output byte i inside a piece
    const long long o1 = (long long)R.hdr_b, o2 = (long long)R.seq_b + lo - s2, o3 = (long long)R.hdr_b + 1 - s3,
                    o4 = (long long)R.qual_b + lo - s4;
    const int head = (int)((4u - (unsigned)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u);
    const int nwords = total > head ? (total - head) >> 2 : 0;
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
    for (int t = lane; t < nwords; t += 32) {
        const int i = head + 4 * t, j = i + 3;
        long long off;
        if (j < H) off = o1;
        else if (i >= s2 && j < e2) off = o2;
        else if (i >= s4 && j < e4) off = o4;
        else if (i >= s3 && j < e3) off = o3;
        else continue;                                        // touches a boundary: the bytewise pass below
        const unsigned char* p = src + (off + i);
        const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(p - (sh >> 3));
        const uint32_t a = pw[0];
        dw[t] = sh ? __funnelshift_r(a, pw[1], sh) : a;
    }
    // bytewise: four runs of up to 8 bytes around the boundaries (lanes 0-7, 8-15, 16-23, 24-31) ...
    {
        const int run = lane >> 3, k = lane & 7;
        const int start = run == 0 ? H - 3 : run == 1 ? e2 - 3 : run == 2 ? e3 - 3 : e4 - 3;
        const int stop = run == 0 ? H + 3 : run == 1 ? e2 + 4 : run == 2 ? e3 + 3 : e4;       // inclusive
        const int i = start + k;
        if (i >= 0 && i <= stop && i < total) dst[i] = fq_out_byte(src, R, lo, hi, (uint32_t)i);
    }
    // ... the bytes before the first aligned word and after the last one
    if (lane < head && lane < total) dst[lane] = fq_out_byte(src, R, lo, hi, (uint32_t)lane);
    {
        const int i = head + 4 * nwords + lane;
        if (lane < 4 && i < total) dst[i] = fq_out_byte(src, R, lo, hi, (uint32_t)i);
    }
}

__global__ void __launch_bounds__(256) k_fq_format(const unsigned char* __restrict__ text, const FqRec* __restrict__ recs,
                                                   const uint16_t* __restrict__ fwin, const long long* __restrict__ out_off,
                                                   long long n_rec, unsigned char* __restrict__ out, FqInfo* __restrict__ info) {
    __shared__ __align__(128) unsigned char s_text[FQF_TILE + 16];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = (long long)blockIdx.x * FQF_RECS;
    const int cnt = (int)(n_rec - r0 < FQF_RECS ? n_rec - r0 : FQF_RECS);
    // the CTA's span of the text: from the first record's header to the end of the last record's qualities
    const FqRec Rf = recs[r0], Rl = recs[r0 + cnt - 1];
    const uint32_t a_begin = Rf.hdr_b & ~15u;
    const uint32_t span = ((Rl.qual_b + (uint32_t)Rl.seq_len - a_begin) + 15u) & ~15u;
    const bool use_tma = span <= FQF_TILE && span > 0 && ((reinterpret_cast<uintptr_t>(text) & 15) == 0);
    if (tid == 0 && use_tma) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && use_tma) tma_load_1d(s_text, text + a_begin, span, &s_bar);
    // src[pos] = text[pos]: the staged copy, or the text itself
    const unsigned char* src = use_tma ? (const unsigned char*)s_text - a_begin : text;
    bool waited = !use_tma;
    for (int k = warp; k < cnt; k += 8) {
        const long long r = r0 + k;
        const FqRec R = recs[r];
        const int lo = fwin[2 * r], hi = fwin[2 * r + 1];
        const int total = lo <= hi ? (int)fq_out_len(R, lo, hi) : 0;
        unsigned char* dst = out + out_off[r];
        if (!waited) { mbar_wait(&s_bar, 0); waited = true; }
        if (total) fq_format_record(src, R, lo, hi, total, dst, lane);
        if (r == n_rec - 1 && lane == 0) info->out_bytes = (unsigned long long)(out_off[r] + total);
    }
}

// ---- paired-end ------------------------------------------------------------------------------------------------
struct FqPeCounters {
    unsigned long long records, insert_matches, with_adapters[2], bp_in[2], bp_out[2], overflow, invalid;
    unsigned long long records_corrected, bp_corrected[2], correction_errors;
};

// sequence_names_match for every pair (io/seqio.py:448-452, :773-791)
__global__ void __launch_bounds__(256) k_pe_names(const unsigned char* __restrict__ t1, const FqRec* __restrict__ r1,
                                                  const unsigned char* __restrict__ t2, const FqRec* __restrict__ r2,
                                                  long long n, FqInfo* __restrict__ info) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int nm = fq_names_match(t1, r1[r], t2, r2[r]);
    if (nm) fq_report(info, r * 16 + 8, nm == 1 ? ATR_FQ_PAIR_NAMES : ATR_FQ_EMPTY_NAME);
}

// windows of the per-read fallback (adapter.match_to only where match_insert returned None: modifiers.py:401-406)
__global__ void __launch_bounds__(256) k_pe_prepare(const atr_insert_result* __restrict__ ins, const FqRec* __restrict__ r1,
                                                    const FqRec* __restrict__ r2, long long n, int min_insert_len,
                                                    uint16_t* __restrict__ win1, uint16_t* __restrict__ win2) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int l1 = r1[r].seq_len, l2 = r2[r].seq_len;
    const bool need = l1 >= min_insert_len && l2 >= min_insert_len && ins[r].insert.status == ATR_ST_NONE;
    win1[2 * r] = 0; win1[2 * r + 1] = need ? (uint16_t)l1 : (uint16_t)0;
    win2[2 * r] = 0; win2[2 * r + 1] = need ? (uint16_t)l2 : (uint16_t)0;
}

// InsertAdapterCutter.__call__ after the alignments + trim() + the adapters' statistics, one thread per pair
__global__ void __launch_bounds__(256) k_pe_apply(unsigned char* __restrict__ t1, const FqRec* __restrict__ r1,
                                                  unsigned char* __restrict__ t2, const FqRec* __restrict__ r2,
                                                  const atr_insert_result* __restrict__ ins, const atr_match* __restrict__ fb1,
                                                  const atr_match* __restrict__ fb2, long long n, int symmetric, int min_insert_len,
                                                  int max_len, int max_errors, uint16_t* __restrict__ fwin1, uint16_t* __restrict__ fwin2,
                                                  unsigned long long* __restrict__ hist1, unsigned long long* __restrict__ hist2,
                                                  unsigned long long* __restrict__ adj1, unsigned long long* __restrict__ adj2,
                                                  FqPeCounters* __restrict__ ctr, const __grid_constant__ atr_read_ops ops,
                                                  FqOpsCounters* __restrict__ oc, int mismatch_action,
                                                  const unsigned char* __restrict__ comp, unsigned char* __restrict__ pflags) {
    // pflags != nullptr: MergeOverlapping follows (k_pe_merge_apply). The filters wait for it; this kernel leaves the
    // windows and, per pair, what the merge stage and the filters need to know (FQ_PF_*)
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // per-thread increments, added once per warp at the end (every lane reaches it)
    unsigned n_invalid = 0, n_hit = 0, n_cerr = 0, n_corr = 0, bpc1 = 0, bpc2 = 0, with1 = 0, with2 = 0, n_over = 0, bpn1 = 0, bpn2 = 0;
    int flt = -1;
    unsigned long long *h1 = nullptr, *h2 = nullptr, *a1 = nullptr, *a2 = nullptr;
    if (r < n) {
        const FqRec A = r1[r], B = r2[r];
        int len1 = A.seq_len;
        const int len2 = B.seq_len;
        PeMatch m1, m2;
        int hit = 0, invalid = 0, im[4];
        bool correct = false;
        fq_pe_decide(ins[r], fb1[r], fb2[r], len1, len2, min_insert_len, symmetric, mismatch_action, m1, m2, hit, invalid, correct, im);
        n_invalid = invalid; n_hit = hit;
        if (correct) {                                     // error correction edits this chunk's copy of the text in place
            int c1 = 0, c2 = 0, nl1 = len1;
            if (!fq_pe_correct(t1 + A.seq_b, t1 + A.qual_b, len1, t2 + B.seq_b, t2 + B.qual_b, len2, im[0], im[1], im[2], im[3],
                               mismatch_action, comp, c1, c2, nl1)) n_cerr = 1;
            n_corr = (c1 || c2); bpc1 = (unsigned)c1; bpc2 = (unsigned)c2;
            len1 = nl1;
        }
        FqApply ap;
        bool counted;
        const int k1 = fq_pe_trim(m1, len1, t1 + A.seq_b, ap, counted);
        with1 = m1.present;
        if (counted) {
            if (ap.length <= max_len && ap.errors <= max_errors) h1 = &hist1[(size_t)ap.length * (size_t)(max_errors + 1) + (size_t)ap.errors];
            else n_over++;
            a1 = &adj1[ap.adjacent];
        }
        const int k2 = fq_pe_trim(m2, len2, t2 + B.seq_b, ap, counted);
        with2 = m2.present;
        if (counted) {
            if (ap.length <= max_len && ap.errors <= max_errors) h2 = &hist2[(size_t)ap.length * (size_t)(max_errors + 1) + (size_t)ap.errors];
            else n_over++;
            a2 = &adj2[ap.adjacent];
        }
        // NEndTrimmer on both reads, then the pair filters ("any": either read)
        int lo1 = 0, hi1 = k1, lo2 = 0, hi2 = k2;
        if (ops.trim_n) {
            fq_trim_n(t1 + A.seq_b, lo1, hi1, bpn1);
            fq_trim_n(t2 + B.seq_b, lo2, hi2, bpn2);
        }
        if (pflags) {
            pflags[r] = (unsigned char)((m1.present ? FQ_PF_MATCH1 : 0) | (m2.present ? FQ_PF_MATCH2 : 0) | (hit ? FQ_PF_INSERT : 0) |
                                        (bpc1 ? FQ_PF_CORRECTED1 : 0) | (bpc2 ? FQ_PF_CORRECTED2 : 0));
        } else {
            flt = fq_filter(ops, t1 + A.seq_b, lo1, hi1, m1.present != 0, t2 + B.seq_b, lo2, hi2, m2.present != 0, true);
            if (flt) { lo1 = lo2 = 1; hi1 = hi2 = 0; }
        }
        fwin1[2 * r] = (uint16_t)lo1; fwin1[2 * r + 1] = (uint16_t)hi1;
        fwin2[2 * r] = (uint16_t)lo2; fwin2[2 * r + 1] = (uint16_t)hi2;
    }
    const int lane = threadIdx.x & 31;
    unsigned long long* addrs[4] = {h1, h2, a1, a2};
#pragma unroll
    for (int q = 0; q < 4; q++) {                          // mates trim at the same place: the same bins in both histograms
        const unsigned peers = __match_any_sync(0xffffffffu, (unsigned long long)addrs[q]);
        if (addrs[q] != nullptr && lane == __ffs((int)peers) - 1) atomicAdd(addrs[q], (unsigned long long)__popc(peers));
    }
    fq_warp_add(&ctr->invalid, n_invalid);
    fq_warp_add(&ctr->insert_matches, n_hit);
    fq_warp_add(&ctr->correction_errors, n_cerr);
    fq_warp_add(&ctr->records_corrected, n_corr);
    fq_warp_add(&ctr->bp_corrected[0], bpc1);
    fq_warp_add(&ctr->bp_corrected[1], bpc2);
    fq_warp_add(&ctr->with_adapters[0], with1);
    fq_warp_add(&ctr->with_adapters[1], with2);
    fq_warp_add(&ctr->overflow, n_over);
    fq_warp_add(&oc->bp_n_ends[0], bpn1);
    fq_warp_add(&oc->bp_n_ends[1], bpn2);
    fq_warp_add(&oc->records_written, flt == 0);
    fq_warp_add(&oc->too_short, flt == 1);
    fq_warp_add(&oc->too_long, flt == 2);
    fq_warp_add(&oc->too_many_n, flt == 3);
    fq_warp_add(&oc->discarded_trimmed, flt == 4);
    fq_warp_add(&oc->discarded_untrimmed, flt == 5);
}

// adapter mode: NEndTrimmer + the pair filters after two independent AdapterCutters (flags: bit 0 = read.match is set)
__global__ void __launch_bounds__(256) k_pe_post(const unsigned char* __restrict__ t1, const FqRec* __restrict__ r1,
                                                 const unsigned char* __restrict__ t2, const FqRec* __restrict__ r2, long long n,
                                                 const __grid_constant__ atr_read_ops ops, uint16_t* __restrict__ fwin1,
                                                 uint16_t* __restrict__ fwin2, const unsigned char* __restrict__ flags1,
                                                 const unsigned char* __restrict__ flags2, FqOpsCounters* __restrict__ oc,
                                                 unsigned char* __restrict__ pflags) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bpn1 = 0, bpn2 = 0;
    int flt = -1;
    if (r < n) {
        const FqRec A = r1[r], B = r2[r];
        int lo1 = fwin1[2 * r], hi1 = fwin1[2 * r + 1], lo2 = fwin2[2 * r], hi2 = fwin2[2 * r + 1];
        if (ops.trim_n) {
            fq_trim_n(t1 + A.seq_b, lo1, hi1, bpn1);
            fq_trim_n(t2 + B.seq_b, lo2, hi2, bpn2);
        }
        if (pflags) {                                       // MergeOverlapping follows: the filters wait for it
            pflags[r] = (unsigned char)((flags1[r] ? FQ_PF_MATCH1 : 0) | (flags2[r] ? FQ_PF_MATCH2 : 0));
        } else {
            flt = fq_filter(ops, t1 + A.seq_b, lo1, hi1, flags1[r] != 0, t2 + B.seq_b, lo2, hi2, flags2[r] != 0, true);
            if (flt) { lo1 = lo2 = 1; hi1 = hi2 = 0; }
        }
        fwin1[2 * r] = (uint16_t)lo1; fwin1[2 * r + 1] = (uint16_t)hi1;
        fwin2[2 * r] = (uint16_t)lo2; fwin2[2 * r + 1] = (uint16_t)hi2;
    }
    fq_warp_add(&oc->bp_n_ends[0], bpn1);
    fq_warp_add(&oc->bp_n_ends[1], bpn2);
    fq_warp_add(&oc->records_written, flt == 0);
    fq_warp_add(&oc->too_short, flt == 1);
    fq_warp_add(&oc->too_long, flt == 2);
    fq_warp_add(&oc->too_many_n, flt == 3);
    fq_warp_add(&oc->discarded_trimmed, flt == 4);
    fq_warp_add(&oc->discarded_untrimmed, flt == 5);
}

// ---- MergeOverlapping behind the paired-end modifiers (fastq_core.cuh: FqMergeRec, fq_merge_decide) ----------------
#define FQ_MERGE_BINS 12
#ifndef ATR_MERGE_MAX_READ
#define ATR_MERGE_MAX_READ 4000
#endif
struct FqMergeCounters {
    unsigned long long merged, merged_written, bp_merged, records_corrected, bp_corrected[2], raises, correction_errors;
};

// window lengths of both reads (-> offsets of the contiguous copies the merge kernels read), the pairs' insert_matched
// bytes and the chunk's longest windows (h_max: two ints in mapped pinned memory, zeroed by the host before the launch)
__global__ void __launch_bounds__(256) k_pe_merge_len(const uint16_t* __restrict__ fwin1, const uint16_t* __restrict__ fwin2,
                                                      const unsigned char* __restrict__ pflags, long long n, long long* __restrict__ len1,
                                                      long long* __restrict__ len2, unsigned char* __restrict__ insert_matched,
                                                      int* __restrict__ d_max, const unsigned short* __restrict__ minov,
                                                      unsigned char* __restrict__ keys, int* __restrict__ d_hist) {
    // keys / d_hist: the pair's class for k_merge_warp's two-pair mode, ceil(len2 / 16) if the pair is aligned at all (both
    // reads >= its minimum overlap) else 0, and how many pairs each class has (FQ_MERGE_BINS bins)
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int l1 = 0, l2 = 0, key = -1;
    if (r < n) {
        l1 = (int)fwin1[2 * r + 1] - (int)fwin1[2 * r];
        l2 = (int)fwin2[2 * r + 1] - (int)fwin2[2 * r];
        len1[r] = l1; len2[r] = l2;
        insert_matched[r] = (pflags[r] & FQ_PF_INSERT) ? 1 : 0;
        const int lm = l1 < l2 ? l1 : l2;
        const int mo = (int)minov[lm > ATR_MERGE_MAX_READ ? ATR_MERGE_MAX_READ : lm];
        key = (l1 >= mo && l2 >= mo) ? (l2 + 15) >> 4 : 0;
        if (key > FQ_MERGE_BINS - 1) key = FQ_MERGE_BINS - 1;
        keys[r] = (unsigned char)key;
    }
    const int m1 = __reduce_max_sync(0xffffffffu, l1), m2 = __reduce_max_sync(0xffffffffu, l2);
    if ((threadIdx.x & 31) == 0) {
        if (m1) atomicMax(&d_max[0], m1);
        if (m2) atomicMax(&d_max[1], m2);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && (int)(threadIdx.x & 31) == __ffs((int)peers) - 1) atomicAdd(&d_hist[key], __popc(peers));
}
// bins -> running cursors (exclusive prefix), then every pair takes a slot of its class: `order` lists the pairs class by
// class, so that neighbours in k_merge_warp need the same rows per lane (any order inside a class will do)
__global__ void k_pe_merge_bins(const int* __restrict__ d_hist, int* __restrict__ d_cursor) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < FQ_MERGE_BINS; b++) { d_cursor[b] = acc; acc += d_hist[b]; }
    }
}
__global__ void __launch_bounds__(256) k_pe_merge_order(const unsigned char* __restrict__ keys, long long n, int* __restrict__ d_cursor,
                                                        uint32_t* __restrict__ order) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int key = r < n ? (int)keys[r] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int lane = threadIdx.x & 31, leader = __ffs((int)peers) - 1;
    int base = 0;
    if (key >= 0 && lane == leader) base = atomicAdd(&d_cursor[key], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (key >= 0) order[base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)r;
}
__global__ void k_pe_merge_publish(const int* __restrict__ d_max, int* h_max) {
    if (blockIdx.x == 0 && threadIdx.x < 2) h_max[threadIdx.x] = d_max[threadIdx.x];
}

// the windows as one contiguous ASCII batch (warp per read)
__global__ void __launch_bounds__(256) k_pe_merge_gather(const unsigned char* __restrict__ text, const FqRec* __restrict__ recs,
                                                         const uint16_t* __restrict__ fwin, const long long* __restrict__ offsets,
                                                         long long n, unsigned char* __restrict__ ascii) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const int lane = threadIdx.x & 31;
    const int lo = fwin[2 * r], hi = fwin[2 * r + 1];
    const unsigned char* src = text + recs[r].seq_b + lo;
    unsigned char* dst = ascii + offsets[r];
    for (int i = lane; i < hi - lo; i += 32) dst[i] = src[i];
}

// MergeOverlapping's decision per pair (+ the correction of the overlap), MergedReadFilter, then the filters that had
// to wait (k_pe_apply / k_pe_post with pflags): a merged pair leaves both paired outputs (window mark lo > hi) and
// gets a record length in the merged output; every other pair meets the remaining filters as usual
__global__ void __launch_bounds__(256) k_pe_merge_apply(unsigned char* __restrict__ t1, const FqRec* __restrict__ r1,
                                                        unsigned char* __restrict__ t2, const FqRec* __restrict__ r2,
                                                        const atr_merge_result* __restrict__ mres, const unsigned char* __restrict__ pflags,
                                                        long long n, const __grid_constant__ atr_read_ops ops, int mismatch_action,
                                                        const unsigned char* __restrict__ comp, int write_merged,
                                                        uint16_t* __restrict__ fwin1, uint16_t* __restrict__ fwin2,
                                                        FqMergeRec* __restrict__ mrec, long long* __restrict__ mlen64,
                                                        FqOpsCounters* __restrict__ oc, FqMergeCounters* __restrict__ mc) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int flt = -1;
    unsigned merged = 0, raises = 0, cerr = 0, corr = 0, bpc1 = 0, bpc2 = 0, bpm = 0;
    if (r < n) {
        const FqRec A = r1[r], B = r2[r];
        int lo1 = fwin1[2 * r], hi1 = fwin1[2 * r + 1], lo2 = fwin2[2 * r], hi2 = fwin2[2 * r + 1];
        const int pf = pflags[r];
        FqMergeRec M;
        int c1 = 0, c2 = 0;
        const int d = fq_merge_decide(mres[r], pf, mismatch_action, t1, A, lo1, hi1, t2, B, lo2, hi2, comp, M, c1, c2);
        if (d == -1) raises = 1;
        if (d == -2) cerr = 1;
        corr = (c1 || c2); bpc1 = (unsigned)c1; bpc2 = (unsigned)c2;
        mrec[r] = M;
        if (d == 1) {
            merged = 1;
            if (write_merged) { mlen64[r] = (long long)fq_merged_out_len(A.hdr_len, A.name2, M.mlen); bpm = M.mlen; }
            lo1 = lo2 = 1; hi1 = hi2 = 0;
        } else {
            flt = fq_filter(ops, t1 + A.seq_b, lo1, hi1, (pf & FQ_PF_MATCH1) != 0, t2 + B.seq_b, lo2, hi2, (pf & FQ_PF_MATCH2) != 0, true);
            if (flt) { lo1 = lo2 = 1; hi1 = hi2 = 0; }
        }
        fwin1[2 * r] = (uint16_t)lo1; fwin1[2 * r + 1] = (uint16_t)hi1;
        fwin2[2 * r] = (uint16_t)lo2; fwin2[2 * r + 1] = (uint16_t)hi2;
    }
    fq_warp_add(&mc->merged, merged);
    fq_warp_add(&mc->merged_written, merged && write_merged);
    fq_warp_add(&mc->bp_merged, bpm);
    fq_warp_add(&mc->raises, raises);
    fq_warp_add(&mc->correction_errors, cerr);
    fq_warp_add(&mc->records_corrected, corr);
    fq_warp_add(&mc->bp_corrected[0], bpc1);
    fq_warp_add(&mc->bp_corrected[1], bpc2);
    fq_warp_add(&oc->records_written, flt == 0);
    fq_warp_add(&oc->too_short, flt == 1);
    fq_warp_add(&oc->too_long, flt == 2);
    fq_warp_add(&oc->too_many_n, flt == 3);
    fq_warp_add(&oc->discarded_trimmed, flt == 4);
    fq_warp_add(&oc->discarded_untrimmed, flt == 5);
}

// the merged reads' records (warp per merged pair; read 2's bases come from the copy made BEFORE the correction)
__global__ void __launch_bounds__(256) k_pe_merge_format(const unsigned char* __restrict__ t1, const FqRec* __restrict__ r1,
                                                         const unsigned char* __restrict__ t2, const FqRec* __restrict__ r2,
                                                         const unsigned char* __restrict__ ascii2, const long long* __restrict__ off2,
                                                         const FqMergeRec* __restrict__ mrec, const long long* __restrict__ out_off,
                                                         const unsigned char* __restrict__ comp, long long n,
                                                         unsigned char* __restrict__ out, FqInfo* __restrict__ info) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const int lane = threadIdx.x & 31;
    const FqMergeRec M = mrec[r];
    const FqRec A = r1[r], B = r2[r];
    const uint32_t total = M.action ? fq_merged_out_len(A.hdr_len, A.name2, M.mlen) : 0u;
    unsigned char* dst = out + out_off[r];
    for (uint32_t i = lane; i < total; i += 32)
        dst[i] = fq_merged_out_byte(t1 + A.hdr_b, A.hdr_len, A.name2, t1 + A.seq_b + M.lo1, t1 + A.qual_b + M.lo1, ascii2 + off2[r],
                                    t2 + B.qual_b + M.lo2, comp, M, i);
    if (r == n - 1 && lane == 0) info->out_bytes = (unsigned long long)(out_off[r] + total);
}

// bytes consumed by the first n records (n < the chunk's complete records): -> mapped pinned host memory
__global__ void k_fq_consumed(const uint32_t* __restrict__ nl, long long n, FqInfo* h_info) {
    if (blockIdx.x == 0 && threadIdx.x == 0) h_info->consumed = n > 0 ? (long long)nl[4 * n - 1] + 1 : 0;
}

// NEndTrimmer + the filters after the adapter stage, single-end: final window, or the "discarded" mark (lo > hi)
__global__ void __launch_bounds__(256) k_fq_post(const unsigned char* __restrict__ text, const FqRec* __restrict__ recs, long long n_rec,
                                                 const __grid_constant__ atr_read_ops ops, uint16_t* __restrict__ fwin,
                                                 const unsigned char* __restrict__ flags, FqOpsCounters* __restrict__ oc) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bp_n = 0;
    int f = -1;                                        // -1 no record, 0 kept, 1..5 the filter that fired
    if (r < n_rec) {
        const FqRec R = recs[r];
        int lo = fwin[2 * r], hi = fwin[2 * r + 1];
        const unsigned char* seq = text + R.seq_b;
        if (ops.trim_n) fq_trim_n(seq, lo, hi, bp_n);
        f = fq_filter(ops, seq, lo, hi, flags[r] != 0, seq, 0, 0, false, false);
        if (f) { lo = 1; hi = 0; }
        fwin[2 * r] = (uint16_t)lo; fwin[2 * r + 1] = (uint16_t)hi;
    }
    fq_warp_add(&oc->bp_n_ends[0], bp_n);
    fq_warp_add(&oc->records_written, f == 0);
    fq_warp_add(&oc->too_short, f == 1);
    fq_warp_add(&oc->too_long, f == 2);
    fq_warp_add(&oc->too_many_n, f == 3);
    fq_warp_add(&oc->discarded_trimmed, f == 4);
    fq_warp_add(&oc->discarded_untrimmed, f == 5);
}
