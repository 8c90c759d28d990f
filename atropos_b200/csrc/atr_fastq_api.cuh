// atr_fastq_api.cuh -- host side of atr_trim_fastq_host (included by atr_api.cu after its helpers).
//
// Chunked pipeline over two slots (streams). Per chunk:
//   front:  H2D text -> k_fq_nl_count -> scan -> k_fq_nl_fill -> k_fq_info -> D2H of the 64-byte FqInfo
//           (the host needs the record count to size the launches, and `consumed` to know where the next chunk
//            starts: a chunk is cut at an arbitrary byte, its partial last record is re-sent with the next one)
//   back:   k_fq_frame -> scan -> k_fq_gather -> pack -> [adapter kernels -> k_fq_apply] x times -> k_fq_outlen
//           -> scan -> k_fq_format -> D2H FqInfo -> D2H formatted text
// The H2D of chunk c+1 is issued as soon as chunk c's front is known, so it runs under chunk c's kernels and
// chunk c-1's D2H (PCIe is full duplex).
#pragma once
#include <time.h>
#include "fastq_kernels.cuh"

namespace {

struct FqChunk {
    int slot = 0;
    int64_t start = 0, len = 0;
    bool last = false;
    int64_t n_rec = 0, n_nl = 0, consumed = 0;
    int lines_left = 0;
    int n_tiles = 0;
};

struct FqStatsLayout {
    size_t n_adapters, H;            // H = (max_len+1)*(max_errors+1)
    size_t o_ctr, o_front, o_back, o_adj, o_flags, o_ops, total;
};

FqStatsLayout fq_layout(size_t n_adapters, int max_len, int max_errors) {
    FqStatsLayout L;
    L.n_adapters = n_adapters;
    L.H = (size_t)(max_len + 1) * (size_t)(max_errors + 1);
    L.o_ctr = 0;
    L.o_front = 64;
    L.o_back = L.o_front + n_adapters * L.H * 8;
    L.o_adj = L.o_back + n_adapters * L.H * 8;
    L.o_flags = L.o_adj + n_adapters * 5 * 8;
    L.o_ops = L.o_flags + ((n_adapters + 15) & ~(size_t)15);
    L.total = L.o_ops + ((sizeof(FqOpsCounters) + 15) & ~(size_t)15);
    return L;
}

void fq_add_ops(atr_read_ops_stats& dst, const FqOpsCounters& c) {
    for (int i = 0; i < 2; i++) {
        dst.bp_cut[i] += (int64_t)c.bp_cut[i];
        dst.bp_quality[i] += (int64_t)c.bp_quality[i];
        dst.bp_n_ends[i] += (int64_t)c.bp_n_ends[i];
        dst.bp_nextseq[i] += (int64_t)c.bp_nextseq[i];
    }
    dst.too_short += (int64_t)c.too_short;
    dst.too_long += (int64_t)c.too_long;
    dst.too_many_n += (int64_t)c.too_many_n;
    dst.discarded_trimmed += (int64_t)c.discarded_trimmed;
    dst.discarded_untrimmed += (int64_t)c.discarded_untrimmed;
    dst.records_written += (int64_t)c.records_written;
}

int fq_scan_u32(atr_ctx* ctx, cudaStream_t st, DevBuf& tmp, const unsigned* in, unsigned* out, int n) {
    size_t bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
    int rc = tmp.ensure(bytes);
    if (rc) return fail(ctx, rc, "out of device memory (scan)");
    CU(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, st));
    ctx->launches += 2;
    return ATR_OK;
}

int fq_scan_i64(atr_ctx* ctx, cudaStream_t st, DevBuf& tmp, const long long* in, long long* out, int64_t n) {
    size_t bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st));
    int rc = tmp.ensure(bytes);
    if (rc) return fail(ctx, rc, "out of device memory (scan)");
    CU(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n, st));
    ctx->launches += 2;
    return ATR_OK;
}

// 4-bit packing from the chunk text through the record table (the twin of pack_on_stream)
int fq_pack_on_stream(atr_ctx* ctx, cudaStream_t st, DevBuf& counts, DevBuf& scan_tmp, const unsigned char* d_text,
                      const FqRec* d_recs, int64_t n, int fold_case, uint32_t* d_codes, uint32_t* d_woff, uint16_t* d_len) {
    if (n <= 0) return ATR_OK;
    int rc = counts.ensure((size_t)(n + 1) * sizeof(uint32_t));
    if (rc) return fail(ctx, rc, "out of device memory (pack counts)");
    CU(cudaMemsetAsync(counts.p, 0, (size_t)(n + 1) * sizeof(uint32_t), st));
    k_fq_word_counts<<<grid_for(n, 256), 256, 0, st>>>(d_recs, n, counts.as<uint32_t>());
    LAUNCHED(ctx);
    rc = fq_scan_u32(ctx, st, scan_tmp, counts.as<unsigned>(), d_woff, (int)(n + 1));
    if (rc) return rc;
    k_fq_pack<<<(unsigned)std::min<int64_t>(grid_for(n, ATR_PK_READS), 148 * 8), 256, 0, st>>>(d_text, d_recs, n, fold_case, ctx->d_tables, d_woff, d_codes, d_len);
    LAUNCHED(ctx);
    return ATR_OK;
}

// newline index of the chunk (re-runnable: a too small index buffer is grown and the pass repeated)
int fq_index(atr_ctx* ctx, Slot& s, int side, const FqChunk& c, int final_text, int unterminated) {
    cudaStream_t st = s.stream;
    FqSide& f = s.fq[side];
    FqInfo* d_info = f.info.as<FqInfo>();
    CU(cudaMemsetAsync(d_info, 0, sizeof(FqInfo), st));
    CU(cudaMemsetAsync(d_info, 0xFF, sizeof(unsigned long long), st));       // err_key = "none"
    CU(cudaMemsetAsync(f.tiles.p, 0, (size_t)(c.n_tiles + 1) * sizeof(unsigned), st));
    if (c.n_tiles > 0) {
        k_fq_nl_count<<<(unsigned)c.n_tiles, FQ_THREADS, 0, st>>>(f.text.as<unsigned char>(), c.len, final_text,
                                                                  f.tiles.as<unsigned>(), d_info);
        LAUNCHED(ctx);
    }
    int rc = fq_scan_u32(ctx, st, s.scan_tmp, f.tiles.as<unsigned>(), f.tile_offs.as<unsigned>(), c.n_tiles + 1);
    if (rc) return rc;
    const long long nl_cap = (long long)(f.nl.cap / sizeof(uint32_t));
    if (c.n_tiles > 0) {
        k_fq_nl_fill<<<(unsigned)c.n_tiles, FQ_THREADS, 0, st>>>(f.text.as<unsigned char>(), c.len, f.tile_offs.as<unsigned>(),
                                                                 f.nl.as<uint32_t>(), nl_cap, d_info);
        LAUNCHED(ctx);
    }
    k_fq_info<<<1, 32, 0, st>>>(f.tile_offs.as<unsigned>(), c.n_tiles, f.nl.as<uint32_t>(), nl_cap, c.len, unterminated, d_info,
                                f.hinfo);
    LAUNCHED(ctx);
    return ATR_OK;
}

int fq_front(atr_ctx* ctx, Slot& s, int side, FqChunk& c, const uint8_t* text, int final_text, int unterminated) {
    FqSide& f = s.fq[side];
    c.n_tiles = (int)((c.len + FQ_TILE - 1) / FQ_TILE);
    int rc = f.text.ensure((size_t)c.len + 64);
    if (!rc) rc = f.tiles.ensure((size_t)(c.n_tiles + 2) * sizeof(unsigned));
    if (!rc) rc = f.tile_offs.ensure((size_t)(c.n_tiles + 2) * sizeof(unsigned));
    if (!rc) rc = f.nl.ensure((size_t)(c.len / 8 + 1024) * sizeof(uint32_t));      // >= 8 bytes per line on average; grown on demand
    if (!rc) rc = f.info.ensure(sizeof(FqInfo));
    if (rc) return fail(ctx, rc, "out of device memory (FASTQ chunk)");
    if (!f.hinfo) {                   // mapped pinned (UVA: the same pointer is valid in kernels)
        CU(cudaHostAlloc((void**)&f.hinfo, sizeof(FqInfo), cudaHostAllocMapped));
        CU(cudaStreamCreateWithFlags(&f.out_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&f.ev_d2h, cudaEventDisableTiming));
    }
    if (c.len) CU(cudaMemcpyAsync(f.text.p, text + c.start, (size_t)c.len, cudaMemcpyHostToDevice, s.stream));
    return fq_index(ctx, s, side, c, final_text, unterminated);
}

int fq_back(atr_ctx* ctx, Slot& s, const FqChunk& c, const atr_adapterset* set, const atr_trim_opts* o, const FqStatsLayout& L,
            char* d_stats) {
    cudaStream_t st = s.stream;
    FqSide& f = s.fq[0];
    const int64_t n = c.n_rec;
    FqInfo* d_info = f.info.as<FqInfo>();
    const unsigned char* d_text = f.text.as<unsigned char>();
    FqCounters* d_ctr = (FqCounters*)(d_stats + L.o_ctr);
    int rc = f.recs.ensure((size_t)(n + 1) * sizeof(FqRec));
    if (!rc) rc = f.len64.ensure((size_t)(n + 2) * sizeof(long long));
    if (!rc) rc = s.offsets.ensure((size_t)(n + 2) * sizeof(int64_t));
    if (!rc) rc = f.outoff.ensure((size_t)(n + 2) * sizeof(long long));
    if (!rc) rc = s.ascii.ensure((size_t)c.len + 64);
    if (!rc) rc = s.codes.ensure((size_t)(c.len / 8 + n + 2) * sizeof(uint32_t));
    if (!rc) rc = s.woff.ensure((size_t)(n + 1) * sizeof(uint32_t));
    if (!rc) rc = s.len.ensure((size_t)(n + 1) * sizeof(uint16_t));
    if (!rc) rc = s.out.ensure((size_t)(n + 1) * sizeof(atr_match));
    if (!rc) rc = s.win.ensure((size_t)(n + 1) * 2 * sizeof(uint16_t));
    if (!rc) rc = f.fwin.ensure((size_t)(n + 1) * 2 * sizeof(uint16_t));
    if (!rc) rc = f.outtext.ensure((size_t)c.len + 64);
    if (!rc) rc = f.flags.ensure((size_t)n + 16);
    if (rc) return fail(ctx, rc, "out of device memory (FASTQ records)");
    FqOpsCounters* d_ops = (FqOpsCounters*)(d_stats + L.o_ops);
    // the formatted text of the chunk that used this slot before may still be on its way to the host
    if (f.d2h_pending) { CU(cudaStreamWaitEvent(st, f.ev_d2h, 0)); f.d2h_pending = 0; }
    // frame + validate (one extra thread for a trailing partial record)
    CU(cudaMemsetAsync(f.len64.p, 0, (size_t)(n + 2) * sizeof(long long), st));
    k_fq_frame<<<grid_for(n + 1, 256), 256, 0, st>>>(d_text, f.nl.as<uint32_t>(), c.n_nl, c.len, n, c.lines_left,
                                                     f.recs.as<FqRec>(), f.len64.as<long long>(), d_info, 4, 0);
    LAUNCHED(ctx);
    if (n > 0) {
        // cut / quality-trim before the adapters: from here on a record IS what these modifiers left of it
        k_fq_pre<<<grid_for(n, 256), 256, 0, st>>>(d_text, f.recs.as<FqRec>(), n, o->ops, 0, f.len64.as<long long>(), &d_ctr->bp_in, d_ops);
        LAUNCHED(ctx);
        CU(cudaMemsetAsync(f.flags.p, 0, (size_t)n, st));
        rc = fq_scan_i64(ctx, st, s.scan_tmp, f.len64.as<long long>(), (long long*)s.offsets.p, n + 1);
        if (rc) return rc;
        // the reads are upper-cased for matching only (adapters/__init__.py:349): fold_case = 1
        rc = fq_pack_on_stream(ctx, st, s.counts, s.scan_tmp, d_text, f.recs.as<FqRec>(), n, 1, s.codes.as<uint32_t>(),
                               s.woff.as<uint32_t>(), s.len.as<uint16_t>());
        if (rc) return rc;
        int gather_all = 0;
        for (const atr::HostAdapter& h : set->host) if (!h.k1a_ok) gather_all = 1;
        if (o->linked_back) for (const atr::HostAdapter& h : o->linked_back->host) if (!h.k1a_ok) gather_all = 1;
        k_fq_gather<<<grid_for(n * 32, 256), 256, 0, st>>>(d_text, f.recs.as<FqRec>(), (const long long*)s.offsets.p, s.len.as<uint16_t>(), n,
                                                           gather_all, s.ascii.as<unsigned char>());
        LAUNCHED(ctx);
        k_fq_init_win<<<grid_for(n, 256), 256, 0, st>>>(f.recs.as<FqRec>(), n, f.fwin.as<uint16_t>(), &d_ctr->records);
        LAUNCHED(ctx);
        const atr_adapterset* back = o->linked_back;
        for (int round = 0; round < o->times; round++) {
            rc = locate_on_stream(ctx, s, set, s.codes.as<uint32_t>(), s.woff.as<uint32_t>(), s.len.as<uint16_t>(),
                                  round ? s.win.as<uint16_t>() : nullptr, s.ascii.as<uint8_t>(), s.offsets.as<int64_t>(), 0, 1, n,
                                  s.out.as<atr_match>());
            if (rc) return rc;
            k_fq_apply<<<grid_for(n, 256), 256, 0, st>>>(d_text, f.recs.as<FqRec>(), s.out.as<atr_match>(), n, round,
                                                         (round + 1 < o->times || back) ? 1 : 0, (const signed char*)(d_stats + L.o_flags),
                                                         o->max_len, o->max_errors, f.fwin.as<uint16_t>(), s.win.as<uint16_t>(),
                                                         f.flags.as<unsigned char>(), (unsigned long long*)(d_stats + L.o_front),
                                                         (unsigned long long*)(d_stats + L.o_back),
                                                         (unsigned long long*)(d_stats + L.o_adj), d_ctr, 0);
            LAUNCHED(ctx);
        }
        if (back) {
            // LinkedAdapter.match_to (:671-690): the back adapter inside what the front adapter left, only where the front
            // adapter matched (the round's windows are empty elsewhere)
            rc = locate_on_stream(ctx, s, back, s.codes.as<uint32_t>(), s.woff.as<uint32_t>(), s.len.as<uint16_t>(), s.win.as<uint16_t>(),
                                  s.ascii.as<uint8_t>(), s.offsets.as<int64_t>(), 0, 1, n, s.out.as<atr_match>());
            if (rc) return rc;
            k_fq_apply<<<grid_for(n, 256), 256, 0, st>>>(d_text, f.recs.as<FqRec>(), s.out.as<atr_match>(), n, 1, 0,
                                                         (const signed char*)(d_stats + L.o_flags), o->max_len, o->max_errors,
                                                         f.fwin.as<uint16_t>(), s.win.as<uint16_t>(), f.flags.as<unsigned char>(),
                                                         (unsigned long long*)(d_stats + L.o_front), (unsigned long long*)(d_stats + L.o_back),
                                                         (unsigned long long*)(d_stats + L.o_adj), d_ctr, 1);
            LAUNCHED(ctx);
        }
        // N-end trimming and the filters
        k_fq_post<<<grid_for(n, 256), 256, 0, st>>>(d_text, f.recs.as<FqRec>(), n, o->ops, f.fwin.as<uint16_t>(), f.flags.as<unsigned char>(), d_ops);
        LAUNCHED(ctx);
        CU(cudaMemsetAsync(f.len64.p, 0, (size_t)(n + 2) * sizeof(long long), st));
        k_fq_outlen<<<grid_for(n, 256), 256, 0, st>>>(f.recs.as<FqRec>(), f.fwin.as<uint16_t>(), n, f.len64.as<long long>(), &d_ctr->bp_out);
        LAUNCHED(ctx);
        rc = fq_scan_i64(ctx, st, s.scan_tmp, f.len64.as<long long>(), f.outoff.as<long long>(), n + 1);
        if (rc) return rc;
        k_fq_format<<<grid_for(n, FQF_RECS), 256, 0, st>>>(d_text, f.recs.as<FqRec>(), f.fwin.as<uint16_t>(),
                                                           f.outoff.as<long long>(), n, f.outtext.as<unsigned char>(), d_info);
        LAUNCHED(ctx);
    }
    k_fq_publish<<<1, 32, 0, st>>>(d_info, f.hinfo);
    LAUNCHED(ctx);
    return ATR_OK;
}

// describe the first malformed line of chunk c (key = line index << 8 | kind) in coordinates of the call's text
int fq_describe(atr_ctx* ctx, FqSide& f, const FqChunk& c, int64_t line, int kind, int64_t records_before, atr_fastq_error* err) {
    err->kind = kind;
    err->record = records_before + line / 4;
    err->line_in_record = (int32_t)(line % 4);
    const int64_t total_lines = c.n_nl + ((c.n_rec * 4 + c.lines_left) > c.n_nl ? 1 : 0);
    if (line >= total_lines) {           // "ended prematurely": there is no such line
        err->line_begin = err->line_end = c.start + c.len;
        err->terminated = 0;
        return ATR_OK;
    }
    uint32_t prev = 0, cur = 0;
    if (line > 0) CU(cudaMemcpy(&prev, f.nl.as<uint32_t>() + (line - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
    int64_t b = line > 0 ? (int64_t)prev + 1 : 0, e = c.len;
    err->terminated = 0;
    if (line < c.n_nl) {
        CU(cudaMemcpy(&cur, f.nl.as<uint32_t>() + line, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        e = cur;
        err->terminated = 1;
    }
    err->line_begin = c.start + b;
    err->line_end = c.start + e;         // a "\r" before the "\n" is left to the caller (it has the text)
    return ATR_OK;
}

}  // namespace

// inside the chunk loops an error must not return: D2H copies into the caller's buffers may still be in flight on the
// slots' out_streams -- record it and leave the loop; the common tail synchronises everything before returning
#define CUB(call) { cudaError_t e_ = (call); if (e_ != cudaSuccess) { result = cuda_fail(ctx, e_, #call); break; } }
extern "C" int atr_trim_fastq_host(atr_ctx* ctx, const atr_adapterset* set, const atr_trim_opts* opts, const uint8_t* text,
                                   int64_t nbytes, uint8_t* out_text, int64_t out_cap, int64_t* out_bytes, int64_t* consumed,
                                   atr_trim_stats* stats, atr_fastq_error* err) {
    if (!ctx || !set || !opts || nbytes < 0 || (nbytes > 0 && (!text || !out_text)) || !out_bytes || !consumed || !stats || !err)
        return fail(ctx, ATR_E_ARG, "bad arguments to atr_trim_fastq_host");
    if (set->ctx != ctx) return fail(ctx, ATR_E_ARG, "adapter set belongs to another context");
    if (opts->times < 1 || opts->max_len < 0 || opts->max_len > ATR_MAX_READ || opts->max_errors < 0 || opts->max_errors > 4095)
        return fail(ctx, ATR_E_ARG, "bad atr_trim_opts (times >= 1, 0 <= max_len <= 32767, 0 <= max_errors <= 4095)");
    for (const atr::HostAdapter& h : set->host)
        if (!h.desc.match_to_semantics) return fail(ctx, ATR_E_ARG, "atr_trim_fastq_host needs adapters created with match_to_semantics = 1");
    const atr_adapterset* back = opts->linked_back;
    if (back) {
        if (back->ctx != ctx || set->host.size() != 1 || back->host.size() != 1 || opts->times != 1 || !back->host[0].desc.match_to_semantics)
            return fail(ctx, ATR_E_ARG, "a linked adapter is one front and one back adapter (match_to_semantics = 1) with times == 1");
        const int wf = set->host[0].desc.flags, wb = back->host[0].desc.flags;
        if (!(wf == 8 || wf == 11) || !(wb == 14 || wb == 2))
            return fail(ctx, ATR_E_ARG, "a linked adapter needs a 5' (PREFIX / FRONT) front adapter and a 3' (BACK / SUFFIX) back adapter");
    }
    CU(cudaSetDevice(ctx->device));
    memset(err, 0, sizeof(*err));
    *out_bytes = 0;
    *consumed = 0;
    int64_t chunk = opts->chunk_bytes > 0 ? opts->chunk_bytes : ((int64_t)64 << 20);
    chunk = std::max<int64_t>(4096, std::min<int64_t>(chunk, (int64_t)1 << 30));
    const size_t nA = set->host.size() + (back ? 1 : 0);
    const FqStatsLayout L = fq_layout(nA, opts->max_len, opts->max_errors);
    int rc = ctx->fq_stats.ensure(L.total);
    if (rc) return fail(ctx, rc, "out of device memory (statistics)");
    char* d_stats = ctx->fq_stats.as<char>();
    CU(cudaMemset(d_stats, 0, L.total));
    {
        std::vector<signed char> ff(nA);
        for (size_t a = 0; a < nA; a++) {                // Adapter.__init__: adapters/__init__.py:301-304
            const int w = (back && a == nA - 1) ? back->host[0].desc.flags : set->host[a].desc.flags;
            ff[a] = (w == ATR_SEMIGLOBAL) ? (signed char)-1 : ((w == 14 || w == 2) ? (signed char)0 : (signed char)1);
        }
        CU(cudaMemcpy(d_stats + L.o_flags, ff.data(), nA, cudaMemcpyHostToDevice));
    }
    int64_t opos = 0, records_before = 0, done = 0;
    // measurement knob: ATR_FQ_TIMING=1 prints the kernels' device time and the host's waits to stderr
    static const bool dbg_timing = getenv("ATR_FQ_TIMING") != nullptr;
    double t_front = 0, t_back = 0, t_wait_front = 0, t_wait_back = 0;
    cudaEvent_t evs[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    if (dbg_timing) for (int a = 0; a < 2; a++) for (int b = 0; b < 4; b++) cudaEventCreate(&evs[a][b]);
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const bool final_call = opts->final_chunk != 0;
    auto make_chunk = [&](int slot, int64_t start) {
        FqChunk c;
        c.slot = slot; c.start = start;
        c.len = std::min(chunk, nbytes - start);
        c.last = (start + c.len == nbytes);
        return c;
    };
    auto flags_of = [&](const FqChunk& c, int& final_text, int& unterminated) {
        final_text = (c.last && final_call) ? 1 : 0;
        unterminated = (final_text && c.len > 0 && text[c.start + c.len - 1] != '\n') ? 1 : 0;
    };
    FqChunk cur = make_chunk(0, 0), nxt;
    int ft = 0, ut = 0;
    flags_of(cur, ft, ut);
    rc = fq_front(ctx, ctx->slot[0], 0, cur, text, ft, ut);
    if (rc) return rc;
    int result = ATR_OK;
    while (true) {
        Slot& s = ctx->slot[cur.slot];
        FqSide& f = s.fq[0];
        NvtxRange nvtx_chunk("atr_trim_fastq_host: chunk");
        const double tw0 = now();
        CUB(cudaStreamSynchronize(s.stream));
        t_wait_front += now() - tw0;
        if (f.hinfo->nl_overflow) {                   // more lines than the index was sized for: grow, redo the index
            // (an error leaves the loop instead of returning: copies into the caller's buffers may still be in flight)
            rc = f.nl.ensure((size_t)(f.hinfo->n_nl + 16) * sizeof(uint32_t));
            if (rc) { result = fail(ctx, rc, "out of device memory (newline index)"); break; }
            flags_of(cur, ft, ut);
            rc = fq_index(ctx, s, 0, cur, ft, ut);
            if (rc) { result = rc; break; }
            CUB(cudaStreamSynchronize(s.stream));
        }
        const FqInfo hi = *f.hinfo;
        cur.n_rec = hi.n_rec; cur.n_nl = hi.n_nl; cur.consumed = hi.consumed;
        flags_of(cur, ft, ut);
        cur.lines_left = ft ? hi.lines_left : 0;
        if (hi.bare_cr) {
            err->kind = ATR_FQ_BARE_CR; err->record = -1; err->line_begin = err->line_end = cur.start;
            result = fail(ctx, ATR_E_FORMAT, "FASTQ text holds a carriage return that is not followed by a newline");
            break;
        }
        if (!cur.last && cur.n_rec == 0) {
            err->kind = ATR_FQ_TOO_LONG; err->record = records_before; err->line_begin = err->line_end = cur.start;
            result = fail(ctx, ATR_E_FORMAT, "one FASTQ record is larger than the chunk size");
            break;
        }
        if (cur.n_rec > (int64_t)0x7ffffff0) { result = fail(ctx, ATR_E_LIMIT, "too many records in one chunk"); break; }
        if (!cur.last) {                                 // next chunk starts right after the last complete record
            nxt = make_chunk(cur.slot ^ 1, cur.start + cur.consumed);
            int ft2, ut2;
            flags_of(nxt, ft2, ut2);
            rc = fq_front(ctx, ctx->slot[nxt.slot], 0, nxt, text, ft2, ut2);
            if (rc) { result = rc; break; }
        }
        if (dbg_timing) cudaEventRecord(evs[cur.slot][2], s.stream);
        rc = fq_back(ctx, s, cur, set, opts, L, d_stats);
        if (rc) { result = rc; break; }
        if (dbg_timing) cudaEventRecord(evs[cur.slot][3], s.stream);
        const double tw1 = now();
        CUB(cudaStreamSynchronize(s.stream));
        t_wait_back += now() - tw1;
        if (dbg_timing) { float ms = 0; cudaEventElapsedTime(&ms, evs[cur.slot][2], evs[cur.slot][3]); t_back += ms; }
        const FqInfo hb = *f.hinfo;
        if (hb.err_key != ~0ull) {
            rc = fq_describe(ctx, f, cur, (int64_t)(hb.err_key >> 8), (int)(hb.err_key & 0xFF), records_before, err);
            result = rc ? rc : fail(ctx, ATR_E_FORMAT, "malformed FASTQ (see atr_fastq_error)");
            break;
        }
        if (opos + (int64_t)hb.out_bytes > out_cap) { result = fail(ctx, ATR_E_ARG, "out_cap too small for the trimmed text (nbytes + 1 always suffices)"); break; }
        // own stream: the next chunk's H2D into this slot must not queue behind this copy
        if (hb.out_bytes) {
            CUB(cudaMemcpyAsync(out_text + opos, f.outtext.p, (size_t)hb.out_bytes, cudaMemcpyDeviceToHost, f.out_stream));
            CUB(cudaEventRecord(f.ev_d2h, f.out_stream));
            f.d2h_pending = 1;
        }
        opos += (int64_t)hb.out_bytes;
        records_before += cur.n_rec;
        done = cur.start + ((cur.last && final_call) ? cur.len : cur.consumed);
        if (cur.last) break;
        cur = nxt;
    }
    for (int k = 0; k < 2; k++) {
        CU(cudaStreamSynchronize(ctx->slot[k].stream));
        if (ctx->slot[k].fq[0].out_stream) CU(cudaStreamSynchronize(ctx->slot[k].fq[0].out_stream));
        ctx->slot[k].fq[0].d2h_pending = 0;
    }
    if (dbg_timing) {
        fprintf(stderr, "[atr_trim_fastq_host] back kernels %.2f ms, host waits: front %.2f ms, back %.2f ms (front kernels %.2f)\n",
                t_back, t_wait_front, t_wait_back, t_front);
        for (int a = 0; a < 2; a++) for (int b = 0; b < 4; b++) cudaEventDestroy(evs[a][b]);
    }
    if (result != ATR_OK) return result;
    // statistics: device block -> added to the caller's arrays
    std::vector<char> hst(L.total);
    CU(cudaMemcpy(hst.data(), d_stats, L.total, cudaMemcpyDeviceToHost));
    const FqCounters* hc = (const FqCounters*)(hst.data() + L.o_ctr);
    if (hc->invalid) {
        err->kind = ATR_FQ_INVALID_MATCH; err->record = -1;
        return fail(ctx, ATR_E_FORMAT, "an alignment of length <= errors: Match.__init__ raises ValueError in the reference");
    }
    stats->records += (int64_t)hc->records;
    stats->with_adapters += (int64_t)hc->with_adapters;
    stats->bp_in += (int64_t)hc->bp_in;
    stats->bp_out += (int64_t)hc->bp_out;
    stats->overflow += (int64_t)hc->overflow;
    const unsigned long long* hf = (const unsigned long long*)(hst.data() + L.o_front);
    const unsigned long long* hbk = (const unsigned long long*)(hst.data() + L.o_back);
    const unsigned long long* ha = (const unsigned long long*)(hst.data() + L.o_adj);
    if (stats->errors_front) for (size_t i = 0; i < nA * L.H; i++) stats->errors_front[i] += (int64_t)hf[i];
    if (stats->errors_back) for (size_t i = 0; i < nA * L.H; i++) stats->errors_back[i] += (int64_t)hbk[i];
    if (stats->adjacent_bases) for (size_t i = 0; i < nA * 5; i++) stats->adjacent_bases[i] += (int64_t)ha[i];
    fq_add_ops(stats->ops, *(const FqOpsCounters*)(hst.data() + L.o_ops));
    *out_bytes = opos;
    *consumed = done;
    ctx->last_ms = -1.f;
    return ATR_OK;
}

// =====================================================================================================================
// Paired-end twin: two texts read in lockstep. Per step one chunk of each text on the same slot; the number of pairs
// is the smaller of the two chunks' complete records, and each side resumes right after its n-th record.
// =====================================================================================================================
namespace {

struct PeStep {
    int slot = 0;
    FqChunk c[2];
    int64_t n = 0;
};

// statistics block: pair counters | per read: AdapterCutter counters, back / front histograms, adjacent bases, front flags
struct PeLayout { size_t H, nA[2], o_ctr, o_side[2], o_hist[2], o_front[2], o_adj[2], o_flags[2], o_ops, o_merge, total; };

// MergeOverlapping behind the modifiers (atr_trim_fastq_pe_merge_host): what pe_back needs of it
struct PeMerge {
    MergeTables tb;
    double error_rate = 0;
    int write_merged = 0;            // a merged output exists; else merged pairs are discarded
};

PeLayout pe_layout(int max_len, int max_errors, size_t nA0, size_t nA1) {
    PeLayout L;
    L.H = (size_t)(max_len + 1) * (size_t)(max_errors + 1);
    L.nA[0] = nA0 ? nA0 : 1; L.nA[1] = nA1 ? nA1 : 1;
    L.o_ctr = 0;
    size_t o = 128;
    for (int f = 0; f < 2; f++) {
        L.o_side[f] = o; o += 64;
        L.o_hist[f] = o; o += L.nA[f] * L.H * 8;
        L.o_front[f] = o; o += L.nA[f] * L.H * 8;
        L.o_adj[f] = o; o += ((L.nA[f] * 5 * 8 + 63) & ~(size_t)63);
        L.o_flags[f] = o; o += ((L.nA[f] + 63) & ~(size_t)63);
    }
    L.o_ops = o;
    L.o_merge = L.o_ops + ((sizeof(FqOpsCounters) + 15) & ~(size_t)15);
    L.total = L.o_merge + ((sizeof(FqMergeCounters) + 15) & ~(size_t)15);
    return L;
}

// per-side buffers of a slot that the single-end path keeps in unnumbered fields
struct PeSide {
    DevBuf &ascii, &offsets, &counts, &woff, &codes, &len, &win, &out;
};
PeSide pe_side(Slot& s, int f) {
    if (f == 0) return PeSide{s.ascii, s.offsets, s.counts, s.woff, s.codes, s.len, s.win, s.out};
    return PeSide{s.ascii2, s.offsets2, s.counts2, s.woff2, s.codes2, s.len2, s.win2, s.out2};
}

int pe_back(atr_ctx* ctx, Slot& s, const PeStep& P, const atr_insertset* iset, const atr_adapterset* const sets[2],
            const atr_trim_pe_opts* o, const PeLayout& L, char* d_stats, const PeMerge* mg) {
    cudaStream_t st = s.stream;
    const int64_t n = P.n;
    FqPeCounters* d_ctr = (FqPeCounters*)(d_stats + L.o_ctr);
    FqOpsCounters* d_ops = (FqOpsCounters*)(d_stats + L.o_ops);
    FqInfo* d_info0 = s.fq[0].info.as<FqInfo>();
    for (int f = 0; f < 2; f++) {
        FqSide& q = s.fq[f];
        PeSide b = pe_side(s, f);
        const int64_t len = P.c[f].len;
        int rc = q.recs.ensure((size_t)(n + 1) * sizeof(FqRec));
        if (!rc) rc = q.len64.ensure((size_t)(n + 2) * sizeof(long long));
        if (!rc) rc = b.offsets.ensure((size_t)(n + 2) * sizeof(int64_t));
        if (!rc) rc = q.outoff.ensure((size_t)(n + 2) * sizeof(long long));
        if (!rc) rc = b.ascii.ensure((size_t)len + 64);
        if (!rc) rc = b.codes.ensure((size_t)(len / 8 + n + 2) * sizeof(uint32_t));
        if (!rc) rc = b.woff.ensure((size_t)(n + 1) * sizeof(uint32_t));
        if (!rc) rc = b.len.ensure((size_t)(n + 1) * sizeof(uint16_t));
        if (!rc) rc = b.out.ensure((size_t)(n + 1) * sizeof(atr_match));
        if (!rc) rc = b.win.ensure((size_t)(n + 1) * 2 * sizeof(uint16_t));
        if (!rc) rc = q.fwin.ensure((size_t)(n + 1) * 2 * sizeof(uint16_t));
        if (!rc) rc = q.outtext.ensure((size_t)len + 64);
        if (!rc) rc = q.flags.ensure((size_t)n + 16);
        if (rc) return fail(ctx, rc, "out of device memory (paired FASTQ records)");
        if (q.d2h_pending) { CU(cudaStreamWaitEvent(st, q.ev_d2h, 0)); q.d2h_pending = 0; }
    }
    int rc = iset ? s.ins_out.ensure((size_t)(n + 1) * sizeof(atr_insert_result)) : ATR_OK;
    if (rc) return fail(ctx, rc, "out of device memory (insert results)");
    // merged reads: pair flags, insert_matched bytes, merge results, merge records, the third formatted text
    FqSide& qm = s.fqm;
    unsigned char* d_pflags = nullptr;
    if (mg) {
        rc = qm.flags.ensure((size_t)n + 16);
        if (!rc) rc = qm.text.ensure((size_t)(n + 1) * sizeof(atr_merge_result));
        if (!rc) rc = qm.recs.ensure((size_t)(n + 1) * sizeof(FqMergeRec));
        if (!rc) rc = qm.len64.ensure((size_t)(n + 2) * sizeof(long long));
        if (!rc) rc = qm.outoff.ensure((size_t)(n + 2) * sizeof(long long));
        if (!rc) rc = qm.nl.ensure(256);                                   // longest windows (2 ints), class histogram, cursors
        if (!rc) rc = qm.tile_offs.ensure((size_t)(n + 1) * sizeof(uint32_t));   // the pairs in class order
        if (!rc) rc = qm.tiles.ensure(2 * ((size_t)n + 16));               // insert_matched bytes | class keys
        if (!rc) rc = qm.info.ensure(sizeof(FqInfo));
        if (!rc && mg->write_merged) rc = qm.outtext.ensure((size_t)(P.c[0].len + P.c[1].len) + 64);
        if (rc) return fail(ctx, rc, "out of device memory (merged reads)");
        if (!qm.hinfo) {
            CU(cudaHostAlloc((void**)&qm.hinfo, sizeof(FqInfo), cudaHostAllocMapped));
            CU(cudaStreamCreateWithFlags(&qm.out_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&qm.ev_d2h, cudaEventDisableTiming));
        }
        if (qm.d2h_pending) { CU(cudaStreamWaitEvent(st, qm.ev_d2h, 0)); qm.d2h_pending = 0; }
        d_pflags = qm.flags.as<unsigned char>();
    }
    // frame + validate both sides, then the names; every error goes to side 0's key
    for (int f = 0; f < 2; f++) {
        FqSide& q = s.fq[f];
        CU(cudaMemsetAsync(q.len64.p, 0, (size_t)(n + 2) * sizeof(long long), st));
        k_fq_frame<<<grid_for(n, 256), 256, 0, st>>>(q.text.as<unsigned char>(), q.nl.as<uint32_t>(), P.c[f].n_nl, P.c[f].len, n, 0,
                                                     q.recs.as<FqRec>(), q.len64.as<long long>(), d_info0, 16, 4 * f);
        LAUNCHED(ctx);
    }
    k_pe_names<<<grid_for(n, 256), 256, 0, st>>>(s.fq[0].text.as<unsigned char>(), s.fq[0].recs.as<FqRec>(),
                                                 s.fq[1].text.as<unsigned char>(), s.fq[1].recs.as<FqRec>(), n, d_info0);
    LAUNCHED(ctx);
    for (int f = 0; f < 2; f++) {
        FqSide& q = s.fq[f];
        PeSide b = pe_side(s, f);
        k_fq_pre<<<grid_for(n, 256), 256, 0, st>>>(q.text.as<unsigned char>(), q.recs.as<FqRec>(), n, o->ops, f, q.len64.as<long long>(),
                                                   &d_ctr->bp_in[f], d_ops);
        LAUNCHED(ctx);
        rc = fq_scan_i64(ctx, st, s.scan_tmp, q.len64.as<long long>(), (long long*)b.offsets.p, n + 1);
        if (rc) return rc;
        // match_insert compares the reads as they are (no upper-casing, align/__init__.py:250-267): fold_case = 0;
        // lower-case reads come out "escaped" and take the byte-exact kernels in both stages
        // (adapter mode only has match_to, which upper-cases: fold there)
        rc = fq_pack_on_stream(ctx, st, b.counts, s.scan_tmp, q.text.as<unsigned char>(), q.recs.as<FqRec>(), n, iset ? 0 : 1,
                               b.codes.as<uint32_t>(), b.woff.as<uint32_t>(), b.len.as<uint16_t>());
        if (rc) return rc;
        k_fq_gather<<<grid_for(n * 32, 256), 256, 0, st>>>(q.text.as<unsigned char>(), q.recs.as<FqRec>(), (const long long*)b.offsets.p,
                                                           b.len.as<uint16_t>(), n, 1, b.ascii.as<unsigned char>());
        LAUNCHED(ctx);
        k_fq_init_win<<<grid_for(n, 256), 256, 0, st>>>(q.recs.as<FqRec>(), n, q.fwin.as<uint16_t>(), f == 0 ? &d_ctr->records : nullptr);
        LAUNCHED(ctx);
    }
    PeSide b0 = pe_side(s, 0), b1 = pe_side(s, 1);
    if (iset) {
        rc = insert_on_stream(ctx, st, iset, b0.codes.as<uint32_t>(), b0.woff.as<uint32_t>(), b0.len.as<uint16_t>(),
                              b1.codes.as<uint32_t>(), b1.woff.as<uint32_t>(), b1.len.as<uint16_t>(),
                              b0.ascii.as<uint8_t>(), b0.offsets.as<int64_t>(), 0, b1.ascii.as<uint8_t>(), b1.offsets.as<int64_t>(), 0, n,
                              s.ins_out.as<atr_insert_result>());
        if (rc) return rc;
        k_pe_prepare<<<grid_for(n, 256), 256, 0, st>>>(s.ins_out.as<atr_insert_result>(), s.fq[0].recs.as<FqRec>(), s.fq[1].recs.as<FqRec>(), n,
                                                       o->min_insert_overlap, b0.win.as<uint16_t>(), b1.win.as<uint16_t>());
        LAUNCHED(ctx);
        for (int f = 0; f < 2; f++) {               // adapter{1,2}.match_to(read{1,2}) where there was no insert match
            PeSide b = pe_side(s, f);
            rc = locate_on_stream(ctx, s, sets[f], b.codes.as<uint32_t>(), b.woff.as<uint32_t>(), b.len.as<uint16_t>(), b.win.as<uint16_t>(),
                                  b.ascii.as<uint8_t>(), b.offsets.as<int64_t>(), 0, 1, n, b.out.as<atr_match>());
            if (rc) return rc;
        }
        k_pe_apply<<<grid_for(n, 256), 256, 0, st>>>(s.fq[0].text.as<unsigned char>(), s.fq[0].recs.as<FqRec>(),
                                                     s.fq[1].text.as<unsigned char>(), s.fq[1].recs.as<FqRec>(),
                                                     s.ins_out.as<atr_insert_result>(), b0.out.as<atr_match>(), b1.out.as<atr_match>(), n,
                                                     o->symmetric, o->min_insert_overlap, o->max_len, o->max_errors,
                                                     s.fq[0].fwin.as<uint16_t>(), s.fq[1].fwin.as<uint16_t>(),
                                                     (unsigned long long*)(d_stats + L.o_hist[0]), (unsigned long long*)(d_stats + L.o_hist[1]),
                                                     (unsigned long long*)(d_stats + L.o_adj[0]), (unsigned long long*)(d_stats + L.o_adj[1]), d_ctr,
                                                     o->ops, d_ops, o->mismatch_action, iset->dev.comp, d_pflags);
        LAUNCHED(ctx);
    } else {
        // "--aligner adapter": two independent AdapterCutters (the single-end adapter stage per read), then the pair filters
        for (int f = 0; f < 2; f++) {
            FqSide& q = s.fq[f];
            PeSide b = pe_side(s, f);
            CU(cudaMemsetAsync(q.flags.p, 0, (size_t)n, st));
            if (!sets[f]) continue;
            for (int round = 0; round < o->times; round++) {
                rc = locate_on_stream(ctx, s, sets[f], b.codes.as<uint32_t>(), b.woff.as<uint32_t>(), b.len.as<uint16_t>(),
                                      round ? b.win.as<uint16_t>() : nullptr, b.ascii.as<uint8_t>(), b.offsets.as<int64_t>(), 0, 1, n,
                                      b.out.as<atr_match>());
                if (rc) return rc;
                k_fq_apply<<<grid_for(n, 256), 256, 0, st>>>(q.text.as<unsigned char>(), q.recs.as<FqRec>(), b.out.as<atr_match>(), n, round,
                                                             round + 1 < o->times ? 1 : 0, (const signed char*)(d_stats + L.o_flags[f]),
                                                             o->max_len, o->max_errors, q.fwin.as<uint16_t>(), b.win.as<uint16_t>(),
                                                             q.flags.as<unsigned char>(), (unsigned long long*)(d_stats + L.o_front[f]),
                                                             (unsigned long long*)(d_stats + L.o_hist[f]),
                                                             (unsigned long long*)(d_stats + L.o_adj[f]), (FqCounters*)(d_stats + L.o_side[f]), 0);
                LAUNCHED(ctx);
            }
        }
        k_pe_post<<<grid_for(n, 256), 256, 0, st>>>(s.fq[0].text.as<unsigned char>(), s.fq[0].recs.as<FqRec>(),
                                                    s.fq[1].text.as<unsigned char>(), s.fq[1].recs.as<FqRec>(), n, o->ops,
                                                    s.fq[0].fwin.as<uint16_t>(), s.fq[1].fwin.as<uint16_t>(),
                                                    s.fq[0].flags.as<unsigned char>(), s.fq[1].flags.as<unsigned char>(), d_ops, d_pflags);
        LAUNCHED(ctx);
    }
    if (mg) {
        // MergeOverlapping, the last modifier (commands/trim/__init__.py:546-552): what trimming left of the two reads goes
        // to the merge kernels as two contiguous ASCII batches; then the decision, MergedReadFilter and the waiting filters
        FqMergeCounters* d_mc = (FqMergeCounters*)(d_stats + L.o_merge);
        int* d_max = qm.nl.as<int>();
        int* h_max = &qm.hinfo->nl_overflow;                 // two consecutive ints of the mapped FqInfo (unused for this side)
        int* d_hist = d_max + 2;
        int* d_cursor = d_hist + FQ_MERGE_BINS;
        unsigned char* d_keys = qm.tiles.as<unsigned char>() + n + 16;
        CU(cudaMemsetAsync(d_max, 0, (2 + 2 * FQ_MERGE_BINS) * sizeof(int), st));
        for (int f = 0; f < 2; f++) CU(cudaMemsetAsync(s.fq[f].len64.p, 0, (size_t)(n + 2) * sizeof(long long), st));
        k_pe_merge_len<<<grid_for(n, 256), 256, 0, st>>>(s.fq[0].fwin.as<uint16_t>(), s.fq[1].fwin.as<uint16_t>(), d_pflags, n,
                                                         s.fq[0].len64.as<long long>(), s.fq[1].len64.as<long long>(),
                                                         qm.tiles.as<unsigned char>(), d_max, mg->tb.minov, d_keys, d_hist);
        LAUNCHED(ctx);
        k_pe_merge_publish<<<1, 32, 0, st>>>(d_max, h_max);
        LAUNCHED(ctx);
        k_pe_merge_bins<<<1, 32, 0, st>>>(d_hist, d_cursor);
        LAUNCHED(ctx);
        k_pe_merge_order<<<grid_for(n, 256), 256, 0, st>>>(d_keys, n, d_cursor, qm.tile_offs.as<uint32_t>());
        LAUNCHED(ctx);
        for (int f = 0; f < 2; f++) {
            PeSide b = pe_side(s, f);
            rc = fq_scan_i64(ctx, st, s.scan_tmp, s.fq[f].len64.as<long long>(), (long long*)b.offsets.p, n + 1);
            if (rc) return rc;
            k_pe_merge_gather<<<grid_for(n * 32, 256), 256, 0, st>>>(s.fq[f].text.as<unsigned char>(), s.fq[f].recs.as<FqRec>(),
                                                                    s.fq[f].fwin.as<uint16_t>(), (const long long*)b.offsets.p, n,
                                                                    b.ascii.as<unsigned char>());
            LAUNCHED(ctx);
        }
        CU(cudaStreamSynchronize(st));                       // the longest windows decide which merge kernel runs
        const int max1 = h_max[0], max2 = h_max[1];
        if (max1 > ATR_MERGE_MAX_READ || max2 > ATR_MERGE_MAX_READ) return fail(ctx, ATR_E_LIMIT, "read longer than 4000 nt (merge)");
        const MergePlan plan = merge_plan(max1, max2, mg->error_rate);
        rc = merge_launch(ctx, s, st, plan, max2, b0.ascii.as<unsigned char>(), b0.offsets.as<int64_t>(), 0, b1.ascii.as<unsigned char>(),
                          b1.offsets.as<int64_t>(), 0, qm.tiles.as<unsigned char>(), n, mg->tb,
                          (plan.use_warp && max2 <= 160) ? qm.tile_offs.as<uint32_t>() : nullptr, qm.text.as<atr_merge_result>());
        if (rc) return rc;
        CU(cudaMemsetAsync(qm.len64.p, 0, (size_t)(n + 2) * sizeof(long long), st));
        CU(cudaMemsetAsync(qm.info.p, 0, sizeof(FqInfo), st));
        k_pe_merge_apply<<<grid_for(n, 256), 256, 0, st>>>(s.fq[0].text.as<unsigned char>(), s.fq[0].recs.as<FqRec>(),
                                                           s.fq[1].text.as<unsigned char>(), s.fq[1].recs.as<FqRec>(),
                                                           qm.text.as<atr_merge_result>(), d_pflags, n, o->ops, o->mismatch_action,
                                                           mg->tb.comp, mg->write_merged, s.fq[0].fwin.as<uint16_t>(),
                                                           s.fq[1].fwin.as<uint16_t>(), qm.recs.as<FqMergeRec>(), qm.len64.as<long long>(),
                                                           d_ops, d_mc);
        LAUNCHED(ctx);
        if (mg->write_merged) {
            rc = fq_scan_i64(ctx, st, s.scan_tmp, qm.len64.as<long long>(), qm.outoff.as<long long>(), n + 1);
            if (rc) return rc;
            k_pe_merge_format<<<grid_for(n * 32, 256), 256, 0, st>>>(s.fq[0].text.as<unsigned char>(), s.fq[0].recs.as<FqRec>(),
                                                                    s.fq[1].text.as<unsigned char>(), s.fq[1].recs.as<FqRec>(),
                                                                    b1.ascii.as<unsigned char>(), (const long long*)b1.offsets.p,
                                                                    qm.recs.as<FqMergeRec>(), qm.outoff.as<long long>(), mg->tb.comp, n,
                                                                    qm.outtext.as<unsigned char>(), qm.info.as<FqInfo>());
            LAUNCHED(ctx);
        }
        k_fq_publish<<<1, 32, 0, st>>>(qm.info.as<FqInfo>(), qm.hinfo);
        LAUNCHED(ctx);
    }
    for (int f = 0; f < 2; f++) {
        FqSide& q = s.fq[f];
        CU(cudaMemsetAsync(q.len64.p, 0, (size_t)(n + 2) * sizeof(long long), st));
        k_fq_outlen<<<grid_for(n, 256), 256, 0, st>>>(q.recs.as<FqRec>(), q.fwin.as<uint16_t>(), n, q.len64.as<long long>(), &d_ctr->bp_out[f]);
        LAUNCHED(ctx);
        rc = fq_scan_i64(ctx, st, s.scan_tmp, q.len64.as<long long>(), q.outoff.as<long long>(), n + 1);
        if (rc) return rc;
        k_fq_format<<<grid_for(n, FQF_RECS), 256, 0, st>>>(q.text.as<unsigned char>(), q.recs.as<FqRec>(), q.fwin.as<uint16_t>(),
                                                           q.outoff.as<long long>(), n, q.outtext.as<unsigned char>(), q.info.as<FqInfo>());
        LAUNCHED(ctx);
        k_fq_publish<<<1, 32, 0, st>>>(q.info.as<FqInfo>(), q.hinfo);
        LAUNCHED(ctx);
    }
    return ATR_OK;
}

// frame `n_full` complete records (0 or 1) or the partial record of side f starting at record index `first`: the key
// of the first error, ~0 if none. Synchronous (only used at the very end of the inputs).
int pe_probe(atr_ctx* ctx, Slot& s, int f, const FqChunk& c, int64_t first, int lines_left, unsigned long long* key) {
    FqSide& q = s.fq[f];
    cudaStream_t st = s.stream;
    int rc = q.recs.ensure((size_t)(first + 2) * sizeof(FqRec));
    if (!rc) rc = q.len64.ensure((size_t)(first + 3) * sizeof(long long));
    if (rc) return fail(ctx, rc, "out of device memory");
    CU(cudaMemsetAsync(q.info.p, 0xFF, sizeof(unsigned long long), st));
    // records [0, first) are framed again (cheap: `first` is 0 or 1 here); the thread of record `first` is the probe
    k_fq_frame<<<grid_for(first + 1, 256), 256, 0, st>>>(q.text.as<unsigned char>(), q.nl.as<uint32_t>(), c.n_nl, c.len, first, lines_left,
                                                         q.recs.as<FqRec>(), q.len64.as<long long>(), q.info.as<FqInfo>(), 4, 0);
    LAUNCHED(ctx);
    k_fq_publish<<<1, 32, 0, st>>>(q.info.as<FqInfo>(), q.hinfo);
    LAUNCHED(ctx);
    CU(cudaStreamSynchronize(st));
    *key = q.hinfo->err_key;
    return ATR_OK;
}

}  // namespace

namespace {
int trim_fastq_pe_impl(atr_ctx* ctx, const atr_insertset* iset, const atr_adapterset* set1, const atr_adapterset* set2,
                       const atr_trim_pe_opts* opts, const atr_merge_opts* mopts, const uint8_t* text1, int64_t nbytes1,
                       const uint8_t* text2, int64_t nbytes2, uint8_t* out1, int64_t out_cap1, uint8_t* out2, int64_t out_cap2,
                       uint8_t* out_merged, int64_t out_cap_merged, int64_t* out_bytes /* [3] */, int64_t* consumed,
                       atr_trim_pe_stats* stats, atr_merge_stats* mstats, atr_fastq_error* err) {
    if (!ctx || (iset && (!set1 || !set2)) || !opts || nbytes1 < 0 || nbytes2 < 0 || (nbytes1 > 0 && (!text1 || !out1)) ||
        (nbytes2 > 0 && (!text2 || !out2)) || !out_bytes || !consumed || !stats || !err || (mopts && !mstats))
        return fail(ctx, ATR_E_ARG, "bad arguments to atr_trim_fastq_pe_host");
    if (mopts && (!(mopts->min_overlap > 0) || !(mopts->error_rate >= 0) || mopts->error_rate > 1 || out_cap_merged < 0))
        return fail(ctx, ATR_E_ARG, "bad atr_merge_opts");
    if ((set1 && set1->ctx != ctx) || (set2 && set2->ctx != ctx) || (iset && iset->ctx != ctx))
        return fail(ctx, ATR_E_ARG, "adapter / insert set belongs to another context");
    if (!iset && opts->times < 1) return fail(ctx, ATR_E_ARG, "atr_trim_pe_opts.times must be >= 1 in adapter mode");
    if (opts->mismatch_action < 0 || opts->mismatch_action > 3 || (opts->mismatch_action && !iset && !mopts))
        return fail(ctx, ATR_E_ARG, "mismatch_action is 0..3 and needs the insert aligner or MergeOverlapping");
    if (opts->max_len < 0 || opts->max_len > ATR_MAX_READ || opts->max_errors < 0 || opts->max_errors > 4095 || opts->min_insert_overlap < 0)
        return fail(ctx, ATR_E_ARG, "bad atr_trim_pe_opts");
    const atr_adapterset* sets[2] = {set1, set2};
    for (int f = 0; f < 2; f++) {
        if (iset && (sets[f]->host.size() != 1 || !sets[f]->host[0].desc.match_to_semantics || sets[f]->host[0].desc.flags != 14))
            return fail(ctx, ATR_E_ARG, "insert mode takes one 3' (BACK) adapter per read, created with match_to_semantics = 1");
        if (sets[f]) for (const atr::HostAdapter& h : sets[f]->host)
            if (!h.desc.match_to_semantics) return fail(ctx, ATR_E_ARG, "atr_trim_fastq_pe_host needs adapters created with match_to_semantics = 1");
    }
    CU(cudaSetDevice(ctx->device));
    memset(err, 0, sizeof(*err));
    out_bytes[0] = out_bytes[1] = out_bytes[2] = 0;
    consumed[0] = consumed[1] = 0;
    const uint8_t* text[2] = {text1, text2};
    const int64_t nbytes[2] = {nbytes1, nbytes2};
    uint8_t* outp[2] = {out1, out2};
    const int64_t out_cap[2] = {out_cap1, out_cap2};
    int64_t chunk = opts->chunk_bytes > 0 ? opts->chunk_bytes : ((int64_t)32 << 20);
    chunk = std::max<int64_t>(4096, std::min<int64_t>(chunk, (int64_t)1 << 30));
    const PeLayout L = pe_layout(opts->max_len, opts->max_errors, set1 ? set1->host.size() : 0, set2 ? set2->host.size() : 0);
    int rc = ctx->fq_stats.ensure(L.total);
    if (rc) return fail(ctx, rc, "out of device memory (statistics)");
    char* d_stats = ctx->fq_stats.as<char>();
    CU(cudaMemset(d_stats, 0, L.total));
    PeMerge merge;
    if (mopts) {
        // thresholds and minimum overlaps for every read length the merge kernels take (ctx->misc; nothing else in this call uses it)
        rc = merge_tables(ctx, ctx->slot[0].stream, ATR_MERGE_MAX_READ, mopts->min_overlap, mopts->error_rate, merge.tb);
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->slot[0].stream));
        merge.error_rate = mopts->error_rate;
        merge.write_merged = out_merged != nullptr;
    }
    const PeMerge* mg = mopts ? &merge : nullptr;
    int64_t opos_m = 0;
    for (int f = 0; f < 2; f++) {
        if (!sets[f]) continue;
        std::vector<signed char> ff(sets[f]->host.size());
        for (size_t a = 0; a < ff.size(); a++) {
            const int w = sets[f]->host[a].desc.flags;
            ff[a] = (w == ATR_SEMIGLOBAL) ? (signed char)-1 : ((w == 14 || w == 2) ? (signed char)0 : (signed char)1);
        }
        CU(cudaMemcpy(d_stats + L.o_flags[f], ff.data(), ff.size(), cudaMemcpyHostToDevice));
    }
    const bool final_call = opts->final_chunk != 0;
    int64_t pos[2] = {0, 0}, opos[2] = {0, 0}, records_before = 0;
    auto make_step = [&](int slot) {
        PeStep P;
        P.slot = slot;
        for (int f = 0; f < 2; f++) {
            P.c[f].slot = slot; P.c[f].start = pos[f];
            P.c[f].len = std::min(chunk, nbytes[f] - pos[f]);
            P.c[f].last = (pos[f] + P.c[f].len == nbytes[f]);
        }
        return P;
    };
    auto flags_of = [&](const FqChunk& c, int f, int& final_text, int& unterminated) {
        final_text = (c.last && final_call) ? 1 : 0;
        unterminated = (final_text && c.len > 0 && text[f][c.start + c.len - 1] != '\n') ? 1 : 0;
    };
    auto front = [&](PeStep& P) -> int {
        for (int f = 0; f < 2; f++) {
            int ft, ut;
            flags_of(P.c[f], f, ft, ut);
            int r2 = fq_front(ctx, ctx->slot[P.slot], f, P.c[f], text[f], ft, ut);
            if (r2) return r2;
        }
        return ATR_OK;
    };
    auto line_error = [&](Slot& s, int f, const FqChunk& c, int64_t line, int kind, int64_t rec_base) -> int {
        int r2 = fq_describe(ctx, s.fq[f], c, line, kind, rec_base, err);
        err->file = f;
        return r2 ? r2 : fail(ctx, ATR_E_FORMAT, "malformed FASTQ (see atr_fastq_error)");
    };
    PeStep cur = make_step(0), nxt;
    rc = front(cur);
    if (rc) return rc;
    int result = ATR_OK;
    while (true) {
        Slot& s = ctx->slot[cur.slot];
        CUB(cudaStreamSynchronize(s.stream));
        bool redo = false;
        for (int f = 0; f < 2 && result == ATR_OK; f++) {
            if (s.fq[f].hinfo->nl_overflow) {
                rc = s.fq[f].nl.ensure((size_t)(s.fq[f].hinfo->n_nl + 16) * sizeof(uint32_t));
                if (rc) { result = fail(ctx, rc, "out of device memory (newline index)"); break; }
                int ft, ut;
                flags_of(cur.c[f], f, ft, ut);
                rc = fq_index(ctx, s, f, cur.c[f], ft, ut);
                if (rc) { result = rc; break; }
                redo = true;
            }
        }
        if (result != ATR_OK) break;
        if (redo) CUB(cudaStreamSynchronize(s.stream));
        bool bare = false;
        for (int f = 0; f < 2; f++) {
            const FqInfo hi = *s.fq[f].hinfo;
            int ft, ut;
            flags_of(cur.c[f], f, ft, ut);
            cur.c[f].n_rec = hi.n_rec; cur.c[f].n_nl = hi.n_nl; cur.c[f].consumed = hi.consumed;
            cur.c[f].lines_left = ft ? hi.lines_left : 0;
            if (hi.bare_cr && !bare) {
                bare = true;
                err->kind = ATR_FQ_BARE_CR; err->file = f; err->record = -1;
            }
        }
        if (bare) { result = fail(ctx, ATR_E_FORMAT, "FASTQ text holds a carriage return that is not followed by a newline"); break; }
        const int64_t n = std::min(cur.c[0].n_rec, cur.c[1].n_rec);
        cur.n = n;
        if (n > (int64_t)0x7ffffff0) { result = fail(ctx, ATR_E_LIMIT, "too many records in one chunk"); break; }
        if (n == 0) {
            // No complete pair in this step. Either a record does not fit a chunk, or an input is exhausted.
            bool too_long = false;
            for (int f = 0; f < 2; f++) if (cur.c[f].n_rec == 0 && !cur.c[f].last) too_long = true;
            if (too_long) {
                err->kind = ATR_FQ_TOO_LONG; err->record = records_before;
                result = fail(ctx, ATR_E_FORMAT, "one FASTQ record is larger than the chunk size");
                break;
            }
            if (!final_call) break;                  // the caller supplies more text later
            // end of the files, in the reader's order (io/seqio.py:431-447): next(it1), then next(it2)
            unsigned long long key = ~0ull;
            const bool more1 = cur.c[0].n_rec > 0, more2 = cur.c[1].n_rec > 0;
            if (more1) {                             // read 1 exists: it must be well-formed; then file 2 ends (or fails)
                rc = pe_probe(ctx, s, 0, cur.c[0], 1, 0, &key);
                if (rc) { result = rc; break; }
                if (key != ~0ull) { result = line_error(s, 0, cur.c[0], (int64_t)(key >> 8), (int)(key & 0xFF), records_before); break; }
                if (cur.c[1].lines_left) {
                    rc = pe_probe(ctx, s, 1, cur.c[1], 0, cur.c[1].lines_left, &key);
                    if (rc) { result = rc; break; }
                    result = line_error(s, 1, cur.c[1], (int64_t)(key >> 8), (int)(key & 0xFF), records_before);
                    break;
                }
                err->kind = ATR_FQ_MORE_IN_1; err->file = 0; err->record = records_before;
                result = fail(ctx, ATR_E_FORMAT, "Reads are improperly paired. There are more reads in file 1 than in file 2.");
                break;
            }
            if (cur.c[0].lines_left) {
                rc = pe_probe(ctx, s, 0, cur.c[0], 0, cur.c[0].lines_left, &key);
                if (rc) { result = rc; break; }
                result = line_error(s, 0, cur.c[0], (int64_t)(key >> 8), (int)(key & 0xFF), records_before);
                break;
            }
            if (more2) {
                rc = pe_probe(ctx, s, 1, cur.c[1], 1, 0, &key);
                if (rc) { result = rc; break; }
                if (key != ~0ull) { result = line_error(s, 1, cur.c[1], (int64_t)(key >> 8), (int)(key & 0xFF), records_before); break; }
                err->kind = ATR_FQ_MORE_IN_2; err->file = 1; err->record = records_before;
                result = fail(ctx, ATR_E_FORMAT, "Reads are improperly paired. There are more reads in file 2 than in file 1.");
                break;
            }
            if (cur.c[1].lines_left) {
                rc = pe_probe(ctx, s, 1, cur.c[1], 0, cur.c[1].lines_left, &key);
                if (rc) { result = rc; break; }
                result = line_error(s, 1, cur.c[1], (int64_t)(key >> 8), (int)(key & 0xFF), records_before);
                break;
            }
            pos[0] = nbytes[0]; pos[1] = nbytes[1];  // both inputs end cleanly
            break;
        }
        // where does each side resume? its own complete-record count may exceed n
        bool need_sync = false;
        for (int f = 0; f < 2; f++) {
            if (cur.c[f].n_rec != n) {
                k_fq_consumed<<<1, 32, 0, s.stream>>>(s.fq[f].nl.as<uint32_t>(), n, s.fq[f].hinfo);
                LAUNCHED(ctx);
                need_sync = true;
            }
        }
        if (need_sync) {
            CUB(cudaStreamSynchronize(s.stream));
            for (int f = 0; f < 2; f++) cur.c[f].consumed = s.fq[f].hinfo->consumed;
        }
        for (int f = 0; f < 2; f++) pos[f] = cur.c[f].start + cur.c[f].consumed;
        // the next step's texts start crossing PCIe now
        nxt = make_step(cur.slot ^ 1);
        rc = front(nxt);
        if (rc) { result = rc; break; }
        rc = pe_back(ctx, s, cur, iset, sets, opts, L, d_stats, mg);
        if (rc) { result = rc; break; }
        CUB(cudaStreamSynchronize(s.stream));
        const unsigned long long key = s.fq[0].hinfo->err_key;
        if (key != ~0ull) {
            const int64_t code = (int64_t)(key >> 8), r = code / 16, sub = code % 16;
            const int kind = (int)(key & 0xFF);
            if (sub < 8) {
                const int f = sub < 4 ? 0 : 1;
                result = line_error(s, f, cur.c[f], 4 * r + (sub & 3), kind, records_before);
            } else {                                 // names: both header lines
                FqRec R[2];
                for (int f = 0; f < 2; f++) CUB(cudaMemcpy(&R[f], s.fq[f].recs.as<FqRec>() + r, sizeof(FqRec), cudaMemcpyDeviceToHost));
                if (result) break;                   // (the CUB above only left the inner loop)
                err->kind = kind; err->file = 0; err->record = records_before + r; err->terminated = 1;
                err->line_begin = cur.c[0].start + R[0].hdr_b; err->line_end = err->line_begin + R[0].hdr_len;
                err->line_begin2 = cur.c[1].start + R[1].hdr_b; err->line_end2 = err->line_begin2 + R[1].hdr_len;
                result = fail(ctx, ATR_E_FORMAT, "Reads are improperly paired (read names differ)");
            }
            break;
        }
        bool cap_bad = false;
        for (int f = 0; f < 2; f++) {
            FqSide& q = s.fq[f];
            const int64_t ob = (int64_t)q.hinfo->out_bytes;
            if (opos[f] + ob > out_cap[f]) { cap_bad = true; break; }
            if (ob) {
                CUB(cudaMemcpyAsync(outp[f] + opos[f], q.outtext.p, (size_t)ob, cudaMemcpyDeviceToHost, q.out_stream));
                CUB(cudaEventRecord(q.ev_d2h, q.out_stream));
                q.d2h_pending = 1;
            }
            opos[f] += ob;
        }
        if (result) break;                           // (a CUB inside the loop over the two sides)
        if (mg && mg->write_merged && !cap_bad) {
            FqSide& q = s.fqm;
            const int64_t ob = (int64_t)q.hinfo->out_bytes;
            if (opos_m + ob > out_cap_merged) cap_bad = true;
            else {
                if (ob) {
                    CUB(cudaMemcpyAsync(out_merged + opos_m, q.outtext.p, (size_t)ob, cudaMemcpyDeviceToHost, q.out_stream));
                    CUB(cudaEventRecord(q.ev_d2h, q.out_stream));
                    q.d2h_pending = 1;
                }
                opos_m += ob;
            }
            if (result) break;
        }
        if (cap_bad) { result = fail(ctx, ATR_E_ARG, "out_cap too small for the trimmed text (nbytes + 1 always suffices)"); break; }
        records_before += n;
        cur = nxt;
    }
    for (int k = 0; k < 2; k++) {
        CU(cudaStreamSynchronize(ctx->slot[k].stream));
        for (int f = 0; f < 2; f++) {
            if (ctx->slot[k].fq[f].out_stream) CU(cudaStreamSynchronize(ctx->slot[k].fq[f].out_stream));
            ctx->slot[k].fq[f].d2h_pending = 0;
        }
        if (ctx->slot[k].fqm.out_stream) CU(cudaStreamSynchronize(ctx->slot[k].fqm.out_stream));
        ctx->slot[k].fqm.d2h_pending = 0;
    }
    if (result != ATR_OK) return result;
    std::vector<char> hst(L.total);
    CU(cudaMemcpy(hst.data(), d_stats, L.total, cudaMemcpyDeviceToHost));
    const FqPeCounters* hc = (const FqPeCounters*)(hst.data() + L.o_ctr);
    if (hc->invalid) {
        err->kind = ATR_FQ_INVALID_MATCH; err->record = -1;
        return fail(ctx, ATR_E_FORMAT, "a pair for which the reference raises (Match with length <= errors, or a byte reverse_complement rejects)");
    }
    if (hc->correction_errors) {
        err->kind = ATR_FQ_CORRECTION; err->record = -1;
        return fail(ctx, ATR_E_FORMAT, "error correction would raise in the reference (reads of unequal length or bytes outside the complement table)");
    }
    stats->records_corrected += (int64_t)hc->records_corrected;
    stats->bp_corrected[0] += (int64_t)hc->bp_corrected[0];
    stats->bp_corrected[1] += (int64_t)hc->bp_corrected[1];
    stats->records += (int64_t)hc->records;
    stats->insert_matches += (int64_t)hc->insert_matches;
    stats->overflow += (int64_t)hc->overflow;
    for (int f = 0; f < 2; f++) {
        const FqCounters* sc = (const FqCounters*)(hst.data() + L.o_side[f]);     // adapter mode: the read's AdapterCutter
        if (sc->invalid) {
            err->kind = ATR_FQ_INVALID_MATCH; err->record = -1;
            return fail(ctx, ATR_E_FORMAT, "an alignment of length <= errors: Match.__init__ raises ValueError in the reference");
        }
        stats->with_adapters[f] += (int64_t)sc->with_adapters;
        stats->overflow += (int64_t)sc->overflow;
        const unsigned long long* hfr = (const unsigned long long*)(hst.data() + L.o_front[f]);
        if (stats->errors_front[f]) for (size_t i = 0; i < L.nA[f] * L.H; i++) stats->errors_front[f][i] += (int64_t)hfr[i];
        stats->with_adapters[f] += (int64_t)hc->with_adapters[f];
        stats->bp_in[f] += (int64_t)hc->bp_in[f];
        stats->bp_out[f] += (int64_t)hc->bp_out[f];
        const unsigned long long* hh = (const unsigned long long*)(hst.data() + L.o_hist[f]);
        const unsigned long long* ha = (const unsigned long long*)(hst.data() + L.o_adj[f]);
        if (stats->errors_back[f]) for (size_t i = 0; i < L.nA[f] * L.H; i++) stats->errors_back[f][i] += (int64_t)hh[i];
        if (stats->adjacent_bases[f]) for (size_t i = 0; i < L.nA[f] * 5; i++) stats->adjacent_bases[f][i] += (int64_t)ha[i];
    }
    fq_add_ops(stats->ops, *(const FqOpsCounters*)(hst.data() + L.o_ops));
    if (mg) {
        const FqMergeCounters* mc = (const FqMergeCounters*)(hst.data() + L.o_merge);
        if (mc->raises) {
            err->kind = ATR_FQ_INVALID_MATCH; err->record = -1;
            return fail(ctx, ATR_E_FORMAT, "a pair for which MergeOverlapping raises in the reference (a byte reverse_complement rejects, or an invalid alignment)");
        }
        if (mc->correction_errors) {
            err->kind = ATR_FQ_CORRECTION; err->record = -1;
            return fail(ctx, ATR_E_FORMAT, "error correction would raise in the reference (bytes outside the complement table)");
        }
        mstats->merged += (int64_t)mc->merged;
        mstats->merged_written += (int64_t)mc->merged_written;
        mstats->bp_merged_written += (int64_t)mc->bp_merged;
        mstats->records_corrected += (int64_t)mc->records_corrected;
        mstats->bp_corrected[0] += (int64_t)mc->bp_corrected[0];
        mstats->bp_corrected[1] += (int64_t)mc->bp_corrected[1];
    }
    out_bytes[0] = opos[0]; out_bytes[1] = opos[1]; out_bytes[2] = opos_m;
    consumed[0] = pos[0]; consumed[1] = pos[1];
    ctx->last_ms = -1.f;
    return ATR_OK;
}
}  // namespace

extern "C" int atr_trim_fastq_pe_host(atr_ctx* ctx, const atr_insertset* iset, const atr_adapterset* set1, const atr_adapterset* set2,
                                      const atr_trim_pe_opts* opts, const uint8_t* text1, int64_t nbytes1, const uint8_t* text2,
                                      int64_t nbytes2, uint8_t* out1, int64_t out_cap1, uint8_t* out2, int64_t out_cap2,
                                      int64_t* out_bytes, int64_t* consumed, atr_trim_pe_stats* stats, atr_fastq_error* err) {
    int64_t ob[3] = {0, 0, 0};
    const int rc = trim_fastq_pe_impl(ctx, iset, set1, set2, opts, nullptr, text1, nbytes1, text2, nbytes2, out1, out_cap1, out2, out_cap2,
                                      nullptr, 0, out_bytes ? ob : nullptr, consumed, stats, nullptr, err);
    if (out_bytes) { out_bytes[0] = ob[0]; out_bytes[1] = ob[1]; }
    return rc;
}

extern "C" int atr_trim_fastq_pe_merge_host(atr_ctx* ctx, const atr_insertset* iset, const atr_adapterset* set1, const atr_adapterset* set2,
                                            const atr_trim_pe_opts* opts, const atr_merge_opts* mopts, const uint8_t* text1,
                                            int64_t nbytes1, const uint8_t* text2, int64_t nbytes2, uint8_t* out1, int64_t out_cap1,
                                            uint8_t* out2, int64_t out_cap2, uint8_t* out_merged, int64_t out_cap_merged,
                                            int64_t* out_bytes, int64_t* consumed, atr_trim_pe_stats* stats, atr_merge_stats* mstats,
                                            atr_fastq_error* err) {
    return trim_fastq_pe_impl(ctx, iset, set1, set2, opts, mopts, text1, nbytes1, text2, nbytes2, out1, out_cap1, out2, out_cap2,
                              out_merged, out_cap_merged, out_bytes, consumed, stats, mstats, err);
}
