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
    size_t o_ctr, o_front, o_back, o_adj, o_flags, total;
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
    L.total = L.o_flags + ((n_adapters + 15) & ~(size_t)15);
    return L;
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

// newline index of the chunk (re-runnable: a too small index buffer is grown and the pass repeated)
int fq_index(atr_ctx* ctx, Slot& s, const FqChunk& c, int final_text, int unterminated) {
    cudaStream_t st = s.stream;
    FqInfo* d_info = s.fq_info.as<FqInfo>();
    CU(cudaMemsetAsync(d_info, 0, sizeof(FqInfo), st));
    CU(cudaMemsetAsync(d_info, 0xFF, sizeof(unsigned long long), st));       // err_key = "none"
    CU(cudaMemsetAsync(s.fq_tiles.p, 0, (size_t)(c.n_tiles + 1) * sizeof(unsigned), st));
    if (c.n_tiles > 0) {
        k_fq_nl_count<<<(unsigned)c.n_tiles, FQ_THREADS, 0, st>>>(s.fq_text.as<unsigned char>(), c.len, final_text,
                                                                  s.fq_tiles.as<unsigned>(), d_info);
        LAUNCHED(ctx);
    }
    int rc = fq_scan_u32(ctx, st, s.scan_tmp, s.fq_tiles.as<unsigned>(), s.fq_tile_offs.as<unsigned>(), c.n_tiles + 1);
    if (rc) return rc;
    const long long nl_cap = (long long)(s.fq_nl.cap / sizeof(uint32_t));
    if (c.n_tiles > 0) {
        k_fq_nl_fill<<<(unsigned)c.n_tiles, FQ_THREADS, 0, st>>>(s.fq_text.as<unsigned char>(), c.len, s.fq_tile_offs.as<unsigned>(),
                                                                 s.fq_nl.as<uint32_t>(), nl_cap, d_info);
        LAUNCHED(ctx);
    }
    k_fq_info<<<1, 32, 0, st>>>(s.fq_tile_offs.as<unsigned>(), c.n_tiles, s.fq_nl.as<uint32_t>(), nl_cap, c.len, unterminated, d_info,
                                s.fq_hinfo);
    LAUNCHED(ctx);
    return ATR_OK;
}

int fq_front(atr_ctx* ctx, Slot& s, FqChunk& c, const uint8_t* text, int final_text, int unterminated) {
    c.n_tiles = (int)((c.len + FQ_TILE - 1) / FQ_TILE);
    int rc = s.fq_text.ensure((size_t)c.len + 64);
    if (!rc) rc = s.fq_tiles.ensure((size_t)(c.n_tiles + 2) * sizeof(unsigned));
    if (!rc) rc = s.fq_tile_offs.ensure((size_t)(c.n_tiles + 2) * sizeof(unsigned));
    if (!rc) rc = s.fq_nl.ensure((size_t)(c.len / 8 + 1024) * sizeof(uint32_t));      // >= 8 bytes per line on average; grown on demand
    if (!rc) rc = s.fq_info.ensure(sizeof(FqInfo));
    if (rc) return fail(ctx, rc, "out of device memory (FASTQ chunk)");
    if (!s.fq_hinfo) {                   // mapped pinned (UVA: the same pointer is valid in kernels)
        CU(cudaHostAlloc((void**)&s.fq_hinfo, sizeof(FqInfo), cudaHostAllocMapped));
        CU(cudaStreamCreateWithFlags(&s.fq_out_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s.fq_ev_d2h, cudaEventDisableTiming));
    }
    if (c.len) CU(cudaMemcpyAsync(s.fq_text.p, text + c.start, (size_t)c.len, cudaMemcpyHostToDevice, s.stream));
    return fq_index(ctx, s, c, final_text, unterminated);
}

int fq_back(atr_ctx* ctx, Slot& s, const FqChunk& c, const atr_adapterset* set, const atr_trim_opts* o, const FqStatsLayout& L,
            char* d_stats) {
    cudaStream_t st = s.stream;
    const int64_t n = c.n_rec;
    FqInfo* d_info = s.fq_info.as<FqInfo>();
    const unsigned char* d_text = s.fq_text.as<unsigned char>();
    FqCounters* d_ctr = (FqCounters*)(d_stats + L.o_ctr);
    int rc = s.fq_recs.ensure((size_t)(n + 1) * sizeof(FqRec));
    if (!rc) rc = s.fq_len64.ensure((size_t)(n + 2) * sizeof(long long));
    if (!rc) rc = s.offsets.ensure((size_t)(n + 2) * sizeof(int64_t));
    if (!rc) rc = s.fq_outoff.ensure((size_t)(n + 2) * sizeof(long long));
    if (!rc) rc = s.ascii.ensure((size_t)c.len + 64);
    if (!rc) rc = s.codes.ensure((size_t)(c.len / 8 + n + 2) * sizeof(uint32_t));
    if (!rc) rc = s.woff.ensure((size_t)(n + 1) * sizeof(uint32_t));
    if (!rc) rc = s.len.ensure((size_t)(n + 1) * sizeof(uint16_t));
    if (!rc) rc = s.out.ensure((size_t)(n + 1) * sizeof(atr_match));
    if (!rc) rc = s.win.ensure((size_t)(n + 1) * 2 * sizeof(uint16_t));
    if (!rc) rc = s.fq_fwin.ensure((size_t)(n + 1) * 2 * sizeof(uint16_t));
    if (!rc) rc = s.fq_outtext.ensure((size_t)c.len + 64);
    if (rc) return fail(ctx, rc, "out of device memory (FASTQ records)");
    // the formatted text of the chunk that used this slot before may still be on its way to the host
    if (s.fq_d2h_pending) { CU(cudaStreamWaitEvent(st, s.fq_ev_d2h, 0)); s.fq_d2h_pending = 0; }
    // frame + validate (one extra thread for a trailing partial record)
    CU(cudaMemsetAsync(s.fq_len64.p, 0, (size_t)(n + 2) * sizeof(long long), st));
    k_fq_frame<<<grid_for(n + 1, 256), 256, 0, st>>>(d_text, s.fq_nl.as<uint32_t>(), c.n_nl, c.len, n, c.lines_left,
                                                     s.fq_recs.as<FqRec>(), s.fq_len64.as<long long>(), d_info);
    LAUNCHED(ctx);
    if (n > 0) {
        rc = fq_scan_i64(ctx, st, s.scan_tmp, s.fq_len64.as<long long>(), (long long*)s.offsets.p, n + 1);
        if (rc) return rc;
        k_fq_gather<<<grid_for(n * 32, 256), 256, 0, st>>>(d_text, s.fq_recs.as<FqRec>(), (const long long*)s.offsets.p, n,
                                                           s.ascii.as<unsigned char>());
        LAUNCHED(ctx);
        // the reads are upper-cased for matching only (adapters/__init__.py:349): fold_case = 1
        rc = pack_on_stream(ctx, st, s.counts, s.scan_tmp, s.ascii.as<uint8_t>(), s.offsets.as<int64_t>(), 0, n, 1,
                            s.codes.as<uint32_t>(), s.woff.as<uint32_t>(), s.len.as<uint16_t>());
        if (rc) return rc;
        k_fq_init_win<<<grid_for(n, 256), 256, 0, st>>>(s.fq_recs.as<FqRec>(), n, s.fq_fwin.as<uint16_t>(), d_ctr);
        LAUNCHED(ctx);
        for (int round = 0; round < o->times; round++) {
            rc = locate_on_stream(ctx, s, set, s.codes.as<uint32_t>(), s.woff.as<uint32_t>(), s.len.as<uint16_t>(),
                                  round ? s.win.as<uint16_t>() : nullptr, s.ascii.as<uint8_t>(), s.offsets.as<int64_t>(), 0, 1, n,
                                  s.out.as<atr_match>());
            if (rc) return rc;
            k_fq_apply<<<grid_for(n, 256), 256, 0, st>>>(d_text, s.fq_recs.as<FqRec>(), s.out.as<atr_match>(), n, round,
                                                         round + 1 < o->times ? 1 : 0, (const signed char*)(d_stats + L.o_flags),
                                                         o->max_len, o->max_errors, s.fq_fwin.as<uint16_t>(), s.win.as<uint16_t>(),
                                                         (unsigned long long*)(d_stats + L.o_front), (unsigned long long*)(d_stats + L.o_back),
                                                         (unsigned long long*)(d_stats + L.o_adj), d_ctr);
            LAUNCHED(ctx);
        }
        CU(cudaMemsetAsync(s.fq_len64.p, 0, (size_t)(n + 2) * sizeof(long long), st));
        k_fq_outlen<<<grid_for(n, 256), 256, 0, st>>>(s.fq_recs.as<FqRec>(), s.fq_fwin.as<uint16_t>(), n, s.fq_len64.as<long long>(), d_ctr);
        LAUNCHED(ctx);
        rc = fq_scan_i64(ctx, st, s.scan_tmp, s.fq_len64.as<long long>(), s.fq_outoff.as<long long>(), n + 1);
        if (rc) return rc;
        k_fq_format<<<grid_for(n * 32, 256), 256, 0, st>>>(d_text, s.fq_recs.as<FqRec>(), s.fq_fwin.as<uint16_t>(),
                                                           s.fq_outoff.as<long long>(), n, s.fq_outtext.as<unsigned char>(), d_info);
        LAUNCHED(ctx);
    }
    k_fq_publish<<<1, 32, 0, st>>>(d_info, s.fq_hinfo);
    LAUNCHED(ctx);
    return ATR_OK;
}

// describe the first malformed line of chunk c (key = line index << 8 | kind) in coordinates of the call's text
int fq_describe(atr_ctx* ctx, Slot& s, const FqChunk& c, unsigned long long key, int64_t records_before, atr_fastq_error* err) {
    const int64_t line = (int64_t)(key >> 8);
    err->kind = (int32_t)(key & 0xFF);
    err->record = records_before + line / 4;
    err->line_in_record = (int32_t)(line % 4);
    const int64_t total_lines = c.n_nl + ((c.n_rec * 4 + c.lines_left) > c.n_nl ? 1 : 0);
    if (line >= total_lines) {           // "ended prematurely": there is no such line
        err->line_begin = err->line_end = c.start + c.len;
        err->terminated = 0;
        return ATR_OK;
    }
    uint32_t prev = 0, cur = 0;
    if (line > 0) CU(cudaMemcpy(&prev, s.fq_nl.as<uint32_t>() + (line - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
    int64_t b = line > 0 ? (int64_t)prev + 1 : 0, e = c.len;
    err->terminated = 0;
    if (line < c.n_nl) {
        CU(cudaMemcpy(&cur, s.fq_nl.as<uint32_t>() + line, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        e = cur;
        err->terminated = 1;
    }
    err->line_begin = c.start + b;
    err->line_end = c.start + e;         // a "\r" before the "\n" is left to the caller (it has the text)
    return ATR_OK;
}

}  // namespace

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
    CU(cudaSetDevice(ctx->device));
    memset(err, 0, sizeof(*err));
    *out_bytes = 0;
    *consumed = 0;
    int64_t chunk = opts->chunk_bytes > 0 ? opts->chunk_bytes : ((int64_t)64 << 20);
    chunk = std::max<int64_t>(4096, std::min<int64_t>(chunk, (int64_t)1 << 30));
    const size_t nA = set->host.size();
    const FqStatsLayout L = fq_layout(nA, opts->max_len, opts->max_errors);
    int rc = ctx->fq_stats.ensure(L.total);
    if (rc) return fail(ctx, rc, "out of device memory (statistics)");
    char* d_stats = ctx->fq_stats.as<char>();
    CU(cudaMemset(d_stats, 0, L.total));
    {
        std::vector<signed char> ff(nA);
        for (size_t a = 0; a < nA; a++) {                // Adapter.__init__: adapters/__init__.py:301-304
            const int w = set->host[a].desc.flags;
            ff[a] = (w == ATR_SEMIGLOBAL) ? (signed char)-1 : ((w == 14 || w == 2) ? (signed char)0 : (signed char)1);
        }
        CU(cudaMemcpy(d_stats + L.o_flags, ff.data(), nA, cudaMemcpyHostToDevice));
    }
    int64_t opos = 0, records_before = 0, done = 0;
    // measurement knobs (never set in production): skip the D2H of the text / print per-phase device times
    static const bool dbg_no_d2h = getenv("ATR_FQ_NO_D2H") != nullptr;
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
    rc = fq_front(ctx, ctx->slot[0], cur, text, ft, ut);
    if (rc) return rc;
    int result = ATR_OK;
    while (true) {
        Slot& s = ctx->slot[cur.slot];
        const double tw0 = now();
        CU(cudaStreamSynchronize(s.stream));
        t_wait_front += now() - tw0;
        if (s.fq_hinfo->nl_overflow) {                   // more lines than the index was sized for: grow, redo the index
            rc = s.fq_nl.ensure((size_t)(s.fq_hinfo->n_nl + 16) * sizeof(uint32_t));
            if (rc) return fail(ctx, rc, "out of device memory (newline index)");
            flags_of(cur, ft, ut);
            rc = fq_index(ctx, s, cur, ft, ut);
            if (rc) return rc;
            CU(cudaStreamSynchronize(s.stream));
        }
        const FqInfo hi = *s.fq_hinfo;
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
            rc = fq_front(ctx, ctx->slot[nxt.slot], nxt, text, ft2, ut2);
            if (rc) { result = rc; break; }
        }
        if (dbg_timing) cudaEventRecord(evs[cur.slot][2], s.stream);
        rc = fq_back(ctx, s, cur, set, opts, L, d_stats);
        if (rc) { result = rc; break; }
        if (dbg_timing) cudaEventRecord(evs[cur.slot][3], s.stream);
        const double tw1 = now();
        CU(cudaStreamSynchronize(s.stream));
        t_wait_back += now() - tw1;
        if (dbg_timing) { float ms = 0; cudaEventElapsedTime(&ms, evs[cur.slot][2], evs[cur.slot][3]); t_back += ms; }
        const FqInfo hb = *s.fq_hinfo;
        if (hb.err_key != ~0ull) {
            rc = fq_describe(ctx, s, cur, hb.err_key, records_before, err);
            result = rc ? rc : fail(ctx, ATR_E_FORMAT, "malformed FASTQ (see atr_fastq_error)");
            break;
        }
        if (opos + (int64_t)hb.out_bytes > out_cap) { result = fail(ctx, ATR_E_ARG, "out_cap too small for the trimmed text"); break; }
        // own stream: the next chunk's H2D into this slot must not queue behind this copy
        if (hb.out_bytes && !dbg_no_d2h) {
            CU(cudaMemcpyAsync(out_text + opos, s.fq_outtext.p, (size_t)hb.out_bytes, cudaMemcpyDeviceToHost, s.fq_out_stream));
            CU(cudaEventRecord(s.fq_ev_d2h, s.fq_out_stream));
            s.fq_d2h_pending = 1;
        }
        opos += (int64_t)hb.out_bytes;
        records_before += cur.n_rec;
        done = cur.start + ((cur.last && final_call) ? cur.len : cur.consumed);
        if (cur.last) break;
        cur = nxt;
    }
    for (int k = 0; k < 2; k++) {
        CU(cudaStreamSynchronize(ctx->slot[k].stream));
        if (ctx->slot[k].fq_out_stream) CU(cudaStreamSynchronize(ctx->slot[k].fq_out_stream));
        ctx->slot[k].fq_d2h_pending = 0;
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
    *out_bytes = opos;
    *consumed = done;
    ctx->last_ms = -1.f;
    return ATR_OK;
}
