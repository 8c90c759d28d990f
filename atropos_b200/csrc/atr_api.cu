// atr_api.cu -- the C ABI declared in include/atropos_b200.h: contexts, adapter/insert sets,
// chunked double-buffered host entry points and the kernel launch logic.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <nvtx3/nvToolsExt.h>        // header-only; ranges cost nothing unless a tool (nsys / ncu --nvtx) is attached

#include <algorithm>
#include <cfenv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "atr_kernels.cuh"

static thread_local std::string g_last_error;

// NVTX range around a host-side chunk / entry point (shows the double-buffered pipeline in a timeline)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return ATR_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); p = nullptr; return ATR_E_NOMEM; }
        cap = want;
        return ATR_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

}  // namespace
struct FqInfo;                           // fastq_kernels.cuh
namespace {

// one FASTQ text stream of a slot: chunk text, newline index, record table, windows, formatted text
struct FqSide {
    DevBuf text, tiles, tile_offs, nl, info, recs, len64, outoff, fwin, outtext, flags;
    FqInfo* hinfo = nullptr;             // mapped pinned
    cudaStream_t out_stream = nullptr;   // D2H of the formatted text
    cudaEvent_t ev_d2h = nullptr;
    int d2h_pending = 0;
    void release() {
        DevBuf* all[] = {&text, &tiles, &tile_offs, &nl, &info, &recs, &len64, &outoff, &fwin, &outtext, &flags};
        for (DevBuf* b : all) b->release();
        if (hinfo) { cudaFreeHost(hinfo); hinfo = nullptr; }
        if (out_stream) { cudaStreamSynchronize(out_stream); cudaStreamDestroy(out_stream); out_stream = nullptr; }
        if (ev_d2h) { cudaEventDestroy(ev_d2h); ev_d2h = nullptr; }
    }
};

// per-stream working set of the host entry points
struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;          // k_wide runs here, concurrently with k_band on `stream`
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DevBuf ascii, offsets, win, counts, woff, codes, len, out, scan_tmp, gen_scratch, lists;
    DevBuf ascii2, offsets2, counts2, woff2, codes2, len2;     // second mate (insert aligner)
    // FASTQ paths (atr_fastq_api.cuh): one FqSide per input text (single-end uses fq[0]); second mate's windows,
    // fallback matches and the insert-aligner results of the paired-end path
    FqSide fq[2];
    FqSide fqm;                          // merged reads of the paired-end path (third output: atr_trim_fastq_pe_merge_host)
    DevBuf win2, out2, ins_out;
    void release() {
        DevBuf* all[] = {&ascii, &offsets, &win, &counts, &woff, &codes, &len, &out, &scan_tmp, &gen_scratch, &lists,
                         &ascii2, &offsets2, &counts2, &woff2, &codes2, &len2, &win2, &out2, &ins_out};
        for (DevBuf* b : all) b->release();
        fq[0].release(); fq[1].release(); fqm.release();
    }
};

}  // namespace

struct atr_ctx {
    int device = 0;
    Slot slot[2];                    // slot[0].stream is "the ctx stream" of the *_device entry points
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    AtrTables* d_tables = nullptr;
    AtrTables h_tables;
    std::string err;
    int64_t launches = 0;
    float last_ms = -1.f;
    int profile = 0, phases_valid = 0;
    cudaEvent_t pev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // start, after refine, after band, after wide, after filter
    const char* phase_names[4] = {"", "", "", ""};   // kernels behind atr_ctx_last_phase_ms' four intervals
    int disable_sa = 0;
    int disable_qg = 0;              // ATR_DISABLE_QG=1: Shift-And first stage even where the q-gram form is eligible (A/B measurements)
    int sm_count = 148;
    int64_t panel_chunk = (int64_t)1 << 40;   // ATR_PANEL_CHUNK: reads per piece of a multi-adapter device call. Off by default: pieces of
                                     // 512 Ki reads keep the later passes in L2 but multiply the launches (9.97 vs 5.68 ms per 5 M reads x 8 adapters)
    int qg_ctas = 5;                 // persistent CTAs per SM of k_filter_qg (51 registers, 42 KB shared memory: 5 fit); ATR_QG_CTAS overrides
    int disable_fused = 0;           // ATR_DISABLE_FUSED=1: always use the plain register-DP kernel (A/B measurements)
    DevBuf misc;                     // small single-call scratch (compare_prefixes, multi_locate)
    DevBuf fq_stats;                 // counters + histograms of atr_trim_fastq_host
};

struct atr_adapterset {
    atr_ctx* ctx = nullptr;
    std::vector<atr::HostAdapter> host;
    std::vector<AdapterK1a> k1a;     // valid where host[i].k1a_ok
    std::vector<AdapterGen> gen;
    std::vector<void*> dev_allocs;
    std::vector<char> shadowed;      // adapter i repeats an earlier adapter of the set in every respect: it can never win
                                     // (AdapterCutter._best_match keeps the first on equal matches, modifiers.py:120) and is skipped
    int max_m = 0;
};

struct atr_insertset {
    atr_ctx* ctx = nullptr;
    int kmax = 0;                    // largest error budget of the set: k_by_len[max_len]
    InsertDev dev;
    std::vector<void*> dev_allocs;
};

namespace {

int fail(atr_ctx* ctx, int code, const std::string& msg) {
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}

int cuda_fail(atr_ctx* ctx, cudaError_t e, const char* what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return fail(ctx, e == cudaErrorMemoryAllocation ? ATR_E_NOMEM : ATR_E_CUDA, msg);
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); } while (0)
#define LAUNCHED(ctx) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, "kernel launch"); } while (0)

template <class T>
int upload(atr_ctx* ctx, std::vector<void*>& allocs, const T* src, size_t count, const T** out) {
    void* p = nullptr;
    CU(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)));
    allocs.push_back(p);
    if (count) CU(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return ATR_OK;
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

using atr::thr_mul_of;

// ---- packing on a stream -------------------------------------------------------------------
// d_offsets: n+1 int64 (absolute; `base` is subtracted to index d_ascii). Fills woff (n+1), codes, len.
int pack_on_stream(atr_ctx* ctx, cudaStream_t st, DevBuf& counts, DevBuf& scan_tmp, const uint8_t* d_ascii,
                   const int64_t* d_offsets, int64_t base, int64_t n, int fold_case, uint32_t* d_codes, uint32_t* d_woff,
                   uint16_t* d_len) {
    if (n <= 0) return ATR_OK;
    int rc = counts.ensure((size_t)(n + 1) * sizeof(uint32_t));
    if (rc) return fail(ctx, rc, "out of device memory (pack counts)");
    CU(cudaMemsetAsync(counts.p, 0, (size_t)(n + 1) * sizeof(uint32_t), st));
    k_word_counts<<<grid_for(n, 256), 256, 0, st>>>(d_offsets, n, counts.as<uint32_t>());
    LAUNCHED(ctx);
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.as<uint32_t>(), d_woff, (int)(n + 1), st));
    rc = scan_tmp.ensure(tmp_bytes);
    if (rc) return fail(ctx, rc, "out of device memory (scan)");
    CU(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, counts.as<uint32_t>(), d_woff, (int)(n + 1), st));
    ctx->launches += 2;              // CUB's scan is two kernels (library plumbing, counted for honesty)
    k_pack<<<(unsigned)std::min<int64_t>(grid_for(n, ATR_PK_READS), 148 * 8), 256, 0, st>>>(d_ascii, d_offsets, base, n, fold_case, ctx->d_tables, d_woff, d_codes, d_len);
    LAUNCHED(ctx);
    return ATR_OK;
}

// ---- K1 launch logic -------------------------------------------------------------------------
int locate_on_stream(atr_ctx* ctx, Slot& slot, const atr_adapterset* set,
                     const uint32_t* d_codes, const uint32_t* d_woff, const uint16_t* d_len, const uint16_t* d_win,
                     const uint8_t* d_ascii, const int64_t* d_offsets, int64_t base, int fold_case, int64_t n,
                     atr_match* d_out) {
    if (n <= 0) return ATR_OK;
    cudaStream_t st = slot.stream;
    DevBuf& gen_scratch = slot.gen_scratch;
    DevBuf& lists = slot.lists;
    const bool have_ascii = d_ascii != nullptr && d_offsets != nullptr;
    const bool have_packed = d_codes != nullptr && d_woff != nullptr && d_len != nullptr;
    // general-kernel geometry (grid-stride; scratch is per thread)
    const int gblock = 128;
    int64_t gthreads = std::min<int64_t>((n + gblock - 1) / gblock, 148 * 8) * gblock;
    const size_t per_thread = (size_t)(set->max_m + 1) * sizeof(GCell);
    const size_t budget = (size_t)768 << 20;
    if ((size_t)gthreads * per_thread > budget)
        gthreads = std::max<int64_t>(gblock, (int64_t)(budget / per_thread) / gblock * gblock);
    if (have_ascii) {
        int rc = gen_scratch.ensure((size_t)gthreads * per_thread);
        if (rc) return fail(ctx, rc, "out of device memory (general-kernel scratch)");
    }
    for (size_t a = 0; a < set->host.size(); a++) {
        const atr::HostAdapter& h = set->host[a];
        if (a < set->shadowed.size() && set->shadowed[a]) continue;
        if (h.k1a_ok && have_packed) {
            AdapterK1a p = set->k1a[a];
            p.reduce = a > 0;
            p.mark_routed = !have_ascii;
            if (p.fused_ok && !ctx->disable_fused && n < (int64_t)0x7fffffff) {
                // filter -> survivor lists -> banded / windowed DP over the survivors only
                int rc = lists.ensure((size_t)n * 3 * sizeof(Survivor) + 64);
                if (rc) return fail(ctx, rc, "out of device memory (survivor lists)");
                int* counters = lists.as<int>();
                Survivor* narrow = (Survivor*)(lists.as<char>() + 64);
                Survivor* wide = narrow + n;
                Survivor* refine = wide + n;
                CU(cudaMemsetAsync(counters, 0, 4 * sizeof(int), st));
                p.split8 = 1;                                  // bands of <= 8 diagonals: back of the narrow list, k_band<8>
                const bool prof = ctx->profile && set->host.size() == 1;
                if (prof) CU(cudaEventRecord(ctx->pev[0], st));
                const unsigned g = grid_for(n, ATR_K1F_THREADS);
                const unsigned gp = (unsigned)std::min<int64_t>((n + 127) / 128, 148 * 12);
                const bool use_sa = p.sa_ok && !ctx->disable_sa;
                const bool use_qg = p.qg_ok && !ctx->disable_qg && (p.qg_wide || (use_sa && p.qg_step == 3));   // at step 2 (7-row pieces) the automaton wins: 0.38 vs 0.48 ms per 5 M reads for the 21-mer
                if (prof) {
                    ctx->phase_names[0] = use_qg ? "k_filter_qg" : (use_sa ? "k_filter_sa" : ((p.sa_front && !ctx->disable_sa) ? "k_filter_front" : "k_filter"));
                    ctx->phase_names[1] = (use_sa || use_qg) ? "k_refine" : "";
                    ctx->phase_names[2] = "k_band<16>+k_band<8>";
                    ctx->phase_names[3] = "k_wide";
                }
                if (use_qg) {
                    const unsigned gq = (unsigned)std::min<int64_t>((n + ATR_QG_THREADS - 1) / ATR_QG_THREADS, (int64_t)ctx->sm_count * ctx->qg_ctas);
                    if (p.qg_wide) {
                        if (p.qg_step == 3) k_filter_qg<3, unsigned long long><<<gq, ATR_QG_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                        else k_filter_qg<2, unsigned long long><<<gq, ATR_QG_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                    } else {
                        if (p.qg_step == 3) k_filter_qg<3, unsigned><<<gq, ATR_QG_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                        else k_filter_qg<2, unsigned><<<gq, ATR_QG_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                    }
                } else if (use_sa) {
                    if (h.and_mode) k_filter_sa<true><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                    else k_filter_sa<false><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                } else if (p.sa_front && !ctx->disable_sa) {
                    if (h.and_mode) k_filter_front<true><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                    else k_filter_front<false><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                } else if (h.m <= 32) {
                    if (h.and_mode) k_filter<unsigned int, true><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                    else k_filter<unsigned int, false><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                } else {
                    if (h.and_mode) k_filter<unsigned long long, true><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                    else k_filter<unsigned long long, false><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                }
                LAUNCHED(ctx);
                if (prof) CU(cudaEventRecord(ctx->pev[4], st));
                if (use_sa || use_qg) {
                    if (h.m <= 32) {
                        if (h.and_mode) k_refine<unsigned int, true><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, refine, narrow, wide, counters);
                        else k_refine<unsigned int, false><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, refine, narrow, wide, counters);
                    } else {
                        if (h.and_mode) k_refine<unsigned long long, true><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, refine, narrow, wide, counters);
                        else k_refine<unsigned long long, false><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, refine, narrow, wide, counters);
                    }
                    LAUNCHED(ctx);
                }
                if (prof) CU(cudaEventRecord(ctx->pev[1], st));
                // the two DP kernels work on disjoint survivors and records: unless the per-kernel timing is on,
                // the (small, latency-bound) windowed kernel runs on a side stream underneath the banded one
                cudaStream_t sw = prof ? st : slot.aux;
                if (!prof) { CU(cudaEventRecord(slot.ev_fork, st)); CU(cudaStreamWaitEvent(sw, slot.ev_fork, 0)); }
                if (h.and_mode) k_band<true, 16><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, narrow, counters + 0);
                else k_band<false, 16><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, narrow, counters + 0);
                LAUNCHED(ctx);
                if (h.and_mode) k_band<true, 8><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, wide, counters + 3);
                else k_band<false, 8><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, wide, counters + 3);
                LAUNCHED(ctx);
                if (prof) CU(cudaEventRecord(ctx->pev[2], st));
                if (h.and_mode) k_wide<true><<<gp, 128, 0, sw>>>(p, d_codes, d_woff, d_len, d_win, d_out, wide, counters + 1);
                else k_wide<false><<<gp, 128, 0, sw>>>(p, d_codes, d_woff, d_len, d_win, d_out, wide, counters + 1);
                if (!prof) { LAUNCHED(ctx); ctx->launches--; CU(cudaEventRecord(slot.ev_join, sw)); CU(cudaStreamWaitEvent(st, slot.ev_join, 0)); }
                if (prof) { LAUNCHED(ctx); CU(cudaEventRecord(ctx->pev[3], st)); ctx->launches--; ctx->phases_valid = 1; }
            }
            else if (p.filter_only && !p.anchor_ok && !ctx->disable_fused && n < (int64_t)0x7fffffff) {
                // funnel shape, dearer indels: the funnel's filter kernel, then the register DP over all survivor lists
                int rc = lists.ensure((size_t)n * 3 * sizeof(Survivor) + 64);
                if (rc) return fail(ctx, rc, "out of device memory (survivor lists)");
                int* counters = lists.as<int>();
                Survivor* narrow = (Survivor*)(lists.as<char>() + 64);
                Survivor* wide = narrow + n;
                Survivor* refine = wide + n;
                CU(cudaMemsetAsync(counters, 0, 4 * sizeof(int), st));
                const unsigned g = grid_for(n, ATR_K1F_THREADS);
                const unsigned gp = (unsigned)std::min<int64_t>((n + 127) / 128, 148 * 12);
                const bool use_sa = p.sa_ok && !ctx->disable_sa;
                const bool use_qg = p.qg_ok && !ctx->disable_qg && (p.qg_wide || (use_sa && p.qg_step == 3));   // at step 2 (7-row pieces) the automaton wins: 0.38 vs 0.48 ms per 5 M reads for the 21-mer
                if (use_qg) {
                    const unsigned gq = (unsigned)std::min<int64_t>((n + ATR_QG_THREADS - 1) / ATR_QG_THREADS, (int64_t)ctx->sm_count * ctx->qg_ctas);
                    if (p.qg_wide) {
                        if (p.qg_step == 3) k_filter_qg<3, unsigned long long><<<gq, ATR_QG_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                        else k_filter_qg<2, unsigned long long><<<gq, ATR_QG_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                    } else {
                        if (p.qg_step == 3) k_filter_qg<3, unsigned><<<gq, ATR_QG_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                        else k_filter_qg<2, unsigned><<<gq, ATR_QG_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                    }
                } else if (use_sa) {
                    if (h.and_mode) k_filter_sa<true><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                    else k_filter_sa<false><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, refine, counters);
                } else if (h.m <= 32) {
                    if (h.and_mode) k_filter<unsigned int, true><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                    else k_filter<unsigned int, false><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                } else {
                    if (h.and_mode) k_filter<unsigned long long, true><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                    else k_filter<unsigned long long, false><<<g, ATR_K1F_THREADS, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, narrow, wide, counters);
                }
                LAUNCHED(ctx);
                Survivor* ls[3] = {narrow, wide, refine};
                for (int li = 0; li < ((use_sa || use_qg) ? 3 : 2); li++) {
                    // narrow stays empty here (band_ok = 0 without unit indel cost): wide and refine entries carry windows
                    if (h.and_mode) k_anchor_dp<true><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, ls[li], counters + li, li > 0);
                    else k_anchor_dp<false><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, ls[li], counters + li, li > 0);
                    if (li + 1 < ((use_sa || use_qg) ? 3 : 2)) LAUNCHED(ctx);
                }
            }
            else if (p.anchor_ok && !ctx->disable_fused && n < (int64_t)0x7fffffff) {
                // anchored adapter with indels: fixed-position piece filter, register DP over the survivors only
                int rc = lists.ensure((size_t)n * 3 * sizeof(Survivor) + 64);
                if (rc) return fail(ctx, rc, "out of device memory (survivor lists)");
                int* counters = lists.as<int>();
                Survivor* surv = (Survivor*)(lists.as<char>() + 64);
                CU(cudaMemsetAsync(counters, 0, 4 * sizeof(int), st));
                const unsigned gp = (unsigned)std::min<int64_t>((n + 127) / 128, 148 * 12);
                if (h.and_mode) k_filter_anchor<true><<<grid_for(n, 256), 256, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, surv, counters);
                else k_filter_anchor<false><<<grid_for(n, 256), 256, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out, surv, counters);
                LAUNCHED(ctx);
                if (h.and_mode) k_anchor_dp<true><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, surv, counters, 0);
                else k_anchor_dp<false><<<gp, 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, d_out, surv, counters, 0);
            }
            else if (h.and_mode) k_locate_k1a<true><<<grid_for(n, 128), 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out);
            else k_locate_k1a<false><<<grid_for(n, 128), 128, 0, st>>>(p, d_codes, d_woff, d_len, d_win, n, d_out);
            LAUNCHED(ctx);
            if (have_ascii) {
                AdapterGen g = set->gen[a];
                g.reduce = a > 0;
                k_locate_gen<<<(unsigned)(gthreads / gblock), gblock, 0, st>>>(g, h.and_mode ? 1 : 0, 0, ctx->d_tables, d_ascii,
                                                                              d_offsets, base, d_len, d_win, fold_case, n,
                                                                              gen_scratch.as<GCell>(), d_out);
                LAUNCHED(ctx);
            }
        } else {
            if (!have_ascii)
                return fail(ctx, ATR_E_ARG, "this adapter needs the byte-exact kernel (longer than 64 nt, error budget > 126, or "
                                            "letters outside the IUPAC set): pass d_ascii/d_offsets");
            AdapterGen g = set->gen[a];
            g.reduce = a > 0;
            k_locate_gen<<<(unsigned)(gthreads / gblock), gblock, 0, st>>>(g, 0, 1, ctx->d_tables, d_ascii, d_offsets, base,
                                                                          d_len, d_win, fold_case, n, gen_scratch.as<GCell>(), d_out);
            LAUNCHED(ctx);
        }
    }
    return ATR_OK;
}

int insert_on_stream(atr_ctx* ctx, cudaStream_t st, const atr_insertset* set,
                     const uint32_t* c1, const uint32_t* w1, const uint16_t* l1,
                     const uint32_t* c2, const uint32_t* w2, const uint16_t* l2,
                     const uint8_t* a1, const int64_t* o1, int64_t base1,
                     const uint8_t* a2, const int64_t* o2, int64_t base2, int64_t n, atr_insert_result* d_out) {
    if (n <= 0) return ATR_OK;
    k_insert_packed<<<grid_for(n, ATR_K2_THREADS), ATR_K2_THREADS, 0, st>>>(set->dev, c1, w1, l1, c2, w2, l2, n, d_out);
    LAUNCHED(ctx);
    if (a1 && o1 && a2 && o2) {
        const unsigned g = (unsigned)std::min<int64_t>((n + 127) / 128, 148 * 8);
        k_insert_bytes<<<g, 128, 0, st>>>(set->dev, a1, o1, base1, a2, o2, base2, n, d_out);
        LAUNCHED(ctx);
    }
    return ATR_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int atr_abi_version(void) { return ATR_ABI_VERSION; }

int atr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* atr_last_error(const atr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int atr_ctx_create(int device, atr_ctx** out) {
    if (!out) return fail(nullptr, ATR_E_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, ATR_E_CUDA, "no CUDA device: this engine has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(nullptr, ATR_E_ARG, "bad device index");
    atr_ctx* ctx = new (std::nothrow) atr_ctx();
    if (!ctx) return fail(nullptr, ATR_E_NOMEM, "out of host memory");
    ctx->device = device;
    { const char* e = getenv("ATR_DISABLE_FUSED"); ctx->disable_fused = (e && e[0] == '1'); }
    { const char* e = getenv("ATR_DISABLE_SA"); ctx->disable_sa = (e && e[0] == '1'); }
    { const char* e = getenv("ATR_DISABLE_QG"); ctx->disable_qg = (e && e[0] == '1'); }
    { const char* e = getenv("ATR_QG_CTAS"); if (e && atoi(e) > 0) ctx->qg_ctas = atoi(e); }
    { const char* e = getenv("ATR_PANEL_CHUNK"); if (e && atoll(e) > 0) ctx->panel_chunk = atoll(e); }
    CU(cudaSetDevice(device));
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ctx->sm_count = v; }
    for (int s = 0; s < 2; s++) {
        CU(cudaStreamCreateWithFlags(&ctx->slot[s].stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->slot[s].aux, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->slot[s].ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->slot[s].ev_join, cudaEventDisableTiming));
    }
    CU(cudaEventCreate(&ctx->ev0));
    CU(cudaEventCreate(&ctx->ev1));
    atr::build_tables(ctx->h_tables);
    CU(cudaMalloc((void**)&ctx->d_tables, sizeof(AtrTables)));
    CU(cudaMemcpy(ctx->d_tables, &ctx->h_tables, sizeof(AtrTables), cudaMemcpyHostToDevice));
    *out = ctx;
    return ATR_OK;
}

void atr_ctx_destroy(atr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int s = 0; s < 2; s++) {
        if (ctx->slot[s].stream) { cudaStreamSynchronize(ctx->slot[s].stream); cudaStreamDestroy(ctx->slot[s].stream); }
        if (ctx->slot[s].aux) { cudaStreamSynchronize(ctx->slot[s].aux); cudaStreamDestroy(ctx->slot[s].aux); }
        if (ctx->slot[s].ev_fork) cudaEventDestroy(ctx->slot[s].ev_fork);
        if (ctx->slot[s].ev_join) cudaEventDestroy(ctx->slot[s].ev_join);
        ctx->slot[s].release();
    }
    ctx->misc.release();
    ctx->fq_stats.release();
    for (int i = 0; i < 5; i++) if (ctx->pev[i]) cudaEventDestroy(ctx->pev[i]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    delete ctx;
}

int atr_ctx_sync(atr_ctx* ctx) {
    if (!ctx) return fail(nullptr, ATR_E_ARG, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    for (int s = 0; s < 2; s++) CU(cudaStreamSynchronize(ctx->slot[s].stream));
    return ATR_OK;
}

void* atr_ctx_stream(atr_ctx* ctx) { return ctx ? (void*)ctx->slot[0].stream : nullptr; }

int64_t atr_ctx_launch_count(atr_ctx* ctx, int reset) {
    if (!ctx) return 0;
    const int64_t v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

int atr_ctx_set_profiling(atr_ctx* ctx, int on) {
    if (!ctx) return fail(nullptr, ATR_E_ARG, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    if (on && !ctx->pev[0]) for (int i = 0; i < 5; i++) CU(cudaEventCreate(&ctx->pev[i]));
    ctx->profile = on ? 1 : 0;
    ctx->phases_valid = 0;
    return ATR_OK;
}

int atr_ctx_last_phase_ms(atr_ctx* ctx, float* out_ms, int n) {
    if (!ctx || !out_ms || !ctx->profile || !ctx->phases_valid) return 0;
    cudaSetDevice(ctx->device);
    if (cudaEventSynchronize(ctx->pev[3]) != cudaSuccess) { cudaGetLastError(); return 0; }
    // events: 0 start, 4 after the filter kernel, 1 after the refine kernel (if any), 2 after band, 3 after wide
    const int order[5] = {0, 4, 1, 2, 3};
    int k = 0;
    for (; k < 4 && k < n; k++)
        if (cudaEventElapsedTime(&out_ms[k], ctx->pev[order[k]], ctx->pev[order[k + 1]]) != cudaSuccess) { cudaGetLastError(); break; }
    return k;
}

const char* atr_ctx_last_phase_name(atr_ctx* ctx, int i) {
    if (!ctx || i < 0 || i >= 4 || !ctx->phases_valid) return "";
    return ctx->phase_names[i];
}

float atr_ctx_last_kernel_ms(atr_ctx* ctx) {
    if (!ctx) return -1.f;
    if (ctx->last_ms == -2.f) {      // events recorded by the last *_device call, not read yet
        float ms = -1.f;
        cudaSetDevice(ctx->device);
        if (cudaEventSynchronize(ctx->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess)
            ctx->last_ms = ms;
        else { cudaGetLastError(); ctx->last_ms = -1.f; }
    }
    return ctx->last_ms;
}

// ---- adapter sets ---------------------------------------------------------------------------
int atr_adapterset_create(atr_ctx* ctx, int32_t n_adapters, const atr_adapter_desc* descs, atr_adapterset** out) {
    if (!ctx || !out || !descs || n_adapters < 1) return fail(ctx, ATR_E_ARG, "bad arguments to atr_adapterset_create");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    atr_adapterset* set = new (std::nothrow) atr_adapterset();
    if (!set) return fail(ctx, ATR_E_NOMEM, "out of host memory");
    set->ctx = ctx;
    set->host.resize((size_t)n_adapters);
    set->k1a.resize((size_t)n_adapters);
    set->gen.resize((size_t)n_adapters);
    for (int a = 0; a < n_adapters; a++) {
        std::string msg;
        atr::HostAdapter& h = set->host[(size_t)a];
        int rc = atr::prepare_adapter(descs[a], ctx->h_tables, h, msg);
        if (rc) { atr_adapterset_destroy(set); return fail(ctx, rc, msg); }
        set->max_m = std::max(set->max_m, h.m);
        const unsigned char *d_ref = nullptr, *d_lit = nullptr, *d_rmp = nullptr;
        const unsigned short *d_mul = nullptr, *d_div = nullptr;
        rc = upload(ctx, set->dev_allocs, h.ref_gen.data(), h.ref_gen.size(), &d_ref);
        if (!rc) rc = upload(ctx, set->dev_allocs, (const unsigned char*)h.seq.data(), h.seq.size(), &d_lit);
        if (!rc) rc = upload(ctx, set->dev_allocs, h.thr_mul.data(), h.thr_mul.size(), &d_mul);
        if (!rc) rc = upload(ctx, set->dev_allocs, h.thr_div.data(), h.thr_div.size(), &d_div);
        if (!rc && !h.rmp_ok.empty()) rc = upload(ctx, set->dev_allocs, h.rmp_ok.data(), h.rmp_ok.size(), &d_rmp);
        if (rc) { atr_adapterset_destroy(set); return rc; }
        atr::fill_gen(h, a, a > 0, d_ref, d_lit, d_mul, d_div, d_rmp, set->gen[(size_t)a]);
        if (h.k1a_ok) {
            AdapterK1a& ka = set->k1a[(size_t)a];
            atr::fill_k1a(h, ctx->h_tables, a, a > 0, d_rmp, ka);
            std::vector<unsigned char> qtab;
            if (atr::build_qg(ka, qtab)) {
                const unsigned char* d_qtab = nullptr;
                rc = upload(ctx, set->dev_allocs, qtab.data(), qtab.size(), &d_qtab);
                if (rc) { atr_adapterset_destroy(set); return rc; }
                ka.qg_tab = d_qtab;
            }
        }
    }
    set->shadowed.assign((size_t)n_adapters, 0);
    for (int a = 1; a < n_adapters; a++) {
        const atr::HostAdapter& x = set->host[(size_t)a];
        for (int b = 0; b < a && !set->shadowed[(size_t)a]; b++) {
            const atr::HostAdapter& y = set->host[(size_t)b];
            const atr_adapter_desc &p = x.desc, &q = y.desc;
            if (x.seq == y.seq && p.max_error_rate == q.max_error_rate && p.flags == q.flags && p.wildcard_ref == q.wildcard_ref &&
                p.wildcard_query == q.wildcard_query && p.min_overlap == q.min_overlap && p.indel_cost == q.indel_cost &&
                p.match_to_semantics == q.match_to_semantics && p.no_indels == q.no_indels && x.rmp_ok == y.rmp_ok)
                set->shadowed[(size_t)a] = 1;
        }
    }
    *out = set;
    return ATR_OK;
}

void atr_adapterset_destroy(atr_adapterset* set) {
    if (!set) return;
    if (set->ctx) cudaSetDevice(set->ctx->device);
    for (void* p : set->dev_allocs) cudaFree(p);
    delete set;
}

// ---- packing ----------------------------------------------------------------------------------
int64_t atr_packed_words(const int64_t* offsets, int64_t n) {
    int64_t w = 0;
    for (int64_t i = 0; i < n; i++) w += (offsets[i + 1] - offsets[i] + 7) >> 3;
    return w;
}

int atr_pack_device(atr_ctx* ctx, const uint8_t* d_ascii, const int64_t* d_offsets, int64_t n, int fold_case,
                    uint32_t* d_codes, uint32_t* d_woff, uint16_t* d_len) {
    if (!ctx || !d_ascii || !d_offsets || !d_codes || !d_woff || !d_len || n < 0) return fail(ctx, ATR_E_ARG, "bad arguments to atr_pack_device");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[0];
    // offsets are absolute into d_ascii: base 0
    return pack_on_stream(ctx, s.stream, s.counts, s.scan_tmp, d_ascii, d_offsets, 0, n, fold_case, d_codes, d_woff, d_len);
}

// ---- locate -------------------------------------------------------------------------------------
int atr_locate_batch_device(atr_ctx* ctx, const atr_adapterset* set, const uint32_t* d_codes, const uint32_t* d_woff,
                            const uint16_t* d_len, const uint16_t* d_win, const uint8_t* d_ascii, const int64_t* d_offsets,
                            int fold_case, int64_t n, atr_match* d_out) {
    if (!ctx || !set || !d_out || n < 0) return fail(ctx, ATR_E_ARG, "bad arguments to atr_locate_batch_device");
    if (set->ctx != ctx) return fail(ctx, ATR_E_ARG, "adapter set belongs to another context");
    CU(cudaSetDevice(ctx->device));
    {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, d_out) != cudaSuccess || pa.type != cudaMemoryTypeDevice || pa.device != ctx->device) {
            cudaGetLastError();
            return fail(ctx, ATR_E_ARG, "d_out is not device memory of this context's GPU");
        }
        if (reinterpret_cast<uintptr_t>(d_out) & 15) return fail(ctx, ATR_E_ARG, "d_out must be 16-byte aligned");
    }
    Slot& s = ctx->slot[0];
    CU(cudaEventRecord(ctx->ev0, s.stream));
    int rc = ATR_OK;
    if (set->host.size() > 1 && n > ctx->panel_chunk) {
        // a panel runs one pass per adapter: in pieces of panel_chunk reads (packed reads + records ~ 50 MB) every pass
        // after the first finds the reads in the 126 MB L2 instead of streaming them from HBM again
        for (int64_t c0 = 0; c0 < n && !rc; c0 += ctx->panel_chunk) {
            const int64_t cn = std::min<int64_t>(ctx->panel_chunk, n - c0);
            rc = locate_on_stream(ctx, s, set, d_codes, d_woff ? d_woff + c0 : nullptr, d_len ? d_len + c0 : nullptr,
                                  d_win ? d_win + 2 * c0 : nullptr, d_ascii, d_offsets ? d_offsets + c0 : nullptr, 0,
                                  fold_case, cn, d_out + c0);
        }
    } else {
        rc = locate_on_stream(ctx, s, set, d_codes, d_woff, d_len, d_win, d_ascii, d_offsets, 0, fold_case, n, d_out);
    }
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev1, s.stream));
    ctx->last_ms = -2.f;             // resolved lazily by atr_ctx_last_kernel_ms()
    return ATR_OK;
}

int atr_locate_batch_host(atr_ctx* ctx, const atr_adapterset* set, const uint8_t* ascii, const int64_t* offsets,
                          const uint16_t* win, int64_t n, int fold_case, atr_match* out) {
    if (!ctx || !set || !offsets || !out || n < 0 || (n > 0 && !ascii && offsets[n] > offsets[0]))
        return fail(ctx, ATR_E_ARG, "bad arguments to atr_locate_batch_host");
    CU(cudaSetDevice(ctx->device));
    // chunks of <= 1 Mi reads and <= 256 MiB of bases, alternating between the two slots so that the H2D copy
    // of chunk c+1 overlaps the kernels of chunk c and the D2H copy of chunk c-1
    const int64_t max_reads = 1 << 20, max_bytes = (int64_t)256 << 20;
    int64_t c0 = 0;
    int which = 0;
    while (c0 < n) {
        int64_t c1 = std::min(n, c0 + max_reads);
        while (c1 > c0 + 1 && offsets[c1] - offsets[c0] > max_bytes) c1 = c0 + (c1 - c0) / 2;
        const int64_t cn = c1 - c0, bytes = offsets[c1] - offsets[c0];
        NvtxRange nvtx_chunk("atr_locate_batch_host: chunk (H2D, pack, locate, D2H enqueue)");
        bool uniform = true;
        const int64_t len0 = offsets[c0 + 1] - offsets[c0];
        for (int64_t i = c0; i < c1; i++) {
            const int64_t li = offsets[i + 1] - offsets[i];
            if (li > ATR_MAX_READ || li < 0) return fail(ctx, ATR_E_LIMIT, "read longer than 32767 nt (or offsets not monotone)");
            uniform = uniform && li == len0;
        }
        Slot& s = ctx->slot[which];
        int rc = s.ascii.ensure((size_t)bytes + 16);
        if (!rc) rc = s.offsets.ensure((size_t)(cn + 1) * sizeof(int64_t));
        if (!rc) rc = s.woff.ensure((size_t)(cn + 1) * sizeof(uint32_t));
        if (!rc) rc = s.codes.ensure((size_t)(bytes / 8 + cn + 2) * sizeof(uint32_t));
        if (!rc) rc = s.len.ensure((size_t)cn * sizeof(uint16_t));
        if (!rc) rc = s.out.ensure((size_t)cn * sizeof(atr_match));
        if (!rc && win) rc = s.win.ensure((size_t)cn * 2 * sizeof(uint16_t));
        if (rc) return fail(ctx, rc, "out of device memory (host entry point staging)");
        if (bytes) CU(cudaMemcpyAsync(s.ascii.p, ascii + offsets[c0], (size_t)bytes, cudaMemcpyHostToDevice, s.stream));
        if (uniform) {      // all reads of the chunk have one length: build the offsets on the device
            k_make_offsets<<<grid_for(cn + 1, 256), 256, 0, s.stream>>>(s.offsets.as<int64_t>(), cn, offsets[c0], len0);
            LAUNCHED(ctx);
        } else {
            CU(cudaMemcpyAsync(s.offsets.p, offsets + c0, (size_t)(cn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        }
        if (win) CU(cudaMemcpyAsync(s.win.p, win + 2 * c0, (size_t)cn * 2 * sizeof(uint16_t), cudaMemcpyHostToDevice, s.stream));
        rc = pack_on_stream(ctx, s.stream, s.counts, s.scan_tmp, s.ascii.as<uint8_t>(), s.offsets.as<int64_t>(), offsets[c0], cn,
                            fold_case, s.codes.as<uint32_t>(), s.woff.as<uint32_t>(), s.len.as<uint16_t>());
        if (rc) return rc;
        rc = locate_on_stream(ctx, s, set, s.codes.as<uint32_t>(), s.woff.as<uint32_t>(),
                              s.len.as<uint16_t>(), win ? s.win.as<uint16_t>() : nullptr, s.ascii.as<uint8_t>(),
                              s.offsets.as<int64_t>(), offsets[c0], fold_case, cn, s.out.as<atr_match>());
        if (rc) return rc;
        CU(cudaMemcpyAsync(out + c0, s.out.p, (size_t)cn * sizeof(atr_match), cudaMemcpyDeviceToHost, s.stream));
        c0 = c1;
        which ^= 1;
        // a slot's buffers are reused two chunks later: same stream, so stream order protects them
    }
    for (int s = 0; s < 2; s++) CU(cudaStreamSynchronize(ctx->slot[s].stream));
    ctx->last_ms = -1.f;
    return ATR_OK;
}

int atr_locate_batch_host_packed(atr_ctx* ctx, const atr_adapterset* set, const uint32_t* codes, const uint32_t* woff,
                                 const uint16_t* len, const uint16_t* win, const uint8_t* ascii, const int64_t* offsets,
                                 int fold_case, int64_t n, atr_match* out) {
    if (!ctx || !set || !codes || !woff || !len || !out || n < 0) return fail(ctx, ATR_E_ARG, "bad arguments to atr_locate_batch_host_packed");
    if (set->ctx != ctx) return fail(ctx, ATR_E_ARG, "adapter set belongs to another context");
    CU(cudaSetDevice(ctx->device));
    const bool have_ascii = ascii != nullptr && offsets != nullptr;
    const int64_t max_reads = 1 << 20;
    std::vector<int64_t> esc_offsets;
    std::vector<uint8_t> esc_bytes;
    int64_t c0 = 0;
    int which = 0;
    while (c0 < n) {
        int64_t c1 = std::min(n, c0 + max_reads);
        while (c1 > c0 + 1 && (int64_t)(woff[c1] - woff[c0]) > ((int64_t)64 << 20)) c1 = c0 + (c1 - c0) / 2;     // <= 256 MiB of codes
        const int64_t cn = c1 - c0;
        NvtxRange nvtx_chunk("atr_locate_batch_host_packed: chunk");
        const uint32_t w0 = woff[c0], nwords = woff[c1] - w0;
        bool uniform = true, any_esc = false;
        for (int64_t i = c0; i < c1; i++) {
            uniform = uniform && len[i] == len[c0] && woff[i + 1] - woff[i] == woff[c0 + 1] - woff[c0];
            any_esc = any_esc || (len[i] & ATR_ESC_BIT);
        }
        Slot& s = ctx->slot[which];
        int rc = s.codes.ensure(((size_t)nwords + 16) * sizeof(uint32_t));
        if (!rc) rc = s.woff.ensure((size_t)(cn + 1) * sizeof(uint32_t));
        if (!rc) rc = s.len.ensure((size_t)cn * sizeof(uint16_t));
        if (!rc) rc = s.out.ensure((size_t)cn * sizeof(atr_match));
        if (!rc && win) rc = s.win.ensure((size_t)cn * 2 * sizeof(uint16_t));
        if (rc) return fail(ctx, rc, "out of device memory (packed host entry point staging)");
        // the chunk's words land (w0 & 3) words into the buffer, so that `base + woff[i]` addresses them with the caller's
        // own word offsets and 16-byte alignment (the TMA tile copies) is the same as in the caller's array
        uint32_t* d_words = s.codes.as<uint32_t>() + (w0 & 3u);
        const uint32_t* d_base = d_words - w0;
        if (nwords) CU(cudaMemcpyAsync(d_words, codes + w0, (size_t)nwords * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
        if (uniform) {
            k_make_packed_index<<<grid_for(cn + 1, 256), 256, 0, s.stream>>>(s.woff.as<uint32_t>(), s.len.as<uint16_t>(), cn, w0,
                                                                              woff[c0 + 1] - woff[c0], len[c0]);
            LAUNCHED(ctx);
        } else {
            CU(cudaMemcpyAsync(s.woff.p, woff + c0, (size_t)(cn + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
            CU(cudaMemcpyAsync(s.len.p, len + c0, (size_t)cn * sizeof(uint16_t), cudaMemcpyHostToDevice, s.stream));
        }
        if (win) CU(cudaMemcpyAsync(s.win.p, win + 2 * c0, (size_t)cn * 2 * sizeof(uint16_t), cudaMemcpyHostToDevice, s.stream));
        const uint8_t* d_ascii = nullptr;
        const int64_t* d_offsets = nullptr;
        if (any_esc && have_ascii) {
            // only the escaped reads' bytes travel: offsets of zero-length entries for every other read of the chunk
            CU(cudaStreamSynchronize(s.stream));          // the staging vectors of the previous use of this slot are free
            esc_offsets.assign((size_t)cn + 1, 0);
            esc_bytes.clear();
            for (int64_t i = 0; i < cn; i++) {
                esc_offsets[(size_t)i] = (int64_t)esc_bytes.size();
                if (len[c0 + i] & ATR_ESC_BIT) esc_bytes.insert(esc_bytes.end(), ascii + offsets[c0 + i], ascii + offsets[c0 + i + 1]);
            }
            esc_offsets[(size_t)cn] = (int64_t)esc_bytes.size();
            rc = s.ascii.ensure(esc_bytes.size() + 16);
            if (!rc) rc = s.offsets.ensure((size_t)(cn + 1) * sizeof(int64_t));
            if (rc) return fail(ctx, rc, "out of device memory (escaped reads)");
            CU(cudaMemcpyAsync(s.ascii.p, esc_bytes.data(), esc_bytes.size(), cudaMemcpyHostToDevice, s.stream));
            CU(cudaMemcpyAsync(s.offsets.p, esc_offsets.data(), (size_t)(cn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
            CU(cudaStreamSynchronize(s.stream));          // pageable staging: the vectors are reused by the next chunk
            d_ascii = s.ascii.as<uint8_t>();
            d_offsets = s.offsets.as<int64_t>();
        }
        rc = locate_on_stream(ctx, s, set, d_base, s.woff.as<uint32_t>(), s.len.as<uint16_t>(), win ? s.win.as<uint16_t>() : nullptr,
                              d_ascii, d_offsets, 0, fold_case, cn, s.out.as<atr_match>());
        if (rc) return rc;
        CU(cudaMemcpyAsync(out + c0, s.out.p, (size_t)cn * sizeof(atr_match), cudaMemcpyDeviceToHost, s.stream));
        c0 = c1;
        which ^= 1;
    }
    for (int s = 0; s < 2; s++) CU(cudaStreamSynchronize(ctx->slot[s].stream));
    ctx->last_ms = -1.f;
    return ATR_OK;
}

// ---- compare_prefixes -----------------------------------------------------------------------------
int atr_compare_prefixes(atr_ctx* ctx, const char* ref, int32_t m, const char* query, int32_t n, int wildcard_ref,
                         int wildcard_query, int32_t* out6) {
    if (!ctx || !out6 || m < 0 || n < 0 || (m && !ref) || (n && !query)) return fail(ctx, ATR_E_ARG, "bad arguments to atr_compare_prefixes");
    CU(cudaSetDevice(ctx->device));
    const int length = std::min(m, n);
    int matches = 0;
    if (length > 0) {
        cudaStream_t st = ctx->slot[0].stream;
        int rc = ctx->misc.ensure((size_t)2 * length + 64);
        if (rc) return fail(ctx, rc, "out of device memory");
        unsigned char* d = ctx->misc.as<unsigned char>();
        int* d_cnt = (int*)(d + (((size_t)2 * length + 15) & ~(size_t)15));
        CU(cudaMemcpyAsync(d, ref, (size_t)length, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + length, query, (size_t)length, cudaMemcpyHostToDevice, st));
        const int mode = (wildcard_ref || wildcard_query) ? 1 : 0;
        const unsigned char* tr = wildcard_ref ? ctx->d_tables->iupac : ctx->d_tables->acgt;       // _align.pyx:521-530
        const unsigned char* tq = wildcard_query ? ctx->d_tables->iupac : ctx->d_tables->acgt;
        k_compare_prefixes<<<1, 256, 0, st>>>(d, d + length, length, mode, tr, tq, d_cnt);
        LAUNCHED(ctx);
        CU(cudaMemcpyAsync(&matches, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    out6[0] = 0; out6[1] = length; out6[2] = 0; out6[3] = length; out6[4] = matches; out6[5] = length - matches;
    return ATR_OK;
}

// ---- insert aligner -------------------------------------------------------------------------------
int atr_insertset_create(atr_ctx* ctx, const atr_insert_desc* d, atr_insertset** out) {
    if (!ctx || !d || !out) return fail(ctx, ATR_E_ARG, "bad arguments to atr_insertset_create");
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    atr::HostInsert h;
    std::string msg;
    int rc = atr::prepare_insert(*d, ctx->h_tables, h, msg);
    if (rc) return fail(ctx, rc, msg);
    atr_insertset* set = new (std::nothrow) atr_insertset();
    if (!set) return fail(ctx, ATR_E_NOMEM, "out of host memory");
    set->ctx = ctx;
    set->kmax = h.k_by_len.empty() ? 0 : (int)h.k_by_len.back();
    set->dev = h.dev;
    InsertDev& v = set->dev;
    rc = upload(ctx, set->dev_allocs, h.k_by_len.data(), h.k_by_len.size(), &v.k_by_len);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.thr_ins.data(), h.thr_ins.size(), &v.thr_ins);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.maxmm.data(), h.maxmm.size(), &v.maxmm);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.a1_code.data(), h.a1_code.size(), &v.a1_code);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.a2_code.data(), h.a2_code.size(), &v.a2_code);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.a1_pack.data(), h.a1_pack.size(), &v.a1_pack);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.a2_pack.data(), h.a2_pack.size(), &v.a2_pack);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.a1_ascii.data(), h.a1_ascii.size(), &v.a1_ascii);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.a2_ascii.data(), h.a2_ascii.size(), &v.a2_ascii);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.insert_prob.data(), h.insert_prob.size(), &v.insert_prob);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.adapter_prob.data(), h.adapter_prob.size(), &v.adapter_prob);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.comp.data(), h.comp.size(), &v.comp);
    if (!rc) rc = upload(ctx, set->dev_allocs, h.ov_tab.data(), h.ov_tab.size(), &v.ov_tab);
    if (rc) { atr_insertset_destroy(set); return rc; }
    *out = set;
    return ATR_OK;
}

void atr_insertset_destroy(atr_insertset* set) {
    if (!set) return;
    if (set->ctx) cudaSetDevice(set->ctx->device);
    for (void* p : set->dev_allocs) cudaFree(p);
    delete set;
}

int atr_match_insert_batch_device(atr_ctx* ctx, const atr_insertset* set,
                                  const uint32_t* d_codes1, const uint32_t* d_woff1, const uint16_t* d_len1,
                                  const uint32_t* d_codes2, const uint32_t* d_woff2, const uint16_t* d_len2,
                                  const uint8_t* d_ascii1, const int64_t* d_offsets1,
                                  const uint8_t* d_ascii2, const int64_t* d_offsets2,
                                  int64_t n, atr_insert_result* d_out) {
    if (!ctx || !set || !d_codes1 || !d_woff1 || !d_len1 || !d_codes2 || !d_woff2 || !d_len2 || !d_out || n < 0)
        return fail(ctx, ATR_E_ARG, "bad arguments to atr_match_insert_batch_device");
    CU(cudaSetDevice(ctx->device));
    Slot& s = ctx->slot[0];
    CU(cudaEventRecord(ctx->ev0, s.stream));
    int rc = insert_on_stream(ctx, s.stream, set, d_codes1, d_woff1, d_len1, d_codes2, d_woff2, d_len2, d_ascii1, d_offsets1, 0,
                              d_ascii2, d_offsets2, 0, n, d_out);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev1, s.stream));
    ctx->last_ms = -2.f;
    return ATR_OK;
}

int atr_match_insert_batch_host(atr_ctx* ctx, const atr_insertset* set, const uint8_t* ascii1, const int64_t* offsets1,
                                const uint8_t* ascii2, const int64_t* offsets2, int64_t n, atr_insert_result* out) {
    if (!ctx || !set || !offsets1 || !offsets2 || !out || n < 0) return fail(ctx, ATR_E_ARG, "bad arguments to atr_match_insert_batch_host");
    CU(cudaSetDevice(ctx->device));
    const int64_t max_pairs = 1 << 19;
    int64_t c0 = 0;
    int which = 0;
    while (c0 < n) {
        const int64_t c1 = std::min(n, c0 + max_pairs), cn = c1 - c0;
        NvtxRange nvtx_chunk("atr_match_insert_batch_host: chunk");
        const int64_t b1 = offsets1[c1] - offsets1[c0], b2 = offsets2[c1] - offsets2[c0];
        for (int64_t i = c0; i < c1; i++) {
            const int64_t la = offsets1[i + 1] - offsets1[i], lb = offsets2[i + 1] - offsets2[i];
            if (la < 0 || lb < 0 || la > ATR_MAX_READ || lb > ATR_MAX_READ) return fail(ctx, ATR_E_LIMIT, "read longer than 32767 nt");
            if (std::min(la, lb) > set->dev.max_len) return fail(ctx, ATR_E_LIMIT, "read longer than the insert set's max_len tables");
        }
        Slot& s = ctx->slot[which];
        int rc = s.ascii.ensure((size_t)b1 + 16);
        if (!rc) rc = s.ascii2.ensure((size_t)b2 + 16);
        if (!rc) rc = s.offsets.ensure((size_t)(cn + 1) * sizeof(int64_t));
        if (!rc) rc = s.offsets2.ensure((size_t)(cn + 1) * sizeof(int64_t));
        if (!rc) rc = s.woff.ensure((size_t)(cn + 1) * sizeof(uint32_t));
        if (!rc) rc = s.woff2.ensure((size_t)(cn + 1) * sizeof(uint32_t));
        if (!rc) rc = s.codes.ensure((size_t)(b1 / 8 + cn + 2) * sizeof(uint32_t));
        if (!rc) rc = s.codes2.ensure((size_t)(b2 / 8 + cn + 2) * sizeof(uint32_t));
        if (!rc) rc = s.len.ensure((size_t)cn * sizeof(uint16_t));
        if (!rc) rc = s.len2.ensure((size_t)cn * sizeof(uint16_t));
        if (!rc) rc = s.out.ensure((size_t)cn * sizeof(atr_insert_result));
        if (rc) return fail(ctx, rc, "out of device memory (insert staging)");
        if (b1) CU(cudaMemcpyAsync(s.ascii.p, ascii1 + offsets1[c0], (size_t)b1, cudaMemcpyHostToDevice, s.stream));
        if (b2) CU(cudaMemcpyAsync(s.ascii2.p, ascii2 + offsets2[c0], (size_t)b2, cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.offsets.p, offsets1 + c0, (size_t)(cn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.offsets2.p, offsets2 + c0, (size_t)(cn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        rc = pack_on_stream(ctx, s.stream, s.counts, s.scan_tmp, s.ascii.as<uint8_t>(), s.offsets.as<int64_t>(), offsets1[c0], cn, 0,
                            s.codes.as<uint32_t>(), s.woff.as<uint32_t>(), s.len.as<uint16_t>());
        if (!rc) rc = pack_on_stream(ctx, s.stream, s.counts2, s.scan_tmp, s.ascii2.as<uint8_t>(), s.offsets2.as<int64_t>(), offsets2[c0],
                                     cn, 0, s.codes2.as<uint32_t>(), s.woff2.as<uint32_t>(), s.len2.as<uint16_t>());
        if (!rc) rc = insert_on_stream(ctx, s.stream, set, s.codes.as<uint32_t>(), s.woff.as<uint32_t>(), s.len.as<uint16_t>(),
                                       s.codes2.as<uint32_t>(), s.woff2.as<uint32_t>(), s.len2.as<uint16_t>(),
                                       s.ascii.as<uint8_t>(), s.offsets.as<int64_t>(), offsets1[c0],
                                       s.ascii2.as<uint8_t>(), s.offsets2.as<int64_t>(), offsets2[c0], cn,
                                       s.out.as<atr_insert_result>());
        if (rc) return rc;
        CU(cudaMemcpyAsync(out + c0, s.out.p, (size_t)cn * sizeof(atr_insert_result), cudaMemcpyDeviceToHost, s.stream));
        c0 = c1;
        which ^= 1;
    }
    for (int s = 0; s < 2; s++) CU(cudaStreamSynchronize(ctx->slot[s].stream));
    ctx->last_ms = -1.f;
    return ATR_OK;
}

int atr_multi_locate(atr_ctx* ctx, const char* reference, int32_t m, const char* query, int32_t n, double max_error_rate,
                     int32_t flags, int32_t min_overlap, int32_t max_matches, int32_t* out6, int32_t* n_out) {
    if (!ctx || !out6 || !n_out || m < 0 || n < 0 || max_matches < 1 || (m && !reference) || (n && !query))
        return fail(ctx, ATR_E_ARG, "bad arguments to atr_multi_locate");
    if (m > ATR_MAX_READ || n > ATR_MAX_READ) return fail(ctx, ATR_E_LIMIT, "sequence longer than 32767 nt");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slot[0].stream;
    std::vector<unsigned short> thr((size_t)m + 1);
    for (int l = 0; l <= m; l++) thr[l] = thr_mul_of(l, max_error_rate);
    const int k = (int)(max_error_rate * m);
    const size_t n_tuples = (size_t)max_matches + m + 2;
    // layout of the scratch: [ref m][query n][pad][thr (m+1) u16][pad][col (m+1) GCellM][out6][n_out]
    size_t o_ref = 0, o_q = o_ref + (size_t)m, o_thr = (o_q + (size_t)n + 15) & ~(size_t)15;
    size_t o_col = (o_thr + ((size_t)m + 1) * 2 + 15) & ~(size_t)15;
    size_t o_out = o_col + ((size_t)m + 1) * sizeof(GCellM);
    size_t o_cnt = o_out + n_tuples * 6 * sizeof(int);
    int rc = ctx->misc.ensure(o_cnt + 16);
    if (rc) return fail(ctx, rc, "out of device memory");
    unsigned char* d = ctx->misc.as<unsigned char>();
    if (m) CU(cudaMemcpyAsync(d + o_ref, reference, (size_t)m, cudaMemcpyHostToDevice, st));
    if (n) CU(cudaMemcpyAsync(d + o_q, query, (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + o_thr, thr.data(), thr.size() * 2, cudaMemcpyHostToDevice, st));
    k_multi_locate<<<1, 32, 0, st>>>(d + o_ref, m, d + o_q, n, k, (const unsigned short*)(d + o_thr), flags, min_overlap,
                                     max_matches, (GCellM*)(d + o_col), (int*)(d + o_out), (int*)(d + o_cnt));
    LAUNCHED(ctx);
    int cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d + o_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (cnt > 0) CU(cudaMemcpy(out6, d + o_out, (size_t)cnt * 6 * sizeof(int), cudaMemcpyDeviceToHost));
    *n_out = cnt;
    return ATR_OK;
}

}  // extern "C"

#include "atr_merge_api.cuh"
#include "atr_fastq_api.cuh"
