// atr_merge_api.cuh -- atr_merge_overlap_batch_host (included by atr_api.cu after its helpers): the alignment and the
// decision of MergeOverlapping.__call__ (reference: atropos/commands/trim/modifiers.py:864-931) for a batch of pairs.
//
// k_merge_overlap: one thread = one pair (merge_core.cuh). The DP column of a thread is (len2 + 1) cells of 8 bytes,
// element i of thread t at col[i * stride + t]: the threads of a warp walk the rows in lockstep, so a row is one
// coalesced 256-byte access. When the batch's longest read 2 allows it the columns of a CTA live in shared memory
// (dynamic, up to 200 KB: reads up to 199 nt at 128 threads); longer reads use a global scratch of the same shape.
#pragma once
#include "merge_core.cuh"

#define ATR_MERGE_THREADS 128
#define ATR_MERGE_MAX_READ 4000

template <bool SHARED>
__global__ void __launch_bounds__(ATR_MERGE_THREADS) k_merge_overlap(const unsigned char* __restrict__ ascii1, const int64_t* __restrict__ offsets1, int64_t base1,
                                                                     const unsigned char* __restrict__ ascii2, const int64_t* __restrict__ offsets2, int64_t base2,
                                                                     const unsigned char* __restrict__ insert_matched, int64_t n, const MergeTables tb,
                                                                     GCell* __restrict__ scratch, atr_merge_result* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char s_merge[];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    GCell* col = SHARED ? reinterpret_cast<GCell*>(s_merge) + threadIdx.x : scratch + t;
    const long stride = SHARED ? (long)blockDim.x : (long)nthreads;
    for (int64_t p = t; p < n; p += nthreads) {
        const int64_t a0 = offsets1[p] - base1, b0 = offsets2[p] - base2;
        const int len1 = (int)(offsets1[p + 1] - base1 - a0), len2 = (int)(offsets2[p + 1] - base2 - b0);
        merge_pair(ascii1 + a0, len1, ascii2 + b0, len2, insert_matched ? (int)insert_matched[p] : 0, tb, col, stride, out + p);
    }
}

namespace {

// thr_mul / minov / comp tables of one call, uploaded next to each other into ctx->misc
int merge_tables(atr_ctx* ctx, cudaStream_t st, int max_len, double min_overlap, double error_rate, MergeTables& tb) {
    std::vector<unsigned short> h;
    unsigned char comp[256];
    atr::build_merge_tables(max_len, min_overlap, error_rate, h, comp);
    const size_t o_comp = (h.size() * 2 + 15) & ~(size_t)15;
    int rc = ctx->misc.ensure(o_comp + 256);
    if (rc) return fail(ctx, rc, "out of device memory (merge tables)");
    unsigned char* d = ctx->misc.as<unsigned char>();
    // pageable sources: cudaMemcpyAsync stages them before returning
    CU(cudaMemcpyAsync(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d + o_comp, comp, 256, cudaMemcpyHostToDevice, st));
    tb.thr_mul = (const unsigned short*)d;
    tb.minov = (const unsigned short*)d + (max_len + 1);
    tb.comp = d + o_comp;
    tb.max_len = max_len;
    return ATR_OK;
}

}  // namespace

extern "C" int atr_merge_overlap_batch_host(atr_ctx* ctx, const uint8_t* ascii1, const int64_t* offsets1, const uint8_t* ascii2,
                                            const int64_t* offsets2, const uint8_t* insert_matched, int64_t n, double min_overlap,
                                            double error_rate, atr_merge_result* out) {
    if (!ctx || !offsets1 || !offsets2 || !out || n < 0 || !(min_overlap > 0) || !(error_rate >= 0) || error_rate > 1)
        return fail(ctx, ATR_E_ARG, "bad arguments to atr_merge_overlap_batch_host");
    if (n == 0) return ATR_OK;
    CU(cudaSetDevice(ctx->device));
    int max_len = 0, max_len2 = 0;
    for (int64_t i = 0; i < n; i++) {
        const int64_t la = offsets1[i + 1] - offsets1[i], lb = offsets2[i + 1] - offsets2[i];
        if (la < 0 || lb < 0) return fail(ctx, ATR_E_ARG, "offsets must not decrease");
        if (la > ATR_MERGE_MAX_READ || lb > ATR_MERGE_MAX_READ) return fail(ctx, ATR_E_LIMIT, "read longer than 4000 nt (merge)");
        max_len = std::max(max_len, (int)std::max(la, lb));
        max_len2 = std::max(max_len2, (int)lb);
    }
    if ((offsets1[n] > offsets1[0] && !ascii1) || (offsets2[n] > offsets2[0] && !ascii2))
        return fail(ctx, ATR_E_ARG, "bad arguments to atr_merge_overlap_batch_host");
    // the tables are shared by both slots: upload them on slot 0's stream and make slot 1 wait for them
    MergeTables tb;
    int rc = merge_tables(ctx, ctx->slot[0].stream, max_len, min_overlap, error_rate, tb);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev0, ctx->slot[0].stream));
    CU(cudaStreamWaitEvent(ctx->slot[1].stream, ctx->ev0, 0));
    const size_t col_bytes = (size_t)(max_len2 + 1) * sizeof(GCell);
    const size_t smem = col_bytes * ATR_MERGE_THREADS;
    const bool use_shared = smem <= (size_t)200 * 1024;
    if (use_shared) CU(cudaFuncSetAttribute(k_merge_overlap<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t max_pairs = 1 << 19;
    int64_t c0 = 0;
    int which = 0;
    float kernel_ms = 0.f;
    while (c0 < n) {
        const int64_t c1 = std::min(n, c0 + max_pairs), cn = c1 - c0;
        const int64_t b1 = offsets1[c1] - offsets1[c0], b2 = offsets2[c1] - offsets2[c0];
        Slot& s = ctx->slot[which];
        cudaStream_t st = s.stream;
        int64_t blocks = std::min<int64_t>((cn + ATR_MERGE_THREADS - 1) / ATR_MERGE_THREADS, 148 * 8);
        if (!use_shared) {
            const size_t budget = (size_t)768 << 20;
            blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)(budget / (col_bytes * ATR_MERGE_THREADS))));
        }
        rc = s.ascii.ensure((size_t)b1 + 16);
        if (!rc) rc = s.ascii2.ensure((size_t)b2 + 16);
        if (!rc) rc = s.offsets.ensure((size_t)(cn + 1) * sizeof(int64_t));
        if (!rc) rc = s.offsets2.ensure((size_t)(cn + 1) * sizeof(int64_t));
        if (!rc) rc = s.out.ensure((size_t)cn * sizeof(atr_merge_result));
        if (!rc && insert_matched) rc = s.win.ensure((size_t)cn);
        if (!rc && !use_shared) rc = s.gen_scratch.ensure((size_t)blocks * ATR_MERGE_THREADS * col_bytes);
        if (rc) return fail(ctx, rc, "out of device memory (merge staging)");
        if (b1) CU(cudaMemcpyAsync(s.ascii.p, ascii1 + offsets1[c0], (size_t)b1, cudaMemcpyHostToDevice, st));
        if (b2) CU(cudaMemcpyAsync(s.ascii2.p, ascii2 + offsets2[c0], (size_t)b2, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(s.offsets.p, offsets1 + c0, (size_t)(cn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(s.offsets2.p, offsets2 + c0, (size_t)(cn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if (insert_matched) CU(cudaMemcpyAsync(s.win.p, insert_matched + c0, (size_t)cn, cudaMemcpyHostToDevice, st));
        const unsigned char* d_im = insert_matched ? s.win.as<unsigned char>() : nullptr;
        if (ctx->profile) CU(cudaEventRecord(ctx->pev[0], st));      // profiling mode: the kernel timed alone, chunk after chunk
        if (use_shared)
            k_merge_overlap<true><<<(unsigned)blocks, ATR_MERGE_THREADS, smem, st>>>(s.ascii.as<unsigned char>(), s.offsets.as<int64_t>(), offsets1[c0],
                                                                                    s.ascii2.as<unsigned char>(), s.offsets2.as<int64_t>(), offsets2[c0],
                                                                                    d_im, cn, tb, nullptr, s.out.as<atr_merge_result>());
        else
            k_merge_overlap<false><<<(unsigned)blocks, ATR_MERGE_THREADS, 0, st>>>(s.ascii.as<unsigned char>(), s.offsets.as<int64_t>(), offsets1[c0],
                                                                                  s.ascii2.as<unsigned char>(), s.offsets2.as<int64_t>(), offsets2[c0],
                                                                                  d_im, cn, tb, s.gen_scratch.as<GCell>(), s.out.as<atr_merge_result>());
        LAUNCHED(ctx);
        if (ctx->profile) {
            float ms = 0.f;
            CU(cudaEventRecord(ctx->pev[1], st));
            CU(cudaEventSynchronize(ctx->pev[1]));
            CU(cudaEventElapsedTime(&ms, ctx->pev[0], ctx->pev[1]));
            kernel_ms += ms;
        }
        CU(cudaMemcpyAsync(out + c0, s.out.p, (size_t)cn * sizeof(atr_merge_result), cudaMemcpyDeviceToHost, st));
        c0 = c1;
        which ^= 1;
    }
    for (int s = 0; s < 2; s++) CU(cudaStreamSynchronize(ctx->slot[s].stream));
    ctx->last_ms = ctx->profile ? kernel_ms : -1.f;      // atr_ctx_last_kernel_ms: sum of the k_merge_overlap launches
    return ATR_OK;
}
