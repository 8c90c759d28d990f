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

// ---- k_merge_warp: one warp = one pair, the DP as a wavefront over the lanes ------------------------------------------
// Lane l owns R = ceil(len2 / 32) consecutive rows of the DP (rows = rc(read 2)) as packed 32-bit cells in
// registers and works on column s - l in step s: the cell above its first row is the last row of lane l-1 in the same
// column, computed one step earlier and handed down with one shuffle per step; the diagonal one is the value it received
// the step before. Every cell of the matrix is evaluated (no data-dependent band): costs are clamped to k + 1 ("dead"),
// which cannot change an accepted alignment because cells above k never feed one (the argument of K1a,
// locate_core.cuh). Packed cell, compared as one unsigned: cost [24,32) | tie-break priority [22,24) (0 while stored) |
// origin + 1024 [10,22) | matches [0,10); min(diag + SUB, up + INS, left + DEL) reproduces the reference's
// "mismatch, then insertion, then deletion" order (_align.pyx:405-419). Row len2 is offered as a candidate after every
// column (:440-458) by the lane that owns it; the last column's candidates (:461-474) are reduced over the warp with
// the reference's first-best-wins order encoded in the score. Limits: len2 <= 320, k + 1 <= 250, len1 <= 3000
// (else k_merge_overlap).
#define MW_COST_SHIFT 24
#define MW_ORG_SHIFT 10
#define MW_ORG_BIAS 1024
#define MW_SUB (1u << MW_COST_SHIFT)
#define MW_INS ((1u << MW_COST_SHIFT) | (1u << 22))
#define MW_DEL ((1u << MW_COST_SHIFT) | (2u << 22))
#define MW_PRIO_CLEAR (~(3u << 22))
__device__ __forceinline__ unsigned mw_key(int cost, int origin, int matches) {
    return ((unsigned)cost << MW_COST_SHIFT) | ((unsigned)(origin + MW_ORG_BIAS) << MW_ORG_SHIFT) | (unsigned)matches;
}
__device__ __forceinline__ int mw_cost(unsigned key) { return (int)(key >> MW_COST_SHIFT); }
__device__ __forceinline__ int mw_origin(unsigned key) { return (int)((key >> MW_ORG_SHIFT) & 0xFFFu) - MW_ORG_BIAS; }
__device__ __forceinline__ int mw_matches(unsigned key) { return (int)(key & 0x3FFu); }

// one pair with R = ceil(len2 / 32) rows per lane (R is uniform over the warp: the caller switches on it). The first
// `a` = len2 - 32*(R-1) lanes own R rows, the others R - 1, so that the rows add up to len2 exactly: all 32 lanes work
// (for R = 1: len2 lanes), only a lane's last register is conditional, and row len2 is always the last cell the last
// lane computes -- the candidate of every column, with no search for it.
template <int R, int LANES>
__device__ __forceinline__ void mw_pair(const unsigned char* __restrict__ r1, const int nq, const unsigned char* __restrict__ r2, const int m,
                                        const bool siq, const int min_ov, const MergeTables& tb, const int wlane,
                                        atr_merge_result& res, atr_merge_result* __restrict__ dst) {
    // LANES = 32: the whole warp works on this pair; 16: each half of the warp on its own pair (all shuffles, votes
    // and reductions name the half's lanes only, so the halves may run different trip counts)
    const unsigned FULL = LANES == 32 ? 0xFFFFFFFFu : (0xFFFFu << (wlane & 16));
    const int lane = wlane & (LANES - 1);
    const int a = m - LANES * (R - 1);                                 // lanes with R rows (1..LANES)
    const bool full = lane < a;
    const int lanes_used = R > 1 ? LANES : m;
    const int row0 = full ? lane * R + 1 : a * R + (lane - a) * (R - 1) + 1;      // row of register 0
    int refc[R];
    bool bad = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = row0 + r;
        refc[r] = 256;
        if ((r < R - 1 || full) && i <= m) { const unsigned char c = tb.comp[r2[m - i]]; bad = bad || c == 0; refc[r] = c; }
    }
    if (__any_sync(FULL, bad)) { res.status = ATR_ST_KEYERROR; if (lane == 0) *dst = res; return; }    // reverse_complement raises
    const int k = (int)tb.thr_mul[m];
    const unsigned DEAD = (unsigned)(k + 1) << MW_COST_SHIFT;
    const unsigned BIAS0 = (unsigned)MW_ORG_BIAS << MW_ORG_SHIFT;      // origin 0, matches 0, cost 0
    const int max_n = siq ? nq : atr_min(nq, m + k);                   // min_n = 0: both flag sets stop within read 1 (:315-321)
    unsigned key[R];
#pragma unroll
    for (int r = 0; r < R; r++) key[r] = mw_key(0, -(row0 + r), 0);    // column 0 (:338-352): cost 0, origin -i for both flag sets
    unsigned diag_in = mw_key(0, -(row0 - 1), 0);                      // the row above this lane's first row, one column back
    unsigned bottom = 0;
    const bool is_last = lane == lanes_used - 1;
    const unsigned c_hi = lane < lanes_used ? (unsigned)max_n : 0u;
    int b_mat = 0, b_cost = m + nq, b_org = 0, b_ref = m, b_q = nq;    // this lane's best candidate (:358-363)
    const int total = max_n + lanes_used - 1;
    const unsigned char* __restrict__ q = r1 - 1 - lane;               // q[s] = read 1 base of this lane's column in step s
    for (int s = 1; s <= total; s++) {
        const int c = s - lane;                                        // this lane's column in this step
        const unsigned up_sh = __shfl_up_sync(FULL, bottom, 1, LANES);
        if ((unsigned)(c - 1) < c_hi) {
            const int qc = (int)q[s];
            // row 0 (:384-388): free start in read 1 (origin = column) or cost = column
            const unsigned top = siq ? BIAS0 + ((unsigned)c << MW_ORG_SHIFT)
                                     : atr_umin(((unsigned)atr_min(c, k + 1) << MW_COST_SHIFT) | BIAS0, DEAD);
            unsigned up = lane == 0 ? top : up_sh;
            unsigned diag = diag_in;
            diag_in = up;
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r < R - 1 || full) {
                    const unsigned left = key[r];
                    const unsigned mn = atr_umin(atr_umin(diag + MW_SUB, up + MW_INS), left + MW_DEL) & MW_PRIO_CLEAR;
                    unsigned nw = refc[r] == qc ? diag + 1u : mn;
                    nw = atr_umin(nw, DEAD);
                    diag = left;
                    key[r] = nw;
                    up = nw;
                }
            }
            bottom = up;
            if (is_last && up < DEAD) {                                // row m after this column, if it is alive (:440-458)
                const int cost = mw_cost(up), org = mw_origin(up), mat = mw_matches(up);
                const int length = m + atr_min(org, 0);
                if (length >= 1 && cost <= (int)tb.thr_mul[length] && (mat > b_mat || (mat == b_mat && cost < b_cost))) {
                    b_mat = mat; b_cost = cost; b_org = org; b_ref = m; b_q = c;
                }
            }
        }
    }
    // score: valid | matches | 1023 - cost | order (the in-loop best first, then the last column's rows top down)
    unsigned score = b_cost != m + nq ? ((1u << 30) | ((unsigned)b_mat << 20) | ((unsigned)(1023 - b_cost) << 10) | 1023u) : 0u;
    if (siq) {                                                         // max_n == n and STOP_WITHIN_SEQ1: the last column (:461-474)
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = row0 + r;
            if ((r < R - 1 || full) && lane < lanes_used && key[r] < DEAD) {
                const int cost = mw_cost(key[r]), org = mw_origin(key[r]), mat = mw_matches(key[r]);
                const int length = i + atr_min(org, 0);
                if (length >= 1 && cost <= (int)tb.thr_mul[length]) {
                    const unsigned sc = (1u << 30) | ((unsigned)mat << 20) | ((unsigned)(1023 - cost) << 10) | (unsigned)(1023 - i);
                    if (sc > score) { score = sc; b_mat = mat; b_cost = cost; b_org = org; b_ref = i; b_q = nq; }
                }
            }
        }
    }
    const unsigned best = __reduce_max_sync(FULL, score);
    if (best == 0u) { if (lane == 0) *dst = res; return; }             // locate() returned None
    const int src = __ffs(__ballot_sync(FULL, score == best)) - 1;
    b_mat = __shfl_sync(FULL, b_mat, src); b_cost = __shfl_sync(FULL, b_cost, src); b_org = __shfl_sync(FULL, b_org, src);
    b_ref = __shfl_sync(FULL, b_ref, src); b_q = __shfl_sync(FULL, b_q, src);
    if (lane == 0) {
        int start1 = 0, start2 = b_org;
        if (b_org < 0) { start1 = -b_org; start2 = 0; }
        res.r2_start = (uint16_t)start1; res.r2_stop = (uint16_t)b_ref;
        res.r1_start = (uint16_t)start2; res.r1_stop = (uint16_t)b_q;
        res.matches = (uint16_t)b_mat; res.errors = (uint16_t)b_cost;
        if (b_mat >= min_ov) {                                         // :900-927
            res.status = ATR_ST_MATCH;
            if (start1 == 0 && b_ref == m) res.action = ATR_MERGE_KEEP1;
            else if (start2 == 0 && b_q == nq) res.action = ATR_MERGE_TAKE2;
            else if (start2 > 0) res.action = ATR_MERGE_APPEND;
            else if (start1 > 0) res.action = ATR_MERGE_PREPEND;
            else res.status = ATR_ST_INVALID;
        }
        *dst = res;
    }
}

// One warp takes two consecutive pairs. If both need the same number R16 = ceil(len2 / 16) <= 10 of rows per lane on 16
// lanes (always, for reads of one length up to 160 nt), each half of the warp runs its own pair: per wavefront step the
// fixed cost (shuffle, loop, the read-1 base) is paid once for two pairs and a lane does twice the cells, 15 instead of
// 31 steps of fill and drain. Otherwise the two pairs run one after the other on all 32 lanes (for reads of mixed
// lengths the host hands the pairs over sorted by ceil(len2 / 16), `order`, so that neighbours usually agree). RHI = most rows per lane
// in the 32-lane mode: 5 (read 2 up to 160 nt; the kernel with the two-pair mode) or 10 (up to 320 nt).
template <int RHI>
__global__ void __launch_bounds__(256) k_merge_warp(const unsigned char* __restrict__ ascii1, const int64_t* __restrict__ offsets1, int64_t base1,
                                                    const unsigned char* __restrict__ ascii2, const int64_t* __restrict__ offsets2, int64_t base2,
                                                    const unsigned char* __restrict__ insert_matched, int64_t n, const MergeTables tb,
                                                    const uint32_t* __restrict__ order, atr_merge_result* __restrict__ out) {
    const unsigned ALL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, half = lane >> 4;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t pp = 2 * warp; pp < n; pp += 2 * nwarps) {
        // this half's pair
        const bool have = pp + half < n;
        const int64_t p = have ? (order ? (int64_t)order[pp + half] : pp + half) : 0;      // order: pairs sorted by rows per lane
        int64_t a0 = 0, b0 = 0;
        int len1 = 0, len2 = 0, min_ov = 0;
        bool siq = true, work = false;
        atr_merge_result res;
        res.r2_start = res.r2_stop = res.r1_start = res.r1_stop = res.matches = res.errors = 0;
        res.min_overlap = 0; res.status = ATR_ST_NONE; res.action = 0;
        if (have) {
            a0 = offsets1[p] - base1; b0 = offsets2[p] - base2;
            len1 = (int)(offsets1[p + 1] - base1 - a0); len2 = (int)(offsets2[p + 1] - base2 - b0);
            min_ov = (int)tb.minov[atr_min(len1, len2)];
            res.min_overlap = (uint16_t)min_ov;
            work = len1 >= min_ov && len2 >= min_ov;                   // :881-882
            if (!work && (lane & 15) == 0) out[p] = res;
            siq = !(insert_matched && insert_matched[p]);              // SEMIGLOBAL; else START_WITHIN_SEQ1 | STOP_WITHIN_SEQ2 (:886-890)
        }
        const int r16 = work ? (len2 + 15) >> 4 : 0;
        const int r16_other = __shfl_xor_sync(ALL, r16, 16);
        if (RHI == 5 && r16 == r16_other && r16 >= 1 && r16 <= 10) {   // both halves at once (the 320-nt kernel keeps its registers for 10 rows per lane)
            const unsigned char* __restrict__ r1 = ascii1 + a0;
            const unsigned char* __restrict__ r2 = ascii2 + b0;
            switch (r16) {
                case 1: mw_pair<1, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                case 2: mw_pair<2, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                case 3: mw_pair<3, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                case 4: mw_pair<4, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                case 5: mw_pair<5, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                case 6: mw_pair<6, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                case 7: mw_pair<7, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                case 8: mw_pair<8, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                case 9: mw_pair<9, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
                default: mw_pair<10, 16>(r1, len1, r2, len2, siq, min_ov, tb, lane, res, out + p); break;
            }
            __syncwarp();
            continue;
        }
        for (int h = 0; h < 2; h++) {                                  // one pair after the other on the whole warp
            const int src = 16 * h;
            if (!__shfl_sync(ALL, (int)work, src)) continue;
            const int64_t q = __shfl_sync(ALL, p, src);
            const int64_t qa0 = __shfl_sync(ALL, a0, src), qb0 = __shfl_sync(ALL, b0, src);
            const int l1 = __shfl_sync(ALL, len1, src), l2 = __shfl_sync(ALL, len2, src), mo = __shfl_sync(ALL, min_ov, src);
            const bool sq = __shfl_sync(ALL, (int)siq, src) != 0;
            const unsigned char* __restrict__ r1 = ascii1 + qa0;
            const unsigned char* __restrict__ r2 = ascii2 + qb0;
            atr_merge_result rs;
            rs.r2_start = rs.r2_stop = rs.r1_start = rs.r1_stop = rs.matches = rs.errors = 0;
            rs.min_overlap = (uint16_t)mo; rs.status = ATR_ST_NONE; rs.action = 0;
            const int R = (l2 + 31) >> 5;                              // rows per lane, uniform over the warp
            switch (R) {
                case 1: mw_pair<1, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                case 2: mw_pair<2, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                case 3: mw_pair<3, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                case 4: mw_pair<4, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                case 5: mw_pair<5, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                default:
                    if (RHI > 5) {
                        switch (R) {
                            case 6: mw_pair<RHI - 4, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                            case 7: mw_pair<RHI - 3, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                            case 8: mw_pair<RHI - 2, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                            case 9: mw_pair<RHI - 1, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                            default: mw_pair<RHI, 32>(r1, l1, r2, l2, sq, mo, tb, lane, rs, out + q); break;
                        }
                    }
                    break;
            }
        }
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

// which kernel serves a batch whose longest reads are max_len1 / max_len2 (read 2 = the DP's rows)
struct MergePlan {
    bool use_warp = false, use_shared = false;
    size_t col_bytes = 0, smem = 0;
};
MergePlan merge_plan(int max_len1, int max_len2, double error_rate) {
    MergePlan P;
    // warp-per-pair wavefront kernel whenever its limits hold (ATR_MERGE_KERNEL=thread forces the other one: tests)
    const char* force = getenv("ATR_MERGE_KERNEL");
    const int kmax = (int)thr_mul_of(max_len2, error_rate);
    P.use_warp = !(force && force[0] == 't') && max_len2 <= 320 && max_len1 <= 3000 && kmax + 1 <= 250;
    P.col_bytes = (size_t)(max_len2 + 1) * sizeof(GCell);
    P.smem = P.col_bytes * ATR_MERGE_THREADS;
    P.use_shared = P.smem <= (size_t)200 * 1024;
    return P;
}

// one launch over cn pairs that are already on the device (offsets absolute, base subtracted); scratch: the slot's
// gen_scratch when the thread-per-pair kernel keeps its columns in global memory
int merge_launch(atr_ctx* ctx, Slot& s, cudaStream_t st, const MergePlan& P, int max_len2, const unsigned char* d_a1, const int64_t* d_o1,
                 int64_t base1, const unsigned char* d_a2, const int64_t* d_o2, int64_t base2, const unsigned char* d_im, int64_t cn,
                 const MergeTables& tb, const uint32_t* d_order, atr_merge_result* d_out) {
    if (cn <= 0) return ATR_OK;
    if (P.use_warp) {
        const unsigned wblocks = (unsigned)std::min<int64_t>((cn + 15) / 16, 148 * 16);      // 8 warps per CTA, two pairs per warp
        if (max_len2 <= 160) k_merge_warp<5><<<wblocks, 256, 0, st>>>(d_a1, d_o1, base1, d_a2, d_o2, base2, d_im, cn, tb, d_order, d_out);
        else k_merge_warp<10><<<wblocks, 256, 0, st>>>(d_a1, d_o1, base1, d_a2, d_o2, base2, d_im, cn, tb, d_order, d_out);
        LAUNCHED(ctx);
        return ATR_OK;
    }
    int64_t blocks = std::min<int64_t>((cn + ATR_MERGE_THREADS - 1) / ATR_MERGE_THREADS, 148 * 8);
    if (P.use_shared) {
        CU(cudaFuncSetAttribute(k_merge_overlap<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
        k_merge_overlap<true><<<(unsigned)blocks, ATR_MERGE_THREADS, P.smem, st>>>(d_a1, d_o1, base1, d_a2, d_o2, base2, d_im, cn, tb, nullptr, d_out);
    } else {
        const size_t budget = (size_t)768 << 20;
        blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)(budget / (P.col_bytes * ATR_MERGE_THREADS))));
        int rc = s.gen_scratch.ensure((size_t)blocks * ATR_MERGE_THREADS * P.col_bytes);
        if (rc) return fail(ctx, rc, "out of device memory (merge scratch)");
        k_merge_overlap<false><<<(unsigned)blocks, ATR_MERGE_THREADS, 0, st>>>(d_a1, d_o1, base1, d_a2, d_o2, base2, d_im, cn, tb,
                                                                              s.gen_scratch.as<GCell>(), d_out);
    }
    LAUNCHED(ctx);
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
    int max_len1 = 0;
    for (int64_t i = 0; i < n; i++) max_len1 = std::max(max_len1, (int)(offsets1[i + 1] - offsets1[i]));
    const MergePlan plan = merge_plan(max_len1, max_len2, error_rate);
    const bool use_warp = plan.use_warp;
    const int64_t max_pairs = 1 << 19;
    int64_t c0 = 0;
    int which = 0;
    float kernel_ms = 0.f;
    std::vector<uint32_t> order_host;
    while (c0 < n) {
        const int64_t c1 = std::min(n, c0 + max_pairs), cn = c1 - c0;
        const int64_t b1 = offsets1[c1] - offsets1[c0], b2 = offsets2[c1] - offsets2[c0];
        Slot& s = ctx->slot[which];
        cudaStream_t st = s.stream;
        rc = s.ascii.ensure((size_t)b1 + 16);
        if (!rc) rc = s.ascii2.ensure((size_t)b2 + 16);
        if (!rc) rc = s.offsets.ensure((size_t)(cn + 1) * sizeof(int64_t));
        if (!rc) rc = s.offsets2.ensure((size_t)(cn + 1) * sizeof(int64_t));
        if (!rc) rc = s.out.ensure((size_t)cn * sizeof(atr_merge_result));
        if (!rc && insert_matched) rc = s.win.ensure((size_t)cn);
        if (rc) return fail(ctx, rc, "out of device memory (merge staging)");
        if (b1) CU(cudaMemcpyAsync(s.ascii.p, ascii1 + offsets1[c0], (size_t)b1, cudaMemcpyHostToDevice, st));
        if (b2) CU(cudaMemcpyAsync(s.ascii2.p, ascii2 + offsets2[c0], (size_t)b2, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(s.offsets.p, offsets1 + c0, (size_t)(cn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(s.offsets2.p, offsets2 + c0, (size_t)(cn + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if (insert_matched) CU(cudaMemcpyAsync(s.win.p, insert_matched + c0, (size_t)cn, cudaMemcpyHostToDevice, st));
        const unsigned char* d_im = insert_matched ? s.win.as<unsigned char>() : nullptr;
        // mixed read lengths: a counting sort of the chunk's pairs by ceil(len2 / 16) lets neighbours share the two-pair mode
        const uint32_t* d_order = nullptr;
        if (use_warp && max_len2 <= 160) {
            bool mixed = false;
            const int64_t first = offsets2[c0 + 1] - offsets2[c0];
            for (int64_t i = c0 + 1; i < c1 && !mixed; i++) mixed = (offsets2[i + 1] - offsets2[i]) != first;
            if (mixed) {
                size_t start[12] = {0};
                for (int64_t i = c0; i < c1; i++) start[((offsets2[i + 1] - offsets2[i] + 15) >> 4) + 1]++;
                for (int b = 1; b < 12; b++) start[b] += start[b - 1];
                order_host.resize((size_t)cn);
                for (int64_t i = c0; i < c1; i++) order_host[start[(offsets2[i + 1] - offsets2[i] + 15) >> 4]++] = (uint32_t)(i - c0);
                rc = s.woff.ensure((size_t)cn * sizeof(uint32_t));
                if (rc) return fail(ctx, rc, "out of device memory (merge staging)");
                CU(cudaMemcpyAsync(s.woff.p, order_host.data(), (size_t)cn * sizeof(uint32_t), cudaMemcpyHostToDevice, st));   // pageable: staged before the call returns
                d_order = s.woff.as<uint32_t>();
            }
        }
        if (ctx->profile) CU(cudaEventRecord(ctx->pev[0], st));      // profiling mode: the kernel timed alone, chunk after chunk
        rc = merge_launch(ctx, s, st, plan, max_len2, s.ascii.as<unsigned char>(), s.offsets.as<int64_t>(), offsets1[c0], s.ascii2.as<unsigned char>(),
                          s.offsets2.as<int64_t>(), offsets2[c0], d_im, cn, tb, d_order, s.out.as<atr_merge_result>());
        if (rc) return rc;
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
