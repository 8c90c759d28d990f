// merge_core.cuh -- per-pair logic of MergeOverlapping (reference: atropos/commands/trim/modifiers.py:864-931).
//
// The reference builds, for every pair, an Aligner whose *reference* is reverse_complement(read 2)
// (util/__init__.py:479-482) and locates read 1 in it (_align.pyx:266-491: unit costs, no wildcards, so the bytes are
// compared as they are; min_overlap 1; k = int(rate * len2)). The flags are SEMIGLOBAL, or
// START_WITHIN_SEQ1 | STOP_WITHIN_SEQ2 when the insert aligner has already matched the pair (:886-893). What follows
// the alignment (:898-929) is a four-way case split on the coordinates. Here one thread owns one pair: the mate is
// never materialised (row i of the DP compares with comp[read2[len2 - i]]), the column holds (cost, origin, matches)
// cells like the general kernel's (locate_core.cuh: GCell), and Ukkonen's cut-off `last` is kept because with both
// reads of a pair as long as each other it is what makes the work O(k * n) instead of O(n^2).
// __host__ __device__ like the other cores: tests/host_sim runs the same code on the CPU against the reference.
#pragma once
#include "locate_core.cuh"

struct MergeTables {                  // device pointers (kernel parameter, by value)
    const unsigned short* thr_mul;    // [max_len + 1] largest c with c <= l * rate         (_align.pyx:312, :447, :468)
    const unsigned short* minov;      // [max_len + 1] max(2, round(frac * l)), or the fixed int(--merge-min-overlap) (:877-879)
    const unsigned char* comp;        // [256] BASE_COMPLEMENTS, 0 = KeyError               (util/__init__.py:67-88)
    int max_len;
};

struct MergeAd {                      // what consider() (locate_core.cuh) reads
    int min_overlap;
    const unsigned short* thr_mul;
};

// Aligner(reverse_complement(read2), rate, flags).locate(read1): best alignment into `best` (best.cost == m + n: None)
ATR_HD void merge_locate(const unsigned char* __restrict__ r1, int n, const unsigned char* __restrict__ r2, int m, int flags,
                         const MergeAd& ad, const unsigned char* __restrict__ comp, GCell* col, long stride, Best& best) {
    const int k = (int)ad.thr_mul[m];
    const bool start_in_ref = flags & ATR_START_WITHIN_SEQ1, start_in_query = flags & ATR_START_WITHIN_SEQ2;
    const bool stop_in_ref = flags & ATR_STOP_WITHIN_SEQ1, stop_in_query = flags & ATR_STOP_WITHIN_SEQ2;
    int max_n = n, min_n = 0;
    if (!start_in_query) max_n = atr_min(n, m + k);
    if (!stop_in_query) min_n = atr_max(0, n - m - k);
    const int dead = k + 1;
    for (int i = 0; i <= m; i++) {                                           // column min_n (:334-352)
        int cost, origin;
        if (!start_in_ref && !start_in_query) { cost = atr_max(i, min_n); origin = 0; }
        else if (start_in_ref && !start_in_query) { cost = min_n; origin = atr_min(0, min_n - i); }
        else if (!start_in_ref && start_in_query) { cost = i; origin = atr_max(0, min_n - i); }
        else { cost = atr_min(i, min_n); origin = min_n - i; }
        GCell c; c.cost = cost > k ? dead : cost; c.pay = g_pay(origin, 0);
        col[i * stride] = c;
    }
    best.ref_stop = m; best.q_stop = n; best.cost = m + n; best.origin = 0; best.matches = 0;
    int last = start_in_ref ? m : atr_min(m, k + 1);                         // :366-368
    for (int j = min_n + 1; j <= max_n; j++) {
        GCell diag = col[0];
        GCell up = diag;
        if (start_in_query) up.pay = g_pay(j, 0);
        else up.cost = j > k ? dead : j;
        col[0] = up;
        const unsigned char qc = r1[j - 1];
        for (int i = 1; i <= last; i++) {
            const GCell left = col[i * stride];
            GCell nw;
            if (comp[r2[m - i]] == qc) { nw.cost = diag.cost; nw.pay = diag.pay + 1; }
            else {                                                            // :405-419: mismatch, then insertion, then deletion
                const int c_sub = diag.cost + 1, c_del = left.cost + 1, c_ins = up.cost + 1;
                if (c_sub <= c_del && c_sub <= c_ins) { nw.cost = c_sub; nw.pay = diag.pay; }
                else if (c_ins <= c_del) { nw.cost = c_ins; nw.pay = up.pay; }
                else { nw.cost = c_del; nw.pay = left.pay; }
            }
            if (nw.cost > k) nw.cost = dead;
            diag = left;
            col[i * stride] = nw;
            up = nw;
        }
        while (last >= 0 && col[last * stride].cost > k) last--;             // :433-439
        if (last < m) last++;
        else if (stop_in_query) {
            const GCell c = col[m * stride];
            consider(ad, best, c.cost, g_origin(c.pay), g_matches(c.pay), m, j);
        }
    }
    if (max_n == n) {                                                         // :461-474
        for (int i = stop_in_ref ? 0 : m; i <= m; i++) {
            const GCell c = col[i * stride];
            if (c.cost <= k) consider(ad, best, c.cost, g_origin(c.pay), g_matches(c.pay), i, n);
        }
    }
}

// atr_merge_result.action
#define ATR_MERGE_KEEP1   1          // read 2 lies inside read 1: read 1 stays as it is                          (:905-907)
#define ATR_MERGE_TAKE2   2          // read 1 lies inside read 2: read 1 := rc(read 2), reversed qualities       (:908-911)
#define ATR_MERGE_APPEND  3          // read 1 + rc(read 2)[r2_stop:]                                             (:912-916)
#define ATR_MERGE_PREPEND 4          // rc(read 2) + read 1[r1_stop:]                                             (:917-922)

// MergeOverlapping.__call__ for one pair, up to the decision (the strings are put together by the caller)
ATR_HD void merge_pair(const unsigned char* __restrict__ r1, int len1, const unsigned char* __restrict__ r2, int len2,
                       int insert_matched, const MergeTables& tb, GCell* col, long stride, atr_merge_result* out) {
    atr_merge_result r;
    r.r2_start = r.r2_stop = r.r1_start = r.r1_stop = r.matches = r.errors = 0;
    r.min_overlap = 0; r.status = ATR_ST_NONE; r.action = 0;
    const int min_ov = (int)tb.minov[atr_min(len1, len2)];
    r.min_overlap = (uint16_t)min_ov;
    if (len1 < min_ov || len2 < min_ov) { *out = r; return; }                // :881-882
    for (int p = 0; p < len2; p++)
        if (tb.comp[r2[p]] == 0) { r.status = ATR_ST_KEYERROR; *out = r; return; }      // reverse_complement raises
    MergeAd ad; ad.min_overlap = 1; ad.thr_mul = tb.thr_mul;
    Best b;
    merge_locate(r1, len1, r2, len2, insert_matched ? (ATR_START_WITHIN_SEQ1 | ATR_STOP_WITHIN_SEQ2) : ATR_SEMIGLOBAL, ad,
                 tb.comp, col, stride, b);
    if (b.cost == len1 + len2) { *out = r; return; }                         // locate() returned None
    int start1 = 0, start2 = b.origin;
    if (b.origin < 0) { start1 = -b.origin; start2 = 0; }
    r.r2_start = (uint16_t)start1; r.r2_stop = (uint16_t)b.ref_stop;
    r.r1_start = (uint16_t)start2; r.r1_stop = (uint16_t)b.q_stop;
    r.matches = (uint16_t)b.matches; r.errors = (uint16_t)b.cost;
    if (b.matches < min_ov) { *out = r; return; }                            // :900: the pair is left alone
    r.status = ATR_ST_MATCH;
    if (start1 == 0 && b.ref_stop == len2) r.action = ATR_MERGE_KEEP1;
    else if (start2 == 0 && b.q_stop == len1) r.action = ATR_MERGE_TAKE2;
    else if (start2 > 0) r.action = ATR_MERGE_APPEND;
    else if (start1 > 0) r.action = ATR_MERGE_PREPEND;
    else r.status = ATR_ST_INVALID;                                          // :923-927 AtroposError
    *out = r;
}
