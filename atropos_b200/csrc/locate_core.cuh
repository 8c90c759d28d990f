// locate_core.cuh -- per-read alignment logic of the K1 kernels (one thread = one read).
//
// Re-designs Aligner.locate (reference: atropos/align/_align.pyx:266-491) for a GPU thread:
//   * K1a  k1a_locate():  adapter <= 64 nt; the whole DP column lives in registers as packed 32-bit
//          keys (atr_common.cuh), rows fully unrolled, every column evaluated in full and every cost
//          clamped to k+1 -- no data-dependent band (the reference's Ukkonen `last` cut-off and the
//          early exit on an exact hit do not change results: cells with cost > k never feed an
//          accepted alignment). One unsigned min3 implements the reference's tie order.
//   * K1g  gen_locate():  any adapter/read length and the byte-exact ASCII alphabet; column in
//          global scratch, banded like the reference. This is the path for escaped reads,
//          adapters > 64 nt and reads > 4000 nt.
//   * the fast funnel: sa_scan / sa_exact / sa_tail / sa_classify (Shift-And pre-filter), myers_filter
//          (exact bit-vector costs), k1d_band (banded 3-field DP on 16 diagonals), k1a_locate with a
//          column window. k1f_read() chains them for one read exactly as the kernels do.
// The functions are __host__ __device__ so tests/host_sim can run the very same code on the CPU
// build box (no GPU there); the product only ever calls them from the kernels in kernels.cu.
#pragma once
#include "atr_common.cuh"

template <class T> struct SignedOf;
template <> struct SignedOf<unsigned int> { typedef int type; };
template <> struct SignedOf<unsigned long> { typedef long type; };
template <> struct SignedOf<unsigned long long> { typedef long long type; };

struct Best {
    int matches, cost, origin, ref_stop, q_stop;
};

ATR_HD int atr_min(int a, int b) { return a < b ? a : b; }
ATR_HD int atr_max(int a, int b) { return a > b ? a : b; }
ATR_HD unsigned atr_umin(unsigned a, unsigned b) { return a < b ? a : b; }
ATR_HD uint32_t funnel_r32(uint32_t lo, uint32_t hi, unsigned shift) {   // (hi:lo) >> shift, shift in [0,32)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    return shift == 0 ? lo : ((lo >> shift) | (hi << (32 - shift)));
#endif
}
ATR_HD int atr_msb(unsigned x) {          // index of the highest set bit (x != 0)
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
ATR_HD int atr_ctz(unsigned x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// candidate acceptance: _align.pyx:443-449 (row m inside the loop) and :465-474 (last column)
template <class A>
ATR_HD void consider(const A& ad, Best& b, int cost, int origin, int matches, int ref_stop, int q_stop) {
    const int length = ref_stop + atr_min(origin, 0);
    if (length >= ad.min_overlap && cost <= (int)ad.thr_mul[length] &&
        (matches > b.matches || (matches == b.matches && cost < b.cost))) {
        b.matches = matches; b.cost = cost; b.origin = origin; b.ref_stop = ref_stop; b.q_stop = q_stop;
    }
}

// AdapterCutter._best_match's reduction over the adapters of a set (modifiers.py:116-121): the kernels run
// adapter after adapter in list order, a later adapter only replaces the stored match with strictly more matches.
// a record is one 16-byte transaction on the device (every record array the kernels write is 16-byte aligned: the
// library's own buffers, and atr_locate_batch_device checks the caller's)
ATR_HD void store_match(atr_match* out, const atr_match& r) {
#if defined(__CUDA_ARCH__)
    union { atr_match m; uint4 v; } u;
    u.m = r;
    *reinterpret_cast<uint4*>(out) = u.v;
#else
    *out = r;
#endif
}
ATR_HD atr_match load_match(const atr_match* p) {
#if defined(__CUDA_ARCH__)
    union { atr_match m; uint4 v; } u;
    u.v = *reinterpret_cast<const uint4*>(p);
    return u.m;
#else
    return *p;
#endif
}

template <class A>
ATR_HD void emit(const A& ad, const atr_match& r, atr_match* out) {
    if (ad.reduce) {
        const atr_match prev = load_match(out);
        if (prev.status >= ATR_ST_ESCAPED) return;      // ESCAPED / INVALID / KEYERROR are sticky: the host raises
        if (r.status == ATR_ST_NONE) return;
        if (r.status == ATR_ST_MATCH && prev.status == ATR_ST_MATCH && r.matches <= prev.matches) return;
    }
    store_match(out, r);
}

// the literal-hit result of the str.find / startswith / endswith shortcut (adapters/__init__.py:363-367)
template <class A>
ATR_HD void emit_exact(const A& ad, int pos, atr_match* out) {
    atr_match r;
    r.astart = 0; r.astop = (uint16_t)ad.m; r.rstart = (uint16_t)pos; r.rstop = (uint16_t)(pos + ad.m);
    r.matches = (uint16_t)ad.m; r.errors = 0; r.adapter = (int16_t)ad.adapter_index; r.status = ATR_ST_MATCH;
    emit(ad, r, out);
}

// Turn `best` into the result record: _align.pyx:476-491, then (match_to semantics) the post-filter
// of Adapter.match_to (adapters/__init__.py:384-398) and Match.__init__'s checks (align/__init__.py:85-88),
// then AdapterCutter._best_match's reduction over adapters (modifiers.py:116-121).
template <class A>
ATR_HD void finalize(const A& ad, const Best& b, int n, atr_match* out) {
    atr_match r;
    r.astart = r.astop = r.rstart = r.rstop = r.matches = r.errors = 0;
    r.adapter = -1;
    r.status = ATR_ST_NONE;
    if (b.cost != ad.m + n) {
        int start1 = 0, start2 = b.origin;
        if (b.origin < 0) { start1 = -b.origin; start2 = 0; }
        bool ok = true;
        bool invalid = false;
        if (ad.match_to) {
            const int size = b.ref_stop - start1;
            ok = size >= ad.min_overlap && b.cost <= (int)ad.thr_div[size];
            if (ok && ad.rmp_ok != nullptr) {
                const bool exact = ad.exact_bypass && b.cost == 0 && size == ad.m;     // str.find shortcut skips the gate
                if (!exact) ok = ad.rmp_ok[size * (ad.m + 1) + b.matches] != 0;
            }
            if (ok && size - b.cost <= 0) invalid = true;
        }
        if (ok) {
            r.astart = (uint16_t)start1; r.astop = (uint16_t)b.ref_stop;
            r.rstart = (uint16_t)start2; r.rstop = (uint16_t)b.q_stop;
            r.matches = (uint16_t)b.matches; r.errors = (uint16_t)b.cost;
            r.adapter = (int16_t)ad.adapter_index;
            r.status = invalid ? ATR_ST_INVALID : ATR_ST_MATCH;
        }
    }
    emit(ad, r, out);
}

// compare_prefixes / compare_suffixes result -> Best-free direct record (adapters/__init__.py:370-380)
template <class A>
ATR_HD void finalize_cmp(const A& ad, int n, int length, int matches, bool suffix, atr_match* out) {
    Best b;
    b.matches = matches; b.cost = length - matches; b.ref_stop = suffix ? ad.m : length;
    b.q_stop = suffix ? n : length;
    // origin encodes (start1, start2): compare_suffixes returns (m-length, m, n-length, n, ..)
    // at most one of the two starts is non-zero only if lengths are equal to min(m,n); carry both explicitly.
    atr_match r;
    r.astart = r.astop = r.rstart = r.rstop = r.matches = r.errors = 0;
    r.adapter = -1;
    r.status = ATR_ST_NONE;
    const int astart = suffix ? ad.m - length : 0, rstart = suffix ? n - length : 0;
    const int size = length;                                   // astop - astart
    bool ok = size >= ad.min_overlap && b.cost <= (int)ad.thr_div[size > 0 ? size : 0];
    if (ok && ad.rmp_ok != nullptr) {
        const bool exact = ad.exact_bypass && b.cost == 0 && size == ad.m;
        if (!exact) ok = ad.rmp_ok[size * (ad.m + 1) + matches] != 0;
    }
    if (ok) {
        r.astart = (uint16_t)astart; r.astop = (uint16_t)b.ref_stop; r.rstart = (uint16_t)rstart;
        r.rstop = (uint16_t)b.q_stop; r.matches = (uint16_t)matches; r.errors = (uint16_t)b.cost;
        r.adapter = (int16_t)ad.adapter_index;
        r.status = (size <= 0 || size - b.cost <= 0) ? ATR_ST_INVALID : ATR_ST_MATCH;
    }
    emit(ad, r, out);
}

// ---- K1a ----------------------------------------------------------------------------------

// uniform binary-search "tap" of register row m (m is a kernel parameter, so the branches never diverge)
template <int LO, int HI>
struct Tap {
    static ATR_HD unsigned get(const unsigned (&col)[ATR_K1A_MAXM + 1], int m) {
        constexpr int MID = (LO + HI) / 2;
        return m <= MID ? Tap<LO, MID>::get(col, m) : Tap<MID + 1, HI>::get(col, m);
    }
};
template <int I>
struct Tap<I, I> {
    static ATR_HD unsigned get(const unsigned (&col)[ATR_K1A_MAXM + 1], int) { return col[I]; }
};

ATR_HD unsigned k1a_read_code(const uint32_t* __restrict__ codes, int pos) {
    return (codes[pos >> 3] >> ((pos & 7) * 4)) & 15u;
}

// One DP column of K1a: updates col[0..m] in place for read code qc at column j.
template <bool AND_MODE>
ATR_HD void k1a_column(const AdapterK1a& ad, unsigned (&col)[ATR_K1A_MAXM + 1], unsigned qc, int j, bool start_in_query,
                       unsigned CLAMP, unsigned C_SUB, unsigned C_INS, unsigned C_DEL) {
    const int m = ad.m;
    unsigned diag = col[0];
    if (start_in_query) col[0] = k1a_key(0, j, 0);                          // _align.pyx:385-386
    else col[0] = atr_umin(col[0] + ((unsigned)ad.ic << ATR_COST_SHIFT), CLAMP);   // :387-388
#pragma unroll
    for (int g = 0; g < ATR_K1A_MAXM / 8; g++) {
        if (m > g * 8) {
#pragma unroll
            for (int r = 1; r <= 8; r++) {
                const int i = g * 8 + r;
                const unsigned left = col[i];
                unsigned t = atr_umin(atr_umin(left + C_DEL, col[i - 1] + C_INS), diag + C_SUB) & ATR_PRIO_CLEAR;
                const bool eq = AND_MODE ? ((((unsigned)ad.code[i - 1]) & qc) != 0u) : ((unsigned)ad.code[i - 1] == qc);
                unsigned nw = eq ? diag + 1u : t;                           // a match always takes the diagonal (:394-398)
                nw = atr_umin(nw, CLAMP);
                diag = left;
                col[i] = nw;
            }
        }
    }
}

// One read (window [lo, lo+n) of the packed read starting at word `codes`) against one adapter.
// c0/c1: evaluate only DP columns c0+1..c1 (the windowed DP stage k_wide; see myers_filter below). c0 < 0 = all.
template <bool AND_MODE>
ATR_HD void k1a_locate(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, Best& best,
                       int c0 = -1, int c1 = -1) {
    const int m = ad.m, k = ad.k;
    const bool start_in_ref = ad.flags & ATR_START_WITHIN_SEQ1, start_in_query = ad.flags & ATR_START_WITHIN_SEQ2;
    const bool stop_in_ref = ad.flags & ATR_STOP_WITHIN_SEQ1, stop_in_query = ad.flags & ATR_STOP_WITHIN_SEQ2;
    const unsigned CLAMP = (unsigned)(k + 1) << ATR_COST_SHIFT;
    const unsigned C_SUB = 1u << ATR_COST_SHIFT;
    const unsigned C_INS = ((unsigned)ad.ic << ATR_COST_SHIFT) | (1u << ATR_PRIO_SHIFT);
    const unsigned C_DEL = ((unsigned)ad.ic << ATR_COST_SHIFT) | (2u << ATR_PRIO_SHIFT);

    int max_n = n, min_n = 0;                                              // _align.pyx:315-321
    if (!start_in_query) max_n = atr_min(n, m + k);
    if (!stop_in_query) min_n = atr_max(0, n - m - k);
    const bool scan_last = max_n == n;                                     // :461
    // windowed evaluation: start at column c0 as if the read began there. Exact for every cell that can
    // lie on an accepted alignment ending in the window ("Exactness"): costs elsewhere are only
    // ever over-estimated. Only used for start_in_query && !start_in_ref flag sets.
    const bool windowed = c0 > min_n;
    if (c0 >= 0) { min_n = atr_max(min_n, c0); max_n = atr_min(max_n, c1); }

    unsigned col[ATR_K1A_MAXM + 1];
#pragma unroll
    for (int i = 0; i <= ATR_K1A_MAXM; i++) {                               // _align.pyx:333-352
        int cost, origin;
        if (windowed) { cost = i * ad.ic; origin = min_n; }
        else if (!start_in_ref && !start_in_query) { cost = atr_max(i, min_n) * ad.ic; origin = 0; }
        else if (start_in_ref && !start_in_query) { cost = min_n * ad.ic; origin = atr_min(0, min_n - i); }
        else if (!start_in_ref && start_in_query) { cost = i * ad.ic; origin = atr_max(0, min_n - i); }
        else { cost = atr_min(i, min_n) * ad.ic; origin = min_n - i; }
        col[i] = cost > k ? CLAMP : k1a_key(cost, origin, 0);
    }

    best.ref_stop = m; best.q_stop = n; best.cost = m + n; best.origin = 0; best.matches = 0;

    uint32_t w = 0;
    if (min_n < max_n) w = codes[(lo + min_n) >> 3];
#pragma unroll 1
    for (int j = min_n + 1; j <= max_n; j++) {
        const int pos = lo + j - 1;
        if ((pos & 7) == 0) w = codes[pos >> 3];
        unsigned qc = (w >> ((pos & 7) * 4)) & 15u;
        if (AND_MODE && ad.q_single_only) qc = (qc & (qc - 1)) ? 0u : qc;
        k1a_column<AND_MODE>(ad, col, qc, j, start_in_query, CLAMP, C_SUB, C_INS, C_DEL);
        if (stop_in_query) {                                                // :440-458 without the early break
            const unsigned c = Tap<1, ATR_K1A_MAXM>::get(col, m);
            if (c < CLAMP) consider(ad, best, k1a_cost(c), k1a_origin(c), k1a_matches(c), m, j);
        }
    }
    if (scan_last && max_n == n) {                                          // :461-474
        if (stop_in_ref) {
#pragma unroll
            for (int i = 1; i <= ATR_K1A_MAXM; i++) {
                if (i <= m) {
                    const unsigned c = col[i];
                    if (c < CLAMP) consider(ad, best, k1a_cost(c), k1a_origin(c), k1a_matches(c), i, n);
                }
            }
        } else {
            const unsigned c = Tap<1, ATR_K1A_MAXM>::get(col, m);
            if (c < CLAMP) consider(ad, best, k1a_cost(c), k1a_origin(c), k1a_matches(c), m, n);
        }
    }
}

// ---- filter stage (k_filter, k_refine): Myers/Hyyro bit-vector DP ------------------------------------------------------
// Exact unit-cost DP costs (not the tie-broken path). The adapter sits LEFT-ALIGNED in the word: row i is
// bit (WB - m + i - 1), so the bottom row m is always the sign bit; the bits below row 1 are "virtual rows"
// whose Peq bits are all ones and whose vertical deltas stay 0, i.e. they behave exactly like the free row 0.
// Bit r of Pv/Mv = vertical delta +1/-1 between the rows below/at that bit. The filter finds every cell
// the reference could accept,
//   J = { j : D[m][j] <= k }                                   (row m inside the loop; needs stop_in_query)
//   I = { i : D[i][n] <= floor(i*rate), i >= min_overlap }     (last column; all rows if stop_in_ref, else row m)
// and reports the range of DP diagonals (column - row) that alignments ending in them can touch:
// [dlo, dlo + width). false = no acceptable cell = the read has no match.
// Requires start_in_query && !start_in_ref && indel cost 1 (AdapterK1a.fused_ok).
struct FilterHit {
    int dlo, width;      // diagonal band for the banded kernel (K1d)
    int c0, c1;          // column window for the windowed register kernel (fallback when the band is too wide)
};

template <class WORD>
struct MyersState {
    WORD Pv, Mv;
    int score;
};

template <class WORD>
ATR_HD void myers_col(MyersState<WORD>& st, WORD Eq) {
    const WORD Pv = st.Pv, Mv = st.Mv;
    const WORD Xv = Eq | Mv;
    const WORD Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    WORD Ph = Mv | ~(Xh | Pv);
    WORD Mh = Pv & Xh;
    typedef typename SignedOf<WORD>::type SWORD;
    if ((SWORD)Ph < 0) st.score++;                     // bottom row = sign bit
    if ((SWORD)Mh < 0) st.score--;
    Ph <<= 1;                                          // row 0 (and the virtual rows) contribute horizontal delta 0
    Mh <<= 1;
    st.Pv = Mh | ~(Xv | Ph);
    st.Mv = Ph & Xv;
}

template <class WORD>
ATR_HD bool myers_filter(const AdapterK1a& ad, const WORD* __restrict__ peq, const uint32_t* __restrict__ codes, int lo, int n,
                         FilterHit& hit, int cstart = -1, int cstop = -1) {
    // cstart/cstop: evaluate only columns cstart+1..cstop, starting afresh at cstart (k_refine). The costs of
    // cells whose cheap alignments start at or after cstart are exact, all others only grow.
    const int m = ad.m, k = ad.k;
    const int WB = (int)(8 * sizeof(WORD));
    const int sh = WB - m;                             // first real row sits at bit sh
    const bool stop_in_ref = ad.flags & ATR_STOP_WITHIN_SEQ1, stop_in_query = ad.flags & ATR_STOP_WITHIN_SEQ2;
    const bool start_in_ref = ad.flags & ATR_START_WITHIN_SEQ1;
    int min_n = stop_in_query ? 0 : atr_max(0, n - m - k);
    const int n_full = n;
    if (cstart >= 0) { min_n = atr_max(min_n, cstart); n = atr_min(n, cstop); }
    MyersState<WORD> st;
    st.Pv = (sh == 0) ? ~(WORD)0 : (WORD)(~(WORD)0 << sh);   // column min_n: cost(i) = i (_align.pyx:345-348)
    st.Mv = 0;
    st.score = m;
    int jmin = 0x7fffffff, jmax = -1;
    if (start_in_ref) {
        // FRONT / ANYWHERE (:349-352 with min_n = 0): the first column costs 0 in every row, and an alignment
        // may start inside the adapter, so the bound on D[m][j] depends on j (AdapterK1a.thrJ). Plain loop.
        st.Pv = 0; st.score = 0;
        // columns 1..m one by one (the bound changes with j); from column m on the bound is thrJ[m] == k and the
        // word-wise loops below take over (FRONT / ANYWHERE always have stop_in_query)
        const int head = (ad.thrJ[m] == k && stop_in_query) ? atr_min(n, m) : n;
        for (int j = 1; j <= head; j++) {
            const int pos = lo + j - 1;
            myers_col(st, peq[(codes[pos >> 3] >> ((pos & 7) * 4)) & 15u]);
            if (st.score <= (int)ad.thrJ[j < m ? j : m]) { jmin = atr_min(jmin, j); jmax = j; }
        }
        min_n = head;
    }
    int j = min_n;                                     // columns done so far
    int pos = lo + min_n;                              // packed position of the next column's base
    const int pend = lo + n;
    if (start_in_ref) min_n = 0;
    // head: up to the next word boundary
    while (pos < pend && (pos & 7) != 0) {
        const unsigned qc = (codes[pos >> 3] >> ((pos & 7) * 4)) & 15u;
        myers_col(st, peq[qc]);
        j++; pos++;
        if (stop_in_query && st.score <= k) { jmin = atr_min(jmin, j); jmax = j; }
    }
    // body: whole words, 8 columns each, fully unrolled
    while (pos + 8 <= pend) {
        const uint32_t w = codes[pos >> 3];
#pragma unroll
        for (int t = 0; t < 8; t++) {
            myers_col(st, peq[(w >> (4 * t)) & 15u]);
            if (stop_in_query && st.score <= k) { jmin = atr_min(jmin, j + t + 1); jmax = j + t + 1; }
        }
        j += 8; pos += 8;
    }
    // tail
    if (pos < pend) {
        const uint32_t w = codes[pos >> 3];
        for (int t = 0; pos < pend; t++) {
            myers_col(st, peq[(w >> (4 * t)) & 15u]);
            j++; pos++;
            if (stop_in_query && st.score <= k) { jmin = atr_min(jmin, j); jmax = j; }
        }
    }
    // last column: D[i][n] = sum of the vertical deltas of rows 1..i
    int imin = 0, imax = 0;
    if (n == n_full) {
        int d = 0;
        const int first = stop_in_ref ? 1 : m;
        for (int i = 1; i <= m; i++) {
            d += (int)((st.Pv >> (sh + i - 1)) & (WORD)1) - (int)((st.Mv >> (sh + i - 1)) & (WORD)1);
            if (i >= first && i >= ad.min_overlap && d <= (int)ad.thr_mul[i]) { if (imin == 0) imin = i; imax = i; }
        }
    }
    if (jmax < 0 && imax == 0) return false;
    // end diagonals of the candidate cells: (m, j) -> j - m ; (i, n) -> n - i. An alignment of cost <= k
    // that ends on diagonal e stays within [e - k, e + k].
    int elo = 0x7fffffff, ehi = -0x7fffffff;
    if (jmax >= 0) { elo = jmin - m; ehi = jmax - m; }
    if (imax > 0) { elo = atr_min(elo, n - imax); ehi = atr_max(ehi, n - imin); }
    hit.dlo = elo - k;
    hit.width = (ehi - elo) + 2 * k + 1;
    hit.c0 = atr_max(min_n, elo - k);                 // first column any such alignment can touch in row 0
    hit.c1 = imax > 0 ? n : jmax;
    return true;
}

// ---- filter stage (k_filter_sa): Shift-And pre-filter + 32-bit tail Myers ---------------------------------------
// Pigeonhole: an alignment of the adapter's first sa_rows rows with <= k unit-cost errors contains one of the
// k+1 pieces verbatim, so a read without any verbatim piece has no row-m candidate and none in the last column
// below row sa_rows... The Shift-And automaton over all pieces costs ~7 instructions per column (vs ~31 for the
// 64-bit Myers). Candidates in the last column with row <= sa_rows are found exactly by a 32-bit Myers over the
// last sa_rows + k columns. Result classes:
//   0  nothing can match                       -> the read is finished (no match)
//   1  no piece hit, but last-column candidates -> band known exactly (dlo / width, or window if too wide)
//   2  piece hit(s)                             -> exact 64/32-bit Myers over columns [c0, c1] decides (k_refine)
//   3  all hits on one diagonal v and the whole adapter occurs there verbatim -> the result is known: this is
//      the reference's own str.find shortcut (adapters/__init__.py:351-367); an exact occurrence has the
//      maximal number of matches at cost 0, and being the only diagonal with hits it is the leftmost one
struct SaResult {
    int cls;
    int dlo, width;      // class 1
    int c0, c1;          // class 1 (window form) and class 2
    int v;               // class 3: read position of the exact occurrence
};

#ifndef ATR_SA_PAIR_TABLE
#define ATR_SA_PAIR_TABLE 1          // measured on B200: pair table 0.856 ms, single-base table 0.891 ms per 10 M reads
#endif
// eight columns (one packed word) of the Shift-And automaton; G = columns per hit-accumulation group (see sa_scan)
template <int G>
ATR_HD void sa_word(uint32_t w, unsigned& St, unsigned S0, unsigned E, const unsigned* __restrict__ sa_peq,
                    const unsigned long long* __restrict__ sa_pair, int j, int& hmin, int& hmax) {
#pragma unroll
    for (int g = 0; g < 8; g += G) {
        unsigned H = 0;
#pragma unroll
        for (int t = g; t < g + G; t += 2) {
#if ATR_SA_PAIR_TABLE
            const unsigned long long pr = sa_pair[(w >> (4 * t)) & 255u];
            const unsigned p0 = (unsigned)pr, p1 = (unsigned)(pr >> 32);
#else
            // Alternative kept for the record: 16 entries x 4 bytes sit in 16 different banks, so every load is one
            // wavefront, whereas the pair table's random accesses conflict (7 wavefronts per load measured, the
            // shared-memory pipe busy for 64 % of the kernel). It still loses: the kernel is bound by instruction
            // issue, and this form needs one more index extraction per column.
            const unsigned p0 = sa_peq[(w >> (4 * t)) & 15u], p1 = sa_peq[(w >> (4 * t + 4)) & 15u];
#endif
            St = ((St << 1) | S0) & p0;
            H |= (St & E) >> (t - g);
            St = ((St << 1) | S0) & p1;
            H |= (St & E) >> (t + 1 - g);
        }
        if (H) { hmin = atr_min(hmin, j + g - atr_msb(H)); hmax = atr_max(hmax, j + g - atr_ctz(H)); }
    }
}

// (a) Shift-And over all columns: range of hit diagonals and the automaton's final state
// sa_pair: 256-entry table indexed by a byte of the packed read (two bases): low word = Peq of the first base,
// high word = Peq of the second -- one 8-byte shared-memory load serves two columns.
ATR_HD void sa_scan(const AdapterK1a& ad, const unsigned* __restrict__ sa_peq, const unsigned long long* __restrict__ sa_pair,
                    const uint32_t* __restrict__ codes, int lo, int n, int& hmin, int& hmax, unsigned& st_final) {
    const unsigned S0 = ad.sa_start, E = ad.sa_end;
    unsigned St = 0;
    hmin = 0x7fffffff; hmax = -0x7fffffff;           // over hits: (column of the piece end) - (row of the piece end)
    int j = 0, pos = lo;
    const int pend = lo + n;
    while (pos < pend && (pos & 7) != 0) {
        const unsigned qc = (codes[pos >> 3] >> ((pos & 7) * 4)) & 15u;
        St = ((St << 1) | S0) & sa_peq[qc];
        j++; pos++;
        const unsigned hb = St & E;
        if (hb) { hmin = atr_min(hmin, j - 1 - atr_msb(hb)); hmax = atr_max(hmax, j - 1 - atr_ctz(hb)); }
    }
    // Whole words: the piece ends of G consecutive columns are folded into one word H, column t's shifted right by
    // t, so that bit b of H = a piece ending in row r1 at column j + t with r1 - t = b, i.e. on diagonal v = j - b:
    // one test and one msb / ctz per group instead of one per column (the hits sit in a few lanes of a warp, so
    // every instruction spent on them runs almost empty). G = 8 needs every piece to end in row >= 8 (bit >= 7).
    if (ad.sa_end & 0x7Fu) {
        while (pos + 8 <= pend) { sa_word<4>(codes[pos >> 3], St, S0, E, sa_peq, sa_pair, j, hmin, hmax); j += 8; pos += 8; }
    } else {
        while (pos + 8 <= pend) { sa_word<8>(codes[pos >> 3], St, S0, E, sa_peq, sa_pair, j, hmin, hmax); j += 8; pos += 8; }
    }
    if (pos < pend) {
        const uint32_t w = codes[pos >> 3];
        for (int t = 0; pos < pend; t++) {
            St = ((St << 1) | S0) & sa_peq[(w >> (4 * t)) & 15u];
            j++; pos++;
            const unsigned hb = St & E;
            if (hb) { hmin = atr_min(hmin, j - 1 - atr_msb(hb)); hmax = atr_max(hmax, j - 1 - atr_ctz(hb)); }
        }
    }
    st_final = St;
}

// (b) the reference's str.find shortcut: all hits on one diagonal v and the whole adapter verbatim there
ATR_HD bool sa_exact(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, int hmin, int hmax) {
    const int m = ad.m;
    if (!(hmax != -0x7fffffff && hmin == hmax && ad.exact_ok && hmin >= 0 && hmin + m <= n)) return false;
    const int q = lo + hmin;
    const unsigned sh = (unsigned)(q & 7) * 4u;
    bool same = true;
    for (int w = 0; w * 8 < m; w++) {                  // 8 bases per step
        const int rows = atr_min(8, m - 8 * w);        // adapter rows in this word
        const uint32_t w0 = codes[(q >> 3) + w];
        const bool need_hi = ((q & 7) + rows) > 8;     // the rows spill into the next read word
        const uint32_t w1 = need_hi ? codes[(q >> 3) + w + 1] : 0u;
        const uint32_t rd = funnel_r32(w0, w1, sh);
        const uint32_t mask = rows == 8 ? 0xFFFFFFFFu : ((1u << (4 * rows)) - 1u);
        uint32_t x;
        if (ad.and_mode) {                             // every nibble must share a bit
            x = rd & ad.apack[w];
            x |= x >> 1; x |= x >> 2;
            x = (~x) & 0x11111111u & mask;
        } else {
            x = (rd ^ ad.apack[w]) & mask;
        }
        same = same && x == 0u;
    }
    return same;
}

// (c) can there be a candidate in the last column with row <= sa_rows? (see adapter_build.hpp: tail gate)
ATR_HD bool sa_need_tail(const AdapterK1a& ad, int n, int hmax, unsigned st_final) {
    const bool stop_in_ref = ad.flags & ATR_STOP_WITHIN_SEQ1;
    if (!(stop_in_ref || ad.m <= ad.sa_rows)) return false;
    if (!ad.tail_gate_ok) return true;
    if (st_final & ad.tail_mask) return true;
    return hmax != -0x7fffffff && hmax >= n - ad.sa_rows - ad.k;
}

// (d) exact D[i][n] for rows i <= sa_rows from a Myers pass over the last sa_rows + k columns (rows left-aligned in a
// 32-bit word; 64-bit for the wide q-gram form, AdapterK1a.qg_wide)
template <class WORD>
ATR_HD void sa_tail_w(const AdapterK1a& ad, const WORD* __restrict__ tail_peq, const uint32_t* __restrict__ codes, int lo, int n,
                      int& imin, int& imax) {
    const int m = ad.m, k = ad.k, mp = ad.sa_rows;
    const int WB = (int)(8 * sizeof(WORD));
    const bool stop_in_ref = ad.flags & ATR_STOP_WITHIN_SEQ1;
    const int pend = lo + n;
    imin = 0; imax = 0;
    const int sh = WB - mp;
    MyersState<WORD> st;
    st.Pv = (sh == 0) ? ~(WORD)0 : (WORD)(~(WORD)0 << sh);
    st.Mv = 0; st.score = mp;
    // start on a word boundary at or before column n - mp - k (an earlier start is always safe)
    int p = atr_max(lo, (lo + atr_max(0, n - mp - k)) & ~7);
    while (p < pend && (p & 7) != 0) { myers_col(st, tail_peq[(codes[p >> 3] >> ((p & 7) * 4)) & 15u]); p++; }
    while (p + 8 <= pend) {
        const uint32_t w = codes[p >> 3];
#pragma unroll
        for (int t = 0; t < 8; t++) myers_col(st, tail_peq[(w >> (4 * t)) & 15u]);
        p += 8;
    }
    if (p < pend) {
        const uint32_t w = codes[p >> 3];
        for (int t = 0; p < pend; t++, p++) myers_col(st, tail_peq[(w >> (4 * t)) & 15u]);
    }
    // D[i][n] = running sum of the vertical deltas; rows right-aligned so that the shifts are static
    const WORD pv = st.Pv >> sh, mv = st.Mv >> sh;
    const int row_lo = atr_max(stop_in_ref ? 1 : m, ad.min_overlap);
    int d = 0;
#pragma unroll
    for (int i = 1; i <= WB; i++) {
        d += (int)((pv >> (i - 1)) & (WORD)1) - (int)((mv >> (i - 1)) & (WORD)1);
        if (i >= row_lo && i <= mp && d <= (int)ad.thr_mul[i]) { imin = imin == 0 ? i : imin; imax = i; }
    }
}
ATR_HD void sa_tail(const AdapterK1a& ad, const unsigned* __restrict__ tail_peq, const uint32_t* __restrict__ codes, int lo, int n,
                    int& imin, int& imax) {
    sa_tail_w<unsigned>(ad, tail_peq, codes, lo, n, imin, imax);
}

// (e) classes 0 / 1 / 2 from the hit range and the last-column rows
ATR_HD void sa_classify(const AdapterK1a& ad, int lo, int n, int hmin, int hmax, int imin, int imax, SaResult& res) {
    const int m = ad.m, k = ad.k;
    bool have_hit = hmax != -0x7fffffff;
    // Which hits can belong to a candidate? A row-m candidate (m, j), j <= n, passes through a hit diagonal
    // v <= j - m + k; a last-column candidate (i, n) with i > sa_rows (it contains every piece) through one with
    // v <= n - i + k; rows i <= sa_rows of the last column are the tail pass's (imin..imax), hit or not. Hits further
    // right -- a piece of an adapter that sticks out of the read -- bound nothing.
    if (have_hit) {
        const int vmax = (m > ad.sa_rows ? n - ad.sa_rows - 1 : n - m) + k;
        if (hmin > vmax) have_hit = false;
        else hmax = atr_min(hmax, vmax);
    }
    // a last-column candidate in row i has at most thr_mul[i] <= thr_mul[imax] errors, hence as many indels
    const int t = imax > 0 ? (int)ad.thr_mul[imax] : 0;
    if (!have_hit && imax == 0) { res.cls = 0; return; }
    if (!have_hit) {
        res.cls = 1;
        res.dlo = (n - imax) - t;
        res.width = (imax - imin) + 2 * t + 1;
        res.c0 = atr_max(0, n - imax - k);
        res.c1 = n;
        return;
    }
    // a verbatim piece ending at (row r1, column jh), v = jh - r1: the alignment through it starts in row 0 at a
    // column >= v - k and reaches row m at a column in [v + m - k, v + m + k]
    res.cls = 2;
    // An accepted alignment of the whole adapter (cost <= k) contains one of the pieces verbatim (pigeonhole), i.e. it
    // passes through a hit diagonal v in [hmin, hmax], and with at most k indels in total every cell of it lies
    // within [v - k, v + k]. The reference's own path to such a cell is one of these, and inside that band the DP
    // sees it with its exact predecessors (cells outside can only be worse, and a worse neighbour never wins a
    // tie), so the band [hmin - k, hmax + k] (plus the last-column candidates' diagonals) is exact: if it fits the
    // banded kernel the read skips k_refine.
    {
        int blo = hmin - k, bhi = hmax + k;
        if (imax > 0) { blo = atr_min(blo, n - imax - t); bhi = atr_max(bhi, n - imin + t); }
        res.dlo = blo;
        res.width = bhi - blo + 1;
    }
    int c0 = hmin - k - 1;
    int c1 = hmax + m + k + 1;
    if (imax > 0) { c0 = atr_min(c0, n - imax - k - 1); c1 = n; }
    // widen to word boundaries of the packed read (a wider range is always safe; it keeps k_refine in its
    // unrolled whole-word loop)
    c0 = atr_max(0, c0); c1 = atr_min(n, c1);
    res.c0 = atr_max(0, ((lo + c0) & ~7) - lo);
    res.c1 = atr_min(n, ((lo + c1 + 7) & ~7) - lo);
}

// the whole stage for one read (host simulator; the kernel interleaves a block-level compaction before (d))
ATR_HD void sa_filter(const AdapterK1a& ad, const unsigned* __restrict__ sa_peq, const unsigned* __restrict__ tail_peq,
                      const uint32_t* __restrict__ codes, int lo, int n, SaResult& res) {
    int hmin, hmax, imin = 0, imax = 0;
    unsigned st_final;
    unsigned long long sa_pair[256];
    for (int b = 0; b < 256; b++) sa_pair[b] = (unsigned long long)sa_peq[b & 15] | ((unsigned long long)sa_peq[b >> 4] << 32);
    sa_scan(ad, sa_peq, sa_pair, codes, lo, n, hmin, hmax, st_final);
    if (sa_exact(ad, codes, lo, n, hmin, hmax)) { res.cls = 3; res.v = hmin; return; }
    if (sa_need_tail(ad, n, hmax, st_final)) sa_tail(ad, tail_peq, codes, lo, n, imin, imax);
    sa_classify(ad, lo, n, hmin, hmax, imin, imax, res);
}

// ---- first stage for unanchored 5' adapters (AdapterK1a.sa_front; k_filter_front) ---------------------------------
// Candidates are row-m cells (m, j) only. (H) j <= m + k: found exactly by the Myers pass with the all-zero first column
// (an alignment may start inside the adapter) over the first m + k columns, bound thrJ[j]. (F) j > m + k: the alignment
// covers all m rows (one that starts inside the adapter starts in column 0 and reaches at most column m + k), so one of the
// k + 1 pieces is verbatim on a diagonal v with |v - (j - m)| <= k, -k <= v <= n - m + k. The diagonals any accepted
// alignment touches are [e - k, e + k] around the end diagonals e = j - m of (H) and [v - k, v + k] around the hits of (F):
// the same FilterHit as myers_filter's, a superset of its range (a hit need not lead to a candidate; the DP decides).
// peq: the adapter left-aligned in 32 bits with the virtual rows set (as for myers_filter<unsigned>).
ATR_HD bool front_filter(const AdapterK1a& ad, const unsigned* __restrict__ sa_peq, const unsigned long long* __restrict__ sa_pair,
                         const unsigned* __restrict__ peq, const uint32_t* __restrict__ codes, int lo, int n, FilterHit& hit) {
    const int m = ad.m, k = ad.k;
    int hmin, hmax;
    unsigned st_final;
    sa_scan(ad, sa_peq, sa_pair, codes, lo, n, hmin, hmax, st_final);
    // (H)
    MyersState<unsigned> st;
    st.Pv = 0; st.Mv = 0; st.score = 0;
    int jmin = 0x7fffffff, jmax = -1;
    const int head = atr_min(n, m + k);
    for (int j = 1; j <= head; j++) {
        const int pos = lo + j - 1;
        myers_col(st, peq[(codes[pos >> 3] >> ((pos & 7) * 4)) & 15u]);
        if (st.score <= (int)ad.thrJ[j < m ? j : m]) { jmin = atr_min(jmin, j); jmax = j; }
    }
    // (F)
    int vlo = 0x7fffffff, vhi = -0x7fffffff;
    if (hmax != -0x7fffffff) {
        vlo = atr_max(hmin, -k); vhi = atr_min(hmax, n - m + k);
        if (vlo > vhi) { vlo = 0x7fffffff; vhi = -0x7fffffff; }
    }
    if (jmax < 0 && vhi == -0x7fffffff) return false;
    int elo = 0x7fffffff, ehi = -0x7fffffff, c1 = 0;
    if (jmax >= 0) { elo = jmin - m; ehi = jmax - m; c1 = jmax; }
    if (vhi != -0x7fffffff) { elo = atr_min(elo, vlo); ehi = atr_max(ehi, vhi); c1 = atr_max(c1, atr_min(n, vhi + m + k)); }
    hit.dlo = elo - k;
    hit.width = (ehi - elo) + 2 * k + 1;
    hit.c0 = atr_max(0, elo - k);
    hit.c1 = c1;
    return true;
}

// ---- DP stage for narrow bands (k_band): K1d, banded DP along diagonals ------------------------------------------
// B[d] = cell (i, i + dlo + d) of the current row i, d = 0..W-1, as K1a packed keys. Rows run 1..m in a
// rolled loop (the adapter base of a row is warp-uniform), the W diagonals are unrolled in registers:
// diag = B[d] (old), up = B[d+1] (old), left = B[d-1] (new); cells outside the band count as dead.
// Columns left of the first DP column (j <= min_n, including j <= 0) are "virtual": free row 0 and a base
// that matches nothing. With unit indel cost their cells reproduce the reference's first column exactly
// (cost i, 0 matches); their origins may come out below 0 / below the true max(0, min_n - i) only where the
// true origin is 0, hence the clamp at the end (these flag sets never have negative origins).
template <int W> struct BandWin { typedef unsigned long long type; };      // W read codes of 4 bits
template <> struct BandWin<8> { typedef unsigned int type; };
template <bool AND_MODE, int W, bool SIR = false>
ATR_HD void k1d_band(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, int dlo, Best& best) {
    // SIR (start_in_ref: FRONT / ANYWHERE adapters): the first DP column costs 0 in every row with origin -i
    // (:349-352). Modelled by bases left of the read that match EVERYTHING under a free row 0: cell (i, 0) comes
    // out as (cost 0, origin -i, matches i); the i surplus matches are taken off at the end (matches + min(origin, 0)),
    // comparisons never look at matches. Negative origins are real here (the alignment starts inside the adapter).
    const int m = ad.m, k = ad.k;
    const bool stop_in_ref = ad.flags & ATR_STOP_WITHIN_SEQ1, stop_in_query = ad.flags & ATR_STOP_WITHIN_SEQ2;
    const int min_n = stop_in_query ? 0 : atr_max(0, n - m - k);
    const unsigned DEAD = (unsigned)(k + 1) << ATR_COST_SHIFT;
    const unsigned C_SUB = 1u << ATR_COST_SHIFT;
    const unsigned C_INS = (1u << ATR_COST_SHIFT) | (1u << ATR_PRIO_SHIFT);
    const unsigned C_DEL = (1u << ATR_COST_SHIFT) | (2u << ATR_PRIO_SHIFT);
    const unsigned nomatch = (unsigned)ad.nomatch;
    // No per-cell clamp here: a cell costs at most one more than the cell above it on its diagonal, so after
    // <= 64 rows the 8-bit cost field holds at most (k+1) + 64 <= 191. Cells with cost > k are dead whatever
    // their exact value.

    unsigned B[W];
#pragma unroll
    for (int d = 0; d < W; d++) {                      // row 0: cost 0, origin j (:385-386); no column beyond n
        const int j = dlo + d;
        B[d] = j > n ? DEAD : k1a_key(0, atr_max(j, -ATR_ORG_BIAS), 0);
    }
    // Sliding window of the W read codes of the current row: nibble d = base of column i + dlo + d.
    //  * left of the first DP column: a code that matches nothing (see above);
    //  * right of the read end: a base that matches EVERYTHING (vm marks those nibbles). A match always takes
    //    the diagonal, so cell (i, n) travels unchanged (plus one match per step) down its diagonal to row m:
    //    the reference's last-column candidates (:461-474) are read off row m at columns n+1.., no per-row tap.
    typedef typename BandWin<W>::type WIN;
    const WIN ONES = (WIN)0x1111111111111111ull;
    WIN win = 0, vm = 0;
    auto base_at = [&](int p) -> unsigned {            // p = 0-based position in the (windowed) read
        if (p < min_n) return SIR ? 0u : nomatch;
        if (p >= n) return 0u;
        const int q = lo + p;
        unsigned c = (codes[q >> 3] >> ((q & 7) * 4)) & 15u;
        if (AND_MODE && ad.q_single_only) c = (c & (c - 1)) ? 0u : c;
        return c;
    };
#pragma unroll 1
    for (int d = 0; d < W - 1; d++) {                  // row 1 needs columns 1+dlo .. W+dlo -> positions dlo .. dlo+W-1
        win |= (WIN)base_at(dlo + d) << (4 * (d + 1));
        if (dlo + d >= n || (SIR && dlo + d < 0)) vm |= (WIN)0xF << (4 * (d + 1));
    }
    // Rows beyond R = n - dlo have their whole band right of the read end: pure forced-match propagation, which
    // the candidate extraction below accounts for in closed form (partial adapters at the read end stop early).
    const int R = atr_max(0, atr_min(m, n - dlo));
#pragma unroll 1
    for (int i = 1; i <= R; i++) {
        const int pnew = i + dlo + W - 2;
        win = (win >> 4) | ((WIN)base_at(pnew) << (4 * (W - 1)));
        vm = (vm >> 4) | ((pnew >= n || (SIR && pnew < 0)) ? ((WIN)0xF << (4 * (W - 1))) : (WIN)0);
        const unsigned a = (unsigned)ad.code[i - 1];
        // per-nibble (mis)match flags for the whole row at once
        WIN x;
        if (AND_MODE) x = (win & (ONES * a)) | vm;
        else x = (win ^ (ONES * a)) & ~vm;
        x |= x >> 1; x |= x >> 2;                      // bit 4d set <=> nibble d non-zero
        const unsigned xl = (unsigned)x, xh = (unsigned)((unsigned long long)x >> 32);
        unsigned left = DEAD;                          // cell (i, i + dlo - 1): outside the band
#pragma unroll
        for (int d = 0; d < W; d++) {
            const unsigned diag = B[d];
            const unsigned up = (d + 1 < W) ? B[d + 1] : DEAD;
            const unsigned nz = ((d < 8 ? xl : xh) >> (4 * (d & 7))) & 1u;
            const bool eq = AND_MODE ? (nz != 0u) : (nz == 0u);
            const unsigned t = atr_umin(atr_umin(left + C_DEL, up + C_INS), diag + C_SUB) & ATR_PRIO_CLEAR;
            const unsigned nw = eq ? diag + 1u : t;
            B[d] = nw;
            left = nw;
        }
    }
    Best bl;                                           // last-column candidates, merged after the row-m ones
    bl.ref_stop = m; bl.q_stop = n; bl.cost = m + n; bl.origin = 0; bl.matches = 0;
    best = bl;
    if (stop_in_query && R == m) {                     // row m, columns in ascending order (:440-458)
#pragma unroll
        for (int d = 0; d < W; d++) {
            const int j = m + dlo + d;
            const unsigned c = B[d];
            if (j > min_n && j <= n && k1a_cost(c) <= k) {
                const int org = SIR ? k1a_origin(c) : atr_max(k1a_origin(c), 0);
                consider(ad, best, k1a_cost(c), org, k1a_matches(c) + (SIR ? atr_min(org, 0) : 0), m, j);
            }
        }
    }
    {                                                  // last column, rows ascending = virtual columns descending
        const int first_i = stop_in_ref ? 1 : m;
#pragma unroll
        for (int d = W - 1; d >= 0; d--) {
            const int j = R + dlo + d;                 // B[d] = cell (R, j)
            const int i = R - (j - n);                 // cell (i, n) arrived here after j - n forced matches
            const unsigned c = B[d];
            if (j >= n && i >= first_i && i >= 1 && k1a_cost(c) <= k) {
                const int org = SIR ? k1a_origin(c) : atr_max(k1a_origin(c), 0);
                consider(ad, bl, k1a_cost(c), org, k1a_matches(c) - (j - n) + (SIR ? atr_min(org, 0) : 0), i, n);
            }
        }
    }
    // the reference scans the last column after all in-loop candidates; replacement needs a strictly better key
    if (bl.cost != m + n && (bl.matches > best.matches || (bl.matches == best.matches && bl.cost < best.cost))) best = bl;
}

// compare_prefixes / compare_suffixes on packed codes (anchored adapters with indels off)
template <bool AND_MODE>
ATR_HD int k1a_compare(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, bool suffix, int& length) {
    length = atr_min(ad.m, n);
    int matches = 0;
    for (int t = 0; t < length; t++) {
        const int ai = suffix ? ad.m - 1 - t : t;
        const int pos = lo + (suffix ? n - 1 - t : t);
        unsigned qc = k1a_read_code(codes, pos);
        if (AND_MODE && ad.q_single_only) qc = (qc & (qc - 1)) ? 0u : qc;
        const unsigned a = (unsigned)ad.code[ai];          // uniform index: one constant-bank load
        matches += AND_MODE ? ((a & qc) != 0u) : (a == qc);
    }
    return matches;
}

// ---- K1g: general byte-exact kernel ---------------------------------------------------------

struct GCell { int cost, pay; };     // pay = (origin + ATR_G_ORG_BIAS) << 16 | matches

ATR_HD int g_pay(int origin, int matches) { return ((origin + ATR_G_ORG_BIAS) << 16) | matches; }
ATR_HD int g_origin(int pay) { return (int)((unsigned)pay >> 16) - ATR_G_ORG_BIAS; }
ATR_HD int g_matches(int pay) { return pay & 0xFFFF; }

ATR_HD unsigned gen_qcode(const AdapterGen& ad, const AtrTables& tb, unsigned char ch) {
    return ad.q_table == 0 ? ch : (ad.q_table == 1 ? tb.iupac[ch] : tb.acgt[ch]);
}

// col: this thread's DP column, element i at col[i * stride].
ATR_HD unsigned char gen_char(const unsigned char* __restrict__ read, int p, int fold_case) {
    unsigned char c = read[p];
    if (fold_case && c >= 'a' && c <= 'z') c = (unsigned char)(c - 32);      // str.upper() (adapters/__init__.py:349)
    return c;
}

ATR_HD void gen_locate(const AdapterGen& ad, const AtrTables& tb, const unsigned char* __restrict__ read, int n,
                       int fold_case, GCell* col, long stride, Best& best) {
    const int m = ad.m, k = ad.k, ic = ad.ic;
    const bool start_in_ref = ad.flags & ATR_START_WITHIN_SEQ1, start_in_query = ad.flags & ATR_START_WITHIN_SEQ2;
    const bool stop_in_ref = ad.flags & ATR_STOP_WITHIN_SEQ1, stop_in_query = ad.flags & ATR_STOP_WITHIN_SEQ2;
    int max_n = n, min_n = 0;
    if (!start_in_query) max_n = atr_min(n, m + k);
    if (!stop_in_query) min_n = atr_max(0, n - m - k);
    const int dead = k + 1;
    for (int i = 0; i <= m; i++) {
        long cost; int origin;                      // i * ic can exceed int for ic = 100000 and long adapters
        if (!start_in_ref && !start_in_query) { cost = (long)atr_max(i, min_n) * ic; origin = 0; }
        else if (start_in_ref && !start_in_query) { cost = (long)min_n * ic; origin = atr_min(0, min_n - i); }
        else if (!start_in_ref && start_in_query) { cost = (long)i * ic; origin = atr_max(0, min_n - i); }
        else { cost = (long)atr_min(i, min_n) * ic; origin = min_n - i; }
        GCell c; c.cost = cost > k ? dead : (int)cost; c.pay = g_pay(origin, 0);
        col[i * stride] = c;
    }
    best.ref_stop = m; best.q_stop = n; best.cost = m + n; best.origin = 0; best.matches = 0;
    int last = start_in_ref ? m : atr_min(m, k + 1);                        // Ukkonen band (:366-368)
    for (int j = min_n + 1; j <= max_n; j++) {
        GCell diag = col[0];
        GCell up = diag;
        if (start_in_query) up.pay = g_pay(j, 0);
        else { long c0 = (long)j * ic; up.cost = c0 > k ? dead : (int)c0; }
        col[0] = up;
        const unsigned qc = gen_qcode(ad, tb, gen_char(read, j - 1, fold_case));
        for (int i = 1; i <= last; i++) {
            const GCell left = col[i * stride];
            const unsigned a = ad.ref[i - 1];
            const bool eq = ad.and_mode ? ((a & qc) != 0u) : (a == qc);
            GCell nw;
            if (eq) { nw.cost = diag.cost; nw.pay = diag.pay + 1; }
            else {
                const int c_sub = diag.cost + 1, c_del = left.cost + ic, c_ins = up.cost + ic;
                if (c_sub <= c_del && c_sub <= c_ins) { nw.cost = c_sub; nw.pay = diag.pay; }
                else if (c_ins <= c_del) { nw.cost = c_ins; nw.pay = up.pay; }
                else { nw.cost = c_del; nw.pay = left.pay; }
            }
            if (nw.cost > k) nw.cost = dead;
            diag = left;
            col[i * stride] = nw;
            up = nw;
        }
        while (last >= 0 && col[last * stride].cost > k) last--;           // :433-439
        if (last < m) {
            // the row entering the band holds a value > k (initial, or left behind when `last` shrank);
            // like the reference's stale cells only its being > k matters
            last++;
        } else if (stop_in_query) {
            const GCell c = col[m * stride];
            consider(ad, best, c.cost, g_origin(c.pay), g_matches(c.pay), m, j);
        }
    }
    if (max_n == n) {
        for (int i = stop_in_ref ? 0 : m; i <= m; i++) {
            const GCell c = col[i * stride];
            if (c.cost <= k) consider(ad, best, c.cost, g_origin(c.pay), g_matches(c.pay), i, n);
        }
    }
}

ATR_HD int gen_compare(const AdapterGen& ad, const AtrTables& tb, const unsigned char* __restrict__ read, int n,
                       int fold_case, bool suffix, int& length) {
    length = atr_min(ad.m, n);
    int matches = 0;
    for (int t = 0; t < length; t++) {
        const unsigned a = ad.ref[suffix ? ad.m - 1 - t : t];
        const unsigned qc = gen_qcode(ad, tb, gen_char(read, suffix ? n - 1 - t : t, fold_case));
        matches += ad.and_mode ? ((a & qc) != 0u) : (a == qc);
    }
    return matches;
}

// ---- literal search (only for need_find adapters, see adapter_build.hpp) -----------------------
ATR_HD int k1a_find(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n) {
    const int m = ad.m;
    if (m > n) return -1;
    int first = 0, last = n - m;
    if (ad.flags == ATR_STOP_WITHIN_SEQ2) last = 0;                 // PREFIX: startswith
    else if (ad.flags == ATR_START_WITHIN_SEQ2) first = n - m;      // SUFFIX: endswith
    for (int p = first; p <= last; p++) {
        int t = 0;
        while (t < m && (unsigned)ad.lit[t] == k1a_read_code(codes, lo + p + t)) t++;
        if (t == m) return p;
    }
    return -1;
}

ATR_HD int gen_find(const AdapterGen& ad, const unsigned char* __restrict__ read, int n, int fold_case) {
    const int m = ad.m;
    if (m > n) return -1;
    int first = 0, last = n - m;
    if (ad.flags == ATR_STOP_WITHIN_SEQ2) last = 0;
    else if (ad.flags == ATR_START_WITHIN_SEQ2) first = n - m;
    for (int p = first; p <= last; p++) {
        int t = 0;
        while (t < m && ad.lit[t] == gen_char(read, p + t, fold_case)) t++;
        if (t == m) return p;
    }
    return -1;
}

// ---- one read against one adapter, start to finish ---------------------------------------------
template <bool AND_MODE>
ATR_HD void k1a_read(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, atr_match* out) {
    if (ad.need_find) {
        const int pos = k1a_find(ad, codes, lo, n);
        if (pos >= 0) { emit_exact(ad, pos, out); return; }
    }
    if (ad.cmp_only) {
        int length;
        const int matches = k1a_compare<AND_MODE>(ad, codes, lo, n, ad.cmp_only == 2, length);
        finalize_cmp(ad, n, length, matches, ad.cmp_only == 2, out);
    } else {
        Best b;
        k1a_locate<AND_MODE>(ad, codes, lo, n, b);
        finalize(ad, b, n, out);
    }
}

ATR_HD void gen_read(const AdapterGen& ad, const AtrTables& tb, const unsigned char* __restrict__ read, int n,
                     int fold_case, GCell* col, long stride, atr_match* out) {
    if (ad.need_find) {
        const int pos = gen_find(ad, read, n, fold_case);
        if (pos >= 0) { emit_exact(ad, pos, out); return; }
    }
    if (ad.cmp_only) {
        int length;
        const int matches = gen_compare(ad, tb, read, n, fold_case, ad.cmp_only == 2, length);
        finalize_cmp(ad, n, length, matches, ad.cmp_only == 2, out);
    } else {
        Best b;
        gen_locate(ad, tb, read, n, fold_case, col, stride, b);
        finalize(ad, b, n, out);
    }
}

// ---- anchored adapters outside the funnel (k_filter_anchor): fixed-position pigeonhole --------------------------
// PREFIX (flags == stop_in_query): the alignment runs from cell (0, 0) to (m, j), j <= m + k, cost <= k (:315-321,
// :333-352 with neither start flag, candidates only in row m). SUFFIX (flags == start_in_query): from row 0 to cell
// (m, n) exactly (:461-474 with stop_in_ref unset). Either way the whole adapter is aligned with at most k edits, each
// costing >= 1 whatever the indel cost, so of k + 1 pieces one is matched verbatim, and with at most k indels it sits
// within k columns of its anchored position. A read without such a piece cannot match; the others take the full
// register DP (k1a_read). Necessary condition only -- exactness comes from the DP.
ATR_HD bool anchor_piece_equal(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int q, int row, int len) {
    for (int w = 0; w * 8 < len; w++) {
        const int rows = atr_min(8, len - 8 * w);
        const int qq = q + 8 * w, rr = row + 8 * w;
        const uint32_t r0 = codes[qq >> 3];
        const uint32_t r1 = ((qq & 7) + rows > 8) ? codes[(qq >> 3) + 1] : 0u;
        const uint32_t rd = funnel_r32(r0, r1, (unsigned)(qq & 7) * 4u);
        const uint32_t a0 = ad.apack[rr >> 3];
        const uint32_t a1 = ((rr & 7) + rows > 8) ? ad.apack[(rr >> 3) + 1] : 0u;
        const uint32_t ap = funnel_r32(a0, a1, (unsigned)(rr & 7) * 4u);
        const uint32_t mask = rows == 8 ? 0xFFFFFFFFu : ((1u << (4 * rows)) - 1u);
        uint32_t x;
        if (ad.and_mode) {                             // every nibble must share a bit
            x = rd & ap;
            x |= x >> 1; x |= x >> 2;
            x = (~x) & 0x11111111u & mask;
        } else {
            x = (rd ^ ap) & mask;
        }
        if (x != 0u) return false;
    }
    return true;
}

ATR_HD bool anchor_filter(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n) {
    const int m = ad.m, k = ad.k, pieces = k + 1;
    if (m < pieces) return true;                       // degenerate: empty pieces prove nothing
    const int base = (ad.flags == ATR_START_WITHIN_SEQ2) ? n - m : 0;
    int row = 0;
    for (int pc = 0; pc < pieces; pc++) {
        const int len = m / pieces + (pc < m % pieces ? 1 : 0);
        for (int d = -k; d <= k; d++) {
            const int p0 = base + row + d;             // read position of the piece's first base
            if (p0 < 0 || p0 + len > n) continue;
            if (anchor_piece_equal(ad, codes, lo + p0, row, len)) return true;
        }
        row += len;
    }
    return false;
}

template <bool AND_MODE>
ATR_HD void anchor_read(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, atr_match* out);

// ---- the whole funnel for one read (what the kernels do, minus the compaction between the stages) ----
#define ATR_K1D_W 16
// qgram_core.cuh: the q-gram sampling form of the first stage (AdapterK1a.qg_ok), same result classes
ATR_HD void qg_filter(const AdapterK1a& ad, const unsigned* __restrict__ tail_peq, const uint32_t* __restrict__ codes, int wlimit,
                      int lo, int n, SaResult& res);
template <class WORD, bool AND_MODE>
ATR_HD void k1f_read(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, atr_match* out, int* path = nullptr) {
    Best b;
    b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
    FilterHit hit;
    const int WB = (int)(8 * sizeof(WORD)), sh = WB - ad.m;
    WORD peq[16];
    for (int c = 0; c < 16; c++) peq[c] = (WORD)(((WORD)ad.peq[c] << sh) | (sh ? (((WORD)1 << sh) - 1) : 0));
    if (path) *path = 0;
    bool have = false;
    if (ad.sa_ok || ad.qg_ok) {
        unsigned sa_peq[16], tail_peq[16];
        const int mp = ad.sa_rows < 32 ? ad.sa_rows : 32, sh32 = 32 - mp;
        for (int c = 0; c < 16; c++) {
            const unsigned low = (unsigned)(ad.peq[c] & (mp >= 32 ? 0xFFFFFFFFull : ((1ull << mp) - 1)));
            sa_peq[c] = low;
            tail_peq[c] = (sh32 ? (low << sh32) | ((1u << sh32) - 1u) : low);
        }
        SaResult sr;
        if (ad.qg_ok) qg_filter(ad, tail_peq, codes, (lo + n + 7) >> 3, lo, n, sr);
        else sa_filter(ad, sa_peq, tail_peq, codes, lo, n, sr);
        if (sr.cls == 3) {
            b.matches = ad.m; b.cost = 0; b.origin = sr.v; b.ref_stop = ad.m; b.q_stop = sr.v + ad.m;
            if (path) *path = 5;
        }
        else if (sr.cls == 1) { have = true; hit.dlo = sr.dlo; hit.width = sr.width; hit.c0 = sr.c0; hit.c1 = sr.c1; }
        else if (sr.cls == 2) {
            if (ad.band_ok && sr.width <= ATR_K1D_W) { have = true; hit.dlo = sr.dlo; hit.width = sr.width; hit.c0 = sr.c0; hit.c1 = sr.c1; }
            else have = myers_filter<WORD>(ad, peq, codes, lo, n, hit, sr.c0, sr.c1);
        }
    } else if (ad.sa_front) {
        unsigned sa_peq[16], peq32[16];
        const int mp = ad.sa_rows, sh32 = 32 - ad.m;
        for (int c = 0; c < 16; c++) {
            sa_peq[c] = (unsigned)(ad.peq[c] & (mp >= 32 ? 0xFFFFFFFFull : ((1ull << mp) - 1)));
            peq32[c] = (unsigned)(((unsigned)ad.peq[c] << sh32) | (sh32 ? ((1u << sh32) - 1u) : 0u));
        }
        unsigned long long sa_pair[256];
        for (int bb = 0; bb < 256; bb++) sa_pair[bb] = (unsigned long long)sa_peq[bb & 15] | ((unsigned long long)sa_peq[bb >> 4] << 32);
        have = front_filter(ad, sa_peq, sa_pair, peq32, codes, lo, n, hit);
    } else {
        have = myers_filter<WORD>(ad, peq, codes, lo, n, hit);
    }
    if (have) {
        if (ad.band_ok && hit.width <= ATR_K1D_W) {
            if (path) *path = 1;
            if (hit.width <= 8) {                      // k_band8
                if (ad.flags & ATR_START_WITHIN_SEQ1) k1d_band<AND_MODE, 8, true>(ad, codes, lo, n, hit.dlo, b);
                else k1d_band<AND_MODE, 8, false>(ad, codes, lo, n, hit.dlo, b);
            }
            else if (ad.flags & ATR_START_WITHIN_SEQ1) k1d_band<AND_MODE, ATR_K1D_W, true>(ad, codes, lo, n, hit.dlo, b);
            else k1d_band<AND_MODE, ATR_K1D_W, false>(ad, codes, lo, n, hit.dlo, b);
        }
        else { if (path) *path = 2; k1a_locate<AND_MODE>(ad, codes, lo, n, b, hit.c0, hit.c1); }
    }
    if (path && (ad.sa_ok || ad.qg_ok || ad.sa_front)) *path += 10;   // tests: 10 + x = went through the Shift-And / q-gram pre-filter
    finalize(ad, b, n, out);
}


// anchored adapter, one read: filter, then the register DP (what k_filter_anchor + k_anchor_dp do)
template <bool AND_MODE>
ATR_HD void anchor_read(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, atr_match* out) {
    if (anchor_filter(ad, codes, lo, n)) { k1a_read<AND_MODE>(ad, codes, lo, n, out); return; }
    Best b;
    b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
    finalize(ad, b, n, out);
}

// funnel shape with dearer indels, one read: the funnel's first stage as a filter, then the register DP (what
// k_filter_sa / k_filter + k_anchor_dp over the survivor lists do)
template <class WORD, bool AND_MODE>
ATR_HD void icfilter_read(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, atr_match* out, int* path = nullptr) {
    Best b;
    b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
    bool have = false;
    int c0 = -1, c1 = -1;
    if (path) *path = 0;
    if (ad.sa_ok || ad.qg_ok) {
        unsigned sa_peq[16], tail_peq[16];
        const int mp = ad.sa_rows < 32 ? ad.sa_rows : 32, sh32 = 32 - mp;
        for (int c = 0; c < 16; c++) {
            const unsigned low = (unsigned)(ad.peq[c] & (mp >= 32 ? 0xFFFFFFFFull : ((1ull << mp) - 1)));
            sa_peq[c] = low;
            tail_peq[c] = (sh32 ? (low << sh32) | ((1u << sh32) - 1u) : low);
        }
        SaResult sr;
        if (ad.qg_ok) qg_filter(ad, tail_peq, codes, (lo + n + 7) >> 3, lo, n, sr);
        else sa_filter(ad, sa_peq, tail_peq, codes, lo, n, sr);
        if (sr.cls == 3) {
            b.matches = ad.m; b.cost = 0; b.origin = sr.v; b.ref_stop = ad.m; b.q_stop = sr.v + ad.m;
            if (path) *path = 5;
            finalize(ad, b, n, out);
            return;
        }
        have = sr.cls != 0;
        c0 = sr.c0; c1 = sr.c1;
    } else {
        const int WB = (int)(8 * sizeof(WORD)), sh = WB - ad.m;
        WORD peq[16];
        for (int c = 0; c < 16; c++) peq[c] = (WORD)(((WORD)ad.peq[c] << sh) | (sh ? (((WORD)1 << sh) - 1) : 0));
        FilterHit hit;
        have = myers_filter<WORD>(ad, peq, codes, lo, n, hit);
        c0 = hit.c0; c1 = hit.c1;
    }
    // the register DP over the column window the filter stage leaves (what k_anchor_dp does with the survivors' windows)
    if (have) { if (path) *path = 1; k1a_locate<AND_MODE>(ad, codes, lo, n, b, c0, c1); }
    finalize(ad, b, n, out);
}
