// qgram_core.cuh -- q-gram sampling form of the funnel's first stage (k_filter_qg), one thread = one read.
//
// Same necessary condition as the Shift-And stage (locate_core.cuh: sa_scan): an alignment of the adapter's first
// sa_rows rows with <= k unit-cost errors contains one of k+1 pieces verbatim. The automaton finds the pieces by
// touching every column of the read (~7.5 instructions per column, the kernel is bound by instruction issue). Here a
// read position is only looked at if it is a multiple of `step` (2 or 3): a verbatim piece of length
// L >= 6 + step - 1 contains a 6-mer that starts at such a position, whatever the piece's own position. One lookup =
// 24 bits of the packed read -> multiplicative hash -> one byte of an 8 KB table in shared memory (0 = no adapter 6-mer
// hashes here, else which 6-mer of which piece). The few hits are verified by comparing the whole piece at the position
// the 6-mer implies, so the verified hit set -- and with it the range [hmin, hmax] of hit diagonals, the exact-occurrence
// shortcut, the bands and every result -- is exactly that of the automaton (tests/test_hostsim.py compares them).
// ASCII compare mode only (verbatim = equal codes); wildcard modes keep the automaton.
#pragma once
#include "locate_core.cuh"

#define ATR_QG_GROUPS 8               // groups per chunk: 8 lookups each, one 32-bit accumulator per group

// 8 lookups of group g: S = 3 -> 24 columns (3 packed words + 1), S = 2 -> 16 columns (2 words + 1).
// codes: the packed read; word indices >= wlimit are not read (they count as zeros).
template <int S>
ATR_HD uint32_t qg_group(const unsigned char* __restrict__ tab, unsigned mul, const uint32_t* __restrict__ codes,
                         int g, int wlimit) {
    const int w0i = (S == 3 ? 3 : 2) * g;
    uint32_t w0, w1, w2, w3 = 0;
    if (w0i + (S == 3 ? 3 : 2) < wlimit) {               // the common case: every word of the group exists
        w0 = codes[w0i]; w1 = codes[w0i + 1]; w2 = codes[w0i + 2];
        if (S == 3) w3 = codes[w0i + 3];
    } else {
        w0 = w0i < wlimit ? codes[w0i] : 0u;
        w1 = w0i + 1 < wlimit ? codes[w0i + 1] : 0u;
        w2 = w0i + 2 < wlimit ? codes[w0i + 2] : 0u;
        if (S == 3) w3 = w0i + 3 < wlimit ? codes[w0i + 3] : 0u;
    }
    uint32_t x[8];
    if (S == 3) {                                         // nibble offsets 0, 3, 6, ..., 21
        x[0] = w0;                     x[1] = funnel_r32(w0, w1, 12); x[2] = funnel_r32(w0, w1, 24);
        x[3] = funnel_r32(w1, w2, 4);  x[4] = funnel_r32(w1, w2, 16); x[5] = funnel_r32(w1, w2, 28);
        x[6] = funnel_r32(w2, w3, 8);  x[7] = funnel_r32(w2, w3, 20);
    } else {                                              // nibble offsets 0, 2, 4, ..., 14
        x[0] = w0; x[1] = funnel_r32(w0, w1, 8); x[2] = funnel_r32(w0, w1, 16); x[3] = funnel_r32(w0, w1, 24);
        x[4] = w1; x[5] = funnel_r32(w1, w2, 8); x[6] = funnel_r32(w1, w2, 16); x[7] = funnel_r32(w1, w2, 24);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += (uint32_t)tab[(x[i] * mul) >> (32 - ATR_QG_BITS)] << (4 * i);
    return acc;
}

// is the piece of pattern `id` at absolute position a of the packed read? (a + piece length <= end of the read;
// only the words that hold needed bases are loaded)
ATR_HD bool qg_piece_at(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int a, int id) {
    const int plen = ad.qg_plen[id];
    const int wi = a >> 3, sh = a & 7;
    const uint32_t w0 = codes[wi];
    const uint32_t w1 = (sh + atr_min(plen, 8) > 8) ? codes[wi + 1] : 0u;
    uint32_t bad = (funnel_r32(w0, w1, (unsigned)sh * 4u) ^ ad.qg_pw[id][0]) & ad.qg_pm[id][0];
    if (plen > 8) {
        const uint32_t v1 = codes[wi + 1];
        const uint32_t v2 = (sh + plen > 16) ? codes[wi + 2] : 0u;
        bad |= (funnel_r32(v1, v2, (unsigned)sh * 4u) ^ ad.qg_pw[id][1]) & ad.qg_pm[id][1];
    }
    return bad == 0u;
}

// a hit of pattern `id` at sampled absolute position c of the packed read: is the whole piece there, inside the
// window? -> its diagonal v = (column of the piece end) - (row of the piece end), else ATR_QG_NOHIT
#define ATR_QG_NOHIT 0x7fffffff
ATR_HD int qg_hit_diagonal(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, int id, int c) {
    const int prow = ad.qg_prow[id], plen = ad.qg_plen[id];
    const int a = c - (int)ad.qg_poff[id];               // absolute position of the piece's first base
    if (a < lo || a + plen > lo + n) return ATR_QG_NOHIT;
    if (!qg_piece_at(ad, codes, a, id)) return ATR_QG_NOHIT;
    return (a - lo) - prow;
}

ATR_HD void qg_verify_pattern(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, int id, int c,
                              int& hmin, int& hmax) {
    const int v = qg_hit_diagonal(ad, codes, lo, n, id, c);
    if (v != ATR_QG_NOHIT) { hmin = atr_min(hmin, v); hmax = atr_max(hmax, v); }
}

template <int S>
ATR_HD void qg_verify_group(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int lo, int n, int g, uint32_t acc,
                            int& hmin, int& hmax) {
    while (acc) {
        const int i = atr_ctz(acc) >> 2;
        const int id = (int)((acc >> (4 * i)) & 15u);
        acc &= ~(15u << (4 * i));
        const int c = (S == 3 ? 24 : 16) * g + S * i;
        if (id == 15) { for (int t = 1; t <= ad.qg_npat; t++) qg_verify_pattern(ad, codes, lo, n, t, c, hmin, hmax); }
        else qg_verify_pattern(ad, codes, lo, n, id, c, hmin, hmax);
    }
}

// groups that hold a sampled position at which a 6-mer can lie inside the window [lo, lo + n)
template <int S>
ATR_HD void qg_group_range(int lo, int n, int& g0, int& g1) {
    const int G = S == 3 ? 24 : 16;
    g0 = lo / G;
    g1 = n >= ATR_QG_Q ? (lo + n - ATR_QG_Q) / G + 1 : g0;        // exclusive
}

// (a) the whole scan for one read: range of the verified hit diagonals. acc: scratch of ATR_QG_GROUPS words, element
// t at acc[t * stride] (shared memory [group][thread] in the kernel).
template <int S>
ATR_HD void qg_scan(const AdapterK1a& ad, const unsigned char* __restrict__ tab, const uint32_t* __restrict__ codes, int wlimit,
                    int lo, int n, uint32_t* acc, int stride, int& hmin, int& hmax) {
    hmin = 0x7fffffff; hmax = -0x7fffffff;
    int g0, g1;
    qg_group_range<S>(lo, n, g0, g1);
    for (int gb = g0; gb < g1; gb += ATR_QG_GROUPS) {
        const int ng = atr_min(ATR_QG_GROUPS, g1 - gb);
        uint32_t any = 0;
        for (int t = 0; t < ng; t++) {
            const uint32_t a = qg_group<S>(tab, ad.qg_mul, codes, gb + t, wlimit);
            acc[t * stride] = a;
            any |= a;
        }
        if (any) for (int t = 0; t < ng; t++) {
            const uint32_t a = acc[t * stride];
            if (a) qg_verify_group<S>(ad, codes, lo, n, gb + t, a, hmin, hmax);
        }
    }
}

// (c) the need-tail gate of sa_need_tail without the automaton's state after the whole read: a tail_mask bit
// (row i, l rows deep in its piece) only depends on the read's last l <= tail_cols columns, so the automaton is run over
// the last 8 (or 16) columns only. sa_peq: the 16-entry Peq table of the first sa_rows rows (as for sa_scan); WORD is
// unsigned, or unsigned long long for the wide form (sa_rows > 32).
template <class WORD>
ATR_HD bool qg_need_tail(const AdapterK1a& ad, const WORD* __restrict__ sa_peq, const uint32_t* __restrict__ codes,
                         int lo, int n, int hmax) {
    const bool stop_in_ref = ad.flags & ATR_STOP_WITHIN_SEQ1;
    if (!(stop_in_ref || ad.m <= ad.sa_rows)) return false;
    if (!ad.tail_gate_ok || ad.tail_cols > 16) return true;
    if (hmax != -0x7fffffff && hmax >= n - ad.sa_rows - ad.k) return true;
    const int cols = ad.tail_cols > 8 ? 16 : 8;
    if (lo + n < cols) return true;                        // shorter than the look-back: let the exact pass decide
    const int q = lo + n - cols;                           // columns before the window start only add chains: a superset
    const unsigned sh = (unsigned)(q & 7) * 4u;
    const uint32_t w0 = codes[q >> 3];
    const uint32_t w1 = (q & 7) || cols == 16 ? codes[(q >> 3) + 1] : 0u;
    const WORD S0 = (WORD)ad.sa_start64;
    WORD St = 0;
    uint32_t tw = funnel_r32(w0, w1, sh);
#pragma unroll
    for (int t = 0; t < 8; t++) St = ((St << 1) | S0) & sa_peq[(tw >> (4 * t)) & 15u];
    if (cols == 16) {
        const uint32_t w2 = (q & 7) ? codes[(q >> 3) + 2] : 0u;
        tw = funnel_r32(w1, w2, sh);
#pragma unroll
        for (int t = 0; t < 8; t++) St = ((St << 1) | S0) & sa_peq[(tw >> (4 * t)) & 15u];
    }
    return (St & (WORD)ad.tail_mask64) != (WORD)0;
}

// the Peq tables of the first sa_rows rows: right-aligned for the automaton, left-aligned with the virtual rows set for Myers
template <class WORD>
ATR_HD void qg_peq_tables(const AdapterK1a& ad, int c, WORD& sa, WORD& tail) {
    const int WB = (int)(8 * sizeof(WORD)), mp = ad.sa_rows, sh = WB - mp;
    const WORD low = (WORD)(ad.peq[c] & (mp >= 64 ? ~0ull : ((1ull << mp) - 1)));
    sa = low;
    tail = sh ? (WORD)((low << sh) | (((WORD)1 << sh) - 1)) : low;
}

// the whole stage for one read (host simulator; the kernel interleaves block-level compactions)
template <int S, class WORD>
ATR_HD void qg_filter_s(const AdapterK1a& ad, const uint32_t* __restrict__ codes, int wlimit, int lo, int n, SaResult& res) {
    int hmin, hmax, imin = 0, imax = 0;
    uint32_t acc[ATR_QG_GROUPS];
    WORD sa_peq[16], tail_peq[16];
    for (int c = 0; c < 16; c++) qg_peq_tables<WORD>(ad, c, sa_peq[c], tail_peq[c]);
    qg_scan<S>(ad, ad.qg_tab, codes, wlimit, lo, n, acc, 1, hmin, hmax);
    if (sa_exact(ad, codes, lo, n, hmin, hmax)) { res.cls = 3; res.v = hmin; return; }
    if (qg_need_tail<WORD>(ad, sa_peq, codes, lo, n, hmax)) sa_tail_w<WORD>(ad, tail_peq, codes, lo, n, imin, imax);
    sa_classify(ad, lo, n, hmin, hmax, imin, imax, res);
}

ATR_HD void qg_filter(const AdapterK1a& ad, const unsigned* __restrict__ /*tail_peq32*/, const uint32_t* __restrict__ codes, int wlimit,
                      int lo, int n, SaResult& res) {
    if (ad.qg_wide) {
        if (ad.qg_step == 3) qg_filter_s<3, unsigned long long>(ad, codes, wlimit, lo, n, res);
        else qg_filter_s<2, unsigned long long>(ad, codes, wlimit, lo, n, res);
    } else {
        if (ad.qg_step == 3) qg_filter_s<3, unsigned>(ad, codes, wlimit, lo, n, res);
        else qg_filter_s<2, unsigned>(ad, codes, wlimit, lo, n, res);
    }
}
