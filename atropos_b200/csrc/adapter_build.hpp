// adapter_build.hpp -- host-side preparation of the per-adapter device tables (pure C++, no CUDA).
//
// Everything that involves floating point is evaluated HERE, once per adapter, with the very
// expressions of the reference (C doubles, as Cython compiles them), and shipped to the kernels as
// small integer tables:
//   k          = (int)(max_error_rate * m)                      _align.pyx:312
//   thr_mul[l] = max integer c with (double)c <= l * rate       _align.pyx:447, :468  (cost <= length * max_error_rate)
//   thr_div[s] = max integer e with (double)e / s <= rate       adapters/__init__.py:389-392 (errors / size <= max_error_rate)
// Translation tables: _align.pyx:31-83.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cfenv>
#include "atr_common.cuh"
#include "insert_core.cuh"

namespace atr {

inline void build_tables(AtrTables& t) {
    std::memset(&t, 0, sizeof(t));
    auto put = [](unsigned char* tab, char c, int v) {
        tab[(unsigned char)c] = (unsigned char)v;
        tab[(unsigned char)(c + 32)] = (unsigned char)v;
    };
    const int A = 1, C = 2, G = 4, T = 8;
    put(t.acgt, 'A', A); put(t.acgt, 'C', C); put(t.acgt, 'G', G); put(t.acgt, 'T', T); put(t.acgt, 'U', T);
    put(t.iupac, 'X', 0); put(t.iupac, 'A', A); put(t.iupac, 'C', C); put(t.iupac, 'G', G); put(t.iupac, 'T', T);
    put(t.iupac, 'U', T); put(t.iupac, 'R', A | G); put(t.iupac, 'Y', C | T); put(t.iupac, 'S', G | C);
    put(t.iupac, 'W', A | T); put(t.iupac, 'K', G | T); put(t.iupac, 'M', A | C); put(t.iupac, 'B', C | G | T);
    put(t.iupac, 'D', A | G | T); put(t.iupac, 'H', A | C | T); put(t.iupac, 'V', A | C | G);
    put(t.iupac, 'N', A | C | G | T);
}

// The 16 upper-case letters whose 4-bit IUPAC code identifies them uniquely: for these, equality of
// codes == equality of ASCII bytes, so the packed kernel is exact in the reference's ASCII mode.
inline bool exact_symbol(unsigned char c) {
    switch (c) {
        case 'X': case 'A': case 'C': case 'G': case 'T': case 'R': case 'Y': case 'S': case 'W':
        case 'K': case 'M': case 'B': case 'D': case 'H': case 'V': case 'N': return true;
        default: return false;
    }
}

struct HostAdapter {
    atr_adapter_desc desc;            // copy of the caller's arguments (sequence pointer NOT retained)
    std::string seq;
    int m = 0, k = 0, ic_eff = 1;
    bool and_mode = false, k1a_ok = false;
    int q_table = 0;                  // 0 ASCII, 1 IUPAC, 2 ACGT
    int cmp_only = 0;                 // 0 DP, 1 compare_prefixes, 2 compare_suffixes
    // Adapter.match_to tries str.find/startswith/endswith first when adapter_wildcards is off
    // (adapters/__init__.py:351-367). In ASCII mode the DP returns the very same hit (cost 0, m matches,
    // leftmost), so only the RMP bypass has to be honoured; with read wildcards on (AND mode) a cost-0
    // DP hit need not be a literal hit, so the literal search is run explicitly.
    bool need_find = false;
    bool lit_exact = true;            // every adapter letter is one of the 16 exactly-coded symbols
    std::vector<unsigned char> ref_gen;        // K1g operand: ASCII or translated
    std::vector<unsigned short> thr_mul, thr_div;
    std::vector<unsigned char> rmp_ok;         // empty = no gate
};

// returns 0 or ATR_E_*; msg filled on error
inline int prepare_adapter(const atr_adapter_desc& d, const AtrTables& tb, HostAdapter& h, std::string& msg) {
    if (d.sequence == nullptr || d.length < 1) { msg = "empty adapter sequence"; return ATR_E_ARG; }
    if (d.length > ATR_MAX_ADAPTER) { msg = "adapter longer than 4095 nt"; return ATR_E_LIMIT; }
    if (d.min_overlap < 1) { msg = "Minimum overlap must be at least 1"; return ATR_E_ARG; }            // _align.pyx:218-220
    if (d.indel_cost < 1) { msg = "Insertion/deletion cost must be at least 1"; return ATR_E_ARG; }     // _align.pyx:228-230
    if (!(d.max_error_rate >= 0.0)) { msg = "max_error_rate must be >= 0"; return ATR_E_ARG; }
    h.desc = d;
    h.seq.assign(d.sequence, (size_t)d.length);
    h.desc.sequence = nullptr;
    h.desc.rmp_ok = nullptr;
    const int m = h.m = d.length;
    for (unsigned char c : h.seq) if (c >= 128) { msg = "non-ASCII adapter"; return ATR_E_ARG; }
    h.k = (int)(d.max_error_rate * m);
    // indel costs above k+1 all mean "an indel kills the path": clamp so packed costs cannot overflow
    h.ic_eff = d.indel_cost > h.k + 2 ? h.k + 2 : d.indel_cost;
    h.and_mode = d.wildcard_ref || d.wildcard_query;
    h.q_table = !h.and_mode ? 0 : (d.wildcard_query ? 1 : 2);
    h.ref_gen.resize((size_t)m);
    bool all_exact = true;
    for (int i = 0; i < m; i++) {
        const unsigned char c = (unsigned char)h.seq[i];
        if (!h.and_mode) { h.ref_gen[i] = c; all_exact = all_exact && exact_symbol(c); }
        else h.ref_gen[i] = d.wildcard_ref ? tb.iupac[c] : tb.acgt[c];      // _align.pyx:245-248
    }
    // packed keys hold cost <= 255: (k+1) + ic_eff <= 2k+3 must fit
    h.need_find = d.match_to_semantics && !d.wildcard_ref && d.wildcard_query;
    h.lit_exact = true;
    for (unsigned char c : h.seq) h.lit_exact = h.lit_exact && exact_symbol(c);
    h.k1a_ok = m <= ATR_K1A_MAXM && h.k <= 126 && (h.and_mode || all_exact) && (!h.need_find || h.lit_exact);
    h.thr_mul.assign((size_t)m + 1, 0);
    h.thr_div.assign((size_t)m + 1, 0);
    for (int l = 0; l <= m; l++) {
        const double lim = l * d.max_error_rate;
        double f = std::floor(lim);
        if (f > 60000.0) f = 60000.0;
        int c = (int)f;
        while ((double)(c + 1) <= lim && c < 60000) c++;
        while (c > 0 && !((double)c <= lim)) c--;
        h.thr_mul[l] = (unsigned short)c;
        int e = 0;
        if (l > 0) while (e < 60000 && (double)(e + 1) / (double)l <= d.max_error_rate) e++;
        h.thr_div[l] = (unsigned short)e;
    }
    h.cmp_only = 0;
    if (d.match_to_semantics && d.no_indels) {                              // adapters/__init__.py:370-380
        if (d.flags == ATR_STOP_WITHIN_SEQ2) h.cmp_only = 1;                // PREFIX
        else if (d.flags == ATR_START_WITHIN_SEQ2) h.cmp_only = 2;          // SUFFIX
    }
    if (d.rmp_ok != nullptr) h.rmp_ok.assign(d.rmp_ok, d.rmp_ok + (size_t)(m + 1) * (m + 1));
    return ATR_OK;
}

// Fill the by-value parameter block of the register kernel. rmp_ok_dev: device (or host, for the simulator) pointer.
inline void fill_k1a(const HostAdapter& h, const AtrTables& tb, int index, int reduce, const unsigned char* rmp_ok_dev,
                     AdapterK1a& a) {
    std::memset(&a, 0, sizeof(a));
    a.m = h.m; a.k = h.k; a.flags = h.desc.flags; a.ic = h.ic_eff; a.min_overlap = h.desc.min_overlap;
    a.and_mode = h.and_mode; a.q_single_only = h.q_table == 2;
    a.match_to = h.desc.match_to_semantics; a.exact_bypass = !h.desc.wildcard_ref && !h.need_find; a.cmp_only = h.cmp_only;
    a.adapter_index = index; a.reduce = reduce;
    a.need_find = h.need_find;
    for (int i = 0; i < ATR_K1A_MAXM; i++) a.lit[i] = i < h.m ? tb.iupac[(unsigned char)h.seq[i]] : 0x100;
    for (int i = 0; i < ATR_K1A_MAXM; i++) {
        if (i < h.m) a.code[i] = h.and_mode ? h.ref_gen[i] : tb.iupac[(unsigned char)h.seq[i]];
        else a.code[i] = h.and_mode ? 0 : 0x100;                            // pad rows never match
    }
    for (int l = 0; l <= ATR_K1A_MAXM; l++) {
        a.thr_mul[l] = l <= h.m ? h.thr_mul[l] : 0;
        a.thr_div[l] = l <= h.m ? h.thr_div[l] : 0;
    }
    a.rmp_ok = rmp_ok_dev;
    // the filter -> banded DP funnel needs unit indel cost, a free start in the read and an
    // anchored start in the adapter (BACK / SUFFIX style flag sets)
    const bool start_in_ref = h.desc.flags & ATR_START_WITHIN_SEQ1, start_in_query = h.desc.flags & ATR_START_WITHIN_SEQ2;
    const bool stop_q = h.desc.flags & ATR_STOP_WITHIN_SEQ2;
    // BACK / SUFFIX style (anchored start in the adapter), or FRONT / ANYWHERE style (free start in both, needs
    // stop_in_query so that the first DP column is column 0)
    const bool funnel_shape = h.k1a_ok && start_in_query && (!start_in_ref || stop_q) && !h.cmp_only && !h.need_find;
    a.fused_ok = funnel_shape && h.desc.indel_cost == 1;
    // Dearer indels (insert mode's fallback adapters cost 3, --no-indels 100000): every accepted alignment is also one of
    // at most k unit-cost edits, so the unit-cost filter stage is still a valid necessary condition (and its verbatim-
    // occurrence shortcut is still the answer); only the DP itself has to price the indels -- the register DP does.
    a.filter_only = funnel_shape && h.desc.indel_cost != 1;
    // an alignment that ends at (m, j) and covers r adapter rows (it starts inside the adapter at row m - r, or r = m)
    // needs r >= min_overlap and a cost c <= floor(r * rate) (consider(): length = r); r rows over j columns take at
    // least r - j deletions, so r - j <= c. The bound on D[m][j] is the largest budget among the r that can do that;
    // -1 = no r can (columns 1 and 2 at min_overlap 3: the first build allowed "3 rows, cost 0" there, which one column
    // cannot hold, and a quarter of all reads -- last adapter base == first read base -- went on to the band kernel)
    for (int j = 0; j <= ATR_K1A_MAXM; j++) {
        const int jj = j < h.m ? j : h.m;
        int best = -1;
        for (int r = h.desc.min_overlap > 1 ? h.desc.min_overlap : 1; r <= h.m; r++)
            if (r - jj <= (int)h.thr_mul[r]) best = best > (int)h.thr_mul[r] ? best : (int)h.thr_mul[r];
        a.thrJ[j] = (short)best;
    }
    // a read code that no adapter row matches: 0 in AND mode, an unused letter code in ASCII mode
    a.nomatch = 0; a.band_ok = a.fused_ok;
    if (!a.and_mode) {
        int free_code = -1;
        for (int c = 15; c >= 0 && free_code < 0; c--) {
            bool used = false;
            for (int i = 0; i < h.m && i < ATR_K1A_MAXM; i++) used = used || a.code[i] == c;
            if (!used) free_code = c;
        }
        if (free_code < 0) a.band_ok = 0; else a.nomatch = free_code;
    }
    for (int c = 0; c < 16; c++) {
        unsigned cq = (unsigned)c;
        if (a.and_mode && a.q_single_only) cq = (cq & (cq - 1)) ? 0u : cq;
        unsigned long long bits = 0;
        for (int i = 0; i < h.m && i < 64; i++) {
            const unsigned code = (unsigned)a.code[i];
            const bool eq = a.and_mode ? ((code & cq) != 0u) : (code == (unsigned)c);
            if (eq) bits |= 1ull << i;
        }
        a.peq[c] = bits;
    }
    for (int w = 0; w < ATR_K1A_MAXM / 8; w++) a.apack[w] = 0;
    for (int i = 0; i < h.m && i < ATR_K1A_MAXM; i++) a.apack[i >> 3] |= ((unsigned)a.code[i] & 15u) << (4 * (i & 7));
    a.exact_ok = !(a.and_mode && a.q_single_only) && h.m >= h.desc.min_overlap;
    // anchored adapters (PREFIX = stop_in_query only, SUFFIX = start_in_query only) that the funnel does not take:
    // the whole adapter must align at the read start / end with <= k errors, so one of k+1 pieces sits verbatim
    // within k columns of its anchored position (locate_core.cuh: anchor_filter)
    a.anchor_ok = h.k1a_ok && !a.fused_ok && !h.cmp_only && !h.need_find && a.exact_ok &&
                  (h.desc.flags == ATR_STOP_WITHIN_SEQ2 || h.desc.flags == ATR_START_WITHIN_SEQ2);
    // Pieces for the pigeonhole stages: k+1 pieces over the first `rows` adapter rows, each at least `min_len` rows (shorter
    // pieces hit at random too often to be a filter); needs row-m candidates to be reported inside the loop (stop_in_query).
    // Shift-And automaton (k_filter_sa) and q-gram sampling (k_filter_qg): rows = min(m, 32), pieces >= 6 rows.
    // Adapters whose k+1 pieces do not fit 32 rows that way (the 58-nt TruSeq adapter at 0.1: k = 5) get pieces over
    // min(m, 64) rows for the q-gram form alone (qg_wide; 64-bit tail pass and gate).
    a.split8 = 0;
    a.sa_ok = 0; a.sa_rows = h.m < 32 ? h.m : 32; a.sa_start = 0; a.sa_end = 0; a.tail_gate_ok = 0; a.tail_mask = 0;
    a.qg_wide = 0; a.sa_start64 = a.sa_end64 = a.tail_mask64 = 0;
    const bool stop_in_query = h.desc.flags & ATR_STOP_WITHIN_SEQ2;
    const int pieces = h.k + 1;
    // lens: the pieces' lengths in rows (they cover rows 1 .. sum). Returns whether the tail gate works for this layout.
    auto layout = [&](const std::vector<int>& lens) -> int {
        unsigned long long st = 0, en = 0, mask = 0;
        int row = 1;
        for (int len : lens) {
            st |= 1ull << (row - 1);
            en |= 1ull << (row + len - 2);
            row += len;
        }
        const int rows = row - 1;
        // Gate for the exact tail pass. A last-column candidate (i, n), i <= rows, has e <= thr_mul[i] errors
        // over rows 1..i, which contain c(i) complete pieces. e < c: a complete piece is verbatim (a hit near the
        // read end). e == c: either that, or every complete piece is broken and the rows after them -- the begun
        // piece -- are verbatim up to column n, which the automaton shows as bit i-1 of its final state.
        // e > c (or e == c with no begun piece): no cheap certificate -> the tail pass always runs.
        const bool stop_in_ref = h.desc.flags & ATR_STOP_WITHIN_SEQ1;
        int gate = 1;
        const int first = stop_in_ref ? 1 : h.m;
        for (int i = 1; i <= rows; i++) {
            if (i < first || i < h.desc.min_overlap) continue;
            int c = 0, end_c = 0, r = 1;
            for (int len : lens) {
                if (r + len - 1 <= i) { c++; end_c = r + len - 1; }
                r += len;
            }
            const int e = (int)h.thr_mul[i];
            if (e < c) continue;
            if (e == c && i > end_c) { mask |= 1ull << (i - 1); continue; }
            gate = 0;
        }
        a.sa_rows = rows; a.sa_start64 = st; a.sa_end64 = en; a.tail_mask64 = mask; a.tail_gate_ok = gate;
        a.sa_start = (unsigned)st; a.sa_end = (unsigned)en; a.tail_mask = (unsigned)mask;
        return gate;
    };
    auto equal_split = [&](int rows) {
        std::vector<int> lens;
        for (int pc = 0; pc < pieces; pc++) lens.push_back(rows / pieces + (pc < rows % pieces ? 1 : 0));
        return lens;
    };
    // pieces that end right before the rows where the error budget grows (thr_mul[i] steps from p - 1 to p): then the
    // budget never reaches the number of complete pieces AT a piece end, which is the one case the gate cannot certify
    // (equal pieces of exactly 1 / rate rows, the 58-nt adapter at 0.1, have it at every piece end)
    // slack > 0 ends the pieces that many rows earlier still: the first mask row after a piece end then lies slack + 1
    // rows deep in the begun piece, and a gate bit l rows deep fires by chance in 4^-l of the reads (with slack 0 the 58-nt
    // adapter has five mask rows of depth 1: some begun piece "matches" the last base of 3 reads in 4, and the 64-bit
    // tail pass ran for 95 % of the reads -- 55 % of the kernel's instructions)
    auto budget_split = [&](int rows, int slack) {
        std::vector<int> lens;
        int prev = 0;
        for (int pc = 1; pc < pieces; pc++) {
            int i = prev + 1;
            while (i <= rows && (int)h.thr_mul[i] < pc) i++;     // first row whose budget is pc
            const int end = i - 1 - slack;
            lens.push_back(end - prev);
            prev = end;
        }
        lens.push_back(rows - prev);
        return lens;
    };
    // expected fraction of random reads that pass the gate: sum over the mask rows of 4^-(rows into the begun piece)
    auto gate_pass_rate = [&](const std::vector<int>& lens) {
        double p = 0;
        int r = 1;
        for (int len : lens) {
            for (int d = 1; d <= len; d++)
                if ((a.tail_mask64 >> (r + d - 2)) & 1ull) p += std::pow(0.25, d);
            r += len;
        }
        return p;
    };
    const bool shape = (a.fused_ok || a.filter_only) && !start_in_ref && stop_in_query;
    // 5' adapters that may also start before the read (FRONT): every candidate that ends beyond column m + k is a
    // full-length occurrence (an alignment that starts inside the adapter starts in column 0 and has at most k
    // insertions), hence holds a verbatim piece; the others are found exactly by a Myers pass over the first m + k columns
    const bool stop_in_ref_f = h.desc.flags & ATR_STOP_WITHIN_SEQ1;
    a.sa_front = 0;
    if (a.fused_ok && start_in_ref && stop_in_query && !stop_in_ref_f && (h.desc.flags & ATR_START_WITHIN_SEQ2) && h.m <= 32 &&
        pieces <= a.sa_rows && a.sa_rows / pieces >= 6) {
        a.sa_front = 1;
        layout(equal_split(a.sa_rows));
        a.tail_gate_ok = 0; a.tail_mask = 0; a.tail_mask64 = 0;      // no last-column candidates below row m: nothing to gate
    }
    if (shape && pieces <= a.sa_rows && a.sa_rows / pieces >= 6) {
        a.sa_ok = 1;
        layout(equal_split(a.sa_rows));
    } else if (shape && h.m > 32 && !a.and_mode) {
        const int rows = h.m < 64 ? h.m : 64;
        if (pieces <= rows && rows / pieces >= 7) {
            a.qg_wide = 1;
            if (!layout(equal_split(rows))) {
                // among the budget-aligned layouts with a working gate take the one whose gate fires least often by chance
                std::vector<int> best;
                double best_rate = 1e9;
                for (int slack = 0; slack <= 3; slack++) {
                    const std::vector<int> alt = budget_split(rows, slack);
                    bool ok = true;
                    for (int len : alt) ok = ok && len >= 7 && len <= 16;
                    if (!ok || !layout(alt)) continue;
                    const double rate = gate_pass_rate(alt) + (alt[0] < 8 ? 0.02 : 0.0);   // a 7-row piece costs the step-3 sampling
                    if (rate < best_rate) { best_rate = rate; best = alt; }
                }
                if (best.empty() || !layout(best)) layout(equal_split(rows));
            }
        }
    }
}

// ---- q-gram sampling pre-filter tables (qgram_core.cuh) ----------------------------------------------------------
// Needs the Shift-And piece layout of fill_k1a (a.sa_ok). tab: 1 << ATR_QG_BITS bytes, to be placed where the kernels
// (device memory) or the simulator (host memory) can read it; the caller stores that pointer in a.qg_tab.
// Returns false (and leaves a.qg_ok = 0) when the adapter does not qualify.
inline bool build_qg(AdapterK1a& a, std::vector<unsigned char>& tab) {
    a.qg_ok = 0; a.qg_tab = nullptr; a.qg_npat = 0; a.n_tail_cmp = 0; a.tail_cols = 0;
    if ((!a.sa_ok && !a.qg_wide) || a.and_mode) return false;
    // pieces: runs of rows between sa_start bits and sa_end bits
    int pstart[8], plen[8], np = 0, lmin = 1 << 30, lmax = 0;
    for (int r = 0; r < a.sa_rows; r++) {
        if (a.sa_start64 & (1ull << r)) {
            int e = r;
            while (!(a.sa_end64 & (1ull << e))) e++;
            if (np == 8) return false;
            pstart[np] = r; plen[np] = e - r + 1;
            lmin = std::min(lmin, plen[np]); lmax = std::max(lmax, plen[np]);
            np++;
        }
    }
    if (np == 0) return false;
    const int q = ATR_QG_Q;
    int step = lmin - q + 1;                              // q + step - 1 <= shortest piece
    if (step < 2) return false;                           // every column would be sampled: the automaton is as cheap
    if (step > 3) step = 3;
    // patterns: the 6-mers at offsets 0 .. step-1 of every piece. Whatever the piece's position a in the read,
    // exactly one of these offsets o makes a + o a sampled position, and o + 6 <= step - 1 + 6 <= piece length.
    if (np * step > 14 && step == 3) step = 2;
    struct Pat { unsigned x; int piece, off; };
    std::vector<Pat> pats;
    for (int p = 0; p < np; p++)
        for (int o = 0; o < step; o++) {
            unsigned x = 0;
            for (int t = 0; t < q; t++) x |= ((unsigned)a.code[pstart[p] + o + t] & 15u) << (4 * t);
            pats.push_back({x, p, o});
        }
    if (pats.size() > 14) return false;
    // hash multiplier: low 8 bits zero (the 8 bits above the 6-mer in the extracted word must not count); prefer one
    // that keeps the patterns in distinct buckets
    static const unsigned cands[] = {0x02416821u, 0x5b33e441u, 0x23e9f297u, 0x9e3779b1u, 0x85ebca6bu, 0xc2b2ae35u, 0x27d4eb2fu, 0x165667b1u};
    unsigned best_mul = 0; int best_coll = 1 << 30;
    for (unsigned c : cands) {
        const unsigned mul = (c | 1u) << 8;
        std::vector<int> seen;
        int coll = 0;
        for (const Pat& pt : pats) {
            const int key = (int)((pt.x * mul) >> (32 - ATR_QG_BITS));
            for (size_t t = 0; t < seen.size(); t++) if (seen[t] == key && pats[t].x != pt.x) coll++;
            seen.push_back(key);
        }
        if (coll < best_coll) { best_coll = coll; best_mul = mul; }
    }
    tab.assign((size_t)1 << ATR_QG_BITS, 0);
    for (size_t t = 0; t < pats.size(); t++) {
        const unsigned key = (pats[t].x * best_mul) >> (32 - ATR_QG_BITS);
        tab[key] = tab[key] == 0 ? (unsigned char)(t + 1) : 15;       // 15: verify every pattern (repeats inside the adapter land here too)
        a.qg_prow[t + 1] = (unsigned char)pstart[pats[t].piece];
        a.qg_plen[t + 1] = (unsigned char)plen[pats[t].piece];
        a.qg_poff[t + 1] = (unsigned char)pats[t].off;
        for (int h = 0; h < 2; h++) { a.qg_pw[t + 1][h] = 0; a.qg_pm[t + 1][h] = 0; }
        for (int r = 0; r < plen[pats[t].piece] && r < 16; r++) {
            a.qg_pw[t + 1][r >> 3] |= ((unsigned)a.code[pstart[pats[t].piece] + r] & 15u) << (4 * (r & 7));
            a.qg_pm[t + 1][r >> 3] |= 15u << (4 * (r & 7));
        }
    }
    if (lmax > 16) return false;
    a.qg_npat = (int)pats.size();
    a.qg_step = step; a.qg_mul = best_mul;
    // need-tail gate: one compare per tail_mask row (see fill_k1a). Rows whose prefix is longer than 8 bases are
    // tested on their last 8 only (a necessary condition; the gate only decides whether the exact tail pass runs).
    if (a.tail_gate_ok) {
        for (int i = 1; i <= a.sa_rows; i++) {
            if (!(a.tail_mask64 & (1ull << (i - 1)))) continue;
            int p = 0;
            while (p + 1 < np && pstart[p + 1] < i) p++;              // piece containing row i (1-based): pstart[p] < i
            const int l = i - pstart[p];                               // rows pstart[p]+1 .. i
            a.tail_cols = std::max(a.tail_cols, l);
            const int lc = l > 8 ? 8 : l;
            unsigned c = 0;                                            // rows i-lc+1 .. i in the top lc nibbles
            for (int t = 0; t < lc; t++) c |= ((unsigned)a.code[i - lc + t] & 15u) << (4 * (8 - lc + t));
            if (a.n_tail_cmp == 24) { a.n_tail_cmp = -1; break; }      // too many rows: the tail pass always runs
            a.tail_c[a.n_tail_cmp] = c;
            a.tail_m[a.n_tail_cmp] = lc == 8 ? 0xFFFFFFFFu : ~((1u << (4 * (8 - lc))) - 1u);
            a.n_tail_cmp++;
        }
    }
    a.qg_ok = 1;
    return true;
}

inline void fill_gen(const HostAdapter& h, int index, int reduce, const unsigned char* ref_dev, const unsigned char* lit_dev,
                     const unsigned short* thr_mul_dev, const unsigned short* thr_div_dev,
                     const unsigned char* rmp_ok_dev, AdapterGen& g) {
    std::memset(&g, 0, sizeof(g));
    g.m = h.m; g.k = h.k; g.flags = h.desc.flags; g.ic = h.ic_eff; g.min_overlap = h.desc.min_overlap;
    g.and_mode = h.and_mode; g.q_table = h.q_table;
    g.match_to = h.desc.match_to_semantics; g.exact_bypass = !h.desc.wildcard_ref && !h.need_find; g.cmp_only = h.cmp_only;
    g.adapter_index = index; g.reduce = reduce; g.rate = h.desc.max_error_rate;
    g.need_find = h.need_find; g.lit = lit_dev;
    g.ref = ref_dev; g.thr_mul = thr_mul_dev; g.thr_div = thr_div_dev; g.rmp_ok = rmp_ok_dev;
}

// ---- 4-bit packing of one read (the device packer runs the same function per output word) ----
// returns the packed word `widx` of a read of `len` bases; *esc |= 1 if a byte is not exactly representable
ATR_HD uint32_t pack_word(const unsigned char* __restrict__ ascii, int len, int widx, int fold_case,
                          const unsigned char* __restrict__ iupac, int* esc) {
    uint32_t w = 0;
    const int base = widx * 8;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int p = base + t;
        if (p < len) {
            unsigned char c = ascii[p];
            if (fold_case && c >= 'a' && c <= 'z') c = (unsigned char)(c - 32);
            const unsigned code = iupac[c];
            // exact iff c is one of the 16 upper-case IUPAC letters: code != 0 and upper case, or 'X'
            const bool exact = (c >= 'A' && c <= 'Z' && c != 'U' && (code != 0 || c == 'X'));
            if (!exact) *esc |= 1;
            w |= code << (4 * t);
        }
    }
    return w;
}

// thresholds "largest integer c with (double)c <= l * rate" (cost <= length * max_error_rate in C doubles)
inline unsigned short thr_mul_of(int l, double rate) {
    const double lim = l * rate;
    double f = std::floor(lim);
    if (f > 60000.0) f = 60000.0;
    if (f < 0) f = 0;
    int c = (int)f;
    while (c < 60000 && (double)(c + 1) <= lim) c++;
    while (c > 0 && !((double)c <= lim)) c--;
    return (unsigned short)c;
}

// ---- InsertAligner.__init__ (align/__init__.py:206-233): host tables of the K2 kernels ----------
struct HostInsert {
    InsertDev dev;                    // scalar fields filled; pointer fields set by the owner (device or host pointers)
    std::vector<unsigned short> k_by_len, thr_ins, maxmm;
    std::vector<unsigned char> a1_code, a2_code, a1_ascii, a2_ascii, comp, ov_tab;
    std::vector<uint32_t> a1_pack, a2_pack;
    std::vector<double> insert_prob, adapter_prob;
};

inline int prepare_insert(const atr_insert_desc& d, const AtrTables& tb, HostInsert& h, std::string& msg) {
    if (!d.adapter1 || !d.adapter2 || d.adapter1_len < 1 || d.adapter2_len < 1 || d.max_len < 1 || !d.insert_prob ||
        !d.adapter_prob) { msg = "bad arguments to atr_insertset_create"; return ATR_E_ARG; }
    if (d.max_len > ATR_MAX_READ || d.adapter1_len > ATR_MAX_ADAPTER || d.adapter2_len > ATR_MAX_ADAPTER) {
        msg = "insert aligner: read or adapter too long"; return ATR_E_LIMIT;
    }
    if (!(d.max_insert_mismatch_frac >= 0.0) || !(d.max_adapter_mismatch_frac >= 0.0)) {
        msg = "mismatch fractions must be >= 0"; return ATR_E_ARG;
    }
    InsertDev& v = h.dev;
    std::memset(&v, 0, sizeof(v));
    const int L = d.max_len, amax = std::max(d.adapter1_len, d.adapter2_len);
    const double rate = d.max_insert_mismatch_frac;
    v.min_insert_overlap = d.min_insert_overlap; v.min_adapter_overlap = d.min_adapter_overlap;
    v.cutoff = d.adapter_check_cutoff; v.max_len = L; v.kmax = (int)(rate * L);
    v.a1_len = d.adapter1_len; v.a2_len = d.adapter2_len; v.amax = amax;
    v.insert_max_rmp = d.insert_max_rmp; v.adapter_max_rmp = d.adapter_max_rmp;
    const bool aw = d.adapter_wildcards != 0, rw = d.read_wildcards != 0;
    // compare_prefixes(read_overhang, adapter, wildcard_ref=adapter_wildcards, wildcard_query=read_wildcards):
    // the READ overhang is the "ref" argument (align/__init__.py:285-288; _align.pyx:521-530)
    v.and_mode = aw || rw;
    v.ov_single_only = (!aw && rw);
    h.k_by_len.resize((size_t)L + 1); h.thr_ins.resize((size_t)L + 1); h.maxmm.resize((size_t)amax + 1);
    for (int l = 0; l <= L; l++) {
        h.k_by_len[l] = (unsigned short)std::min(60000, (int)(rate * l));            // _align.pyx:634
        h.thr_ins[l] = thr_mul_of(l, rate);                                          // _align.pyx:728
    }
    std::fesetround(FE_TONEAREST);
    for (int a = 0; a <= amax; a++) {                                                // Python round(): half to even
        const double r = std::nearbyint(a * d.max_adapter_mismatch_frac);
        h.maxmm[a] = (unsigned short)std::max(0.0, std::min(60000.0, r));
    }
    bool packed_ok = true;
    auto enc = [&](const char* s, int n, std::vector<unsigned char>& code, std::vector<unsigned char>& asc) {
        code.resize((size_t)n); asc.resize((size_t)n);
        for (int i = 0; i < n; i++) {
            const unsigned char c = (unsigned char)s[i];
            if (!v.and_mode) { code[i] = tb.iupac[c]; asc[i] = c; packed_ok = packed_ok && exact_symbol(c); }
            else { code[i] = asc[i] = rw ? tb.iupac[c] : tb.acgt[c]; }
        }
    };
    enc(d.adapter1, d.adapter1_len, h.a1_code, h.a1_ascii);
    enc(d.adapter2, d.adapter2_len, h.a2_code, h.a2_ascii);
    auto packw = [](const std::vector<unsigned char>& code, std::vector<uint32_t>& out) {
        out.assign(code.size() / 8 + 2, 0u);
        for (size_t i = 0; i < code.size(); i++) out[i >> 3] |= ((uint32_t)code[i] & 15u) << (4 * (i & 7));
    };
    packw(h.a1_code, h.a1_pack);
    packw(h.a2_code, h.a2_pack);
    v.packed_ok = packed_ok;
    h.comp.assign(256, 0); h.ov_tab.assign(256, 0);
    const char* a = "ACRSWKBDN"; const char* b = "TGYSWMVHN";                         // util/__init__.py:67-88
    for (int i = 0; a[i]; i++) {
        h.comp[(unsigned char)a[i]] = (unsigned char)b[i]; h.comp[(unsigned char)b[i]] = (unsigned char)a[i];
        h.comp[(unsigned char)(a[i] + 32)] = (unsigned char)(b[i] + 32);
        h.comp[(unsigned char)(b[i] + 32)] = (unsigned char)(a[i] + 32);
    }
    for (int c = 0; c < 256; c++) h.ov_tab[c] = !v.and_mode ? (unsigned char)c : (aw ? tb.iupac[c] : tb.acgt[c]);
    h.insert_prob.assign(d.insert_prob, d.insert_prob + (size_t)(L + 1) * (v.kmax + 1));
    h.adapter_prob.assign(d.adapter_prob, d.adapter_prob + (size_t)(amax + 1) * (amax + 1));
    return ATR_OK;
}

// ---- MergeOverlapping(min_overlap, error_rate) (commands/trim/modifiers.py:867-879): per-length tables -----------
// h = [thr_mul (max_len + 1) | minov (max_len + 1)], comp = BASE_COMPLEMENTS (util/__init__.py:67-88; 0 = KeyError)
inline void build_merge_tables(int max_len, double min_overlap, double error_rate, std::vector<unsigned short>& h,
                               unsigned char comp[256]) {
    h.assign((size_t)2 * (max_len + 1), 0);
    std::fesetround(FE_TONEAREST);
    // __init__ keeps int(min_overlap) if min_overlap > 1 (:871); __call__ then treats a value <= 1 as a fraction
    // of the shorter read: max(2, round(...)), Python's round = half to even (:877-879)
    const double v = min_overlap > 1 ? std::floor(min_overlap) : min_overlap;
    for (int l = 0; l <= max_len; l++) {
        h[l] = thr_mul_of(l, error_rate);
        const double mo = v <= 1 ? std::max(2.0, std::nearbyint(v * l)) : v;
        h[(size_t)max_len + 1 + l] = (unsigned short)std::min(60000.0, mo);
    }
    memset(comp, 0, 256);
    const char* a = "ACRSWKBDN";
    const char* b = "TGYSWMVHN";
    for (int i = 0; a[i]; i++) {
        comp[(unsigned char)a[i]] = (unsigned char)b[i]; comp[(unsigned char)b[i]] = (unsigned char)a[i];
        comp[(unsigned char)(a[i] + 32)] = (unsigned char)(b[i] + 32); comp[(unsigned char)(b[i] + 32)] = (unsigned char)(a[i] + 32);
    }
}

}  // namespace atr
