// insert_core.cuh -- per-pair logic of the K2 kernels (one thread = one read pair).
//
// Re-designs InsertAligner.match_insert (reference: atropos/align/__init__.py:250-377) and the
// MultiAligner.locate call inside it (atropos/align/_align.pyx:593-772) for a GPU thread.
// With the insert flag set (START_WITHIN_SEQ1|STOP_WITHIN_SEQ2), equal lengths, no indels and ASCII
// equality, the reference's DP is exactly a sliding Hamming distance: for overlap length j = 1..m the
// candidate cell (m, j) has cost Hamming(rc(read2)[m-j:m], read1[0:j]), origin -(m-j), matches j-cost
// (see DESIGN.md for the argument that the Ukkonen band never hides such a cell). On 4-bit codes
//   * reverse-complement of 8 bases is ONE instruction: __brev(word) reverses the nibble order and
//     bit-reverses every nibble (A1<->T8, C2<->G4, IUPAC unions map to their complements, N=15 fixed);
//   * the Hamming distance of 8 bases is XOR, OR-fold, AND 0x11111111, POPC.
// Everything floating point is a host-built table (see atr_api.cu: insertset): k per read length,
// floor(j*rate), round(alen*frac), and the two random-match-probability tables.
// __host__ __device__ so tests/host_sim can run it on the CPU build box.
#pragma once
#include "atr_common.cuh"

#define ATR_K2_MAXW 40             // packed fast path: reads up to 304 nt (38 4-bit words + 2 guard words)
#define ATR_K2_MAXLEN 304
#define ATR_MAX_CAND 100           // MultiAligner.locate(max_matches=100) (_align.pyx:593)

struct InsertDev {
    int min_insert_overlap, min_adapter_overlap, cutoff;
    int kmax, max_len;              // tables cover read lengths 0..max_len, costs 0..kmax
    int a1_len, a2_len, amax;
    int and_mode;                   // overhang-vs-adapter compare: 0 equality, 1 (a & b) != 0
    int ov_single_only;             // overhang under the ACGT table (read_wildcards && !adapter_wildcards)
    int packed_ok;                  // adapters representable for the packed path
    double insert_max_rmp, adapter_max_rmp;
    const unsigned short* k_by_len; // [max_len+1]  (int)(rate * m)                       _align.pyx:634
    const unsigned short* thr_ins;  // [max_len+1]  max c with c <= j * rate               _align.pyx:728
    const unsigned short* maxmm;    // [amax+1]     round(alen * max_adapter_mismatch_frac) align/__init__.py:290
    const unsigned char* a1_code;   // packed-path compare operand per adapter base
    const unsigned char* a2_code;
    const uint32_t* a1_pack;        // the same codes packed 8 per word like a read (word-wise overhang compare)
    const uint32_t* a2_pack;
    const unsigned char* a1_ascii;  // byte path: adapter bytes / translated per mode
    const unsigned char* a2_ascii;
    const double* insert_prob;      // [(max_len+1) * (kmax+1)]  P(matches = size - cost, size)   align/__init__.py:358
    const double* adapter_prob;     // [(amax+1) * (amax+1)]     P(matches, alen)                 align/__init__.py:303-304
    const unsigned char* comp;      // [256] byte complement, 0 = KeyError                        util/__init__.py:67-88
    const unsigned char* ov_tab;    // [256] byte-path translation of the read overhang (identity in ASCII mode)
};

struct Cand { unsigned short j, cost; };
struct GCellM { long long cost; int origin, matches; };

ATR_HD int atr_imin(int a, int b) { return a < b ? a : b; }

ATR_HD void im_clear(atr_match& m) {
    m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0;
    m.adapter = -1; m.status = ATR_ST_NONE;
}

ATR_HD unsigned nib_mismatches(uint32_t x) {          // number of non-zero nibbles
    x |= x >> 1; x |= x >> 2;
    x &= 0x11111111u;
#if defined(__CUDA_ARCH__)
    return (unsigned)__popc(x);
#else
    return (unsigned)__builtin_popcount(x);
#endif
}

ATR_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

ATR_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, unsigned shift) {   // (hi:lo) >> shift, shift in [0,32)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    return shift == 0 ? lo : ((lo >> shift) | (hi << (32 - shift)));
#endif
}

// ---- the packed pair: 2-bit look-ahead filter + exact 4-bit verification ------------------------------------------
// The sliding Hamming distance is evaluated for every overlap length j, and almost every j is a random alignment that
// fails within a few bases. The scan therefore runs on a 2-BIT recoding of both sequences (A 0, C 1, G 2, T 3; N and
// the other IUPAC codes fall on one of these): 16 bases per word, mismatches of 16 bases = one funnel shift, XOR,
// fold, POPC. The recoding is a function of the 4-bit code, so differing 2-bit codes imply differing bases: the
// 2-bit count is a LOWER BOUND of the true cost and never rejects a real candidate. An overlap that survives the
// look-ahead (the first 16..80 bases, more where the error budget is larger) is verified with the exact 4-bit
// comparison straight from the packed reads in global memory (the reverse complement of 8 bases is one __brev).
//   rc(read2[:m]) in 2-bit words lives in shared memory ([word][thread]); the first 80 bases of read 1 in registers.
#define ATR_K2_MAXW2 22            // 2-bit words of rc(read 2): 304 nt = 19 words + guard
#define ATR_K2_QW 5                // look-ahead: up to 5 words = 80 bases

ATR_HD uint32_t conv2(uint32_t w) {                   // 8 bases as 4-bit codes -> 16 bits of 2-bit codes
    uint32_t t = ((w >> 1) | (w >> 3)) & 0x11111111u;
    t |= (((w >> 2) | (w >> 3)) & 0x11111111u) << 1;
    t = (t | (t >> 2)) & 0x0F0F0F0Fu;
    t = (t | (t >> 4)) & 0x00FF00FFu;
    return (t | (t >> 8)) & 0xFFFFu;
}
ATR_HD int k2_msb(unsigned x) {                       // index of the highest set bit (x != 0)
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
ATR_HD unsigned mism2(uint32_t x) {                   // differing 2-bit fields of x = a ^ b
    x = (x | (x >> 1)) & 0x55555555u;
#if defined(__CUDA_ARCH__)
    return (unsigned)__popc(x);
#else
    return (unsigned)__builtin_popcount(x);
#endif
}
// look-ahead words for an error budget: the expected 12 mismatches per random 16 bases must clear it by >= 3.5 sigma
ATR_HD int k2_words_for(int bound) { return bound <= 6 ? 1 : bound <= 15 ? 2 : bound <= 24 ? 3 : bound <= 33 ? 4 : 5; }

struct PackedPair {
    uint32_t* R2;           // rc(seq2[:m]) in 2-bit words, element u at R2[u * stride], zero padded by 2 words
    uint32_t q2[ATR_K2_QW]; // seq1[0:80] in 2-bit words
    const uint32_t* S2;     // forward read2 packed (4-bit)
    const uint32_t* S1;
    const unsigned short* thr;   // thr_ins (shared-memory copy in the kernel)
    int stride;
    int m, W, pad;
    ATR_HD uint32_t rc_word(int i) const {             // word i of rc(seq2[:m]) in 4-bit codes; zero beyond the sequence
        if (i >= W) return 0u;
        uint32_t a = S2[W - 1 - i];
        if (i == 0 && (m & 7)) a &= (1u << ((m & 7) * 4)) - 1u;              // drop bases beyond m in the last word
        const uint32_t b = (i + 1 < W) ? S2[W - 2 - i] : 0u;
        return funnel_r(brev32(a), brev32(b), (unsigned)pad * 4u);
    }
    // exact Hamming(rc2[m-j:m], seq1[0:j]), abandoned once it exceeds `bound`. Out of line: it runs for the few
    // overlaps that survive the look-ahead, and inlined at every call site it multiplied the kernel's code size
    ATR_HD_NOINLINE int overlap_cost(int j, int bound) const {
        const int s = m - j, ws = s >> 3;
        const unsigned bs = (unsigned)(s & 7) * 4u;
        const int nfull = j >> 3;
        unsigned cost = 0;
        int w = 0;
        uint32_t lo = rc_word(ws);
        for (; w < nfull; w++) {
            const uint32_t hi = rc_word(ws + w + 1);
            cost += nib_mismatches(funnel_r(lo, hi, bs) ^ S1[w]);
            lo = hi;
            if ((int)cost > bound) return (int)cost;
        }
        if (j & 7) {
            const uint32_t hi = rc_word(ws + w + 1);
            const uint32_t x = (funnel_r(lo, hi, bs) ^ S1[w]) & ((1u << ((j & 7) * 4)) - 1u);
            cost += nib_mismatches(x);
        }
        return (int)cost;
    }
    // lower bound of the cost of overlap j from its first min(j, 32) bases
    ATR_HD int small_lb(int j) const {
        const int s = m - j, us = s >> 4;
        const unsigned b2 = (unsigned)(s & 15) * 2u;
        const uint32_t r0 = R2[us * stride], r1 = R2[(us + 1) * stride];
        if (j <= 16) {
            const uint32_t mask = j == 16 ? 0xFFFFFFFFu : ((1u << (2 * j)) - 1u);
            return (int)mism2((funnel_r(r0, r1, b2) ^ q2[0]) & mask);
        }
        const uint32_t r2 = R2[(us + 2) * stride];
        const uint32_t mask = j == 32 ? 0xFFFFFFFFu : ((1u << (2 * (j - 16))) - 1u);
        return (int)(mism2(funnel_r(r0, r1, b2) ^ q2[0]) + mism2((funnel_r(r1, r2, b2) ^ q2[1]) & mask));
    }
    // 16 overlaps that share one 2-bit word offset `us` (shifts b = 15..0 within the word), NW look-ahead words, one
    // error budget for the whole group (that of its longest overlap: the look-ahead is a necessary condition only, the
    // verification applies each overlap's own budget). Branch-free: returns the shifts that stay within the budget.
    template <int NW>
    ATR_HD unsigned group(int us, int gbound) const {
        uint32_t r[NW + 1];
#pragma unroll
        for (int t = 0; t <= NW; t++) r[t] = R2[(us + t) * stride];
        unsigned hits = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            unsigned c = 0;
#pragma unroll
            for (int t = 0; t < NW; t++) c += mism2(funnel_r(r[t], r[t + 1], 2u * (unsigned)b) ^ q2[t]);
            hits |= ((int)c <= gbound ? 1u : 0u) << b;
        }
        return hits;
    }
    ATR_HD unsigned group_nw(int nw, int us, int gbound) const {
        switch (nw) {
            case 1: return group<1>(us, gbound);
            case 2: return group<2>(us, gbound);
            case 3: return group<3>(us, gbound);
            case 4: return group<4>(us, gbound);
            default: return group<5>(us, gbound);
        }
    }
    // Enumerate the overlap lengths j = 1..m in ascending order and call emit(j, cost) for every j >= min_overlap
    // with cost <= min(k, thr[j]); emit returns false to stop (_align.pyx:722-745).
    // Two nested stages with ONE call site each (the kernel is otherwise bound by instruction fetch): the look-ahead
    // marks survivors (the real overlap plus noise; everything at once on low-complexity reads), which are verified in
    // ascending j in batches of ATR_K2_SURV.
#define ATR_K2_SURV 16
    template <class F>
    ATR_HD void scan(const InsertDev& d, int k, F&& emit) const {
        unsigned short surv[ATR_K2_SURV];
        int ns = 0;
        bool go_on = true;
        auto flush = [&]() {
            for (int t = 0; t < ns && go_on; t++) {
                const int j = (int)surv[t];
                const int bound = atr_imin(k, (int)thr[j]);
                const int full = overlap_cost(j, bound);
                if (full <= bound) go_on = emit(j, full);
            }
            ns = 0;
        };
        const int jmin = d.min_insert_overlap > 1 ? d.min_insert_overlap : 1;
        const int jsmall = m < 31 ? m : 31;
        for (int j = jmin; j <= jsmall; j++) {
            if (small_lb(j) <= atr_imin(k, (int)thr[j])) {
                surv[ns++] = (unsigned short)j;
                if (ns == ATR_K2_SURV) { flush(); if (!go_on) return; }
            }
        }
        const int jfirst = jmin > 32 ? jmin : 32;              // shortest overlap of the word-wise part
        if (m >= jfirst) {
            const int s_first = m - jfirst;                    // shift of the shortest overlap handled here
            for (int us = s_first >> 4; us >= 0 && go_on; us--) {
                const int bhi = us == (s_first >> 4) ? (s_first & 15) : 15;       // the first group may be partial
                const int gbound = atr_imin(k, (int)thr[m - 16 * us]);
                const int nw = atr_imin(k2_words_for(gbound), (m - (16 * us + bhi)) >> 4);    // never past the shortest overlap
                unsigned hits = group_nw(nw, us, gbound) & ((2u << bhi) - 1u);
                while (hits) {                                 // ascending overlap length = descending shift
                    const int b = k2_msb(hits);
                    hits &= ~(1u << b);
                    surv[ns++] = (unsigned short)(m - (16 * us + b));
                    if (ns == ATR_K2_SURV) { flush(); if (!go_on) break; }
                }
            }
        }
        if (go_on) flush();
    }
    ATR_HD unsigned ov1(int p) const { return (S1[p >> 3] >> ((p & 7) * 4)) & 15u; }
    ATR_HD unsigned ov2(int p) const { return (S2[p >> 3] >> ((p & 7) * 4)) & 15u; }
    ATR_HD const uint32_t* fwd1() const { return S1; }
    ATR_HD const uint32_t* fwd2() const { return S2; }
};


struct BytePair {
    const unsigned char* s1;   // read1 bytes
    const unsigned char* s2;   // read2 bytes (forward)
    const unsigned char* comp;
    const unsigned char* ov_tab;
    int m;
    ATR_HD int overlap_cost(int j, int k) const {      // ref[i] = comp[s2[m-1-i]]
        int cost = 0;
        for (int t = 0; t < j && cost <= k; t++) cost += (comp[s2[j - 1 - t]] != s1[t]);      // ref[m-j+t] = comp[s2[j-1-t]]
        return cost;
    }
    template <class F>
    ATR_HD void scan(const InsertDev& d, int k, F&& emit) const {
        for (int j = 1; j <= m; j++) {
            if (j < d.min_insert_overlap) continue;
            const int bound = atr_imin(k, (int)d.thr_ins[j]);
            const int cost = overlap_cost(j, bound);
            if (cost <= bound && !emit(j, cost)) return;
        }
    }
    ATR_HD unsigned ov1(int p) const { return ov_tab[s1[p]]; }
    ATR_HD unsigned ov2(int p) const { return ov_tab[s2[p]]; }
    ATR_HD const uint32_t* fwd1() const { return nullptr; }
    ATR_HD const uint32_t* fwd2() const { return nullptr; }
};

// compare_prefixes(read[size:], adapter) (align/__init__.py:285-288): mismatches over alen bases
template <class P, bool FIRST>
ATR_HD int overhang_mismatches(const InsertDev& d, const P& pr, int size, int alen, bool packed) {
    const unsigned char* ac = packed ? (FIRST ? d.a1_code : d.a2_code) : (FIRST ? d.a1_ascii : d.a2_ascii);
    int mm = 0;
    for (int t = 0; t < alen; t++) {
        unsigned rc = FIRST ? pr.ov1(size + t) : pr.ov2(size + t);
        if (packed && d.ov_single_only) rc = (rc & (rc - 1)) ? 0u : rc;
        const unsigned a = ac[t];
        mm += d.and_mode ? ((rc & a) == 0u) : (rc != a);
    }
    return mm;
}

// the same on packed words, 8 bases per step (packed path, not the ACGT-filtered overhang mode)
template <bool FIRST>
ATR_HD int overhang_mismatches_words(const InsertDev& d, const uint32_t* __restrict__ S, int size, int alen) {
    const uint32_t* ap = FIRST ? d.a1_pack : d.a2_pack;
    const unsigned sh = (unsigned)(size & 7) * 4u;
    int mm = 0;
    for (int w = 0; w * 8 < alen; w++) {
        const int rows = alen - 8 * w < 8 ? alen - 8 * w : 8;
        const uint32_t w0 = S[(size >> 3) + w];
        const uint32_t w1 = ((size & 7) + rows) > 8 ? S[(size >> 3) + w + 1] : 0u;     // only if the bases spill over
        const uint32_t rd = funnel_r(w0, w1, sh);
        const uint32_t mask = rows == 8 ? 0x11111111u : (0x11111111u & ((1u << (4 * rows)) - 1u));
        uint32_t x = d.and_mode ? (rd & ap[w]) : (rd ^ ap[w]);
        x |= x >> 1; x |= x >> 2;                      // bit 4t set <=> nibble t non-zero
        x = d.and_mode ? (~x & mask) : (x & mask);     // AND mode: a mismatch is a nibble with no common bit
#if defined(__CUDA_ARCH__)
        mm += __popc(x);
#else
        mm += __builtin_popcount(x);
#endif
    }
    return mm;
}

// InsertAligner.match_insert._match (align/__init__.py:269-320) for one candidate.
// returns 0: None, 1: result written
template <class P>
ATR_HD int insert_try(const InsertDev& d, const P& pr, bool packed, int m, int j, int cost, int len1, int len2,
                      atr_insert_result* out) {
    const int offset = m - j, size = j;
    atr_insert_result r;
    im_clear(r.insert); im_clear(r.match1); im_clear(r.match2);
    r.insert.astart = (uint16_t)(m - j); r.insert.astop = (uint16_t)m; r.insert.rstart = 0; r.insert.rstop = (uint16_t)j;
    r.insert.matches = (uint16_t)(j - cost); r.insert.errors = (uint16_t)cost; r.insert.status = ATR_ST_MATCH;
    if (offset < d.min_adapter_overlap) { *out = r; return 1; }              // (insert_match, None, None)
    const int alen1 = offset < d.a1_len ? offset : d.a1_len;
    const int alen2 = offset < d.a2_len ? offset : d.a2_len;
    int mm1, mm2;
    if (packed && !d.ov_single_only) {
        mm1 = overhang_mismatches_words<true>(d, pr.fwd1(), size, alen1);
        mm2 = overhang_mismatches_words<false>(d, pr.fwd2(), size, alen2);
    } else {
        mm1 = overhang_mismatches<P, true>(d, pr, size, alen1, packed);
        mm2 = overhang_mismatches<P, false>(d, pr, size, alen2, packed);
    }
    if (mm1 > (int)d.maxmm[alen1] && mm2 > (int)d.maxmm[alen2]) return 0;    // :297-300
    if ((alen1 < alen2 ? alen1 : alen2) > d.cutoff) {                        // :302-306
        const double p1 = d.adapter_prob[alen1 * (d.amax + 1) + (alen1 - mm1)];
        const double p2 = d.adapter_prob[alen2 * (d.amax + 1) + (alen2 - mm2)];
        if (p1 * p2 > d.adapter_max_rmp) return 0;
    }
    const int mism = mm1 < mm2 ? mm1 : mm2;
    for (int which = 0; which < 2; which++) {                                // _create_match :310-314
        const int slen = which ? len2 : len1;
        int a = which ? alen2 : alen1;
        if (slen - size < a) a = slen - size;
        const int e = a < mism ? a : mism;
        atr_match& mt = which ? r.match2 : r.match1;
        mt.astart = 0; mt.astop = (uint16_t)(a > 0 ? a : 0); mt.rstart = (uint16_t)size; mt.rstop = (uint16_t)slen;
        mt.matches = (uint16_t)(a - e > 0 ? a - e : 0); mt.errors = (uint16_t)e; mt.adapter = (int16_t)which;
        mt.status = (a <= 0 || a - e <= 0) ? ATR_ST_INVALID : ATR_ST_MATCH;  // Match.__init__ :85-88
    }
    *out = r;
    return 1;
}

// The whole decision procedure for one pair. `cand` = per-thread scratch for ATR_MAX_CAND candidates.
template <class P>
ATR_HD void insert_pair(const InsertDev& d, const P& pr, bool packed, int m, int len1, int len2, Cand* cand,
                        atr_insert_result* out) {
    atr_insert_result none;
    im_clear(none.insert); im_clear(none.match1); im_clear(none.match2);
    *out = none;
    if (m <= 0) return;
    const int k = d.k_by_len[m];
    int count = 0;
    // Appendix B of SURVEY.md: a candidate needs cost <= k AND cost <= floor(j * rate) (_align.pyx:722-728)
    pr.scan(d, k, [&](int j, int cost) -> bool {
        if (cost == 0 && j == m) { cand[0].j = (unsigned short)j; cand[0].cost = 0; count = 1; return false; }   // [exact]
        cand[count].j = (unsigned short)j; cand[count].cost = (unsigned short)cost;
        count++;
        return count < ATR_MAX_CAND;
    });
    // (the reference may append the j == m candidate a second time, :746-763; a duplicate cannot change the outcome)
    if (count == 0) return;
    // random-match-probability filter, then candidates in order of probability (stable) :353-375
    unsigned long long done_lo = 0, done_hi = 0;       // visited / rejected bitmap over <= 100 candidates
    int alive = 0;
    for (int c = 0; c < count; c++) {
        const double p = d.insert_prob[(int)cand[c].j * (d.kmax + 1) + (int)cand[c].cost];
        if (p <= d.insert_max_rmp) alive++;
        else { if (c < 64) done_lo |= 1ull << c; else done_hi |= 1ull << (c - 64); }
    }
    while (alive > 0) {
        int bestc = -1; double bestp = 0.0;
        for (int c = 0; c < count; c++) {
            const bool dn = c < 64 ? ((done_lo >> c) & 1ull) : ((done_hi >> (c - 64)) & 1ull);
            if (dn) continue;
            const double p = d.insert_prob[(int)cand[c].j * (d.kmax + 1) + (int)cand[c].cost];
            if (bestc < 0 || p < bestp) { bestc = c; bestp = p; }
        }
        if (insert_try(d, pr, packed, m, cand[bestc].j, cand[bestc].cost, len1, len2, out)) return;
        if (bestc < 64) done_lo |= 1ull << bestc; else done_hi |= 1ull << (bestc - 64);
        alive--;
    }
}

// Build the packed operands of a pair. S1/S2: forward packed reads (4-bit); m = min(len1, len2); W1 = words of read 1.
// Returns 0 if read2[:m] contains code 0 ('X': reverse_complement raises KeyError) -> byte path decides.
ATR_HD int packed_pair_setup(PackedPair& pp, const uint32_t* S1, const uint32_t* S2, int m, int W1) {
    const int W = (m + 7) >> 3;
    const int st = pp.stride;
    int ok = 1;
    for (int w = 0; w < W; w++) {                      // zero-nibble scan of read2[:m]
        const uint32_t x = S2[w];
        const uint32_t valid = (w == W - 1 && (m & 7)) ? ((1u << ((m & 7) * 4)) - 1u) : 0xFFFFFFFFu;
        uint32_t nz = x | (x >> 1); nz |= nz >> 2; nz &= valid & 0x11111111u;
        if (nz != (valid & 0x11111111u)) ok = 0;
    }
    pp.S1 = S1; pp.S2 = S2; pp.m = m; pp.W = W; pp.pad = 8 * W - m;
    // rc(read2[:m]): reversed words, each bit-reversed (A1<->T8, C2<->G4), shifted so that base 0 sits in nibble 0
    const int U = (W + 1) >> 1;
    for (int u = 0; u < U; u++) pp.R2[u * st] = conv2(pp.rc_word(2 * u)) | (conv2(pp.rc_word(2 * u + 1)) << 16);
    pp.R2[U * st] = 0; pp.R2[(U + 1) * st] = 0;
    if (U + 2 < ATR_K2_MAXW2) pp.R2[(U + 2) * st] = 0;
#pragma unroll
    for (int t = 0; t < ATR_K2_QW; t++) {
        const uint32_t a = 2 * t < W1 ? S1[2 * t] : 0u, b = 2 * t + 1 < W1 ? S1[2 * t + 1] : 0u;
        pp.q2[t] = conv2(a) | (conv2(b) << 16);
    }
    return ok;
}

// ---- MultiAligner.locate for ANY flag set (single-call API; not the batched hot path) -----------
// Follows _align.pyx:593-772 literally (banded, diagonal-only recurrence, 100000-per-base overhang
// costs, candidate cap, [exact] collapse, last-column scan in the for/else). col: m+1 cells.
// out6: up to max_matches + m + 2 tuples. Returns the number of tuples.
ATR_HD int gen_multi_locate(const unsigned char* ref, int m, const unsigned char* query, int n, int k,
                            const unsigned short* thr /* floor(l*rate), l = 0..m */, int flags, int min_overlap,
                            int max_matches, GCellM* col, int* out6) {
    const int OVER = 100000;
    const bool start_in_ref = flags & ATR_START_WITHIN_SEQ1, start_in_query = flags & ATR_START_WITHIN_SEQ2;
    const bool stop_in_ref = flags & ATR_STOP_WITHIN_SEQ1, stop_in_query = flags & ATR_STOP_WITHIN_SEQ2;
    int max_n = n, min_n = 0;
    if (!start_in_query) max_n = n < m + k ? n : m + k;
    if (!stop_in_query) min_n = n - m - k > 0 ? n - m - k : 0;
    const long long max_cost = (long long)m + n;
    for (int i = 0; i <= m; i++) {
        long long cost; int origin;
        if (!start_in_ref && !start_in_query) { cost = (long long)(i > min_n ? i : min_n) * OVER; origin = 0; }
        else if (start_in_ref && !start_in_query) { cost = (long long)min_n * OVER; origin = (min_n - i) < 0 ? (min_n - i) : 0; }
        else if (!start_in_ref && start_in_query) { cost = (long long)i * OVER; origin = (min_n - i) > 0 ? (min_n - i) : 0; }
        else { cost = (long long)(i < min_n ? i : min_n) * OVER; origin = min_n - i; }
        col[i].cost = cost; col[i].origin = origin; col[i].matches = 0;
    }
    int last = start_in_ref ? m : (m < k + 1 ? m : k + 1);
    int count = 0, exact = -1;
    bool broke = false;
    for (int j = min_n + 1; j <= max_n; j++) {
        GCellM diag = col[0];
        if (start_in_query) col[0].origin = j; else col[0].cost = (long long)j * OVER;
        for (int i = 1; i <= last; i++) {
            GCellM nw;
            nw.origin = diag.origin;
            if (ref[i - 1] == query[j - 1]) { nw.cost = diag.cost; nw.matches = diag.matches + 1; }
            else { nw.cost = diag.cost + 1; nw.matches = diag.matches; }
            diag = col[i];
            col[i] = nw;
        }
        while (last >= 0 && col[last].cost > k) last--;
        if (last < m) { last++; }
        else if (stop_in_query) {
            const long long cost = col[m].cost;
            if (cost > max_cost) continue;
            const int length = m + (col[m].origin < 0 ? col[m].origin : 0);
            if (length >= min_overlap && cost <= (long long)thr[length]) {
                int* o = out6 + 6 * count;
                const int org = col[m].origin;
                o[0] = org < 0 ? -org : 0; o[1] = m; o[2] = org < 0 ? 0 : org; o[3] = j; o[4] = col[m].matches; o[5] = (int)cost;
                if (cost == 0 && col[m].matches == m) { exact = count; count++; broke = true; break; }
                count++;
                if (count >= max_matches) { broke = true; break; }
            }
        }
    }
    if (!broke && max_n == n) {
        for (int i = stop_in_ref ? 0 : m; i <= m; i++) {
            const long long cost = col[i].cost;
            if (cost > max_cost) continue;
            const int length = i + (col[i].origin < 0 ? col[i].origin : 0);
            if (length >= min_overlap && cost <= (long long)thr[length]) {
                int* o = out6 + 6 * count;
                const int org = col[i].origin;
                o[0] = org < 0 ? -org : 0; o[1] = i; o[2] = org < 0 ? 0 : org; o[3] = n; o[4] = col[i].matches; o[5] = (int)cost;
                count++;
            }
        }
    }
    if (count && exact >= 0) {
        for (int t = 0; t < 6; t++) out6[t] = out6[6 * exact + t];
        return 1;
    }
    return count;
}
