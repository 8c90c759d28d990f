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

#define ATR_K2_MAXW 40             // packed fast path: reads up to 304 nt (38 words + 2 guard words)
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

// Accessors give the pair's two sequences to the shared decision code.
//   PackedPair: R/Q are per-thread word arrays with element w at [w * stride] (shared memory on the GPU)
#ifndef ATR_K2_INLINE_THR
#define ATR_K2_INLINE_THR 18
#endif
// INLINE_HIGH selects, at compile time, the scan that finishes high-bound overlaps inline (see scan_impl): the kernel
// for insert sets whose bounds stay below ATR_K2_INLINE_THR keeps the leaner loop.
template <bool INLINE_HIGH>
struct PackedPairT {
    uint32_t* R;            // rc(seq2[:m]) packed, >= W+2 words, zero padded
    uint32_t* Q;            // seq1 packed (only the first m bases are looked at)
    const uint32_t* S2;     // forward read2 packed (for the overhang)
    const uint32_t* S1;
    int stride;
    int m;
    // Hamming(R[m-j:m], Q[0:j]) continued from whole word w0 (cost so far c0), abandoned once it exceeds `bound`
    ATR_HD int overlap_cost(int j, int bound, int w0 = 0, unsigned c0 = 0) const {
        const int s = m - j, ws = s >> 3;
        const unsigned bs = (unsigned)(s & 7) * 4u;
        const int nfull = j >> 3;                       // whole words
        unsigned cost = c0;
        int w = w0;
        uint32_t lo = R[(ws + w) * stride];
        for (; w < nfull; w++) {
            const uint32_t hi = R[(ws + w + 1) * stride];
            cost += nib_mismatches(funnel_r(lo, hi, bs) ^ Q[w * stride]);
            lo = hi;
            if ((int)cost > bound) return (int)cost;
        }
        if (j & 7) {
            const uint32_t hi = R[(ws + w + 1) * stride];
            const uint32_t x = (funnel_r(lo, hi, bs) ^ Q[w * stride]) & ((1u << ((j & 7) * 4)) - 1u);
            cost += nib_mismatches(x);
        }
        return (int)cost;
    }

    // Enumerate the overlap lengths j = 1..m in ascending order and call emit(j, cost) for every j >= min_overlap
    // with cost <= min(k, thr[j]); emit returns false to stop (_align.pyx:722-745).
    // Overlaps of >= 32 bases are handled in groups of 8 shifts that share the same word offset: the 5 words of
    // rc(read2) a group needs and the first 4 words of read 1 stay in registers, the funnel shifts are static,
    // and the first 16 or 32 bases are compared unconditionally (no per-word exit test, no shared-memory
    // traffic). A random overlap is almost surely over its bound after that; only real overlaps continue with
    // the word-by-word scan.
    template <class F>
    ATR_HD void scan(const InsertDev& d, int k, F&& emit) const {
        const int jsmall = m < 31 ? m : 31;
        for (int j = 1; j <= jsmall; j++) {
            if (j < d.min_insert_overlap) continue;
            const int bound = atr_imin(k, (int)d.thr_ins[j]);
            const int cost = overlap_cost(j, bound);
            if (cost <= bound && !emit(j, cost)) return;
        }
        if (m < 32) return;
        const uint32_t q0 = Q[0], q1 = Q[stride], q2 = Q[2 * stride], q3 = Q[3 * stride];
        int g = (m - 32) >> 3;                         // word offset of the first (shortest) overlap handled here
        int bstart = (m - 32) & 7;
        uint32_t r0 = R[g * stride], r1 = R[(g + 1) * stride], r2 = R[(g + 2) * stride], r3 = R[(g + 3) * stride],
                 r4 = R[(g + 4) * stride];
        // Overlaps that survive the unconditional part are rare (the real one, plus noise): they are parked and
        // finished after the loop, all lanes together, instead of one lane at a time in the middle of it.
        // pend[] holds (j << 16 | words done << 8 | cost so far), in ascending j.
        unsigned pend0 = 0, pend1 = 0, pend2 = 0, pend3 = 0;
        int npend = 0;
        bool go_on = true;
        auto finish = [&](unsigned e) -> bool {        // complete one parked overlap; false = stop everything
            const int j = (int)(e >> 16), wdone = (int)((e >> 8) & 255u);
            const int bound = atr_imin(k, (int)d.thr_ins[j]);
            const int full = overlap_cost(j, bound, wdone, e & 255u);
            return !(full <= bound) || emit(j, full);
        };
        auto flush = [&]() -> bool {
            if (npend > 0 && !finish(pend0)) return false;
            if (npend > 1 && !finish(pend1)) return false;
            if (npend > 2 && !finish(pend2)) return false;
            if (npend > 3 && !finish(pend3)) return false;
            npend = 0;
            return true;
        };
        for (; g >= 0 && go_on; g--) {
#pragma unroll
            for (int b = 7; b >= 0; b--) {
                if (b > bstart || !go_on) continue;
                const int j = m - (8 * g + b);
                const int tj = (int)d.thr_ins[j];
                const int bound = atr_imin(k, tj);
                unsigned cost = nib_mismatches(funnel_r(r0, r1, 4u * b) ^ q0) + nib_mismatches(funnel_r(r1, r2, 4u * b) ^ q1);
                unsigned wdone = 2;
                if (tj > 6) {                          // long overlaps tolerate more mismatches: look at 32 bases
                    cost += nib_mismatches(funnel_r(r2, r3, 4u * b) ^ q2) + nib_mismatches(funnel_r(r3, r4, 4u * b) ^ q3);
                    wdone = 4;
                }
                if (INLINE_HIGH && (int)cost <= bound && j >= d.min_insert_overlap && tj >= ATR_K2_INLINE_THR) {
                    // A bound this high cannot be exceeded within 32 bases often enough (random bases mismatch at
                    // 3/4: 24 +- 2.4 of 32), so nearly every lane would park nearly every overlap: finish it right
                    // here, word by word with the exit test -- all lanes are in the same situation, so this does
                    // not diverge. A real candidate is emitted after the parked (shorter) ones, in order.
                    const int full = overlap_cost(j, bound, (int)wdone, cost);
                    if (full <= bound) {
                        go_on = flush();
                        if (go_on) go_on = emit(j, full);
                    }
                } else if ((int)cost <= bound && j >= d.min_insert_overlap) {
                    if (npend == 4) go_on = flush();   // full (low-complexity read): finish the parked ones in order
                    if (go_on) {
                        const unsigned e = ((unsigned)j << 16) | (wdone << 8) | cost;
                        if (npend == 0) pend0 = e; else if (npend == 1) pend1 = e; else if (npend == 2) pend2 = e; else pend3 = e;
                        npend++;
                    }
                }
            }
            bstart = 7;
            r4 = r3; r3 = r2; r2 = r1; r1 = r0;
            if (g > 0) r0 = R[(g - 1) * stride];
        }
        if (go_on) flush();
    }
    ATR_HD unsigned ov1(int p) const { return (S1[p >> 3] >> ((p & 7) * 4)) & 15u; }
    ATR_HD unsigned ov2(int p) const { return (S2[p >> 3] >> ((p & 7) * 4)) & 15u; }
    ATR_HD const uint32_t* fwd1() const { return S1; }
    ATR_HD const uint32_t* fwd2() const { return S2; }
};
typedef PackedPairT<false> PackedPair;


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

// Build the packed operands of a pair. S1/S2: forward packed reads; m = min(len1, len2).
// Returns 0 if read2[:m] contains code 0 ('X': reverse_complement raises KeyError) -> byte path decides.
template <bool IH>
ATR_HD int packed_pair_setup(PackedPairT<IH>& pp, const uint32_t* S1, const uint32_t* S2, int m) {
    const int W = (m + 7) >> 3;
    const int st = pp.stride;
    int ok = 1;
    // zero-nibble scan of read2[:m]
    for (int w = 0; w < W; w++) {
        uint32_t x = S2[w];
        uint32_t valid = (w == W - 1 && (m & 7)) ? ((1u << ((m & 7) * 4)) - 1u) : 0xFFFFFFFFu;
        uint32_t nz = x | (x >> 1); nz |= nz >> 2; nz &= valid & 0x11111111u;
        if (nz != (valid & 0x11111111u)) ok = 0;
    }
    // reversed words: big[w] = brev(S2m[W-1-w]) holds rc of the 8W-nibble padded read; the m real
    // bases start at nibble 8W - m -> shift right by that many nibbles
    const int pad = 8 * W - m;                      // 0..7
    const unsigned bs = (unsigned)pad * 4u;
    for (int w = 0; w < W; w++) {
        uint32_t a = S2[W - 1 - w];
        if (w == 0 && (m & 7)) a &= (1u << ((m & 7) * 4)) - 1u;         // drop bases beyond m in the last word
        uint32_t b = 0;
        if (w + 1 < W) b = S2[W - 2 - w];
        pp.R[w * st] = funnel_r(brev32(a), brev32(b), bs);
    }
    pp.R[W * st] = 0; pp.R[(W + 1) * st] = 0;
    for (int w = 0; w < W; w++) pp.Q[w * st] = S1[w];
    pp.S1 = S1; pp.S2 = S2; pp.m = m;
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
