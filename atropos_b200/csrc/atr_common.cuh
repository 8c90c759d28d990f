// atr_common.cuh -- device-side structures shared by the kernels and the C-ABI host code.
#pragma once
#include <stdint.h>
#include "../../include/atropos_b200.h"

#if defined(__CUDACC__)
#define ATR_HD __host__ __device__ __forceinline__
#define ATR_D __device__ __forceinline__
#define ATR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ATR_HD inline
#define ATR_D inline
#define ATR_HD_NOINLINE inline
#endif

#define ATR_ESC_BIT 0x8000u          // bit 15 of len[]: read must take the byte-exact general kernel
#define ATR_LEN_MASK 0x7FFFu

// ---- packed-key cell of the register kernel (K1a) ----------------------------------------
// bits [24,32) cost (clamped to k+1)   [22,24) tie-break priority (0 while stored)
//      [7,19)  origin + ATR_ORG_BIAS    [0,7)   matches
// A single unsigned min over (diag+SUB, up+INS, left+DEL) then reproduces the reference's
// "mismatch <= insertion <= deletion" preference (_align.pyx:405-419): equal costs are ordered
// by the priority bits, which are cleared again before the key is stored.
#define ATR_K1A_MAXM 64
#define ATR_K1A_MAXN 4000
#define ATR_COST_SHIFT 24
#define ATR_PRIO_SHIFT 22
#define ATR_ORG_SHIFT 7
#define ATR_ORG_BIAS 64
#define ATR_ORG_MASK 0xFFFu
#define ATR_MAT_MASK 0x7Fu
#define ATR_PRIO_CLEAR (~(3u << ATR_PRIO_SHIFT))

// general kernel (K1g) limits
#define ATR_MAX_ADAPTER 4095
#define ATR_MAX_READ 32767
#define ATR_G_ORG_BIAS 4096

struct AdapterK1a {                   // passed by value as a __grid_constant__ kernel parameter
    int m, k, flags, ic, min_overlap;
    int and_mode;                     // 0: codes compared for equality (ASCII mode); 1: (a & q) != 0
    int q_single_only;                // and_mode with the query under the ACGT table: multi-bit codes -> 0
    int match_to, exact_bypass, cmp_only;   // cmp_only: 0 DP, 1 compare_prefixes, 2 compare_suffixes
    int adapter_index;
    int reduce;                       // 0: overwrite out[i]; 1: keep the previous result unless strictly more matches
    int mark_routed;                  // no general pass follows: flag routed reads ATR_ST_ESCAPED instead of leaving them
    int need_find;                    // match_to with read wildcards only: run the literal str.find shortcut explicitly
    int code[ATR_K1A_MAXM];           // per-row compare operand (4-bit code)
    int lit[ATR_K1A_MAXM];            // literal 4-bit code of the adapter letters (need_find only)
    unsigned short thr_mul[ATR_K1A_MAXM + 1];   // floor(length * rate) in double   (_align.pyx:447, :468)
    unsigned short thr_div[ATR_K1A_MAXM + 1];   // max e with e / size <= rate       (adapters/__init__.py:389-392)
    const unsigned char* rmp_ok;      // [(m+1)*(m+1)] or nullptr
    int fused_ok;                     // eligible for the filter -> banded/windowed DP funnel (atr_kernels.cuh)
    int band_ok, nomatch;             // K1d usable; a 4-bit code that matches no adapter row (virtual columns)
    unsigned long long peq[16];       // bit i-1 of peq[c] set iff adapter row i matches read code c (Shift-And / Myers)
    // Shift-And pre-filter (k_filter_sa): the first sa_rows (<= 32) adapter rows cut into k+1 pieces; an alignment
    // with <= k errors must contain one piece verbatim (pigeonhole)
    int sa_ok, sa_rows;
    int sa_front;                     // unanchored 5' adapters (START_WITHIN_SEQ1 | START_WITHIN_SEQ2 | STOP_WITHIN_SEQ2, m <= 32): the
                                      // pieces find the full-length occurrences, an exact Myers pass over the first m + k columns the
                                      // partial ones at the read start (front_filter, k_filter_front); sa_ok stays 0
    int tail_gate_ok;                 // the Shift-And state can tell when no partial match at the read end is possible
    unsigned tail_mask;               // bit i-1: a candidate (i, n) without a verbatim complete piece leaves this bit set
    unsigned apack[ATR_K1A_MAXM / 8]; // the adapter's compare codes packed like a read (exact-occurrence shortcut)
    int exact_ok;                     // shortcut usable (codes fit nibbles; not the ACGT-filtered query mode)
    short thrJ[ATR_K1A_MAXM + 1];     // start_in_ref adapters: bound on D[m][j] for j = 0..m (j > m uses thrJ[m]); -1 = never
    unsigned sa_start, sa_end;        // bit r-1: row r is the first / last row of a piece
    int split8;                       // fused path: survivors whose band is <= 8 diagonals go to the back of the narrow list (k_band<8>)
    int filter_only;                  // funnel shape but indel cost != 1: the funnel's first stage as a pure filter, then k1a_read
    int anchor_ok;                    // PREFIX / SUFFIX flag set outside the funnel: fixed-position piece filter (k_filter_anchor)
    // q-gram sampling pre-filter (k_filter_qg; qgram_core.cuh): the same pieces as the Shift-And stage, but instead of
    // running an automaton over every column only every qg_step-th read position is looked at: a verbatim piece of
    // length L >= 6 + qg_step - 1 contains a 6-mer that starts at a sampled position. One hashed byte-table lookup
    // per sample; hits are verified by comparing the whole piece, so the hit set is exactly the automaton's.
    int qg_ok, qg_step;               // step 2 or 3 (ASCII compare mode only)
    int qg_wide;                      // pieces over up to 64 rows (adapters whose k+1 pieces do not fit 32 rows): sa_ok = 0, only the
                                      // q-gram form exists; sa_rows > 32, the 64-bit fields below, 64-bit tail Myers
    unsigned long long sa_start64, sa_end64, tail_mask64;   // = sa_start / sa_end / tail_mask where sa_rows <= 32
    unsigned qg_mul;                  // key = (x * qg_mul) >> (32 - ATR_QG_BITS); low 8 bits zero: only 24 bits of x count
    const unsigned char* qg_tab;      // [1 << ATR_QG_BITS]: 0 none, 1..14 pattern index, 15 several patterns share the bucket
    int qg_npat;
    unsigned char qg_prow[16];        // per pattern: 0-based first row of its piece ...
    unsigned char qg_plen[16];        // ... the piece's length ...
    unsigned char qg_poff[16];        // ... and the offset of the 6-mer inside the piece
    unsigned qg_pw[16][2], qg_pm[16][2];  // ... the piece itself packed like a read (up to 16 rows) and its nibble mask
    // need-tail gate without the automaton's final state: tail_mask bit i-1 (row i inside a begun piece p, l = i - first
    // row of p) <=> the read's last l bases equal the first l rows of p: (last8 ^ tail_c[t]) & tail_m[t] == 0
    int n_tail_cmp;
    unsigned tail_c[24], tail_m[24];
    int tail_cols;                    // the same test as an automaton run over the read's last tail_cols columns (<= 16): rows of tail_mask never
                                      // lie deeper than that inside their piece
};
#define ATR_QG_BITS 13
#define ATR_QG_Q 6

struct AdapterGen {                   // general kernel: tables live in global memory
    int m, k, flags, ic, min_overlap;
    int and_mode, q_table;            // q_table: 0 none (ASCII), 1 IUPAC, 2 ACGT translation of the query byte
    int match_to, exact_bypass, cmp_only;
    int adapter_index, reduce;
    int need_find;
    double rate;
    const unsigned char* ref;         // m bytes: ASCII (ascii mode) or translated codes
    const unsigned char* lit;         // m bytes: the adapter's ASCII letters (need_find only)
    const unsigned short* thr_mul;    // m+1
    const unsigned short* thr_div;    // m+1
    const unsigned char* rmp_ok;
};

// translation tables (_align.pyx:31-83), built once on the host and uploaded
struct AtrTables {
    unsigned char iupac[256];
    unsigned char acgt[256];
};

ATR_HD unsigned k1a_key(int cost, int origin, int matches) {
    return ((unsigned)cost << ATR_COST_SHIFT) | ((unsigned)(origin + ATR_ORG_BIAS) << ATR_ORG_SHIFT) | (unsigned)matches;
}
ATR_HD int k1a_cost(unsigned key) { return (int)(key >> ATR_COST_SHIFT); }
ATR_HD int k1a_origin(unsigned key) { return (int)((key >> ATR_ORG_SHIFT) & ATR_ORG_MASK) - ATR_ORG_BIAS; }
ATR_HD int k1a_matches(unsigned key) { return (int)(key & ATR_MAT_MASK); }
