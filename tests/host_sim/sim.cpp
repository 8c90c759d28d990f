// tests/host_sim/sim.cpp -- TEST-ONLY: runs the per-read device functions of
// atropos_b200/csrc/locate_core.cuh on the CPU (they are __host__ __device__) so that the kernel
// logic can be checked against the oracle in the GPU-less build container. Never part of the product.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../atropos_b200/csrc/adapter_build.hpp"
#include "../../atropos_b200/csrc/locate_core.cuh"
#include "../../atropos_b200/csrc/qgram_core.cuh"
#include "../../atropos_b200/csrc/insert_core.cuh"
#include "../../atropos_b200/csrc/fastq_core.cuh"
#include "../../atropos_b200/csrc/merge_core.cuh"

static int g_sim_qg = 1;          // 0: keep the Shift-And first stage even where the q-gram form is eligible (A/B)

extern "C" {

void sim_set_qg(int on) { g_sim_qg = on; }

// First stage of the funnel both ways for one read: the Shift-And automaton (sa_filter) and the q-gram sampling
// (qg_filter). out[0..5] = cls, dlo, width, c0, c1, v of the automaton, out[6..11] of the q-gram form; also the raw
// hit ranges out[12..13] (automaton hmin, hmax), out[14..15] (q-gram). Returns 1 if the adapter has both forms.
int sim_filter_compare(const atr_adapter_desc* d, const unsigned char* read, int len, int lo, int hi, int fold_case, int* out) {
    AtrTables tb;
    atr::build_tables(tb);
    atr::HostAdapter h;
    std::string msg;
    if (atr::prepare_adapter(*d, tb, h, msg) != ATR_OK || !h.k1a_ok) return 0;
    AdapterK1a a;
    atr::fill_k1a(h, tb, 0, 0, nullptr, a);
    std::vector<unsigned char> tab;
    if (!a.sa_ok || !atr::build_qg(a, tab)) return 0;
    a.qg_tab = tab.data();
    if (hi > len) hi = len;
    if (lo > hi) lo = hi;
    const int n = hi - lo, nwords = (len + 7) / 8;
    std::vector<uint32_t> codes((size_t)nwords + 1, 0);
    int esc = 0;
    for (int w = 0; w < nwords; w++) codes[w] = atr::pack_word(read, len, w, fold_case, tb.iupac, &esc);
    if (esc || n > ATR_K1A_MAXN) return 0;
    unsigned sa_peq[16], tail_peq[16];
    const int mp = a.sa_rows, sh32 = 32 - mp;
    for (int c = 0; c < 16; c++) {
        const unsigned low = (unsigned)(a.peq[c] & (mp >= 32 ? 0xFFFFFFFFull : ((1ull << mp) - 1)));
        sa_peq[c] = low;
        tail_peq[c] = (sh32 ? (low << sh32) | ((1u << sh32) - 1u) : low);
    }
    SaResult r1, r2;
    memset(&r1, 0, sizeof r1); memset(&r2, 0, sizeof r2);
    sa_filter(a, sa_peq, tail_peq, codes.data(), lo, n, r1);
    qg_filter(a, tail_peq, codes.data(), (lo + n + 7) >> 3, lo, n, r2);
    const SaResult* rs[2] = {&r1, &r2};
    for (int t = 0; t < 2; t++) {
        int* o = out + 6 * t;
        o[0] = rs[t]->cls;
        o[1] = rs[t]->cls == 1 || rs[t]->cls == 2 ? rs[t]->dlo : 0; o[2] = rs[t]->cls == 1 || rs[t]->cls == 2 ? rs[t]->width : 0;
        o[3] = rs[t]->cls == 1 || rs[t]->cls == 2 ? rs[t]->c0 : 0; o[4] = rs[t]->cls == 1 || rs[t]->cls == 2 ? rs[t]->c1 : 0;
        o[5] = rs[t]->cls == 3 ? rs[t]->v : 0;
    }
    unsigned st_final;
    unsigned long long sa_pair[256];
    for (int b = 0; b < 256; b++) sa_pair[b] = (unsigned long long)sa_peq[b & 15] | ((unsigned long long)sa_peq[b >> 4] << 32);
    sa_scan(a, sa_peq, sa_pair, codes.data(), lo, n, out[12], out[13], st_final);
    uint32_t acc[ATR_QG_GROUPS];
    if (a.qg_step == 3) qg_scan<3>(a, a.qg_tab, codes.data(), (lo + n + 7) >> 3, lo, n, acc, 1, out[14], out[15]);
    else qg_scan<2>(a, a.qg_tab, codes.data(), (lo + n + 7) >> 3, lo, n, acc, 1, out[14], out[15]);
    out[16] = sa_need_tail(a, n, out[13], st_final);
    out[17] = qg_need_tail(a, sa_peq, codes.data(), lo, n, out[15]);
    out[18] = a.qg_step;
    return 1;
}

// route: 0 = as the kernels would (K1f if eligible, else K1a, else K1g), 1 = force the general kernel (K1g),
// 2 = force plain K1a (error if not possible), 3 = require K1f (error if not eligible). *used_k1a: 0 K1g, 1 K1a, 2 K1f
int sim_locate(const atr_adapter_desc* d, int adapter_index, int reduce, const unsigned char* read, int len,
               int lo, int hi, int fold_case, int route, atr_match* out, int* used_k1a) {
    AtrTables tb;
    atr::build_tables(tb);
    atr::HostAdapter h;
    std::string msg;
    int rc = atr::prepare_adapter(*d, tb, h, msg);
    if (rc != ATR_OK) return rc;
    if (hi > len) hi = len;
    if (lo > hi) lo = hi;
    const int n = hi - lo;
    const int nwords = (len + 7) / 8;
    std::vector<uint32_t> codes((size_t)nwords + 1, 0);
    int esc = 0;
    for (int w = 0; w < nwords; w++) codes[w] = atr::pack_word(read, len, w, fold_case, tb.iupac, &esc);
    bool k1a = h.k1a_ok && n <= ATR_K1A_MAXN && !(esc && (!h.and_mode || h.need_find));
    if (route == 1) k1a = false;
    if ((route == 2 || route == 3) && !k1a) return -100;
    if (used_k1a) *used_k1a = k1a;
    const unsigned char* rmp = h.rmp_ok.empty() ? nullptr : h.rmp_ok.data();
    if (k1a) {
        AdapterK1a a;
        atr::fill_k1a(h, tb, adapter_index, reduce, rmp, a);
        std::vector<unsigned char> qg_tab;
        if (g_sim_qg && atr::build_qg(a, qg_tab)) a.qg_tab = qg_tab.data(); else a.qg_ok = 0;
        if (route == 3 && !a.fused_ok) return -101;
        if (used_k1a && a.fused_ok && route != 2) *used_k1a = 2;
        if (a.fused_ok && route != 2) {                // as the library does: the fused kernels whenever eligible
            int path = 0;
            if (h.and_mode) { if (h.m <= 32) k1f_read<unsigned int, true>(a, codes.data(), lo, n, out, &path); else k1f_read<unsigned long long, true>(a, codes.data(), lo, n, out, &path); }
            else { if (h.m <= 32) k1f_read<unsigned int, false>(a, codes.data(), lo, n, out, &path); else k1f_read<unsigned long long, false>(a, codes.data(), lo, n, out, &path); }
            if (used_k1a) *used_k1a = 2 + path;       // 2 filtered out, 3 banded K1d, 4 windowed K1a
        }
        else if (a.filter_only && !a.anchor_ok && route != 2) {   // funnel shape, dearer indels: filter stage, then the register DP
            int path = 0;
            if (h.and_mode) { if (h.m <= 32) icfilter_read<unsigned int, true>(a, codes.data(), lo, n, out, &path); else icfilter_read<unsigned long long, true>(a, codes.data(), lo, n, out, &path); }
            else { if (h.m <= 32) icfilter_read<unsigned int, false>(a, codes.data(), lo, n, out, &path); else icfilter_read<unsigned long long, false>(a, codes.data(), lo, n, out, &path); }
            if (used_k1a) *used_k1a = 30 + path;       // 30 filtered out, 31 register DP, 35 verbatim shortcut
        }
        else if (a.anchor_ok && route != 2) {          // anchored adapter: piece filter, then the register DP
            if (used_k1a) *used_k1a = 20 + (anchor_filter(a, codes.data(), lo, n) ? 1 : 0);
            if (h.and_mode) anchor_read<true>(a, codes.data(), lo, n, out); else anchor_read<false>(a, codes.data(), lo, n, out);
        }
        else if (h.and_mode) k1a_read<true>(a, codes.data(), lo, n, out);
        else k1a_read<false>(a, codes.data(), lo, n, out);
    } else {
        AdapterGen g;
        atr::fill_gen(h, adapter_index, reduce, h.ref_gen.data(), (const unsigned char*)h.seq.data(), h.thr_mul.data(),
                      h.thr_div.data(), rmp, g);
        std::vector<GCell> col((size_t)h.m + 1);
        gen_read(g, tb, read + lo, n, fold_case, col.data(), 1, out);
    }
    return 0;
}

// InsertAligner.match_insert for one pair through the K2 per-pair functions.
// route: 0 as the kernels would (packed path unless escaped / too long / X in read2), 1 force the byte path
int sim_match_insert(const atr_insert_desc* d, const unsigned char* r1, int len1, const unsigned char* r2, int len2,
                     int route, atr_insert_result* out, int* used_packed) {
    AtrTables tb;
    atr::build_tables(tb);
    atr::HostInsert h;
    std::string msg;
    int rc = atr::prepare_insert(*d, tb, h, msg);
    if (rc) return rc;
    InsertDev v = h.dev;
    v.k_by_len = h.k_by_len.data(); v.thr_ins = h.thr_ins.data(); v.maxmm = h.maxmm.data();
    v.a1_pack = h.a1_pack.data(); v.a2_pack = h.a2_pack.data();
    v.a1_code = h.a1_code.data(); v.a2_code = h.a2_code.data(); v.a1_ascii = h.a1_ascii.data(); v.a2_ascii = h.a2_ascii.data();
    v.insert_prob = h.insert_prob.data(); v.adapter_prob = h.adapter_prob.data(); v.comp = h.comp.data(); v.ov_tab = h.ov_tab.data();
    const int m = len1 < len2 ? len1 : len2;
    if (m > v.max_len) return ATR_E_LIMIT;
    std::vector<uint32_t> c1((size_t)(len1 + 7) / 8 + 2, 0), c2((size_t)(len2 + 7) / 8 + 2, 0);
    int esc = 0;
    for (int w = 0; w < (len1 + 7) / 8; w++) c1[w] = atr::pack_word(r1, len1, w, 0, tb.iupac, &esc);
    for (int w = 0; w < (len2 + 7) / 8; w++) c2[w] = atr::pack_word(r2, len2, w, 0, tb.iupac, &esc);
    std::vector<uint32_t> R2(ATR_K2_MAXW2 + 2, 0);
    std::vector<Cand> cand(ATR_MAX_CAND + 1);
    bool routed = esc || m > ATR_K2_MAXLEN || !v.packed_ok || route == 1;
    {
        PackedPair pp;
        pp.R2 = R2.data(); pp.stride = 1; pp.thr = v.thr_ins;
        if (!routed) routed = packed_pair_setup(pp, c1.data(), c2.data(), m, (len1 + 7) / 8) == 0;
        if (used_packed) *used_packed = !routed;
        if (!routed) { insert_pair(v, pp, true, m, len1, len2, cand.data(), out); return 0; }
    }
    BytePair bp;
    bp.s1 = r1; bp.s2 = r2; bp.comp = v.comp; bp.ov_tab = v.ov_tab; bp.m = m;
    for (int p = 0; p < m; p++) if (v.comp[r2[p]] == 0) {
        im_clear(out->insert); im_clear(out->match1); im_clear(out->match2);
        out->insert.status = ATR_ST_KEYERROR;
        return 0;
    }
    insert_pair(v, bp, false, m, len1, len2, cand.data(), out);
    return 0;
}

int sim_multi_locate(const unsigned char* ref, int m, const unsigned char* query, int n, double rate, int flags,
                     int min_overlap, int max_matches, int* out6) {
    std::vector<unsigned short> thr((size_t)m + 1);
    for (int l = 0; l <= m; l++) thr[l] = atr::thr_mul_of(l, rate);
    std::vector<GCellM> col((size_t)m + 1);
    return gen_multi_locate(ref, m, query, n, (int)(rate * m), thr.data(), flags, min_overlap, max_matches, col.data(), out6);
}

// compare_prefixes as atr_compare_prefixes / k_compare_prefixes evaluate it (_align.pyx:501-544)
int sim_compare_prefixes(const unsigned char* ref, int m, const unsigned char* query, int n, int wildcard_ref,
                         int wildcard_query, int* out6) {
    AtrTables tb;
    atr::build_tables(tb);
    const int length = m < n ? m : n;
    const int mode = (wildcard_ref || wildcard_query) ? 1 : 0;
    const unsigned char* tr = wildcard_ref ? tb.iupac : tb.acgt;
    const unsigned char* tq = wildcard_query ? tb.iupac : tb.acgt;
    int matches = 0;
    for (int i = 0; i < length; i++) matches += mode == 0 ? (ref[i] == query[i]) : ((tr[ref[i]] & tq[query[i]]) != 0);
    out6[0] = 0; out6[1] = length; out6[2] = 0; out6[3] = length; out6[4] = matches; out6[5] = length - matches;
    return 0;
}

// The FASTQ-in -> trimmed-FASTQ-out path with the device functions of fastq_core.cuh (framing, window/statistics
// bookkeeping, formatting) and sim_locate for the alignments: what atr_trim_fastq_host computes, one record at a
// time. counters = {records, with_adapters, bp_in, bp_out, overflow}. Returns 0, or ATR_E_FORMAT with *err filled.
static void sim_count_filter(FqOpsCounters& oc, int f) {
    if (f == 1) oc.too_short++; else if (f == 2) oc.too_long++; else if (f == 3) oc.too_many_n++;
    else if (f == 4) oc.discarded_trimmed++; else if (f == 5) oc.discarded_untrimmed++; else oc.records_written++;
}

int sim_trim_fastq(const atr_adapter_desc* descs, int n_adapters, const atr_trim_opts* o, const unsigned char* text,
                   long long nbytes, unsigned char* out_text, long long* out_bytes, long long* consumed, long long* counters,
                   long long* errors_front, long long* errors_back, long long* adjacent, atr_fastq_error* err,
                   FqOpsCounters* oc, const atr_adapter_desc* linked_back /* NULL, or the back adapter of a linked adapter */) {
    std::vector<uint32_t> nl;
    for (long long i = 0; i < nbytes; i++) {
        if (text[i] == '\n') nl.push_back((uint32_t)i);
        if (text[i] == '\r' && !(i + 1 < nbytes && text[i + 1] == '\n') && (o->final_chunk || i + 1 < nbytes)) {
            err->kind = ATR_FQ_BARE_CR; err->record = -1;
            return ATR_E_FORMAT;
        }
    }
    const long long n_nl = (long long)nl.size();
    const int unterminated = o->final_chunk && nbytes > 0 && text[nbytes - 1] != '\n';
    const long long lines = n_nl + unterminated, n_rec = lines / 4;
    const int lines_left = o->final_chunk ? (int)(lines % 4) : 0;
    *consumed = 0;
    if (n_rec > 0) *consumed = (4 * n_rec - 1 < n_nl) ? (long long)nl[4 * n_rec - 1] + 1 : nbytes;
    if (o->final_chunk) *consumed = nbytes;
    nl.push_back(0);
    std::vector<FqRec> recs((size_t)n_rec + 1);
    long long bad_key = -1;
    for (long long r = 0; r <= n_rec; r++) {
        if (r == n_rec && lines_left == 0) break;
        int bad = 0;
        FqRec R;
        const int kind = fq_frame(text, nl.data(), n_nl, nbytes, r, r < n_rec ? 4 : lines_left, R, bad);
        if (r < n_rec) recs[(size_t)r] = R;
        if (kind != ATR_FQ_OK) { bad_key = ((4 * r + bad) << 8) | kind; break; }       // records are visited in file order
    }
    if (bad_key >= 0) {
        const long long line = bad_key >> 8;
        err->kind = (int)(bad_key & 0xFF); err->record = line / 4; err->line_in_record = (int)(line % 4);
        if (line >= lines) { err->line_begin = err->line_end = nbytes; err->terminated = 0; }
        else {
            err->line_begin = line > 0 ? (long long)nl[line - 1] + 1 : 0;
            err->line_end = line < n_nl ? (long long)nl[line] : nbytes;
            err->terminated = line < n_nl;
        }
        return ATR_E_FORMAT;
    }
    const size_t H = (size_t)(o->max_len + 1) * (size_t)(o->max_errors + 1);
    long long opos = 0;
    for (long long r = 0; r < n_rec; r++) {
        FqRec R = recs[(size_t)r];
        counters[0]++; counters[2] += R.seq_len;
        unsigned bpc, bpq, bpg;
        fq_pre_ops(o->ops, 0, text, R, bpc, bpq, bpg);
        oc->bp_cut[0] += bpc; oc->bp_quality[0] += bpq; oc->bp_nextseq[0] += bpg;
        int lo = 0, hi = R.seq_len;
        bool any = false;
        for (int round = 0; round < o->times; round++) {
            atr_match m;
            m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0; m.adapter = -1; m.status = ATR_ST_NONE;
            for (int a = 0; a < n_adapters; a++) {
                int used = 0;
                int rc = sim_locate(&descs[a], a, a > 0, text + R.seq_b, R.seq_len, lo, hi, 1, 0, &m, &used);
                if (rc) return rc;
            }
            if (m.status == ATR_ST_INVALID) { err->kind = ATR_FQ_INVALID_MATCH; err->record = -1; return ATR_E_FORMAT; }
            FqApply ap;
            const int w = m.adapter >= 0 ? descs[m.adapter].flags : 0;
            const int ff = (w == ATR_SEMIGLOBAL) ? -1 : ((w == 14 || w == 2) ? 0 : 1);
            if (!fq_apply(m, ff, lo, hi, text + R.seq_b, ap)) break;
            any = true;
            if (ap.length <= o->max_len && ap.errors <= o->max_errors)
                (ap.front ? errors_front : errors_back)[((size_t)m.adapter * (size_t)(o->max_len + 1) + (size_t)ap.length) * (size_t)(o->max_errors + 1) + (size_t)ap.errors]++;
            else counters[4]++;
            if (!ap.front) adjacent[(size_t)m.adapter * 5 + (size_t)ap.adjacent]++;
            lo = ap.new_lo; hi = ap.new_hi;
            if (linked_back) {                          // LinkedAdapter.match_to: the back adapter in what the front adapter left
                atr_match mb;
                im_clear(mb);
                int used = 0;
                int rc = sim_locate(linked_back, 0, 0, text + R.seq_b, R.seq_len, lo, hi, 1, 0, &mb, &used);
                if (rc) return rc;
                if (mb.status == ATR_ST_INVALID) { err->kind = ATR_FQ_INVALID_MATCH; err->record = -1; return ATR_E_FORMAT; }
                const int wb = linked_back->flags;
                FqApply ab;
                if (fq_apply(mb, (wb == 14 || wb == 2) ? 0 : 1, lo, hi, text + R.seq_b, ab)) {
                    if (ab.length <= o->max_len && ab.errors <= o->max_errors)
                        (ab.front ? errors_front : errors_back)[((size_t)1 * (size_t)(o->max_len + 1) + (size_t)ab.length) * (size_t)(o->max_errors + 1) + (size_t)ab.errors]++;
                    else counters[4]++;
                    if (!ab.front) adjacent[(size_t)1 * 5 + (size_t)ab.adjacent]++;
                    lo = ab.new_lo; hi = ab.new_hi;
                }
            }
        }
        (void)H;
        if (any) counters[1]++;
        if (o->ops.trim_n) { unsigned bpn; fq_trim_n(text + R.seq_b, lo, hi, bpn); oc->bp_n_ends[0] += bpn; }
        const int flt = fq_filter(o->ops, text + R.seq_b, lo, hi, any, text, 0, 0, false, false);
        sim_count_filter(*oc, flt);
        if (flt) continue;
        counters[3] += hi - lo;
        const uint32_t total = fq_out_len(R, lo, hi);
        for (uint32_t i = 0; i < total; i++) out_text[opos + i] = fq_out_byte(text, R, lo, hi, i);
        opos += total;
    }
    *out_bytes = opos;
    return 0;
}

// Paired-end twin (atr_trim_fastq_pe_host): PairedSequenceReader + InsertAdapterCutter + two formatters, one pair at a
// time, with the device functions of fastq_core.cuh, sim_match_insert and sim_locate.
// counters = {records, insert_matches, with1, with2, bp_in1, bp_in2, bp_out1, bp_out2, overflow}
namespace {
struct SimText {
    const unsigned char* text; long long nbytes; std::vector<uint32_t> nl; long long n_nl, lines, n_rec; int lines_left;
    bool index(int final_chunk) {
        for (long long i = 0; i < nbytes; i++) {
            if (text[i] == '\n') nl.push_back((uint32_t)i);
            if (text[i] == '\r' && !(i + 1 < nbytes && text[i + 1] == '\n') && (final_chunk || i + 1 < nbytes)) return false;
        }
        n_nl = (long long)nl.size();
        const int unterminated = final_chunk && nbytes > 0 && text[nbytes - 1] != '\n';
        lines = n_nl + unterminated; n_rec = lines / 4; lines_left = final_chunk ? (int)(lines % 4) : 0;
        nl.push_back(0);
        return true;
    }
    long long consumed_for(long long n) const { return n == 0 ? 0 : (4 * n - 1 < n_nl ? (long long)nl[4 * n - 1] + 1 : nbytes); }
    void describe(long long line, atr_fastq_error* err) const {
        if (line >= lines) { err->line_begin = err->line_end = nbytes; err->terminated = 0; return; }
        err->line_begin = line > 0 ? (long long)nl[line - 1] + 1 : 0;
        err->line_end = line < n_nl ? (long long)nl[line] : nbytes;
        err->terminated = line < n_nl;
    }
};
}

// idesc == NULL: adapter mode, n_ad1 / n_ad2 adapters per read (d1 / d2 arrays), `times` rounds each; ef1 / ef2 = front histograms
int sim_trim_fastq_pe(const atr_insert_desc* idesc, const atr_adapter_desc* d1, int n_ad1, const atr_adapter_desc* d2, int n_ad2,
                      long long* ef1, long long* ef2,
                      const atr_trim_pe_opts* o, const unsigned char* text1, long long nbytes1, const unsigned char* text2,
                      long long nbytes2, unsigned char* out1, unsigned char* out2, long long* out_bytes, long long* consumed,
                      long long* counters, long long* eb1, long long* eb2, long long* adj1, long long* adj2, atr_fastq_error* err,
                      FqOpsCounters* oc, long long* corrected /* {records, bp1, bp2} */,
                      const atr_merge_opts* mopts /* NULL: no MergeOverlapping */, unsigned char* outm /* NULL: merged reads are discarded */,
                      long long* mcounters /* {merged, written, bp, records corrected, bp1, bp2} */) {
    // error correction edits the reads in place: work on private copies of the texts, like the GPU path on its chunk
    std::vector<unsigned char> copy1(text1, text1 + nbytes1), copy2(text2, text2 + nbytes2);
    copy1.push_back(0); copy2.push_back(0);
    unsigned char* const mtext1 = copy1.data();
    unsigned char* const mtext2 = copy2.data();
    text1 = mtext1; text2 = mtext2;
    SimText T[2];
    T[0].text = text1; T[0].nbytes = nbytes1; T[1].text = text2; T[1].nbytes = nbytes2;
    for (int f = 0; f < 2; f++) if (!T[f].index(o->final_chunk)) { err->kind = ATR_FQ_BARE_CR; err->file = f; err->record = -1; return ATR_E_FORMAT; }
    const long long n = T[0].n_rec < T[1].n_rec ? T[0].n_rec : T[1].n_rec;
    auto fail_line = [&](int f, long long r, int bad, int kind) {
        err->kind = kind; err->file = f; err->record = r; err->line_in_record = bad;
        T[f].describe(4 * r + bad, err);
        return ATR_E_FORMAT;
    };
    std::vector<FqRec> R1((size_t)n + 1), R2((size_t)n + 1);
    for (long long r = 0; r < n; r++) {                 // pairs in file order: read 1, read 2, names
        int bad = 0;
        int kind = fq_frame(text1, T[0].nl.data(), T[0].n_nl, nbytes1, r, 4, R1[(size_t)r], bad);
        if (kind) return fail_line(0, r, bad, kind);
        kind = fq_frame(text2, T[1].nl.data(), T[1].n_nl, nbytes2, r, 4, R2[(size_t)r], bad);
        if (kind) return fail_line(1, r, bad, kind);
        const int nm = fq_names_match(text1, R1[(size_t)r], text2, R2[(size_t)r]);
        if (nm) { err->kind = nm == 1 ? ATR_FQ_PAIR_NAMES : ATR_FQ_EMPTY_NAME; err->file = 0; err->record = r;
                  err->line_begin = R1[(size_t)r].hdr_b; err->line_end = R1[(size_t)r].hdr_b + R1[(size_t)r].hdr_len;
                  err->line_begin2 = R2[(size_t)r].hdr_b; err->line_end2 = R2[(size_t)r].hdr_b + R2[(size_t)r].hdr_len;
                  err->terminated = 1; return ATR_E_FORMAT; }
    }
    if (o->final_chunk) {                               // the end of the files: PairedSequenceReader.__iter__ :431-447
        FqRec tmp; int bad = 0;
        const bool more1 = T[0].n_rec > n, more2 = T[1].n_rec > n;
        if (more1) {                                    // next(it1) yields record n (validated), next(it2) ends (or fails)
            int kind = fq_frame(text1, T[0].nl.data(), T[0].n_nl, nbytes1, n, 4, tmp, bad);
            if (kind) return fail_line(0, n, bad, kind);
            if (T[1].lines_left) { kind = fq_frame(text2, T[1].nl.data(), T[1].n_nl, nbytes2, n, T[1].lines_left, tmp, bad); return fail_line(1, n, bad, kind); }
            err->kind = ATR_FQ_MORE_IN_1; err->file = 0; err->record = n; return ATR_E_FORMAT;
        }
        if (T[0].lines_left) { int kind = fq_frame(text1, T[0].nl.data(), T[0].n_nl, nbytes1, n, T[0].lines_left, tmp, bad); return fail_line(0, n, bad, kind); }
        if (more2) {
            int kind = fq_frame(text2, T[1].nl.data(), T[1].n_nl, nbytes2, n, 4, tmp, bad);
            if (kind) return fail_line(1, n, bad, kind);
            err->kind = ATR_FQ_MORE_IN_2; err->file = 1; err->record = n; return ATR_E_FORMAT;
        }
        if (T[1].lines_left) { int kind = fq_frame(text2, T[1].nl.data(), T[1].n_nl, nbytes2, n, T[1].lines_left, tmp, bad); return fail_line(1, n, bad, kind); }
    }
    consumed[0] = o->final_chunk ? nbytes1 : T[0].consumed_for(n);
    consumed[1] = o->final_chunk ? nbytes2 : T[1].consumed_for(n);
    long long opos1 = 0, opos2 = 0, oposm = 0;
    std::vector<unsigned short> mh;
    unsigned char mcomp[256];
    MergeTables mtb;
    if (mopts) {
        atr::build_merge_tables(4000, mopts->min_overlap, mopts->error_rate, mh, mcomp);
        mtb.thr_mul = mh.data(); mtb.minov = mh.data() + 4001; mtb.comp = mcomp; mtb.max_len = 4000;
    }
    for (long long r = 0; r < n; r++) {
        FqRec A = R1[(size_t)r], B = R2[(size_t)r];
        counters[0]++; counters[4] += A.seq_len; counters[5] += B.seq_len;
        int pflags = 0;
        unsigned bpc, bpq, bpg;
        fq_pre_ops(o->ops, 0, text1, A, bpc, bpq, bpg);
        oc->bp_cut[0] += bpc; oc->bp_quality[0] += bpq; oc->bp_nextseq[0] += bpg;
        fq_pre_ops(o->ops, 1, text2, B, bpc, bpq, bpg);
        oc->bp_cut[1] += bpc; oc->bp_quality[1] += bpq; oc->bp_nextseq[1] += bpg;
        int len1 = A.seq_len;
        const int len2 = B.seq_len;
        int k1 = len1, k2 = len2;
        PeMatch m1, m2;
        m1.present = m2.present = 0;
        if (idesc) {
            atr_insert_result ins;
            im_clear(ins.insert); im_clear(ins.match1); im_clear(ins.match2);
            atr_match fb1, fb2;
            im_clear(fb1); im_clear(fb2);
            if (len1 >= o->min_insert_overlap && len2 >= o->min_insert_overlap) {
                int used = 0;
                int rc = sim_match_insert(idesc, text1 + A.seq_b, len1, text2 + B.seq_b, len2, 0, &ins, &used);
                if (rc) return rc;
                if (ins.insert.status == ATR_ST_NONE) {
                    rc = sim_locate(d1, 0, 0, text1 + A.seq_b, len1, 0, len1, 1, 0, &fb1, &used);
                    if (!rc) rc = sim_locate(d2, 0, 0, text2 + B.seq_b, len2, 0, len2, 1, 0, &fb2, &used);
                    if (rc) return rc;
                }
            }
            int hit = 0, invalid = 0, im[4];
            bool correct = false;
            fq_pe_decide(ins, fb1, fb2, len1, len2, o->min_insert_overlap, o->symmetric, o->mismatch_action, m1, m2, hit, invalid, correct, im);
            if (invalid) { err->kind = ATR_FQ_INVALID_MATCH; err->record = r; return ATR_E_FORMAT; }
            counters[1] += hit;
            if (hit) pflags |= FQ_PF_INSERT;
            if (correct) {
                atr::HostInsert hi;
                std::string msg;
                AtrTables tb;
                atr::build_tables(tb);
                if (atr::prepare_insert(*idesc, tb, hi, msg)) return -200;
                int c1 = 0, c2 = 0, nl1 = len1;
                if (!fq_pe_correct(mtext1 + A.seq_b, mtext1 + A.qual_b, len1, mtext2 + B.seq_b, mtext2 + B.qual_b, len2, im[0], im[1], im[2],
                                   im[3], o->mismatch_action, hi.comp.data(), c1, c2, nl1)) { err->kind = ATR_FQ_CORRECTION; err->record = r; return ATR_E_FORMAT; }
                if (c1 || c2) corrected[0]++;
                corrected[1] += c1; corrected[2] += c2;
                if (c1) pflags |= FQ_PF_CORRECTED1;
                if (c2) pflags |= FQ_PF_CORRECTED2;
                len1 = nl1;
            }
            FqApply ap;
            bool counted;
            k1 = fq_pe_trim(m1, len1, text1 + A.seq_b, ap, counted);
            if (m1.present) counters[2]++;
            if (counted) {
                if (ap.length <= o->max_len && ap.errors <= o->max_errors) eb1[(size_t)ap.length * (size_t)(o->max_errors + 1) + (size_t)ap.errors]++; else counters[8]++;
                adj1[ap.adjacent]++;
            }
            k2 = fq_pe_trim(m2, len2, text2 + B.seq_b, ap, counted);
            if (m2.present) counters[3]++;
            if (counted) {
                if (ap.length <= o->max_len && ap.errors <= o->max_errors) eb2[(size_t)ap.length * (size_t)(o->max_errors + 1) + (size_t)ap.errors]++; else counters[8]++;
                adj2[ap.adjacent]++;
            }
        } else {                                         // two independent AdapterCutters (the single-end loop per read)
            for (int f = 0; f < 2; f++) {
                const atr_adapter_desc* descs = f ? d2 : d1;
                const int n_ad = f ? n_ad2 : n_ad1;
                const unsigned char* text = f ? text2 : text1;
                const FqRec& R = f ? B : A;
                long long *efr = f ? ef2 : ef1, *eb = f ? eb2 : eb1, *adj = f ? adj2 : adj1;
                int lo = 0, hi = R.seq_len;
                bool any = false;
                for (int round = 0; round < o->times && n_ad > 0; round++) {
                    atr_match m;
                    im_clear(m);
                    for (int a = 0; a < n_ad; a++) {
                        int used = 0;
                        int rc = sim_locate(&descs[a], a, a > 0, text + R.seq_b, R.seq_len, lo, hi, 1, 0, &m, &used);
                        if (rc) return rc;
                    }
                    if (m.status == ATR_ST_INVALID) { err->kind = ATR_FQ_INVALID_MATCH; err->record = r; return ATR_E_FORMAT; }
                    FqApply ap;
                    const int w = m.adapter >= 0 ? descs[m.adapter].flags : 0;
                    const int ff = (w == ATR_SEMIGLOBAL) ? -1 : ((w == 14 || w == 2) ? 0 : 1);
                    if (!fq_apply(m, ff, lo, hi, text + R.seq_b, ap)) break;
                    any = true;
                    if (ap.length <= o->max_len && ap.errors <= o->max_errors)
                        (ap.front ? efr : eb)[((size_t)m.adapter * (size_t)(o->max_len + 1) + (size_t)ap.length) * (size_t)(o->max_errors + 1) + (size_t)ap.errors]++;
                    else counters[8]++;
                    if (!ap.front) adj[(size_t)m.adapter * 5 + (size_t)ap.adjacent]++;
                    lo = ap.new_lo; hi = ap.new_hi;
                }
                if (any) counters[2 + f]++;
                (f ? m2 : m1).present = any;
                // windows may start beyond 0 (front adapters): carried into the post stage below
                if (f == 0) { k1 = hi; A.seq_b += (uint32_t)lo; A.qual_b += (uint32_t)lo; k1 -= lo; }
                else { k2 = hi; B.seq_b += (uint32_t)lo; B.qual_b += (uint32_t)lo; k2 -= lo; }
            }
        }
        int lo1 = 0, hi1 = k1, lo2 = 0, hi2 = k2;
        if (o->ops.trim_n) {
            unsigned bpn;
            fq_trim_n(text1 + A.seq_b, lo1, hi1, bpn); oc->bp_n_ends[0] += bpn;
            fq_trim_n(text2 + B.seq_b, lo2, hi2, bpn); oc->bp_n_ends[1] += bpn;
        }
        if (mopts) {                                     // MergeOverlapping, then MergedReadFilter in front of the other filters
            if (m1.present) pflags |= FQ_PF_MATCH1;
            if (m2.present) pflags |= FQ_PF_MATCH2;
            const int l1 = hi1 - lo1, l2 = hi2 - lo2;
            if (l1 > 4000 || l2 > 4000) return ATR_E_LIMIT;
            atr_merge_result mres;
            std::vector<GCell> col((size_t)l2 + 1);
            std::vector<unsigned char> before(text2 + B.seq_b + lo2, text2 + B.seq_b + hi2);      // read 2 before any correction
            before.push_back(0);
            merge_pair(text1 + A.seq_b + lo1, l1, text2 + B.seq_b + lo2, l2, (pflags & FQ_PF_INSERT) ? 1 : 0, mtb, col.data(), 1, &mres);
            FqMergeRec M;
            int c1 = 0, c2 = 0;
            const int d = fq_merge_decide(mres, pflags, o->mismatch_action, mtext1, A, lo1, hi1, mtext2, B, lo2, hi2, mcomp, M, c1, c2);
            if (d == -1) { err->kind = ATR_FQ_INVALID_MATCH; err->record = r; return ATR_E_FORMAT; }
            if (d == -2) { err->kind = ATR_FQ_CORRECTION; err->record = r; return ATR_E_FORMAT; }
            if (c1 || c2) mcounters[3]++;
            mcounters[4] += c1; mcounters[5] += c2;
            if (d == 1) {
                mcounters[0]++;
                if (outm) {
                    const uint32_t total = fq_merged_out_len(A.hdr_len, A.name2, M.mlen);
                    for (uint32_t i = 0; i < total; i++)
                        outm[oposm + i] = fq_merged_out_byte(text1 + A.hdr_b, A.hdr_len, A.name2, text1 + A.seq_b + lo1, text1 + A.qual_b + lo1,
                                                             before.data(), text2 + B.qual_b + lo2, mcomp, M, i);
                    oposm += total;
                    mcounters[1]++; mcounters[2] += M.mlen;
                }
                continue;
            }
        }
        const int flt = fq_filter(o->ops, text1 + A.seq_b, lo1, hi1, m1.present != 0, text2 + B.seq_b, lo2, hi2, m2.present != 0, true);
        sim_count_filter(*oc, flt);
        if (flt) continue;
        counters[6] += hi1 - lo1; counters[7] += hi2 - lo2;
        uint32_t total = fq_out_len(A, lo1, hi1);
        for (uint32_t i = 0; i < total; i++) out1[opos1 + i] = fq_out_byte(text1, A, lo1, hi1, i);
        opos1 += total;
        total = fq_out_len(B, lo2, hi2);
        for (uint32_t i = 0; i < total; i++) out2[opos2 + i] = fq_out_byte(text2, B, lo2, hi2, i);
        opos2 += total;
    }
    out_bytes[0] = opos1; out_bytes[1] = opos2;
    if (mopts) out_bytes[2] = oposm;
    return 0;
}

// MergeOverlapping.__call__ up to the decision (merge_core.cuh: merge_pair), one pair
int sim_merge_overlap(const unsigned char* r1, int len1, const unsigned char* r2, int len2, int insert_matched,
                      double min_overlap, double error_rate, atr_merge_result* out) {
    const int max_len = std::max(len1, len2);
    std::vector<unsigned short> h;
    unsigned char comp[256];
    atr::build_merge_tables(max_len, min_overlap, error_rate, h, comp);
    MergeTables tb;
    tb.thr_mul = h.data(); tb.minov = h.data() + (max_len + 1); tb.comp = comp; tb.max_len = max_len;
    std::vector<GCell> col((size_t)len2 + 1);
    merge_pair(r1, len1, r2, len2, insert_matched, tb, col.data(), 1, out);
    return 0;
}

}  // extern "C"
