// fastq_core.cuh -- per-record logic of the FASTQ-in -> trimmed-FASTQ-out path ("next" rows f-1/f-2/f-3 of
// SURVEY.md section 8), __host__ __device__ so that tests/host_sim runs the very same functions on the CPU.
//
// What it restates (paths relative to the reference checkout):
//   fq_frame    FastqReader.__iter__            atropos/io/_seqio.pyx:180-245  (+ Sequence.__init__ :33-44)
//   fq_apply    AdapterCutter.__call__ loop     atropos/commands/trim/modifiers.py:141-149
//               Adapter._trimmed_front / _back  atropos/adapters/__init__.py:413-436
//               Match._guess_is_front           atropos/align/__init__.py:108-114
//   fq_out_len / fq_out_byte   FastqFormat.format_entry   atropos/io/seqio.py:693-700
//
// Text model: the reference opens the file in text mode (io/__init__.py:148-149), i.e. with universal newlines:
// "\r\n" reaches the reader as "\n". A bare '\r' (not followed by '\n') would also split a line there; that is
// not reproduced -- such input is refused (ATR_FQ_BARE_CR) instead of being framed differently.
#pragma once
#include "atr_common.cuh"

struct FqLine {
    uint32_t b, e;       // content [b, e): without the terminator ("\n" or "\r\n")
    int term;            // 1: a '\n' ended the line; 0: the text ended first (only the last line of a file)
};

// 20 bytes per record: where its pieces are in the chunk's text
struct FqRec {
    uint32_t hdr_b;      // '@'
    uint32_t seq_b;
    uint32_t qual_b;
    uint16_t hdr_len;    // "@name" without terminator
    uint16_t seq_len;    // == quality length for every accepted record
    uint16_t name2;      // 1: the '+' line repeats the name
    uint16_t pad;
};

// line li of the chunk; nl[] = ascending positions of '\n', n_nl of them. Line n_nl (if any) is unterminated.
ATR_HD FqLine fq_line(const unsigned char* __restrict__ text, const uint32_t* __restrict__ nl, int64_t n_nl, int64_t nbytes,
                      int64_t li) {
    FqLine L;
    L.b = li == 0 ? 0u : nl[li - 1] + 1u;
    if (li < n_nl) {
        L.e = nl[li];
        L.term = 1;
        if (L.e > L.b && text[L.e - 1] == '\r') L.e--;
    } else {
        L.e = (uint32_t)nbytes;
        L.term = 0;
    }
    return L;
}

// The reader slices every line with `line[:strip]`, strip = -1 (the "\n"): an unterminated line loses its last
// character that way (_seqio.pyx:196-197, :207-209, :213).
ATR_HD uint32_t fq_sliced_end(const FqLine& L) { return L.term ? L.e : (L.e > L.b ? L.e - 1u : L.b); }

// Frame and validate record r (lines 4r .. 4r+3). lines_avail < 4 only for the partial record at the end of the
// input: its lines are still checked in order before the reader reports "ended prematurely" (:244-245).
// Returns ATR_FQ_OK or the ATR_FQ_* kind; *bad_line = index (0..3) of the offending line.
ATR_HD int fq_frame(const unsigned char* __restrict__ text, const uint32_t* __restrict__ nl, int64_t n_nl, int64_t nbytes,
                    int64_t r, int lines_avail, FqRec& R, int& bad_line) {
    R.hdr_b = R.seq_b = R.qual_b = 0; R.hdr_len = R.seq_len = R.name2 = R.pad = 0;
    bad_line = 0;
    // line 0: `if not (line and line[0] == '@')` (:202-206)
    const FqLine h = fq_line(text, nl, n_nl, nbytes, 4 * r);
    if (!(h.e > h.b && text[h.b] == '@')) return ATR_FQ_NO_AT;
    const uint32_t name_b = h.b + 1u, name_e = fq_sliced_end(h);
    if (lines_avail < 2) { bad_line = 1; return ATR_FQ_TRUNCATED; }
    const FqLine s = fq_line(text, nl, n_nl, nbytes, 4 * r + 1);
    const uint32_t seq_e = fq_sliced_end(s);
    if (lines_avail < 3) { bad_line = 2; return ATR_FQ_TRUNCATED; }
    // line 2: '+\n' is the common case; else the sliced line must start with '+', and if it goes on it must
    // repeat the name (:210-230)
    const FqLine p = fq_line(text, nl, n_nl, nbytes, 4 * r + 2);
    const uint32_t plus_e = fq_sliced_end(p);
    bad_line = 2;
    if (!(plus_e > p.b && text[p.b] == '+')) return ATR_FQ_NO_PLUS;
    int name2 = 0;
    if (plus_e - p.b > 1u) {
        const uint32_t l2 = plus_e - (p.b + 1u);
        bool same = l2 == (name_e >= name_b ? name_e - name_b : 0u);
        for (uint32_t t = 0; same && t < l2; t++) same = text[p.b + 1u + t] == text[name_b + t];
        if (!same) return ATR_FQ_NAME_MISMATCH;
        name2 = 1;
    }
    if (lines_avail < 4) { bad_line = 3; return ATR_FQ_TRUNCATED; }
    // line 3: `if len(line) == len(sequence) - strip: qualities = line[:strip] else: line.rstrip('\r\n')`
    // (:231-235): a terminated line is its content; an unterminated one that is exactly one character too long
    // loses that character (and is accepted), anything else is kept whole. Then Sequence.__init__ demands equal
    // lengths (:33-44), reported as "Error creating sequence record" (:236-241).
    const FqLine q = fq_line(text, nl, n_nl, nbytes, 4 * r + 3);
    const uint32_t slen = seq_e - s.b;
    uint32_t qlen = q.e - q.b;
    if (!q.term && qlen == slen + 1u) qlen = slen;
    bad_line = 3;
    if (qlen != slen) return ATR_FQ_LENGTH;
    if (slen > (uint32_t)ATR_MAX_READ || h.e - h.b > 65535u) return ATR_FQ_TOO_LONG;
    R.hdr_b = h.b; R.seq_b = s.b; R.qual_b = q.b;
    R.hdr_len = (uint16_t)(h.e - h.b); R.seq_len = (uint16_t)slen; R.name2 = (uint16_t)name2;
    bad_line = 0;
    return ATR_FQ_OK;
}

// One round of AdapterCutter.__call__ for one read: the match record `m` (coordinates relative to the window
// [lo, hi) the previous rounds left) -> new window and the statistics bin. front_flag: 1 FRONT/PREFIX, 0 BACK/SUFFIX,
// -1 ANYWHERE (front iff rstart == 0). Returns false if the record is not a match (the loop ends for this read).
struct FqApply {
    int front;           // which histogram
    int length;          // lengths_front[rstop] / lengths_back[len(read) - rstart]
    int errors;
    int adjacent;        // 0..3 = A C G T, 4 = '' (anything else, or no base before the adapter); back only
    int new_lo, new_hi;
};

ATR_HD bool fq_apply(const atr_match& m, int front_flag, int lo, int hi, const unsigned char* __restrict__ seq, FqApply& a) {
    if (m.status != ATR_ST_MATCH) return false;
    const int rstart = m.rstart, rstop = m.rstop;
    a.front = front_flag < 0 ? (rstart == 0) : front_flag;
    a.errors = m.errors;
    a.adjacent = 4;
    if (a.front) {
        a.length = rstop;
        a.new_lo = lo + rstop; a.new_hi = hi;
    } else {
        a.length = (hi - lo) - rstart;
        if (rstart >= 1) {
            const unsigned char c = seq[lo + rstart - 1];          // the original letter: 'a' is not in 'ACGT'
            a.adjacent = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
        }
        a.new_lo = lo; a.new_hi = lo + rstart;
    }
    return true;
}

// '@' name '\n' sequence '\n+' name2 '\n' qualities '\n'
ATR_HD uint32_t fq_out_len(const FqRec& R, int lo, int hi) {
    const uint32_t w = (uint32_t)(hi - lo);
    return (uint32_t)R.hdr_len + 1u + w + 1u + (R.name2 ? (uint32_t)R.hdr_len : 1u) + 1u + w + 1u;
}

// byte i of the formatted record
ATR_HD unsigned char fq_out_byte(const unsigned char* __restrict__ text, const FqRec& R, int lo, int hi, uint32_t i) {
    const uint32_t w = (uint32_t)(hi - lo), H = R.hdr_len;
    if (i < H) return text[R.hdr_b + i];
    i -= H;
    if (i == 0) return '\n';
    i -= 1;
    if (i < w) return text[R.seq_b + (uint32_t)lo + i];
    i -= w;
    if (i == 0) return '\n';
    i -= 1;
    const uint32_t P = R.name2 ? H : 1u;
    if (i < P) return i == 0 ? (unsigned char)'+' : text[R.hdr_b + i];
    i -= P;
    if (i == 0) return '\n';
    i -= 1;
    if (i < w) return text[R.qual_b + (uint32_t)lo + i];
    return '\n';
}

// ---- paired-end ("--aligner insert") ------------------------------------------------------------------------------
// sequence_names_match (atropos/io/seqio.py:773-791): first whitespace-delimited token of each name, a trailing '1' /
// '2' dropped when both have one. Returns 0 equal, 1 different, 2 a name without any token (the reference dies with
// an IndexError there). Whitespace = the ASCII characters str.split() splits on.
ATR_HD bool fq_is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }

ATR_HD void fq_name_token(const unsigned char* __restrict__ text, const FqRec& R, uint32_t& b, uint32_t& e) {
    b = R.hdr_b + 1u;
    const uint32_t end = R.hdr_b + R.hdr_len;
    while (b < end && fq_is_space(text[b])) b++;
    e = b;
    while (e < end && !fq_is_space(text[e])) e++;
}

ATR_HD int fq_names_match(const unsigned char* __restrict__ t1, const FqRec& R1, const unsigned char* __restrict__ t2, const FqRec& R2) {
    uint32_t b1, e1, b2, e2;
    fq_name_token(t1, R1, b1, e1);
    fq_name_token(t2, R2, b2, e2);
    if (e1 == b1 || e2 == b2) return 2;
    const unsigned char l1 = t1[e1 - 1], l2 = t2[e2 - 1];
    if ((l1 == '1' || l1 == '2') && (l2 == '1' || l2 == '2')) { e1--; e2--; }
    if (e1 - b1 != e2 - b2) return 1;
    for (uint32_t i = 0; i < e1 - b1; i++) if (t1[b1 + i] != t2[b2 + i]) return 1;
    return 0;
}

// InsertAdapterCutter.__call__ (commands/trim/modifiers.py:391-453) with mismatch_action None, for one pair: which
// match each read ends up with. ins = InsertAligner.match_insert's result, fb1/fb2 = adapter{1,2}.match_to(read{1,2})
// (only consulted without an insert match, :401-406).
struct PeMatch { int present, rstart, rstop, astop, errors; };

ATR_HD void fq_pe_load(const atr_match& m, PeMatch& p, int& invalid) {
    p.present = m.status == ATR_ST_MATCH;
    if (m.status == ATR_ST_INVALID || m.status == ATR_ST_KEYERROR) invalid = 1;
    p.rstart = m.rstart; p.rstop = m.rstop; p.astop = m.astop; p.errors = m.errors;
}

// create_symmetric_match (:421-433)
ATR_HD void fq_pe_symmetric(const PeMatch& src, int read_len, PeMatch& dst) {
    dst = src;
    if (src.rstart > read_len) { dst.present = 0; return; }
    if (dst.rstop < read_len) { dst.astop -= (read_len - dst.rstop); dst.rstop = read_len; }
}

// correct / im: whether ErrorCorrectorMixin.correct_errors runs for this pair and with which insert_match[0..3]
// (mismatch_action set: :399-400 an insert match with mismatches; :407-414 complementary adapter matches; :439-446
// after the symmetric duplication)
ATR_HD void fq_pe_decide(const atr_insert_result& ins, const atr_match& fb1, const atr_match& fb2, int len1, int len2,
                         int min_insert_len, int symmetric, int action, PeMatch& m1, PeMatch& m2, int& insert_hit, int& invalid,
                         bool& correct, int* im) {
    m1.present = m2.present = 0; m1.rstart = m1.rstop = m1.astop = m1.errors = 0; m2 = m1;
    insert_hit = 0;
    correct = false;
    bool have_im = false;
    im[0] = im[1] = im[2] = im[3] = 0;
    if (len1 < min_insert_len || len2 < min_insert_len) return;          // :392-394
    if (ins.insert.status == ATR_ST_INVALID || ins.insert.status == ATR_ST_KEYERROR) invalid = 1;
    if (ins.insert.status == ATR_ST_MATCH) {
        insert_hit = 1;
        fq_pe_load(ins.match1, m1, invalid);
        fq_pe_load(ins.match2, m2, invalid);
        have_im = true;
        im[0] = ins.insert.astart; im[1] = ins.insert.astop; im[2] = ins.insert.rstart; im[3] = ins.insert.rstop;
        correct = action != 0 && ins.insert.errors > 0;
    } else {
        fq_pe_load(fb1, m1, invalid);
        fq_pe_load(fb2, m2, invalid);
        if (action != 0 && m1.present && m2.present && m1.rstart == m2.rstart) {
            im[0] = len2 - m1.rstart; im[1] = len2; im[2] = 0; im[3] = m1.rstart;
            have_im = true;
            correct = true;
        }
    }
    if (symmetric && (m1.present + m2.present) == 1) {                    // :417-437
        if (m1.present) fq_pe_symmetric(m1, len2, m2);
        else fq_pe_symmetric(m2, len1, m1);
        if (action != 0 && !have_im && m1.present && m2.present) {
            im[0] = len2 - m1.rstart; im[1] = len2; im[2] = 0; im[3] = m1.rstart;
            correct = true;
        }
    }
}

// InsertAdapterCutter.trim (:455-496), action 'trim', the adapter being a 3' (BACK) adapter: match.front = False,
// no trimming (and no statistics) when the match starts at or beyond the read end. Returns the new read length.
ATR_HD int fq_pe_trim(const PeMatch& m, int len, const unsigned char* __restrict__ seq, FqApply& a, bool& counted) {
    counted = false;
    if (!m.present || m.rstart >= len) return len;
    counted = true;
    a.front = 0;
    a.length = len - m.rstart;
    a.errors = m.errors;
    a.adjacent = 4;
    if (m.rstart >= 1) {
        const unsigned char c = seq[m.rstart - 1];
        a.adjacent = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
    }
    a.new_lo = 0; a.new_hi = m.rstart;
    return m.rstart;
}

// ---- the modifiers and filters around the adapter stage (atr_read_ops) ------------------------------------------------
// Device-side counters, one block per call.
struct FqOpsCounters {
    unsigned long long bp_cut[2], bp_quality[2], bp_n_ends[2], bp_nextseq[2];
    unsigned long long too_short, too_long, too_many_n, discarded_trimmed, discarded_untrimmed;
    unsigned long long records_written;
};

// quality_trim_index (commands/trim/_qualtrim.pyx:7-49), the BWA rule on both ends
ATR_HD void fq_quality_trim_index(const unsigned char* __restrict__ q, int len, int cutoff_front, int cutoff_back, int base,
                                  int& start, int& stop) {
    start = 0; stop = len;
    int s = 0, max_qual = 0;
    for (int i = 0; i < len; i++) {
        s += cutoff_front - ((int)q[i] - base);
        if (s < 0) break;
        if (s > max_qual) { max_qual = s; start = i + 1; }
    }
    max_qual = 0; s = 0;
    for (int i = len - 1; i >= 0; i--) {
        s += cutoff_back - ((int)q[i] - base);
        if (s < 0) break;
        if (s > max_qual) { max_qual = s; stop = i; }
    }
    if (start >= stop) { start = 0; stop = 0; }
}

// nextseq_trim_index (commands/trim/_qualtrim.pyx:52-84): the 3' BWA rule with every 'G' counted as quality cutoff - 1
ATR_HD int fq_nextseq_trim_index(const unsigned char* __restrict__ seq, const unsigned char* __restrict__ q, int len, int cutoff, int base) {
    int s = 0, max_qual = 0, max_i = len;
    for (int i = len - 1; i >= 0; i--) {
        int qq = (int)q[i] - base;
        if (seq[i] == 'G') qq = cutoff - 1;
        s += cutoff - qq;
        if (s < 0) break;
        if (s > max_qual) { max_qual = s; max_i = i; }
    }
    return max_i;
}

// UnconditionalCutter, NextseqQualityTrimmer, then QualityTrimmer on one record, BEFORE the adapters: the record table entry is narrowed to what
// is left (nothing downstream needs the removed ends). bp_cut / bp_quality get Trimmer.trimmed_bases' increments:
// clip() counts the nominal lengths (modifiers.py:73-82, _seqio.pyx:76-88), subseq() begin + (len - end) (:54-71).
ATR_HD void fq_pre_ops(const atr_read_ops& o, int side, const unsigned char* __restrict__ text, FqRec& R,
                       unsigned& bp_cut, unsigned& bp_quality, unsigned& bp_nextseq) {
    bp_cut = bp_quality = bp_nextseq = 0;
    int lo = 0, hi = R.seq_len;
    const int front = o.cut_front[side], back = o.cut_back[side];
    if ((front || back) && hi - lo > 0) {
        const int L = hi - lo;
        int a = front < L ? front : L;                 // read[front:back] / read[front:]
        int b = back < 0 ? (L + back > 0 ? L + back : 0) : L;
        if (a > b) a = b;
        bp_cut = (unsigned)(front + (back < 0 ? -back : 0));
        hi = lo + b; lo = lo + a;
    }
    if (o.nextseq_trim[side] >= 0 && hi - lo > 0) {          // subseq(read, end=stop): counts len - stop (modifiers.py:742-746)
        const int stop = fq_nextseq_trim_index(text + R.seq_b + lo, text + R.qual_b + lo, hi - lo, o.nextseq_trim[side], o.quality_base);
        bp_nextseq = (unsigned)((hi - lo) - stop);
        hi = lo + stop;
    }
    if ((o.quality_front > 0 || o.quality_back > 0) && hi - lo > 0) {
        int start, stop;
        fq_quality_trim_index(text + R.qual_b + lo, hi - lo, o.quality_front, o.quality_back, o.quality_base, start, stop);
        bp_quality = (unsigned)(start + ((hi - lo) - stop));
        hi = lo + stop; lo = lo + start;
    }
    R.seq_b += (uint32_t)lo; R.qual_b += (uint32_t)lo; R.seq_len = (uint16_t)(hi - lo);
}

// NEndTrimmer on the window the adapters left (modifiers.py:776-784): leading and trailing runs of 'N'
ATR_HD void fq_trim_n(const unsigned char* __restrict__ seq, int& lo, int& hi, unsigned& bp_n) {
    bp_n = 0;
    const int L = hi - lo;
    if (L == 0) return;
    int start = 0;
    while (start < L && seq[lo + start] == 'N') start++;
    int end = L;
    while (end > 0 && seq[lo + end - 1] == 'N') end--;
    // re: '^N+' and 'N+$' are found independently: an all-N read gives start = L and end = 0, both counted
    bp_n = (unsigned)(start + (L - end));
    int a = start, b = end;
    if (a > b) { a = 0; b = 0; }
    hi = lo + b; lo = lo + a;
}

// one read against the filters; returns 0 keep, or 1 too_short, 2 too_long, 3 too_many_n, 4 trimmed, 5 untrimmed: the
// first filter that fires, in the order the command adds them (commands/trim/__init__.py:566-620)
ATR_HD int fq_filter_one(const atr_read_ops& o, const unsigned char* __restrict__ seq, int lo, int hi, bool matched, int which) {
    const int L = hi - lo;
    if (which == 1) return o.minimum_length > 0 && L < o.minimum_length;
    if (which == 2) return o.maximum_length >= 0 && L > o.maximum_length;
    if (which == 3) {
        if (!(o.max_n >= 0.0)) return 0;
        int n = 0;
        for (int i = lo; i < hi; i++) n += (seq[i] == 'N' || seq[i] == 'n');
        if (o.max_n < 1.0) return L == 0 ? 0 : ((double)n / (double)L > o.max_n);
        return (double)n > o.max_n;
    }
    if (which == 4) return o.discard_trimmed && matched;
    return o.discard_untrimmed && !matched;
}

// single-end: SingleWrapper; paired-end: PairedWrapper with min_affected = 1 == either read (filters.py:54-95)
ATR_HD int fq_filter(const atr_read_ops& o, const unsigned char* __restrict__ seq1, int lo1, int hi1, bool matched1,
                     const unsigned char* __restrict__ seq2, int lo2, int hi2, bool matched2, bool paired) {
    for (int which = 1; which <= 5; which++) {
        const int f1 = fq_filter_one(o, seq1, lo1, hi1, matched1, which);
        if (!paired || o.legacy_first) { if (f1) return which; continue; }
        if (o.pair_filter_both) {                        // PairedWrapper(min_affected = 2), filters.py:86-95
            if (f1 && fq_filter_one(o, seq2, lo2, hi2, matched2, which)) return which;
        } else if (f1 || fq_filter_one(o, seq2, lo2, hi2, matched2, which)) return which;
    }
    return 0;
}

// ---- ErrorCorrectorMixin.correct_errors(read1, read2, insert_match, truncate_seqs=True) ---------------------------
// (commands/trim/modifiers.py:219-350), in place on the two reads' bases and qualities. im = insert_match[0..3].
// action: 1 liberal, 2 conservative, 3 'N'; min_qual_difference is the constructor default 1 (:212-216).
// Python semantics kept where the reference leans on them: list indices may be negative (wrap), slices clamp.
// Returns false where the reference would raise; changed1 / changed2 = bases changed; new_len1 = read 1's length
// afterwards (the reference re-assembles a changed read 1 from the TRUNCATED list when it is the longer read, :328-336
// with len1 never updated by :260-269).
ATR_HD bool fq_py_index(int& i, int L) { if (i < 0) i += L; return i >= 0 && i < L; }
ATR_HD void fq_py_slice(int& a, int& b, int L) {
    if (a < 0) { a += L; if (a < 0) a = 0; } else if (a > L) a = L;
    if (b < 0) { b += L; if (b < 0) b = 0; } else if (b > L) b = L;
    if (b < a) b = a;
}

ATR_HD bool fq_pe_correct(unsigned char* s1, unsigned char* q1, int len1_full, unsigned char* s2, unsigned char* q2, int len2_full,
                          int im0, int im1, int im2, int im3, int action, const unsigned char* __restrict__ comp,
                          int& changed1, int& changed2, int& new_len1, bool truncate = true) {
    changed1 = changed2 = 0;
    new_len1 = len1_full;
    int L1 = len1_full, L2 = len2_full, len2 = len2_full;
    if (truncate) {                                      // truncate_seqs=True (:246-255): InsertAdapterCutter; MergeOverlapping passes False
        if (len1_full > len2_full) L1 = len2_full;
        else if (len2_full > len1_full) { L2 = len1_full; len2 = len1_full; }
    }
    const int r1_start = im2, r1_end = im3, r2_start = len2 - im1, r2_end = len2 - im0;
    int n1 = r1_end - r1_start, n2 = r2_end - r2_start;
    if (n1 < 0) n1 = 0;
    if (n2 < 0) n2 = 0;
    const int npairs = n1 < n2 ? n1 : n2;
    const int mqd = 1;
    int n_equal = 0;
    for (int pass = 0; pass < 2; pass++) {
        int toward = 0;                                  // pass 1 (liberal ties): 1 = read 1 is better, 2 = read 2 is better
        if (pass == 1) {
            if (n_equal == 0) break;
            int a1 = r1_start, b1 = r1_end, a2 = r2_start, b2 = r2_end;
            fq_py_slice(a1, b1, L1);
            fq_py_slice(a2, b2, L2);
            if (b1 == a1 || b2 == a2) return false;      // mean() of an empty sequence raises
            long long sum1 = 0, sum2 = 0;
            for (int i = a1; i < b1; i++) sum1 += q1[i];
            for (int j = a2; j < b2; j++) sum2 += q2[j];
            const double diff = (double)sum1 / (double)(b1 - a1) - (double)sum2 / (double)(b2 - a2);
            if (diff > 1) toward = 1; else if (diff < -1) toward = 2; else break;
        }
        for (int t = 0; t < npairs; t++) {
            int i = r1_start + t, j = r2_end - 1 - t;
            if (!fq_py_index(i, L1) || !fq_py_index(j, L2)) return false;
            const unsigned char base1 = s1[i];
            const unsigned char base2 = comp[s2[j]];
            if (base2 == 0) return false;                // KeyError
            if (base1 == base2) continue;
            if (pass == 0) {
                if (action == 3) { s1[i] = 'N'; s2[j] = 'N'; changed1++; changed2++; }
                else if (base1 == 'N') { s1[i] = base2; q1[i] = q2[j]; changed1++; }
                else if (base2 == 'N') { const unsigned char c = comp[base1]; if (c == 0) return false; s2[j] = c; q2[j] = q1[i]; changed2++; }
                else {
                    const int diff = (int)q1[i] - (int)q2[j];
                    if (diff >= mqd) { const unsigned char c = comp[base1]; if (c == 0) return false; s2[j] = c; q2[j] = q1[i]; changed2++; }
                    else if (diff <= -mqd) { s1[i] = base2; q1[i] = q2[j]; changed1++; }
                    else if (action == 1) n_equal++;
                }
            } else {
                // the ties left by pass 0: still mismatching, neither base N, qualities within the margin
                if (base1 == 'N' || base2 == 'N') continue;
                const int diff = (int)q1[i] - (int)q2[j];
                if (diff >= mqd || diff <= -mqd) continue;
                if (toward == 1) { const unsigned char c = comp[base1]; if (c == 0) return false; s2[j] = c; q2[j] = q1[i]; changed2++; }
                else { s1[i] = base2; q1[i] = q2[j]; changed1++; }
            }
        }
        if (action != 1) break;
    }
    if (truncate && changed1 && len1_full > len2_full) new_len1 = len2_full;
    return true;
}

// ---- MergeOverlapping behind the paired-end modifiers (commands/trim/modifiers.py:864-931) -------------------------
// What one merged pair needs for its record in the merged output: the two reads' windows after trimming (relative to
// FqRec.seq_b / qual_b), where the alignment stops in each (r1_stop in read 1, r2_stop in rc(read 2)) and the action.
struct FqMergeRec {                   // 16 bytes
    uint16_t lo1, hi1, lo2, hi2;
    uint16_t r1_stop, r2_stop;
    uint16_t mlen;                    // length of the merged read
    uint8_t action;                   // ATR_MERGE_* (merge_core.cuh), 0 = the pair was not merged
    uint8_t pad;
};
// per-pair flags handed from the adapter stage to the merge stage
#define FQ_PF_MATCH1     1            // read 1 has an adapter match (TrimmedFilter / UntrimmedFilter look at it)
#define FQ_PF_MATCH2     2
#define FQ_PF_INSERT     4            // read.insert_overlap: match_insert returned a match (modifiers.py:397)
#define FQ_PF_CORRECTED1 8            // read.corrected > 0 (modifiers.py:232-233, set by :333)
#define FQ_PF_CORRECTED2 16

// length of the merged read (:905-922); action 1 keep read 1, 2 take rc(read 2), 3 read 1 + rc(read 2)[r2_stop:],
// 4 rc(read 2) + read 1[r1_stop:]
ATR_HD int fq_merged_len(int action, int len1, int len2, int r1_stop, int r2_stop) {
    if (action == 1) return len1;
    if (action == 2) return len2;
    if (action == 3) return len1 + (len2 - r2_stop);
    return len2 + (len1 - r1_stop);
}
// which base / quality of which read is position i of the merged read: src = 1 -> read 1 position j, 2 -> position j of
// rc(read 2) = complement of read 2's base len2 - 1 - j (quality: read 2's quality at len2 - 1 - j)
ATR_HD void fq_merged_source(int action, int len1, int len2, int r1_stop, int r2_stop, int i, int& src, int& j) {
    if (action == 1) { src = 1; j = i; }
    else if (action == 2) { src = 2; j = i; }
    else if (action == 3) { if (i < len1) { src = 1; j = i; } else { src = 2; j = r2_stop + (i - len1); } }
    else { if (i < len2) { src = 2; j = i; } else { src = 1; j = r1_stop + (i - len2); } }
}
// byte i of the merged read's FASTQ record: read 1's header, the merged bases, '+' (and the name again if the input
// repeated it), the merged qualities. s1 / q1 / q2: the reads' windows in the chunk text AFTER error correction;
// s2: read 2's window BEFORE it (the reference computes read2_rc before it corrects, :891 vs :901-903, and a merged
// pair's read 2 is never written, so only this copy is ever used); hdr: read 1's header line.
ATR_HD unsigned char fq_merged_out_byte(const unsigned char* __restrict__ hdr, int H, int name2, const unsigned char* __restrict__ s1,
                                        const unsigned char* __restrict__ q1, const unsigned char* __restrict__ s2,
                                        const unsigned char* __restrict__ q2, const unsigned char* __restrict__ comp,
                                        const FqMergeRec& M, uint32_t i) {
    const uint32_t w = M.mlen;
    const int len1 = (int)M.hi1 - (int)M.lo1, len2 = (int)M.hi2 - (int)M.lo2;
    if (i < (uint32_t)H) return hdr[i];
    i -= (uint32_t)H;
    if (i == 0) return '\n';
    i -= 1;
    int src, j;
    if (i < w) {
        fq_merged_source(M.action, len1, len2, M.r1_stop, M.r2_stop, (int)i, src, j);
        return src == 1 ? s1[j] : comp[s2[len2 - 1 - j]];
    }
    i -= w;
    if (i == 0) return '\n';
    i -= 1;
    const uint32_t P = name2 ? (uint32_t)H : 1u;
    if (i < P) return i == 0 ? (unsigned char)'+' : hdr[i];
    i -= P;
    if (i == 0) return '\n';
    i -= 1;
    if (i < w) {
        fq_merged_source(M.action, len1, len2, M.r1_stop, M.r2_stop, (int)i, src, j);
        return src == 1 ? q1[j] : q2[len2 - 1 - j];
    }
    return '\n';
}
ATR_HD uint32_t fq_merged_out_len(int H, int name2, int mlen) {
    return (uint32_t)H + 1u + (uint32_t)mlen + 1u + (name2 ? (uint32_t)H : 1u) + 1u + (uint32_t)mlen + 1u;
}

// one pair after the merge alignment: the decision, the optional correction of the overlap, the record for the merged
// output. res: atr_merge_result of the pair's windows. Returns 0 not merged, 1 merged, -1 the reference raises
// (KeyError / AtroposError), -2 correction raises. c1 / c2: bases corrected in read 1 / read 2.
ATR_HD int fq_merge_decide(const atr_merge_result& res, int pflags, int mismatch_action, unsigned char* t1, const FqRec& A, int lo1, int hi1,
                           unsigned char* t2, const FqRec& B, int lo2, int hi2, const unsigned char* __restrict__ comp,
                           FqMergeRec& M, int& c1, int& c2) {
    c1 = c2 = 0;
    M.lo1 = (uint16_t)lo1; M.hi1 = (uint16_t)hi1; M.lo2 = (uint16_t)lo2; M.hi2 = (uint16_t)hi2;
    M.r1_stop = res.r1_stop; M.r2_stop = res.r2_stop; M.mlen = 0; M.action = 0; M.pad = 0;
    if (res.status == ATR_ST_KEYERROR || res.status == ATR_ST_INVALID) return -1;
    if (res.status != ATR_ST_MATCH) return 0;
    const int len1 = hi1 - lo1, len2 = hi2 - lo2;
    if (mismatch_action && res.errors > 0 && !(pflags & FQ_PF_INSERT) && !(pflags & (FQ_PF_CORRECTED1 | FQ_PF_CORRECTED2))) {
        int nl1;
        if (!fq_pe_correct(t1 + A.seq_b + lo1, t1 + A.qual_b + lo1, len1, t2 + B.seq_b + lo2, t2 + B.qual_b + lo2, len2,
                           res.r2_start, res.r2_stop, res.r1_start, res.r1_stop, mismatch_action, comp, c1, c2, nl1, false)) return -2;
    }
    M.action = res.action;
    M.mlen = (uint16_t)fq_merged_len(res.action, len1, len2, res.r1_stop, res.r2_stop);
    return 1;
}
