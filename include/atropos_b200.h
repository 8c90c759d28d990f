/*
 * atropos_b200.h -- C ABI of the B200-native adapter-alignment engine.
 *
 * This is the drop-in boundary for ONE hot path of jdidion/atropos: the compiled extension
 * module `atropos.align._align` (Cython) and the two Python methods that call it per read,
 * `Adapter.match_to()` and `InsertAligner.match_insert()` -- plus, at the end of this file, the
 * path's immediate callers as whole-batch calls (FASTQ text in, trimmed FASTQ text out:
 * atr_trim_fastq_host, atr_trim_fastq_pe_host).  Every entry point below names the
 * reference interface it replaces (paths relative to the reference checkout).  The reference-side
 * binding a maintainer would add (a ctypes stub inside atropos/align/__init__.py) is shown in
 * INTEGRATION.md; `atropos_b200/` in this repository is that binding plus batched twins of the
 * reference classes.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types; every function returns 0 on success
 *     or a negative ATR_E_* code, and atr_last_error(ctx) gives a message;
 *   - the caller owns every buffer it passes; `*_host` entry points take HOST pointers (pageable
 *     or pinned) and do the H2D/D2H copies themselves, `*_device` entry points take DEVICE
 *     pointers on the ctx's device and only enqueue kernels on the ctx's stream (a non-blocking stream:
 *     it does not synchronise with the legacy default stream, so inputs produced on other streams must be
 *     complete before the call);
 *   - one atr_ctx per GPU per host thread; no global state; a ctx is not re-entrant (the
 *     reference's Aligner is not either: it owns one DP column, _align.pyx:184, :239-242);
 *   - there is NO CPU fallback: without a CUDA device atr_ctx_create fails with ATR_E_CUDA.
 *
 * Data layout in HBM ("packed reads")
 *   codes   4-bit IUPAC code per base (A=1 C=2 G=4 T=8, unions for R,Y,S,W,K,M,B,D,H,V, N=15,
 *           X=0 -- the bit assignment of _align.pyx:46-83), two bases per byte, base j of a read
 *           in bits 4*(j%8) .. 4*(j%8)+3 of 32-bit word j/8; every read starts on a word boundary.
 *   woff    uint32 per read: index of its first 32-bit word inside `codes`.
 *   len     uint16 per read: length in bases; bit 15 set = "escaped" read: it holds a byte
 *           that the 4-bit code cannot represent exactly for ASCII-compare adapters (anything
 *           outside the 16 upper-case IUPAC letters, e.g. 'U', '.', lower case when case folding
 *           is off). Escaped reads are aligned by the byte-exact kernel from the raw ASCII.
 *   result  one 16-byte atr_match per read.
 */
#ifndef ATROPOS_B200_H
#define ATROPOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATR_ABI_VERSION 1

/* error codes */
#define ATR_OK          0
#define ATR_E_ARG      -1   /* invalid argument (ValueError on the Python side, cf. _align.pyx:214-232) */
#define ATR_E_CUDA     -2   /* CUDA runtime error / no device */
#define ATR_E_NOMEM    -3   /* allocation failed (MemoryError, _align.pyx:240-241) */
#define ATR_E_LIMIT    -4   /* size outside what the kernels support (read > 32767 nt, adapter > 4095 nt) */
#define ATR_E_FORMAT   -5   /* malformed FASTQ: the reader's FormatError (io/_seqio.pyx:180-245); see atr_fastq_error */

/* alignment flags: atropos/align/__init__.py:17-26, _align.pyx:12-16 */
#define ATR_START_WITHIN_SEQ1 1
#define ATR_START_WITHIN_SEQ2 2
#define ATR_STOP_WITHIN_SEQ1  4
#define ATR_STOP_WITHIN_SEQ2  8
#define ATR_SEMIGLOBAL       15

/* atr_match.status */
#define ATR_ST_NONE    0    /* no acceptable alignment: locate() returned None / match_to() returned None */
#define ATR_ST_MATCH   1
#define ATR_ST_ESCAPED 2    /* the read needs the byte-exact kernel but no ASCII was supplied (device entry points only) */
#define ATR_ST_INVALID 3    /* Match.__init__ would raise ValueError (align/__init__.py:85-88) */
#define ATR_ST_KEYERROR 4   /* reverse_complement would raise KeyError: a byte outside the IUPAC table (util/__init__.py:479-482) */

/* One alignment result: the 6-tuple of Aligner.locate (_align.pyx:491) == the fields of
 * Match (align/__init__.py:51-88), plus which adapter of the set won (AdapterCutter._best_match,
 * commands/trim/modifiers.py:107-122). 16 bytes. */
typedef struct atr_match {
    uint16_t astart, astop;   /* refstart, refstop   (within the adapter) */
    uint16_t rstart, rstop;   /* querystart, querystop (within the read / window) */
    uint16_t matches, errors;
    int16_t  adapter;         /* index into the adapter set, -1 if none */
    uint16_t status;          /* ATR_ST_* */
} atr_match;

/* Result of InsertAligner.match_insert (align/__init__.py:250-377) for one pair. 48 bytes.
 * status: ATR_ST_NONE -> returned None; ATR_ST_MATCH -> (insert_match, match1, match2) where
 * match{1,2}.status == ATR_ST_NONE encodes Python None (the "(insert_match, None, None)" case). */
typedef struct atr_insert_result {
    atr_match insert;         /* the winning MultiAligner tuple; .adapter unused */
    atr_match match1, match2; /* Match(0, alen, insert_size, slen, alen-mm, mm) */
} atr_insert_result;

/* Constructor arguments of Aligner (_align.pyx:197-208) + the post-filter of Adapter.match_to
 * (adapters/__init__.py:338-400). */
typedef struct atr_adapter_desc {
    const char* sequence;     /* ASCII, `length` bytes, as the reference would hold it (Adapter upper-cases and U->T first) */
    int32_t length;
    double  max_error_rate;
    int32_t flags;            /* ATR_* flags; BACK=14 FRONT=11 PREFIX=8 SUFFIX=2 ANYWHERE=15 */
    int32_t wildcard_ref;     /* adapter_wildcards */
    int32_t wildcard_query;   /* read_wildcards */
    int32_t min_overlap;      /* >= 1 (ATR_E_ARG otherwise) */
    int32_t indel_cost;       /* >= 1; 100000 is how the reference spells "--no-indels" */
    /* 0: raw Aligner.locate semantics.
     * 1: Adapter.match_to semantics: afterwards require size >= min_overlap and errors/size <= rate
     *    (adapters/__init__.py:386-392), the rmp_ok gate unless the hit is the exact-find shortcut
     *    (:351-367), and anchored no-indel adapters use compare_prefixes/suffixes (:370-380). */
    int32_t match_to_semantics;
    int32_t no_indels;        /* Adapter(indels=False): only meaningful with match_to_semantics */
    /* optional max_rmp gate: rmp_ok[size*(length+1)+matches] != 0 iff
     * match_probability(matches,size) <= max_rmp, for 0 <= matches,size <= length; NULL = no gate.
     * Computed by the host binding with the reference's own big-int arithmetic
     * (util/__init__.py:117-155). */
    const uint8_t* rmp_ok;
} atr_adapter_desc;

/* Constructor arguments of InsertAligner (align/__init__.py:206-233). The probability tables are
 * computed by the host binding with RandomMatchProbability (util/__init__.py:104-174):
 *   insert_prob[size*(kmax+1)+cost]         = match_probability(size-cost,size,**base_probs) (:358)
 *                                             for size <= max_len, cost <= kmax = (int)(frac*max_len)
 *   adapter_prob[alen*(alen_max+1)+matches] = match_probability(matches,alen)               (:303-304)
 * with alen_max = max(len(adapter1), len(adapter2)). Only costs <= (int)(frac*size) are ever looked up. */
typedef struct atr_insert_desc {
    const char* adapter1; int32_t adapter1_len;
    const char* adapter2; int32_t adapter2_len;
    double  insert_max_rmp, adapter_max_rmp;
    int32_t min_insert_overlap;
    double  max_insert_mismatch_frac;
    int32_t min_adapter_overlap;
    double  max_adapter_mismatch_frac;
    int32_t adapter_check_cutoff;
    int32_t adapter_wildcards, read_wildcards;
    int32_t max_len;                 /* longest read the tables cover */
    const double* insert_prob;       /* (max_len+1)*(kmax+1) doubles */
    const double* adapter_prob;      /* (alen_max+1)^2 doubles */
} atr_insert_desc;

typedef struct atr_ctx atr_ctx;
typedef struct atr_adapterset atr_adapterset;
typedef struct atr_insertset atr_insertset;

/* ---- library / context -------------------------------------------------------------------- */
int         atr_abi_version(void);
int         atr_device_count(void);
/* Owns a CUDA stream, events and grow-only device/pinned scratch. Replaces the per-process
 * aligner state of the reference workers (commands/multicore.py:404-414). */
int         atr_ctx_create(int device, atr_ctx** out);
void        atr_ctx_destroy(atr_ctx* ctx);
const char* atr_last_error(const atr_ctx* ctx);          /* ctx may be NULL: last global error */
int         atr_ctx_sync(atr_ctx* ctx);
void*       atr_ctx_stream(atr_ctx* ctx);                /* cudaStream_t, for callers that time with events */
/* kernels launched on this ctx since creation / since the last reset (bench.py's gpu_launches) */
int64_t     atr_ctx_launch_count(atr_ctx* ctx, int reset);
/* device time in ms of the kernels of the last *_device / *_host call, measured with CUDA events
 * on the ctx stream around the kernel launches only (copies excluded) */
float       atr_ctx_last_kernel_ms(atr_ctx* ctx);
/* Per-kernel device times of the last atr_locate_batch_device call on the fast path (single adapter):
 * out_ms[0] = filter kernel, [1] = refine kernel, [2] = banded-DP kernel, [3] = windowed-DP kernel, measured with CUDA events
 * recorded between the launches once atr_ctx_set_profiling(ctx, 1) was called. Returns the number of
 * valid entries (0 if the last call did not take the fast path or profiling is off). */
int         atr_ctx_set_profiling(atr_ctx* ctx, int on);
int         atr_ctx_last_phase_ms(atr_ctx* ctx, float* out_ms, int n);
/* name of the kernel(s) timed by interval i (0..3) of the last profiled call, "" if none */
const char* atr_ctx_last_phase_name(atr_ctx* ctx, int i);

/* ---- adapters: Aligner.__cinit__ / Adapter.__init__ ---------------------------------------- */
/* replaces Aligner(reference, max_error_rate, flags, wildcard_ref, wildcard_query, min_overlap,
 * indel_cost) (_align.pyx:197-249), one per adapter of an AdapterCutter (modifiers.py:100-105). */
int  atr_adapterset_create(atr_ctx* ctx, int32_t n_adapters, const atr_adapter_desc* descs, atr_adapterset** out);
void atr_adapterset_destroy(atr_adapterset* set);

/* ---- reads: 4-bit packing ------------------------------------------------------------------ */
/* Size in 32-bit words of the packed form of n reads given their ASCII offsets (n+1 entries). */
int64_t atr_packed_words(const int64_t* offsets, int64_t n);
/* Device-side packer: d_ascii (bytes), d_offsets (int64[n+1]) -> d_codes/d_woff/d_len as described
 * above; d_woff must hold n+1 entries. fold_case != 0 applies str.upper() first
 * (adapters/__init__.py:349). Replaces `query.encode('ascii')` + bytes.translate (_align.pyx:281-297). */
int  atr_pack_device(atr_ctx* ctx, const uint8_t* d_ascii, const int64_t* d_offsets, int64_t n,
                     int fold_case, uint32_t* d_codes, uint32_t* d_woff, uint16_t* d_len);

/* Host-side twin of atr_pack_device (no GPU involved): the same layout in HOST memory -- codes must hold
 * atr_packed_words(offsets, n) words, woff n + 1 entries, len n entries. n_threads <= 0: every hardware thread.
 * For callers that keep reads packed (their own reader, a packed file format) and ship them with
 * atr_locate_batch_host_packed: 75 + 6 bytes per 150-nt read over PCIe instead of 150. */
int  atr_pack_reads_host(const uint8_t* ascii, const int64_t* offsets, int64_t n, int fold_case, int n_threads,
                         uint32_t* codes, uint32_t* woff, uint16_t* len);

/* ---- Aligner.locate / Adapter.match_to / AdapterCutter._best_match over a batch ------------- */
/* Device-resident form. For every read i (window d_win[2i], d_win[2i+1] if d_win != NULL, else the
 * whole read) aligns every adapter of the set and keeps the best (strictly more matches wins, the
 * first adapter on ties). d_ascii/d_offsets may be NULL if no read is escaped. Asynchronous on the
 * ctx stream. Replaces Aligner.locate (_align.pyx:266-491) called from Adapter.match_to
 * (adapters/__init__.py:382) inside AdapterCutter._best_match (modifiers.py:107-122). */
int  atr_locate_batch_device(atr_ctx* ctx, const atr_adapterset* set, const uint32_t* d_codes,
                             const uint32_t* d_woff, const uint16_t* d_len, const uint16_t* d_win,
                             const uint8_t* d_ascii, const int64_t* d_offsets, int fold_case, int64_t n,
                             atr_match* d_out);
/* Host form (what the Python binding calls): ASCII reads back to back + offsets[n+1] in host
 * memory, results into host memory; copies, packing and kernels inside. win may be NULL. */
int  atr_locate_batch_host(atr_ctx* ctx, const atr_adapterset* set, const uint8_t* ascii,
                           const int64_t* offsets, const uint16_t* win, int64_t n, int fold_case,
                           atr_match* out);

/* The same for reads that are packed already, in HOST memory (atr_pack_reads_host or the caller's own packer): only
 * codes / woff / len cross PCIe (fixed-length chunks: just the codes, the index is rebuilt on the device).
 * ascii / offsets: optional, only read for ESCAPED reads (len bit 15), whose bytes are sent along; NULL: escaped reads
 * come back with status ATR_ST_ESCAPED. */
int  atr_locate_batch_host_packed(atr_ctx* ctx, const atr_adapterset* set, const uint32_t* codes, const uint32_t* woff,
                                  const uint16_t* len, const uint16_t* win, const uint8_t* ascii, const int64_t* offsets,
                                  int fold_case, int64_t n, atr_match* out);

/* ---- compare_prefixes (_align.pyx:501-544) -------------------------------------------------- */
/* One (ref, query) pair per call: compare_prefixes(ref, query, wildcard_ref, wildcard_query);
 * out6 = (0, length, 0, length, matches, length - matches). compare_suffixes
 * (align/__init__.py:28-44) is this on reversed strings; the binding does the reversal. */
int  atr_compare_prefixes(atr_ctx* ctx, const char* ref, int32_t m, const char* query, int32_t n,
                          int wildcard_ref, int wildcard_query, int32_t* out6);

/* ---- MultiAligner.locate / InsertAligner.match_insert --------------------------------------- */
/* replaces InsertAligner.__init__ (align/__init__.py:206-233) */
int  atr_insertset_create(atr_ctx* ctx, const atr_insert_desc* desc, atr_insertset** out);
void atr_insertset_destroy(atr_insertset* set);
/* replaces InsertAligner.match_insert(seq1, seq2) (align/__init__.py:250-377), which wraps
 * reverse_complement (util/__init__.py:479-482), MultiAligner.locate (_align.pyx:593-772) and
 * compare_prefixes. Reads 1 and 2 are two packed batches of the same n. */
int  atr_match_insert_batch_device(atr_ctx* ctx, const atr_insertset* set,
                                   const uint32_t* d_codes1, const uint32_t* d_woff1, const uint16_t* d_len1,
                                   const uint32_t* d_codes2, const uint32_t* d_woff2, const uint16_t* d_len2,
                                   const uint8_t* d_ascii1, const int64_t* d_offsets1,
                                   const uint8_t* d_ascii2, const int64_t* d_offsets2,
                                   int64_t n, atr_insert_result* d_out);
int  atr_match_insert_batch_host(atr_ctx* ctx, const atr_insertset* set,
                                 const uint8_t* ascii1, const int64_t* offsets1,
                                 const uint8_t* ascii2, const int64_t* offsets2,
                                 int64_t n, atr_insert_result* out);
/* replaces MultiAligner(max_error_rate, flags, min_overlap).locate(reference, query, max_matches)
 * (_align.pyx:593-772) for one pair and any flag set: candidates in emission order (incl. the
 * duplicate of the last-column scan and the [exact] collapse). out6 must hold
 * 6*(max_matches+m+2) ints; *n_out = number of tuples (0 == None). */
int  atr_multi_locate(atr_ctx* ctx, const char* reference, int32_t m, const char* query, int32_t n,
                      double max_error_rate, int32_t flags, int32_t min_overlap, int32_t max_matches,
                      int32_t* out6, int32_t* n_out);

/* ---- MergeOverlapping -------------------------------------------------------------------- */
/* One pair's outcome of MergeOverlapping.__call__ (commands/trim/modifiers.py:864-931). 16 bytes.
 * status: ATR_ST_NONE  -> the pair stays as it is (a read shorter than min_overlap :881-882, locate() returned None, or
 *                         matches < min_overlap :900; the six alignment fields are still filled in when there was one);
 *         ATR_ST_MATCH -> merged (read1.merged = True, read 2 dropped), `action` says how read 1 is rebuilt;
 *         ATR_ST_INVALID  -> the reference raises AtroposError("Invalid alignment while trying to merge ...") (:923-927);
 *         ATR_ST_KEYERROR -> reverse_complement(read 2) raises KeyError (util/__init__.py:479-482).
 * action (ATR_ST_MATCH only): 1 read 2 inside read 1, read 1 unchanged (:905-907); 2 read 1 inside read 2, read 1 :=
 * rc(read 2) with reversed qualities (:908-911); 3 read 1 + rc(read 2)[r2_stop:] (:912-916); 4 rc(read 2) +
 * read 1[r1_stop:] (:917-922). */
typedef struct atr_merge_result {
    uint16_t r2_start, r2_stop;   /* refstart, refstop: within reverse_complement(read 2) */
    uint16_t r1_start, r1_stop;   /* querystart, querystop: within read 1 */
    uint16_t matches, errors;
    uint16_t min_overlap;         /* the pair's effective minimum: max(2, round(frac * min(len1, len2))) or int(value) (:877-879) */
    uint8_t  status, action;
} atr_merge_result;
/* replaces, for n pairs, the alignment and the decision of MergeOverlapping(min_overlap, error_rate).__call__: per pair
 * Aligner(reverse_complement(read2), error_rate, flags).locate(read1) (_align.pyx:197-208, :266-491) with flags
 * SEMIGLOBAL, or START_WITHIN_SEQ1 | STOP_WITHIN_SEQ2 where insert_matched[i] != 0 (read.insert_overlap set by
 * InsertAdapterCutter, modifiers.py:884-893); insert_matched may be NULL (all 0). min_overlap as on the command line
 * (--merge-min-overlap: a fraction of the shorter read if <= 1, else a number of bases). Reads up to 4000 nt. */
int  atr_merge_overlap_batch_host(atr_ctx* ctx, const uint8_t* ascii1, const int64_t* offsets1,
                                  const uint8_t* ascii2, const int64_t* offsets2, const uint8_t* insert_matched,
                                  int64_t n, double min_overlap, double error_rate, atr_merge_result* out);

/* ---- FASTQ text in -> trimmed FASTQ text out ("next" rows of the hot path: its reader and its consumer) ---- */
/* atr_fastq_error.kind: the FormatErrors of FastqReader.__iter__ (io/_seqio.pyx:180-245) */
#define ATR_FQ_OK            0
#define ATR_FQ_NO_AT         1   /* "Line 1 in FASTQ file is expected to start with '@', but found ..." (:192-195, :202-206) */
#define ATR_FQ_NO_PLUS       2   /* "Line 3 in FASTQ file is expected to start with '+', but found ..." (:214-217) */
#define ATR_FQ_NAME_MISMATCH 3   /* "At line 3: Sequence descriptions in the FASTQ file don't match" (:219-226) */
#define ATR_FQ_LENGTH        4   /* "Error creating sequence record at line 4": qualities and sequence differ in length (:33-44, :236-241) */
#define ATR_FQ_TRUNCATED     5   /* "FASTQ file ended prematurely" (:244-245) */
#define ATR_FQ_BARE_CR       6   /* a '\r' not followed by '\n' (universal-newline splitting is not reproduced: refused) */
#define ATR_FQ_TOO_LONG      7   /* read > 32767 nt or header > 65535 bytes, or one record larger than a chunk */
#define ATR_FQ_INVALID_MATCH 8   /* Match.__init__ would raise ValueError (align/__init__.py:85-88), or reverse_complement KeyError */
#define ATR_FQ_MORE_IN_1     9   /* "Reads are improperly paired. There are more reads in file 1 than in file 2." (io/seqio.py:444-447) */
#define ATR_FQ_MORE_IN_2    10   /* "... more reads in file 2 than in file 1." (:436-440) */
#define ATR_FQ_PAIR_NAMES   11   /* "Read name '..' in file 1 does not match '..' in file 2." (:448-452, sequence_names_match :773-791) */
#define ATR_FQ_EMPTY_NAME   12   /* a read name without any token: the reference dies with an IndexError in sequence_names_match */
#define ATR_FQ_CORRECTION   13   /* error correction would raise in the reference (IndexError / KeyError / ValueError on reads of
                                    unequal length or bytes outside the complement table) */

typedef struct atr_fastq_error {
    int32_t kind;             /* ATR_FQ_* of the FIRST error in file order */
    int32_t line_in_record;   /* 0..3 */
    int64_t record;           /* index of the record within this call's text */
    int64_t line_begin;       /* byte offsets of the offending line's content within this call's text */
    int64_t line_end;         /* (terminator excluded) */
    int32_t terminated;       /* 1: the line ends with a newline, 0: the text ended first */
    int32_t file;             /* paired-end: 0 = the error is in file 1, 1 = file 2; line offsets are within that text */
    int64_t line_begin2;      /* ATR_FQ_PAIR_NAMES / ATR_FQ_EMPTY_NAME: line_begin/line_end = header line of the read in */
    int64_t line_end2;        /* file 1, line_begin2/line_end2 = header line of its mate in file 2 */
} atr_fastq_error;

/* The modifiers and filters the `trim` command puts around the adapter stage (commands/trim/__init__.py:422-620),
 * in the default operation order "CGQAW" (trim/cli.py:232-241): cut and quality-trim BEFORE the adapters, N-end
 * trimming after them, then the filters in the order the command adds them. All zero / off by default. */
typedef struct atr_read_ops {
    int32_t cut_front[2];     /* -u N / -U N, N > 0 summed: UnconditionalCutter.front_length of read 1 / read 2 (modifiers.py:565-585) */
    int32_t cut_back[2];      /* -u -N / -U -N summed (<= 0): back_length */
    int32_t quality_front;    /* -q F,B: QualityTrimmer cutoffs (modifiers.py:748-764, _qualtrim.pyx:7-49); 0,0 = off */
    int32_t quality_back;
    int32_t quality_base;     /* --quality-base, 33 */
    int32_t trim_n;           /* --trim-n: NEndTrimmer (modifiers.py:766-784) */
    int32_t minimum_length;   /* -m: TooShortReadFilter, 0 = off (filters.py:118-128) */
    int32_t maximum_length;   /* -M: TooLongReadFilter, < 0 = off (filters.py:130-140) */
    int32_t discard_trimmed;  /* --discard-trimmed: TrimmedFilter (filters.py:176-180) */
    int32_t discard_untrimmed;/* --discard-untrimmed: UntrimmedFilter (filters.py:170-174) */
    int32_t nextseq_trim[2];  /* --nextseq-trim CUTOFF per read, < 0 = off (NextseqQualityTrimmer modifiers.py:732-746,
                                 nextseq_trim_index _qualtrim.pyx:52-84): after the cut, before -q ("CGQAW") */
    int32_t legacy_first;     /* paired-end "legacy mode" (paired == 'first', trim/cli.py:629-645: no option touches read 2):
                                 the filters look at read 1 only (SingleWrapper, filters.py:54-61) */
    int32_t pair_filter_both; /* --pair-filter both: PairedWrapper(min_affected = 2), a pair is discarded only if BOTH reads
                                 meet the filter's criterion (filters.py:63-95; read 2 is only looked at when read 1 does) */
    double  max_n;            /* --max-n: NContentFilter, < 0 = off; < 1 is a proportion (filters.py:142-168) */
} atr_read_ops;

/* Trimmer.trimmed_bases per modifier and read (modifiers.py:45-88), FilterWrapper.filtered per filter (filters.py:18-52;
 * paired-end: the pair filter "any", PairedWrapper with min_affected = 1), and what was written. ADDED to. */
typedef struct atr_read_ops_stats {
    int64_t bp_cut[2], bp_quality[2], bp_n_ends[2], bp_nextseq[2];
    int64_t too_short, too_long, too_many_n, discarded_trimmed, discarded_untrimmed;
    int64_t records_written;  /* reads (pairs) that passed every filter; their bases are bp_out of the enclosing struct */
} atr_read_ops_stats;

/* AdapterCutter(adapters, times, action='trim') (commands/trim/modifiers.py:91-105) */
typedef struct atr_trim_opts {
    int32_t times;            /* >= 1: rounds of best-match-and-trim per read (modifiers.py:141-149) */
    int32_t max_len;          /* the statistics cover removed lengths 0..max_len ... */
    int32_t max_errors;       /* ... and error counts 0..max_errors */
    int32_t final_chunk;      /* 1: `text` ends the file (a partial last record is an error, an unterminated last
                                 line is a line); 0: stop after the last complete record and report `consumed` */
    int64_t chunk_bytes;      /* internal H2D chunk size, 0 = default (64 MiB) */
    /* LinkedAdapter (adapters/__init__.py:637-690, "-a FRONT...BACK"): non-NULL = `set` holds the one anchored front
     * adapter and linked_back the one back adapter; the back adapter is only looked for, in what the front adapter
     * leaves, when the front adapter matched. Statistics: adapter index 0 = front, 1 = back. Needs times == 1 (with
     * further adapters or rounds the reference itself fails: LinkedMatch has no `matches`, modifiers.py:120). */
    const atr_adapterset* linked_back;
    atr_read_ops ops;         /* index 0 of the per-read fields */
} atr_trim_opts;

/* What Adapter.trimmed() accumulates (adapters/__init__.py:413-436) and the report prints. The histogram
 * arrays are caller-allocated and ADDED to (zero them first; calls and shards then merge by themselves, like
 * Summary.merge, commands/multicore.py:389). Index: ((a*(max_len+1) + length)*(max_errors+1) + errors;
 * lengths_front/back of the reference are the row sums. */
typedef struct atr_trim_stats {
    int64_t records, with_adapters, bp_in, bp_out;   /* bp_out = bases written (reads that passed the filters) */
    int64_t overflow;         /* matches outside the histogram extents (not counted in the arrays) */
    int64_t* errors_front;    /* [n_adapters][max_len+1][max_errors+1] */
    int64_t* errors_back;     /* same shape */
    int64_t* adjacent_bases;  /* [n_adapters][5]: A, C, G, T, '' (anything else) */
    atr_read_ops_stats ops;
} atr_trim_stats;

/* Replaces, for single-end FASTQ and the adapter-trimming modifier, the per-record pipeline
 *   FastqReader.__iter__ (io/_seqio.pyx:180-245) -> AdapterCutter.__call__ (modifiers.py:124-187; reads are
 *   upper-cased for matching only, adapters/__init__.py:349) -> Adapter.trimmed (adapters/__init__.py:413-436)
 *   -> FastqFormat.format (io/seqio.py:686-700)
 * by one pass on the GPU: newline index, record framing + validation, 4-bit packing, the adapter-alignment
 * kernels, trimming windows + statistics, formatting. `text`/`out_text` are HOST buffers (pinned for full speed);
 * out_cap >= nbytes + 1 always suffices (trimming never grows a record; an unterminated last line gains its
 * newline, FastqFormat.format io/seqio.py:686-700). The adapter set must have been created with
 * match_to_semantics = 1. *consumed = bytes of `text` that were processed (all of it if final_chunk).
 * On ATR_E_FORMAT *err describes the first malformed line. */
int  atr_trim_fastq_host(atr_ctx* ctx, const atr_adapterset* set, const atr_trim_opts* opts,
                         const uint8_t* text, int64_t nbytes, uint8_t* out_text, int64_t out_cap,
                         int64_t* out_bytes, int64_t* consumed, atr_trim_stats* stats, atr_fastq_error* err);

/* ---- paired-end twin: two FASTQ texts in -> two trimmed FASTQ texts out ("--aligner insert") ---- */
/* InsertAdapterCutter(adapter1, adapter2, action='trim', mismatch_action=None, symmetric, min_insert_overlap)
 * (commands/trim/modifiers.py:359-389) */
typedef struct atr_trim_pe_opts {
    int32_t symmetric;          /* duplicate the one good adapter match onto the other read (:417-437) */
    int32_t min_insert_overlap; /* pairs with a shorter read are left alone (:392-394) */
    int32_t max_len;            /* statistics cover removed lengths 0..max_len and error counts 0..max_errors */
    int32_t max_errors;
    int32_t final_chunk;        /* 1: both texts end their files */
    int32_t times;              /* adapter mode (iset == NULL): AdapterCutter(times) of both reads, >= 1 */
    int32_t mismatch_action;    /* insert mode, --correct-mismatches: 0 off, 1 liberal, 2 conservative, 3 N (ErrorCorrectorMixin,
                                   commands/trim/modifiers.py:201-357) */
    int32_t pad;
    int64_t chunk_bytes;        /* per text; 0 = default (32 MiB) */
    atr_read_ops ops;
} atr_trim_pe_opts;

typedef struct atr_trim_pe_stats {
    int64_t records, insert_matches;          /* pairs; pairs for which match_insert returned a match */
    int64_t with_adapters[2], bp_in[2], bp_out[2];
    int64_t overflow;
    int64_t* errors_back[2];                  /* per read: [n_adapters of that read][max_len+1][max_errors+1], ADDED to */
    int64_t* adjacent_bases[2];               /* per read: [n_adapters][5] */
    int64_t* errors_front[2];                 /* adapter mode only (may be NULL): same shape as errors_back */
    int64_t records_corrected, bp_corrected[2];   /* ErrorCorrectorMixin.summarize (modifiers.py:352-357) */
    atr_read_ops_stats ops;
} atr_trim_pe_stats;

/* Replaces, for two FASTQ files read in lockstep, PairedSequenceReader.__iter__ (io/seqio.py:429-453: pairing and
 * name checks) over two FastqReaders, InsertAdapterCutter.__call__ (modifiers.py:391-453 with mismatch_action None:
 * InsertAligner.match_insert first, adapter{1,2}.match_to as the fallback, the symmetric fix-up), its trim() and the
 * adapters' statistics (:455-496, adapters/__init__.py:424-436) and FastqFormat.format for both outputs.
 * set1 / set2: one 3' (BACK) adapter each, created with match_to_semantics = 1 (the fallback adapters).
 *
 * iset == NULL selects the command's default paired-end mode instead ("--aligner adapter", commands/trim/__init__.py:
 * 457-476: modifiers.add_modifier_pair(AdapterCutter, ...)): read 1 is cut with the adapters of set1, read 2 with those of
 * set2 (any adapter types, `times` rounds; either set may be NULL), independently of each other; pairing checks,
 * modifiers, pair filters and outputs as above.
 * consumed[i]: bytes of text i that were processed. On ATR_E_FORMAT *err names the file and the line. */
int  atr_trim_fastq_pe_host(atr_ctx* ctx, const atr_insertset* iset, const atr_adapterset* set1, const atr_adapterset* set2,
                            const atr_trim_pe_opts* opts, const uint8_t* text1, int64_t nbytes1, const uint8_t* text2,
                            int64_t nbytes2, uint8_t* out1, int64_t out_cap1, uint8_t* out2, int64_t out_cap2,
                            int64_t* out_bytes, int64_t* consumed, atr_trim_pe_stats* stats, atr_fastq_error* err);

/* ---- ... and the merged reads as a third output ("--merge-overlapping [--merged-output FILE]") ---- */
/* MergeOverlapping(min_overlap, error_rate, mismatch_action) as the LAST modifier of the paired-end command
 * (commands/trim/__init__.py:546-552) + MergedReadFilter as its FIRST filter (:576-579, filters.py:108-113). */
typedef struct atr_merge_opts {
    double  min_overlap;        /* --merge-min-overlap (0.9): a fraction of the shorter read if <= 1, else bases (modifiers.py:871, :877-879) */
    double  error_rate;         /* --merge-error-rate (0.2) */
} atr_merge_opts;

typedef struct atr_merge_stats {
    int64_t merged;             /* MergedReadFilter.records_filtered: pairs whose read 1 became the merged read (read 2 dropped) */
    int64_t merged_written;     /* of those, records written to the merged output (0 when there is none: they are discarded) */
    int64_t bp_merged_written;  /* their bases (part of the report's bp_written[0]) */
    int64_t records_corrected;  /* ErrorCorrectorMixin counters of the MergeOverlapping instance (not in the reference's report: */
    int64_t bp_corrected[2];    /* ReadPairModifier.summarize comes first in its MRO) */
} atr_merge_stats;

/* atr_trim_fastq_pe_host with MergeOverlapping behind the other modifiers. Per pair, on what trimming left of the two
 * reads: Aligner(reverse_complement(read 2), error_rate, flags).locate(read 1) (modifiers.py:864-931; flags SEMIGLOBAL, or
 * START_WITHIN_SEQ1 | STOP_WITHIN_SEQ2 for a pair the insert aligner matched, read.insert_overlap :397, :882-890); a pair
 * with matches >= min_overlap is merged into read 1 in one of four ways (:905-922) -- after error correction of the
 * overlap when opts->mismatch_action is set, the alignment has errors and the pair was neither insert-matched nor
 * corrected before (:899-903, :232-233) -- and leaves through MergedReadFilter: written to `out_merged` as a single-end
 * record under read 1's name (FastqFormat, io/seqio.py:686-700), or discarded when out_merged is NULL (no
 * --merged-output: Formatters.format counts it as discarded, writers.py:151-154). Every other pair goes on to the
 * remaining filters and the two paired outputs exactly as in atr_trim_fastq_pe_host. Reads up to 4000 nt.
 * out_bytes / consumed: [0], [1] as above; out_bytes[2] = bytes written to out_merged (out_cap_merged >= nbytes1 +
 * nbytes2 always suffices). mopts == NULL: identical to atr_trim_fastq_pe_host (mstats may then be NULL too).
 * Where the reference raises -- reverse_complement KeyError on a byte outside its table (util/__init__.py:479-482),
 * AtroposError("Invalid alignment while trying to merge read ...") (:923-927) -- the call fails with ATR_E_FORMAT,
 * err->kind = ATR_FQ_INVALID_MATCH. */
int  atr_trim_fastq_pe_merge_host(atr_ctx* ctx, const atr_insertset* iset, const atr_adapterset* set1, const atr_adapterset* set2,
                                  const atr_trim_pe_opts* opts, const atr_merge_opts* mopts,
                                  const uint8_t* text1, int64_t nbytes1, const uint8_t* text2, int64_t nbytes2,
                                  uint8_t* out1, int64_t out_cap1, uint8_t* out2, int64_t out_cap2,
                                  uint8_t* out_merged, int64_t out_cap_merged,
                                  int64_t* out_bytes, int64_t* consumed, atr_trim_pe_stats* stats, atr_merge_stats* mstats,
                                  atr_fastq_error* err);

#ifdef __cplusplus
}
#endif
#endif /* ATROPOS_B200_H */
