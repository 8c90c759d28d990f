// atr_kernels.cuh -- the __global__ kernels (sm_100a). Host launch code lives in atr_api.cu.
#pragma once
#include <cuda_runtime.h>
#include "locate_core.cuh"
#include "qgram_core.cuh"
#include "insert_core.cuh"
#include "adapter_build.hpp"

// ---------------------------------------------------------------------------------------------
// pack: ASCII -> 4-bit codes. One warp per read, one lane per output word (8 bases), so a warp reads
// up to 256 contiguous bytes and writes up to 128 contiguous bytes per pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_word_counts(const int64_t* __restrict__ offsets, int64_t n,
                                                     uint32_t* __restrict__ counts) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) counts[r] = (uint32_t)((offsets[r + 1] - offsets[r] + 7) >> 3);
}

// fixed-length chunk: offsets are an arithmetic progression, generated here instead of crossing PCIe (8 B/read)
__global__ void __launch_bounds__(256) k_make_offsets(int64_t* __restrict__ offsets, int64_t n, int64_t first, int64_t len) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) offsets[i] = first + i * len;
}

// fixed-length packed chunk: word offsets and lengths are arithmetic, generated here instead of crossing PCIe (6 B/read)
__global__ void __launch_bounds__(256) k_make_packed_index(uint32_t* __restrict__ woff, uint16_t* __restrict__ len, int64_t n,
                                                           uint32_t first_word, uint32_t words_per_read, uint16_t len_value) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) woff[i] = first_word + (uint32_t)i * words_per_read;
    if (i < n) len[i] = len_value;
}

// Fast path of the packers: the eight bytes of a full output word are all upper-case A / C / G / T (nearly every word of
// real data). Three aligned 32-bit loads and two byte permutes fetch them whatever the read's alignment; bits 1-2 of
// a base (A 00, C 01, T 10, G 11) become the selector of two more permutes, one producing the 4-bit codes, the other
// the letters those selectors stand for -- equal to the input iff the input was A/C/G/T, which is the exactness test.
// Anything else (N, IUPAC, lower case, other bytes) takes the byte-wise pack_word.
// Reads up to 3 bytes before and after the eight (never before the 4-aligned `ascii` base; the caller checks the end).
// `valid` (1..8) = bases of the word that belong to the read: the bytes behind them are read (they exist) but replaced.
__device__ __forceinline__ bool pack8_acgt(const unsigned char* __restrict__ p, int valid, uint32_t& w) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3);
    const uint32_t x0 = q[0], x1 = q[1], x2 = sh ? q[2] : 0u;
    const unsigned take = 0x3210u + 0x1111u * sh;
    uint32_t g0 = __byte_perm(x0, x1, take), g1 = __byte_perm(x1, x2, take);
    if (valid < 8) {                                                   // the read's last word: 'A' behind the read end
        const unsigned long long keep = (1ull << (8 * valid)) - 1ull;
        g0 = (g0 & (uint32_t)keep) | (0x41414141u & ~(uint32_t)keep);
        g1 = (g1 & (uint32_t)(keep >> 32)) | (0x41414141u & ~(uint32_t)(keep >> 32));
    }
    uint32_t out = 0;
    bool ok = true;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t g = h ? g1 : g0;
        uint32_t t = (g >> 1) & 0x03030303u;
        t |= t >> 4;
        const uint32_t sel = __byte_perm(t, 0u, 0x4420);                 // one selector nibble per base
        ok = ok && __byte_perm(0x47544341u, 0u, sel) == g;             // 'A' 'C' 'T' 'G'
        uint32_t c = __byte_perm(0x04080201u, 0u, sel);                // their codes 1 2 8 4
        c |= c >> 4;
        out |= __byte_perm(c, 0u, 0x4420) << (16 * h);
    }
    w = valid < 8 ? out & ((1u << (4 * valid)) - 1u) : out;
    return ok;
}

// One thread per OUTPUT WORD. A CTA takes ATR_PK_READS consecutive reads at a time, keeps their offsets and word offsets
// in shared memory and walks the words of the group with all 256 threads (the read of a word is a 6-step binary search
// in shared memory), so that every lane has loads in flight and the stores are consecutive words. (A warp per read left
// 13 of 32 lanes idle at 150 nt and had one read's dependent loads in flight per warp: 3.6 ms per 10 M reads, latency bound.)
#define ATR_PK_READS 64
__device__ __forceinline__ int pk_find(const uint32_t* __restrict__ s_woff, int cnt, uint32_t x) {   // largest i with s_woff[i] <= x
    int lo = 0, hi = cnt;
#pragma unroll
    for (int it = 0; it < 7; it++) {                       // cnt <= 64: 7 halvings always suffice
        const int mid = (lo + hi + 1) >> 1;
        if (s_woff[mid] <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_pack(const unsigned char* __restrict__ ascii, const int64_t* __restrict__ offsets,
                                              int64_t base, int64_t n, int fold_case,
                                              const AtrTables* __restrict__ tables, const uint32_t* __restrict__ woff,
                                              uint32_t* __restrict__ codes, uint16_t* __restrict__ len_out) {
    __shared__ unsigned char s_iupac[256];
    __shared__ int64_t s_off[ATR_PK_READS + 1];
    __shared__ uint32_t s_woff[ATR_PK_READS + 1];
    __shared__ int s_esc[ATR_PK_READS];
    const int tid = threadIdx.x;
    s_iupac[tid] = tables->iupac[tid];
    const bool aligned = (reinterpret_cast<uintptr_t>(ascii) & 3) == 0;
    const int64_t total = offsets[n] - base;                           // the fast path may read 3 bytes past its word
    for (int64_t r0 = (int64_t)blockIdx.x * ATR_PK_READS; r0 < n; r0 += (int64_t)gridDim.x * ATR_PK_READS) {
        const int cnt = (int)(n - r0 < ATR_PK_READS ? n - r0 : ATR_PK_READS);
        __syncthreads();                                               // the previous group is done with the tables
        if (tid <= cnt) { s_off[tid] = offsets[r0 + tid] - base; s_woff[tid] = woff[r0 + tid]; }
        if (tid < cnt) s_esc[tid] = 0;
        __syncthreads();
        const uint32_t wbeg = s_woff[0], wcnt = s_woff[cnt] - wbeg;
        for (uint32_t j = tid; j < wcnt; j += 256) {
            const int i = pk_find(s_woff, cnt, wbeg + j);
            const int w = (int)(wbeg + j - s_woff[i]);
            const int64_t off = s_off[i];
            const int len = (int)(s_off[i + 1] - off);
            uint32_t v;
            int esc = 0;
            if (!(aligned && off + 8 * w + 11 < total && pack8_acgt(ascii + off + 8 * w, atr_min(8, len - 8 * w), v)))
                v = atr::pack_word(ascii + off, len, w, fold_case, s_iupac, &esc);
            codes[wbeg + j] = v;
            if (esc) s_esc[i] = 1;                                     // same value from every writer
        }
        __syncthreads();
        if (tid < cnt) len_out[r0 + tid] = (uint16_t)((int)(s_off[tid + 1] - s_off[tid]) | (s_esc[tid] ? ATR_ESC_BIT : 0));
    }
}

// ---------------------------------------------------------------------------------------------
// K1a: one thread per read, DP column in registers (see locate_core.cuh).
// ---------------------------------------------------------------------------------------------
template <bool AND_MODE>
__global__ void __launch_bounds__(128) k_locate_k1a(const __grid_constant__ AdapterK1a ad,
                                                    const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff,
                                                    const uint16_t* __restrict__ len, const uint16_t* __restrict__ win,
                                                    int64_t n_reads, atr_match* __restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const unsigned l = len[r];
    const bool esc = (l & ATR_ESC_BIT) != 0;
    int hi = (int)(l & ATR_LEN_MASK), lo = 0;
    if (win != nullptr) {
        const int wlo = win[2 * r], whi = win[2 * r + 1];
        hi = atr_min(hi, whi);
        lo = atr_min(wlo, hi);
    }
    const int n = hi - lo;
    // routed to the byte-exact general kernel (k_locate_gen, launched right after on the same stream)
    if ((esc && (!AND_MODE || ad.need_find)) || n > ATR_K1A_MAXN) {
        if (ad.mark_routed) {
            atr_match m;
            m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0;
            m.adapter = -1; m.status = ATR_ST_ESCAPED;
            out[r] = m;
        }
        return;
    }
    k1a_read<AND_MODE>(ad, codes + woff[r], lo, n, out + r);
}

// ---------------------------------------------------------------------------------------------
// K1g: general kernel on raw ASCII, DP column in global scratch (coalesced: element i of thread t at
// scratch[i * nthreads + t]). all_reads = 0: only the reads K1a skipped; 1: every read (adapter not K1a-able).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_locate_gen(const __grid_constant__ AdapterGen ad, int k1a_and_mode, int all_reads,
                                                    const AtrTables* __restrict__ tables,
                                                    const unsigned char* __restrict__ ascii, const int64_t* __restrict__ offsets,
                                                    int64_t base, const uint16_t* __restrict__ len,
                                                    const uint16_t* __restrict__ win, int fold_case, int64_t n_reads,
                                                    GCell* __restrict__ scratch, atr_match* __restrict__ out) {
    __shared__ AtrTables s_tb;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_tb.iupac[i] = tables->iupac[i]; s_tb.acgt[i] = tables->acgt[i]; }
    __syncthreads();
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t r = t; r < n_reads; r += nthreads) {
        const int full = (int)(offsets[r + 1] - offsets[r]);
        int hi = full, lo = 0;
        if (win != nullptr) {
            const int wlo = win[2 * r], whi = win[2 * r + 1];
            hi = atr_min(hi, whi);
            lo = atr_min(wlo, hi);
        }
        const int n = hi - lo;
        if (!all_reads) {
            const bool esc = len != nullptr && (len[r] & ATR_ESC_BIT) != 0;
            const bool routed = (esc && (!k1a_and_mode || ad.need_find)) || n > ATR_K1A_MAXN;
            if (!routed) continue;
        }
        gen_read(ad, s_tb, ascii + (offsets[r] - base) + lo, n, fold_case, scratch + t, nthreads, out + r);
    }
}

__global__ void k_fill_none(atr_match* __restrict__ out, int64_t n) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        atr_match m;
        m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0;
        m.adapter = -1; m.status = ATR_ST_NONE;
        out[r] = m;
    }
}

// ---------------------------------------------------------------------------------------------
// K2: InsertAligner.match_insert, one thread per pair. The reverse-complemented mate and read 1
// are staged as packed words in shared memory ([word][thread], conflict-free) because the sliding
// overlap needs dynamically indexed funnel shifts.
// ---------------------------------------------------------------------------------------------
#ifndef ATR_K2_THREADS
#define ATR_K2_THREADS 128
#endif
__global__ void __launch_bounds__(ATR_K2_THREADS) k_insert_packed(
        const __grid_constant__ InsertDev d,
        const uint32_t* __restrict__ codes1, const uint32_t* __restrict__ woff1, const uint16_t* __restrict__ len1,
        const uint32_t* __restrict__ codes2, const uint32_t* __restrict__ woff2, const uint16_t* __restrict__ len2,
        int64_t n_pairs, atr_insert_result* __restrict__ out) {
    __shared__ uint32_t sR2[ATR_K2_MAXW2 * ATR_K2_THREADS];     // rc(read 2) in 2-bit words, [word][thread]
    __shared__ unsigned short s_thr[ATR_K2_MAXLEN + 1];         // floor(j * rate) for every overlap length the packed path sees
    for (int i = threadIdx.x; i <= ATR_K2_MAXLEN; i += ATR_K2_THREADS) s_thr[i] = i <= d.max_len ? d.thr_ins[i] : (unsigned short)0;
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pairs) return;
    const unsigned l1 = len1[r], l2 = len2[r];
    const int n1 = (int)(l1 & ATR_LEN_MASK), n2 = (int)(l2 & ATR_LEN_MASK);
    const int m = n1 < n2 ? n1 : n2;
    atr_insert_result* o = out + r;
    bool routed = ((l1 | l2) & ATR_ESC_BIT) != 0 || m > ATR_K2_MAXLEN || !d.packed_ok;
    PackedPair pp;
    pp.R2 = sR2 + threadIdx.x; pp.stride = ATR_K2_THREADS; pp.thr = s_thr;
    if (!routed) routed = packed_pair_setup(pp, codes1 + woff1[r], codes2 + woff2[r], m, (n1 + 7) >> 3) == 0;
    if (routed) {                      // the byte-exact kernel (k_insert_bytes) picks these up
        atr_insert_result e;
        im_clear(e.insert); im_clear(e.match1); im_clear(e.match2);
        e.insert.status = ATR_ST_ESCAPED;
        *o = e;
        return;
    }
    Cand cand[ATR_MAX_CAND];
    insert_pair(d, pp, true, m, n1, n2, cand, o);
}

__global__ void __launch_bounds__(128) k_insert_bytes(
        const __grid_constant__ InsertDev d,
        const unsigned char* __restrict__ ascii1, const int64_t* __restrict__ off1, int64_t base1,
        const unsigned char* __restrict__ ascii2, const int64_t* __restrict__ off2, int64_t base2,
        int64_t n_pairs, atr_insert_result* __restrict__ out) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_pairs; r += nthreads) {
        if (out[r].insert.status != ATR_ST_ESCAPED) continue;
        const int n1 = (int)(off1[r + 1] - off1[r]), n2 = (int)(off2[r + 1] - off2[r]);
        const int m = n1 < n2 ? n1 : n2;
        BytePair bp;
        bp.s1 = ascii1 + (off1[r] - base1); bp.s2 = ascii2 + (off2[r] - base2);
        bp.comp = d.comp; bp.ov_tab = d.ov_tab; bp.m = m;
        bool keyerr = false;
        for (int p = 0; p < m; p++) keyerr = keyerr || d.comp[bp.s2[p]] == 0;
        if (keyerr) {
            atr_insert_result e;
            im_clear(e.insert); im_clear(e.match1); im_clear(e.match2);
            e.insert.status = ATR_ST_KEYERROR;
            out[r] = e;
            continue;
        }
        Cand cand[ATR_MAX_CAND];
        insert_pair(d, bp, false, m, n1, n2, cand, out + r);
    }
}

// single-call MultiAligner.locate (any flags): one thread
__global__ void k_multi_locate(const unsigned char* __restrict__ ref, int m, const unsigned char* __restrict__ query, int n,
                               int k, const unsigned short* __restrict__ thr, int flags, int min_overlap, int max_matches,
                               GCellM* col, int* out6, int* n_out) {
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *n_out = gen_multi_locate(ref, m, query, n, k, thr, flags, min_overlap, max_matches, col, out6);
}

// single-call compare_prefixes (_align.pyx:501-544): one block, lanes stride the common prefix
__global__ void __launch_bounds__(256) k_compare_prefixes(const unsigned char* __restrict__ ref, const unsigned char* __restrict__ query,
                                                          int length, int mode /*0 ascii, 1 and*/, const unsigned char* __restrict__ tr,
                                                          const unsigned char* __restrict__ tq, int* matches_out) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < length; i += blockDim.x) {
        if (mode == 0) c += ref[i] == query[i];
        else c += (tr[ref[i]] & tq[query[i]]) != 0;
    }
    atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) *matches_out = s_cnt;
}

// ---------------------------------------------------------------------------------------------
// The fast path (AdapterK1a.fused_ok: unit indel cost, free start in the read, adapter <= 64 nt) is a funnel of
// kernels on one stream; survivors are compacted into lists in global memory so that every stage runs dense:
//   k_filter_sa  (adapters whose k+1 pieces are >= 6 rows, BACK/SUFFIX style) one CTA = one tile of 256 reads,
//                staged in shared memory by a single TMA bulk copy (cp.async.bulk completing on an mbarrier):
//                Shift-And over verbatim pieces, the str.find shortcut, an exact 32-bit Myers over the read tail
//                for the reads that can have a partial match there (compacted inside the CTA);
//   k_filter     (every other eligible adapter, incl. FRONT/ANYWHERE) same tiling; exact 32/64-bit Myers/Hyyro
//                bit-vector DP over the whole read;
//   k_refine     dense over the reads whose piece hits do not pin a narrow band: exact Myers on the columns around
//                the hits;
//   k_band       dense over the "narrow" survivors: tie-broken 3-field DP on 16 diagonals (k1d_band);
//   k_wide       dense over the rest: the register-column DP restricted to a column window (k1a_locate); runs on a
//                side stream underneath k_band.
// ---------------------------------------------------------------------------------------------
#define ATR_K1F_THREADS 256
#define ATR_K1F_TILE_WORDS 6144      // 24 KB: 256 reads of up to 192 nt

struct Survivor {                    // 8 bytes
    uint32_t read;
    short a, b;                      // narrow: a = dlo ; wide: a = c0, b = c1
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

// warp-aggregated append: one atomicAdd per warp per list
__device__ __forceinline__ void list_append(bool want, Survivor sv, Survivor* __restrict__ list, int* __restrict__ counter) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) list[base + __popc(mask & ((1u << lane) - 1u))] = sv;
}

// the same from the end of a list downwards: end[-1], end[-2], ... (the 8-diagonal survivors share the narrow list's
// storage with the 16-diagonal ones, which fill it from the front)
__device__ __forceinline__ void list_append_back(bool want, Survivor sv, Survivor* __restrict__ end, int* __restrict__ counter) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) end[-1 - (base + __popc(mask & ((1u << lane) - 1u)))] = sv;
}

__device__ __forceinline__ void read_extent(const uint16_t* __restrict__ len, const uint16_t* __restrict__ win, int64_t r,
                                            int& lo, int& n, bool& esc) {
    const unsigned l = len[r];
    esc = (l & ATR_ESC_BIT) != 0;
    int hi = (int)(l & ATR_LEN_MASK);
    lo = 0;
    if (win != nullptr) {
        const int wlo = win[2 * r], whi = win[2 * r + 1];
        hi = atr_min(hi, whi);
        lo = atr_min(wlo, hi);
    }
    n = hi - lo;
}

template <class WORD, bool AND_MODE>
__global__ void __launch_bounds__(ATR_K1F_THREADS) k_filter(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, int64_t n_reads, atr_match* __restrict__ out,
        Survivor* __restrict__ narrow, Survivor* __restrict__ wide, int* __restrict__ counters) {
    __shared__ __align__(128) uint32_t s_tile[ATR_K1F_TILE_WORDS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ WORD s_peq[16];

    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * ATR_K1F_THREADS;
    const int cnt = (int)(n_reads - t0 < ATR_K1F_THREADS ? n_reads - t0 : ATR_K1F_THREADS);
    const uint32_t w_begin = woff[t0], w_end = woff[t0 + cnt];
    const uint32_t a_begin = w_begin & ~3u;                         // 16-byte aligned start of the span
    const uint32_t span = ((w_end - a_begin) + 3u) & ~3u;           // words, multiple of 4
    // TMA needs 16-byte aligned addresses and sizes; the last tile may not read past the end of `codes`
    const bool last_tile = (t0 + cnt == n_reads);
    const bool fits = span <= ATR_K1F_TILE_WORDS;
    const bool use_tma = fits && !last_tile && span > 0 && ((reinterpret_cast<uintptr_t>(codes) & 15) == 0);
    if (tid < 16) {
        const int WB = (int)(8 * sizeof(WORD)), sh = WB - ad.m;     // left-aligned pattern, virtual rows all ones
        s_peq[tid] = (WORD)(((WORD)ad.peq[tid] << sh) | (sh ? (((WORD)1 << sh) - 1) : 0));
    }
    if (tid == 0 && use_tma) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (use_tma) {
        if (tid == 0) tma_load_1d(s_tile, codes + a_begin, span * 4u, &s_bar);
    } else if (fits) {                                              // edge tile: coalesced cooperative copy
        for (uint32_t w = tid; w < w_end - a_begin; w += ATR_K1F_THREADS) s_tile[w] = codes[a_begin + w];
    }
    // per-read bookkeeping while the copy is in flight
    const int64_t r = t0 + tid;
    const bool mine = tid < cnt;
    bool routed = false, esc = false;
    int lo = 0, n = 0;
    uint32_t wr = a_begin;
    if (mine) {
        read_extent(len, win, r, lo, n, esc);
        wr = woff[r];
        routed = (esc && !AND_MODE) || n > ATR_K1A_MAXN;            // the byte-exact general kernel picks these up
        if (routed && ad.mark_routed) {
            atr_match m;
            m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0;
            m.adapter = -1; m.status = ATR_ST_ESCAPED;
            out[r] = m;
        }
    }
    if (use_tma) mbar_wait(&s_bar, 0);
    else __syncthreads();
    const uint32_t* rd = codes + wr;                       // generic pointer: shared tile or global
    if (fits) rd = s_tile + (wr - a_begin);
    bool to_narrow = false, to_wide = false;
    Survivor sv;
    sv.read = (uint32_t)r; sv.a = 0; sv.b = 0;
    if (mine && !routed) {
        FilterHit hit;
        if (myers_filter<WORD>(ad, s_peq, rd, lo, n, hit)) {
            if (ad.band_ok && hit.width <= ATR_K1D_W) { to_narrow = true; sv.a = (short)hit.dlo; }
            else { to_wide = true; sv.a = (short)hit.c0; sv.b = (short)hit.c1; }
        } else {
            Best b;
            b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
            finalize(ad, b, n, out + r);                            // "no match" (respects the panel reduction)
        }
    }
    list_append(to_narrow, sv, narrow, counters + 0);
    list_append(to_wide, sv, wide, counters + 1);
}

// ---------------------------------------------------------------------------------------------
// k_filter_front: first stage for unanchored 5' adapters (AdapterK1a.sa_front; locate_core.cuh: front_filter). k_filter's
// frame (one CTA = 256 reads, TMA tile) with the whole-read Myers pass (~14 instructions per column) replaced by the
// Shift-And automaton over the pieces (~7.5 per column) + an exact Myers pass over the first m + k columns only.
// ---------------------------------------------------------------------------------------------
template <bool AND_MODE>
__global__ void __launch_bounds__(ATR_K1F_THREADS) k_filter_front(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, int64_t n_reads, atr_match* __restrict__ out,
        Survivor* __restrict__ narrow, Survivor* __restrict__ wide, int* __restrict__ counters) {
    __shared__ __align__(128) uint32_t s_tile[ATR_K1F_TILE_WORDS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned s_peq[16], s_sa_peq[16];
    __shared__ unsigned long long s_sa_pair[256];

    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * ATR_K1F_THREADS;
    const int cnt = (int)(n_reads - t0 < ATR_K1F_THREADS ? n_reads - t0 : ATR_K1F_THREADS);
    const uint32_t w_begin = woff[t0], w_end = woff[t0 + cnt];
    const uint32_t a_begin = w_begin & ~3u;
    const uint32_t span = ((w_end - a_begin) + 3u) & ~3u;
    const bool last_tile = (t0 + cnt == n_reads);
    const bool fits = span <= ATR_K1F_TILE_WORDS;
    const bool use_tma = fits && !last_tile && span > 0 && ((reinterpret_cast<uintptr_t>(codes) & 15) == 0);
    const unsigned long long mk = ad.sa_rows >= 32 ? 0xFFFFFFFFull : ((1ull << ad.sa_rows) - 1);
    if (tid < 16) {
        const int sh = 32 - ad.m;
        s_peq[tid] = ((unsigned)ad.peq[tid] << sh) | (sh ? ((1u << sh) - 1u) : 0u);
        s_sa_peq[tid] = (unsigned)(ad.peq[tid] & mk);
    }
    s_sa_pair[tid] = (ad.peq[tid & 15] & mk) | ((ad.peq[tid >> 4] & mk) << 32);
    if (tid == 0 && use_tma) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (use_tma) {
        if (tid == 0) tma_load_1d(s_tile, codes + a_begin, span * 4u, &s_bar);
    } else if (fits) {
        for (uint32_t w = tid; w < w_end - a_begin; w += ATR_K1F_THREADS) s_tile[w] = codes[a_begin + w];
    }
    const int64_t r = t0 + tid;
    const bool mine = tid < cnt;
    bool routed = false, esc = false;
    int lo = 0, n = 0;
    uint32_t wr = a_begin;
    if (mine) {
        read_extent(len, win, r, lo, n, esc);
        wr = woff[r];
        routed = (esc && !AND_MODE) || n > ATR_K1A_MAXN;
        if (routed && ad.mark_routed) {
            atr_match m;
            m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0;
            m.adapter = -1; m.status = ATR_ST_ESCAPED;
            out[r] = m;
        }
    }
    if (use_tma) mbar_wait(&s_bar, 0);
    else __syncthreads();
    const uint32_t* rd = codes + wr;
    if (fits) rd = s_tile + (wr - a_begin);
    bool to_narrow = false, to_wide = false, narrow8 = false;
    Survivor sv;
    sv.read = (uint32_t)r; sv.a = 0; sv.b = 0;
    if (mine && !routed) {
        FilterHit hit;
        if (front_filter(ad, s_sa_peq, s_sa_pair, s_peq, rd, lo, n, hit)) {
            if (ad.band_ok && hit.width <= ATR_K1D_W) { to_narrow = true; sv.a = (short)hit.dlo; narrow8 = ad.split8 && hit.width <= 8; }
            else { to_wide = true; sv.a = (short)hit.c0; sv.b = (short)hit.c1; }
        } else {
            Best b;
            b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
            finalize(ad, b, n, out + r);
        }
    }
    list_append(to_narrow && !narrow8, sv, narrow, counters + 0);
    list_append_back(narrow8, sv, wide, counters + 3);
    list_append(to_wide, sv, wide, counters + 1);
}

// ---------------------------------------------------------------------------------------------
// k_filter_sa: the cheaper first stage used when the adapter's pieces are long enough (AdapterK1a.sa_ok):
// Shift-And over k+1 verbatim pieces of the first <= 32 adapter rows (~7 instructions per column) plus an exact
// 32-bit Myers over the last columns for the partial matches at the read end. Reads with a piece hit go to
// k_refine (exact Myers on the few columns around the hits), which feeds the same narrow / wide lists.
// ---------------------------------------------------------------------------------------------
// The tail pass is a separate (non-inlined) device function: it is executed by few, compacted warps and must
// not inflate the register allocation of the Shift-And scan that every thread runs.
template <class WORD>
__device__ __noinline__ int sa_tail_packed_w(const AdapterK1a* ad, const WORD* tail_peq, const uint32_t* rd, int lo, int n) {
    int imin, imax;
    sa_tail_w<WORD>(*ad, tail_peq, rd, lo, n, imin, imax);
    return (imin << 16) | imax;
}
__device__ __forceinline__ int sa_tail_packed(const AdapterK1a* ad, const unsigned* tail_peq, const uint32_t* rd, int lo, int n) {
    return sa_tail_packed_w<unsigned>(ad, tail_peq, rd, lo, n);
}

template <bool AND_MODE>
__global__ void __launch_bounds__(ATR_K1F_THREADS) k_filter_sa(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, int64_t n_reads, atr_match* __restrict__ out,
        Survivor* __restrict__ narrow, Survivor* __restrict__ wide, Survivor* __restrict__ refine, int* __restrict__ counters) {
    __shared__ __align__(128) uint32_t s_tile[ATR_K1F_TILE_WORDS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned s_sa_peq[16], s_tail_peq[16];
    __shared__ unsigned long long s_sa_pair[256];      // Peq of two bases per byte of the packed read
    __shared__ int s_im[ATR_K1F_THREADS];
    __shared__ unsigned short s_tail_list[ATR_K1F_THREADS];
    __shared__ int s_tail_count;

    const int tid = threadIdx.x;
    const int64_t t0 = (int64_t)blockIdx.x * ATR_K1F_THREADS;
    const int cnt = (int)(n_reads - t0 < ATR_K1F_THREADS ? n_reads - t0 : ATR_K1F_THREADS);
    const uint32_t w_begin = woff[t0], w_end = woff[t0 + cnt];
    const uint32_t a_begin = w_begin & ~3u;
    const uint32_t span = ((w_end - a_begin) + 3u) & ~3u;
    const bool last_tile = (t0 + cnt == n_reads);
    const bool fits = span <= ATR_K1F_TILE_WORDS;
    const bool use_tma = fits && !last_tile && span > 0 && ((reinterpret_cast<uintptr_t>(codes) & 15) == 0);
    if (tid < 16) {
        const int mp = ad.sa_rows, sh32 = 32 - mp;
        const unsigned low = (unsigned)(ad.peq[tid] & (mp >= 32 ? 0xFFFFFFFFull : ((1ull << mp) - 1)));
        s_sa_peq[tid] = low;
        s_tail_peq[tid] = sh32 ? ((low << sh32) | ((1u << sh32) - 1u)) : low;
    }
    if (tid == 0) s_tail_count = 0;
    {
        const int mp = ad.sa_rows;
        const unsigned long long mk = mp >= 32 ? 0xFFFFFFFFull : ((1ull << mp) - 1);
        s_sa_pair[tid] = (ad.peq[tid & 15] & mk) | ((ad.peq[tid >> 4] & mk) << 32);
    }
    if (tid == 0 && use_tma) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (use_tma) {
        if (tid == 0) tma_load_1d(s_tile, codes + a_begin, span * 4u, &s_bar);
    } else if (fits) {
        for (uint32_t w = tid; w < w_end - a_begin; w += ATR_K1F_THREADS) s_tile[w] = codes[a_begin + w];
    }
    const int64_t r = t0 + tid;
    const bool mine = tid < cnt;
    bool routed = false, esc = false;
    int lo = 0, n = 0;
    uint32_t wr = a_begin;
    if (mine) {
        read_extent(len, win, r, lo, n, esc);
        wr = woff[r];
        routed = (esc && !AND_MODE) || n > ATR_K1A_MAXN;
        if (routed && ad.mark_routed) {
            atr_match m;
            m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0;
            m.adapter = -1; m.status = ATR_ST_ESCAPED;
            out[r] = m;
        }
    }
    if (use_tma) mbar_wait(&s_bar, 0);
    else __syncthreads();
    const uint32_t* rd = codes + wr;                       // generic pointer: shared tile or global
    if (fits) rd = s_tile + (wr - a_begin);
    bool to_narrow = false, to_wide = false, to_refine = false, narrow8 = false;
    Survivor sv;
    sv.read = (uint32_t)r; sv.a = 0; sv.b = 0;
    // ---- phase A (every thread): Shift-And scan, exact-occurrence shortcut, does this read need the tail pass? ----
    int hmin = 0x7fffffff, hmax = -0x7fffffff;
    bool exact = false, need_tail = false;
    if (mine && !routed) {
        unsigned st_final;
        sa_scan(ad, s_sa_peq, s_sa_pair, rd, lo, n, hmin, hmax, st_final);
        exact = sa_exact(ad, rd, lo, n, hmin, hmax);
        if (exact) {                                                // verbatim occurrence: result known (str.find shortcut)
            Best b;
            b.matches = ad.m; b.cost = 0; b.origin = hmin; b.ref_stop = ad.m; b.q_stop = hmin + ad.m;
            finalize(ad, b, n, out + r);
        } else {
            need_tail = sa_need_tail(ad, n, hmax, st_final);
        }
    }
    // ---- phase B: the exact 32-bit Myers over the read tail, only for the reads that can have a partial match at
    // the end (~20 %), compacted so that the warps running it are full ----
    s_im[tid] = 0;
    if (need_tail) s_tail_list[atomicAdd(&s_tail_count, 1)] = (unsigned short)tid;
    __syncthreads();
    for (int e = tid; e < s_tail_count; e += ATR_K1F_THREADS) {
        const int t2 = s_tail_list[e];
        int lo2, n2; bool esc2;
        read_extent(len, win, t0 + t2, lo2, n2, esc2);
        const uint32_t wr2 = woff[t0 + t2];
        const uint32_t* rd2 = codes + wr2;
        if (fits) rd2 = s_tile + (wr2 - a_begin);
        s_im[t2] = sa_tail_packed(&ad, s_tail_peq, rd2, lo2, n2);
    }
    __syncthreads();
    // ---- phase C (every thread): classify ----
    if (mine && !routed && !exact) {
        const int im = s_im[tid];
        SaResult sr;
        sa_classify(ad, lo, n, hmin, hmax, im >> 16, im & 0xFFFF, sr);
        if (sr.cls == 0) {
            Best b;
            b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
            finalize(ad, b, n, out + r);
        } else if (sr.cls == 1) {
            if (ad.band_ok && sr.width <= ATR_K1D_W) { to_narrow = true; sv.a = (short)sr.dlo; narrow8 = ad.split8 && sr.width <= 8; }
            else { to_wide = true; sv.a = (short)sr.c0; sv.b = (short)sr.c1; }
        } else if (ad.band_ok && sr.width <= ATR_K1D_W) {
            to_narrow = true; sv.a = (short)sr.dlo;                 // band known from the hits: no exact pass needed
            narrow8 = ad.split8 && sr.width <= 8;
        } else {
            to_refine = true; sv.a = (short)sr.c0; sv.b = (short)sr.c1;
        }
    }
    list_append(to_narrow && !narrow8, sv, narrow, counters + 0);
    list_append_back(narrow8, sv, wide, counters + 3);             // `wide` starts where the narrow list's storage ends
    list_append(to_wide, sv, wide, counters + 1);
    list_append(to_refine, sv, refine, counters + 2);
}

// ---------------------------------------------------------------------------------------------
// k_filter_qg: the first stage for adapters with a q-gram form (AdapterK1a.qg_ok; qgram_core.cuh) -- the headline
// configuration. Same outputs as k_filter_sa (finished reads, narrow / 8-diagonal / wide / refine survivor lists) for a
// fraction of the instructions: every `step`-th read position costs one hashed byte-table lookup instead of every
// column an automaton step, and everything that only some reads need runs DENSE over a queue in shared memory instead
// of divergently inside the scan (the first build verified hits inline: 3.4 of 32 lanes active, 37 % of the kernel's
// instructions).
//   * persistent CTAs (a few per SM) walk the tiles of 256 reads; the 8 KB lookup table and the Peq tables are built
//     in shared memory once per CTA;
//   * a tile's packed reads are one contiguous span of `codes`, fetched by one TMA bulk copy (cp.async.bulk +
//     mbarrier) into the CTA's tile buffer; the other CTAs of the SM compute meanwhile;
//   * A1 scan (every thread, uniform, no branches): the 8 lookups of a group of 24 (16) columns leave one word of
//     pattern indices in shared memory;
//   * A2 verification, one thread per read WITH hits (a compacted list, about half of the tile): compare the whole
//     piece of every hit, then the verbatim-occurrence shortcut where all hits lie on one diagonal;
//   * A3 per read: verbatim-occurrence shortcut, need-tail gate; reads that need the exact tail pass go to a queue that
//     lives ACROSS tiles, all others are classified and stored / appended right away;
//   * whenever 256 tail candidates have accumulated: the exact 32-bit Myers over the read tail with full warps
//     (the reads come back from L2), then their classification.
// ---------------------------------------------------------------------------------------------
#ifndef ATR_QG_THREADS
#define ATR_QG_THREADS 256
#define ATR_QG_TILE_WORDS 4992       // 19.5 KB: 256 reads of up to 156 nt; 5 CTAs fit an SM
#define ATR_QG_MINCTAS 5
#endif                               // (longer reads: the tile is read from global memory)
#define ATR_QG_PAD 8                 // a group of lookups reads up to 3 words past the read's last word
#define ATR_QG_NONE_LO 0x7fffffff

struct QgTailItem { uint32_t read; short hmin, hmax; };

// warp-aggregated push into the CTA's tail queue (one shared-memory atomic per warp instead of one per thread: the
// per-thread form serialised, 4.7 active lanes per atomic in the profile). Called by every thread of the warp.
__device__ __forceinline__ void qg_tail_push(bool want, const QgTailItem& q, QgTailItem* __restrict__ s_tq, int* __restrict__ s_tq_count) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(s_tq_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) s_tq[base + __popc(m & ((1u << lane) - 1u))] = q;
}

// classification + record store / list appends of one read. Called by EVERY thread of the CTA (warp ballots inside);
// `active`: this thread holds a read. im = (imin << 16) | imax of the tail pass, 0 if it did not run.
__device__ __forceinline__ void qg_finish(const AdapterK1a& ad, bool active, uint32_t read, int lo, int n, int hmin, int hmax,
                                          bool exact, int im, atr_match* __restrict__ out, Survivor* __restrict__ narrow,
                                          Survivor* __restrict__ wide, Survivor* __restrict__ refine, int* __restrict__ counters) {
    bool to_narrow = false, to_wide = false, to_refine = false, narrow8 = false, finished = false;
    Survivor sv;
    sv.read = read; sv.a = 0; sv.b = 0;
    Best b;
    b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;           // "no match"
    if (active) {
        if (exact) {                                                    // verbatim occurrence: the str.find shortcut
            b.matches = ad.m; b.cost = 0; b.origin = hmin; b.q_stop = hmin + ad.m;
            finished = true;
        } else {
            SaResult sr;
            sa_classify(ad, lo, n, hmin, hmax, im >> 16, im & 0xFFFF, sr);
            if (sr.cls == 0) finished = true;
            else if (sr.cls == 1) {
                if (ad.band_ok && sr.width <= ATR_K1D_W) { to_narrow = true; sv.a = (short)sr.dlo; narrow8 = ad.split8 && sr.width <= 8; }
                else { to_wide = true; sv.a = (short)sr.c0; sv.b = (short)sr.c1; }
            } else if (ad.band_ok && sr.width <= ATR_K1D_W) {
                to_narrow = true; sv.a = (short)sr.dlo;                 // band known from the hits: no exact pass needed
                narrow8 = ad.split8 && sr.width <= 8;
            } else {
                to_refine = true; sv.a = (short)sr.c0; sv.b = (short)sr.c1;
            }
        }
    }
    if (finished) finalize(ad, b, n, out + read);
    list_append(to_narrow && !narrow8, sv, narrow, counters + 0);
    list_append_back(narrow8, sv, wide, counters + 3);                 // `wide` starts where the narrow list's storage ends
    list_append(to_wide, sv, wide, counters + 1);
    list_append(to_refine, sv, refine, counters + 2);
}

// groups per chunk: 192 columns at step 3, 160 at step 2 (one chunk for reads up to that length; 48 KB of static shared memory)
template <int S> struct QgChunk { static const int NG = S == 3 ? 8 : 10; };

template <int S, class WORD>
__global__ void __launch_bounds__(ATR_QG_THREADS, ATR_QG_MINCTAS) k_filter_qg(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, int64_t n_reads, atr_match* __restrict__ out,
        Survivor* __restrict__ narrow, Survivor* __restrict__ wide, Survivor* __restrict__ refine, int* __restrict__ counters) {
    constexpr int NG = QgChunk<S>::NG;
    constexpr int GCOLS = S == 3 ? 24 : 16;
    __shared__ __align__(128) uint32_t s_tile[ATR_QG_TILE_WORDS + ATR_QG_PAD];
    __shared__ __align__(16) unsigned char s_qtab[1 << ATR_QG_BITS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_acc[NG * ATR_QG_THREADS];          // [group of the chunk][thread]: the group's 8 lookups, 4 bits each
    __shared__ QgTailItem s_tq[2 * ATR_QG_THREADS];
    __shared__ uint2 s_meta[ATR_QG_THREADS];                 // x: first word of the read relative to the tile, y: lo | n << 16
    __shared__ int s_hmin[ATR_QG_THREADS], s_hmax[ATR_QG_THREADS];
    __shared__ unsigned short s_gm[ATR_QG_THREADS];          // groups of the chunk with hits
    __shared__ WORD s_sa_peq[16], s_tail_peq[16];
    __shared__ unsigned char s_hl[ATR_QG_THREADS];           // reads of the tile with hits
    __shared__ int s_tq_count, s_hl_count;

    const int tid = threadIdx.x;
    // ---- once per CTA: tables, zeroed tile buffer, barrier ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(ad.qg_tab);
        uint4* dst = reinterpret_cast<uint4*>(s_qtab);
        for (int i = tid; i < (1 << ATR_QG_BITS) / 16; i += ATR_QG_THREADS) dst[i] = src[i];
        for (int i = tid; i < ATR_QG_TILE_WORDS + ATR_QG_PAD; i += ATR_QG_THREADS) s_tile[i] = 0u;
    }
    if (tid < 16) qg_peq_tables<WORD>(ad, tid, s_sa_peq[tid], s_tail_peq[tid]);
    if (tid == 0) {
        s_tq_count = 0; s_hl_count = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    uint32_t parity = 0;
    const int64_t n_tiles = (n_reads + ATR_QG_THREADS - 1) / ATR_QG_THREADS;
    const bool aligned = (reinterpret_cast<uintptr_t>(codes) & 15) == 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * ATR_QG_THREADS;
        const int cnt = (int)(n_reads - t0 < ATR_QG_THREADS ? n_reads - t0 : ATR_QG_THREADS);
        const uint32_t w_begin = woff[t0], w_end = woff[t0 + cnt];
        const uint32_t a_begin = w_begin & ~3u;                        // TMA: 16-byte aligned address and size
        const uint32_t span = ((w_end - a_begin) + 3u) & ~3u;
        const bool last_tile = (t0 + cnt == n_reads);                  // may not read past the end of `codes`
        const bool fits = span <= ATR_QG_TILE_WORDS;
        const bool use_tma = fits && !last_tile && span > 0 && aligned;
        __syncthreads();                       // everybody is done with the previous tile (and with the set-up above)
        if (tid == 0 && use_tma) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic reads of the tile before the async write
            tma_load_1d(s_tile, codes + a_begin, span * 4u, &s_bar);
        }
        if (fits && !use_tma)
            for (uint32_t w = tid; w < w_end - a_begin; w += ATR_QG_THREADS) s_tile[w] = codes[a_begin + w];
        // per-read bookkeeping while the copy is in flight
        const int64_t r = t0 + tid;
        const bool mine = tid < cnt;
        bool routed = false, esc = false;
        int lo = 0, n = 0, nw = 0;
        uint32_t wr = a_begin;
        if (mine) {
            nw = (int)(((len[r] & ATR_LEN_MASK) + 7u) >> 3);
            read_extent(len, win, r, lo, n, esc);
            wr = woff[r];
            routed = esc || n > ATR_K1A_MAXN;                          // ASCII compare mode: escaped reads take k_locate_gen
            if (routed && ad.mark_routed) {
                atr_match m;
                m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0;
                m.adapter = -1; m.status = ATR_ST_ESCAPED;
                store_match(out + r, m);
            }
        }
        const bool live = mine && !routed;
        s_meta[tid] = make_uint2(wr - a_begin, (unsigned)lo | ((unsigned)n << 16));
        s_hmin[tid] = 0x7fffffff; s_hmax[tid] = -0x7fffffff;
        __syncthreads();                       // slots and the cooperative copy are visible
        if (use_tma) { mbar_wait(&s_bar, parity); parity ^= 1u; }
        const uint32_t* rd = codes + wr;       // generic pointer: shared tile or global
        int wlimit = nw;
        if (fits) { rd = s_tile + (wr - a_begin); wlimit = (int)(ATR_QG_TILE_WORDS + ATR_QG_PAD - (wr - a_begin)); }
        int g0 = 0, g1 = 0;
        if (live) qg_group_range<S>(lo, n, g0, g1);
        int gb = 0;
        bool more;
        do {
            // ---- A1 (every thread, uniform, branch-free): 8 lookups per group -> one word of pattern indices ----
            unsigned gm = 0;
#pragma unroll
            for (int j = 0; j < NG; j++) {
                const int g = gb + j;
                uint32_t acc = 0;
                if (g >= g0 && g < g1) acc = qg_group<S>(s_qtab, ad.qg_mul, rd, g, wlimit);
                s_acc[j * ATR_QG_THREADS + tid] = acc;
                gm |= acc ? (1u << j) : 0u;
            }
            s_gm[tid] = (unsigned short)gm;
            more = gb + NG < g1;                                        // this read continues in the next chunk
            const bool last_chunk = !__syncthreads_or(more);            // (also: the chunk's words are visible)
            // reads with hits (in this or an earlier chunk) -> the chunk's list, one atomic per warp. In the last chunk
            // whoever takes a read from the list also finishes it; a read that never had a hit is finished by its own thread.
            const bool listed = live && (gm != 0u || (last_chunk && gb > 0 && s_hmax[tid] != -0x7fffffff));
            {
                const unsigned m = __ballot_sync(0xffffffffu, listed);
                if (m) {
                    const int lane = tid & 31, leader = __ffs(m) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(&s_hl_count, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (listed) s_hl[base + __popc(m & ((1u << lane) - 1u))] = (unsigned char)tid;
                }
            }
            {
                bool push = false;
                if (last_chunk && live && !listed) {                    // no piece anywhere: a partial match at the read end?
                    push = qg_need_tail<WORD>(ad, s_sa_peq, rd, lo, n, -0x7fffffff);
                    if (!push) {
                        Best b;
                        b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
                        finalize(ad, b, n, out + r);
                    }
                }
                QgTailItem q;
                q.read = (uint32_t)r; q.hmin = 32767; q.hmax = -32768;
                qg_tail_push(push, q, s_tq, &s_tq_count);
            }
            __syncthreads();
            // ---- A2: one thread per read WITH hits (dense): verify every hit of the chunk; in the last chunk the verbatim-
            // occurrence shortcut, the need-tail gate and the read's classification follow at once ----
            const int n_hl = s_hl_count;
            {
                const bool act = tid < n_hl;
                int t = 0, lo_t = 0, n_t = 0, vmin = 0x7fffffff, vmax = -0x7fffffff;
                bool exact = false, need_tail = false;
                if (act) {
                    t = s_hl[tid];
                    const uint2 mt = s_meta[t];
                    const uint32_t* rdt = fits ? s_tile + mt.x : codes + (a_begin + mt.x);
                    lo_t = (int)(mt.y & 0xFFFFu); n_t = (int)(mt.y >> 16);
                    vmin = s_hmin[t]; vmax = s_hmax[t];
                    // one flat loop over the read's hits: every trip of the warp verifies one hit in every lane that still
                    // has one (a loop per group ran at 6 of 32 lanes: the lanes' hits sit in different groups)
                    unsigned gmt = s_gm[t];
                    uint32_t acc = 0;
                    int j = 0;
                    for (;;) {
                        if (acc == 0u) {
                            if (gmt == 0u) break;
                            j = atr_ctz(gmt);
                            gmt &= gmt - 1u;
                            acc = s_acc[j * ATR_QG_THREADS + t];
                        }
                        const int i = atr_ctz(acc) >> 2;
                        const int id = (int)((acc >> (4 * i)) & 15u);
                        acc &= ~(15u << (4 * i));
                        const int c = GCOLS * (gb + j) + S * i;
                        if (id == 15) { for (int p2 = 1; p2 <= ad.qg_npat; p2++) qg_verify_pattern(ad, rdt, lo_t, n_t, p2, c, vmin, vmax); }
                        else qg_verify_pattern(ad, rdt, lo_t, n_t, id, c, vmin, vmax);
                    }
                    if (!last_chunk) { s_hmin[t] = vmin; s_hmax[t] = vmax; }
                    else {
                        exact = sa_exact(ad, rdt, lo_t, n_t, vmin, vmax);
                        if (!exact) need_tail = qg_need_tail<WORD>(ad, s_sa_peq, rdt, lo_t, n_t, vmax);
                    }
                }
                {
                    QgTailItem q;
                    q.read = (uint32_t)(t0 + t);
                    q.hmin = (short)(vmax == -0x7fffffff ? 32767 : vmin);
                    q.hmax = (short)(vmax == -0x7fffffff ? -32768 : vmax);
                    qg_tail_push(need_tail, q, s_tq, &s_tq_count);
                }
                if (last_chunk)
                    qg_finish(ad, act && !need_tail, (uint32_t)(t0 + t), lo_t, n_t, vmin, vmax, exact, 0, out, narrow, wide, refine, counters);
            }
            gb += NG;
            __syncthreads();                                            // s_acc / s_hl are free again; tail pushes visible
            if (tid == 0) s_hl_count = 0;
            if (last_chunk) break;
        } while (true);
        // ---- B: 256 tail candidates accumulated -> exact tail Myers + classification, full warps ----
        const int n_tail = s_tq_count;
        if (n_tail >= ATR_QG_THREADS) {
            const QgTailItem q = s_tq[n_tail - ATR_QG_THREADS + tid];
            int lo2, n2; bool esc2;
            read_extent(len, win, q.read, lo2, n2, esc2);
            const int im = sa_tail_packed_w<WORD>(&ad, s_tail_peq, codes + woff[q.read], lo2, n2);
            const bool nohit = q.hmax == -32768;
            qg_finish(ad, true, q.read, lo2, n2, nohit ? 0x7fffffff : (int)q.hmin, nohit ? -0x7fffffff : (int)q.hmax, false, im,
                      out, narrow, wide, refine, counters);
            __syncthreads();
            if (tid == 0) s_tq_count = n_tail - ATR_QG_THREADS;
        }
    }
    // ---- the tail candidates left over ----
    __syncthreads();
    {
        const int n_tail = s_tq_count;
        const bool active = tid < n_tail;
        QgTailItem q;
        q.read = 0; q.hmin = 32767; q.hmax = -32768;
        int lo2 = 0, n2 = 0, im = 0;
        if (active) {
            q = s_tq[tid];
            bool esc2;
            read_extent(len, win, q.read, lo2, n2, esc2);
            im = sa_tail_packed_w<WORD>(&ad, s_tail_peq, codes + woff[q.read], lo2, n2);
        }
        const bool nohit = q.hmax == -32768;
        qg_finish(ad, active, q.read, lo2, n2, nohit ? 0x7fffffff : (int)q.hmin, nohit ? -0x7fffffff : (int)q.hmax, false, im,
                  out, narrow, wide, refine, counters);
    }
}

// k_refine: exact Myers over the column range the piece hits point at; dense over the refine list
template <class WORD, bool AND_MODE>
__global__ void __launch_bounds__(128) k_refine(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, atr_match* __restrict__ out, const Survivor* __restrict__ list,
        Survivor* __restrict__ narrow, Survivor* __restrict__ wide, int* __restrict__ counters) {
    __shared__ WORD s_peq[16];
    if (threadIdx.x < 16) {
        const int WB = (int)(8 * sizeof(WORD)), sh = WB - ad.m;
        s_peq[threadIdx.x] = (WORD)(((WORD)ad.peq[threadIdx.x] << sh) | (sh ? (((WORD)1 << sh) - 1) : 0));
    }
    __syncthreads();
    const int count = counters[2];
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (count + stride - 1) / stride;          // uniform trip count: the appends use warp ballots
    for (int it = 0; it < rounds; it++) {
        const int s = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool to_narrow = false, to_wide = false, narrow8 = false;
        Survivor sv;
        sv.read = 0; sv.a = 0; sv.b = 0;
        if (s < count) {
            const Survivor in = list[s];
            int lo, n; bool esc;
            read_extent(len, win, in.read, lo, n, esc);
            FilterHit hit;
            sv.read = in.read;
            if (myers_filter<WORD>(ad, s_peq, codes + woff[in.read], lo, n, hit, (int)in.a, (int)in.b)) {
                if (ad.band_ok && hit.width <= ATR_K1D_W) { to_narrow = true; sv.a = (short)hit.dlo; narrow8 = ad.split8 && hit.width <= 8; }
                else { to_wide = true; sv.a = (short)hit.c0; sv.b = (short)hit.c1; }
            } else {
                Best b;
                b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
                finalize(ad, b, n, out + in.read);
            }
        }
        list_append(to_narrow && !narrow8, sv, narrow, counters + 0);
        list_append_back(narrow8, sv, wide, counters + 3);
        list_append(to_wide, sv, wide, counters + 1);
    }
}

// W = 16: the narrow list from the front; W = 8: the survivors appended from its end (AdapterK1a.split8), list = that end
#define ATR_BAND_BATCH 256
template <bool AND_MODE, int W>
__global__ void __launch_bounds__(128) k_band(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, atr_match* __restrict__ out,
        const Survivor* __restrict__ list, const int* __restrict__ counter) {
    // The DP of a survivor runs R = min(m, n - dlo) rows: all m for an adapter inside the read, fewer for one that sticks
    // out of the read end, and a warp runs as long as its longest lane (22 of 32 lanes active on the benchmark reads). A CTA
    // therefore takes ATR_BAND_BATCH survivors at a time, counting-sorts them by R in shared memory and hands them out in
    // that order, so that the lanes of a warp finish together.
    __shared__ int s_hist[ATR_K1A_MAXM + 2];
    __shared__ unsigned short s_order[ATR_BAND_BATCH];
    __shared__ unsigned char s_key[ATR_BAND_BATCH];
    __shared__ Survivor s_sv[ATR_BAND_BATCH];
    __shared__ unsigned s_ext[ATR_BAND_BATCH];             // lo | n << 16
    const int count = *counter;
    const int tid = threadIdx.x;
    for (int base = blockIdx.x * ATR_BAND_BATCH; base < count; base += gridDim.x * ATR_BAND_BATCH) {
        const int nb = count - base < ATR_BAND_BATCH ? count - base : ATR_BAND_BATCH;
        for (int i = tid; i < ATR_K1A_MAXM + 2; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int e = tid; e < nb; e += blockDim.x) {
            const Survivor sv = W == 8 ? list[-1 - (base + e)] : list[base + e];
            int lo, n; bool esc;
            read_extent(len, win, sv.read, lo, n, esc);
            int R = n - (int)sv.a;
            R = R < 0 ? 0 : (R > ad.m ? ad.m : R);
            s_key[e] = (unsigned char)R;
            s_sv[e] = sv;
            s_ext[e] = (unsigned)lo | ((unsigned)n << 16);
            atomicAdd(&s_hist[R + 1], 1);
        }
        __syncthreads();
        if (tid == 0) { int acc = 0; for (int i = 0; i <= ad.m + 1; i++) { acc += s_hist[i]; s_hist[i] = acc; } }   // s_hist[R] = first slot of key R
        __syncthreads();
        for (int e = tid; e < nb; e += blockDim.x) s_order[atomicAdd(&s_hist[s_key[e]], 1)] = (unsigned short)e;
        __syncthreads();
        for (int e = tid; e < nb; e += blockDim.x) {
            const int o = s_order[e];
            const Survivor sv = s_sv[o];
            const int lo = (int)(s_ext[o] & 0xFFFFu), n = (int)(s_ext[o] >> 16);
            Best b;
            if (ad.flags & ATR_START_WITHIN_SEQ1) k1d_band<AND_MODE, W, true>(ad, codes + woff[sv.read], lo, n, (int)sv.a, b);
            else k1d_band<AND_MODE, W, false>(ad, codes + woff[sv.read], lo, n, (int)sv.a, b);
            finalize(ad, b, n, out + sv.read);
        }
        __syncthreads();
    }
}

template <bool AND_MODE>
__global__ void __launch_bounds__(128) k_wide(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, atr_match* __restrict__ out,
        const Survivor* __restrict__ list, const int* __restrict__ counter) {
    const int count = *counter;
    const int stride = gridDim.x * blockDim.x;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < count; s += stride) {
        const Survivor sv = list[s];
        int lo, n; bool esc;
        read_extent(len, win, sv.read, lo, n, esc);
        Best b;
        k1a_locate<AND_MODE>(ad, codes + woff[sv.read], lo, n, b, (int)sv.a, (int)sv.b);
        finalize(ad, b, n, out + sv.read);
    }
}


// ---------------------------------------------------------------------------------------------
// Anchored adapters with indels (PREFIX / SUFFIX flag sets the funnel does not take): a fixed-position piece
// filter over every read, then the register DP over the survivors only (dense list).
// ---------------------------------------------------------------------------------------------
template <bool AND_MODE>
__global__ void __launch_bounds__(256) k_filter_anchor(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, int64_t n_reads, atr_match* __restrict__ out,
        Survivor* __restrict__ list, int* __restrict__ counter) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false;
    Survivor sv;
    sv.read = (uint32_t)r; sv.a = 0; sv.b = 0;
    if (r < n_reads) {
        int lo, n; bool esc;
        read_extent(len, win, r, lo, n, esc);
        const bool routed = (esc && !AND_MODE) || n > ATR_K1A_MAXN;
        if (routed) {
            if (ad.mark_routed) {
                atr_match m;
                m.astart = m.astop = m.rstart = m.rstop = m.matches = m.errors = 0;
                m.adapter = -1; m.status = ATR_ST_ESCAPED;
                out[r] = m;
            }
        } else if (anchor_filter(ad, codes + woff[r], lo, n)) {
            want = true;
        } else {
            Best b;
            b.ref_stop = ad.m; b.q_stop = n; b.cost = ad.m + n; b.origin = 0; b.matches = 0;
            finalize(ad, b, n, out + r);
        }
    }
    list_append(want, sv, list, counter);
}

template <bool AND_MODE>
__global__ void __launch_bounds__(128) k_anchor_dp(const __grid_constant__ AdapterK1a ad,
        const uint32_t* __restrict__ codes, const uint32_t* __restrict__ woff, const uint16_t* __restrict__ len,
        const uint16_t* __restrict__ win, atr_match* __restrict__ out,
        const Survivor* __restrict__ list, const int* __restrict__ counter, int windowed) {
    // windowed: the survivors of a funnel filter stage (dearer indels) carry the column window [a, b] every accepted
    // alignment lies in (sa_classify / myers_filter: c0, c1); the register DP only runs over it, like k_wide
    const int count = *counter;
    const int stride = gridDim.x * blockDim.x;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < count; s += stride) {
        const Survivor sv = list[s];
        int lo, n; bool esc;
        read_extent(len, win, sv.read, lo, n, esc);
        if (windowed) {
            Best b;
            k1a_locate<AND_MODE>(ad, codes + woff[sv.read], lo, n, b, (int)sv.a, (int)sv.b);
            finalize(ad, b, n, out + sv.read);
        } else {
            k1a_read<AND_MODE>(ad, codes + woff[sv.read], lo, n, out + sv.read);
        }
    }
}
