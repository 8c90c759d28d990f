// host_pack.cpp -- host-side twin of k_pack (atr_pack_reads_host): ASCII reads -> the packed layout of
// include/atropos_b200.h in HOST memory, so that a caller who keeps reads packed ships 75 + 6 bytes per 150-nt read over
// PCIe instead of 150 (atr_locate_batch_host_packed). Plain C++ (g++), threaded; AVX2 where the CPU has it (runtime
// dispatch), scalar otherwise. Bit-for-bit the device packer's output (tests/test_abi.py, tests/test_gpu_pack.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/atropos_b200.h"

namespace {

struct Tables {
    unsigned char iupac[256];     // 4-bit code (0 for everything outside the IUPAC letters; X = 0 too)
    unsigned char exact[256];     // 1: the byte is one of the 16 upper-case IUPAC letters (code == identity of the letter)
    unsigned char upper[256];
    Tables() {
        std::memset(iupac, 0, sizeof iupac);
        std::memset(exact, 0, sizeof exact);
        for (int c = 0; c < 256; c++) upper[c] = (unsigned char)((c >= 'a' && c <= 'z') ? c - 32 : c);
        const char* letters = "XACMGRSVTWYHKDBN";            // index = code: A=1 C=2 G=4 T=8 and their unions (_align.pyx:46-83)
        for (int code = 0; code < 16; code++) {
            const unsigned char ch = (unsigned char)letters[code];
            iupac[ch] = (unsigned char)code; iupac[ch + 32] = (unsigned char)code;
            exact[ch] = 1;
        }
        iupac[(unsigned char)'U'] = 8; iupac[(unsigned char)'u'] = 8;     // U = T for the wildcard modes, never exact
    }
};
const Tables T;

// one read, scalar: words [0, nw) of its packed form; returns 1 if a byte is not exactly representable
inline int pack_read_scalar(const uint8_t* s, int len, int fold, uint32_t* out) {
    int esc = 0;
    const int nw = (len + 7) >> 3;
    for (int w = 0; w < nw; w++) {
        uint32_t v = 0;
        const int hi = std::min(8, len - 8 * w);
        for (int t = 0; t < hi; t++) {
            unsigned char c = s[8 * w + t];
            if (fold) c = T.upper[c];
            v |= (uint32_t)T.iupac[c] << (4 * t);
            esc |= !T.exact[c];
        }
        out[w] = v;
    }
    return esc;
}

#if defined(__x86_64__)
// 32 bases -> 4 packed words. Fast path for upper-case A/C/G/T/N: the low nibble of the byte picks code and expected
// letter from two 16-entry tables; any other byte fails the letter test and the block is redone by the scalar code.
__attribute__((target("avx2"))) inline bool pack32_avx2(const uint8_t* s, uint32_t* out) {
    const __m256i x = _mm256_loadu_si256((const __m256i*)s);
    const __m256i lo = _mm256_and_si256(x, _mm256_set1_epi8(0x0F));
    // low nibble: 'A' 1, 'C' 3, 'G' 7, 'T' 4, 'N' 0xE
    const __m256i code_lut = _mm256_setr_epi8(0, 1, 0, 2, 8, 0, 0, 4, 0, 0, 0, 0, 0, 0, 15, 0, 0, 1, 0, 2, 8, 0, 0, 4, 0, 0, 0, 0, 0, 0, 15, 0);
    // unused slots hold a byte whose own low nibble differs from the slot index: no input byte can equal it
#define U(i) (char)(0x80 | (((i) + 1) & 15))
    const __m256i char_lut = _mm256_setr_epi8(U(0), 'A', U(2), 'C', 'T', U(5), U(6), 'G', U(8), U(9), U(10), U(11), U(12), U(13), 'N', U(15),
                                              U(0), 'A', U(2), 'C', 'T', U(5), U(6), 'G', U(8), U(9), U(10), U(11), U(12), U(13), 'N', U(15));
#undef U
    const __m256i code = _mm256_shuffle_epi8(code_lut, lo);
    const __m256i want = _mm256_shuffle_epi8(char_lut, lo);
    if (_mm256_movemask_epi8(_mm256_cmpeq_epi8(want, x)) != -1) return false;
    // two codes per byte: even position = low nibble
    const __m256i pairs = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x1001));       // lo*1 + hi*16 in every 16-bit lane
    const __m256i bytes = _mm256_packus_epi16(pairs, pairs);                             // lanes: [0..7 | 0..7 | 8..15 | 8..15]
    const __m256i ord = _mm256_permute4x64_epi64(bytes, 0x08);                           // qwords 0, 2 -> the low 128 bits
    _mm_storeu_si128((__m128i*)out, _mm256_castsi256_si128(ord));
    return true;
}

// the last 1..31 bases of a read, same fast path: 32 bytes are loaded (the caller guarantees that they are readable: they
// belong to the following reads of the batch), the lanes beyond `rem` are excluded from the letter test and give code 0,
// and only the words the read owns are stored (a 150-nt read ends with 22 such bases: the scalar tail was 2/3 of its cost)
__attribute__((target("avx2"))) inline bool pack_tail_avx2(const uint8_t* s, int rem, uint32_t* out) {
    const __m256i x = _mm256_loadu_si256((const __m256i*)s);
    const __m256i iota = _mm256_setr_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31);
    const __m256i valid = _mm256_cmpgt_epi8(_mm256_set1_epi8((char)rem), iota);         // 0xFF where lane < rem
    const __m256i lo = _mm256_and_si256(x, _mm256_set1_epi8(0x0F));
    const __m256i code_lut = _mm256_setr_epi8(0, 1, 0, 2, 8, 0, 0, 4, 0, 0, 0, 0, 0, 0, 15, 0, 0, 1, 0, 2, 8, 0, 0, 4, 0, 0, 0, 0, 0, 0, 15, 0);
#define U(i) (char)(0x80 | (((i) + 1) & 15))
    const __m256i char_lut = _mm256_setr_epi8(U(0), 'A', U(2), 'C', 'T', U(5), U(6), 'G', U(8), U(9), U(10), U(11), U(12), U(13), 'N', U(15),
                                              U(0), 'A', U(2), 'C', 'T', U(5), U(6), 'G', U(8), U(9), U(10), U(11), U(12), U(13), 'N', U(15));
#undef U
    const __m256i want = _mm256_shuffle_epi8(char_lut, lo);
    const __m256i ok = _mm256_or_si256(_mm256_cmpeq_epi8(want, x), _mm256_andnot_si256(valid, _mm256_set1_epi8((char)0xFF)));
    if (_mm256_movemask_epi8(ok) != -1) return false;
    const __m256i code = _mm256_and_si256(_mm256_shuffle_epi8(code_lut, lo), valid);
    const __m256i pairs = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x1001));
    const __m256i bytes = _mm256_packus_epi16(pairs, pairs);
    const __m256i ord = _mm256_permute4x64_epi64(bytes, 0x08);
    alignas(16) uint32_t tmp[4];
    _mm_store_si128((__m128i*)tmp, _mm256_castsi256_si128(ord));
    const int nw = (rem + 7) >> 3;
    for (int w = 0; w < nw; w++) out[w] = tmp[w];
    return true;
}

// slack: bytes readable behind the read (>= 31 lets the tail take the vector path)
__attribute__((target("avx2"))) int pack_read_avx2(const uint8_t* s, int len, int fold, uint32_t* out, int64_t slack) {
    int esc = 0, p = 0;
    for (; p + 32 <= len; p += 32) {
        if (!pack32_avx2(s + p, out + (p >> 3))) esc |= pack_read_scalar(s + p, 32, fold, out + (p >> 3));
    }
    if (p < len) {
        if (slack < 32 - (len - p) || !pack_tail_avx2(s + p, len - p, out + (p >> 3)))
            esc |= pack_read_scalar(s + p, len - p, fold, out + (p >> 3));
    }
    return esc;
}
#endif

int pack_read_plain(const uint8_t* s, int len, int fold, uint32_t* out, int64_t /*slack*/) { return pack_read_scalar(s, len, fold, out); }

typedef int (*pack_fn)(const uint8_t*, int, int, uint32_t*, int64_t);

pack_fn pick() {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) return pack_read_avx2;
#endif
    return pack_read_plain;
}

}  // namespace

extern "C" int atr_pack_reads_host(const uint8_t* ascii, const int64_t* offsets, int64_t n, int fold_case, int n_threads,
                                   uint32_t* codes, uint32_t* woff, uint16_t* len) {
    if (!offsets || !codes || !woff || !len || n < 0 || (n > 0 && !ascii && offsets[n] > offsets[0])) return ATR_E_ARG;
    if (n == 0) { woff[0] = 0; return ATR_OK; }
    unsigned hw = std::thread::hardware_concurrency();
    int nt = n_threads > 0 ? n_threads : (int)(hw ? hw : 1);
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n / 4096 + 1));
    // word offsets: exclusive scan of ceil(len / 8), in two threaded passes (per-range sums, then the offsets; the serial
    // scan was a quarter of the call at 16 threads)
    std::vector<uint64_t> part((size_t)nt + 1, 0);
    std::vector<int> bad((size_t)nt, 0);
    auto run = [&](auto fn) {
        if (nt == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(fn, t);
        for (auto& x : th) x.join();
    };
    run([&](int t) {
        uint64_t w = 0;
        for (int64_t i = n * t / nt; i < n * (t + 1) / nt; i++) {
            const int64_t l = offsets[i + 1] - offsets[i];
            if (l < 0 || l > 32767) { bad[(size_t)t] = 1; return; }
            w += (uint64_t)((l + 7) >> 3);
        }
        part[(size_t)t + 1] = w;
    });
    for (int t = 0; t < nt; t++) {
        if (bad[(size_t)t]) return ATR_E_LIMIT;
        part[(size_t)t + 1] += part[(size_t)t];
    }
    if (part[(size_t)nt] > 0xFFFFFFF0ull) return ATR_E_LIMIT;
    woff[n] = (uint32_t)part[(size_t)nt];
    const pack_fn pack = pick();
    const int64_t end = offsets[n];
    run([&](int t) {
        uint64_t w = part[(size_t)t];
        for (int64_t i = n * t / nt; i < n * (t + 1) / nt; i++) {
            const int l = (int)(offsets[i + 1] - offsets[i]);
            woff[i] = (uint32_t)w;
            const int esc = pack(ascii + offsets[i], l, fold_case, codes + w, end - offsets[i + 1]);
            len[i] = (uint16_t)((unsigned)l | (esc ? 0x8000u : 0u));
            w += (uint64_t)((l + 7) >> 3);
        }
    });
    return ATR_OK;
}
