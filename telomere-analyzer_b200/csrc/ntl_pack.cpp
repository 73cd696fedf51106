/*
 * ntl_pack.cpp -- host side of the boundary: ASCII reads -> planar bit streams in pinned memory.
 *
 * Replaces, for the GPU path, what Biostrings does when readDNAStringSet() encodes letters and when
 * reverseComplement() is applied to the chunk (NanoTel.R:2213, 2219-2221): --rc is folded into the packer, so all
 * coordinates the kernels produce are already in the reverse-complemented frame, as in the reference.
 * Layout: see ntl_dev.h (position p = bit (p - 1) of the read's stream, one {lo, hi} or {A, C, G, T} record per 32
 * positions).  AVX2 path: 32 letters -> two movemasks (bit 1 and bit 2 of the ASCII code are the 2-bit code of
 * A/C/G/T in either case); scalar path otherwise.
 */
#include "ntl_pack.h"
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

/* Biostrings DNA codes, low nibble only (gap letters - + . have no base bit). 0xFF = not DNA. */
static uint8_t g_nib[256];
static uint8_t g_is_acgt[256];
static bool g_tables = false;

static void init_tables()
{
    if (g_tables) return;
    memset(g_nib, 0xFF, sizeof g_nib);
    memset(g_is_acgt, 0, sizeof g_is_acgt);
    const char *letters = "ACGTMRWSYKVHDBN";
    const uint8_t codes[] = {1, 2, 4, 8, 3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15};
    for (int i = 0; letters[i]; i++) {
        g_nib[(unsigned char)letters[i]] = codes[i];
        g_nib[(unsigned char)(letters[i] | 0x20)] = codes[i];
    }
    g_nib[(unsigned char)'-'] = 0; g_nib[(unsigned char)'+'] = 0; g_nib[(unsigned char)'.'] = 0;
    const char *acgt = "ACGTacgt";
    for (int i = 0; acgt[i]; i++) g_is_acgt[(unsigned char)acgt[i]] = 1;
    g_tables = true;
}

int ntl_pattern_nibble(char c)
{
    init_tables();
    uint8_t n = g_nib[(unsigned char)c];
    if (n == 0xFF || n == 0) return -1;     /* gap letters are not accepted in patterns */
    return n;
}

static inline uint8_t comp_nib(uint8_t n)   /* A<->T, C<->G: reverse the 4 bits */
{
    return (uint8_t)(((n & 1) << 3) | ((n & 2) << 1) | ((n & 4) >> 1) | ((n & 8) >> 3));
}

/* 32 letters starting at base index b0 (0-based, in the OUTPUT frame) -> lo/hi masks; returns false if a letter is
 * not A/C/G/T.  n <= 32 letters are valid, the rest of the masks is zero. */
static inline bool block_scalar(const char *s, int64_t L, int rc, int64_t b0, int n, uint32_t *lo, uint32_t *hi)
{
    uint32_t l = 0, h = 0;
    for (int i = 0; i < n; i++) {
        unsigned char c = rc ? (unsigned char)s[L - 1 - (b0 + i)] : (unsigned char)s[b0 + i];
        if (!g_is_acgt[c]) return false;
        uint32_t code = (c >> 1) & 3u;
        if (rc) code ^= 2u;
        l |= (code & 1u) << i;
        h |= (code >> 1) << i;
    }
    *lo = l; *hi = h;
    return true;
}

#if defined(__x86_64__)
__attribute__((target("avx2")))
static inline bool block_avx2(const char *s, int64_t L, int rc, int64_t b0, uint32_t *lo, uint32_t *hi)
{
    __m256i v;
    if (rc) {
        v = _mm256_loadu_si256((const __m256i *)(s + (L - 32 - b0)));
        const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                             15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
        v = _mm256_shuffle_epi8(v, rev);
        v = _mm256_permute2x128_si256(v, v, 0x01);
    } else {
        v = _mm256_loadu_si256((const __m256i *)(s + b0));
    }
    /* A/C/G/T in either case <=> (c | 0x20) equals the letter its low nibble selects: a=0x61 c=0x63 g=0x67 t=0x74 */
    const __m256i lut = _mm256_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0,
                                         0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i u = _mm256_or_si256(v, _mm256_set1_epi8(0x20));
    const __m256i want = _mm256_shuffle_epi8(lut, _mm256_and_si256(v, _mm256_set1_epi8(0x0f)));
    if (_mm256_movemask_epi8(_mm256_cmpeq_epi8(u, want)) != -1) return false;
    uint32_t l = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));   /* ASCII bit 1 -> code bit 0 */
    uint32_t h = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));   /* ASCII bit 2 -> code bit 1 */
    if (rc) h = ~h;
    *lo = l; *hi = h;
    return true;
}
#endif

#if defined(__x86_64__)
/* 128 letters -> four position words {lo, hi} x 4 = 32 bytes per iteration, branch-free inside: the letters (byte-
 * reversed for --rc), ONE validity test (c | 0x20 must equal the letter its low nibble selects: a=0x61 c=0x63 g=0x67
 * t=0x74; pshufb zeroes bytes >= 0x80 itself), the two planes (ASCII bit 1 -> lo, bit 2 -> hi; complement = ~hi), and
 * four 8-byte stores {lo, hi} straight from the mask registers (a detour through a stack array costs a failed store
 * forward per group).  Returns the number of 128-letter groups written, or -1 if a letter is not A/C/G/T.
 * Two builds of the same loop: AVX2 (32 letters per vector, two movemasks per plane pair) and AVX-512 BW + VBMI (64
 * letters per vector, vptestmb yields a plane of 64 positions at once, vpermb reverses the bytes): in cache 12.6 /
 * 20 GB/s per thread for --rc against 9.4 for the round-1 loop. */
struct PackTune { int pf_dist; bool nt; int hint; };
#define NTL_PF2(p, hint) do { if ((hint) == 1) { _mm_prefetch((p), _MM_HINT_T1); _mm_prefetch((p) + 64, _MM_HINT_T1); } \
    else if ((hint) == 2) { _mm_prefetch((p), _MM_HINT_T2); _mm_prefetch((p) + 64, _MM_HINT_T2); } \
    else { _mm_prefetch((p), _MM_HINT_T0); _mm_prefetch((p) + 64, _MM_HINT_T0); } } while (0)
static const PackTune &pack_tune()
{
    /* software prefetch of the input, two lines per group: the hardware streamer alone feeds one core with ~8 GB/s on
     * the B200 hosts, 4 KiB of explicit look-ahead reach 11 GB/s (NTL_PACK_PF overrides); the packed words are written
     * once and read next by the DMA engine: non-temporal stores save the write-allocate traffic (NTL_PACK_NT=0) */
    static const PackTune t = {getenv("NTL_PACK_PF") ? atoi(getenv("NTL_PACK_PF")) : 4096,
                               !(getenv("NTL_PACK_NT") && getenv("NTL_PACK_NT")[0] == '0'),
                               getenv("NTL_PACK_HINT") ? atoi(getenv("NTL_PACK_HINT")) : 0};
    return t;
}

template <bool RC>
__attribute__((target("avx2")))
static int64_t pack_groups_avx2_t(const char *s, int64_t L, uint32_t *dst, int64_t w0)
{
    /* the groups cover position words w0, w0 + 4, ... (w0 < 4 head words are left to the caller so that the 32 bytes
     * of a group are aligned: a read starts at a multiple of 8 bytes, not of 32) */
    const int64_t full = (L - 32 * w0) >> 7;
    if (full <= 0) return 0;
    dst += 2 * w0;
    if (RC) L -= 32 * w0; else s += 32 * w0;     /* output word w0 = letters [32 w0, ..) forward, [.., L - 32 w0) reversed */
    const __m256i lut = _mm256_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0,
                                         0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                         15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i c20 = _mm256_set1_epi8(0x20);
    const int pf_dist = pack_tune().pf_dist, hint = pack_tune().hint;
    /* non-temporal stores only for the 64-byte lines this call fills completely: a read's words start and end in the
     * middle of a line (spans are 200 bytes), and a partially written line leaves the write-combining buffer as a
     * string of small uncached writes; those lines take ordinary stores */
    const bool nt_on = pack_tune().nt && ((uintptr_t)dst & 7) == 0;
    const uintptr_t nt_lo = ((uintptr_t)dst + 63) & ~(uintptr_t)63, nt_hi = (uintptr_t)(dst + full * 8) & ~(uintptr_t)63;
    for (int64_t qq = 0; qq < full; qq++) {
        /* --rc: output group q is made of input bytes [L - 128 (q + 1), L - 128 q); the groups are produced last to
         * first so that the INPUT is read at ascending addresses (the hardware prefetchers of the host follow an
         * ascending stream better: +7 % at 16 threads, +18 % on one). */
        const int64_t q = RC ? full - 1 - qq : qq;
        __m256i v[4];
        const char *b = RC ? s + (L - ((q + 1) << 7)) : s + (q << 7);
        if (pf_dist > 0) NTL_PF2(b + pf_dist, hint);
        if (!RC) {
            for (int j = 0; j < 4; j++) v[j] = _mm256_loadu_si256((const __m256i *)(b + 32 * j));
        } else {
            for (int j = 0; j < 4; j++) {               /* the two 16-byte halves swap places at the load */
                const char *p = b + 32 * (3 - j);
                const __m256i x = _mm256_inserti128_si256(_mm256_castsi128_si256(_mm_loadu_si128((const __m128i *)(p + 16))),
                                                          _mm_loadu_si128((const __m128i *)p), 1);
                v[j] = _mm256_shuffle_epi8(x, rev);
            }
        }
        __m256i ok = _mm256_cmpeq_epi8(_mm256_or_si256(v[0], c20), _mm256_shuffle_epi8(lut, v[0]));
        for (int j = 1; j < 4; j++)
            ok = _mm256_and_si256(ok, _mm256_cmpeq_epi8(_mm256_or_si256(v[j], c20), _mm256_shuffle_epi8(lut, v[j])));
        if (_mm256_movemask_epi8(ok) != -1) return -1;
        uint64_t *o = reinterpret_cast<uint64_t *>(dst + q * 8);
        const bool nt_store = nt_on && (uintptr_t)o >= nt_lo && (uintptr_t)o + 32 <= nt_hi;
        for (int j = 0; j < 4; j++) {
            const uint64_t l = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v[j], 6));   /* ASCII bit 1 -> code bit 0 */
            const uint64_t h = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v[j], 5));   /* ASCII bit 2 -> code bit 1 */
            uint64_t w = l | (h << 32);
            if (RC) w ^= 0xffffffff00000000ull;
            if (nt_store) _mm_stream_si64(reinterpret_cast<long long *>(o + j), (long long)w);
            else memcpy(o + j, &w, 8);
        }
    }
    return full;
}

template <bool RC>
__attribute__((target("avx512f,avx512bw,avx512vbmi")))
static int64_t pack_groups_avx512_t(const char *s, int64_t L, uint32_t *dst, int64_t w0)
{
    const int64_t full = (L - 32 * w0) >> 7;
    if (full <= 0) return 0;
    dst += 2 * w0;
    if (RC) L -= 32 * w0; else s += 32 * w0;
    const __m512i bit1 = _mm512_set1_epi8(2), bit2 = _mm512_set1_epi8(4), c20 = _mm512_set1_epi8(0x20);
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0));
    alignas(64) static const uint8_t revidx[64] = {
        63, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 50, 49, 48, 47, 46, 45, 44, 43, 42, 41, 40, 39, 38, 37, 36, 35, 34, 33, 32,
        31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0};
    const __m512i rev = _mm512_load_si512(revidx);
    const int pf_dist = pack_tune().pf_dist, hint = pack_tune().hint;
    const bool nt_on = pack_tune().nt && ((uintptr_t)dst & 7) == 0;     /* whole lines only, see above */
    const uintptr_t nt_lo = ((uintptr_t)dst + 63) & ~(uintptr_t)63, nt_hi = (uintptr_t)(dst + full * 8) & ~(uintptr_t)63;
    for (int64_t qq = 0; qq < full; qq++) {
        const int64_t q = RC ? full - 1 - qq : qq;      /* --rc: input read at ascending addresses, see above */
        const char *b = RC ? s + (L - ((q + 1) << 7)) : s + (q << 7);
        if (pf_dist > 0) NTL_PF2(b + pf_dist, hint);
        __m512i v0, v1;
        if (!RC) { v0 = _mm512_loadu_si512(b); v1 = _mm512_loadu_si512(b + 64); }
        else { v0 = _mm512_permutexvar_epi8(rev, _mm512_loadu_si512(b + 64)); v1 = _mm512_permutexvar_epi8(rev, _mm512_loadu_si512(b)); }
        const __mmask64 k0 = _mm512_cmpeq_epi8_mask(_mm512_or_si512(v0, c20), _mm512_shuffle_epi8(lut, v0));
        const __mmask64 k1 = _mm512_cmpeq_epi8_mask(_mm512_or_si512(v1, c20), _mm512_shuffle_epi8(lut, v1));
        if ((k0 & k1) != ~0ull) return -1;
        /* a plane of 64 positions per test: l = {lo of word 2i | lo of word 2i + 1 << 32} */
        const uint64_t l0 = _mm512_test_epi8_mask(v0, bit1), l1 = _mm512_test_epi8_mask(v1, bit1);
        uint64_t h0 = _mm512_test_epi8_mask(v0, bit2), h1 = _mm512_test_epi8_mask(v1, bit2);
        if (RC) { h0 = ~h0; h1 = ~h1; }
        const uint64_t w[4] = {(l0 & 0xffffffffu) | (h0 << 32), (l0 >> 32) | (h0 & 0xffffffff00000000ull),
                               (l1 & 0xffffffffu) | (h1 << 32), (l1 >> 32) | (h1 & 0xffffffff00000000ull)};
        uint64_t *o = reinterpret_cast<uint64_t *>(dst + q * 8);
        const bool nt_store = nt_on && (uintptr_t)o >= nt_lo && (uintptr_t)o + 32 <= nt_hi;
        if (nt_store) for (int j = 0; j < 4; j++) _mm_stream_si64(reinterpret_cast<long long *>(o + j), (long long)w[j]);
        else memcpy(o, w, 32);
    }
    return full;
}
#endif

static bool g_have_avx2 =
#if defined(__x86_64__)
    __builtin_cpu_supports("avx2");
#else
    false;
#endif
/* NTL_PACK_ISA=avx2 keeps the 256-bit loop on a host that has AVX-512 (measurements) */
static bool g_have_avx512 =
#if defined(__x86_64__)
    __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
    !(getenv("NTL_PACK_ISA") && strcmp(getenv("NTL_PACK_ISA"), "avx2") == 0);
#else
    false;
#endif

/* The group loops write with non-temporal stores and do NOT fence: a fence per read drains the write-combining
 * buffers every 20 kb and cost ~0.4 us per read; the caller issues ntl_pack_fence() once per batch of reads, before
 * anyone else (the DMA engine, another thread) is told that the words are there. */
void ntl_pack_fence(void)
{
#if defined(__x86_64__)
    _mm_sfence();
#endif
}

int ntl_pack_read_2bit(const char *s, int64_t L, int rc, uint32_t *dst, int64_t n_words)
{
    init_tables();
    const int64_t n_blocks = (L + 31) >> 5;          /* position words that hold letters */
    int64_t k = 0, k_skip_lo = 0, k_skip_hi = 0;     /* words [k_skip_lo, k_skip_hi) were written by the vector loop */
#if defined(__x86_64__)
    if (g_have_avx2 && ((uintptr_t)dst & 7) == 0) {
        /* whole groups of 128 letters from word 0 on (8-byte stores: no alignment beyond the read's own) */
        /* the groups start at the first 32-byte boundary of the output (w0 <= 3 head words are left to the loop below):
         * a group's four 8-byte stores then never straddle a line (measured on the B200 hosts, 16 threads: 124 against
         * 114 GB/s for 20 kb reads, 61 against 34 GB/s for 2 kb reads) */
        const int64_t w0 = (int64_t)(((32 - ((uintptr_t)dst & 31)) & 31) >> 3);
        const int64_t full = g_have_avx512 ? (rc ? pack_groups_avx512_t<true>(s, L, dst, w0) : pack_groups_avx512_t<false>(s, L, dst, w0))
                                           : (rc ? pack_groups_avx2_t<true>(s, L, dst, w0) : pack_groups_avx2_t<false>(s, L, dst, w0));
        if (full < 0) return 1;
        if (full > 0) { k_skip_lo = w0; k_skip_hi = w0 + (full << 2); }
    }
#endif
    /* at most three whole words at either end and one partial word are left */
    for (; k < n_blocks; k++) {
        if (k == k_skip_lo && k_skip_hi > k_skip_lo) { k = k_skip_hi - 1; continue; }
        uint32_t lo, hi;
        const int64_t b0 = k << 5;
        const int n = L - b0 >= 32 ? 32 : (int)(L - b0);
        bool ok;
#if defined(__x86_64__)
        if (g_have_avx2) {
            if (n == 32) ok = block_avx2(s, L, rc, b0, &lo, &hi);
            else {
                /* the partial word through the same vector code: its n letters in a 32-byte buffer padded with 'A'
                 * (forward: at the front; --rc: at the back, the block is the reversal of the read's first n letters) */
                char tmp[32];
                memset(tmp, 'A', 32);
                if (rc) memcpy(tmp + 32 - n, s, (size_t)n); else memcpy(tmp, s + b0, (size_t)n);
                ok = block_avx2(tmp, 32, rc, 0, &lo, &hi);
                const uint32_t m = (1u << n) - 1u;
                lo &= m; hi &= m;
            }
        } else
#endif
            ok = block_scalar(s, L, rc, b0, n, &lo, &hi);
        if (!ok) return 1;
        dst[2 * k] = lo; dst[2 * k + 1] = hi;        /* position p = b0 + i + 1 is bit i of word k */
    }
    if (n_words > n_blocks) memset(dst + 2 * n_blocks, 0, (size_t)(n_words - n_blocks) * 8);   /* the span's padding */
    return 0;
}

#if defined(__x86_64__)
/* Full 32-letter blocks of a read with IUPAC letters -> the four nibble planes, 32 letters per iteration: letter ->
 * Biostrings nibble through two 16-entry byte shuffles ((c | 0x20) - 0x60 is 1..26 for a letter), --rc = byte
 * reversal + a third shuffle that mirrors the nibble (A<->T, C<->G), one movemask per plane.  Returns the number of
 * blocks done; stops early (blocks done so far) at the first block that holds anything but IUPAC letters -- gap
 * characters and errors are left to the scalar loop. */
__attribute__((target("avx2")))
static int64_t pack_blocks_4bit_avx2(const char *s, int64_t L, int rc, uint32_t *dst)
{
    alignas(32) uint8_t t0[32], t1[32], tc[32];
    for (int i = 0; i < 16; i++) {
        t0[i] = t0[i + 16] = g_nib[0x60 + i] == 0 ? 0xFF : g_nib[0x60 + i];          /* '`' a .. o */
        t1[i] = t1[i + 16] = (0x70 + i) <= 'z' ? g_nib[0x70 + i] : 0xFF;             /* p .. z     */
        tc[i] = tc[i + 16] = comp_nib((uint8_t)i);
    }
    t0[0] = t0[16] = 0xFF;
    const __m256i T0 = _mm256_load_si256((const __m256i *)t0), T1 = _mm256_load_si256((const __m256i *)t1);
    const __m256i TC = _mm256_load_si256((const __m256i *)tc);
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                         15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i c20 = _mm256_set1_epi8(0x20), c60 = _mm256_set1_epi8(0x60), m0f = _mm256_set1_epi8(0x0f);
    const int64_t nb = L >> 5;
    for (int64_t k = 0; k < nb; k++) {
        __m256i v;
        if (rc) {
            v = _mm256_loadu_si256((const __m256i *)(s + (L - 32 - (k << 5))));
            v = _mm256_shuffle_epi8(v, rev);
            v = _mm256_permute2x128_si256(v, v, 0x01);
        } else v = _mm256_loadu_si256((const __m256i *)(s + (k << 5)));
        const __m256i idx = _mm256_sub_epi8(_mm256_or_si256(v, c20), c60);           /* 1..26 for letters */
        const __m256i in = _mm256_and_si256(_mm256_cmpgt_epi8(idx, _mm256_setzero_si256()),
                                            _mm256_cmpgt_epi8(_mm256_set1_epi8(27), idx));
        const __m256i lo4 = _mm256_and_si256(idx, m0f);
        __m256i nib = _mm256_blendv_epi8(_mm256_shuffle_epi8(T0, lo4), _mm256_shuffle_epi8(T1, lo4), _mm256_slli_epi16(idx, 3));
        /* not a letter, or a letter outside the IUPAC alphabet (table entry 0xFF): leave the block to the scalar loop */
        const __m256i bad = _mm256_or_si256(_mm256_cmpeq_epi8(nib, _mm256_set1_epi8(-1)), _mm256_xor_si256(in, _mm256_set1_epi8(-1)));
        if (_mm256_movemask_epi8(bad) != 0) return k;
        if (rc) nib = _mm256_shuffle_epi8(TC, nib);
        for (int b = 0; b < 4; b++)
            dst[4 * k + b] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(nib, 7 - b));
    }
    return nb;
}
#endif

int ntl_pack_read_4bit(const char *s, int64_t L, int rc, uint32_t *dst, int64_t n_words)
{
    init_tables();
    int64_t done = 0;                                   /* 32-letter blocks (= position words) finished */
#if defined(__x86_64__)
    if (g_have_avx2) done = pack_blocks_4bit_avx2(s, L, rc, dst);
#endif
    /* the words from `done` on: zero, then the remaining letters one by one */
    if (n_words > done) memset(dst + 4 * done, 0, (size_t)(n_words - done) * 16);
    for (int64_t p = (done << 5); p < L; p++) {         /* p: 0-based position in the output frame */
        unsigned char c = rc ? (unsigned char)s[L - 1 - p] : (unsigned char)s[p];
        uint8_t nb = g_nib[c];
        if (nb == 0xFF) return -1;
        if (rc) nb = comp_nib(nb);
        const uint32_t bit = 1u << (p & 31);
        uint32_t *q = dst + 4 * (p >> 5);
        if (nb & 1) q[0] |= bit;
        if (nb & 2) q[1] |= bit;
        if (nb & 4) q[2] |= bit;
        if (nb & 8) q[3] |= bit;
    }
    return 0;
}

void ntl_parallel_for(int64_t n, int n_threads, int64_t grain, const std::function<void(int64_t, int64_t)> &fn)
{
    if (n <= 0) return;
    if (n_threads <= 1 || n <= grain) { fn(0, n); return; }
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(grain);
            if (b >= n) break;
            int64_t e = b + grain < n ? b + grain : n;
            fn(b, e);
        }
    };
    std::vector<std::thread> th;
    int nt = n_threads;
    if ((int64_t)nt > (n + grain - 1) / grain) nt = (int)((n + grain - 1) / grain);
    th.reserve(nt);
    for (int t = 1; t < nt; t++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
}

/* Diagnostics: the rate at which the host can stream the ASCII input (see nanotel_b200.h). */
#include <chrono>
__attribute__((target("avx2")))
static uint64_t read_range_avx2(const unsigned char *p, int64_t b, int64_t e)
{
    __m256i v0 = _mm256_setzero_si256(), v1 = v0;
    int64_t i = b;
    for (; i + 64 <= e; i += 64) {
        _mm_prefetch((const char *)p + i + 4096, _MM_HINT_T0);
        v0 = _mm256_xor_si256(v0, _mm256_loadu_si256((const __m256i *)(p + i)));
        v1 = _mm256_xor_si256(v1, _mm256_loadu_si256((const __m256i *)(p + i + 32)));
    }
    v0 = _mm256_xor_si256(v0, v1);
    uint64_t acc = (uint64_t)_mm256_extract_epi64(v0, 0) ^ (uint64_t)_mm256_extract_epi64(v0, 1) ^
                   (uint64_t)_mm256_extract_epi64(v0, 2) ^ (uint64_t)_mm256_extract_epi64(v0, 3);
    for (; i < e; i++) acc ^= p[i];
    return acc;
}
static uint64_t read_range_scalar(const unsigned char *p, int64_t b, int64_t e)
{
    uint64_t acc = 0;
    int64_t i = b;
    for (; i + 8 <= e; i += 8) { uint64_t v; memcpy(&v, p + i, 8); acc ^= v; }
    for (; i < e; i++) acc ^= p[i];
    return acc;
}

/* the packer's traffic without its arithmetic: 128 bytes read, 32 bytes written with a non-temporal store */
__attribute__((target("avx2")))
static uint64_t read_write_range_avx2(const unsigned char *p, int64_t b, int64_t e, unsigned char *out)
{
    __m256i acc = _mm256_setzero_si256();
    int64_t i = b;
    for (; i + 128 <= e; i += 128) {
        _mm_prefetch((const char *)p + i + 4096, _MM_HINT_T0);
        _mm_prefetch((const char *)p + i + 4096 + 64, _MM_HINT_T0);
        const __m256i a0 = _mm256_loadu_si256((const __m256i *)(p + i)), a1 = _mm256_loadu_si256((const __m256i *)(p + i + 32));
        const __m256i a2 = _mm256_loadu_si256((const __m256i *)(p + i + 64)), a3 = _mm256_loadu_si256((const __m256i *)(p + i + 96));
        const __m256i x = _mm256_xor_si256(_mm256_xor_si256(a0, a1), _mm256_xor_si256(a2, a3));
        _mm256_stream_si256((__m256i *)(out + (i >> 2)), x);
        acc = _mm256_xor_si256(acc, x);
    }
    _mm_sfence();
    return (uint64_t)_mm256_extract_epi64(acc, 0) ^ (uint64_t)_mm256_extract_epi64(acc, 3);
}

extern "C" double ntl_host_read_gbs(const void *buf, int64_t bytes, int32_t threads, int32_t reps)
{
    if (!buf || bytes <= 0) return 0.0;
    if (threads < 1) threads = 1;
    /* reps < 0: the read + quarter-size non-temporal write pattern of the packer, -reps passes */
    const bool with_writes = reps < 0 && g_have_avx2;
    if (reps < 0) reps = -reps;
    if (reps < 1) reps = 1;
    const unsigned char *p = (const unsigned char *)buf;
    unsigned char *out = nullptr;
    if (with_writes) {
        if (posix_memalign((void **)&out, 64, (size_t)(bytes >> 2) + 256) != 0) return 0.0;
        memset(out, 0, (size_t)(bytes >> 2) + 256);                        /* touch the pages before timing */
    }
    double best = 0.0;
    std::atomic<uint64_t> sink(0);
    for (int r = 0; r < reps; r++) {
        const auto t0 = std::chrono::steady_clock::now();
        ntl_parallel_for(bytes, threads, 1 << 20, [&](int64_t b, int64_t e) {
            sink.fetch_xor(with_writes ? read_write_range_avx2(p, b, e, out)
                                       : g_have_avx2 ? read_range_avx2(p, b, e) : read_range_scalar(p, b, e));
        });
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (s > 0 && (double)bytes / s / 1e9 > best) best = (double)bytes / s / 1e9;
    }
    free(out);
    return sink.load() == 0x5a5a5a5a5a5a5a5bULL ? best + 1e-9 : best;     /* keep the reads alive */
}
