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
/* 128 letters -> four position words {lo, hi} x 4 = 32 bytes per iteration, branch-free inside: four loads (byte-
 * reversed for --rc), ONE validity test (c | 0x20 must equal the letter its low nibble selects: a=0x61 c=0x63 g=0x67
 * t=0x74), eight movemasks (ASCII bit 1 -> lo plane, bit 2 -> hi plane; complement = ~hi), one 32-byte store.
 * Returns the number of 128-letter groups written, or -1 if a letter is not A/C/G/T. */
template <bool RC>
__attribute__((target("avx2")))
static int64_t pack_groups_avx2_t(const char *s, int64_t L, uint32_t *dst, int64_t w0)
{
    /* the groups cover position words w0, w0 + 4, ... (w0 < 4 head words are left to the caller so that every
     * 32-byte store of a group is aligned: a read starts at a multiple of 8 bytes, not of 32) */
    const int64_t full = (L - 32 * w0) >> 7;
    if (full <= 0) return 0;
    dst += 2 * w0;
    if (RC) L -= 32 * w0; else s += 32 * w0;     /* output word w0 = letters [32 w0, ..) forward, [.., L - 32 w0) reversed */
    const __m256i lut = _mm256_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0,
                                         0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                         15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i m0f = _mm256_set1_epi8(0x0f), c20 = _mm256_set1_epi8(0x20);
    /* software prefetch of the input, two lines per group: the hardware streamer alone feeds one core with ~8 GB/s on
     * the B200 hosts, 4 KiB of explicit look-ahead reach 11 GB/s (16 threads: 87 -> 112 GB/s); NTL_PACK_PF overrides */
    static const int pf_dist = getenv("NTL_PACK_PF") ? atoi(getenv("NTL_PACK_PF")) : 4096;
    /* the packed words are written once and read next by the DMA engine: non-temporal stores save the
     * write-allocate traffic; NTL_PACK_NT=0 disables */
    static const bool nt_enabled = !(getenv("NTL_PACK_NT") && getenv("NTL_PACK_NT")[0] == '0');
    const bool nt_store = nt_enabled && ((uintptr_t)dst & 31) == 0;
    for (int64_t qq = 0; qq < full; qq++) {
        /* --rc: output group q is made of input bytes [L - 128 (q + 1), L - 128 q); the groups are produced last to
         * first so that the INPUT is read at ascending addresses (the hardware prefetchers of the host follow an
         * ascending stream better: +7 % at 16 threads, +18 % on one). */
        const int64_t q = RC ? full - 1 - qq : qq;
        __m256i v[4];
        const char *b = RC ? s + (L - ((q + 1) << 7)) : s + (q << 7);
        if (pf_dist > 0) { _mm_prefetch(b + pf_dist, _MM_HINT_T0); _mm_prefetch(b + pf_dist + 64, _MM_HINT_T0); }
        if (!RC) {
            for (int j = 0; j < 4; j++) v[j] = _mm256_loadu_si256((const __m256i *)(b + 32 * j));
        } else {
            for (int j = 0; j < 4; j++) {
                __m256i x = _mm256_loadu_si256((const __m256i *)(b + 32 * (3 - j)));
                x = _mm256_shuffle_epi8(x, rev);
                v[j] = _mm256_permute2x128_si256(x, x, 0x01);
            }
        }
        __m256i ok = _mm256_set1_epi8(-1);
        for (int j = 0; j < 4; j++)
            ok = _mm256_and_si256(ok, _mm256_cmpeq_epi8(_mm256_or_si256(v[j], c20),
                                                        _mm256_shuffle_epi8(lut, _mm256_and_si256(v[j], m0f))));
        if (_mm256_movemask_epi8(ok) != -1) return -1;
        alignas(32) uint32_t out[8];
        for (int j = 0; j < 4; j++) {
            out[2 * j] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v[j], 6));        /* ASCII bit 1 -> code bit 0 */
            const uint32_t h = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v[j], 5));  /* ASCII bit 2 -> code bit 1 */
            out[2 * j + 1] = RC ? ~h : h;
        }
        if (nt_store) _mm256_stream_si256((__m256i *)(dst + q * 8), _mm256_load_si256((const __m256i *)out));
        else memcpy(dst + q * 8, out, 32);
    }
    if (nt_store) _mm_sfence();
    return full;
}
#endif

static bool g_have_avx2 =
#if defined(__x86_64__)
    __builtin_cpu_supports("avx2");
#else
    false;
#endif

int ntl_pack_read_2bit(const char *s, int64_t L, int rc, uint32_t *dst, int64_t n_words)
{
    init_tables();
    const int64_t n_blocks = (L + 31) >> 5;          /* position words that hold letters */
    int64_t k = 0, k_skip_lo = 0, k_skip_hi = 0;     /* words [k_skip_lo, k_skip_hi) were written by the vector loop */
#if defined(__x86_64__)
    if (g_have_avx2) {
        const int64_t w0 = ((uintptr_t)dst & 7) == 0 ? (int64_t)(((32 - ((uintptr_t)dst & 31)) & 31) >> 3) : 0;
        const int64_t full = rc ? pack_groups_avx2_t<true>(s, L, dst, w0) : pack_groups_avx2_t<false>(s, L, dst, w0);
        if (full < 0) return 1;
        if (full > 0) { k_skip_lo = w0; k_skip_hi = w0 + (full << 2); }
    }
#endif
    for (; k < n_blocks; k++) {
        if (k == k_skip_lo && k_skip_hi > k_skip_lo) { k = k_skip_hi - 1; continue; }
        uint32_t lo, hi;
        const int64_t b0 = k << 5;
        const int n = L - b0 >= 32 ? 32 : (int)(L - b0);
        bool ok;
#if defined(__x86_64__)
        if (n == 32 && g_have_avx2) ok = block_avx2(s, L, rc, b0, &lo, &hi);
        else
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
