/*
 * span_model.cpp -- TEST INFRASTRUCTURE: the span scan (telomere-analyzer_b200/csrc/ntl_scan.cuh) compiled by g++ as a
 * host model (NTL_HOST_MODEL: the same arithmetic, funnel shifts and popcounts as plain C++), driven exactly as the
 * kernel drives it: main pass over every span that is neither TAIL nor SKIP, tail pass per read.  Built by
 * tests/test_span_host_model.py from the source text the library generates for NVRTC (ntl_jit_get_source), so the CPU
 * suite checks the very text that is compiled for the GPU.  Never part of the product.
 */
#define NTL_HOST_MODEL 1
#define __device__
#define __forceinline__ inline
#include <stdint.h>
#include NTL_GENERATED_SOURCE

/* class bytes (one per span and track, blocks per span <= 8): optional */
static uint32_t blk_thr = 0xffffffffu;
static uint8_t *cls_planes[3] = {nullptr, nullptr, nullptr};
static uint8_t **cls = nullptr;
extern "C" int span_model_set_cls(uint32_t thr, uint8_t *c0, uint8_t *c1, uint8_t *c2)
{
    blk_thr = thr; cls_planes[0] = c0; cls_planes[1] = c1; cls_planes[2] = c2; cls = c0 ? cls_planes : nullptr;
    return 0;
}

template <int NPL>
static void run_arena(const uint32_t *arena, int64_t n_spans, const uint8_t *flags, int n_reads, const int32_t *len,
                      const int64_t *woff, const uint8_t *fmt, int64_t cnt_base, uint16_t *c0, uint16_t *c1, uint16_t *c2)
{
    constexpr int W = NTL_J_W, SG = NTL_J_SG, BPS = 32 * W / SG;
    uint16_t *cnt[3] = {c0, c1, c2};
    for (int64_t s = 0; s < n_spans; s++) {
        const unsigned f = flags[s];
        if (f & (NTL_SPAN_TAIL | NTL_SPAN_SKIP)) continue;
        uint16_t *out[3];
        for (int t = 0; t < 3; t++) out[t] = t < NTL_J_NTRACKS ? cnt[t] + (cnt_base + s * BPS) : nullptr;
        u32 cb[3];
        ntl_span<NPL, false, W, SG>(arena + s * W * NPL, (f & NTL_SPAN_FIRST) != 0, 0, out, blk_thr, cb);
        if (cls && BPS <= 8) for (int t = 0; t < NTL_J_NTRACKS; t++) cls[t][cnt_base / BPS + s] = (uint8_t)cb[t];
    }
    for (int r = 0; r < n_reads; r++) {
        if ((fmt[r] != 0) != (NPL == 4)) continue;
        const int L = len[r];
        if (L <= 0) continue;
        const int64_t s_first = woff[r] / W;
        const int nsp = (L + 32 * W - 1) / (32 * W);
        const int ltail = L - (nsp - 1) * 32 * W;
        const int s_lo = (nsp >= 2 && ltail < NTL_DEV_MAX_LEN) ? nsp - 2 : nsp - 1;
        for (int s = s_lo; s < nsp; s++) {
            uint16_t *out[3];
            for (int t = 0; t < 3; t++) out[t] = t < NTL_J_NTRACKS ? cnt[t] + (cnt_base + (s_first + s) * BPS) : nullptr;
            u32 cb[3];
            ntl_span<NPL, true, W, SG>(arena + (s_first + s) * W * NPL, s == 0, L - s * 32 * W, out, blk_thr, cb);
            if (cls && BPS <= 8) for (int t = 0; t < NTL_J_NTRACKS; t++) cls[t][cnt_base / BPS + s_first + s] = (uint8_t)cb[t];
        }
    }
}

extern "C" int span_model_geometry(int *W, int *SG, int *T) { *W = NTL_J_W; *SG = NTL_J_SG; *T = NTL_J_NTRACKS; return 0; }

/* arena pointers address position word 0 of span 0 (at least 4 readable words before them, one span after the last) */
extern "C" int span_model_run(const uint32_t *arena2, int64_t n_spans2, const uint8_t *flags2, const uint32_t *arena4,
                              int64_t n_spans4, const uint8_t *flags4, int n_reads, const int32_t *len,
                              const int64_t *woff, const uint8_t *fmt, uint16_t *c0, uint16_t *c1, uint16_t *c2)
{
    constexpr int BPS = 32 * NTL_J_W / NTL_J_SG;
    run_arena<2>(arena2, n_spans2, flags2, n_reads, len, woff, fmt, 0, c0, c1, c2);
    if (n_spans4 > 0) run_arena<4>(arena4, n_spans4, flags4, n_reads, len, woff, fmt, n_spans2 * BPS, c0, c1, c2);
    return 0;
}
