/*
 * ntl_api.cpp -- the C ABI of libnanotel_b200.so (include/nanotel_b200.h): contexts, pinned/device buffers, the
 * batch pipeline  pack -> H2D -> [filter] -> scan -> locate -> D2H  on one or several GPUs, and the host-side Serial
 * logic.
 *
 * Replaces the body of NanoTel.R's chunk loop (NanoTel.R:2219-2258) for one --nrec chunk per call.  The reference's
 * only parallelism, the 8-way fan-out of a chunk over forked workers (NanoTel.R:2207, :2234-2254), becomes: the chunk
 * is cut into contiguous shards balanced by bases, one shard per GPU of ntl_params.device_ids, every shard runs the
 * same pipeline on its own stream with its own packer threads, and the records are gathered on the host in input
 * order.  No collective: reads are independent.
 * There is no CPU fallback: every compute entry point needs a CUDA device and fails with NTL_ERR_CUDA otherwise.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nanotel_b200.h"
#include "ntl_dev.h"
#include "ntl_jit.h"
#include "ntl_pack.h"

extern "C" {
cudaError_t ntl_k_set_params(const ntl_dev_params *p, cudaStream_t st);
cudaError_t ntl_k_scan_generic(const ntl_read_args *a, int grid, cudaStream_t st);
cudaError_t ntl_k_scan_generic_occupancy(int *blocks_per_sm);
cudaError_t ntl_k_filter(const ntl_read_args *a, cudaStream_t st);
cudaError_t ntl_k_items(const uint8_t *active, int n_items, int32_t *items, uint32_t *counter, cudaStream_t st);
cudaError_t ntl_k_triage(const ntl_read_args *a, cudaStream_t st);
cudaError_t ntl_k_locate(const ntl_read_args *a, int grid, cudaStream_t st);
cudaError_t ntl_k_cls(const ntl_read_args *a, long long n_bytes, int T, uint32_t blk_thr, cudaStream_t st);
cudaError_t ntl_k_locate_occupancy(int *blocks_per_sm);
cudaError_t ntl_k_gather_windows(const ntl_read_args *a, const int64_t *list, int n_list, uint16_t *dst, int T, cudaStream_t st);
}

namespace {

thread_local char g_create_err[512] = "";

struct PinnedBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes, bool keep = false)
    {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = bytes + bytes / 4 + 4096;
        void *np = nullptr;
        cudaError_t e = cudaHostAlloc(&np, ncap, cudaHostAllocPortable);
        if (e != cudaSuccess) return e;
        if (p) { if (keep) memcpy(np, p, cap); cudaFreeHost(p); }
        p = np; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    uint64_t gen = 0;                       /* bumped by every (re)allocation: a new buffer holds nothing */
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = bytes + bytes / 4 + 4096;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, ncap);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = ncap; gen++;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

#define NTL_EVENT_RING 256
#define ARENA_LEAD 256          /* bytes in front of an arena: the scan's bulk copy starts 16 bytes before span 0 */
enum { ST_EMPTY = 0, ST_PACKED = 1, ST_UPLOADED = 2, ST_RAN = 3, ST_DOWNLOADED = 4 };

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

/* NTL_TRACE=1: phase times of the host side on stderr (development aid) */
struct Trace {
    bool on; double t;
    Trace() : on(getenv("NTL_TRACE") != nullptr), t(now_ms()) {}
    void mark(const char *what) { if (on) { const double n = now_ms(); fprintf(stderr, "[ntl] %-22s %8.3f ms\n", what, n - t); t = n; } }
};

int32_t count_windows(int64_t L, int32_t S)
{
    if (L < 1 || S < 1) return 0;
    int64_t n = (L - 1) / S + 1;
    int64_t last = 1 + (n - 1) * (int64_t)S;
    if (2 * (L - last) < S) n -= 1;          /* NanoTel.R:220: L - last(idx_start) < sub_length / 2 */
    return (int32_t)n;
}

int gcd_i(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

/* Span geometry for a subseq_length (ntl_dev.h): the largest block SG that divides S, with at most 64 blocks per
 * window, such that spans of W <= 32 position words hold whole blocks (32 W multiple of SG).  Among the possible W the
 * one whose shared-memory stride conflicts least (gcd(W, 16) small), then the longest.  BPS = 0: no such geometry (S
 * with a large odd part, S < 16): the generic scan kernel handles the batch, blocks = windows. */
void choose_geometry(int S, ntl_dev_params *d)
{
    d->SG = S; d->Q = 1; d->W = 2; d->BPS = 0; d->cls_bps = 8;
    for (int q = 1; q <= 64; q++) {
        if (S % q) continue;
        const int sg = S / q;
        if (sg < 16) break;
        const int base = sg / gcd_i(sg, 32);
        if (base > 32) continue;
        int best_w = 0, best_c = 99;
        for (int k = 32 / base; k >= 1; k--) {
            const int w = base * k;
            if (w < 12 && best_w) break;
            const int c = gcd_i(w, 16);
            if (c < best_c) { best_c = c; best_w = w; }
        }
        if (32 * best_w / sg > 64) continue;
        d->SG = sg; d->Q = q; d->W = best_w; d->BPS = 32 * best_w / sg;
        d->cls_bps = d->BPS <= 8 ? d->BPS : 8;              /* class bits: one byte per span, else dense (ntl_dev.h) */
        return;
    }
}

void generic_geometry(ntl_dev_params *d) { d->SG = d->S; d->Q = 1; d->W = 2; d->BPS = 0; d->cls_bps = 8; }

} // namespace

/* ===================================================================================================== one device */
struct ntl_dev_ctx {
    ntl_params prm;
    ntl_dev_params dev;
    int device = 0, n_sms = 0, generic_grid = 0, locate_grid = 0, host_threads = 1;
    int span_align = 1;                     /* reads start at span indices that are multiples of this */
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr};
    cudaEvent_t ring[NTL_EVENT_RING][5] = {{nullptr}};
    int pending = 0, pending_launches = 0;
    char err[512] = "";
    ntl_jit_kernel *jit = nullptr;
    int scan_path = NTL_SCAN_GENERIC;
    std::string scan_note;

    int state = ST_EMPTY;
    int32_t n_reads = 0, n2 = 0, n4 = 0;
    int64_t words2 = 0, words4 = 0;         /* position words of the 2-bit / 4-bit arena (whole spans)             */
    int64_t spans2 = 0, spans4 = 0;         /* spans per arena, rounded up to whole work items                      */
    int64_t total_blocks = 0, bases = 0;    /* entries of one track's block-count plane                             */
    size_t dens_offset = 0;
    size_t meta_bytes = 0, off_len = 0, off_woff = 0, off_cntoff = 0, off_order = 0, off_fmt = 0, off_fl2 = 0, off_fl4 = 0;
    size_t arena4_off = 0;                  /* byte offset of the 4-bit arena (its lead included) inside the packed buffers */

    PinnedBuf h_packed, h_meta, h_results, h_cnt, h_stages, h_list;
    DevBuf d_packed, d_meta, d_results, d_cnt, d_cls, d_stages, d_pass, d_counter, d_thr, d_flags, d_kept, d_list, d_items;
    std::vector<int64_t> kept_off;          /* per read: first element of its rows in h_cnt, -1 = not downloaded */
    std::vector<uint16_t> win_scratch;      /* ntl_get_windows of a read that was not kept: fetched on demand     */
    ntl_timings tm;
};

namespace {

int fail(ntl_dev_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    char *dst = c ? c->err : g_create_err;
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(c, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail((c), NTL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                        __FILE__, __LINE__);                                                          \
    } while (0)

/* --patterns / --tvr_patterns -> device pattern table (unique(), NanoTel.R:328,362; sorted by length so that
 * patterns of equal length share one dilation in K2) */
int digest_patterns(ntl_dev_ctx *c, const std::vector<std::string> &in, ntl_dev_pat *out, int32_t *n_out,
                    int32_t *n_groups, int32_t *group_begin)
{
    std::vector<std::string> uniq;
    for (const auto &s : in)
        if (std::find(uniq.begin(), uniq.end(), s) == uniq.end()) uniq.push_back(s);
    std::stable_sort(uniq.begin(), uniq.end(), [](const std::string &a, const std::string &b) { return a.size() < b.size(); });
    *n_out = (int32_t)uniq.size();
    *n_groups = 0;
    for (size_t i = 0; i < uniq.size(); i++) {
        const std::string &s = uniq[i];
        if (s.empty() || s.size() > NTL_MAX_PATLEN)
            return fail(c, NTL_ERR_PATTERN, "pattern '%s': length must be 1..%d (NanoTel.R:589)", s.c_str(), NTL_MAX_PATLEN);
        ntl_dev_pat &d = out[i];
        memset(&d, 0, sizeof d);
        d.m = (int32_t)s.size();
        d.fixed = 1;
        for (size_t j = 0; j < s.size(); j++) {
            int nb = ntl_pattern_nibble(s[j]);
            if (nb < 0) return fail(c, NTL_ERR_PATTERN, "pattern '%s': letter '%c' is not an IUPAC DNA letter", s.c_str(), s[j]);
            if (strchr("WSMKRYBDHVN", s[j])) d.fixed = 0;       /* case-sensitive, as str_detect at NanoTel.R:334 */
            d.nib[j] = (uint8_t)nb;
        }
        for (int j = 0; j < d.m; j++) {
            static const uint8_t base_nib[4] = {1, 2, 8, 4};    /* 2-bit code 0..3 = A, C, T, G */
            for (int code = 0; code < 4; code++) {
                bool acc = d.fixed ? (d.nib[j] == base_nib[code]) : ((d.nib[j] & base_nib[code]) != 0);
                d.mux2[j][code] = acc ? 0xffffffffu : 0u;
            }
            for (int b = 0; b < 4; b++) {
                d.mux4[j][b] = (d.nib[j] & (1 << b)) ? 0xffffffffu : 0u;
                if (d.nib[j] & (1 << b)) d.q4[b] |= 1u << j;
            }
        }
        if (i == 0 || uniq[i - 1].size() != s.size()) group_begin[(*n_groups)++] = (int32_t)i;
    }
    group_begin[*n_groups] = (int32_t)uniq.size();
    return NTL_OK;
}

void set_span_align(ntl_dev_ctx *c)
{
    /* every read starts on a multiple of 8 block entries (16 bytes of a count plane) */
    c->span_align = c->dev.BPS > 0 ? 8 / gcd_i(c->dev.BPS, 8) : 1;
}

/* argument checks of ntl_create + the device parameter block; no CUDA call */
int digest_params(ntl_dev_ctx *c, const ntl_params *p)
{
    if (p->n_patterns < 1 || p->n_patterns > NTL_MAX_PATTERNS || !p->patterns)
        return fail(c, NTL_ERR_ARG, "n_patterns must be 1..%d", NTL_MAX_PATTERNS);
    if (p->n_tvr < 0 || p->n_tvr > NTL_MAX_PATTERNS || (p->n_tvr > 0 && !p->tvr_patterns))
        return fail(c, NTL_ERR_ARG, "n_tvr must be 0..%d", NTL_MAX_PATTERNS);
    /* the merged last window of split_telo is up to S + ceil(S / 2) - 1 wide and the telomeric-count thresholds are
     * kept as uint16: S + ceil(S / 2) - 1 <= 65535 */
    if (p->subseq_length < 1 || p->subseq_length > NTL_MAX_SUBSEQ)
        return fail(c, NTL_ERR_ARG, "subseq_length must be 1..%d", NTL_MAX_SUBSEQ);
    if (!(p->min_density == p->min_density)) return fail(c, NTL_ERR_ARG, "min_density is NaN");
    c->prm = *p;
    std::vector<std::string> pat_store, tvr_store;
    for (int i = 0; i < p->n_patterns; i++) {
        if (!p->patterns[i]) return fail(c, NTL_ERR_ARG, "NULL pattern");
        pat_store.push_back(p->patterns[i]);
    }
    for (int i = 0; i < p->n_tvr; i++) {
        if (!p->tvr_patterns[i]) return fail(c, NTL_ERR_ARG, "NULL tvr pattern");
        tvr_store.push_back(p->tvr_patterns[i]);
    }
    c->prm.patterns = nullptr; c->prm.tvr_patterns = nullptr; c->prm.device_ids = nullptr;

    ntl_dev_params &d = c->dev;
    memset(&d, 0, sizeof d);
    int rc = digest_patterns(c, pat_store, d.main_pat, &d.n_main, &d.n_main_groups, d.main_group_begin);
    if (rc == NTL_OK) rc = digest_patterns(c, tvr_store, d.tvr_pat, &d.n_tvr, &d.n_tvr_groups, d.tvr_group_begin);
    if (rc != NTL_OK) return rc;
    d.n_tracks = p->n_tvr > 0 ? 3 : 2;
    /* track A keeps the raw hit list iff --patterns is ONE token without ambiguity letters (NanoTel.R:347-354) */
    d.raw_hits_A = (p->n_patterns == 1 && d.main_pat[0].fixed) ? 1 : 0;
    d.S = p->subseq_length;
    d.right_edge = p->right_edge ? 1 : 0;
    d.use_filter = p->use_filter ? 1 : 0;
    d.debug_stages = (p->options & NTL_OPT_DEBUG_STAGES) ? 1 : 0;
    d.min_density = p->min_density;
    d.filter_threshold = p->min_density * 0.8;                 /* NanoTel.R:2143 global_min_density*0.8 */
    {   /* smallest count c with !(c / S < min_density): a width-S window is telomeric from there on (:751-752) */
        int64_t cnt = (int64_t)(p->min_density * (double)d.S) - 2;
        if (cnt < 0) cnt = 0;
        while (cnt <= d.S && ((double)cnt / (double)d.S < p->min_density)) cnt++;
        d.thr_reg = (int32_t)cnt;
    }
    choose_geometry(d.S, &d);
    if (p->options & NTL_OPT_NO_JIT) generic_geometry(&d);      /* generic kernel, generic layout */
    set_span_align(c);
    return NTL_OK;
}

void dev_destroy(ntl_dev_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->jit) ntl_jit_free(c->jit);
    c->h_packed.release(); c->h_meta.release(); c->h_results.release(); c->h_cnt.release(); c->h_stages.release();
    c->h_list.release(); c->d_kept.release(); c->d_list.release(); c->d_items.release();
    c->d_packed.release(); c->d_meta.release(); c->d_results.release(); c->d_cnt.release(); c->d_cls.release(); c->d_stages.release();
    c->d_pass.release(); c->d_counter.release(); c->d_thr.release(); c->d_flags.release();
    for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int k = 0; k < NTL_EVENT_RING; k++)
        for (int i = 0; i < 5; i++) if (c->ring[k][i]) cudaEventDestroy(c->ring[k][i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int dev_create(ntl_dev_ctx **out, const ntl_params *p, int device, int host_threads)
{
    *out = nullptr;
    ntl_dev_ctx *c = new (std::nothrow) ntl_dev_ctx();
    if (!c) return fail(nullptr, NTL_ERR_NOMEM, "out of memory");
    int rc = digest_params(c, p);
    if (rc != NTL_OK) { strncpy(g_create_err, c->err, sizeof g_create_err); delete c; return rc; }
    c->host_threads = host_threads;

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        fail(nullptr, NTL_ERR_CUDA, "cudaSetDevice(%d) failed: %s -- libnanotel_b200 has no CPU fallback", device,
             cudaGetErrorString(e));
        delete c;
        return NTL_ERR_CUDA;
    }
    c->device = device;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        fail(nullptr, NTL_ERR_CUDA, "CUDA initialisation failed: %s -- libnanotel_b200 has no CPU fallback", cudaGetErrorString(e));
        delete c;
        return NTL_ERR_CUDA;
    }

    /* ---- the specialised scan kernel for this pattern set and geometry: precompiled, cached or NVRTC */
    if (p->options & NTL_OPT_NO_JIT) {
        c->scan_path = NTL_SCAN_GENERIC;
        c->scan_note = "generic scan kernel requested (NTL_OPT_NO_JIT)";
    } else if (c->dev.BPS <= 0) {
        c->scan_path = NTL_SCAN_GENERIC;
        char b[200];
        snprintf(b, sizeof b, "subseq_length %d has no span geometry (no divisor >= 16 whose odd part is <= 32): generic scan kernel", c->dev.S);
        c->scan_note = b;
    } else {
        std::string jerr;
        c->jit = ntl_jit_build(&c->dev, prop.major, prop.minor, &jerr);
        if (c->jit) {
            const int o = ntl_jit_origin(c->jit);
            c->scan_path = o == NTL_JIT_ORIGIN_PRECOMPILED ? NTL_SCAN_PRECOMPILED : o == NTL_JIT_ORIGIN_CACHE ? NTL_SCAN_CACHED : NTL_SCAN_NVRTC;
        } else {
            if (p->options & NTL_OPT_REQUIRE_JIT) {
                fail(nullptr, NTL_ERR_JIT, "no specialised scan kernel: %s", jerr.c_str());
                dev_destroy(c);
                return NTL_ERR_JIT;
            }
            /* loud, not silent: the generic kernel is several times slower */
            c->scan_path = NTL_SCAN_GENERIC;
            c->scan_note = "no precompiled cubin for this pattern set / subseq_length and NVRTC failed (" + jerr +
                           "): generic scan kernel, several times slower";
            fprintf(stderr, "libnanotel_b200: WARNING: %s\n", c->scan_note.c_str());
            generic_geometry(&c->dev);                           /* the generic kernel uses the generic layout */
            set_span_align(c);
        }
    }

    for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreate(&c->ev[i]);
    if (e == cudaSuccess) e = ntl_k_set_params(&c->dev, c->stream);
    int bps = 0;
    if (e == cudaSuccess) e = ntl_k_scan_generic_occupancy(&bps);
    if (e == cudaSuccess) e = c->d_counter.ensure(64);
    if (e == cudaSuccess) {
        /* thr[w] = smallest covered count for which a window of width w is telomeric, i.e. the smallest c with
         * !((double)c / (double)w < min_density) (NanoTel.R:467, :751-752) -- found with that very division so the
         * kernels can classify windows with an integer compare.  Widths reach S + S/2 (merged last window). */
        const int32_t S = c->dev.S;
        std::vector<uint16_t> thr((size_t)2 * S + 2, 0);
        for (int32_t w = 1; w <= 2 * S + 1; w++) {
            int64_t cnt = (int64_t)(p->min_density * (double)w) - 2;
            if (cnt < 0) cnt = 0;
            while (cnt <= w && ((double)cnt / (double)w < p->min_density)) cnt++;
            thr[w] = (uint16_t)(cnt > 65535 ? 65535 : cnt);
        }
        /* dens[c] = (double)c / (double)S, the density of a width-S window holding c covered bases (:467) */
        std::vector<double> dens((size_t)S + 1);
        for (int32_t cc = 0; cc <= S; cc++) dens[cc] = (double)cc / (double)S;
        const size_t thr_bytes = (thr.size() * 2 + 15) & ~(size_t)15;
        e = c->d_thr.ensure(thr_bytes + dens.size() * 8);
        if (e == cudaSuccess)
            e = cudaMemcpy(c->d_thr.p, thr.data(), thr.size() * 2, cudaMemcpyHostToDevice);
        if (e == cudaSuccess)
            e = cudaMemcpy((char *)c->d_thr.p + thr_bytes, dens.data(), dens.size() * 8, cudaMemcpyHostToDevice);
        c->dens_offset = thr_bytes;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        fail(nullptr, NTL_ERR_CUDA, "CUDA initialisation failed: %s", cudaGetErrorString(e));
        dev_destroy(c);
        return NTL_ERR_CUDA;
    }
    c->n_sms = prop.multiProcessorCount;
    if (bps < 1) bps = 1;
    c->generic_grid = c->n_sms * bps;
    { int lb = 0; if (ntl_k_locate_occupancy(&lb) != cudaSuccess || lb < 1) lb = 4; c->locate_grid = c->n_sms * lb; }
    *out = c;
    return NTL_OK;
}

/* ---- sizes */
inline int64_t spans_of(int64_t L, int W) { return (L + 32 * (int64_t)W - 1) / (32 * (int64_t)W); }
inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

/* bytes of one track's class-bit plane (ntl_dev.h), a multiple of 16 */
size_t cls_plane_bytes(const ntl_dev_ctx *c)
{
    return (((size_t)c->total_blocks + c->dev.cls_bps - 1) / c->dev.cls_bps + 15 + 16) & ~(size_t)15;
}
inline uint32_t blk_thr_of(const ntl_dev_ctx *c) { return (uint32_t)((c->dev.thr_reg + c->dev.Q - 1) / c->dev.Q); }

int ensure_device_buffers(ntl_dev_ctx *c)
{
    const int32_t n = c->n_reads;
    const int T = c->dev.n_tracks;
    const size_t packed = c->arena4_off + ARENA_LEAD + (size_t)c->words4 * 16 + 64;
    CK(c, c->d_packed.ensure(packed));
    CK(c, c->d_meta.ensure(c->meta_bytes + 16));
    CK(c, c->d_results.ensure((size_t)n * sizeof(ntl_read_result) + 64));
    CK(c, c->d_cnt.ensure((size_t)c->total_blocks * 2 * T + 64));
    CK(c, c->d_cls.ensure(cls_plane_bytes(c) * T + 64));
    CK(c, c->d_pass.ensure((size_t)n + 64));
    CK(c, c->d_flags.ensure((size_t)n * 20 + 128));       /* candidate list + per-candidate join state of the locate kernel */
    const size_t n_items = (size_t)(c->spans2 + c->spans4) / NTL_ITEM_SPANS + 2;
    CK(c, c->d_items.ensure(n_items * 5 + 64));          /* item_active bytes, then the int32 item lists */
    if (c->dev.debug_stages) CK(c, c->d_stages.ensure((size_t)n * 3 * sizeof(ntl_stage) + 64));
    return NTL_OK;
}

/* Span flags of one read that starts at span s0 of its arena. */
void write_span_flags(uint8_t *fl, int64_t s0, int64_t L, int W)
{
    const int64_t nsp = spans_of(L, W);
    const int64_t ltail = L - (nsp - 1) * 32 * W;
    const int64_t tail_from = (nsp >= 2 && ltail < NTL_DEV_MAX_LEN) ? nsp - 2 : nsp - 1;
    for (int64_t s = 0; s < nsp; s++)
        fl[s0 + s] = (uint8_t)((s == 0 ? NTL_SPAN_FIRST : 0) | (s >= tail_from ? NTL_SPAN_TAIL : 0));
}

/* overlap = true: the packed words are copied to the device in 8 MiB pieces while the remaining reads are still
 * being packed (the calling thread issues the copies between its own grains), so that PCIe time hides behind the
 * packer; the batch ends up in state UPLOADED. */
int dev_pack(ntl_dev_ctx *c, const char *const *seq, const int64_t *len, int32_t n, bool overlap)
{
    if (!seq || !len || n < 0) return fail(c, NTL_ERR_ARG, "ntl_batch_pack: bad arguments");
    CK(c, cudaSetDevice(c->device));
    const double t0 = now_ms();
    Trace tr;
    c->state = ST_EMPTY;
    c->n_reads = n;
    const int rcflag = c->prm.rc ? 1 : 0;
    const int W = c->dev.W, SG = c->dev.SG, BPS = c->dev.BPS;
    const bool spans = BPS > 0;
    const int64_t align = c->span_align;

    /* ---- first pass: lengths, span / word offsets of every read in the 2-bit arena */
    int64_t sp = 0, bases = 0;
    std::vector<int64_t> first_span((size_t)n);
    for (int32_t i = 0; i < n; i++) {
        const int64_t L = len[i];
        if (L < 0) return fail(c, NTL_ERR_SEQUENCE, "read %d has a negative length", i);
        if (L > (1LL << 30)) return fail(c, NTL_ERR_SEQUENCE, "read %d is longer than 2^30 bases", i);
        if (L > 0 && !seq[i]) return fail(c, NTL_ERR_ARG, "read %d: NULL sequence", i);
        /* a zero-length read: NanoTel.R drops it in filter_reads (width < 1000, :2124) and stops on it otherwise
         * (seq(1, 0, by = S), :216): it gets no words and is marked FILTERED / REF_ERROR after the kernels */
        first_span[i] = sp;
        sp += round_up(L > 0 ? spans_of(L, W) : 0, align);
        bases += L;
    }
    c->spans2 = round_up(sp, NTL_ITEM_SPANS);
    c->spans4 = 0;
    c->words2 = (c->spans2 + 1) * W;                      /* + one span: the word after the last one is readable */
    c->words4 = 0;
    c->arena4_off = (ARENA_LEAD + (size_t)c->words2 * 8 + 255) & ~(size_t)255;
    c->bases = bases;

    /* ---- table layout inside one pinned block (one H2D copy) */
    size_t off = 0;
    c->off_len = off;    off += ((size_t)n * 4 + 15) & ~(size_t)15;
    c->off_woff = off;   off += (size_t)n * 8;
    c->off_cntoff = off; off += (size_t)n * 8;
    c->off_order = off;  off += ((size_t)n * 4 + 15) & ~(size_t)15;
    c->off_fmt = off;    off += ((size_t)n + 15) & ~(size_t)15;
    c->off_fl2 = off;    off += (size_t)c->spans2;         /* multiples of 32 */
    c->off_fl4 = off;    off += (size_t)c->spans2 + NTL_ITEM_SPANS;   /* worst case: every read moves to the 4-bit arena */
    const size_t meta_cap = off;
    c->meta_bytes = c->off_fl4;                            /* grows if the batch has 4-bit reads */
    CK(c, c->h_meta.ensure(meta_cap + 16));
    char *mb = (char *)c->h_meta.p;
    int32_t *h_len = (int32_t *)(mb + c->off_len);
    int64_t *h_woff = (int64_t *)(mb + c->off_woff);
    int64_t *h_cntoff = (int64_t *)(mb + c->off_cntoff);
    int32_t *h_order = (int32_t *)(mb + c->off_order);
    uint8_t *h_fmt = (uint8_t *)(mb + c->off_fmt);
    uint8_t *h_fl2 = (uint8_t *)(mb + c->off_fl2);
    uint8_t *h_fl4 = (uint8_t *)(mb + c->off_fl4);

    int64_t gblocks = 0;                                   /* generic geometry: blocks are numbered per read */
    for (int32_t i = 0; i < n; i++) {
        h_len[i] = (int32_t)len[i];
        h_woff[i] = first_span[i] * W;
        h_fmt[i] = 0;
        if (spans) h_cntoff[i] = first_span[i] * BPS;
        else { h_cntoff[i] = gblocks; gblocks += round_up((len[i] + SG - 1) / SG, 8); }
    }
    c->total_blocks = spans ? c->spans2 * BPS : gblocks + 8;
    /* the span flags of a read (2.9 M bytes for cfg2) are written by the thread that packs it (pack_range); here only
     * the rounding to whole items */
    if (spans) memset(h_fl2 + sp, NTL_SPAN_SKIP, (size_t)(c->spans2 - sp));
    tr.mark("tables");
    CK(c, c->h_packed.ensure(c->arena4_off + 64));
    uint32_t *hp = (uint32_t *)((char *)c->h_packed.p + ARENA_LEAD);
    memset(c->h_packed.p, 0, ARENA_LEAD);

    /* ---- work order of the read-wise kernels: longest reads first */
    auto work_order = [&]() {
        int32_t maxc = 0;
        std::vector<int32_t> chunks(n);
        for (int32_t i = 0; i < n; i++) {
            chunks[i] = (int32_t)(((int64_t)h_len[i] + 4095) >> 12);
            if (chunks[i] > maxc) maxc = chunks[i];
        }
        std::vector<int64_t> cnt((size_t)maxc + 2, 0), pos((size_t)maxc + 2, 0);
        for (int32_t i = 0; i < n; i++) cnt[chunks[i]]++;
        int64_t acc = 0;
        for (int32_t k = maxc; k >= 0; k--) { pos[k] = acc; acc += cnt[k]; }   /* descending */
        for (int32_t i = 0; i < n; i++) h_order[pos[chunks[i]]++] = i;
    };

    /* ---- 2-bit packing, all host threads; reads with other letters are queued for the 4-bit arena */
    std::vector<int32_t> iupac;
    std::mutex mu;
    auto pack_range = [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; i++) {
            if (len[i] == 0) continue;
            if (spans) {
                write_span_flags(h_fl2, first_span[i], len[i], W);
                const int64_t nsp = spans_of(len[i], W);
                for (int64_t sgap = nsp; sgap < round_up(nsp, align); sgap++) h_fl2[first_span[i] + sgap] = NTL_SPAN_SKIP;   /* alignment gap */
            }
            const int64_t nw = round_up(spans_of(len[i], W), align) * W;
            if (ntl_pack_read_2bit(seq[i], len[i], rcflag, hp + 2 * h_woff[i], nw) != 0) {
                std::lock_guard<std::mutex> g(mu);
                iupac.push_back((int32_t)i);
            }
        }
        ntl_pack_fence();                       /* the packer's non-temporal stores, once per grain */
    };
    int64_t up_words = 0;                      /* position words already handed to cudaMemcpyAsync (overlap mode) */
    uint64_t packed_gen = 0;
    if (!overlap) {
        ntl_parallel_for(n, c->host_threads, 64, pack_range);
    } else {
        const int64_t grain = 64, ngr = (n + grain - 1) / grain;
        const int64_t piece = 1 << 20;          /* 8 MiB of 2-bit position words */
        std::unique_ptr<std::atomic<uint8_t>[]> done(new std::atomic<uint8_t>[(size_t)ngr + 1]);
        for (int64_t g = 0; g <= ngr; g++) done[g].store(0, std::memory_order_relaxed);
        std::atomic<int64_t> next(0);
        auto do_grain = [&](int64_t g) {
            pack_range(g * grain, std::min<int64_t>(n, (g + 1) * grain));
            done[g].store(1, std::memory_order_release);
        };
        auto worker = [&]() {
            for (;;) {
                const int64_t g = next.fetch_add(1);
                if (g >= ngr) break;
                do_grain(g);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < c->host_threads && t < ngr; t++) th.emplace_back(worker);
        work_order();                           /* the packers do not need it */
        cudaError_t cerr = cudaSuccess;
        {
            const int rc0 = ensure_device_buffers(c);
            if (rc0 != NTL_OK || cudaEventRecord(c->ev[0], c->stream) != cudaSuccess) {
                next.store(ngr);                                        /* stop the packers, then report */
                for (auto &t : th) t.join();
                return rc0 != NTL_OK ? rc0 : fail(c, NTL_ERR_CUDA, "cudaEventRecord failed");
            }
        }
        packed_gen = c->d_packed.gen;
        cerr = cudaMemsetAsync(c->d_packed.p, 0, ARENA_LEAD, c->stream);
        uint32_t *dp = (uint32_t *)((char *)c->d_packed.p + ARENA_LEAD);
        const int64_t data_words = sp * W;                               /* words that the packers write */
        int64_t uf = 0;
        for (;;) {
            while (uf < ngr && done[uf].load(std::memory_order_acquire)) uf++;
            const int64_t fw = uf == ngr ? data_words : h_woff[uf * grain];
            if (cerr == cudaSuccess && (fw - up_words >= piece || (uf == ngr && fw > up_words))) {
                cerr = cudaMemcpyAsync(dp + 2 * up_words, hp + 2 * up_words, (size_t)(fw - up_words) * 8,
                                       cudaMemcpyHostToDevice, c->stream);
                up_words = fw;
            }
            if (uf == ngr) break;
            const int64_t g = next.fetch_add(1);
            if (g < ngr) do_grain(g);
            else std::this_thread::yield();
        }
        for (auto &t : th) t.join();
        /* the tables go last: the packers wrote the span flags */
        if (cerr == cudaSuccess) cerr = cudaMemcpyAsync(c->d_meta.p, c->h_meta.p, c->meta_bytes, cudaMemcpyHostToDevice, c->stream);
        if (cerr != cudaSuccess) return fail(c, NTL_ERR_CUDA, "cudaMemcpyAsync (packed reads) failed: %s", cudaGetErrorString(cerr));
    }
    tr.mark("pack 2-bit (+copies)");
    if (!iupac.empty()) {
        /* reads with IUPAC letters move to the 4-bit arena (own span numbering; their 2-bit spans are skipped) */
        std::sort(iupac.begin(), iupac.end());
        int64_t sp4 = 0;
        std::vector<int64_t> fs4(iupac.size());
        for (size_t k = 0; k < iupac.size(); k++) {
            fs4[k] = sp4;
            sp4 += round_up(spans_of(len[iupac[k]], W), align);
        }
        c->spans4 = round_up(sp4, NTL_ITEM_SPANS);
        c->words4 = (c->spans4 + 1) * W;
        CK(c, c->h_packed.ensure(c->arena4_off + ARENA_LEAD + (size_t)c->words4 * 16 + 64, /*keep=*/true));
        hp = (uint32_t *)((char *)c->h_packed.p + ARENA_LEAD);
        uint32_t *hp4 = (uint32_t *)((char *)c->h_packed.p + c->arena4_off + ARENA_LEAD);
        memset((char *)c->h_packed.p + c->arena4_off, 0, ARENA_LEAD);
        if (spans) memset(h_fl4, NTL_SPAN_SKIP, (size_t)c->spans4);
        int bad = -1;
        ntl_parallel_for((int64_t)iupac.size(), c->host_threads, 4, [&](int64_t b, int64_t e) {
            for (int64_t k = b; k < e; k++) {
                const int32_t i = iupac[k];
                const int64_t nw = round_up(spans_of(len[i], W), align) * W;
                if (ntl_pack_read_4bit(seq[i], len[i], rcflag, hp4 + 4 * fs4[k] * W, nw) != 0) bad = i;
            }
        });
        if (bad >= 0) return fail(c, NTL_ERR_SEQUENCE, "read %d holds a letter outside the DNA alphabet", bad);
        for (size_t k = 0; k < iupac.size(); k++) {
            const int32_t i = iupac[k];
            if (spans) {
                const int64_t nsp = spans_of(len[i], W);
                for (int64_t s = 0; s < nsp; s++) h_fl2[first_span[i] + s] = NTL_SPAN_SKIP;
                write_span_flags(h_fl4, fs4[k], len[i], W);
                h_cntoff[i] = (c->spans2 + fs4[k]) * BPS;
            }
            h_woff[i] = fs4[k] * W;
            h_fmt[i] = 1;
        }
        if (spans) c->total_blocks = (c->spans2 + c->spans4) * BPS;
        c->meta_bytes = c->off_fl4 + (size_t)c->spans4;
    }
    c->n4 = (int32_t)iupac.size(); c->n2 = n - c->n4;
    tr.mark("pack 4-bit");

    /* overlap mode built the order and sent the tables while the workers were packing, assuming 2-bit reads only */
    const bool tables_sent = overlap && iupac.empty();
    if (!overlap) work_order();
    tr.mark("work order");
    c->tm = ntl_timings();
    c->tm.pack_ms = now_ms() - t0;
    c->tm.bases = bases;
    c->tm.packed_bytes = sp * W * 8 + c->words4 * 16;
    c->tm.window_bytes = c->total_blocks * 2 * c->dev.n_tracks;
    c->state = ST_PACKED;
    if (overlap) {
        /* the 4-bit arena (if any) and the tables follow; a re-allocated device buffer holds nothing: start over */
        int rc1 = ensure_device_buffers(c);
        if (rc1 != NTL_OK) return rc1;
        const int64_t data_words = sp * W;
        if (c->d_packed.gen != packed_gen) { up_words = 0; CK(c, cudaMemsetAsync(c->d_packed.p, 0, ARENA_LEAD, c->stream)); }
        if (data_words > up_words)
            CK(c, cudaMemcpyAsync((char *)c->d_packed.p + ARENA_LEAD + (size_t)up_words * 8,
                                  (char *)c->h_packed.p + ARENA_LEAD + (size_t)up_words * 8,
                                  (size_t)(data_words - up_words) * 8, cudaMemcpyHostToDevice, c->stream));
        if (c->words4 > 0)
            CK(c, cudaMemcpyAsync((char *)c->d_packed.p + c->arena4_off, (char *)c->h_packed.p + c->arena4_off,
                                  ARENA_LEAD + (size_t)c->words4 * 16, cudaMemcpyHostToDevice, c->stream));
        if (c->meta_bytes > 0 && !tables_sent)
            CK(c, cudaMemcpyAsync(c->d_meta.p, c->h_meta.p, c->meta_bytes, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaEventRecord(c->ev[1], c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        CK(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
        c->tm.h2d_ms = ms;                       /* first copy issued -> last copy done; overlaps pack_ms */
        c->tm.h2d_bytes = data_words * 8 + c->words4 * 16 + (int64_t)c->meta_bytes;
        c->tm.pack_ms = now_ms() - t0;
        tr.mark("tail copies + sync");
        c->state = ST_UPLOADED;
    }
    return NTL_OK;
}

int dev_upload(ntl_dev_ctx *c)
{
    if (c->state < ST_PACKED) return fail(c, NTL_ERR_STATE, "ntl_batch_upload before ntl_batch_pack");
    CK(c, cudaSetDevice(c->device));
    { int rc0 = ensure_device_buffers(c); if (rc0 != NTL_OK) return rc0; }
    CK(c, cudaEventRecord(c->ev[0], c->stream));
    const size_t b2 = ARENA_LEAD + (size_t)c->words2 * 8;
    CK(c, cudaMemcpyAsync(c->d_packed.p, c->h_packed.p, b2, cudaMemcpyHostToDevice, c->stream));
    if (c->words4 > 0)
        CK(c, cudaMemcpyAsync((char *)c->d_packed.p + c->arena4_off, (char *)c->h_packed.p + c->arena4_off,
                              ARENA_LEAD + (size_t)c->words4 * 16, cudaMemcpyHostToDevice, c->stream));
    if (c->meta_bytes > 0)
        CK(c, cudaMemcpyAsync(c->d_meta.p, c->h_meta.p, c->meta_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaEventRecord(c->ev[1], c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    c->tm.h2d_ms = ms;
    c->tm.h2d_bytes = (int64_t)b2 + c->words4 * 16 + (int64_t)c->meta_bytes;
    c->state = ST_UPLOADED;
    return NTL_OK;
}

void fill_read_args(const ntl_dev_ctx *c, ntl_read_args *out)
{
    const int32_t n = c->n_reads;
    const int T = c->dev.n_tracks;
    char *dm = (char *)c->d_meta.p;
    ntl_read_args ra;
    memset(&ra, 0, sizeof ra);
    ra.arena2 = (const uint32_t *)((const char *)c->d_packed.p + ARENA_LEAD);
    ra.arena4 = (const uint32_t *)((const char *)c->d_packed.p + c->arena4_off + ARENA_LEAD);
    ra.len = (const int32_t *)(dm + c->off_len);
    ra.woff = (const int64_t *)(dm + c->off_woff);
    ra.cnt_off = (const int64_t *)(dm + c->off_cntoff);
    ra.fmt = (const uint8_t *)(dm + c->off_fmt);
    ra.pass = c->dev.use_filter ? (uint8_t *)c->d_pass.p : nullptr;
    for (int t = 0; t < 3; t++) ra.cnt[t] = t < T ? (uint16_t *)c->d_cnt.p + (size_t)t * c->total_blocks : nullptr;
    for (int t = 0; t < 3; t++) ra.cls[t] = t < T ? (uint8_t *)c->d_cls.p + (size_t)t * cls_plane_bytes(c) : nullptr;
    ra.results = c->d_results.p;
    ra.thr = (const uint16_t *)c->d_thr.p;
    ra.dens = (const double *)((const char *)c->d_thr.p + c->dens_offset);
    ra.order = (const int32_t *)(dm + c->off_order);
    ra.cand = (int32_t *)c->d_flags.p;
    ra.cand_state = (int32_t *)c->d_flags.p + (((size_t)n + 3) & ~(size_t)3);
    ra.counters = (uint32_t *)c->d_counter.p + 4;
    ra.stages = c->dev.debug_stages ? c->d_stages.p : nullptr;
    if (c->dev.BPS > 0) {
        ra.span_flags[0] = (uint8_t *)(dm + c->off_fl2);
        ra.span_flags[1] = (uint8_t *)(dm + c->off_fl4);
        ra.item_active[0] = (uint8_t *)c->d_items.p;
        ra.item_active[1] = (uint8_t *)c->d_items.p + c->spans2 / NTL_ITEM_SPANS;
    }
    ra.n_reads = n;
    *out = ra;
}

/* Enqueue one pass of the hot path (filter, scan, locate) on the context stream without waiting.  Event quads live
 * in a ring so that a timed loop of back-to-back passes still yields per-kernel device times. */
int dev_enqueue(ntl_dev_ctx *c)
{
    if (c->state < ST_UPLOADED) return fail(c, NTL_ERR_STATE, "ntl_batch_enqueue before ntl_batch_upload");
    CK(c, cudaSetDevice(c->device));
    if (c->pending >= NTL_EVENT_RING) return fail(c, NTL_ERR_STATE, "more than %d passes enqueued without ntl_batch_wait", NTL_EVENT_RING);
    const int32_t n = c->n_reads;
    const int T = c->dev.n_tracks;
    cudaEvent_t *ev = c->ring[c->pending];
    for (int i = 0; i < 5; i++)
        if (!ev[i]) CK(c, cudaEventCreate(&ev[i]));

    ntl_read_args ra;
    fill_read_args(c, &ra);
    const bool spans = c->dev.BPS > 0;
    const bool filt = c->dev.use_filter && n > 0;
    const int64_t items2 = c->spans2 / NTL_ITEM_SPANS, items4 = c->spans4 / NTL_ITEM_SPANS;
    int32_t *item_list = (int32_t *)((char *)c->d_items.p + (((size_t)(items2 + items4) + 15) & ~(size_t)15));

    int launches = 0;
    CK(c, cudaMemsetAsync(c->d_counter.p, 0, 64, c->stream));
    CK(c, cudaEventRecord(ev[0], c->stream));
    if (filt) {
        if (spans) CK(c, cudaMemsetAsync(c->d_items.p, 0, (size_t)(items2 + items4), c->stream));
        CK(c, ntl_k_filter(&ra, c->stream)); launches++;
        if (spans) {
            if (c->n2 > 0) { CK(c, ntl_k_items(ra.item_active[0], (int)items2, item_list, ra.counters + 2, c->stream)); launches++; }
            if (c->n4 > 0) { CK(c, ntl_k_items(ra.item_active[1], (int)items4, item_list + items2, ra.counters + 3, c->stream)); launches++; }
        }
    }
    CK(c, cudaEventRecord(ev[1], c->stream));
    c->tm.scan_is_jit = 0;
    if (spans && c->jit) {
        for (int arena = 0; arena < 2; arena++) {
            if ((arena == 0 ? c->n2 : c->n4) <= 0) continue;
            ntl_scan_args sa;
            memset(&sa, 0, sizeof sa);
            sa.arena = arena ? ra.arena4 : ra.arena2;
            sa.flags = ra.span_flags[arena];
            if (filt) { sa.items = item_list + (arena ? items2 : 0); sa.n_items_dev = ra.counters + 2 + arena; }
            sa.n_items = (int32_t)(arena ? items4 : items2);
            sa.work_counter = ra.counters + 5 + arena;
            sa.n_reads = n; sa.len = ra.len; sa.woff = ra.woff; sa.fmt = ra.fmt; sa.pass = ra.pass;
            sa.cnt_base = arena ? c->spans2 * c->dev.BPS : 0;
            for (int t = 0; t < 3; t++) sa.cnt[t] = ra.cnt[t];
            for (int t = 0; t < 3; t++) sa.cls[t] = c->dev.BPS <= 8 ? ra.cls[t] : nullptr;
            sa.cls_base = arena ? c->spans2 : 0;
            sa.blk_thr = (int32_t)blk_thr_of(c);
            cudaError_t e = ntl_jit_launch(c->jit, &sa, arena, c->n_sms, c->stream);
            if (e != cudaSuccess) return fail(c, NTL_ERR_CUDA, "span scan kernel launch failed: %s", cudaGetErrorString(e));
            launches++;
        }
        c->tm.scan_is_jit = 1;
    } else if (n > 0) {
        CK(c, cudaMemsetAsync(c->d_cnt.p, 0, (size_t)c->total_blocks * 2 * T, c->stream));
        CK(c, ntl_k_scan_generic(&ra, c->generic_grid, c->stream));
        launches++;
    }
    if (n > 0 && !(spans && c->jit && c->dev.BPS <= 8)) {     /* the scan kernel did not write the class bits itself */
        CK(c, ntl_k_cls(&ra, (long long)((c->total_blocks + 7) / 8), T, blk_thr_of(c), c->stream));
        launches++;
    }
    CK(c, cudaEventRecord(ev[2], c->stream));
    if (n > 0) {
        CK(c, ntl_k_triage(&ra, c->stream)); launches++;
        CK(c, cudaEventRecord(ev[4], c->stream));
        CK(c, ntl_k_locate(&ra, c->locate_grid, c->stream)); launches++;
    } else CK(c, cudaEventRecord(ev[4], c->stream));
    CK(c, cudaEventRecord(ev[3], c->stream));
    c->pending_launches += launches;
    c->pending++;
    return NTL_OK;
}

/* Wait for the enqueued passes; timings hold the SUMS over those passes, tm.steps their number. */
int dev_wait(ntl_dev_ctx *c)
{
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    double f = 0, s = 0, l = 0, g = 0;
    for (int k = 0; k < c->pending; k++) {
        float ms = 0.f;
        cudaEvent_t *ev = c->ring[k];
        CK(c, cudaEventElapsedTime(&ms, ev[0], ev[1])); f += ms;
        CK(c, cudaEventElapsedTime(&ms, ev[1], ev[2])); s += ms;
        CK(c, cudaEventElapsedTime(&ms, ev[2], ev[3])); l += ms;
        CK(c, cudaEventElapsedTime(&ms, ev[2], ev[4])); g += ms;
    }
    if (c->pending > 0) {
        c->tm.filter_ms = f; c->tm.scan_ms = s; c->tm.locate_ms = l; c->tm.triage_ms = g;
        c->tm.steps = c->pending;
        c->tm.kernel_launches = c->pending_launches;
        { uint32_t nc = 0; CK(c, cudaMemcpy(&nc, (uint32_t *)c->d_counter.p + 4, 4, cudaMemcpyDeviceToHost)); c->tm.candidates = (int32_t)nc; }
        c->state = ST_RAN;
    }
    c->pending = 0; c->pending_launches = 0;
    return NTL_OK;
}

inline int64_t blocks_of(const ntl_dev_ctx *c, int64_t L) { return (L + c->dev.SG - 1) / c->dev.SG; }

/* Results come back in two steps: the 64-byte records (and the debug stages) first; then, for the reads the keep rule
 * retained -- the only ones whose window tables the caller needs (NanoTel.R:1876-1918) -- the block counts, gathered
 * on the device into one contiguous block.  The tables of all other reads stay on the device until the next batch
 * and are fetched on demand by ntl_get_windows(). */
int dev_download(ntl_dev_ctx *c)
{
    if (c->state < ST_RAN) return fail(c, NTL_ERR_STATE, "ntl_batch_download before ntl_batch_run");
    CK(c, cudaSetDevice(c->device));
    const int32_t n = c->n_reads;
    const int T = c->dev.n_tracks;
    const size_t rbytes = (size_t)n * sizeof(ntl_read_result);
    const size_t sbytes = c->dev.debug_stages ? (size_t)n * 3 * sizeof(ntl_stage) : 0;
    CK(c, c->h_results.ensure(rbytes + 64));
    if (sbytes) CK(c, c->h_stages.ensure(sbytes + 64));
    CK(c, cudaEventRecord(c->ev[6], c->stream));
    if (rbytes) CK(c, cudaMemcpyAsync(c->h_results.p, c->d_results.p, rbytes, cudaMemcpyDeviceToHost, c->stream));
    if (sbytes) CK(c, cudaMemcpyAsync(c->h_stages.p, c->d_stages.p, sbytes, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));

    ntl_read_result *res = (ntl_read_result *)c->h_results.p;
    const int32_t *h_len = (const int32_t *)((const char *)c->h_meta.p + c->off_len);
    c->kept_off.assign((size_t)n, -1);
    int64_t n_kept = 0, elems = 0;
    /* one pass over the records; a kept read went through the locate kernel, so the candidate count bounds the list */
    const int64_t list_cap = std::min<int64_t>(n, std::max<int64_t>(c->tm.candidates, 0));
    CK(c, c->h_list.ensure((size_t)list_cap * 16 + 16));
    {
        int64_t *list = (int64_t *)c->h_list.p;
        for (int32_t i = 0; i < n; i++) {
            if (h_len[i] == 0) {
                /* zero-length read: filter_reads drops it (:2124); without the filter NanoTel.R stops on it (:216) */
                memset(&res[i], 0, sizeof res[i]);
                res[i].status = c->dev.use_filter ? NTL_READ_FILTERED : (NTL_READ_REF_ERROR | NTL_READ_NO_WINDOWS);
                continue;
            }
            if (!(res[i].status & NTL_READ_KEEP)) continue;
            if (n_kept >= list_cap) return fail(c, NTL_ERR_STATE, "more kept reads than locate candidates");
            list[2 * n_kept] = i; list[2 * n_kept + 1] = elems; n_kept++;
            c->kept_off[(size_t)i] = elems;
            elems += round_up(blocks_of(c, h_len[i]), 8) * T;
        }
    }
    if (n_kept > 0) {
        int64_t *list = (int64_t *)c->h_list.p;
        CK(c, c->d_list.ensure((size_t)n_kept * 16));
        CK(c, c->d_kept.ensure((size_t)elems * 2 + 64));
        CK(c, c->h_cnt.ensure((size_t)elems * 2 + 64));
        CK(c, cudaMemcpyAsync(c->d_list.p, list, (size_t)n_kept * 16, cudaMemcpyHostToDevice, c->stream));
        ntl_read_args ra;
        fill_read_args(c, &ra);
        CK(c, ntl_k_gather_windows(&ra, (const int64_t *)c->d_list.p, (int)n_kept, (uint16_t *)c->d_kept.p, T, c->stream));
        CK(c, cudaMemcpyAsync(c->h_cnt.p, c->d_kept.p, (size_t)elems * 2, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(c, cudaEventRecord(c->ev[7], c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(c, cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]));
    c->tm.d2h_ms = ms;
    c->tm.d2h_bytes = (int64_t)(rbytes + (size_t)elems * 2 + sbytes);
    c->state = ST_DOWNLOADED;
    return NTL_OK;
}

/* covered bases of window k (0-based) of a read from its block counts (ntl_dev.h: Q blocks, the last window takes the rest) */
inline int32_t window_count(const uint16_t *blk, int32_t k, int32_t n_win, int64_t n_blocks, int Q)
{
    const int64_t b0 = (int64_t)k * Q, b1 = k == n_win - 1 ? n_blocks : b0 + Q;
    int32_t c = 0;
    for (int64_t b = b0; b < b1; b++) c += blk[b];
    return c;
}

int dev_get_windows(ntl_dev_ctx *c, int32_t read_idx, int32_t track, int32_t cap, int32_t *start_index,
                    int32_t *end_index, int32_t *covered, double *density)
{
    if (c->state < ST_DOWNLOADED) return fail(c, NTL_ERR_STATE, "ntl_get_windows before the batch was downloaded");
    if (read_idx < 0 || read_idx >= c->n_reads || track < 0 || track >= c->dev.n_tracks)
        return fail(c, NTL_ERR_ARG, "ntl_get_windows: read or track out of range");
    const ntl_read_result *r = (const ntl_read_result *)c->h_results.p + read_idx;
    if (r->status & NTL_READ_FILTERED) return 0;
    const int32_t n = r->n_win, S = c->dev.S;
    const int32_t L = ((const int32_t *)((const char *)c->h_meta.p + c->off_len))[read_idx];
    const int64_t nb = blocks_of(c, L);
    const uint16_t *blk;
    if (c->kept_off[(size_t)read_idx] >= 0) {
        blk = (const uint16_t *)c->h_cnt.p + c->kept_off[(size_t)read_idx] + (size_t)track * (size_t)round_up(nb, 8);
    } else {                                    /* not a kept read: its table is still on the device */
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, NTL_ERR_CUDA, "cudaSetDevice failed");
        c->win_scratch.resize((size_t)nb + 8);
        const uint16_t *src = (const uint16_t *)c->d_cnt.p + (size_t)track * c->total_blocks + r->win_offset;
        cudaError_t e = cudaMemcpyAsync(c->win_scratch.data(), src, (size_t)nb * 2, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return fail(c, NTL_ERR_CUDA, "ntl_get_windows: %s", cudaGetErrorString(e));
        blk = c->win_scratch.data();
    }
    for (int32_t k = 0; k < n && k < cap; k++) {
        const int32_t ws = 1 + k * S, we = (k == n - 1) ? L : (k + 1) * S;
        const int32_t cnt = window_count(blk, k, n, nb, c->dev.Q);
        if (start_index) start_index[k] = ws;
        if (end_index) end_index[k] = we;
        if (covered) covered[k] = cnt;
        if (density) density[k] = (double)cnt / (double)(we - ws + 1);     /* NanoTel.R:467 */
    }
    return n;
}

/* every window count of every read of the shard, compact, into out[0 .. cap); returns the number of entries */
int64_t dev_get_window_counts(ntl_dev_ctx *c, int32_t track, uint16_t *out, int64_t cap)
{
    if (c->state < ST_DOWNLOADED) return fail(c, NTL_ERR_STATE, "ntl_get_window_counts before the batch was downloaded");
    if (track < 0 || track >= c->dev.n_tracks || (!out && cap > 0)) return fail(c, NTL_ERR_ARG, "ntl_get_window_counts: bad arguments");
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, NTL_ERR_CUDA, "cudaSetDevice failed");
    std::vector<uint16_t> raw;
    if (cap > 0) {
        raw.resize((size_t)c->total_blocks + 8);
        if (c->total_blocks > 0) {
            const uint16_t *src = (const uint16_t *)c->d_cnt.p + (size_t)track * c->total_blocks;
            cudaError_t e = cudaMemcpyAsync(raw.data(), src, (size_t)c->total_blocks * 2, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return fail(c, NTL_ERR_CUDA, "ntl_get_window_counts: %s", cudaGetErrorString(e));
        }
    }
    const ntl_read_result *res = (const ntl_read_result *)c->h_results.p;
    const int32_t *h_len = (const int32_t *)((const char *)c->h_meta.p + c->off_len);
    int64_t pos = 0;
    for (int32_t i = 0; i < c->n_reads; i++) {
        const int32_t n = res[i].n_win;
        if (cap > 0 && n > 0) {
            const uint16_t *blk = raw.data() + res[i].win_offset;
            const int64_t nb = blocks_of(c, h_len[i]);
            for (int32_t k = 0; k < n; k++)
                if (pos + k < cap) out[pos + k] = (uint16_t)window_count(blk, k, n, nb, c->dev.Q);
        }
        pos += n;
    }
    return pos;
}

} // namespace

/* ================================================================================================ the public context:
 * one or several devices.  A batch is cut into contiguous shards balanced by bases (one per device); every staged
 * entry point runs on all shards; records are gathered in input order. */
struct ntl_ctx {
    std::vector<ntl_dev_ctx *> dev;
    std::vector<int32_t> bound;             /* shard g = reads [bound[g], bound[g + 1]) of the last batch */
    int32_t n_reads = 0;
    std::vector<ntl_read_result> gathered;  /* n_devices > 1: all records in input order */
    char err[640] = "";
    ntl_timings tm;
    int options = 0;
};

namespace {

int cfail(ntl_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

/* run fn(g) for every shard, in parallel threads when there is more than one; first error wins */
template <class F> int for_shards(ntl_ctx *c, F fn)
{
    const int G = (int)c->dev.size();
    if (G == 1) { const int rc = fn(0); if (rc != NTL_OK) snprintf(c->err, sizeof c->err, "%s", c->dev[0]->err); return rc; }
    std::vector<int> rcs((size_t)G, NTL_OK);
    std::vector<std::thread> th;
    for (int g = 1; g < G; g++) th.emplace_back([&, g]() { rcs[g] = fn(g); });
    rcs[0] = fn(0);
    for (auto &t : th) t.join();
    for (int g = 0; g < G; g++)
        if (rcs[g] != NTL_OK) { snprintf(c->err, sizeof c->err, "device %d: %s", c->dev[g]->device, c->dev[g]->err); return rcs[g]; }
    return NTL_OK;
}

/* contiguous shards balanced by cumulative bases (replaces the round-robin split of NanoTel.R:2242-2243, whose only
 * purpose was load balance; the partition is invisible in the output because Serials are assigned after the gather) */
void shard_bounds(const int64_t *len, int32_t n, int G, std::vector<int32_t> *bound)
{
    bound->assign((size_t)G + 1, n);
    (*bound)[0] = 0;
    int64_t total = 0;
    for (int32_t i = 0; i < n; i++) total += len[i];
    int64_t acc = 0;
    int g = 1;
    for (int32_t i = 0; i < n && g < G; i++) {
        acc += len[i];
        while (g < G && acc * G >= total * g) { (*bound)[g] = i + 1; g++; }
    }
}

void merge_timings(ntl_ctx *c)
{
    ntl_timings t = ntl_timings();
    for (size_t g = 0; g < c->dev.size(); g++) {
        const ntl_timings &s = c->dev[g]->tm;
        t.pack_ms = std::max(t.pack_ms, s.pack_ms); t.h2d_ms = std::max(t.h2d_ms, s.h2d_ms);
        t.filter_ms = std::max(t.filter_ms, s.filter_ms); t.scan_ms = std::max(t.scan_ms, s.scan_ms);
        t.locate_ms = std::max(t.locate_ms, s.locate_ms); t.triage_ms = std::max(t.triage_ms, s.triage_ms);
        t.d2h_ms = std::max(t.d2h_ms, s.d2h_ms); t.total_ms = std::max(t.total_ms, s.total_ms);
        t.bases += s.bases; t.packed_bytes += s.packed_bytes; t.window_bytes += s.window_bytes;
        t.h2d_bytes += s.h2d_bytes; t.d2h_bytes += s.d2h_bytes;
        t.kernel_launches += s.kernel_launches; t.candidates += s.candidates;
        t.scan_is_jit = g == 0 ? s.scan_is_jit : (t.scan_is_jit && s.scan_is_jit);
        t.steps = std::max(t.steps, s.steps);
    }
    const double total = c->tm.total_ms;
    c->tm = t;
    if (total > t.total_ms) c->tm.total_ms = total;
}

int gather_results(ntl_ctx *c, const ntl_read_result **results)
{
    const int G = (int)c->dev.size();
    if (G == 1) { if (results) *results = (const ntl_read_result *)c->dev[0]->h_results.p; return NTL_OK; }
    c->gathered.resize((size_t)c->n_reads + 1);
    for (int g = 0; g < G; g++) {
        const int32_t b = c->bound[g], e = c->bound[g + 1];
        if (e > b) memcpy(&c->gathered[(size_t)b], c->dev[g]->h_results.p, (size_t)(e - b) * sizeof(ntl_read_result));
    }
    if (results) *results = c->gathered.data();
    return NTL_OK;
}

int shard_of(const ntl_ctx *c, int32_t read_idx, int32_t *local)
{
    for (size_t g = 0; g + 1 < c->bound.size(); g++)
        if (read_idx >= c->bound[g] && read_idx < c->bound[g + 1]) { *local = read_idx - c->bound[g]; return (int)g; }
    return -1;
}

} // namespace

/* ============================================================================================== lifecycle */
extern "C" int ntl_version(void) { return NTL_VERSION; }

extern "C" const char *ntl_last_error(const ntl_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int32_t ntl_count_windows(int64_t length, int32_t subseq_length) { return count_windows(length, subseq_length); }

extern "C" int ntl_create(ntl_ctx **out, const ntl_params *p)
{
    g_create_err[0] = 0;
    if (!out || !p) return fail(nullptr, NTL_ERR_ARG, "ntl_create: NULL argument");
    *out = nullptr;
    std::vector<int32_t> ids;
    if (p->n_devices > 0) {
        if (!p->device_ids || p->n_devices > NTL_MAX_DEVICES) return fail(nullptr, NTL_ERR_ARG, "n_devices must be 0..%d with device_ids", NTL_MAX_DEVICES);
        /* a device may be listed more than once: every entry is a shard with its own stream and buffers */
        for (int i = 0; i < p->n_devices; i++) ids.push_back(p->device_ids[i]);
    } else ids.push_back(p->device);

    int nt = p->host_threads;
    if (nt <= 0) nt = (int)std::thread::hardware_concurrency();
    if (nt <= 0) nt = 1;
    if (nt > 256) nt = 256;
    const int per_dev = std::max(1, std::min(64, nt / (int)ids.size()));

    ntl_ctx *c = new (std::nothrow) ntl_ctx();
    if (!c) return fail(nullptr, NTL_ERR_NOMEM, "out of memory");
    c->options = (int)p->options;
    for (size_t g = 0; g < ids.size(); g++) {
        ntl_dev_ctx *d = nullptr;
        const int rc = dev_create(&d, p, ids[g], per_dev);
        if (rc != NTL_OK) { ntl_destroy(c); return rc; }
        c->dev.push_back(d);
    }
    c->bound.assign(ids.size() + 1, 0);
    c->tm = ntl_timings();
    *out = c;
    return NTL_OK;
}

extern "C" void ntl_destroy(ntl_ctx *c)
{
    if (!c) return;
    for (ntl_dev_ctx *d : c->dev) dev_destroy(d);
    delete c;
}

extern "C" int ntl_scan_path(const ntl_ctx *c) { return c && !c->dev.empty() ? c->dev[0]->scan_path : NTL_ERR_ARG; }

extern "C" const char *ntl_scan_path_note(const ntl_ctx *c) { return c && !c->dev.empty() ? c->dev[0]->scan_note.c_str() : ""; }

extern "C" int ntl_device_count(const ntl_ctx *c) { return c ? (int)c->dev.size() : NTL_ERR_ARG; }

extern "C" int ntl_get_geometry(const ntl_ctx *c, int32_t *block, int32_t *blocks_per_window, int32_t *words_per_span,
                                int32_t *blocks_per_span)
{
    if (!c || c->dev.empty()) return NTL_ERR_ARG;
    const ntl_dev_params &d = c->dev[0]->dev;
    if (block) *block = d.SG;
    if (blocks_per_window) *blocks_per_window = d.Q;
    if (words_per_span) *words_per_span = d.W;
    if (blocks_per_span) *blocks_per_span = d.BPS;
    return NTL_OK;
}

/* Diagnostics (no device needed): the host packer on one read. */
extern "C" long ntl_pack_read(const char *seq, int64_t len, int32_t rc, uint32_t *words, int64_t capacity, int32_t *four_bit)
{
    if (!seq || !words || len < 1 || len > (1LL << 30)) return NTL_ERR_ARG;
    const int64_t nw = (len + 31) >> 5;
    if (four_bit) *four_bit = 0;
    if (capacity < nw * 2) return NTL_ERR_NOMEM;
    if (ntl_pack_read_2bit(seq, len, rc ? 1 : 0, words, nw) == 0) { ntl_pack_fence(); return (long)(nw * 2); }
    if (capacity < nw * 4) return NTL_ERR_NOMEM;
    if (ntl_pack_read_4bit(seq, len, rc ? 1 : 0, words, nw) != 0) return NTL_ERR_SEQUENCE;
    if (four_bit) *four_bit = 1;
    return (long)(nw * 4);
}

static int digest_only(const ntl_params *p, ntl_dev_ctx *tmp, char *log, int log_cap)
{
    const int rc = digest_params(tmp, p);
    if (rc != NTL_OK && log && log_cap > 0) snprintf(log, (size_t)log_cap, "%s", tmp->err);
    return rc;
}

/* Diagnostics (no device needed): NVRTC-compile the specialised scan kernel for `arch`, optionally saving the cubin. */
extern "C" long ntl_jit_compile_check(const ntl_params *p, const char *arch, char *log, int log_cap, const char *cubin_path)
{
    if (!p || !arch) return NTL_ERR_ARG;
    ntl_dev_ctx tmp;
    int rc = digest_only(p, &tmp, log, log_cap);
    if (rc != NTL_OK) return rc;
    std::string cubin, lg;
    long n = ntl_jit_compile(&tmp.dev, arch, &cubin, &lg);
    if (log && log_cap > 0) snprintf(log, (size_t)log_cap, "%s", lg.c_str());
    if (n > 0 && cubin_path) {
        FILE *f = fopen(cubin_path, "wb");
        if (f) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
    }
    return n > 0 ? n : NTL_ERR_JIT;
}

extern "C" long ntl_jit_precompile_to(const ntl_params *p, const char *arch, const char *dir, char *log, int log_cap)
{
    if (!p || !arch || !dir) return NTL_ERR_ARG;
    ntl_dev_ctx tmp;
    int rc = digest_only(p, &tmp, log, log_cap);
    if (rc != NTL_OK) return rc;
    std::string lg;
    long n = ntl_jit_precompile(&tmp.dev, arch, dir, &lg);
    if (log && log_cap > 0) snprintf(log, (size_t)log_cap, "%s", lg.c_str());
    return n > 0 ? n : NTL_ERR_JIT;
}

extern "C" long ntl_jit_get_source(const ntl_params *p, char *buf, long cap)
{
    if (!p) return NTL_ERR_ARG;
    ntl_dev_ctx tmp;
    int rc = digest_only(p, &tmp, buf, (int)std::min<long>(cap, 1 << 20));
    if (rc != NTL_OK) return rc;
    if (tmp.dev.BPS <= 0) return NTL_ERR_JIT;
    const std::string s = ntl_jit_source(&tmp.dev);
    if (buf && cap > 0) snprintf(buf, (size_t)cap, "%s", s.c_str());
    return (long)s.size();
}

/* ============================================================================================== staged entry points */
static int pack_shards(ntl_ctx *c, const char *const *seq, const int64_t *len, int32_t n, bool overlap)
{
    if (!seq || !len || n < 0) return cfail(c, NTL_ERR_ARG, "ntl_batch_pack: bad arguments");
    c->n_reads = n;
    shard_bounds(len, n, (int)c->dev.size(), &c->bound);
    return for_shards(c, [&](int g) {
        const int32_t b = c->bound[g], e = c->bound[g + 1];
        return dev_pack(c->dev[g], seq + b, len + b, e - b, overlap);
    });
}

extern "C" int ntl_batch_pack(ntl_ctx *c, const char *const *seq, const int64_t *len, int32_t n)
{
    if (!c) return NTL_ERR_ARG;
    const int rc = pack_shards(c, seq, len, n, false);
    merge_timings(c);
    return rc;
}

extern "C" int ntl_batch_upload(ntl_ctx *c)
{
    if (!c) return NTL_ERR_ARG;
    const int rc = for_shards(c, [&](int g) { return dev_upload(c->dev[g]); });
    merge_timings(c);
    return rc;
}

extern "C" int ntl_batch_enqueue(ntl_ctx *c)
{
    if (!c) return NTL_ERR_ARG;
    for (size_t g = 0; g < c->dev.size(); g++) {          /* launches are asynchronous: no threads needed */
        const int rc = dev_enqueue(c->dev[g]);
        if (rc != NTL_OK) { snprintf(c->err, sizeof c->err, "device %d: %s", c->dev[g]->device, c->dev[g]->err); return rc; }
    }
    return NTL_OK;
}

extern "C" int ntl_batch_wait(ntl_ctx *c)
{
    if (!c) return NTL_ERR_ARG;
    for (size_t g = 0; g < c->dev.size(); g++) {
        const int rc = dev_wait(c->dev[g]);
        if (rc != NTL_OK) { snprintf(c->err, sizeof c->err, "device %d: %s", c->dev[g]->device, c->dev[g]->err); return rc; }
    }
    merge_timings(c);
    return NTL_OK;
}

extern "C" int ntl_batch_run(ntl_ctx *c)
{
    int rc = ntl_batch_enqueue(c);
    if (rc == NTL_OK) rc = ntl_batch_wait(c);
    return rc;
}

extern "C" int ntl_batch_download(ntl_ctx *c, const ntl_read_result **results)
{
    if (!c) return NTL_ERR_ARG;
    int rc = for_shards(c, [&](int g) { return dev_download(c->dev[g]); });
    if (rc == NTL_OK) rc = gather_results(c, results);
    merge_timings(c);
    return rc;
}

extern "C" int ntl_scan_batch(ntl_ctx *c, const char *const *seq, const int64_t *len, int32_t n,
                              const ntl_read_result **results)
{
    if (!c) return NTL_ERR_ARG;
    if (!seq || !len || n < 0) return cfail(c, NTL_ERR_ARG, "ntl_scan_batch: bad arguments");
    const double t0 = now_ms();
    Trace tr;
    c->n_reads = n;
    shard_bounds(len, n, (int)c->dev.size(), &c->bound);
    /* every shard runs its whole pipeline on its own thread, stream and packer threads */
    int rc = for_shards(c, [&](int g) {
        ntl_dev_ctx *d = c->dev[g];
        const int32_t b = c->bound[g], e = c->bound[g + 1];
        int r = dev_pack(d, seq + b, len + b, e - b, /*overlap=*/true);
        if (r == NTL_OK) r = dev_enqueue(d);
        if (r == NTL_OK) r = dev_wait(d);
        if (r == NTL_OK) r = dev_download(d);
        return r;
    });
    tr.mark("shards");
    if (rc == NTL_OK) rc = gather_results(c, results);
    tr.mark("gather");
    c->tm.total_ms = 0;
    merge_timings(c);
    c->tm.total_ms = now_ms() - t0;
    return rc;
}

extern "C" int ntl_scan_batch_concat(ntl_ctx *c, const char *buf, const int64_t *offsets, int32_t n,
                                     const ntl_read_result **results)
{
    if (!c) return NTL_ERR_ARG;
    if (!buf || !offsets || n < 0) return cfail(c, NTL_ERR_ARG, "ntl_scan_batch_concat: bad arguments");
    std::vector<const char *> seq((size_t)n);
    std::vector<int64_t> len((size_t)n);
    for (int32_t i = 0; i < n; i++) { seq[i] = buf + offsets[i]; len[i] = offsets[i + 1] - offsets[i]; }
    return ntl_scan_batch(c, seq.data(), len.data(), n, results);
}

/* R hands over an XStringSet without copying it: one shared pool of bytes + start (1-based) + width per read.  The
 * pool holds Biostrings' internal DNA codes (A 1, C 2, G 4, T 8, IUPAC = OR of the bits, '-' 16, '+' 32, '.' 64) when
 * biostrings_codes != 0, ASCII otherwise. */
extern "C" int ntl_scan_batch_pool(ntl_ctx *c, const unsigned char *pool, const int32_t *start, const int32_t *width,
                                   int32_t n, int32_t biostrings_codes, const ntl_read_result **results)
{
    if (!c) return NTL_ERR_ARG;
    if (!pool || !start || !width || n < 0) return cfail(c, NTL_ERR_ARG, "ntl_scan_batch_pool: bad arguments");
    std::vector<const char *> seq((size_t)n);
    std::vector<int64_t> len((size_t)n);
    std::vector<char> ascii;
    if (!biostrings_codes) {
        for (int32_t i = 0; i < n; i++) { seq[i] = (const char *)pool + (start[i] - 1); len[i] = width[i]; }
    } else {
        /* decode into one ASCII buffer (multi-threaded table lookup); the pool itself is never modified */
        static const char dec[17] = {'-', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N', '-'};
        std::vector<int64_t> off((size_t)n + 1, 0);
        for (int32_t i = 0; i < n; i++) off[i + 1] = off[i] + (width[i] > 0 ? width[i] : 0);
        ascii.resize((size_t)off[n] + 1);
        std::atomic<int> bad(-1);
        ntl_parallel_for(n, c->dev[0]->host_threads * (int)c->dev.size(), 64, [&](int64_t b, int64_t e) {
            for (int64_t i = b; i < e; i++) {
                const unsigned char *s = pool + (start[i] - 1);
                char *d = ascii.data() + off[i];
                for (int64_t j = 0; j < width[i]; j++) {
                    const unsigned char v = s[j];
                    if (v <= 16) d[j] = dec[v]; else if (v == 32) d[j] = '+'; else if (v == 64) d[j] = '.'; else { d[j] = '?'; bad.store((int)i); }
                }
            }
        });
        if (bad.load() >= 0) return cfail(c, NTL_ERR_SEQUENCE, "read %d holds a byte that is not a Biostrings DNA code", bad.load());
        for (int32_t i = 0; i < n; i++) { seq[i] = ascii.data() + off[i]; len[i] = width[i]; }
    }
    return ntl_scan_batch(c, seq.data(), len.data(), n, results);
}

extern "C" int ntl_get_timings(const ntl_ctx *c, ntl_timings *out)
{
    if (!c || !out) return NTL_ERR_ARG;
    *out = c->tm;
    return NTL_OK;
}

extern "C" void *ntl_stream(const ntl_ctx *c) { return c && !c->dev.empty() ? (void *)c->dev[0]->stream : nullptr; }

extern "C" int ntl_get_shards(const ntl_ctx *c, int32_t *bounds, int32_t *devices, int32_t cap)
{
    if (!c) return NTL_ERR_ARG;
    const int G = (int)c->dev.size();
    for (int g = 0; g <= G && g <= cap; g++) if (bounds) bounds[g] = c->bound[g];
    for (int g = 0; g < G && g < cap; g++) if (devices) devices[g] = c->dev[g]->device;
    return G;
}

/* ============================================================================================== window tables */
extern "C" int ntl_get_windows(const ntl_ctx *c, int32_t read_idx, int32_t track, int32_t cap, int32_t *start_index,
                               int32_t *end_index, int32_t *covered, double *density)
{
    if (!c) return NTL_ERR_ARG;
    ntl_ctx *mc = const_cast<ntl_ctx *>(c);
    int32_t local = 0;
    const int g = shard_of(c, read_idx, &local);
    if (g < 0) return cfail(mc, NTL_ERR_ARG, "ntl_get_windows: read out of range");
    const int rc = dev_get_windows(c->dev[g], local, track, cap, start_index, end_index, covered, density);
    if (rc < 0) snprintf(mc->err, sizeof mc->err, "%s", c->dev[g]->err);
    return rc;
}

extern "C" int64_t ntl_get_window_counts(const ntl_ctx *c, int32_t track, uint16_t *out, int64_t cap)
{
    if (!c) return NTL_ERR_ARG;
    ntl_ctx *mc = const_cast<ntl_ctx *>(c);
    int64_t pos = 0;
    for (size_t g = 0; g < c->dev.size(); g++) {
        const int64_t left = cap > pos ? cap - pos : 0;
        const int64_t k = dev_get_window_counts(c->dev[g], track, left > 0 ? out + pos : nullptr, left);
        if (k < 0) { snprintf(mc->err, sizeof mc->err, "%s", c->dev[g]->err); return k; }
        pos += k;
    }
    return pos;
}

extern "C" int ntl_get_stages(const ntl_ctx *c, int32_t read_idx, int32_t track, ntl_stage *out)
{
    if (!c || !out) return NTL_ERR_ARG;
    ntl_ctx *mc = const_cast<ntl_ctx *>(c);
    int32_t local = 0;
    const int g = shard_of(c, read_idx, &local);
    if (g < 0) return cfail(mc, NTL_ERR_ARG, "ntl_get_stages: read out of range");
    const ntl_dev_ctx *d = c->dev[g];
    if (!d->dev.debug_stages) return cfail(mc, NTL_ERR_STATE, "context was created without NTL_OPT_DEBUG_STAGES");
    if (d->state < ST_DOWNLOADED) return cfail(mc, NTL_ERR_STATE, "ntl_get_stages before the batch was downloaded");
    if (track < 0 || track >= d->dev.n_tracks) return cfail(mc, NTL_ERR_ARG, "ntl_get_stages: track out of range");
    *out = ((const ntl_stage *)d->h_stages.p)[(size_t)local * 3 + track];
    return NTL_OK;
}

/* ============================================================================================== Serial logic
 * search_patterns' counter (NanoTel.R:2050-2069) under the 8-way round-robin split (NanoTel.R:2234-2258). */
extern "C" int ntl_assign_serials(const ntl_read_result *res, int32_t n_reads, int32_t serial_start, int32_t *serial,
                                  int32_t *row_order, int32_t *next_serial_start)
{
    if (!res || !serial || !row_order || n_reads < 0) return NTL_ERR_ARG;
    std::vector<int32_t> idx;                                  /* the chunk after filter_reads subset it (:2154) */
    idx.reserve((size_t)n_reads);
    for (int32_t i = 0; i < n_reads; i++) {
        serial[i] = 0;
        if (!(res[i].status & NTL_READ_FILTERED)) idx.push_back(i);
    }
    const int32_t n = (int32_t)idx.size();
    int32_t rows = 0, mx = 0;
    if (n < 8) {                                               /* :2236-2239 sequential branch */
        int32_t cur = serial_start;
        for (int32_t j = 0; j < n; j++)
            if (res[idx[j]].status & NTL_READ_KEEP) { serial[idx[j]] = cur; row_order[rows++] = idx[j]; mx = cur; cur++; }
    } else {                                                   /* :2242-2254 group g takes reads g, g+8, ... */
        int32_t offset = 0;
        for (int32_t g = 0; g < 8; g++) {
            int32_t cur = serial_start + offset, size = 0;
            for (int32_t j = g; j < n; j += 8) {
                size++;
                if (res[idx[j]].status & NTL_READ_KEEP) {
                    serial[idx[j]] = cur; row_order[rows++] = idx[j];
                    if (cur > mx) mx = cur;
                    cur++;
                }
            }
            offset += size;
        }
    }
    /* :2258 serial_start <- max(df_summary$Serial) + 1.  With no row at all R computes -Inf (reference bug,
     * SURVEY A.11): documented in DESIGN.md, not imitated -- the start is left unchanged. */
    if (next_serial_start) *next_serial_start = rows > 0 ? std::max(mx + 1, serial_start) : serial_start;
    return rows;
}
