/*
 * ntl_api.cpp -- the C ABI of libnanotel_b200.so (include/nanotel_b200.h): context, pinned/device buffers, the
 * batch pipeline  pack -> H2D -> [filter] -> scan -> locate -> D2H,  and the host-side Serial logic.
 *
 * Replaces the body of NanoTel.R's chunk loop (NanoTel.R:2219-2258) for one --nrec chunk per call.
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
cudaError_t ntl_k_scan(const ntl_scan_args *a, int four_bit, int grid, cudaStream_t st);
cudaError_t ntl_k_scan_occupancy(int *blocks_per_sm);
cudaError_t ntl_k_filter(const ntl_read_args *a, cudaStream_t st);
cudaError_t ntl_k_triage(const ntl_read_args *a, cudaStream_t st);
cudaError_t ntl_k_locate(const ntl_read_args *a, int grid, cudaStream_t st);
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
        cudaError_t e = cudaHostAlloc(&np, ncap, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        if (p) { if (keep) memcpy(np, p, cap); cudaFreeHost(p); }
        p = np; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = bytes + bytes / 4 + 4096;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, ncap);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

#define NTL_EVENT_RING 256
enum { ST_EMPTY = 0, ST_PACKED = 1, ST_UPLOADED = 2, ST_RAN = 3, ST_DOWNLOADED = 4 };

} // namespace

struct ntl_ctx {
    ntl_params prm;
    std::vector<std::string> pat_store, tvr_store;
    ntl_dev_params dev;
    int device = 0, n_sms = 0, scan_grid = 0, scan_grid4 = 0, locate_grid = 0, host_threads = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr};
    cudaEvent_t ring[NTL_EVENT_RING][5] = {{nullptr}};
    int pending = 0, pending_launches = 0;
    char err[512] = "";
    ntl_jit_kernel *jit = nullptr;

    int state = ST_EMPTY;
    int32_t n_reads = 0, n2 = 0, n4 = 0;
    int64_t total_words = 0, total_windows = 0, bases = 0;
    size_t dens_offset = 0;
    size_t meta_bytes = 0, off_len = 0, off_woff = 0, off_winoff = 0, off_order = 0, off_fmt = 0;

    PinnedBuf h_packed, h_meta, h_results, h_cum, h_stages, h_list;
    DevBuf d_packed, d_meta, d_results, d_cum, d_stages, d_pass, d_counter, d_thr, d_flags, d_kept, d_list;
    std::vector<int64_t> kept_off;          /* per read: first element of its rows in h_cum, -1 = not downloaded */
    std::vector<uint16_t> win_scratch;      /* ntl_get_windows of a read that was not kept: fetched on demand     */
    ntl_timings tm;
};

namespace {

int fail(ntl_ctx *c, int code, const char *fmt, ...)
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

/* --patterns / --tvr_patterns -> device pattern table (unique(), NanoTel.R:328,362; sorted by length so that
 * patterns of equal length share one dilation in K2) */
int digest_patterns(ntl_ctx *c, const std::vector<std::string> &in, ntl_dev_pat *out, int32_t *n_out,
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

/* argument checks of ntl_create + the device parameter block; no CUDA call */
int digest_params(ntl_ctx *c, const ntl_params *p)
{
    if (p->n_patterns < 1 || p->n_patterns > NTL_MAX_PATTERNS || !p->patterns)
        return fail(c, NTL_ERR_ARG, "n_patterns must be 1..%d", NTL_MAX_PATTERNS);
    if (p->n_tvr < 0 || p->n_tvr > NTL_MAX_PATTERNS || (p->n_tvr > 0 && !p->tvr_patterns))
        return fail(c, NTL_ERR_ARG, "n_tvr must be 0..%d", NTL_MAX_PATTERNS);
    if (p->subseq_length < 1 || p->subseq_length > 65535)
        return fail(c, NTL_ERR_ARG, "subseq_length must be 1..65535");
    if (!(p->min_density == p->min_density)) return fail(c, NTL_ERR_ARG, "min_density is NaN");
    c->prm = *p;
    for (int i = 0; i < p->n_patterns; i++) {
        if (!p->patterns[i]) return fail(c, NTL_ERR_ARG, "NULL pattern");
        c->pat_store.push_back(p->patterns[i]);
    }
    for (int i = 0; i < p->n_tvr; i++) {
        if (!p->tvr_patterns[i]) return fail(c, NTL_ERR_ARG, "NULL tvr pattern");
        c->tvr_store.push_back(p->tvr_patterns[i]);
    }
    c->prm.patterns = nullptr; c->prm.tvr_patterns = nullptr;

    ntl_dev_params &d = c->dev;
    memset(&d, 0, sizeof d);
    int rc = digest_patterns(c, c->pat_store, d.main_pat, &d.n_main, &d.n_main_groups, d.main_group_begin);
    if (rc == NTL_OK) rc = digest_patterns(c, c->tvr_store, d.tvr_pat, &d.n_tvr, &d.n_tvr_groups, d.tvr_group_begin);
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
    return NTL_OK;
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
    ntl_ctx *c = new (std::nothrow) ntl_ctx();
    if (!c) return fail(nullptr, NTL_ERR_NOMEM, "out of memory");
    int rc = digest_params(c, p);
    if (rc != NTL_OK) { strncpy(g_create_err, c->err, sizeof g_create_err); delete c; return rc; }

    int nt = p->host_threads;
    if (nt <= 0) nt = (int)std::thread::hardware_concurrency();
    if (nt <= 0) nt = 1;
    if (nt > 64) nt = 64;
    c->host_threads = nt;

    /* ---- device */
    cudaError_t e = cudaSetDevice(p->device);
    if (e != cudaSuccess) {
        fail(nullptr, NTL_ERR_CUDA, "cudaSetDevice(%d) failed: %s -- libnanotel_b200 has no CPU fallback", p->device,
             cudaGetErrorString(e));
        delete c;
        return NTL_ERR_CUDA;
    }
    c->device = p->device;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, p->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreate(&c->ev[i]);
    if (e == cudaSuccess) e = ntl_k_set_params(&c->dev, c->stream);
    int bps = 0;
    if (e == cudaSuccess) e = ntl_k_scan_occupancy(&bps);
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
        ntl_destroy(c);
        return NTL_ERR_CUDA;
    }
    c->n_sms = prop.multiProcessorCount;
    if (bps < 1) bps = 1;
    c->scan_grid = c->n_sms * bps;                             /* persistent grid: every CTA resident, 148 x occupancy */
    { int lb = 0; if (ntl_k_locate_occupancy(&lb) != cudaSuccess || lb < 1) lb = 4; c->locate_grid = c->n_sms * lb; }

    /* ---- NVRTC specialisation of the scan kernel for this pattern set */
    if (!(p->options & NTL_OPT_NO_JIT)) {
        std::string jerr;
        c->jit = ntl_jit_build(&c->dev, prop.major, prop.minor, &jerr);
        if (!c->jit && (p->options & NTL_OPT_REQUIRE_JIT)) {
            fail(nullptr, NTL_ERR_JIT, "NVRTC specialisation failed: %s", jerr.c_str());
            ntl_destroy(c);
            return NTL_ERR_JIT;
        }
        if (c->jit) {
            int jb = ntl_jit_blocks_per_sm(c->jit, 0);
            if (jb >= 1) c->scan_grid = c->n_sms * jb;
            jb = ntl_jit_blocks_per_sm(c->jit, 1);
            c->scan_grid4 = c->n_sms * (jb >= 1 ? jb : 1);
        } else {
            snprintf(c->err, sizeof c->err, "note: JIT unavailable (%s); using the runtime-pattern scan kernel", jerr.c_str());
        }
    }
    *out = c;
    return NTL_OK;
}

/* Diagnostics (no device needed): the host packer on one read. */
extern "C" long ntl_pack_read(const char *seq, int64_t len, int32_t rc, uint32_t *words, int64_t capacity, int32_t *four_bit)
{
    if (!seq || !words || len < 1 || len > (1LL << 30)) return NTL_ERR_ARG;
    const int64_t quads = ((((len >> 5) + 1) + 3) >> 2);
    if (four_bit) *four_bit = 0;
    if (capacity < quads * 8) return NTL_ERR_NOMEM;
    if (ntl_pack_read_2bit(seq, len, rc ? 1 : 0, words) == 0) return (long)(quads * 8);
    if (capacity < quads * 16) return NTL_ERR_NOMEM;
    if (ntl_pack_read_4bit(seq, len, rc ? 1 : 0, words) != 0) return NTL_ERR_SEQUENCE;
    if (four_bit) *four_bit = 1;
    return (long)(quads * 16);
}

/* Diagnostics (no device needed): NVRTC-compile the specialised scan kernel for `arch`, optionally saving the cubin. */
extern "C" long ntl_jit_compile_check(const ntl_params *p, const char *arch, char *log, int log_cap, const char *cubin_path)
{
    if (!p || !arch) return NTL_ERR_ARG;
    ntl_ctx tmp;
    int rc = digest_params(&tmp, p);
    if (rc != NTL_OK) { if (log && log_cap > 0) snprintf(log, (size_t)log_cap, "%s", tmp.err); return rc; }
    std::string cubin, lg;
    long n = ntl_jit_compile(&tmp.dev, arch, &cubin, &lg);
    if (log && log_cap > 0) snprintf(log, (size_t)log_cap, "%s", lg.c_str());
    if (n > 0 && cubin_path) {
        FILE *f = fopen(cubin_path, "wb");
        if (f) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
    }
    return n > 0 ? n : NTL_ERR_JIT;
}

extern "C" void ntl_destroy(ntl_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->jit) ntl_jit_free(c->jit);
    c->h_packed.release(); c->h_meta.release(); c->h_results.release(); c->h_cum.release(); c->h_stages.release();
    c->h_list.release(); c->d_kept.release(); c->d_list.release();
    c->d_packed.release(); c->d_meta.release(); c->d_results.release(); c->d_cum.release(); c->d_stages.release();
    c->d_pass.release(); c->d_counter.release(); c->d_thr.release(); c->d_flags.release();
    for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int k = 0; k < NTL_EVENT_RING; k++)
        for (int i = 0; i < 5; i++) if (c->ring[k][i]) cudaEventDestroy(c->ring[k][i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

/* ============================================================================================== pack */
static int ensure_device_buffers(ntl_ctx *c, int64_t packed_words)
{
    const int32_t n = c->n_reads;
    const int T = c->dev.n_tracks;
    CK(c, c->d_packed.ensure((size_t)packed_words * 4 + 64));
    CK(c, c->d_meta.ensure(c->meta_bytes + 16));
    CK(c, c->d_results.ensure((size_t)n * sizeof(ntl_read_result) + 64));
    CK(c, c->d_cum.ensure((size_t)c->total_windows * 2 * T + 64));
    CK(c, c->d_pass.ensure((size_t)n + 64));
    CK(c, c->d_flags.ensure((size_t)n * 20 + 128));       /* candidate list + per-candidate join state of the locate kernel */
    if (c->dev.debug_stages) CK(c, c->d_stages.ensure((size_t)n * 3 * sizeof(ntl_stage) + 64));
    return NTL_OK;
}

/* overlap = true: the packed words are copied to the device in 8 MiB pieces while the remaining reads are still
 * being packed (the calling thread issues the copies between its own grains), so that PCIe time hides behind the
 * packer; the batch ends up in state UPLOADED. */
static int pack_internal(ntl_ctx *c, const char *const *seq, const int64_t *len, int32_t n, bool overlap)
{
    if (!c) return NTL_ERR_ARG;
    if (!seq || !len || n < 0) return fail(c, NTL_ERR_ARG, "ntl_batch_pack: bad arguments");
    CK(c, cudaSetDevice(c->device));
    const double t0 = now_ms();
    Trace tr;
    c->state = ST_EMPTY;
    c->n_reads = n;
    const int32_t S = c->dev.S;
    const int rcflag = c->prm.rc ? 1 : 0;

    /* ---- table layout inside one pinned block (one H2D copy) */
    size_t off = 0;
    c->off_len = off;    off += ((size_t)n * 4 + 15) & ~(size_t)15;
    c->off_woff = off;   off += (size_t)n * 8;
    c->off_winoff = off; off += (size_t)n * 8;
    c->off_order = off;  off += ((size_t)n * 4 + 15) & ~(size_t)15;
    c->off_fmt = off;    off += ((size_t)n + 15) & ~(size_t)15;
    c->meta_bytes = off;
    CK(c, c->h_meta.ensure(off + 16));
    char *mb = (char *)c->h_meta.p;
    int32_t *h_len = (int32_t *)(mb + c->off_len);
    int64_t *h_woff = (int64_t *)(mb + c->off_woff);
    int64_t *h_winoff = (int64_t *)(mb + c->off_winoff);
    int32_t *h_order = (int32_t *)(mb + c->off_order);
    uint8_t *h_fmt = (uint8_t *)(mb + c->off_fmt);

    /* what the packer needs (lengths, word offsets) first; the window offsets, which cost a division per read, are
     * computed by win_tables() -- in overlap mode while the worker threads are already packing */
    int64_t words = 0, bases = 0;
    for (int32_t i = 0; i < n; i++) {
        const int64_t L = len[i];
        if (L < 1) return fail(c, NTL_ERR_SEQUENCE, "read %d has length %lld: NanoTel.R stops on empty reads (seq(1, 0, by = S), :216)", i, (long long)L);
        if (L > (1LL << 30)) return fail(c, NTL_ERR_SEQUENCE, "read %d is longer than 2^30 bases", i);
        if (!seq[i]) return fail(c, NTL_ERR_ARG, "read %d: NULL sequence", i);
        h_len[i] = (int32_t)L;
        h_woff[i] = words;
        h_fmt[i] = 0;
        const int64_t n_words = (L >> 5) + 1;
        words += ((n_words + 3) >> 2) * 8;
        bases += L;
    }
    auto win_tables = [&]() {
        int64_t wins = 0;
        for (int32_t i = 0; i < n; i++) {
            h_winoff[i] = wins;
            wins += ((int64_t)count_windows(h_len[i], S) + 7) & ~(int64_t)7;   /* each read starts on a 16-byte boundary of the uint16 planes */
        }
        c->total_windows = wins;
    };
    if (!overlap) win_tables();
    tr.mark("tables");
    const int64_t main_words = words;
    c->bases = bases;
    CK(c, c->h_packed.ensure((size_t)main_words * 4 + 64));
    uint32_t *hp = (uint32_t *)c->h_packed.p;

    /* ---- work order: longest reads first (by 4096-position steps), 2-bit reads then 4-bit reads */
    auto work_order = [&]() {
        int32_t maxc = 0;
        std::vector<int32_t> chunks(n);
        for (int32_t i = 0; i < n; i++) {
            const int64_t nq = ((((int64_t)h_len[i] >> 5) + 1) + 3) >> 2;
            chunks[i] = (int32_t)((nq + 31) >> 5);
            if (chunks[i] > maxc) maxc = chunks[i];
        }
        std::vector<int64_t> cnt2((size_t)maxc + 2, 0), cnt4((size_t)maxc + 2, 0);
        int32_t n2 = 0, n4 = 0;
        for (int32_t i = 0; i < n; i++) { if (h_fmt[i]) { cnt4[chunks[i]]++; n4++; } else { cnt2[chunks[i]]++; n2++; } }
        /* descending: position of bucket k = number of reads with more chunks */
        std::vector<int64_t> pos2((size_t)maxc + 2, 0), pos4((size_t)maxc + 2, 0);
        int64_t a2 = 0, a4 = 0;
        for (int32_t k = maxc; k >= 0; k--) { pos2[k] = a2; a2 += cnt2[k]; pos4[k] = a4; a4 += cnt4[k]; }
        for (int32_t i = 0; i < n; i++) {
            if (h_fmt[i]) h_order[n2 + pos4[chunks[i]]++] = i;
            else h_order[pos2[chunks[i]]++] = i;
        }
        c->n2 = n2; c->n4 = n4;
    };

    /* ---- 2-bit packing, all host threads; reads with other letters are queued for the 4-bit arena */
    std::vector<int32_t> iupac;
    std::mutex mu;
    auto pack_range = [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; i++) {
            if (ntl_pack_read_2bit(seq[i], len[i], rcflag, hp + h_woff[i]) != 0) {
                std::lock_guard<std::mutex> g(mu);
                iupac.push_back((int32_t)i);
            }
        }
    };
    int64_t up_words = 0;                      /* words already handed to cudaMemcpyAsync (overlap mode) */
    if (!overlap) {
        ntl_parallel_for(n, c->host_threads, 64, pack_range);
    } else {
        c->meta_bytes = off;
        const int64_t grain = 64, ngr = (n + grain - 1) / grain;
        const int64_t piece = 2 << 20;          /* 8 MiB */
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
        win_tables();                           /* the packers do not need these */
        work_order();                           /* h_fmt is all zero here; redone below if a read needs 4 bits */
        cudaError_t cerr = cudaSuccess;
        {
            const int rc0 = ensure_device_buffers(c, main_words);      /* needs total_windows */
            if (rc0 != NTL_OK || cudaEventRecord(c->ev[0], c->stream) != cudaSuccess) {
                next.store(ngr);                                        /* stop the packers, then report */
                for (auto &t : th) t.join();
                return rc0 != NTL_OK ? rc0 : fail(c, NTL_ERR_CUDA, "cudaEventRecord failed");
            }
        }
        cerr = cudaMemcpyAsync(c->d_meta.p, c->h_meta.p, c->meta_bytes, cudaMemcpyHostToDevice, c->stream);
        int64_t uf = 0;
        for (;;) {
            while (uf < ngr && done[uf].load(std::memory_order_acquire)) uf++;
            const int64_t fw = uf == ngr ? main_words : h_woff[uf * grain];
            if (cerr == cudaSuccess && (fw - up_words >= piece || (uf == ngr && fw > up_words))) {
                cerr = cudaMemcpyAsync((uint32_t *)c->d_packed.p + up_words, hp + up_words, (size_t)(fw - up_words) * 4,
                                       cudaMemcpyHostToDevice, c->stream);
                up_words = fw;
            }
            if (uf == ngr) break;
            const int64_t g = next.fetch_add(1);
            if (g < ngr) do_grain(g);
            else std::this_thread::yield();
        }
        for (auto &t : th) t.join();
        if (cerr != cudaSuccess) return fail(c, NTL_ERR_CUDA, "cudaMemcpyAsync (packed reads) failed: %s", cudaGetErrorString(cerr));
    }
    tr.mark("pack 2-bit (+copies)");
    if (!iupac.empty()) {
        std::sort(iupac.begin(), iupac.end());
        std::vector<int64_t> aoff(iupac.size());
        for (size_t k = 0; k < iupac.size(); k++) {
            const int64_t L = len[iupac[k]];
            aoff[k] = words;
            words += ((((L >> 5) + 1) + 3) >> 2) * 16;
        }
        CK(c, c->h_packed.ensure((size_t)words * 4 + 64, /*keep=*/true));
        hp = (uint32_t *)c->h_packed.p;
        int bad = -1;
        ntl_parallel_for((int64_t)iupac.size(), c->host_threads, 4, [&](int64_t b, int64_t e) {
            for (int64_t k = b; k < e; k++) {
                const int32_t i = iupac[k];
                if (ntl_pack_read_4bit(seq[i], len[i], rcflag, hp + aoff[k]) != 0) bad = i;
                h_woff[i] = aoff[k];
                h_fmt[i] = 1;
            }
        });
        if (bad >= 0) return fail(c, NTL_ERR_SEQUENCE, "read %d holds a letter outside the DNA alphabet", bad);
    }
    c->total_words = words;
    tr.mark("pack 4-bit");

    /* overlap mode built the order and sent the tables while the workers were packing, assuming 2-bit reads only */
    const bool tables_sent = overlap && iupac.empty();
    if (!tables_sent) work_order();
    tr.mark("work order");
    c->tm = ntl_timings();
    c->tm.pack_ms = now_ms() - t0;
    c->tm.bases = bases;
    c->tm.packed_bytes = words * 4;
    c->tm.window_bytes = c->total_windows * 2 * c->dev.n_tracks;
    c->state = ST_PACKED;
    if (overlap) {
        /* the 4-bit arena (if any) and the tables follow; a grown device buffer means starting the copy over */
        void *before = c->d_packed.p;
        int rc1 = ensure_device_buffers(c, words);
        if (rc1 != NTL_OK) return rc1;
        if (c->d_packed.p != before) up_words = 0;
        if (words > up_words)
            CK(c, cudaMemcpyAsync((uint32_t *)c->d_packed.p + up_words, (uint32_t *)c->h_packed.p + up_words,
                                  (size_t)(words - up_words) * 4, cudaMemcpyHostToDevice, c->stream));
        if (c->meta_bytes > 0 && !tables_sent)
            CK(c, cudaMemcpyAsync(c->d_meta.p, c->h_meta.p, c->meta_bytes, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaEventRecord(c->ev[1], c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        CK(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
        c->tm.h2d_ms = ms;                       /* first copy issued -> last copy done; overlaps pack_ms */
        c->tm.h2d_bytes = words * 4 + (int64_t)c->meta_bytes;
        c->tm.pack_ms = now_ms() - t0;
        tr.mark("tail copies + sync");
        c->state = ST_UPLOADED;
    }
    return NTL_OK;
}

extern "C" int ntl_batch_pack(ntl_ctx *c, const char *const *seq, const int64_t *len, int32_t n)
{
    return pack_internal(c, seq, len, n, false);
}

/* ============================================================================================== upload */
extern "C" int ntl_batch_upload(ntl_ctx *c)
{
    if (!c) return NTL_ERR_ARG;
    if (c->state < ST_PACKED) return fail(c, NTL_ERR_STATE, "ntl_batch_upload before ntl_batch_pack");
    CK(c, cudaSetDevice(c->device));
    { int rc0 = ensure_device_buffers(c, c->total_words); if (rc0 != NTL_OK) return rc0; }
    CK(c, cudaEventRecord(c->ev[0], c->stream));
    if (c->total_words > 0)
        CK(c, cudaMemcpyAsync(c->d_packed.p, c->h_packed.p, (size_t)c->total_words * 4, cudaMemcpyHostToDevice, c->stream));
    if (c->meta_bytes > 0)
        CK(c, cudaMemcpyAsync(c->d_meta.p, c->h_meta.p, c->meta_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaEventRecord(c->ev[1], c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    c->tm.h2d_ms = ms;
    c->tm.h2d_bytes = c->total_words * 4 + (int64_t)c->meta_bytes;
    c->state = ST_UPLOADED;
    return NTL_OK;
}

static void fill_read_args(const ntl_ctx *c, ntl_read_args *out)
{
    const int32_t n = c->n_reads;
    const int T = c->dev.n_tracks;
    char *dm = (char *)c->d_meta.p;
    ntl_read_args ra;
    memset(&ra, 0, sizeof ra);
    ra.packed = (const uint32_t *)c->d_packed.p;
    ra.len = (const int32_t *)(dm + c->off_len);
    ra.woff = (const int64_t *)(dm + c->off_woff);
    ra.win_off = (const int64_t *)(dm + c->off_winoff);
    ra.fmt = (const uint8_t *)(dm + c->off_fmt);
    ra.pass = c->dev.use_filter ? (uint8_t *)c->d_pass.p : nullptr;
    for (int t = 0; t < 3; t++) ra.cum[t] = t < T ? (const uint16_t *)c->d_cum.p + (size_t)t * c->total_windows : nullptr;
    ra.results = c->d_results.p;
    ra.thr = (const uint16_t *)c->d_thr.p;
    ra.dens = (const double *)((const char *)c->d_thr.p + c->dens_offset);
    ra.order = (const int32_t *)(dm + c->off_order);
    ra.cand = (int32_t *)c->d_flags.p;
    ra.cand_state = (int32_t *)c->d_flags.p + (((size_t)n + 3) & ~(size_t)3);
    ra.counters = (uint32_t *)c->d_counter.p + 4;
    ra.stages = c->dev.debug_stages ? c->d_stages.p : nullptr;
    ra.n_reads = n;
    *out = ra;
}

/* ============================================================================================== run */
/* Enqueue one pass of the hot path (filter, scan, locate) on the context stream without waiting.  Event quads live
 * in a ring so that a timed loop of back-to-back passes still yields per-kernel device times. */
extern "C" int ntl_batch_enqueue(ntl_ctx *c)
{
    if (!c) return NTL_ERR_ARG;
    if (c->state < ST_UPLOADED) return fail(c, NTL_ERR_STATE, "ntl_batch_enqueue before ntl_batch_upload");
    CK(c, cudaSetDevice(c->device));
    if (c->pending >= NTL_EVENT_RING) return fail(c, NTL_ERR_STATE, "more than %d passes enqueued without ntl_batch_wait", NTL_EVENT_RING);
    const int32_t n = c->n_reads;
    const int T = c->dev.n_tracks;
    char *dm = (char *)c->d_meta.p;
    cudaEvent_t *ev = c->ring[c->pending];
    for (int i = 0; i < 5; i++)
        if (!ev[i]) CK(c, cudaEventCreate(&ev[i]));

    ntl_read_args ra;
    fill_read_args(c, &ra);

    ntl_scan_args sa;
    memset(&sa, 0, sizeof sa);
    sa.packed = ra.packed; sa.len = ra.len; sa.woff = ra.woff; sa.win_off = ra.win_off;
    sa.pass = ra.pass;
    for (int t = 0; t < 3; t++) sa.cum[t] = t < T ? (uint16_t *)c->d_cum.p + (size_t)t * c->total_windows : nullptr;

    int launches = 0;
    CK(c, cudaMemsetAsync(c->d_counter.p, 0, 64, c->stream));
    CK(c, cudaEventRecord(ev[0], c->stream));
    if (c->dev.use_filter && n > 0) { CK(c, ntl_k_filter(&ra, c->stream)); launches++; }
    CK(c, cudaEventRecord(ev[1], c->stream));
    c->tm.scan_is_jit = 0;
    if (c->n2 > 0) {
        sa.order = (const int32_t *)(dm + c->off_order);
        sa.n_items = c->n2;
        sa.counter = (uint32_t *)c->d_counter.p;
        if (c->jit) {
            cudaError_t e = ntl_jit_launch(c->jit, &sa, 0, c->scan_grid, c->stream);
            if (e != cudaSuccess) return fail(c, NTL_ERR_CUDA, "JIT scan kernel launch failed: %s", cudaGetErrorString(e));
            c->tm.scan_is_jit = 1;
        } else {
            CK(c, ntl_k_scan(&sa, 0, c->scan_grid, c->stream));
        }
        launches++;
    }
    if (c->n4 > 0) {
        sa.order = (const int32_t *)(dm + c->off_order) + c->n2;
        sa.n_items = c->n4;
        sa.counter = (uint32_t *)c->d_counter.p + 8;
        if (c->jit) {
            cudaError_t e = ntl_jit_launch(c->jit, &sa, 1, c->scan_grid4, c->stream);
            if (e != cudaSuccess) return fail(c, NTL_ERR_CUDA, "JIT scan kernel (IUPAC reads) launch failed: %s", cudaGetErrorString(e));
            c->tm.scan_is_jit = 1;
        } else {
            CK(c, ntl_k_scan(&sa, 1, c->scan_grid, c->stream));
        }
        launches++;
    }
    CK(c, cudaEventRecord(ev[2], c->stream));
    if (n > 0) {
        CK(c, ntl_k_triage(&ra, c->stream)); launches++;
        if (c->pending == 0) CK(c, cudaEventRecord(ev[4], c->stream));   /* first pass only: an event between two kernels costs a few us */
        CK(c, ntl_k_locate(&ra, c->locate_grid, c->stream)); launches++;
    } else if (c->pending == 0) CK(c, cudaEventRecord(ev[4], c->stream));
    CK(c, cudaEventRecord(ev[3], c->stream));
    c->pending_launches += launches;
    c->pending++;
    return NTL_OK;
}

/* Wait for the enqueued passes; timings hold the SUMS over those passes, tm.steps their number. */
extern "C" int ntl_batch_wait(ntl_ctx *c)
{
    if (!c) return NTL_ERR_ARG;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    double f = 0, s = 0, l = 0, g = 0;
    for (int k = 0; k < c->pending; k++) {
        float ms = 0.f;
        cudaEvent_t *ev = c->ring[k];
        CK(c, cudaEventElapsedTime(&ms, ev[0], ev[1])); f += ms;
        CK(c, cudaEventElapsedTime(&ms, ev[1], ev[2])); s += ms;
        CK(c, cudaEventElapsedTime(&ms, ev[2], ev[3])); l += ms;
        if (k == 0) { CK(c, cudaEventElapsedTime(&ms, ev[2], ev[4])); g = ms; }
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

extern "C" int ntl_batch_run(ntl_ctx *c)
{
    int rc = ntl_batch_enqueue(c);
    if (rc == NTL_OK) rc = ntl_batch_wait(c);
    return rc;
}

/* ============================================================================================== download */
/* Results come back in two steps: the 64-byte records (and the debug stages) first; then, for the reads the keep rule
 * retained -- the only ones whose window tables the caller needs (NanoTel.R:1876-1918) -- the window prefixes, gathered
 * on the device into one contiguous block.  The tables of all other reads stay on the device until the next batch
 * and are fetched on demand by ntl_get_windows(). */
extern "C" int ntl_batch_download(ntl_ctx *c, const ntl_read_result **results)
{
    if (!c) return NTL_ERR_ARG;
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

    const ntl_read_result *res = (const ntl_read_result *)c->h_results.p;
    c->kept_off.assign((size_t)n, -1);
    int64_t n_kept = 0, elems = 0;
    /* one pass over the records; a kept read went through the locate kernel, so the candidate count bounds the list */
    const int64_t list_cap = std::min<int64_t>(n, std::max<int64_t>(c->tm.candidates, 0));
    CK(c, c->h_list.ensure((size_t)list_cap * 16 + 16));
    {
        int64_t *list = (int64_t *)c->h_list.p;
        for (int32_t i = 0; i < n; i++) {
            if (!(res[i].status & NTL_READ_KEEP)) continue;
            if (n_kept >= list_cap) return fail(c, NTL_ERR_STATE, "more kept reads than locate candidates");
            list[2 * n_kept] = i; list[2 * n_kept + 1] = elems; n_kept++;
            c->kept_off[(size_t)i] = elems;
            elems += (((int64_t)res[i].n_win + 7) & ~(int64_t)7) * T;
        }
    }
    if (n_kept > 0) {
        int64_t *list = (int64_t *)c->h_list.p;
        CK(c, c->d_list.ensure((size_t)n_kept * 16));
        CK(c, c->d_kept.ensure((size_t)elems * 2 + 64));
        CK(c, c->h_cum.ensure((size_t)elems * 2 + 64));
        CK(c, cudaMemcpyAsync(c->d_list.p, list, (size_t)n_kept * 16, cudaMemcpyHostToDevice, c->stream));
        ntl_read_args ra;
        fill_read_args(c, &ra);
        CK(c, ntl_k_gather_windows(&ra, (const int64_t *)c->d_list.p, (int)n_kept, (uint16_t *)c->d_kept.p, T, c->stream));
        CK(c, cudaMemcpyAsync(c->h_cum.p, c->d_kept.p, (size_t)elems * 2, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(c, cudaEventRecord(c->ev[7], c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(c, cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]));
    c->tm.d2h_ms = ms;
    c->tm.d2h_bytes = (int64_t)(rbytes + (size_t)elems * 2 + sbytes);
    c->state = ST_DOWNLOADED;
    if (results) *results = res;
    return NTL_OK;
}

extern "C" int ntl_scan_batch(ntl_ctx *c, const char *const *seq, const int64_t *len, int32_t n,
                              const ntl_read_result **results)
{
    if (!c) return NTL_ERR_ARG;
    const double t0 = now_ms();
    Trace tr;
    int rc = pack_internal(c, seq, len, n, /*overlap=*/true);
    tr.mark("pack_internal");
    if (rc == NTL_OK) rc = ntl_batch_run(c);
    tr.mark("run");
    if (rc == NTL_OK) rc = ntl_batch_download(c, results);
    tr.mark("download");
    c->tm.total_ms = now_ms() - t0;
    return rc;
}

extern "C" int ntl_scan_batch_concat(ntl_ctx *c, const char *buf, const int64_t *offsets, int32_t n,
                                     const ntl_read_result **results)
{
    if (!c) return NTL_ERR_ARG;
    if (!buf || !offsets || n < 0) return fail(c, NTL_ERR_ARG, "ntl_scan_batch_concat: bad arguments");
    std::vector<const char *> seq((size_t)n);
    std::vector<int64_t> len((size_t)n);
    for (int32_t i = 0; i < n; i++) { seq[i] = buf + offsets[i]; len[i] = offsets[i + 1] - offsets[i]; }
    return ntl_scan_batch(c, seq.data(), len.data(), n, results);
}

extern "C" int ntl_get_timings(const ntl_ctx *c, ntl_timings *out)
{
    if (!c || !out) return NTL_ERR_ARG;
    *out = c->tm;
    return NTL_OK;
}

extern "C" void *ntl_stream(const ntl_ctx *c) { return c ? (void *)c->stream : nullptr; }

/* ============================================================================================== window tables */
extern "C" int ntl_get_windows(const ntl_ctx *c, int32_t read_idx, int32_t track, int32_t cap, int32_t *start_index,
                               int32_t *end_index, int32_t *covered, double *density)
{
    if (!c) return NTL_ERR_ARG;
    ntl_ctx *mc = const_cast<ntl_ctx *>(c);
    if (c->state < ST_DOWNLOADED) return fail(mc, NTL_ERR_STATE, "ntl_get_windows before the batch was downloaded");
    if (read_idx < 0 || read_idx >= c->n_reads || track < 0 || track >= c->dev.n_tracks)
        return fail(mc, NTL_ERR_ARG, "ntl_get_windows: read or track out of range");
    const ntl_read_result *r = (const ntl_read_result *)c->h_results.p + read_idx;
    if (r->status & NTL_READ_FILTERED) return 0;
    const int32_t n = r->n_win, S = c->dev.S;
    const int32_t L = ((const int32_t *)((const char *)c->h_meta.p + c->off_len))[read_idx];
    const uint16_t *cum;
    if (c->kept_off[(size_t)read_idx] >= 0) {
        cum = (const uint16_t *)c->h_cum.p + c->kept_off[(size_t)read_idx] + (size_t)track * (((size_t)n + 7) & ~(size_t)7);
    } else {                                    /* not a kept read: its table is still on the device */
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(mc, NTL_ERR_CUDA, "cudaSetDevice failed");
        mc->win_scratch.resize((size_t)n + 8);
        const uint16_t *src = (const uint16_t *)c->d_cum.p + (size_t)track * c->total_windows + r->win_offset;
        cudaError_t e = cudaMemcpyAsync(mc->win_scratch.data(), src, (size_t)n * 2, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return fail(mc, NTL_ERR_CUDA, "ntl_get_windows: %s", cudaGetErrorString(e));
        cum = mc->win_scratch.data();
    }
    for (int32_t k = 0; k < n && k < cap; k++) {
        const int32_t ws = 1 + k * S, we = (k == n - 1) ? L : (k + 1) * S;
        const int32_t cnt = (int32_t)((uint32_t)(cum[k] - (k ? cum[k - 1] : 0)) & 0xffffu);
        if (start_index) start_index[k] = ws;
        if (end_index) end_index[k] = we;
        if (covered) covered[k] = cnt;
        if (density) density[k] = (double)cnt / (double)(we - ws + 1);     /* NanoTel.R:467 */
    }
    return n;
}

extern "C" int64_t ntl_get_window_counts(const ntl_ctx *c, int32_t track, uint16_t *out, int64_t cap)
{
    if (!c) return NTL_ERR_ARG;
    ntl_ctx *mc = const_cast<ntl_ctx *>(c);
    if (c->state < ST_DOWNLOADED) return fail(mc, NTL_ERR_STATE, "ntl_get_window_counts before the batch was downloaded");
    if (track < 0 || track >= c->dev.n_tracks || (!out && cap > 0)) return fail(mc, NTL_ERR_ARG, "ntl_get_window_counts: bad arguments");
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(mc, NTL_ERR_CUDA, "cudaSetDevice failed");
    std::vector<uint16_t> raw((size_t)c->total_windows + 8);
    if (c->total_windows > 0) {
        const uint16_t *src = (const uint16_t *)c->d_cum.p + (size_t)track * c->total_windows;
        cudaError_t e = cudaMemcpyAsync(raw.data(), src, (size_t)c->total_windows * 2, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return fail(mc, NTL_ERR_CUDA, "ntl_get_window_counts: %s", cudaGetErrorString(e));
    }
    const ntl_read_result *res = (const ntl_read_result *)c->h_results.p;
    int64_t pos = 0;
    for (int32_t i = 0; i < c->n_reads; i++) {
        const int32_t n = res[i].n_win;
        const uint16_t *cum = raw.data() + res[i].win_offset;
        for (int32_t k = 0; k < n; k++, pos++)
            if (pos < cap) out[pos] = (uint16_t)(cum[k] - (k ? cum[k - 1] : 0));
    }
    return pos;
}

extern "C" int ntl_get_stages(const ntl_ctx *c, int32_t read_idx, int32_t track, ntl_stage *out)
{
    if (!c || !out) return NTL_ERR_ARG;
    ntl_ctx *mc = const_cast<ntl_ctx *>(c);
    if (!c->dev.debug_stages) return fail(mc, NTL_ERR_STATE, "context was created without NTL_OPT_DEBUG_STAGES");
    if (c->state < ST_DOWNLOADED) return fail(mc, NTL_ERR_STATE, "ntl_get_stages before the batch was downloaded");
    if (read_idx < 0 || read_idx >= c->n_reads || track < 0 || track >= c->dev.n_tracks)
        return fail(mc, NTL_ERR_ARG, "ntl_get_stages: read or track out of range");
    *out = ((const ntl_stage *)c->h_stages.p)[(size_t)read_idx * 3 + track];
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
