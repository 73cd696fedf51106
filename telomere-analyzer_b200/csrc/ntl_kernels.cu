/*
 * ntl_kernels.cu -- sm_100a kernels of libnanotel_b200 other than the compile-time-specialised span scan (ntl_scan.cuh):
 *   K2g ntl_scan_generic_kernel  any pattern set / any subseq_length, patterns in __constant__ memory: the path taken
 *                              when no specialised build exists (NTL_OPT_NO_JIT, no NVRTC and no precompiled cubin,
 *                              or a subseq_length without a span geometry); one warp per read, block counts by atomics
 *   K4  ntl_filter_kernel      --use_filter edge filter            (filter_reads/filter_density, NanoTel.R:2083-2163);
 *                              also marks the spans / work items of the reads it keeps for the span scan
 *       ntl_items_kernel       compacts the marked work items into the list the span scan walks
 *   K3a ntl_triage_kernel      eight lanes per read: proves that a read has no telomeric window on any track and no
 *                              hit in its first 18 bases and writes its (trivial) record, or hands it to K3b
 *   K3b ntl_locate_kernel      one warp per (candidate read, track): locator and refinement (find_telo_position_wraper
 *                              NanoTel.R:1080-1155 and everything it calls, analyze_read's densities and keep rule
 *                              :1840-1868)
 *       ntl_gather_windows_kernel   block counts of the kept reads, packed for the device-to-host copy
 *
 * Control flow inside a team / warp is uniform.  Where the reference consults its range list (get_accurate_start/end,
 * get_sub_density on arbitrary intervals) K3b re-derives hits and coverage locally from the packed read -- one word
 * (32 positions) per lane, bit-parallel -- so that K2 never has to spill per-base masks to HBM.  All fp64 expressions
 * are written exactly as NanoTel.R evaluates them (int/int divisions in double, sums in window order); this file must
 * be compiled with --fmad=false.
 *
 * Positions: K3/K4 address a read through VIRTUAL words whose bit b holds position 32 w + b (1-based position =
 * bit index, bit 0 of word 0 is a pad), i.e. the packed stream (position p = bit p - 1, ntl_dev.h) shifted up by one
 * bit on the fly (rv_word); that keeps every interval computation below in the reference's 1-based coordinates.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "ntl_dev.h"
#include "../../include/nanotel_b200.h"

typedef unsigned int u32;
#define NTL_FULL 0xffffffffu

__constant__ ntl_dev_params c_prm;

__device__ __forceinline__ int ntl_nwin(int L, int S)
{
    /* split_telo, NanoTel.R:216-224: drop the last window if  L - last_start < S / 2  (real division) */
    int n = (L - 1) / S + 1;
    int last = 1 + (n - 1) * S;
    if (2 * (L - last) < S) n -= 1;
    return n;
}

/* bits b of the virtual word starting at bit index wpos with 1 <= wpos + b <= L: the low  clamp(L - wpos + 1, 0, 32)
 * bits (one clamped funnel shift), minus bit 0 of the read's very first word (the pad position) */
__device__ __forceinline__ u32 ntl_valid_word(int wpos, int L)
{
    int nb = L - wpos + 1;
    if (nb < 0) nb = 0;
    u32 m = __funnelshift_lc(NTL_FULL, 0u, nb);
    if (wpos == 0) m &= ~1u;
    return m;
}

/* =============================================================================================================
 * Shared by K3 and K4: random access into a packed read
 * ============================================================================================================= */
#define NTL_NONE (-999999999)
#define NTL_TRIAGE_MAX_WIN 8192
#define NTL_IMAX 2147483647

struct ReadView {
    const u32 *base;  /* record of the read's position word 0: {lo, hi} (fmt 0) or {A, C, G, T} (fmt 1) per word */
    int L;
    int fmt;        /* 0: 2-bit records of 2 words, 1: 4-bit records of 4 words */
    int n_words;    /* virtual words: (L >> 5) + 1 */
    int n_raw;      /* position words that hold letters: (L + 31) >> 5 */
    /* locate kernel only: the two coverage blocks (one word per lane, warp_cov) that get_accurate_start / _end
     * computed for this (read, track), kept in shared memory so that the partial windows of the final density are
     * not derived from the read a second time.  ccov[slot * 32 + lane], cwb[slot] = first word or NTL_NONE. */
    u32 *ccov;
    int *cwb;
};

__device__ __forceinline__ void rv_init(ReadView &rv, const ntl_read_args &a, int r)
{
    rv.L = a.len[r]; rv.fmt = a.fmt[r];
    rv.base = rv.fmt ? a.arena4 + (size_t)a.woff[r] * 4 : a.arena2 + (size_t)a.woff[r] * 2;
    rv.n_words = (rv.L >> 5) + 1; rv.n_raw = (rv.L + 31) >> 5;
    rv.ccov = nullptr; rv.cwb = nullptr;
}

__device__ __forceinline__ u32 rv_raw(const ReadView &rv, int plane, int x)
{
    if (x < 0 || x >= rv.n_raw) return 0u;
    return rv.fmt ? rv.base[(size_t)x * 4 + plane] : rv.base[(size_t)x * 2 + plane];
}

/* virtual word w of a plane: bit b = position 32 w + b */
__device__ __forceinline__ u32 rv_word(const ReadView &rv, int plane, int w)
{
    if (w < 0 || w >= rv.n_words) return 0u;
    return __funnelshift_l(rv_raw(rv, plane, w - 1), rv_raw(rv, plane, w), 1);
}

/* ---- one word per lane: planes of read word w as A/C/G/T bit masks, zero outside [1, L] */
__device__ __forceinline__ void word_planes(const ReadView &rv, int w, u32 (&pl)[4])
{
    if (w < 0 || w >= rv.n_words) { pl[0] = pl[1] = pl[2] = pl[3] = 0u; return; }
    const u32 vm = ntl_valid_word(w << 5, rv.L);
    if (rv.fmt == 0) {
        const u32 lo = rv_word(rv, 0, w), hi = rv_word(rv, 1, w);
        pl[0] = ~hi & ~lo & vm; pl[1] = ~hi & lo & vm; pl[2] = hi & lo & vm; pl[3] = hi & ~lo & vm;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) pl[k] = rv_word(rv, k, w) & vm;
    }
}

/* exact / <=1-mismatch hit starts of one pattern inside the lane's word (letters may run into the next word) */
__device__ __forceinline__ void word_hits(const ntl_dev_pat &pt, const u32 (&pw)[4], const u32 (&pn)[4], u32 *exact,
                                          u32 *le1)
{
    u32 ones = 0u, twos = 0u;
    const bool fx = pt.fixed != 0;
#pragma unroll 1
    for (int j = 0; j < pt.m; j++) {
        const u32 mA = pt.mux4[j][0], mC = pt.mux4[j][1], mG = pt.mux4[j][2], mT = pt.mux4[j][3];
        u32 ew, en;
        if (fx) {
            ew = ~((pw[0] ^ mA) | (pw[1] ^ mC) | (pw[2] ^ mG) | (pw[3] ^ mT));
            en = ~((pn[0] ^ mA) | (pn[1] ^ mC) | (pn[2] ^ mG) | (pn[3] ^ mT));
        } else {
            ew = (pw[0] & mA) | (pw[1] & mC) | (pw[2] & mG) | (pw[3] & mT);
            en = (pn[0] & mA) | (pn[1] & mC) | (pn[2] & mG) | (pn[3] & mT);
        }
        const u32 x = ~__funnelshift_r(ew, en, j);
        twos |= ones & x;
        ones ^= x;
    }
    *exact = ~(ones | twos);
    *le1 = ~twos;
}

/* Coverage of track t (0 exact, 1 one mismatch, 2 one mismatch + TVR; the union of the trimmed hit intervals of
 * get_density_iranges, NanoTel.R:308-397) for read words wbase .. wbase+31, one word per lane: lane l holds positions
 * 32 (wbase + l) .. +31.  Lane 0 lacks the spill of word wbase-1: callers start one word early and ignore lane 0.
 * *hs = exact hit starts of main pattern 0 (the raw hit list of NanoTel.R:349-354). */
/* planes A, C, G, T (one-hot or IUPAC) of virtual words wbase + lane (pw) and wbase + lane + 1 (pn), zero outside
 * [1, L]: every lane loads ONE raw record (plus lanes 0 / 31 the record before / after the block) and takes its
 * neighbours' by shuffle -- virtual word w = raw words (w - 1, w) shifted up by one bit */
__device__ __forceinline__ void warp_word_planes(const ReadView &rv, int wbase, int lane, u32 (&pw)[4], u32 (&pn)[4])
{
    const int w = wbase + lane;
    const int NP = rv.fmt ? 4 : 2;
    u32 own[4] = {0u, 0u, 0u, 0u}, ext[4] = {0u, 0u, 0u, 0u};
    const int xe = lane == 0 ? wbase - 1 : wbase + 32;             /* only lanes 0 and 31 use theirs */
    if (rv.fmt == 0) {
        if (w >= 0 && w < rv.n_raw) { const uint2 v = *reinterpret_cast<const uint2 *>(rv.base + (size_t)w * 2); own[0] = v.x; own[1] = v.y; }
        if ((lane == 0 || lane == 31) && xe >= 0 && xe < rv.n_raw) { const uint2 v = *reinterpret_cast<const uint2 *>(rv.base + (size_t)xe * 2); ext[0] = v.x; ext[1] = v.y; }
    } else {
        if (w >= 0 && w < rv.n_raw) { const uint4 v = *reinterpret_cast<const uint4 *>(rv.base + (size_t)w * 4); own[0] = v.x; own[1] = v.y; own[2] = v.z; own[3] = v.w; }
        if ((lane == 0 || lane == 31) && xe >= 0 && xe < rv.n_raw) { const uint4 v = *reinterpret_cast<const uint4 *>(rv.base + (size_t)xe * 4); ext[0] = v.x; ext[1] = v.y; ext[2] = v.z; ext[3] = v.w; }
    }
    u32 vw[4], vn[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < NP) {
            u32 prev = __shfl_up_sync(NTL_FULL, own[k], 1), next = __shfl_down_sync(NTL_FULL, own[k], 1);
            if (lane == 0) prev = ext[k];
            if (lane == 31) next = ext[k];
            vw[k] = __funnelshift_l(prev, own[k], 1);
            vn[k] = __funnelshift_l(own[k], next, 1);
        } else { vw[k] = 0u; vn[k] = 0u; }
    }
    const u32 mw = (w < 0 || w >= rv.n_words) ? 0u : ntl_valid_word(w << 5, rv.L);
    const u32 mn = (w + 1 < 0 || w + 1 >= rv.n_words) ? 0u : ntl_valid_word((w + 1) << 5, rv.L);
    if (rv.fmt == 0) {
        pw[0] = ~vw[1] & ~vw[0] & mw; pw[1] = ~vw[1] & vw[0] & mw; pw[2] = vw[1] & vw[0] & mw; pw[3] = vw[1] & ~vw[0] & mw;
        pn[0] = ~vn[1] & ~vn[0] & mn; pn[1] = ~vn[1] & vn[0] & mn; pn[2] = vn[1] & vn[0] & mn; pn[3] = vn[1] & ~vn[0] & mn;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) { pw[k] = vw[k] & mw; pn[k] = vn[k] & mn; }
    }
}

__device__ __noinline__ u32 warp_cov(const ReadView &rv, int t, int wbase, int lane, u32 *hs)
{
    const int w = wbase + lane;
    u32 pw[4], pn[4];
    warp_word_planes(rv, wbase, lane, pw, pn);
    u32 cov = 0u, h0 = 0u;
#pragma unroll 1
    for (int p = 0; p < c_prm.n_main; p++) {
        u32 ex, le;
        word_hits(c_prm.main_pat[p], pw, pn, &ex, &le);
        if (p == 0) h0 = ex;
        const u32 H = t >= 1 ? le : ex;
        u32 Hp = __shfl_up_sync(NTL_FULL, H, 1);
        if (lane == 0) Hp = 0u;
#pragma unroll 1
        for (int j = 0; j < c_prm.main_pat[p].m; j++) cov |= __funnelshift_l(Hp, H, j);
    }
    if (t == 2) {
#pragma unroll 1
        for (int p = 0; p < c_prm.n_tvr; p++) {
            u32 ex, le;
            word_hits(c_prm.tvr_pat[p], pw, pn, &ex, &le);
            u32 Hp = __shfl_up_sync(NTL_FULL, ex, 1);
            if (lane == 0) Hp = 0u;
#pragma unroll 1
            for (int j = 0; j < c_prm.tvr_pat[p].m; j++) cov |= __funnelshift_l(Hp, ex, j);
        }
    }
    *hs = h0;
    if (w < 0 || w >= rv.n_words) return 0u;
    return cov & ntl_valid_word(w << 5, rv.L);          /* trim() to [1, L] */
}

/* =============================================================================================================
 * K4: edge filter (filter_reads / filter_density, NanoTel.R:2083-2163)
 * ============================================================================================================= */
__device__ __forceinline__ u32 range_word_mask(int w, int vlo, int vhi)
{
    const int p0 = w << 5;
    int lb = vlo - p0; if (lb < 0) lb = 0;
    int hb = vhi - p0; if (hb > 31) hb = 31;
    return hb < lb ? 0u : ((NTL_FULL >> (31 - hb)) & (NTL_FULL << lb));
}

/* Eight lanes per read, one word of the 200-base slice per lane (the slice spans at most 8 words): exact hits with
 * fixed = FALSE (letters match iff their IUPAC sets intersect), union of the hit intervals, covered / 200. */
__global__ void __launch_bounds__(256) ntl_filter_kernel(const ntl_read_args a)
{
    const int lane = threadIdx.x & 31;
    const int sub = lane & 7;
    const u32 tmask = 0xffu << (lane & 24);
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (r >= a.n_reads) return;
    ReadView rv;
    rv_init(rv, a, r);
    int keep = 0;
    if (rv.L >= 1000) {                                         /* :2124 */
        int lo, hi;
        if (c_prm.right_edge) { hi = rv.L - 70; lo = hi - 199; }   /* subseq(end = -(70+1), width = 200) :2131-2134 */
        else { lo = 71; hi = 270; }                                 /* subseq(start = 71, width = 200)   :2136       */
        const int w = (lo >> 5) + sub;
        /* the slice is the whole subject of matchPattern: letters outside it do not exist (k = 0: no hit there) */
        const u32 vw = range_word_mask(w, lo, hi), vn = range_word_mask(w + 1, lo, hi);
        u32 pw[4], pn[4];
        word_planes(rv, w, pw);
        word_planes(rv, w + 1, pn);
#pragma unroll
        for (int k = 0; k < 4; k++) { pw[k] &= vw; pn[k] &= vn; }
        u32 cov = 0u;
#pragma unroll 1
        for (int p = 0; p < c_prm.n_main; p++) {                    /* fixed = FALSE always, exact (:2091-2096) */
            const ntl_dev_pat &pt = c_prm.main_pat[p];
            u32 mis = 0u;
#pragma unroll 1
            for (int j = 0; j < pt.m; j++) {
                const u32 mA = pt.mux4[j][0], mC = pt.mux4[j][1], mG = pt.mux4[j][2], mT = pt.mux4[j][3];
                const u32 ew = (pw[0] & mA) | (pw[1] & mC) | (pw[2] & mG) | (pw[3] & mT);
                const u32 en = (pn[0] & mA) | (pn[1] & mC) | (pn[2] & mG) | (pn[3] & mT);
                mis |= ~__funnelshift_r(ew, en, j);
            }
            const u32 H = ~mis;
            u32 Hp = __shfl_up_sync(tmask, H, 1, 8);
            if (sub == 0) Hp = 0u;
#pragma unroll 1
            for (int j = 0; j < pt.m; j++) cov |= __funnelshift_l(Hp, H, j);
        }
        int covered = __popc(cov & vw);
        covered += __shfl_xor_sync(tmask, covered, 1, 8);
        covered += __shfl_xor_sync(tmask, covered, 2, 8);
        covered += __shfl_xor_sync(tmask, covered, 4, 8);
        const double total_density = (double)covered / (double)(hi - lo + 1);   /* :2100 */
        keep = total_density >= c_prm.filter_threshold ? 1 : 0;                  /* :2101, :2143 */
    }
    if (sub == 0) a.pass[r] = (uint8_t)keep;
    /* the span scan walks spans, not reads: give this read's spans their verdict and, for a read that stays, mark the
     * work items (groups of 32 spans) that hold them */
    if (c_prm.BPS > 0 && a.span_flags[rv.fmt] != nullptr) {
        const int W = c_prm.W;
        const int64_t s0 = a.woff[r] / W;
        const int nsp = (rv.L + 32 * W - 1) / (32 * W);
        const int ltail = rv.L - (nsp - 1) * 32 * W;
        const int tail_from = (nsp >= 2 && ltail < NTL_DEV_MAX_LEN) ? nsp - 2 : nsp - 1;
        uint8_t *fl = a.span_flags[rv.fmt] + s0;
        uint8_t *act = a.item_active[rv.fmt];
        for (int s = sub; s < nsp; s += 8) {
            fl[s] = (uint8_t)((s == 0 ? NTL_SPAN_FIRST : 0) | (s >= tail_from ? NTL_SPAN_TAIL : 0) | (keep ? 0 : NTL_SPAN_SKIP));
            if (keep) act[(s0 + s) / NTL_ITEM_SPANS] = 1;
        }
    }
}

/* the work items the filter marked, as a list (order does not matter) */
__global__ void __launch_bounds__(256) ntl_items_kernel(const uint8_t *active, int n_items, int32_t *items, u32 *counter)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const bool on = i < n_items && active[i] != 0;
    const u32 m = __ballot_sync(NTL_FULL, on);
    if (m == 0u) return;
    int base = 0;
    const int leader = __ffs((int)m) - 1;
    if (lane == leader) base = (int)atomicAdd(counter, (u32)__popc(m));
    base = __shfl_sync(NTL_FULL, base, leader);
    if (on) items[base + __popc(m & ((1u << lane) - 1u))] = i;
}

/* =============================================================================================================
 * K3: locator.  ONE THREAD per (candidate read, track): the window state machines of the reference are sequential
 * scans with data-dependent exits, and a candidate needs only a few hundred bases of locally recomputed coverage, so
 * a warp per item spends its 32 lanes on the same scalar work (measured: 4 300 warp-instructions per item, 41 % issue
 * utilisation, bound by dependent latency); per thread the same item is a few thousand scalar instructions, 32 items
 * share every warp-instruction, and the lanes of a warp (neighbouring tracks of the same reads) follow similar paths.
 * ============================================================================================================= */
/* one copy of the IEEE double division (a ~80-instruction sequence) for the whole locate kernel */
__device__ __noinline__ double k3_div(double a, double b) { return a / b; }

struct WinTab {                 /* the window table of one track (analyze_subtelos :737-764), never materialised */
    const uint16_t *cnt;        /* covered bases per block of SG positions (K2); window k = blocks k Q .. k Q + Q - 1, */
    int n, nb, Q, SG, S, L;     /* the last window = every remaining block (n windows, nb blocks)                     */
    int thr_reg, thr_last;      /* smallest telomeric count of a regular window / of this read's last window */
};
__device__ __forceinline__ int wt_start(const WinTab &w, int k) { return 1 + k * w.S; }
__device__ __forceinline__ int wt_end(const WinTab &w, int k) { return k == w.n - 1 ? w.L : (k + 1) * w.S; }
__device__ __forceinline__ int wt_count(const WinTab &w, int k)
{
    if (w.Q == 1 && k < w.n - 1) return (int)w.cnt[k];
    const int b0 = k * w.Q, b1 = k == w.n - 1 ? w.nb : b0 + w.Q;
    int c = 0;
    for (int b = b0; b < b1; b++) c += (int)w.cnt[b];
    return c;
}
/* class == CCCTAA (NanoTel.R:751-758), i.e. !(count / width < min_density): the smallest such count per width was
 * found on the host with the same double division, so the test is an integer compare here. */
__device__ __forceinline__ bool wt_telo_count(const WinTab &w, int k, int count)
{
    return count >= (k == w.n - 1 ? w.thr_last : w.thr_reg);
}
__device__ __forceinline__ double wt_density_of_count(const WinTab &w, int k, int count)
{
    /* get_sub_density (NanoTel.R:467): count / width in double */
    return k3_div((double)count, (double)(wt_end(w, k) - wt_start(w, k) + 1));
}

/* First telomeric window at or after k0 (n if none) / last telomeric window at or before k0 (-1 if none).  The
 * reference's scans walk window by window; between two telomeric windows nothing happens but resets, so the walk may
 * jump.  When a window is one block (Q == 1) the regular windows are skipped 32 at a time: four independent 16-byte
 * loads of 8 counts (a read's block range starts on a multiple of 8 entries), two compares per pair of counts. */
__device__ __forceinline__ bool pair_has_telo(u32 x, u32 thr16) { return (x << 16) >= thr16 || x >= thr16; }
__device__ __forceinline__ bool quad_has_telo(const uint4 &v, u32 thr16)
{
    return pair_has_telo(v.x, thr16) || pair_has_telo(v.y, thr16) || pair_has_telo(v.z, thr16) || pair_has_telo(v.w, thr16);
}

__device__ __noinline__ int next_telo_fwd(const WinTab &w, int k0)
{
    const int n = w.n;
    int k = k0 < 0 ? 0 : k0;
    if (w.Q == 1) {
        const int nr = n - 1;                                            /* regular windows 0 .. nr - 1 */
        const u32 thr16 = (u32)w.thr_reg << 16;
        while (k < nr && (k & 7)) { if ((int)w.cnt[k] >= w.thr_reg) return k; k++; }
        const uint4 *cv = reinterpret_cast<const uint4 *>(w.cnt);
        while (k + 32 <= nr) {
            const uint4 v0 = __ldg(cv + (k >> 3)), v1 = __ldg(cv + (k >> 3) + 1), v2 = __ldg(cv + (k >> 3) + 2), v3 = __ldg(cv + (k >> 3) + 3);
            if (quad_has_telo(v0, thr16) || quad_has_telo(v1, thr16) || quad_has_telo(v2, thr16) || quad_has_telo(v3, thr16)) break;
            k += 32;
        }
        while (k < nr) { if ((int)w.cnt[k] >= w.thr_reg) return k; k++; }
        if (k == nr && nr >= 0 && wt_telo_count(w, nr, wt_count(w, nr))) return nr;
        return n;
    }
    for (; k < n; k++) if (wt_telo_count(w, k, wt_count(w, k))) return k;
    return n;
}

__device__ __noinline__ int next_telo_bwd(const WinTab &w, int k0)
{
    const int n = w.n;
    int k = k0 >= n ? n - 1 : k0;
    if (k < 0) return -1;
    if (w.Q == 1) {
        const u32 thr16 = (u32)w.thr_reg << 16;
        if (k == n - 1) { if (wt_telo_count(w, k, wt_count(w, k))) return k; k--; }
        while (k >= 0 && (k & 7) != 7) { if ((int)w.cnt[k] >= w.thr_reg) return k; k--; }
        const uint4 *cv = reinterpret_cast<const uint4 *>(w.cnt);
        while (k >= 31) {                                                /* windows k - 31 .. k, k % 8 == 7 */
            const int g = k >> 3;
            const uint4 v0 = __ldg(cv + g), v1 = __ldg(cv + g - 1), v2 = __ldg(cv + g - 2), v3 = __ldg(cv + g - 3);
            if (quad_has_telo(v0, thr16) || quad_has_telo(v1, thr16) || quad_has_telo(v2, thr16) || quad_has_telo(v3, thr16)) break;
            k -= 32;
        }
        while (k >= 0) { if ((int)w.cnt[k] >= w.thr_reg) return k; k--; }
        return -1;
    }
    for (; k >= 0; k--) if (wt_telo_count(w, k, wt_count(w, k))) return k;
    return -1;
}

/* find_telo_position (NanoTel.R:973-1077): forward scan for the first run of telomeric windows with in_a_row >= R
 * and score >= T, then the backward scan for the end.  Windows are 0-based here; non-telomeric windows only reset
 * the run state, so stretches of them are jumped over (next_telo_fwd / _bwd). */
__device__ __noinline__ void find_telo_position(const WinTab &w, double R, double T, int *ps, int *pe)
{
    const int n = w.n;
    double score = 0.0;
    int start = -1, end = -1, in_a_row = 0;
    int end_position = 0;                                                /* 1-based i + 1 (:1022) */
    for (int k = 0; k < n;) {                                            /* :1003-1025 */
        const int c = wt_count(w, k);
        if (!wt_telo_count(w, k, c)) {
            score = 0.0; start = -1; in_a_row = 0;
            if (0.0 >= R && 0.0 >= T) { end_position = k + 2; break; }   /* never with the reference's R >= 3 */
            k = next_telo_fwd(w, k + 1);
            continue;
        }
        in_a_row += 1;
        score = score + wt_density_of_count(w, k, c);                    /* :1014 */
        if (start == -1) start = wt_start(w, k);
        if ((double)in_a_row >= R && score >= T) { end_position = k + 2; break; }
        k++;
    }
    if (end_position == 0) { *ps = -1; *pe = -1; return; }               /* :1026-1028 */
    end = -1; score = 0.0; in_a_row = 0;
    if ((double)end_position >= (double)n - R + 1.0) {                   /* :1037-1044 */
        /* i = n; while (i > end_position && window i is not telomeric) i--  (1-based) */
        int i = next_telo_bwd(w, n - 1) + 1;                             /* 1-based, 0 if none */
        if (i < end_position) i = end_position;
        end = wt_end(w, (i < n ? i : n) - 1);                            /* end_position may be n + 1: the loop does not run */
    } else {                                                             /* :1046-1068 */
        for (int i = n; i >= end_position;) {
            const int c = wt_count(w, i - 1);
            if (!wt_telo_count(w, i - 1, c)) {
                score = 0.0; end = -1; in_a_row = 0;
                const int j = next_telo_bwd(w, i - 2) + 1;               /* 1-based */
                i = j;
                continue;
            }
            in_a_row += 1;
            score = score + wt_density_of_count(w, i - 1, c);
            if (end == -1) end = wt_end(w, i - 1);
            if ((double)in_a_row >= R && score >= T) break;
            i--;
        }
    }
    if (start > end) end = start + (wt_end(w, 0) - wt_start(w, 0));      /* :1072-1074 */
    *ps = start; *pe = end;
}

/* find_left_telo (NanoTel.R:906-959): walk from window 0 to the first telomeric window, give up as soon as a window
 * starts more than max_diff = 200 from the edge, then extend over the telomeric run. */
__device__ __noinline__ void find_left_telo(const WinTab &w, int *ps, int *pe)
{
    const int n = w.n;
    int start = 1, end = 1, last_i = 0;
    for (int i = 0; i < n; i++) {
        if (wt_start(w, i) > 200) { *ps = -1; *pe = -1; return; }
        if (!wt_telo_count(w, i, wt_count(w, i))) continue;
        start = wt_start(w, i); last_i = i; break;
    }
    const int last_i_start = last_i;
    for (int i = last_i; i < n; i++) {
        if (!wt_telo_count(w, i, wt_count(w, i))) break;
        end = wt_end(w, i);
    }
    if (start > end && n > 0) end = start + (wt_end(w, last_i_start) - wt_start(w, last_i_start));
    *ps = start; *pe = end;
}

/* find_right_telo (NanoTel.R:843-899); n == 0 is the caller's REF_ERROR case */
__device__ __noinline__ void find_right_telo(const WinTab &w, int *ps, int *pe)
{
    const int n = w.n;
    int start = 1, end = 1, last_i = 0;
    for (int i = n - 1; i >= 0; i--) {
        if (wt_end(w, i) < w.L - 200) { *ps = -1; *pe = -1; return; }
        if (!wt_telo_count(w, i, wt_count(w, i))) continue;
        end = wt_end(w, i); last_i = i; break;
    }
    for (int i = last_i; i >= 0; i--) {
        if (!wt_telo_count(w, i, wt_count(w, i))) break;
        start = wt_start(w, i); last_i = i;
    }
    if (start > end) end = start + (wt_end(w, last_i) - wt_start(w, last_i));
    *ps = start; *pe = end;
}

/* A TEAM of eight lanes works on one (candidate, track) item: the scalar state machines run on all eight lanes alike,
 * and where coverage has to be recomputed from the read every lane takes one word.
 * team_cov: coverage of track t (0 exact, 1 one mismatch, 2 one mismatch + TVR; the union of the trimmed hit intervals
 * of get_density_iranges, NanoTel.R:308-397) for virtual words w0 - 1 + sub (bit b of word w = position 32 w + b), one
 * word per lane: hit starts of every pattern in (previous word, word) dilated by the pattern length.  Lane 0 lacks the
 * hits of the word before its own: its word (w0 - 1) is only the look-behind; words w0 .. w0 + 6 are complete.
 * *hs = exact hit starts of main pattern 0 in the lane's word (the raw hit list of NanoTel.R:349-354). */
#define NTL_TEAM 8
__device__ __noinline__ u32 team_cov(const ReadView &rv, int t, int w0, int sub, u32 tmask, u32 *hs)
{
    const int w = w0 - 1 + sub;
    u32 pw[4], pn[4];
    word_planes(rv, w, pw);
    word_planes(rv, w + 1, pn);
    u32 cov = 0u, h0 = 0u;
#pragma unroll 1
    for (int p = 0; p < c_prm.n_main; p++) {
        u32 ex, le;
        word_hits(c_prm.main_pat[p], pw, pn, &ex, &le);
        if (p == 0) h0 = ex;
        const u32 H = t >= 1 ? le : ex;
        u32 Hp = __shfl_up_sync(tmask, H, 1, NTL_TEAM);
        if (sub == 0) Hp = 0u;
#pragma unroll 1
        for (int j = 0; j < c_prm.main_pat[p].m; j++) cov |= __funnelshift_l(Hp, H, j);
    }
    if (t == 2) {
#pragma unroll 1
        for (int p = 0; p < c_prm.n_tvr; p++) {
            u32 ex, le;
            word_hits(c_prm.tvr_pat[p], pw, pn, &ex, &le);
            u32 Hp = __shfl_up_sync(tmask, ex, 1, NTL_TEAM);
            if (sub == 0) Hp = 0u;
#pragma unroll 1
            for (int j = 0; j < c_prm.tvr_pat[p].m; j++) cov |= __funnelshift_l(Hp, ex, j);
        }
    }
    *hs = h0;
    if (w < 0 || w >= rv.n_words) return 0u;
    return cov & ntl_valid_word(w << 5, rv.L);          /* trim() to [1, L] */
}

/* bits of virtual word w that lie inside positions [lo, hi] */
__device__ __forceinline__ u32 word_range_mask(int w, int lo, int hi)
{
    const int p0 = w << 5;
    int lb = lo - p0; if (lb < 0) lb = 0;
    int hb = hi - p0; if (hb > 31) hb = 31;
    return hb < lb ? 0u : ((NTL_FULL >> (31 - hb)) & (NTL_FULL << lb));
}
__device__ __forceinline__ int team_sum(int v, u32 tmask)
{
    v += __shfl_xor_sync(tmask, v, 1, NTL_TEAM);
    v += __shfl_xor_sync(tmask, v, 2, NTL_TEAM);
    v += __shfl_xor_sync(tmask, v, 4, NTL_TEAM);
    return v;
}
__device__ __forceinline__ int team_max(int v, u32 tmask)
{
    v = max(v, __shfl_xor_sync(tmask, v, 1, NTL_TEAM));
    v = max(v, __shfl_xor_sync(tmask, v, 2, NTL_TEAM));
    v = max(v, __shfl_xor_sync(tmask, v, 4, NTL_TEAM));
    return v;
}
__device__ __forceinline__ int team_min(int v, u32 tmask)
{
    v = min(v, __shfl_xor_sync(tmask, v, 1, NTL_TEAM));
    v = min(v, __shfl_xor_sync(tmask, v, 2, NTL_TEAM));
    v = min(v, __shfl_xor_sync(tmask, v, 4, NTL_TEAM));
    return v;
}
/* bits of the lanes' words (lane sub holds word w0 - 1 + sub; lane 0 is never used) inside [lo, hi]: largest /
 * smallest position, number of set bits */
__device__ __forceinline__ int team_bits_max(u32 bits, int w0, int sub, u32 tmask, int lo, int hi)
{
    const u32 m = sub == 0 ? 0u : bits & word_range_mask(w0 - 1 + sub, lo, hi);
    const int r = team_max(m ? ((w0 - 1 + sub) << 5) + 31 - __clz((int)m) : -NTL_IMAX, tmask);
    return r == -NTL_IMAX ? NTL_NONE : r;
}
__device__ __forceinline__ int team_bits_min(u32 bits, int w0, int sub, u32 tmask, int lo, int hi)
{
    const u32 m = sub == 0 ? 0u : bits & word_range_mask(w0 - 1 + sub, lo, hi);
    const int r = team_min(m ? ((w0 - 1 + sub) << 5) + __ffs((int)m) - 1 : NTL_IMAX, tmask);
    return r == NTL_IMAX ? NTL_NONE : r;
}
__device__ __forceinline__ int team_bits_popc(u32 bits, int w0, int sub, u32 tmask, int lo, int hi)
{
    return team_sum(sub == 0 ? 0 : __popc(bits & word_range_mask(w0 - 1 + sub, lo, hi)), tmask);
}

/* covered positions of track t inside [lo, hi] (1 <= lo <= hi <= L), recomputed from the read, 7 words per step */
__device__ __noinline__ int local_count(const ReadView &rv, int t, int lo, int hi, int sub, u32 tmask)
{
    int total = 0;
    for (int w0 = lo >> 5; w0 <= (hi >> 5); w0 += NTL_TEAM - 1) {
        u32 hs;
        const u32 cov = team_cov(rv, t, w0, sub, tmask, &hs);
        total += team_bits_popc(cov, w0, sub, tmask, lo, hi);
    }
    return total;
}

/* covered bases of track t inside [a, b] (get_sub_density's numerator, NanoTel.R:467): whole blocks come from K2's
 * counts, the partial blocks at the two ends are recomputed from the read. */
__device__ __noinline__ int covered_in(const ReadView &rv, const WinTab &w, int t, int a, int b, int sub, u32 tmask)
{
    const int lo = a < 1 ? 1 : a, hi = b > rv.L ? rv.L : b;
    if (hi < lo) return 0;
    const int SG = w.SG;
    const int blo = (lo - 1) / SG, bhi = (hi - 1) / SG;
    /* block j = positions [j SG + 1, min((j + 1) SG, L)] */
    const int blo_s = blo * SG + 1, bhi_e = (bhi + 1) * SG < rv.L ? (bhi + 1) * SG : rv.L;
    if (blo == bhi) {
        if (lo == blo_s && hi == bhi_e) return (int)w.cnt[blo];
        return local_count(rv, t, lo, hi, sub, tmask);
    }
    int total = 0, bf = blo, bl = bhi;
    if (lo != blo_s) { total += local_count(rv, t, lo, (blo + 1) * SG, sub, tmask); bf = blo + 1; }
    if (hi != bhi_e) { total += local_count(rv, t, bhi * SG + 1, hi, sub, tmask); bl = bhi - 1; }
    for (int j = bf; j <= bl; j++) total += (int)w.cnt[j];
    return total;
}

__device__ __forceinline__ double density_of(const ReadView &rv, const WinTab &w, int t, int a, int b, int sub, u32 tmask)
{
    const int cv = covered_in(rv, w, t, a, b, sub, tmask);
    return cv == 0 ? 0.0 : k3_div((double)cv, (double)(b - a + 1));      /* 0 / width is +0.0 exactly */
}

/* get_accurate_end (NanoTel.R:1692-1721).  `ranges` are the raw exact hits of the single fixed pattern on track A
 * (NanoTel.R:349-354) and the reduced runs of the coverage otherwise (:341-345). */
__device__ __noinline__ int get_accurate_end(const ReadView &rv, int t, int telo_end, int sub, u32 tmask)
{
    if (telo_end == -1) return -1;
    /* range ends inside [e - 99, e + 50]: the words holding [e - 99, e + 51] (at most 6) are lanes 1 .. 7 */
    const int w0 = (telo_end - 99) >> 5;
    u32 hs;
    const u32 cov = team_cov(rv, t, w0, sub, tmask, &hs);
    u32 en;
    if (t == 0 && c_prm.raw_hits_A) {
        const u32 hp = __shfl_up_sync(tmask, hs, 1, NTL_TEAM);
        en = __funnelshift_l(hp, hs, c_prm.main_pat[0].m - 1);           /* end = start + m - 1 */
    } else {
        u32 nx = __shfl_down_sync(tmask, cov, 1, NTL_TEAM);
        if (sub == NTL_TEAM - 1) nx = 0u;                                /* lane 7's word is beyond e + 51: only a look-ahead */
        en = cov & ~((cov >> 1) | (nx << 31));
    }
    int e_index = telo_end;
    const int m1 = team_bits_max(en, w0, sub, tmask, telo_end - 99, telo_end);
    if (m1 != NTL_NONE) e_index = m1;
    const int m2 = team_bits_max(en, w0, sub, tmask, telo_end + 1, telo_end + 50);
    if (m2 != NTL_NONE) e_index = m2;
    return e_index;
}

/* get_accurate_start (NanoTel.R:1726-1764) */
__device__ __noinline__ int get_accurate_start(const ReadView &rv, int t, int telo_start, int sub, u32 tmask)
{
    if (telo_start == -1) return telo_start;
    const int s = telo_start;
    /* range starts inside [s - 36, s + 99]: the words holding [s - 37, s + 99] (at most 6) are lanes 1 .. 7 */
    const int w0 = (s - 37) >> 5;
    u32 hs;
    const u32 cov = team_cov(rv, t, w0, sub, tmask, &hs);
    u32 st;
    if (t == 0 && c_prm.raw_hits_A) st = hs;
    else {
        const u32 pv = __shfl_up_sync(tmask, cov, 1, NTL_TEAM);          /* lane 1 looks behind into lane 0's (incomplete) word:
                                                                            only bit 32 w0 could be wrong, and it is < s - 36 */
        st = cov & ~((cov << 1) | (pv >> 31));
    }
    const int c50 = team_bits_popc(cov, w0, sub, tmask, s, s + 49);
    const double first_50 = k3_div((double)c50, 50.0);                         /* IRanges(start, width = 50) :1732 */
    if (first_50 < 0.3) {
        const int a = team_bits_min(st, w0, sub, tmask, s + 48, s + 99);
        if (a != NTL_NONE) telo_start = a;
        const int b = team_bits_min(st, w0, sub, tmask, s + 33, s + 48);
        if (b != NTL_NONE) telo_start = b;
    } else {
        const int a = team_bits_min(st, w0, sub, tmask, s, s + 99);
        if (a != NTL_NONE) telo_start = a;
        if (first_50 >= 0.72) {
            const int b = team_bits_min(st, w0, sub, tmask, s - 36, s - 1);
            if (b != NTL_NONE) telo_start = b;
        }
    }
    return telo_start;
}

/* One 18-bp window of search_left/right_patterns (multi_pattern_step_*, NanoTel.R:496-575, :614, :676):
 * matchPattern on subseq(read, a, b) with the default fixed = TRUE, the window's own out-of-bounds rule, hits not
 * trimmed.  Returns false if no pattern hits.
 * The window and every alignment that can hit it (starts a - k .. b - m + 1 + k) fit one 32-bit word whose bit i is
 * position a - 1 + i: the word is cut out of two position words (the same for all lanes), letters outside [a, b] carry
 * zero masks (mismatches), and all alignment starts of a pattern are decided together by a two-plane mismatch counter
 * (Shift-And).  No shuffles: every lane computes the same value. */
__device__ __noinline__ bool step_window(const ReadView &rv, int a, int b, int k, bool use_tvr, int *min_start, int *max_end)
{
    *min_start = 0; *max_end = 0;
    if (b < a) return false;
    const int q0 = a - 1;                               /* position of bit 0 */
    const int rb = q0 - 1;                              /* its index in the packed stream (position p = bit p - 1) */
    u32 pl[4];
    {
        const int x = rb >> 5, sh = rb & 31;            /* rb = -1 (a = 1): x = -1, sh = 31 -> the stream shifted up by one */
        const int NP = rv.fmt ? 4 : 2;
        u32 w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int p = 0; p < 4; p++)
            if (p < NP) w[p] = __funnelshift_r(rv_raw(rv, p, x), rv_raw(rv, p, x + 1), sh);
        const u32 vm = ((2u << (b - a)) - 1u) << 1;     /* bits 1 .. b - a + 1: the window itself */
        if (rv.fmt == 0) {
            pl[0] = ~w[1] & ~w[0] & vm; pl[1] = ~w[1] & w[0] & vm; pl[2] = w[1] & w[0] & vm; pl[3] = w[1] & ~w[0] & vm;
        } else {
#pragma unroll
            for (int p = 0; p < 4; p++) pl[p] = w[p] & vm;
        }
    }
    const u32 vmask = ((2u << (b - a)) - 1u) << 1;
    bool any = false;
    int mn = 0, mx = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int np = pass == 0 ? c_prm.n_main : (use_tvr ? c_prm.n_tvr : 0);
        const int kk = pass == 0 ? k : 0;
        for (int p = 0; p < np; p++) {
            const ntl_dev_pat &pt = pass == 0 ? c_prm.main_pat[p] : c_prm.tvr_pat[p];
            const int m = pt.m;
            /* alignment starts a - kk .. b - m + 1 + kk  <->  bits 1 - kk .. b - a + 2 - m + kk */
            const int hi_bit = b - a + 2 - m + kk;
            if (hi_bit < 1 - kk) continue;
            u32 ones = 0u, twos = 0u;
            for (int j = 0; j < m; j++) {
                /* fixed = TRUE: the read's code must EQUAL the pattern letter's code */
                const u32 mA = pt.mux4[j][0], mC = pt.mux4[j][1], mG = pt.mux4[j][2], mT = pt.mux4[j][3];
                const u32 e = ~((pl[0] ^ mA) | (pl[1] ^ mC) | (pl[2] ^ mG) | (pl[3] ^ mT)) & vmask;
                const u32 x = ~(e >> j);
                twos |= ones & x;
                ones ^= x;
            }
            const u32 sm = ((2u << hi_bit) - 1u) & ~((1u << (1 - kk)) - 1u);
            const u32 h = (kk ? ~twos : ~(ones | twos)) & sm;
            if (h) {
                const int lo = q0 + __ffs((int)h) - 1;
                const int hi = q0 + 31 - __clz((int)h) + m - 1;
                if (!any || lo < mn) mn = lo;
                if (!any || hi > mx) mx = hi;
                any = true;
            }
        }
    }
    *min_start = mn; *max_end = mx;
    return any;
}

/* search_left_patterns (NanoTel.R:576-633) */
__device__ __noinline__ int search_left(const ReadView &rv, int start_index, int k, bool use_tvr)
{
    int subseq_start = start_index - 18 > 1 ? start_index - 18 : 1;
    int new_start = start_index;
    for (int i = 0; i < 4; i++) {
        const int curr_end = subseq_start + 17 < rv.L ? subseq_start + 17 : rv.L;
        int mn, mx;
        if (!step_window(rv, subseq_start, curr_end, k, use_tvr, &mn, &mx)) break;
        new_start = mn;
        const int nn = subseq_start - 9 > 1 ? subseq_start - 9 : 1;
        if (nn == subseq_start) break;
        subseq_start = nn;
    }
    return new_start;
}

/* search_right_patterns (NanoTel.R:635-697) */
__device__ __noinline__ int search_right(const ReadView &rv, int end_index, int k, bool use_tvr)
{
    int subseq_end = end_index + 18 < rv.L ? end_index + 18 : rv.L;
    int new_end = end_index;
    for (int i = 0; i < 4; i++) {
        const int curr_start = subseq_end - 17 > 1 ? subseq_end - 17 : 1;
        int mn, mx;
        if (!step_window(rv, curr_start, subseq_end, k, use_tvr, &mn, &mx)) break;
        new_end = mx;
        const int nn = subseq_end + 11 < rv.L ? subseq_end + 11 : rv.L;
        if (nn == subseq_end) break;
        subseq_end = nn;
    }
    return new_end;
}

/* =============================================================================================================
 * K3a: triage, one THREAD per read.  A read whose window tables hold no telomeric window on any track takes the
 * same path through find_telo_position_wraper every time: find_telo_position = (-1,-1) (:1026-1028), the
 * get_accurate_* calls return -1, width 1 < 100 sends it to find_left/right_telo, which return (-1,-1) as soon as a
 * window lies beyond max_diff = 200 of the chosen edge (:861, :921), search_right_patterns looks at s[1..18]
 * (:1141 with end_index 0) and, finding nothing, leaves (-1, 0): width 2 < 30, not telomeric (:1847).  The thread
 * verifies each of those conditions (integer window classes, one bit-parallel match over the first word) and writes
 * the record; every other read goes on the candidate list of the warp-per-read locate kernel.
 * ============================================================================================================= */
__device__ __forceinline__ bool triage_first_window_hit(u32 lo, u32 hi, int T)
{
    /* subseq(read, 1, 18), matchPattern with fixed = TRUE (NanoTel.R:614/:676 via :1141): positions outside the
     * window are mismatches, alignment starts 1-k .. 18-m+1+k */
    const u32 vm = 0x7fffeu;                                  /* bits 1..18 */
    const u32 pA = ~hi & ~lo & vm, pC = ~hi & lo & vm, pG = hi & lo & vm, pT = hi & ~lo & vm;
    bool any = false;
    for (int pass = 0; pass < 2; pass++) {
        const int np = pass == 0 ? c_prm.n_main : (T == 3 ? c_prm.n_tvr : 0);
        for (int p = 0; p < np; p++) {
            const ntl_dev_pat &pt = pass == 0 ? c_prm.main_pat[p] : c_prm.tvr_pat[p];
            u32 ones = 0u, twos = 0u;
            for (int j = 0; j < pt.m; j++) {
                const u32 nb = pt.nib[j];
                const u32 e = nb == 1u ? pA : nb == 2u ? pC : nb == 4u ? pG : nb == 8u ? pT : 0u;
                const u32 x = ~(e >> j);
                twos |= ones & x;
                ones ^= x;
            }
            /* exact alignments start in [1, 19-m]; with one mismatch (main patterns on tracks B, C) in [0, 20-m] */
            const int hi0 = 19 - pt.m;
            const u32 sm0 = hi0 >= 1 ? ((2u << hi0) - 2u) : 0u;
            any |= (~(ones | twos) & sm0) != 0u;
            if (pass == 0) {
                const int hi1 = 20 - pt.m;
                const u32 sm1 = hi1 >= 0 ? ((2u << hi1) - 1u) : 0u;
                any |= (~twos & sm1) != 0u;
            }
        }
    }
    return any;
}

/* Eight lanes per read ("team"), four reads per warp: the team walks the block counts of the read's LAST track
 * (see below) 32 16-byte groups at a time (four independent loads per lane, 512 contiguous bytes per team and step),
 * so the longest read costs n_win / 256 dependent steps, and the loads of a team coalesce. */
__global__ void __launch_bounds__(256) ntl_triage_kernel(const ntl_read_args a)
{
    const int lane = threadIdx.x & 31;
    const int sub = lane & 7;
    const u32 tmask = 0xffu << (lane & 24);
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    bool cand = false;
    int r = -1;
    if (slot < a.n_reads) {
        r = a.order[slot];
        const int S = c_prm.S, T = c_prm.n_tracks, Q = c_prm.Q, SG = c_prm.SG;
        const int L = a.len[r];
        const int n_win = ntl_nwin(L, S);
        const int nb = (L + SG - 1) / SG;
        const long long wo = a.cnt_off[r];
        const int fmt = a.fmt[r];
        /* first position word of the read (bits 0..17 hold s[1..18]); issued now so that its DRAM latency overlaps
         * the window loop.  For a 4-bit read these are other planes: read but not used. */
        const u32 *q0 = fmt ? a.arena4 + (size_t)a.woff[r] * 4 : a.arena2 + (size_t)a.woff[r] * 2;
        const u32 lo0 = __ldg(q0), hi0 = __ldg(q0 + 1);
        int status = 0;
        bool simple = true;
        if (a.pass != nullptr && a.pass[r] == 0) status = NTL_READ_FILTERED;
        else {
            simple = fmt == 0 && n_win >= 1 && n_win <= NTL_TRIAGE_MAX_WIN;
            if (simple)
                simple = c_prm.right_edge ? ((n_win >= 2 ? S : L) < L - 200) : (1 + (n_win - 1) * S > 200);
            if (simple) {
                /* any window with  !(count / width < min_density)  on any track?  Every read's block range starts
                 * on a multiple of 8 entries, so groups of 8 counts are 16-byte aligned. */
                const int thr_reg = c_prm.thr_reg;
                const u32 thr16 = (u32)thr_reg << 16;
                const int thr_last = (int)a.thr[L - (n_win - 1) * S];
                bool tel = false;
                /* The coverage of the tracks is nested (exact hits c <=1-mismatch hits c those + TVR hits), so a
                 * window's covered count can only grow from track to track: some track has a telomeric window iff
                 * the LAST track has one.  Only that plane is walked: half (a third) of the bytes. */
                const uint16_t *cm = a.cnt[T - 1] + wo;
                int k_done = 0;                                  /* regular windows decided by the vector walk */
                if (Q == 1) {
                    const uint4 *cv = reinterpret_cast<const uint4 *>(cm);
                    const int full = (n_win - 1) >> 3;           /* groups made of width-S windows only */
                    constexpr int NJ = 4;                        /* independent 16-byte loads in flight per lane */
                    for (int g0 = 0; g0 < full; g0 += 8 * NJ) {
                        uint4 vv[NJ];
#pragma unroll
                        for (int j = 0; j < NJ; j++) {
                            const int g = g0 + 8 * j + sub;
                            vv[j] = g < full ? __ldg(cv + g) : make_uint4(0u, 0u, 0u, 0u);
                        }
#pragma unroll
                        for (int j = 0; j < NJ; j++) {
                            /* two counts per word: the upper one decides  x >= thr << 16, the lower one after a shift */
                            const u32 x[4] = {vv[j].x, vv[j].y, vv[j].z, vv[j].w};
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                tel |= x[q] >= thr16;
                                tel |= (x[q] << 16) >= thr16;
                            }
                        }
                    }
                    k_done = full << 3;
                }
                /* the remaining windows, one per lane and step: regular windows are Q blocks, the last one (own width
                 * and threshold) is every block that is left */
                for (int k = k_done + sub; k < n_win; k += 8) {
                    const int b0 = k * Q, b1 = k == n_win - 1 ? nb : b0 + Q;
                    int c = 0;
                    for (int b = b0; b < b1; b++) c += (int)cm[b];
                    tel |= c >= (k == n_win - 1 ? thr_last : thr_reg);
                }
                simple = (__ballot_sync(tmask, tel) & tmask) == 0u;
            }
            if (simple) simple = !triage_first_window_hit(lo0 << 1, hi0 << 1, T);
            if (!simple) cand = true;
        }
        if (!cand) {
            /* the 64-byte record goes out as four 16-byte stores from lanes 0..3 of the team */
            alignas(16) ntl_read_result o;       /* stored below as four 16-byte pieces */
            o.status = status;
            o.n_win = n_win > 0 ? n_win : 0;
            for (int t = 0; t < 3; t++) {
                const bool live = status == 0 && t < T;
                o.track[t].start = live ? -1 : 0; o.track[t].end = 0; o.track[t].density = 0.0;
            }
            o.win_offset = wo;
            if (sub < 4)
                reinterpret_cast<uint4 *>(reinterpret_cast<ntl_read_result *>(a.results) + r)[sub] =
                    reinterpret_cast<const uint4 *>(&o)[sub];
            if (a.stages != nullptr && status == 0 && sub < T) {
                ntl_stage s;
                s.coarse_start = s.coarse_end = s.acc_start = s.acc_end = s.edge_start = s.edge_end = -1;
                s.acc_density = 0.0;
                reinterpret_cast<ntl_stage *>(a.stages)[(size_t)r * 3 + sub] = s;
            }
        }
    }
    /* warp-aggregated append to the candidate list (one entry per team) */
    const u32 cm = __ballot_sync(NTL_FULL, cand && sub == 0);
    if (cm) {
        int base = 0;
        const int leader = __ffs((int)cm) - 1;
        if (lane == leader) base = (int)atomicAdd(&a.counters[0], (u32)__popc(cm));
        base = __shfl_sync(NTL_FULL, base, leader);
        if (cand && sub == 0) {
            const int pos = base + __popc(cm & ((1u << lane) - 1u));
            a.cand[pos] = r;
            reinterpret_cast<int4 *>(a.cand_state)[pos] = make_int4(0, 0, 0, 0);
        }
    }
}

/* =============================================================================================================
 * K3b: full locator, one THREAD per (candidate read, track); the thread that completes a read's last track writes
 * the record head.
 * ============================================================================================================= */
/* One item per team, four teams per warp.  The four items of a warp are walked in PHASES with a warp barrier between
 * them: inside a phase the teams run the same code (the coverage recomputation has fixed trip counts), and teams that
 * diverged in a data-dependent phase (the window scans) meet again at the next barrier instead of executing the rest
 * of their items one after the other.  valid = false: a team without an item (it only keeps the barriers company). */
__device__ void locate_items(const ntl_read_args &a, bool valid, int r, int t, int *state, int sub, u32 tmask)
{
    ReadView rv;
    WinTab w;
    rv.L = 0; rv.fmt = 0; rv.base = nullptr; rv.n_words = 0; rv.n_raw = 0; rv.ccov = nullptr; rv.cwb = nullptr;
    w.cnt = nullptr; w.n = 0; w.nb = 0; w.Q = 1; w.SG = 1; w.S = 1; w.L = 0; w.thr_reg = 0; w.thr_last = 0;
    const int S = c_prm.S, T = c_prm.n_tracks;
    int n_win = 0, status = 0;
    ntl_track out;
    out.start = 0; out.end = 0; out.density = 0.0;
    bool err = false;
    int width = 0, ts = -1, te = -1, cs = -1, ce = -1, as = -1, ae = -1, s2 = 0, e2 = 0;
    double acc_density = 0.0;
    const int k = t >= 1 ? 1 : 0;
    const bool use_tvr = t == 2;
    ntl_stage *stg = nullptr;

    /* ---- phase 0: tables, coarse interval (window scans: data dependent) */
    if (valid) {
        stg = a.stages ? reinterpret_cast<ntl_stage *>(a.stages) + (size_t)r * 3 : nullptr;
        rv_init(rv, a, r);
        n_win = ntl_nwin(rv.L, S);
        status = rv.fmt ? NTL_READ_IUPAC : 0;
        if (n_win <= 0) status |= NTL_READ_NO_WINDOWS;
        w.cnt = a.cnt[t] + a.cnt_off[r]; w.n = n_win > 0 ? n_win : 0; w.S = S; w.L = rv.L;
        w.Q = c_prm.Q; w.SG = c_prm.SG; w.nb = (rv.L + c_prm.SG - 1) / c_prm.SG;
        w.thr_reg = c_prm.thr_reg;
        w.thr_last = w.n > 0 ? (int)a.thr[wt_end(w, w.n - 1) - wt_start(w, w.n - 1) + 1] : 0;
        find_telo_position(w, 3.0, 2.0, &ts, &te);                                     /* :1084-1086 */
        /* the coarse interval is made of whole windows: its density needs no coverage recomputation */
        const double telo_density = density_of(rv, w, t, ts, te, sub, tmask);          /* :1099 */
        const int num_rows = (te - ts + 1) / S;                                        /* :1103 */
        if (telo_density < 0.85 && num_rows > 5) {                                     /* :1104-1110 */
            const double min_rows = num_rows <= 7 ? (double)(num_rows - 2) : 7.0;
            const double min_score = 0.6 * min_rows;
            find_telo_position(w, min_rows, min_score, &ts, &te);
        }
        cs = ts; ce = te;
    }
    __syncwarp();
    /* ---- phase 1, 2: accurate start / end (one coverage block each) */
    int start_acc = -1, end_acc = -1;
    if (valid) start_acc = get_accurate_start(rv, t, ts, sub, tmask);                  /* :1119 */
    __syncwarp();
    if (valid) end_acc = get_accurate_end(rv, t, te, sub, tmask);                      /* :1120 */
    __syncwarp();
    if (valid) {
        if (start_acc > end_acc) end_acc = start_acc;                                  /* :1122-1124 */
        ts = start_acc; te = end_acc;
        as = ts; ae = te;
        if (stg) acc_density = density_of(rv, w, t, ts, te, sub, tmask);
        if (te - ts + 1 < 100) {                                                       /* :1129-1136 */
            if (c_prm.right_edge) {
                if (w.n == 0) err = true;                                              /* R stops at :859-861 */
                else find_right_telo(w, &ts, &te);
            } else find_left_telo(w, &ts, &te);
        }
        if (!err && stg && sub == 0) {
            ntl_stage sg;
            sg.coarse_start = cs; sg.coarse_end = ce; sg.acc_start = as; sg.acc_end = ae;
            sg.edge_start = ts; sg.edge_end = te; sg.acc_density = acc_density;
            stg[t] = sg;
        }
    }
    __syncwarp();
    /* ---- phase 3, 4: the 18-bp re-match searches */
    if (valid && !err) e2 = te < rv.L ? search_right(rv, te + 1, k, use_tvr) : te;      /* :1140-1144 */
    __syncwarp();
    if (valid && !err) s2 = ts > 1 ? search_left(rv, ts - 1, k, use_tvr) : ts;          /* :1145-1149 */
    __syncwarp();
    /* ---- phase 5: final density (:1840-1844) */
    if (valid && !err) {
        if (e2 < s2 - 1) err = true;                                                   /* IRanges() would stop */
        else {
            out.start = s2; out.end = e2;
            out.density = density_of(rv, w, t, s2, e2, sub, tmask);
            width = e2 - s2 + 1;
        }
    }
    __syncwarp();
    /* ---- this track is done; the team that completes the read's last track writes the record head:
     *      keep iff max interval width >= 30 over the tracks (:1847, :1857) */
    if (!valid || sub != 0) return;
    ntl_read_result *res = reinterpret_cast<ntl_read_result *>(a.results) + r;
    res->track[t] = out;
    /* state = {tracks done, width or -1 (error) of track 0, 1, 2}: own slot, fence, one atomic; the team that
     * arrives last reads the other slots past L1 */
    *reinterpret_cast<volatile int *>(&state[1 + t]) = err ? -1 : width;
    __threadfence();
    if (atomicAdd(&state[0], 1) == T - 1) {
        int any_err = 0, mw = 0;
        for (int tt = 0; tt < T; tt++) {
            const int wdt = __ldcg(&state[1 + tt]);
            if (wdt < 0) any_err = 1; else if (wdt > mw) mw = wdt;
        }
        if (any_err) status |= NTL_READ_REF_ERROR;
        else if (mw >= 30) status |= NTL_READ_KEEP;
        res->status = status;
        res->n_win = n_win > 0 ? n_win : 0;
        res->win_offset = a.cnt_off[r];
        ntl_track zero;
        zero.start = 0; zero.end = 0; zero.density = 0.0;
        for (int tt = T; tt < 3; tt++) res->track[tt] = zero;
    }
}

__global__ void __launch_bounds__(64) ntl_locate_kernel(const ntl_read_args a)
{
    /* work item = (candidate read, track): the tracks of a read are independent until the keep rule;
     * cand_state[c] = {tracks done, width (or -1: error) of track 0, 1, 2} joins them.  Reads that the filter
     * dropped never get here (the triage kernel writes their record). */
    const int T = c_prm.n_tracks;
    const int n_items = (int)a.counters[0] * T;
    const int lane = threadIdx.x & 31, sub = lane & (NTL_TEAM - 1);
    const u32 tmask = 0xffu << (lane & 24);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * 4; base < n_items; base += n_warps * 4) {     /* warp-uniform: every team keeps the barriers */
        const int i = base + (lane >> 3);
        const bool valid = i < n_items;
        const int c = valid ? i / T : 0;
        locate_items(a, valid, valid ? a.cand[c] : 0, valid ? i - c * T : 0, a.cand_state + 4 * (size_t)c, sub, tmask);
    }
}

/* =============================================================================================================
 * K2g: generic scan (any pattern set, any subseq_length; patterns in __constant__ memory).  One warp per read, longest
 * first; coverage words come from warp_cov (31 words per step), every word adds its covered bits to the blocks it
 * overlaps with 32-bit atomics on the uint16 pairs (the planes are zeroed before the launch).
 * ============================================================================================================= */
__global__ void __launch_bounds__(256) ntl_scan_generic_kernel(const ntl_read_args a)
{
    const int lane = threadIdx.x & 31;
    const int T = c_prm.n_tracks, SG = c_prm.SG;
    for (;;) {
        int i = 0;
        if (lane == 0) i = (int)atomicAdd(&a.counters[4], 1u);
        i = __shfl_sync(NTL_FULL, i, 0);
        if (i >= a.n_reads) break;
        const int r = a.order[i];
        if (a.pass != nullptr && a.pass[r] == 0) continue;
        ReadView rv;
        rv_init(rv, a, r);
        const int L = rv.L;
        for (int t = 0; t < T; t++) {
            u32 *plane = reinterpret_cast<u32 *>(a.cnt[t] + a.cnt_off[r]);      /* cnt_off is a multiple of 8 entries */
            for (int wb = -1; ((wb + 1) << 5) <= L; wb += 31) {
                u32 hs;
                const u32 cov = warp_cov(rv, t, wb, lane, &hs);
                if (lane == 0 || cov == 0u) continue;
                const int p0 = (wb + lane) << 5;                                 /* position of bit 0 */
                const int pa = p0 < 1 ? 1 : p0, pb = p0 + 31 > L ? L : p0 + 31;
                for (int blk = (pa - 1) / SG; blk <= (pb - 1) / SG; blk++) {
                    const int lo = blk * SG + 1 > pa ? blk * SG + 1 : pa, hi = (blk + 1) * SG < pb ? (blk + 1) * SG : pb;
                    const u32 m = (NTL_FULL >> (31 - (hi - p0))) & (NTL_FULL << (lo - p0));
                    const u32 c = (u32)__popc(cov & m);
                    if (c) atomicAdd(plane + (blk >> 1), c << (16 * (blk & 1)));
                }
            }
        }
    }
}

/* =============================================================================================================
 * launchers (called from ntl_api.cpp)
 * ============================================================================================================= */
extern "C" cudaError_t ntl_k_set_params(const ntl_dev_params *p, cudaStream_t st)
{
    return cudaMemcpyToSymbolAsync(c_prm, p, sizeof(ntl_dev_params), 0, cudaMemcpyHostToDevice, st);
}

extern "C" cudaError_t ntl_k_scan_generic(const ntl_read_args *a, int grid, cudaStream_t st)
{
    if (a->n_reads <= 0) return cudaSuccess;
    ntl_scan_generic_kernel<<<grid, 256, 0, st>>>(*a);
    return cudaGetLastError();
}

extern "C" cudaError_t ntl_k_scan_generic_occupancy(int *blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, ntl_scan_generic_kernel, 256, 0);
}

extern "C" cudaError_t ntl_k_filter(const ntl_read_args *a, cudaStream_t st)
{
    if (a->n_reads <= 0) return cudaSuccess;
    ntl_filter_kernel<<<(a->n_reads * 8 + 255) / 256, 256, 0, st>>>(*a);     /* 8 lanes per read */
    return cudaGetLastError();
}

extern "C" cudaError_t ntl_k_items(const uint8_t *active, int n_items, int32_t *items, uint32_t *counter, cudaStream_t st)
{
    if (n_items <= 0) return cudaSuccess;
    ntl_items_kernel<<<(n_items + 255) / 256, 256, 0, st>>>(active, n_items, items, counter);
    return cudaGetLastError();
}

extern "C" cudaError_t ntl_k_triage(const ntl_read_args *a, cudaStream_t st)
{
    if (a->n_reads <= 0) return cudaSuccess;
    ntl_triage_kernel<<<(a->n_reads * 8 + 255) / 256, 256, 0, st>>>(*a);     /* 8 lanes per read */
    return cudaGetLastError();
}

extern "C" cudaError_t ntl_k_locate(const ntl_read_args *a, int grid, cudaStream_t st)
{
    if (a->n_reads <= 0) return cudaSuccess;
    /* one thread per (candidate, track); the number of candidates is only known on the device: enough small CTAs for
     * every read to be one (those beyond the candidate list leave at once), spread over all SMs */
    const long long want = ((long long)a->n_reads * 3 * 8 + 63) / 64;        /* eight lanes per item */
    ntl_locate_kernel<<<(int)(want < grid ? want : grid), 64, 0, st>>>(*a);
    return cudaGetLastError();
}

/* Block counts of the kept reads only, packed for the device-to-host copy: entry i of `list` = {read, first
 * destination element}; a read's T rows of ceil8(blocks) elements follow each other.  One warp per entry, 16-byte
 * moves (source and destination rows start on multiples of 8 elements). */
__global__ void __launch_bounds__(128) ntl_gather_windows_kernel(const ntl_read_args a, const int64_t *list, int n_list,
                                                                 uint16_t *dst, int T)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_list) return;
    const int r = (int)list[2 * i];
    const int groups = ((a.len[r] + c_prm.SG - 1) / c_prm.SG + 7) >> 3;
    for (int t = 0; t < T; t++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.cnt[t] + a.cnt_off[r]);
        uint4 *out = reinterpret_cast<uint4 *>(dst + list[2 * i + 1]) + (size_t)t * groups;
        for (int g = lane; g < groups; g += 32) out[g] = __ldg(src + g);
    }
}

extern "C" cudaError_t ntl_k_gather_windows(const ntl_read_args *a, const int64_t *list, int n_list, uint16_t *dst, int T,
                                            cudaStream_t st)
{
    if (n_list <= 0) return cudaSuccess;
    ntl_gather_windows_kernel<<<(n_list * 32 + 127) / 128, 128, 0, st>>>(*a, list, n_list, dst, T);
    return cudaGetLastError();
}

extern "C" cudaError_t ntl_k_locate_occupancy(int *blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, ntl_locate_kernel, 64, 0);
}
