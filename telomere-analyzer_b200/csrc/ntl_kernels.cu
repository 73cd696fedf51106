/*
 * ntl_kernels.cu -- sm_100a kernels of libnanotel_b200 other than the compile-time-specialised span scan (ntl_scan.cuh):
 *   K2g ntl_scan_generic_kernel  any pattern set / any subseq_length, patterns in __constant__ memory: the path taken
 *                              when no specialised build exists (NTL_OPT_NO_JIT, no NVRTC and no precompiled cubin,
 *                              or a subseq_length without a span geometry); one warp per read, block counts by atomics
 *   K4  ntl_filter_kernel      --use_filter edge filter            (filter_reads/filter_density, NanoTel.R:2083-2163);
 *                              also marks the spans / work items of the reads it keeps for the span scan
 *       ntl_items_kernel       compacts the marked work items into the list the span scan walks
 *   K3a ntl_triage_kernel      one thread per read: proves that a read has no telomeric window on any track and no
 *                              hit in its first 18 bases and writes its (trivial) record, or hands it to K3b
 *   K3b ntl_locate_kernel      a team of eight lanes per (candidate read, track): locator and refinement (find_telo_position_wraper
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

#ifdef NTL_K3_CLOCK
/* development aid (build.py with NTL_EXTRA_NVCC_FLAGS=-DNTL_K3_CLOCK): cycles per phase of the locate kernel, per warp:
 * [phase][0] sum, [1] max, [2] samples */
__device__ unsigned long long g_k3clk[16][3];
#define K3CLK(ph) do { __syncwarp(); const long long now_ = clock64(); if ((threadIdx.x & 31) == 0) { \
    atomicAdd(&g_k3clk[ph][0], (unsigned long long)(now_ - clk_)); atomicMax(&g_k3clk[ph][1], (unsigned long long)(now_ - clk_)); \
    atomicAdd(&g_k3clk[ph][2], 1ull); } clk_ = clock64(); } while (0)
#define K3CLK_BEGIN long long clk_ = clock64()
extern "C" int ntl_debug_k3_clock(unsigned long long *out)
{
    if (cudaMemcpyFromSymbol(out, g_k3clk, sizeof g_k3clk) != cudaSuccess) return -1;
    unsigned long long z[16][3] = {};
    cudaMemcpyToSymbol(g_k3clk, z, sizeof z);
    return 16;
}
#else
#define K3CLK(ph) do { } while (0)
#define K3CLK_BEGIN do { } while (0)
#endif

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
 * K3: locator.  The window state machines of the reference are sequential scans with data-dependent exits, and a
 * candidate needs only a few hundred bases of locally recomputed coverage.  Measured over two rounds: a warp per
 * (read, track) item spends its 32 lanes on the same scalar work (4 300 warp-instructions per item, 41 % issue
 * utilisation, bound by dependent latency); one thread per item divides the instruction count by 32 but leaves every
 * load -> divide -> compare chain fully exposed.  The kernel below gives an item a TEAM of eight lanes: the state
 * machines run on all eight alike, and everything that has data parallelism in it is spread over them -- the class-bit
 * searches (512 windows per step), the window counts and their densities (eight windows per load and division), the
 * coverage words recomputed from the read (one per lane), the whole-block sums.
 * ============================================================================================================= */
/* one copy of the IEEE double division (a ~80-instruction sequence) for the whole locate kernel */
__device__ __noinline__ double k3_div(double a, double b) { return a / b; }

struct WinTab {                 /* the window table of one track (analyze_subtelos :737-764), never materialised */
    const uint16_t *cnt;        /* covered bases per block of SG positions (K2); window k = blocks k Q .. k Q + Q - 1, */
    int n, nb, Q, SG, S, L;     /* the last window = every remaining block (n windows, nb blocks)                     */
    int thr_reg, thr_last;      /* smallest telomeric count of a regular window / of this read's last window */
    const double *dens;         /* dens[c] = (double)c / (double)S: the density of a regular window (host table)       */
    const u32 *cls;             /* the track's class-bit plane (ntl_dev.h) as 32-bit words                            */
    long long bit0;             /* bit address of the read's block 0                                                  */
    int bps;                    /* blocks per class byte                                                              */
};
/* bit address of block b of the read / block of a bit address */
__device__ __forceinline__ long long wt_bit_of_block(const WinTab &w, int b)
{
    return w.bps == 8 ? w.bit0 + b : w.bit0 + (long long)(b / w.bps) * 8 + b % w.bps;
}
__device__ __forceinline__ int wt_block_of_bit(const WinTab &w, long long a)
{
    const int rel = (int)(a - w.bit0);
    return w.bps == 8 ? rel : (rel >> 3) * w.bps + (rel & 7);
}

/* First / last set bit of a class plane inside the bit addresses [a_lo, a_hi] (-1: none), searched by a TEAM of eight
 * lanes: 512 bits per step, two independent 32-bit loads per lane.  Every lane of the team returns the same value. */
#define NTL_TEAM 8
__device__ __forceinline__ u32 cls_word(const u32 *pl, long long wi, long long a_lo, long long a_hi)
{
    if (wi < (a_lo >> 5) || wi > (a_hi >> 5)) return 0u;
    u32 v = __ldg(pl + wi);
    if (wi == (a_lo >> 5)) v &= NTL_FULL << (int)(a_lo & 31);
    if (wi == (a_hi >> 5)) v &= NTL_FULL >> (31 - (int)(a_hi & 31));
    return v;
}
__device__ __noinline__ long long team_first_bit(const u32 *pl, long long a_lo, long long a_hi, int sub, u32 tmask, int lane)
{
    if (a_hi < a_lo) return -1;
    for (long long w0 = a_lo >> 5; w0 <= (a_hi >> 5); w0 += 2 * NTL_TEAM) {
        const u32 v0 = cls_word(pl, w0 + sub, a_lo, a_hi), v1 = cls_word(pl, w0 + NTL_TEAM + sub, a_lo, a_hi);
        const u32 m0 = (__ballot_sync(tmask, v0 != 0u) >> (lane & 24)) & 0xffu;
        const u32 m1 = (__ballot_sync(tmask, v1 != 0u) >> (lane & 24)) & 0xffu;
        if (m0) {
            const int src = __ffs((int)m0) - 1;
            return ((w0 + src) << 5) + __ffs((int)__shfl_sync(tmask, v0, src, NTL_TEAM)) - 1;
        }
        if (m1) {
            const int src = __ffs((int)m1) - 1;
            return ((w0 + NTL_TEAM + src) << 5) + __ffs((int)__shfl_sync(tmask, v1, src, NTL_TEAM)) - 1;
        }
    }
    return -1;
}
__device__ __noinline__ long long team_last_bit(const u32 *pl, long long a_lo, long long a_hi, int sub, u32 tmask, int lane)
{
    if (a_hi < a_lo) return -1;
    for (long long w0 = a_hi >> 5; w0 >= (a_lo >> 5); w0 -= 2 * NTL_TEAM) {
        const u32 v0 = cls_word(pl, w0 - sub, a_lo, a_hi), v1 = cls_word(pl, w0 - NTL_TEAM - sub, a_lo, a_hi);
        const u32 m0 = (__ballot_sync(tmask, v0 != 0u) >> (lane & 24)) & 0xffu;
        const u32 m1 = (__ballot_sync(tmask, v1 != 0u) >> (lane & 24)) & 0xffu;
        if (m0) {
            const int src = __ffs((int)m0) - 1;                      /* the lowest lane holds the highest word */
            return ((w0 - src) << 5) + 31 - __clz((int)__shfl_sync(tmask, v0, src, NTL_TEAM));
        }
        if (m1) {
            const int src = __ffs((int)m1) - 1;
            return ((w0 - NTL_TEAM - src) << 5) + 31 - __clz((int)__shfl_sync(tmask, v1, src, NTL_TEAM));
        }
    }
    return -1;
}
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
    /* get_sub_density (NanoTel.R:467): count / width in double; a regular window has width S, and count / S comes
     * from the table the host filled with that same IEEE division (a cached load instead of ~80 instructions) */
    if (k < w.n - 1) return __ldg(w.dens + count);
    return k3_div((double)count, (double)(wt_end(w, k) - wt_start(w, k) + 1));
}

/* First telomeric window at or after k0 (n if none) / last telomeric window at or before k0 (-1 if none).  The
 * reference's scans walk window by window; between two telomeric windows nothing happens but resets, so the walk may
 * jump: the regular windows are searched through the class bits of their blocks (exact for Q == 1; for Q > 1 a set
 * bit names a window whose count is then checked), 512 blocks per team step; the last window (own width and
 * threshold) is always decided from its count. */
__device__ __noinline__ int next_telo_fwd(const WinTab &w, int k0, int sub, u32 tmask, int lane)
{
    const int n = w.n, nr = n - 1;                                       /* regular windows 0 .. nr - 1 */
    int k = k0 < 0 ? 0 : k0;
    while (k < nr) {
        const long long a = team_first_bit(w.cls, wt_bit_of_block(w, k * w.Q), wt_bit_of_block(w, nr * w.Q - 1), sub, tmask, lane);
        if (a < 0) { k = nr; break; }
        const int kk = wt_block_of_bit(w, a) / w.Q;
        if (w.Q == 1 || wt_count(w, kk) >= w.thr_reg) return kk;
        k = kk + 1;
    }
    if (k == nr && nr >= 0 && wt_telo_count(w, nr, wt_count(w, nr))) return nr;
    return n;
}

__device__ __noinline__ int next_telo_bwd(const WinTab &w, int k0, int sub, u32 tmask, int lane)
{
    const int n = w.n;
    int k = k0 >= n ? n - 1 : k0;
    if (k < 0) return -1;
    if (k == n - 1) { if (wt_telo_count(w, k, wt_count(w, k))) return k; k--; }
    while (k >= 0) {
        const long long a = team_last_bit(w.cls, wt_bit_of_block(w, 0), wt_bit_of_block(w, k * w.Q + w.Q - 1), sub, tmask, lane);
        if (a < 0) return -1;
        const int kk = wt_block_of_bit(w, a) / w.Q;
        if (w.Q == 1 || wt_count(w, kk) >= w.thr_reg) return kk;
        k = kk - 1;
    }
    return -1;
}

/* Eight consecutive windows (base, base + dir, ...), one per lane of the team: count and density (get_sub_density,
 * NanoTel.R:467) are fetched / divided by all lanes at once and handed out by shuffles, so a run of telomeric windows
 * costs one load and one division per eight windows instead of one dependent pair per window. */
struct WinBuf { int base, dir, c; double d; };
__device__ __forceinline__ void wb_get(WinBuf &b, const WinTab &w, int k, int dir, int sub, u32 tmask, int *c, double *d)
{
    int idx = (k - b.base) * dir;
    if (b.dir != dir || idx < 0 || idx >= NTL_TEAM) {
        b.base = k; b.dir = dir; idx = 0;
        const int kk = k + dir * sub;
        b.c = 0; b.d = 0.0;
        if (kk >= 0 && kk < w.n) { b.c = wt_count(w, kk); b.d = wt_density_of_count(w, kk, b.c); }
    }
    *c = __shfl_sync(tmask, b.c, idx, NTL_TEAM);
    *d = __shfl_sync(tmask, b.d, idx, NTL_TEAM);
}

/* find_telo_position (NanoTel.R:973-1077): forward scan for the first run of telomeric windows with in_a_row >= R
 * and score >= T, then the backward scan for the end.  Windows are 0-based here; non-telomeric windows only reset
 * the run state, so stretches of them are jumped over (next_telo_fwd / _bwd). */
__device__ __noinline__ void find_telo_position(const WinTab &w, double R, double T, int *ps, int *pe, int sub, u32 tmask, int lane)
{
    const int n = w.n;
    double score = 0.0;
    int start = -1, end = -1, in_a_row = 0;
    int end_position = 0;                                                /* 1-based i + 1 (:1022) */
    WinBuf wb;
    wb.base = 0; wb.dir = 0; wb.c = 0; wb.d = 0.0;
    /* leading non-telomeric windows only reset a state that is still the initial one: start at the first telomeric
     * window (with the reference's R >= 3 the degenerate exit below cannot fire on them) */
    const bool jump = R > 0.0 || T > 0.0;
    for (int k = jump ? next_telo_fwd(w, 0, sub, tmask, lane) : 0; k < n;) {        /* :1003-1025 */
        int c; double dk;
        wb_get(wb, w, k, 1, sub, tmask, &c, &dk);
        if (!wt_telo_count(w, k, c)) {
            score = 0.0; start = -1; in_a_row = 0;
            if (0.0 >= R && 0.0 >= T) { end_position = k + 2; break; }   /* never with the reference's R >= 3 */
            k = next_telo_fwd(w, k + 1, sub, tmask, lane);
            continue;
        }
        in_a_row += 1;
        score = score + dk;                                              /* :1014 */
        if (start == -1) start = wt_start(w, k);
        if ((double)in_a_row >= R && score >= T) { end_position = k + 2; break; }
        k++;
    }
    if (end_position == 0) { *ps = -1; *pe = -1; return; }               /* :1026-1028 */
    end = -1; score = 0.0; in_a_row = 0;
    if ((double)end_position >= (double)n - R + 1.0) {                   /* :1037-1044 */
        /* i = n; while (i > end_position && window i is not telomeric) i--  (1-based) */
        int i = next_telo_bwd(w, n - 1, sub, tmask, lane) + 1;                             /* 1-based, 0 if none */
        if (i < end_position) i = end_position;
        end = wt_end(w, (i < n ? i : n) - 1);                            /* end_position may be n + 1: the loop does not run */
    } else {                                                             /* :1046-1068 */
        /* the same from the other end: trailing non-telomeric windows only reset the initial state */
        for (int i = next_telo_bwd(w, n - 1, sub, tmask, lane) + 1; i >= end_position;) {
            int c; double dk;
            wb_get(wb, w, i - 1, -1, sub, tmask, &c, &dk);
            if (!wt_telo_count(w, i - 1, c)) {
                score = 0.0; end = -1; in_a_row = 0;
                const int j = next_telo_bwd(w, i - 2, sub, tmask, lane) + 1;               /* 1-based */
                i = j;
                continue;
            }
            in_a_row += 1;
            score = score + dk;
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
 * and where coverage has to be recomputed from the read every lane takes one word (team_cov2 below): coverage of track
 * t (0 exact, 1 one mismatch, 2 one mismatch + TVR; the union of the trimmed hit intervals of get_density_iranges,
 * NanoTel.R:308-397) for virtual words w0 - 1 + sub (bit b of word w = position 32 w + b): hit starts of every pattern
 * in (previous word, word) dilated by the pattern length.  Lane 0 lacks the hits of the word before its own: its word
 * (w0 - 1) is only the look-behind; words w0 .. w0 + 6 are complete.  hs = exact hit starts of main pattern 0 in the
 * lane's word (the raw hit list of NanoTel.R:349-354). */
/* hit starts of one pattern in TWO words at once (word_hits for two independent positions: twice the instruction-level
 * parallelism, one walk over the pattern's letter masks) */
__device__ __forceinline__ void word_hits2(const ntl_dev_pat &pt, const u32 (&pwa)[4], const u32 (&pna)[4], const u32 (&pwb)[4],
                                           const u32 (&pnb)[4], u32 *exa, u32 *lea, u32 *exb, u32 *leb)
{
    u32 o1 = 0u, t1 = 0u, o2 = 0u, t2 = 0u;
    const bool fx = pt.fixed != 0;
#pragma unroll 1
    for (int j = 0; j < pt.m; j++) {
        const u32 mA = pt.mux4[j][0], mC = pt.mux4[j][1], mG = pt.mux4[j][2], mT = pt.mux4[j][3];
        u32 ewa, ena, ewb, enb;
        if (fx) {
            ewa = ~((pwa[0] ^ mA) | (pwa[1] ^ mC) | (pwa[2] ^ mG) | (pwa[3] ^ mT));
            ena = ~((pna[0] ^ mA) | (pna[1] ^ mC) | (pna[2] ^ mG) | (pna[3] ^ mT));
            ewb = ~((pwb[0] ^ mA) | (pwb[1] ^ mC) | (pwb[2] ^ mG) | (pwb[3] ^ mT));
            enb = ~((pnb[0] ^ mA) | (pnb[1] ^ mC) | (pnb[2] ^ mG) | (pnb[3] ^ mT));
        } else {
            ewa = (pwa[0] & mA) | (pwa[1] & mC) | (pwa[2] & mG) | (pwa[3] & mT);
            ena = (pna[0] & mA) | (pna[1] & mC) | (pna[2] & mG) | (pna[3] & mT);
            ewb = (pwb[0] & mA) | (pwb[1] & mC) | (pwb[2] & mG) | (pwb[3] & mT);
            enb = (pnb[0] & mA) | (pnb[1] & mC) | (pnb[2] & mG) | (pnb[3] & mT);
        }
        const u32 xa = ~__funnelshift_r(ewa, ena, j), xb = ~__funnelshift_r(ewb, enb, j);
        t1 |= o1 & xa; o1 ^= xa;
        t2 |= o2 & xb; o2 ^= xb;
    }
    *exa = ~(o1 | t1); *lea = ~t1;
    *exb = ~(o2 | t2); *leb = ~t2;
}

/* team_cov for two independent word ranges (w0a - 1 + sub and w0b - 1 + sub) in one pass: the loads of both are in
 * flight together and the two dependency chains interleave -- the locate kernel is bound by latency, not by issue. */
__device__ __noinline__ void team_cov2(const ReadView &rv, int t, int w0a, int w0b, int sub, u32 tmask, u32 *cova, u32 *hsa,
                                       u32 *covb, u32 *hsb)
{
    const int wa = w0a - 1 + sub, wb = w0b - 1 + sub;
    u32 pwa[4], pna[4], pwb[4], pnb[4];
    word_planes(rv, wa, pwa);
    word_planes(rv, wa + 1, pna);
    word_planes(rv, wb, pwb);
    word_planes(rv, wb + 1, pnb);
    u32 ca = 0u, cb = 0u, h0a = 0u, h0b = 0u;
#pragma unroll 1
    for (int p = 0; p < c_prm.n_main; p++) {
        u32 exa, lea, exb, leb;
        word_hits2(c_prm.main_pat[p], pwa, pna, pwb, pnb, &exa, &lea, &exb, &leb);
        if (p == 0) { h0a = exa; h0b = exb; }
        const u32 Ha = t >= 1 ? lea : exa, Hb = t >= 1 ? leb : exb;
        u32 Hpa = __shfl_up_sync(tmask, Ha, 1, NTL_TEAM), Hpb = __shfl_up_sync(tmask, Hb, 1, NTL_TEAM);
        if (sub == 0) { Hpa = 0u; Hpb = 0u; }
#pragma unroll 1
        for (int j = 0; j < c_prm.main_pat[p].m; j++) { ca |= __funnelshift_l(Hpa, Ha, j); cb |= __funnelshift_l(Hpb, Hb, j); }
    }
    if (t == 2) {
#pragma unroll 1
        for (int p = 0; p < c_prm.n_tvr; p++) {
            u32 exa, lea, exb, leb;
            word_hits2(c_prm.tvr_pat[p], pwa, pna, pwb, pnb, &exa, &lea, &exb, &leb);
            u32 Hpa = __shfl_up_sync(tmask, exa, 1, NTL_TEAM), Hpb = __shfl_up_sync(tmask, exb, 1, NTL_TEAM);
            if (sub == 0) { Hpa = 0u; Hpb = 0u; }
#pragma unroll 1
            for (int j = 0; j < c_prm.tvr_pat[p].m; j++) { ca |= __funnelshift_l(Hpa, exa, j); cb |= __funnelshift_l(Hpb, exb, j); }
        }
    }
    *hsa = h0a; *hsb = h0b;
    *cova = (wa < 0 || wa >= rv.n_words) ? 0u : ca & ntl_valid_word(wa << 5, rv.L);      /* trim() to [1, L] */
    *covb = (wb < 0 || wb >= rv.n_words) ? 0u : cb & ntl_valid_word(wb << 5, rv.L);
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

/* covered positions of track t inside [lo1, hi1] and [lo2, hi2] (each 1 <= lo <= hi <= L, or empty: hi < lo),
 * recomputed from the read, 7 words of each range per step */
__device__ __noinline__ int local_count2(const ReadView &rv, int t, int lo1, int hi1, int lo2, int hi2, int sub, u32 tmask)
{
    int total = 0;
    int wa = lo1 >> 5, wb = lo2 >> 5;
    const int ea = hi1 < lo1 ? wa - 1 : hi1 >> 5, eb = hi2 < lo2 ? wb - 1 : hi2 >> 5;
    while (wa <= ea || wb <= eb) {
        u32 ca, cb, ha, hb;
        team_cov2(rv, t, wa, wb, sub, tmask, &ca, &ha, &cb, &hb);
        int part = 0;
        if (sub != 0) {
            if (wa <= ea) part += __popc(ca & word_range_mask(wa - 1 + sub, lo1, hi1));
            if (wb <= eb) part += __popc(cb & word_range_mask(wb - 1 + sub, lo2, hi2));
        }
        total += team_sum(part, tmask);
        wa += NTL_TEAM - 1; wb += NTL_TEAM - 1;
    }
    return total;
}

/* covered bases of track t inside [a, b] (get_sub_density's numerator, NanoTel.R:467): whole blocks come from K2's
 * counts, the partial blocks at the two ends are recomputed from the read (both in one pass). */
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
        return local_count2(rv, t, lo, hi, 1, 0, sub, tmask);
    }
    int bf = blo, bl = bhi;
    int l1 = 1, h1 = 0, l2 = 1, h2 = 0;
    if (lo != blo_s) { l1 = lo; h1 = (blo + 1) * SG; bf = blo + 1; }
    if (hi != bhi_e) { l2 = bhi * SG + 1; h2 = hi; bl = bhi - 1; }
    int part = 0;                                       /* whole blocks: every lane of the team sums an eighth of them */
    for (int j = bf + sub; j <= bl; j += NTL_TEAM) part += (int)w.cnt[j];
    int total = team_sum(part, tmask);
    if (h1 >= l1 || h2 >= l2) total += local_count2(rv, t, l1, h1, l2, h2, sub, tmask);
    return total;
}

__device__ __forceinline__ double density_of(const ReadView &rv, const WinTab &w, int t, int a, int b, int sub, u32 tmask)
{
    const int cv = covered_in(rv, w, t, a, b, sub, tmask);
    return cv == 0 ? 0.0 : k3_div((double)cv, (double)(b - a + 1));      /* 0 / width is +0.0 exactly */
}

/* get_accurate_start (NanoTel.R:1726-1764) and get_accurate_end (:1692-1721) together: the two are independent, so
 * their coverage blocks are computed in one pass (team_cov2).  `ranges` are the raw exact hits of the single fixed
 * pattern on track A (NanoTel.R:349-354) and the reduced runs of the coverage otherwise (:341-345). */
__device__ __noinline__ void get_accurate_both(const ReadView &rv, int t, int telo_start, int telo_end, int sub, u32 tmask,
                                               int *p_start, int *p_end)
{
    if (telo_start == -1 && telo_end == -1) { *p_start = -1; *p_end = -1; return; }
    const int s = telo_start;
    /* range starts inside [s - 36, s + 99]: the words holding [s - 37, s + 99] (at most 6) are lanes 1 .. 7;
     * range ends inside [e - 99, e + 50]: the words holding [e - 99, e + 51] (at most 6) are lanes 1 .. 7 */
    const int w0s = (s - 37) >> 5, w0e = (telo_end - 99) >> 5;
    u32 cov, hs, cove, hse;
    team_cov2(rv, t, w0s, w0e, sub, tmask, &cov, &hs, &cove, &hse);
    /* ---- start */
    if (telo_start != -1) {
        u32 st;
        if (t == 0 && c_prm.raw_hits_A) st = hs;
        else {
            const u32 pv = __shfl_up_sync(tmask, cov, 1, NTL_TEAM);      /* lane 1 looks behind into lane 0's (incomplete) word:
                                                                            only bit 32 w0 could be wrong, and it is < s - 36 */
            st = cov & ~((cov << 1) | (pv >> 31));
        }
        const int c50 = team_bits_popc(cov, w0s, sub, tmask, s, s + 49);
        const double first_50 = k3_div((double)c50, 50.0);               /* IRanges(start, width = 50) :1732 */
        if (first_50 < 0.3) {
            const int a = team_bits_min(st, w0s, sub, tmask, s + 48, s + 99);
            if (a != NTL_NONE) telo_start = a;
            const int b = team_bits_min(st, w0s, sub, tmask, s + 33, s + 48);
            if (b != NTL_NONE) telo_start = b;
        } else {
            const int a = team_bits_min(st, w0s, sub, tmask, s, s + 99);
            if (a != NTL_NONE) telo_start = a;
            if (first_50 >= 0.72) {
                const int b = team_bits_min(st, w0s, sub, tmask, s - 36, s - 1);
                if (b != NTL_NONE) telo_start = b;
            }
        }
    }
    /* ---- end */
    int e_index = telo_end;
    if (telo_end != -1) {
        u32 en;
        if (t == 0 && c_prm.raw_hits_A) {
            const u32 hp = __shfl_up_sync(tmask, hse, 1, NTL_TEAM);
            en = __funnelshift_l(hp, hse, c_prm.main_pat[0].m - 1);      /* end = start + m - 1 */
        } else {
            u32 nx = __shfl_down_sync(tmask, cove, 1, NTL_TEAM);
            if (sub == NTL_TEAM - 1) nx = 0u;                            /* lane 7's word is beyond e + 51: only a look-ahead */
            en = cove & ~((cove >> 1) | (nx << 31));
        }
        const int m1 = team_bits_max(en, w0e, sub, tmask, telo_end - 99, telo_end);
        if (m1 != NTL_NONE) e_index = m1;
        const int m2 = team_bits_max(en, w0e, sub, tmask, telo_end + 1, telo_end + 50);
        if (m2 != NTL_NONE) e_index = m2;
    }
    *p_start = telo_start; *p_end = e_index;
}

/* One 18-bp window of search_left/right_patterns (multi_pattern_step_*, NanoTel.R:496-575, :614, :676):
 * matchPattern on subseq(read, a, b) with the default fixed = TRUE, the window's own out-of-bounds rule, hits not
 * trimmed.
 * The window and every alignment that can hit it (starts a - k .. b - m + 1 + k) fit one 32-bit word whose bit i is
 * position a - 1 + i: the word is cut out of two position words (the same for all lanes), letters outside [a, b] carry
 * zero masks (mismatches), and all alignment starts of a pattern are decided together by a two-plane mismatch counter
 * (Shift-And).  No shuffles: every lane computes the same value.
 * The right-hand and the left-hand search of a track are independent, so one call serves a window of each (R, L):
 * their loads are in flight together and the two match chains interleave.  on = false: that side is finished. */
__device__ __forceinline__ void window_planes(const ReadView &rv, bool on, int a, int b, u32 (&pl)[4])
{
    pl[0] = pl[1] = pl[2] = pl[3] = 0u;
    if (!on || b < a) return;
    const int rb = a - 2;                               /* index of bit 0 (position a - 1) in the packed stream (p = bit p - 1) */
    const int x = rb >> 5, sh = rb & 31;                /* rb = -1 (a = 1): x = -1, sh = 31 -> the stream shifted up by one */
    const int NP = rv.fmt ? 4 : 2;
    u32 w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int p = 0; p < 4; p++)
        if (p < NP) w[p] = __funnelshift_r(rv_raw(rv, p, x), rv_raw(rv, p, x + 1), sh);
    const u32 vm = ((2u << (b - a)) - 1u) << 1;         /* bits 1 .. b - a + 1: the window itself */
    if (rv.fmt == 0) {
        pl[0] = ~w[1] & ~w[0] & vm; pl[1] = ~w[1] & w[0] & vm; pl[2] = w[1] & w[0] & vm; pl[3] = w[1] & ~w[0] & vm;
    } else {
#pragma unroll
        for (int p = 0; p < 4; p++) pl[p] = w[p] & vm;
    }
}

struct WinHit { bool any; int mn, mx; };
__device__ __forceinline__ void hit_update(WinHit &h, u32 bits, int q0, int m)
{
    if (!bits) return;
    const int lo = q0 + __ffs((int)bits) - 1;
    const int hi = q0 + 31 - __clz((int)bits) + m - 1;
    if (!h.any || lo < h.mn) h.mn = lo;
    if (!h.any || hi > h.mx) h.mx = hi;
    h.any = true;
}

__device__ __noinline__ void step_window2(const ReadView &rv, bool onR, int aR, int bR, bool onL, int aL, int bL, int k,
                                          bool use_tvr, WinHit *hR, WinHit *hL)
{
    WinHit R, Lh;
    R.any = false; R.mn = 0; R.mx = 0; Lh.any = false; Lh.mn = 0; Lh.mx = 0;
    onR = onR && bR >= aR; onL = onL && bL >= aL;
    u32 pr[4], pq[4];
    window_planes(rv, onR, aR, bR, pr);
    window_planes(rv, onL, aL, bL, pq);
    const u32 vmR = onR ? ((2u << (bR - aR)) - 1u) << 1 : 0u, vmL = onL ? ((2u << (bL - aL)) - 1u) << 1 : 0u;
    for (int pass = 0; pass < 2; pass++) {
        const int np = pass == 0 ? c_prm.n_main : (use_tvr ? c_prm.n_tvr : 0);
        const int kk = pass == 0 ? k : 0;
        for (int p = 0; p < np; p++) {
            const ntl_dev_pat &pt = pass == 0 ? c_prm.main_pat[p] : c_prm.tvr_pat[p];
            const int m = pt.m;
            /* alignment starts a - kk .. b - m + 1 + kk  <->  bits 1 - kk .. b - a + 2 - m + kk */
            const int hbR = bR - aR + 2 - m + kk, hbL = bL - aL + 2 - m + kk;
            const bool doR = onR && hbR >= 1 - kk, doL = onL && hbL >= 1 - kk;
            if (!doR && !doL) continue;
            u32 o1 = 0u, t1 = 0u, o2 = 0u, t2 = 0u;
            for (int j = 0; j < m; j++) {
                /* fixed = TRUE: the read's code must EQUAL the pattern letter's code */
                const u32 mA = pt.mux4[j][0], mC = pt.mux4[j][1], mG = pt.mux4[j][2], mT = pt.mux4[j][3];
                const u32 e1 = ~((pr[0] ^ mA) | (pr[1] ^ mC) | (pr[2] ^ mG) | (pr[3] ^ mT)) & vmR;
                const u32 e2 = ~((pq[0] ^ mA) | (pq[1] ^ mC) | (pq[2] ^ mG) | (pq[3] ^ mT)) & vmL;
                const u32 x1 = ~(e1 >> j), x2 = ~(e2 >> j);
                t1 |= o1 & x1; o1 ^= x1;
                t2 |= o2 & x2; o2 ^= x2;
            }
            if (doR) {
                const u32 sm = ((2u << hbR) - 1u) & ~((1u << (1 - kk)) - 1u);
                hit_update(R, (kk ? ~t1 : ~(o1 | t1)) & sm, aR - 1, m);
            }
            if (doL) {
                const u32 sm = ((2u << hbL) - 1u) & ~((1u << (1 - kk)) - 1u);
                hit_update(Lh, (kk ? ~t2 : ~(o2 | t2)) & sm, aL - 1, m);
            }
        }
    }
    *hR = R; *hL = Lh;
}

/* search_right_patterns (NanoTel.R:635-697) from end_index and search_left_patterns (:576-633) from start_index,
 * side by side.  doR / doL = false: that search is not wanted (the caller keeps its own value). */
__device__ __noinline__ void search_both(const ReadView &rv, bool doR, int end_index, bool doL, int start_index, int k,
                                         bool use_tvr, int *p_end, int *p_start)
{
    int subseq_end = end_index + 18 < rv.L ? end_index + 18 : rv.L;
    int new_end = end_index;
    int subseq_start = start_index - 18 > 1 ? start_index - 18 : 1;
    int new_start = start_index;
    bool onR = doR, onL = doL;
    for (int i = 0; i < 4 && (onR || onL); i++) {
        const int curr_start = subseq_end - 17 > 1 ? subseq_end - 17 : 1;              /* right: [curr_start, subseq_end] */
        const int curr_end = subseq_start + 17 < rv.L ? subseq_start + 17 : rv.L;      /* left:  [subseq_start, curr_end] */
        WinHit hR, hL;
        step_window2(rv, onR, curr_start, subseq_end, onL, subseq_start, curr_end, k, use_tvr, &hR, &hL);
        if (onR) {
            if (!hR.any) onR = false;
            else {
                new_end = hR.mx;
                const int nn = subseq_end + 11 < rv.L ? subseq_end + 11 : rv.L;
                if (nn == subseq_end) onR = false;
                subseq_end = nn;
            }
        }
        if (onL) {
            if (!hL.any) onL = false;
            else {
                new_start = hL.mn;
                const int nn = subseq_start - 9 > 1 ? subseq_start - 9 : 1;
                if (nn == subseq_start) onL = false;
                subseq_start = nn;
            }
        }
    }
    *p_end = new_end; *p_start = new_start;
}

/* =============================================================================================================
 * K3a: triage, one THREAD per read.  A read whose window tables hold no telomeric window on any track takes the
 * same path through find_telo_position_wraper every time: find_telo_position = (-1,-1) (:1026-1028), the
 * get_accurate_* calls return -1, width 1 < 100 sends it to find_left/right_telo, which return (-1,-1) as soon as a
 * window lies beyond max_diff = 200 of the chosen edge (:861, :921), search_right_patterns looks at s[1..18]
 * (:1141 with end_index 0) and, finding nothing, leaves (-1, 0): width 2 < 30, not telomeric (:1847).  The thread
 * verifies each of those conditions (integer window classes, one bit-parallel match over the first word) and writes
 * the record; every other read goes on the candidate list of the locate kernel.
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

/* One THREAD per read, reads in input order (coalesced tables and records).  The thread walks the class bits of the
 * regular windows' blocks on the read's LAST track (see below): 32 blocks per word, four independent loads per step;
 * a 20 kb read with 100-base windows is seven words. */
__global__ void __launch_bounds__(128) ntl_triage_kernel(const ntl_read_args a)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    bool cand = false;
    if (r < a.n_reads) {
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
                /* any window with  !(count / width < min_density)  on any track?
                 * The coverage of the tracks is nested (exact hits c <=1-mismatch hits c those + TVR hits), so a
                 * window's covered count can only grow from track to track: some track has a telomeric window iff
                 * the LAST track has one.  The regular windows are decided by the class bits of that track's blocks
                 * (for Q > 1 a set bit only says "maybe": the read goes to the locate kernel, which does not mind);
                 * the last window (own width and threshold) by its count. */
                const int thr_last = (int)a.thr[L - (n_win - 1) * S];
                const int bps = c_prm.cls_bps;
                const u32 *cl = reinterpret_cast<const u32 *>(a.cls[T - 1]);
                const int nrb = (n_win - 1) * Q;                 /* blocks of the regular windows */
                int c_last = 0;                                  /* requested before the bit walk, used after it */
                {
                    const uint16_t *cm = a.cnt[T - 1] + wo;
                    for (int b = nrb; b < nb; b++) c_last += (int)cm[b];
                }
                u32 acc = 0u;
                if (nrb > 0) {
                    const long long b0 = bps == 8 ? wo : wo / bps * 8;
                    const long long a_lo = b0, a_hi = bps == 8 ? b0 + nrb - 1 : b0 + (long long)((nrb - 1) / bps) * 8 + (nrb - 1) % bps;
                    const long long w_lo = a_lo >> 5, w_hi = a_hi >> 5;
                    /* first and last word under their masks, the words between them whole */
                    acc = cls_word(cl, w_lo, a_lo, a_hi) | cls_word(cl, w_hi, a_lo, a_hi);
                    long long wi = w_lo + 1;
                    for (; wi + 4 <= w_hi; wi += 4)
                        acc |= (__ldg(cl + wi) | __ldg(cl + wi + 1)) | (__ldg(cl + wi + 2) | __ldg(cl + wi + 3));
                    for (; wi < w_hi; wi++) acc |= __ldg(cl + wi);
                }
                simple = acc == 0u && c_last < thr_last;
            }
            if (simple) simple = !triage_first_window_hit(lo0 << 1, hi0 << 1, T);
            if (!simple) cand = true;
        }
        if (!cand) {
            alignas(16) ntl_read_result o;       /* stored as four 16-byte pieces */
            o.status = status;
            o.n_win = n_win > 0 ? n_win : 0;
            for (int t = 0; t < 3; t++) {
                const bool live = status == 0 && t < T;
                o.track[t].start = live ? -1 : 0; o.track[t].end = 0; o.track[t].density = 0.0;
            }
            o.win_offset = wo;
            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<ntl_read_result *>(a.results) + r);
#pragma unroll
            for (int q = 0; q < 4; q++) dst[q] = reinterpret_cast<const uint4 *>(&o)[q];
            if (a.stages != nullptr && status == 0) {
                ntl_stage s;
                s.coarse_start = s.coarse_end = s.acc_start = s.acc_end = s.edge_start = s.edge_end = -1;
                s.acc_density = 0.0;
                for (int t = 0; t < T; t++) reinterpret_cast<ntl_stage *>(a.stages)[(size_t)r * 3 + t] = s;
            }
        }
    }
    /* warp-aggregated append to the candidate list */
    const u32 cm = __ballot_sync(NTL_FULL, cand);
    if (cm) {
        int base = 0;
        const int leader = __ffs((int)cm) - 1;
        if (lane == leader) base = (int)atomicAdd(&a.counters[0], (u32)__popc(cm));
        base = __shfl_sync(NTL_FULL, base, leader);
        if (cand) {
            const int pos = base + __popc(cm & ((1u << lane) - 1u));
            a.cand[pos] = r;
            reinterpret_cast<int4 *>(a.cand_state)[pos] = make_int4(0, 0, 0, 0);
        }
    }
}

/* =============================================================================================================
 * K3b: full locator, a team of eight lanes per (candidate read, track); the team that completes a read's last track
 * writes the record head.
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
    w.cls = nullptr; w.bit0 = 0; w.bps = 8; w.dens = nullptr;
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

    K3CLK_BEGIN;
    /* ---- phase 0: tables, coarse interval (window scans: data dependent) */
    if (valid) {
        stg = a.stages ? reinterpret_cast<ntl_stage *>(a.stages) + (size_t)r * 3 : nullptr;
        rv_init(rv, a, r);
        n_win = ntl_nwin(rv.L, S);
        status = rv.fmt ? NTL_READ_IUPAC : 0;
        if (n_win <= 0) status |= NTL_READ_NO_WINDOWS;
        w.cnt = a.cnt[t] + a.cnt_off[r]; w.n = n_win > 0 ? n_win : 0; w.S = S; w.L = rv.L;
        w.Q = c_prm.Q; w.SG = c_prm.SG; w.nb = (rv.L + c_prm.SG - 1) / c_prm.SG;
        w.thr_reg = c_prm.thr_reg; w.dens = a.dens;
        w.cls = reinterpret_cast<const u32 *>(a.cls[t]); w.bps = c_prm.cls_bps; w.bit0 = c_prm.cls_bps == 8 ? a.cnt_off[r] : a.cnt_off[r] / c_prm.cls_bps * 8;
        w.thr_last = w.n > 0 ? (int)a.thr[wt_end(w, w.n - 1) - wt_start(w, w.n - 1) + 1] : 0;
    }
    K3CLK(0);
    if (valid) {
        find_telo_position(w, 3.0, 2.0, &ts, &te, sub, tmask, threadIdx.x & 31);                                     /* :1084-1086 */
    }
    K3CLK(1);
    double telo_density = 0.0;
    if (valid) {
        /* the coarse interval is made of whole windows: its density needs no coverage recomputation */
        telo_density = density_of(rv, w, t, ts, te, sub, tmask);                       /* :1099 */
    }
    K3CLK(2);
    if (valid) {
        const int num_rows = (te - ts + 1) / S;                                        /* :1103 */
        if (telo_density < 0.85 && num_rows > 5) {                                     /* :1104-1110 */
            const double min_rows = num_rows <= 7 ? (double)(num_rows - 2) : 7.0;
            const double min_score = 0.6 * min_rows;
            find_telo_position(w, min_rows, min_score, &ts, &te, sub, tmask, threadIdx.x & 31);
        }
        cs = ts; ce = te;
    }
    __syncwarp();
    K3CLK(3);
    /* ---- phase 1: accurate start and end (one coverage block each, computed side by side) */
    int start_acc = -1, end_acc = -1;
    if (valid) get_accurate_both(rv, t, ts, te, sub, tmask, &start_acc, &end_acc);     /* :1119-1120 */
    __syncwarp();
    K3CLK(4);
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
    K3CLK(6);
    /* ---- phase 2: the 18-bp re-match searches, right and left side by side */
    if (valid && !err) {                                                               /* :1140-1149 */
        e2 = te; s2 = ts;
        const bool doR = te < rv.L, doL = ts > 1;
        if (doR || doL) {
            int e3, s3;
            search_both(rv, doR, te + 1, doL, ts - 1, k, use_tvr, &e3, &s3);
            if (doR) e2 = e3;
            if (doL) s2 = s3;
        }
    }
    __syncwarp();
    K3CLK(7);
    /* ---- phase 3: final density (:1840-1844) */
    if (valid && !err) {
        if (e2 < s2 - 1) err = true;                                                   /* IRanges() would stop */
        else {
            out.start = s2; out.end = e2;
            out.density = density_of(rv, w, t, s2, e2, sub, tmask);
            width = e2 - s2 + 1;
        }
    }
    __syncwarp();
    K3CLK(9);
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

/* 16 CTAs of two warps per SM (64 registers, a few spilled values): with three tracks or a filtered batch there are
 * more item groups than the 22 warps per SM of an unbounded build (92 registers) can hold in one wave -- measured:
 * cfg3 0.143 -> 0.119 ms, cfg4 0.192 -> 0.155, cfg2 (one wave either way) unchanged */
#ifndef NTL_LOCATE_MINB
#define NTL_LOCATE_MINB 16
#endif
__global__ void __launch_bounds__(64, NTL_LOCATE_MINB) ntl_locate_kernel(const ntl_read_args a)
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
        K3CLK_BEGIN;
        locate_items(a, valid, valid ? a.cand[c] : 0, valid ? i - c * T : 0, a.cand_state + 4 * (size_t)c, sub, tmask);
        K3CLK(15);
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

/* Class bits from the block counts, dense layout (cls_bps = 8: bit g = block g): one thread per byte.  Runs after a
 * scan kernel that does not write the bits itself (the generic kernel; spans of more than 8 blocks). */
__global__ void __launch_bounds__(256) ntl_cls_kernel(const ntl_read_args a, long long n_bytes, int T, u32 blk_thr)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bytes) return;
    for (int t = 0; t < T; t++) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(a.cnt[t]) + i);
        const u32 x[4] = {v.x, v.y, v.z, v.w};
        u32 b = 0u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if ((x[q] & 0xffffu) >= blk_thr) b |= 1u << (2 * q);
            if ((x[q] >> 16) >= blk_thr) b |= 2u << (2 * q);
        }
        a.cls[t][i] = (uint8_t)b;
    }
}

extern "C" cudaError_t ntl_k_cls(const ntl_read_args *a, long long n_bytes, int T, uint32_t blk_thr, cudaStream_t st)
{
    if (n_bytes <= 0) return cudaSuccess;
    ntl_cls_kernel<<<(unsigned)((n_bytes + 255) / 256), 256, 0, st>>>(*a, n_bytes, T, blk_thr);
    return cudaGetLastError();
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
    ntl_triage_kernel<<<(a->n_reads + 127) / 128, 128, 0, st>>>(*a);             /* one thread per read */
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
