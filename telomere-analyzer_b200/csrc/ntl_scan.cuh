/*
 * ntl_scan.cuh -- K2, the streaming kernel of the hot path (sm_100a).
 *
 * For every read it produces, in ONE pass over the packed bases, what NanoTel.R computes with
 *   get_density_iranges (NanoTel.R:308-397)  -> pattern hits (exact / <=1 mismatch / +TVR) and their union,
 *   split_telo          (NanoTel.R:199-227)  -> the window table,
 *   get_sub_density     (NanoTel.R:449-468)  -> covered bases per window (numerator of the density),
 * for tracks A (exact), B (<= 1 mismatch) and C (B + exact TVR patterns).
 *
 * Mapping to the machine (B200: 148 SMs, 32-wide warps, LDG.256, 64-lane/clk integer pipe):
 *   - one warp streams one read, longest reads first (dynamic work counter), 4096 positions per step:
 *     lane l owns 128 consecutive positions = one 32-byte quad {lo[4], hi[4]}, fetched with one LDG.256
 *     (1 KiB per warp per step, fully coalesced); two register buffers swap roles every step, so the quad of step
 *     c + 2 is requested as soon as the hits of step c exist and nothing waits for a load it has just issued;
 *   - matching is Shift-And in its position-parallel form: one 32-bit word holds 32 text positions, a pattern
 *     letter is one boolean function of the two bit-planes (LOP3), its j-th letter is aligned with a funnel
 *     shift, and letters are combined three at a time into an (all equal, at most one differs) pair
 *     (AND3 / MAJ3, one LOP3 each); pairs combine as  EX' = EX & ex,  LE' = (EX & le) | (LE & ex):
 *         exact hit  = EX   (track A)          <=1-mismatch hit = LE   (track B)
 *     positions outside [1, L] carry an all-zero "equal" mask, which reproduces Biostrings' rule that
 *     out-of-bounds letters are mismatches (hits may start at 0 or end at L+1 with one mismatch, App. B.3);
 *   - coverage (IRanges::union of the hit intervals, then trim) = hit-start mask dilated by the pattern length
 *     with ternary / doubling funnel-shift steps -- or, for a group of patterns whose exact hits cannot overlap
 *     (NTL_J_MAIN/TVR_UNBORDERED, decided on the host), the single 160-bit subtraction (H << m) - H; the <= 17 bits
 *     that spill into the next lane travel by one shuffle;
 *   - window counts: per-lane popcounts, one packed warp scan per step (its adds are IMADs, the integer ALU pipe is
 *     the kernel's limiter), and every window end that falls into a
 *     lane's 128 positions is written by that lane as the running prefix "covered bases in [1, window end]"
 *     (uint16, mod 2^16).  Each prefix is written exactly once: no atomics, no zero-fill.  Consumers take
 *     differences.
 *
 * Compiled twice from this one source: by nvcc with the patterns in __constant__ memory (any pattern set), and by
 * NVRTC at ntl_create() with the pattern set and subseq_length baked in (NTL_JIT), which turns every pattern
 * letter into a single LOP3, shares equal letters, and fully unrolls the letter and window-end loops.
 */
#ifndef NTL_SCAN_CUH
#define NTL_SCAN_CUH

#include "ntl_dev.h"

typedef unsigned int u32;

#ifndef NTL_JIT
/* c_prm (__constant__ ntl_dev_params) is defined by the including .cu file before this header */
#define PRM_S        (c_prm.S)
#define PRM_NTRACKS  (c_prm.n_tracks)
#define PRM_THR_REG  (c_prm.thr_reg)
#define PRM_MAIN_LEN(p) (c_prm.main_pat[p].m)
#define PRM_TVR_LEN(p)  (c_prm.tvr_pat[p].m)
#define PRM_NMAIN_GROUPS (c_prm.n_main_groups)
#define PRM_NTVR_GROUPS  (c_prm.n_tvr_groups)
#define PRM_MAIN_GBEGIN(g) (c_prm.main_group_begin[g])
#define PRM_TVR_GBEGIN(g)  (c_prm.tvr_group_begin[g])
#define NTL_UNROLL_PAT _Pragma("unroll 1")
#define NTL_UNROLL_LET _Pragma("unroll 1")
#define NTL_UNROLL_WEND _Pragma("unroll 1")
#else
#define PRM_S        NTL_J_S
#define PRM_NTRACKS  NTL_J_NTRACKS
#define PRM_THR_REG  NTL_J_THR_REG
#define PRM_MAIN_LEN(p) (NTL_J_MAIN_LEN[p])
#define PRM_TVR_LEN(p)  (NTL_J_TVR_LEN[p])
#define PRM_NMAIN_GROUPS NTL_J_NMAIN_GROUPS
#define PRM_NTVR_GROUPS  NTL_J_NTVR_GROUPS
#define PRM_MAIN_GBEGIN(g) (NTL_J_MAIN_GBEGIN[g])
#define PRM_TVR_GBEGIN(g)  (NTL_J_TVR_GBEGIN[g])
#define NTL_UNROLL_PAT _Pragma("unroll")
#define NTL_UNROLL_LET _Pragma("unroll")
#define NTL_UNROLL_WEND _Pragma("unroll")
#endif

#define NTL_FULL 0xffffffffu

__device__ __forceinline__ int ntl_nwin(int L, int S)
{
    /* split_telo, NanoTel.R:216-224: drop the last window if  L - last_start < S / 2  (real division) */
    int n = (L - 1) / S + 1;
    int last = 1 + (n - 1) * S;
    if (2 * (L - last) < S) n -= 1;
    return n;
}

__device__ __forceinline__ void ntl_ldg256(const u32 *p, u32 (&v)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}

/* bits b of the word starting at bit index wpos with 1 <= wpos + b <= L: the low  clamp(L - wpos + 1, 0, 32)  bits
 * (one clamped funnel shift), minus bit 0 of the read's very first word (the pad position) */
__device__ __forceinline__ u32 ntl_valid_word(int wpos, int L)
{
    int nb = L - wpos + 1;
    if (nb < 0) nb = 0;
    u32 m = __funnelshift_lc(NTL_FULL, 0u, nb);
    if (wpos == 0) m &= ~1u;
    return m;
}

/* "letter accepts this base" over 32 positions of a 2-bit read; t[c] = all-ones iff code c (= hi*2+lo) accepted */
__device__ __forceinline__ u32 ntl_eq2(u32 hi, u32 lo, u32 v, u32 t0, u32 t1, u32 t2, u32 t3)
{
    u32 a = (lo & t1) | (~lo & t0);
    u32 b = (lo & t3) | (~lo & t2);
    return ((hi & b) | (~hi & a)) & v;
}

/* n <= 3 aligned letter masks -> (all equal, at most one differs) */
__device__ __forceinline__ void ntl_group3(int n, u32 a, u32 b, u32 c, u32 &ex, u32 &le)
{
    if (n == 1) { ex = a; le = NTL_FULL; }
    else if (n == 2) { ex = a & b; le = a | b; }
    else { ex = a & b & c; le = (a & b) | (a & c) | (b & c); }
}

/* Funnel shifts.  The scan kernel is bound by the integer ALU pipe (LOP3 / SHF), while the FMA pipe idles: with
 * NTL_FSHIFT_IMAD the shifts are computed there instead, as  (lo >> j) | (hi << (32 - j))  =  mulhi(lo, K) + hi * K
 * with K = 2^(32-j) read from constant memory (so that the compiler cannot turn the multiplications back into shifts). */
#ifdef NTL_FSHIFT_IMAD
__constant__ u32 ntl_pow2[33] = {1u, 2u, 4u, 8u, 16u, 32u, 64u, 128u, 256u, 512u, 1024u, 2048u, 4096u, 8192u, 16384u,
                                 32768u, 65536u, 131072u, 262144u, 524288u, 1048576u, 2097152u, 4194304u, 8388608u,
                                 16777216u, 33554432u, 67108864u, 134217728u, 268435456u, 536870912u, 1073741824u,
                                 2147483648u, 0u};
__device__ __forceinline__ u32 ntl_fsr(u32 lo, u32 hi, int j)          /* low word of (hi:lo) >> j, 0 <= j < 32 */
{
#ifdef NTL_FSHIFT_IMAD_ALIGN
    if (j == 0) return lo;
    const u32 K = ntl_pow2[32 - j];
    return __umulhi(lo, K) + hi * K;
#else
    return __funnelshift_r(lo, hi, j);
#endif
}
__device__ __forceinline__ u32 ntl_fsl(u32 lo, u32 hi, int j)          /* high word of (hi:lo) << j, 0 <= j < 32 */
{
#ifdef NTL_FSHIFT_IMAD_DILATE
    if (j == 0) return hi;
    const u32 K = ntl_pow2[j];
    return __umulhi(lo, K) + hi * K;
#else
    return __funnelshift_l(lo, hi, j);
#endif
}
#else
__device__ __forceinline__ u32 ntl_fsr(u32 lo, u32 hi, int j) { return __funnelshift_r(lo, hi, j); }
__device__ __forceinline__ u32 ntl_fsl(u32 lo, u32 hi, int j) { return __funnelshift_l(lo, hi, j); }
#endif

/* Dilate hit starts forward by m positions, in place, over the lane's 4 words; word 4 receives what spills past them:
 * its bit k is set iff a hit starts in the top m - 1 - k bits of word 3, i.e. it is those m - 1 bits smeared towards
 * bit 0 (count-leading-zeros on the XU pipe plus one clamped shift instead of a fifth dilation). */
__device__ __forceinline__ void ntl_dilate5(u32 (&d)[5], int m)
{
    u32 lz;                         /* leading zeros of those bits; 0xffffffff if none is set, which clamps to 32 */
    asm("bfind.shiftamt.u32 %0, %1;" : "=r"(lz) : "r"(__funnelshift_rc(d[3], 0u, 33 - m)));
    d[4] = __funnelshift_rc(NTL_FULL, 0u, lz);
    int w = 1;
    while (3 * w <= m) {            /* width w -> 3 w: two shifts and one three-input OR (a single LOP3) per word */
#pragma unroll
        for (int i = 3; i >= 1; i--)
            d[i] = d[i] | ntl_fsl(d[i - 1], d[i], w) | ntl_fsl(d[i - 1], d[i], 2 * w);
        d[0] = d[0] | (d[0] << w) | (d[0] << (2 * w));
        w *= 3;
    }
    while (2 * w <= m) {
#pragma unroll
        for (int i = 3; i >= 1; i--) d[i] |= ntl_fsl(d[i - 1], d[i], w);
        d[0] |= d[0] << w;
        w *= 2;
    }
    int s = m - w;
    if (s > 0) {
#pragma unroll
        for (int i = 3; i >= 1; i--) d[i] |= ntl_fsl(d[i - 1], d[i], s);
        d[0] |= d[0] << s;
    }
}

/* Coverage of hit starts whose intervals cannot overlap (see NTL_J_MAIN_UNBORDERED): every hit bit 2^s becomes the
 * run 2^(s+m) - 2^s, so the whole array is (H << m) - H: four shifts and one 160-bit borrow chain. */
__device__ __forceinline__ void ntl_cover_sub(u32 (&d)[5], int m)
{
    u32 s[5];
    s[0] = d[0] << m;
#pragma unroll
    for (int i = 1; i < 4; i++) s[i] = __funnelshift_l(d[i - 1], d[i], m);
    s[4] = __funnelshift_l(d[3], 0u, m);
    asm("sub.cc.u32 %0, %5, %10;\n\t"
        "subc.cc.u32 %1, %6, %11;\n\t"
        "subc.cc.u32 %2, %7, %12;\n\t"
        "subc.cc.u32 %3, %8, %13;\n\t"
        "subc.u32 %4, %9, 0;"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4])
        : "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]));
}

/* aligned "equal" mask of letter j of one pattern for the lane's 4 words: x[i] bit b <=> position (32 i + b + j)
 * carries a base the letter accepts.  pl: planes [NPL][5], v: validity [5]. */
template <int NPL, bool TVR>
__device__ __forceinline__ void ntl_letter(int p, int j, const u32 (&pl)[NPL][5], const u32 (&v)[5], u32 (&x)[4])
{
    u32 e[5];
#ifndef NTL_JIT
    const ntl_dev_pat &pt = TVR ? c_prm.tvr_pat[p] : c_prm.main_pat[p];
    if constexpr (NPL == 2) {
        const u32 t0 = pt.mux2[j][0], t1 = pt.mux2[j][1], t2 = pt.mux2[j][2], t3 = pt.mux2[j][3];
#pragma unroll
        for (int i = 0; i < 5; i++) e[i] = ntl_eq2(pl[1][i], pl[0][i], v[i], t0, t1, t2, t3);
    } else {
        const u32 nb = pt.nib[j];
        const u32 mA = (nb & 1u) ? NTL_FULL : 0u, mC = (nb & 2u) ? NTL_FULL : 0u;
        const u32 mG = (nb & 4u) ? NTL_FULL : 0u, mT = (nb & 8u) ? NTL_FULL : 0u;
        const bool fx = pt.fixed != 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            u32 any = (pl[0][i] & mA) | (pl[1][i] & mC) | (pl[2][i] & mG) | (pl[3][i] & mT);
            u32 dif = (pl[0][i] ^ mA) | (pl[1][i] ^ mC) | (pl[2][i] ^ mG) | (pl[3][i] ^ mT);
            e[i] = (fx ? ~dif : any) & v[i];
        }
    }
#else
#pragma unroll
    for (int i = 0; i < 5; i++) {
        if constexpr (NPL == 2)
            e[i] = TVR ? ntl_jit_eq_tvr(p, j, pl[1][i], pl[0][i], v[i]) : ntl_jit_eq_main(p, j, pl[1][i], pl[0][i], v[i]);
        else
            e[i] = TVR ? ntl_jit_eq4_tvr(p, j, pl[0][i], pl[1][i], pl[2][i], pl[3][i], v[i])
                       : ntl_jit_eq4_main(p, j, pl[0][i], pl[1][i], pl[2][i], pl[3][i], v[i]);
    }
#endif
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = ntl_fsr(e[i], e[i + 1], j);
}

/* hit-start masks of one pattern over the lane's 4 words: EX = exact, LE = at most one mismatch */
template <int NPL, bool TVR>
__device__ __forceinline__ void ntl_pattern_hits(int p, int m, const u32 (&pl)[NPL][5], const u32 (&v)[5],
                                                 u32 (&EX)[4], u32 (&LE)[4])
{
#pragma unroll
    for (int i = 0; i < 4; i++) { EX[i] = NTL_FULL; LE[i] = NTL_FULL; }
    NTL_UNROLL_LET
    for (int j0 = 0; j0 < m; j0 += 3) {
        const int n = m - j0 < 3 ? m - j0 : 3;
        u32 x0[4], x1[4] = {0u, 0u, 0u, 0u}, x2[4] = {0u, 0u, 0u, 0u};
        ntl_letter<NPL, TVR>(p, j0, pl, v, x0);
        if (n > 1) ntl_letter<NPL, TVR>(p, j0 + 1, pl, v, x1);
        if (n > 2) ntl_letter<NPL, TVR>(p, j0 + 2, pl, v, x2);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u32 ex, le;
            ntl_group3(n, x0[i], x1[i], x2[i], ex, le);
            if (!TVR) LE[i] = (EX[i] & le) | (LE[i] & ex);
            EX[i] &= ex;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * NPL = 2: ACGT reads, planes {lo, hi};  NPL = 4: IUPAC reads, planes {A, C, G, T} (Biostrings code bits).
 * ------------------------------------------------------------------------------------------------------------ */
template <int NPL>
__device__ __forceinline__ void ntl_scan_read(const ntl_scan_args &a, int r, int lane, int kq_init, int off_init,
                                              int adv_a, int adv_b)
{
    const int S = PRM_S;
    const int T = PRM_NTRACKS;
    const int L = a.len[r];
    const int n_win = ntl_nwin(L, S);
    if (n_win <= 0) return;
    const u32 *base = a.packed + a.woff[r];
    const long long wo = a.win_off[r];
    const int n_words = (L >> 5) + 1;
    const int n_quads = (n_words + 3) >> 2;
    const int n_chunks = (n_quads + 31) >> 5;
    const int last_chunk = L >> 12;                /* chunk holding position L = end of the last window */
    constexpr int QW = NPL * 4;                    /* words per quad */

    u32 carry[3] = {0u, 0u, 0u};                   /* coverage spill of the previous chunk's lane 31 -> lane 0 */
    u32 run[3] = {0u, 0u, 0u};                     /* covered bases before this chunk (warp-uniform)          */
    int kq = kq_init, off = off_init;              /* first multiple of S at or after this lane's first bit   */

    const u32 *qp = base + (size_t)lane * QW;
    uint16_t *cump[3];                             /* this read's first window in each track's prefix plane   */
#pragma unroll
    for (int t = 0; t < 3; t++) cump[t] = t < T ? a.cum[t] + wo : nullptr;
    uint16_t *wp[3];                               /* &cump[t][kq - 1]: this lane's first regular window end   */
#pragma unroll
    for (int t = 0; t < 3; t++) wp[t] = cump[t] + (kq_init - 1);
    const u32 v0_first = lane == 0 ? ~1u : NTL_FULL;
    u32 ge[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {                  /* opaque 0/1 factors: keeps the warp scan's adds as IMADs  */
        const u32 g = lane >= (1 << k) ? 1u : 0u;
        asm("mov.u32 %0, %1;" : "=r"(ge[k]) : "r"(g));
    }
    /* Two quad buffers that swap roles every step: while step c works on one (chunk c), the other already holds
     * chunk c + 1 (its first word is lane 31's fifth word), and as soon as the letter masks of chunk c are formed its
     * registers receive chunk c + 2.  Every load is issued more than a full step before its first use, and no
     * register is copied.  Quads beyond the read are not loaded: stale bits there are masked by v[]. */
    u32 bufA[QW], bufB[QW];
#pragma unroll
    for (int i = 0; i < QW; i++) { bufA[i] = 0u; bufB[i] = 0u; }
    auto load_quad = [&](u32 (&X)[QW], const u32 *src) {
        if constexpr (NPL == 2) ntl_ldg256(src, *reinterpret_cast<u32(*)[8]>(&X[0]));
        else {
            ntl_ldg256(src, *reinterpret_cast<u32(*)[8]>(&X[0]));
            ntl_ldg256(src + 8, *reinterpret_cast<u32(*)[8]>(&X[8]));
        }
    };
    if (lane < n_quads) load_quad(bufA, qp);
    if (lane + 32 < n_quads) load_quad(bufB, qp + 32 * QW);

    auto step = [&](u32 (&cur)[QW], u32 (&nxt)[QW], const int c) {
        /* ---- planes: 4 own words + the first word of the next lane (next chunk for lane 31) */
        u32 pl[NPL][5];
#pragma unroll
        for (int k = 0; k < NPL; k++) {
#pragma unroll
            for (int i = 0; i < 4; i++) pl[k][i] = cur[k * 4 + i];
            /* one rotation by a lane: lane 0 offers the first word of the NEXT chunk, which is what lane 31 needs */
            pl[k][4] = __shfl_sync(NTL_FULL, lane == 0 ? nxt[k * 4] : cur[k * 4], (lane + 1) & 31);
        }
        const int pos0 = (c * 32 + lane) * NTL_LANE_BITS;      /* bit index = 1-based position */
        u32 v[5];
        if ((c + 1) * NTL_CHUNK_BITS + 32 > L) {               /* only the read's tail has positions beyond L */
#pragma unroll
            for (int i = 0; i < 5; i++) {
                int nb = L - (pos0 + 32 * i) + 1;
                if (nb < 0) nb = 0;
                v[i] = __funnelshift_lc(NTL_FULL, 0u, nb);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 5; i++) v[i] = NTL_FULL;
        }
        if (c == 0) v[0] &= v0_first;                          /* position 0 is the pad bit */

        u32 cov[3][5];
#pragma unroll
        for (int t = 0; t < 3; t++)
#pragma unroll
            for (int i = 0; i < 5; i++) cov[t][i] = 0u;
        /* once the hits of the last pattern exist the planes of this chunk are dead: their registers take chunk c + 2 */
        auto reload = [&]() {
            qp += 32 * QW;
            if ((c + 2) * 32 + lane < n_quads) load_quad(cur, qp + 32 * QW);
        };

        /* ---- main patterns: tracks A and B (get_density_iranges :327-356) */
        NTL_UNROLL_PAT
        for (int g = 0; g < PRM_NMAIN_GROUPS; g++) {
            const int m = PRM_MAIN_LEN(PRM_MAIN_GBEGIN(g));
            u32 hA[5] = {0u, 0u, 0u, 0u, 0u}, hB[5] = {0u, 0u, 0u, 0u, 0u};
            NTL_UNROLL_PAT
            for (int p = PRM_MAIN_GBEGIN(g); p < PRM_MAIN_GBEGIN(g + 1); p++) {
                u32 EX[4], LE[4];
                ntl_pattern_hits<NPL, false>(p, m, pl, v, EX, LE);
#pragma unroll
                for (int i = 0; i < 4; i++) { hA[i] |= EX[i]; hB[i] |= LE[i]; }
            }
            if (T < 3 && g == PRM_NMAIN_GROUPS - 1) reload();
#ifdef NTL_JIT
            if (NPL == 2 && NTL_J_MAIN_UNBORDERED[g]) ntl_cover_sub(hA, m); else      /* IUPAC read letters can make hits overlap */
#endif
            ntl_dilate5(hA, m);
            ntl_dilate5(hB, m);
#pragma unroll
            for (int i = 0; i < 5; i++) { cov[0][i] |= hA[i]; cov[1][i] |= hB[i]; }
        }
        /* ---- TVR patterns, exact: track C = B + TVR (get_density_iranges :360-393) */
        if (T == 3) {
            NTL_UNROLL_PAT
            for (int g = 0; g < PRM_NTVR_GROUPS; g++) {
                const int m = PRM_TVR_LEN(PRM_TVR_GBEGIN(g));
                u32 hC[5] = {0u, 0u, 0u, 0u, 0u};
                NTL_UNROLL_PAT
                for (int p = PRM_TVR_GBEGIN(g); p < PRM_TVR_GBEGIN(g + 1); p++) {
                    u32 EX[4], LE[4];
                    ntl_pattern_hits<NPL, true>(p, m, pl, v, EX, LE);
#pragma unroll
                    for (int i = 0; i < 4; i++) hC[i] |= EX[i];
                }
                if (g == PRM_NTVR_GROUPS - 1) reload();
#ifdef NTL_JIT
                if (NPL == 2 && NTL_J_TVR_UNBORDERED[g]) ntl_cover_sub(hC, m); else
#endif
                ntl_dilate5(hC, m);
#pragma unroll
                for (int i = 0; i < 5; i++) cov[2][i] |= hC[i];
            }
#pragma unroll
            for (int i = 0; i < 5; i++) cov[2][i] |= cov[1][i];
        }

        /* ---- coverage spill from the previous lane, trim to [1, L], popcounts */
        u32 pc[3][4], tot[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            tot[t] = 0u;
#pragma unroll
            for (int i = 0; i < 4; i++) pc[t][i] = 0u;
            if (t < T) {
                u32 sp = __shfl_up_sync(NTL_FULL, cov[t][4], 1);
                if (lane == 0) sp = carry[t];
                carry[t] = __shfl_sync(NTL_FULL, cov[t][4], 31);
                cov[t][0] |= sp;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    cov[t][i] &= v[i];
                    pc[t][i] = (u32)__popc(cov[t][i]);
                    tot[t] += pc[t][i];
                }
            }
        }
        /* ---- packed warp scan: (A | B << 16) and C; each lane total <= 128, chunk total <= 4096 */
        u32 s01 = tot[0] | (tot[1] << 16), s2 = tot[2];
#pragma unroll
        for (int k = 0; k < 5; k++) {                          /* s += (lane >= 2^k) * y: one IMAD, off the ALU pipe */
            u32 y01 = __shfl_up_sync(NTL_FULL, s01, 1 << k);
            s01 = y01 * ge[k] + s01;
            if (T == 3) {
                u32 y2 = __shfl_up_sync(NTL_FULL, s2, 1 << k);
                s2 = y2 * ge[k] + s2;
            }
        }
        u32 ex[3];
        ex[0] = run[0] + (s01 & 0xffffu) - tot[0];
        ex[1] = run[1] + (s01 >> 16) - tot[1];
        ex[2] = run[2] + s2 - tot[2];
        {
            u32 l01 = __shfl_sync(NTL_FULL, s01, 31);
            run[0] += l01 & 0xffffu; run[1] += l01 >> 16;
            if (T == 3) run[2] += __shfl_sync(NTL_FULL, s2, 31);
        }
        /* bytes 1..3 of pre[t] = covered bases of this lane before word 1..3 (each <= 96) */
        u32 pre[3];
#pragma unroll
        for (int t = 0; t < 3; t++) pre[t] = pc[t][0] * 0x01010100u + pc[t][1] * 0x01010000u + pc[t][2] * 0x01000000u;

        /* ---- window ends inside this lane's 128 positions.  Regular ends are the multiples of S with index
         *      kq-1 <= n_win-2; the last window ends at L (index n_win-1; split_telo :218, :223). */
        const int n_ends = (NTL_LANE_BITS - 1) / S + 1;        /* most multiples of S inside 128 positions */
        NTL_UNROLL_WEND
        for (int it = 0; it <= n_ends; it++) {
            int rel, idx;
            bool act;
            if (it < n_ends) {
                rel = off + it * S; idx = kq + it - 1;
                act = rel < NTL_LANE_BITS && idx >= 0 && idx <= n_win - 2;
            } else {
                if (c != last_chunk) break;                    /* warp-uniform */
                rel = L - pos0; idx = n_win - 1;
                act = rel >= 0 && rel < NTL_LANE_BITS;
            }
            {
                const u32 bm = __funnelshift_r(NTL_FULL, 0u, ~rel);      /* bits 0 .. rel & 31 */
                if (it < n_ends && ((it * S) >> 5) >= 3) {
                    /* an active end of this rank can only lie in the lane's last word */
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        if (t < T) {
                            const u32 val = ex[t] + (pre[t] >> 24) + (u32)__popc(cov[t][3] & bm);
                            if (act) wp[t][it] = (uint16_t)val;
                        }
                    }
                } else {
                    /* branch-free: word (rel >> 5) of the lane's four is picked with clamped funnel shifts, the
                     * count before it with one byte permute */
                    const int sh1 = rel & 32, sh2 = rel & 64;  /* shift counts clamp at 32: 64 selects the high word */
                    const u32 bsel = 0x4440u + ((u32)rel >> 5);  /* active ends have 0 <= rel < 128 */
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        if (t < T) {
                            const u32 wv = __funnelshift_rc(__funnelshift_rc(cov[t][0], cov[t][1], sh1),
                                                            __funnelshift_rc(cov[t][2], cov[t][3], sh1), sh2);
                            const u32 val = ex[t] + __byte_perm(pre[t], 0u, bsel) + (u32)__popc(wv & bm);
                            if (act) { if (it < n_ends) wp[t][it] = (uint16_t)val; else cump[t][idx] = (uint16_t)val; }
                        }
                    }
                }
            }
        }
        /* ---- advance this lane's window-end cursor by one chunk (4096 = adv_a * S + adv_b) */
        {
            const bool wrap = off < adv_b;
            const int adv = adv_a + (wrap ? 1 : 0);
            off += wrap ? S - adv_b : -adv_b;
            kq += adv;
#pragma unroll
            for (int t = 0; t < 3; t++) if (t < T) wp[t] += adv;
        }

    };

    for (int c = 0; c < n_chunks; c += 2) {
        step(bufA, bufB, c);
        if (c + 1 >= n_chunks) break;
        step(bufB, bufA, c + 1);
    }
}

template <int NPL>
__device__ __forceinline__ void ntl_scan_body(const ntl_scan_args &a)
{
    const int lane = threadIdx.x & 31;
    const int S = PRM_S;
    const int kq_init = (NTL_LANE_BITS * lane + S - 1) / S;
    const int off_init = kq_init * S - NTL_LANE_BITS * lane;
    const int adv_a = NTL_CHUNK_BITS / S, adv_b = NTL_CHUNK_BITS % S;
    /* the next work item is claimed (atomic + order[] lookup) while the current read is being scanned, so that only
     * the read's own tables and first quads are waited for between two reads */
    auto claim = [&]() -> int {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(a.counter, 1u);
        item = __shfl_sync(NTL_FULL, item, 0);
        return item < a.n_items ? a.order[item] : -1;
    };
    int r = claim();
    while (r >= 0) {
        const int r_next = claim();
        if (a.pass == nullptr || a.pass[r] != 0) ntl_scan_read<NPL>(a, r, lane, kq_init, off_init, adv_a, adv_b);
        r = r_next;
    }
}

#endif /* NTL_SCAN_CUH */
