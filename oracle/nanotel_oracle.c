/*
 * nanotel_oracle.c -- CPU restatement of NanoTel.R's per-read telomere detection (plain C11).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see nanotel_oracle.h).  Every function cites the lines of
 * /root/reference/NanoTel.R it follows.  The restatement is deliberately literal: it builds the same
 * range lists ("IRanges") the R code builds and walks them the way the R code does, so that accidental
 * behaviour (raw hits vs reduced runs, untrimmed out-of-bounds hits in the 18-bp re-match, the
 * "start-1 / end+1 when nothing is found" shifts) is reproduced rather than idealised.
 */
#include "nanotel_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

#define CLS_TELO (-5)   /* classes$CCCTAA, NanoTel.R:749 */
#define CLS_NONE 1
#define CLS_SKIP 0

/* ---------------------------------------------------------------- Biostrings DNA codes (App. B.1) */
static uint8_t g_code[256];
static uint8_t g_comp[256];      /* ASCII complement, IUPAC aware */
static int g_init_done = 0;

static void init_tables(void)
{
    if (g_init_done) return;
    memset(g_code, 0xFF, sizeof g_code);
    const char *letters = "ACGTMRWSYKVHDBN";
    const uint8_t codes[] = {1, 2, 4, 8, 3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15};
    for (int i = 0; letters[i]; i++) {
        g_code[(unsigned char)letters[i]] = codes[i];
        g_code[(unsigned char)(letters[i] + 32)] = codes[i];   /* lower case stored as upper */
    }
    g_code['-'] = 16; g_code['+'] = 32; g_code['.'] = 64;
    for (int c = 0; c < 256; c++) g_comp[c] = (uint8_t)c;
    const char *from = "ACGTMRWSYKVHDBN", *to = "TGCAKYWSRMBDHVN";
    for (int i = 0; from[i]; i++) {
        g_comp[(unsigned char)from[i]] = (uint8_t)to[i];
        g_comp[(unsigned char)(from[i] + 32)] = (uint8_t)to[i]; /* Biostrings stores upper case */
    }
    g_init_done = 1;
}

/* ---------------------------------------------------------------- range lists (IRanges) */
typedef struct { int32_t *start, *end; int32_t n, cap; } rlist;

static void rl_init(rlist *r) { r->start = r->end = NULL; r->n = r->cap = 0; }
static void rl_free(rlist *r) { free(r->start); free(r->end); rl_init(r); }
static void rl_push(rlist *r, int32_t s, int32_t e)
{
    if (r->n == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 256;
        r->start = (int32_t *)realloc(r->start, sizeof(int32_t) * (size_t)r->cap);
        r->end = (int32_t *)realloc(r->end, sizeof(int32_t) * (size_t)r->cap);
    }
    r->start[r->n] = s; r->end[r->n] = e; r->n++;
}

typedef struct { int32_t s, e; } pair_t;
static int cmp_pair(const void *a, const void *b)
{
    const pair_t *x = (const pair_t *)a, *y = (const pair_t *)b;
    if (x->s != y->s) return x->s < y->s ? -1 : 1;
    if (x->e != y->e) return x->e < y->e ? -1 : 1;
    return 0;
}

/* IRanges::reduce (App. B.4): sort, merge overlapping AND adjacent ranges, drop empty ones. */
static void rl_reduce(rlist *r)
{
    if (r->n == 0) return;
    pair_t *p = (pair_t *)malloc(sizeof(pair_t) * (size_t)r->n);
    int32_t m = 0;
    for (int32_t i = 0; i < r->n; i++)
        if (r->end[i] >= r->start[i]) { p[m].s = r->start[i]; p[m].e = r->end[i]; m++; }
    qsort(p, (size_t)m, sizeof(pair_t), cmp_pair);
    int32_t k = 0;
    for (int32_t i = 0; i < m; i++) {
        if (k > 0 && p[i].s <= r->end[k - 1] + 1) {
            if (p[i].e > r->end[k - 1]) r->end[k - 1] = p[i].e;
        } else { r->start[k] = p[i].s; r->end[k] = p[i].e; k++; }
    }
    r->n = k;
    free(p);
}

/* IRanges::union(x, y) = reduce(c(x, y)) */
static void rl_union_into(rlist *acc, const rlist *y)
{
    for (int32_t i = 0; i < y->n; i++) rl_push(acc, y->start[i], y->end[i]);
    rl_reduce(acc);
}

/* Biostrings::trim on views: clip to [1, L] (NanoTel.R:266-302 author's note, :338, :352) */
static void rl_trim(rlist *r, int32_t L)
{
    for (int32_t i = 0; i < r->n; i++) {
        if (r->start[i] < 1) r->start[i] = 1;
        if (r->end[i] > L) r->end[i] = L;
    }
}

/* ---------------------------------------------------------------- Biostrings::matchPattern (App. B.2, B.3)
 * s: codes, 1-based (s[1..L]); all alignments i in [1-k, L-m+1+k] with at most k mismatches, positions outside
 * [1, L] counting as mismatches; ascending start order; overlapping hits included. */
static void match_pattern(const uint8_t *s, int32_t L, const uint8_t *pat, int32_t m, int32_t k, int fixed,
                          rlist *out)
{
    for (int32_t i = 1 - k; i <= L - m + 1 + k; i++) {
        int32_t mis = 0;
        for (int32_t j = 0; j < m && mis <= k; j++) {
            int32_t pos = i + j;
            if (pos < 1 || pos > L) { mis++; continue; }
            uint8_t c = s[pos], p = pat[j];
            int eq = fixed ? (p == c) : ((p & c) != 0);
            if (!eq) mis++;
        }
        if (mis <= k) rl_push(out, i, i + m - 1);
    }
}

/* ---------------------------------------------------------------- patterns */
typedef struct {
    uint8_t code[NTLO_MAX_PATLEN];
    int32_t m;
    int fixed;          /* !str_detect(pat, "[WSMKRYBDHVN]")  NanoTel.R:334,348,368,384 (case sensitive) */
} pat_t;

typedef struct {
    pat_t pat[NTLO_MAX_PATTERNS]; int32_t n_pat;     /* in CLI order, duplicates kept */
    pat_t upat[NTLO_MAX_PATTERNS]; int32_t n_upat;   /* unique() of the list          */
    int is_list;                                      /* length(cur_patterns) > 1  NanoTel.R:2324 */
    pat_t tvr[NTLO_MAX_PATTERNS]; int32_t n_tvr;
    pat_t utvr[NTLO_MAX_PATTERNS]; int32_t n_utvr;
    int tvr_is_list;
} patset_t;

static int parse_pat(const char *str, pat_t *out)
{
    size_t m = strlen(str);
    if (m == 0 || m > NTLO_MAX_PATLEN) return -1;
    out->m = (int32_t)m;
    out->fixed = 1;
    for (size_t j = 0; j < m; j++) {
        uint8_t c = g_code[(unsigned char)str[j]];
        if (c == 0xFF || c > 15) return -1;
        out->code[j] = c;
        if (strchr("WSMKRYBDHVN", str[j])) out->fixed = 0;
    }
    return 0;
}

static int build_patset(const ntlo_params *p, patset_t *ps)
{
    init_tables();
    if (p->n_patterns < 1 || p->n_patterns > NTLO_MAX_PATTERNS) return -1;
    if (p->n_tvr < 0 || p->n_tvr > NTLO_MAX_PATTERNS) return -1;
    ps->n_pat = p->n_patterns; ps->is_list = p->n_patterns > 1;
    ps->n_upat = 0;
    for (int i = 0; i < p->n_patterns; i++) {
        if (parse_pat(p->patterns[i], &ps->pat[i])) return -2;
        int dup = 0;
        for (int j = 0; j < i; j++) if (strcmp(p->patterns[i], p->patterns[j]) == 0) dup = 1;
        if (!dup) ps->upat[ps->n_upat++] = ps->pat[i];
    }
    ps->n_tvr = p->n_tvr; ps->tvr_is_list = p->n_tvr > 1; ps->n_utvr = 0;
    for (int i = 0; i < p->n_tvr; i++) {
        if (parse_pat(p->tvr[i], &ps->tvr[i])) return -2;
        int dup = 0;
        for (int j = 0; j < i; j++) if (strcmp(p->tvr[i], p->tvr[j]) == 0) dup = 1;
        if (!dup) ps->utvr[ps->n_utvr++] = ps->tvr[i];
    }
    return 0;
}

/* ---------------------------------------------------------------- get_density_iranges (NanoTel.R:308-397)
 * Returns the `ranges` element (mp_all).  with_mismatch -> max.mismatch = 1.  use_tvr -> tvr_patterns != NULL. */
static void get_density_iranges(const uint8_t *s, int32_t L, const patset_t *ps, int with_mismatch, int use_tvr,
                                rlist *mp_all)
{
    int32_t k = with_mismatch ? 1 : 0;
    rl_init(mp_all);
    if (ps->is_list) {                                        /* :327-345 */
        for (int i = 0; i < ps->n_upat; i++) {
            rlist cur; rl_init(&cur);
            match_pattern(s, L, ps->upat[i].code, ps->upat[i].m, k, ps->upat[i].fixed, &cur);
            if (!ps->upat[i].fixed || k > 0) rl_trim(&cur, L);
            rl_union_into(mp_all, &cur);
            rl_free(&cur);
        }
        rl_reduce(mp_all);                                    /* :345 */
    } else {                                                  /* :347-356 */
        match_pattern(s, L, ps->pat[0].code, ps->pat[0].m, k, ps->pat[0].fixed, mp_all);
        if (!ps->pat[0].fixed || k > 0) { rl_trim(mp_all, L); rl_reduce(mp_all); }
        /* else: raw, un-merged hit list is kept (:349-354 skips the union) */
    }
    if (use_tvr) {                                            /* :360-393, exact (default max.mismatch) */
        if (ps->tvr_is_list) {
            for (int i = 0; i < ps->n_utvr; i++) {
                rlist cur; rl_init(&cur);
                match_pattern(s, L, ps->utvr[i].code, ps->utvr[i].m, 0, ps->utvr[i].fixed, &cur);
                if (!ps->utvr[i].fixed || k > 0) rl_trim(&cur, L);
                rl_union_into(mp_all, &cur);
                rl_free(&cur);
            }
            rl_reduce(mp_all);
        } else {
            rlist cur; rl_init(&cur);
            match_pattern(s, L, ps->tvr[0].code, ps->tvr[0].m, 0, ps->tvr[0].fixed, &cur);
            if (!ps->tvr[0].fixed || k > 0) {                 /* :387-390: only unioned in under this condition */
                rl_trim(&cur, L);
                rl_union_into(mp_all, &cur);
            }
            rl_reduce(mp_all);                                /* :391 */
            rl_free(&cur);
        }
    }
}

/* ---------------------------------------------------------------- coverage prefix used for get_sub_density
 * get_sub_density (NanoTel.R:449-468): sum(width(intersect(sub_irange, ranges))) / width(sub_irange).
 * intersect() normalises, so the numerator is the number of positions of [a, b] covered by >= 1 range. All range
 * lists are inside [1, L] here (raw exact hits cannot overhang; everything else is trimmed). */
static int32_t *build_cov_prefix(const rlist *r, int32_t L)
{
    int32_t *d = (int32_t *)calloc((size_t)L + 2, sizeof(int32_t));
    for (int32_t i = 0; i < r->n; i++) {
        int32_t a = r->start[i] < 1 ? 1 : r->start[i], b = r->end[i] > L ? L : r->end[i];
        if (b < a) continue;
        d[a] += 1; d[b + 1] -= 1;
    }
    int32_t depth = 0, acc = 0;
    for (int32_t i = 1; i <= L; i++) { depth += d[i]; acc += depth > 0; d[i] = acc; }
    d[0] = 0;
    return d;                 /* d[i] = covered positions in [1, i] */
}

static int32_t covered(const int32_t *pre, int32_t L, int32_t a, int32_t b)
{
    int32_t lo = a < 1 ? 1 : a, hi = b > L ? L : b;
    if (hi < lo) return 0;
    return pre[hi] - pre[lo - 1];
}
static double sub_density(const int32_t *pre, int32_t L, int32_t a, int32_t b)
{
    return (double)covered(pre, L, a, b) / (double)(b - a + 1);
}

/* ---------------------------------------------------------------- split_telo (NanoTel.R:199-227) */
int32_t ntlo_split_telo(int32_t len, int32_t S, int32_t *starts, int32_t *ends, int32_t max_win)
{
    if (len < 1 || S < 1) return 0;
    int32_t n = (len - 1) / S + 1;                     /* seq(1, len, by = S) */
    int32_t last_start = 1 + (n - 1) * S;
    if ((double)(len - last_start) < (double)S / 2.0) n -= 1;   /* :220-224 */
    for (int32_t i = 0; i < n && i < max_win; i++) {
        starts[i] = 1 + i * S;
        ends[i] = (i == n - 1) ? len : starts[i] + S - 1;
    }
    return n;
}
int32_t ntlo_count_windows(int32_t len, int32_t S) { return ntlo_split_telo(len, S, NULL, NULL, 0); }

/* ---------------------------------------------------------------- window table (analyze_subtelos :717-766) */
typedef struct {
    int32_t n;
    int32_t *ws, *we, *cls;    /* 1-based: index 1..n */
    double *den;
} wtab;

static void wtab_build(wtab *w, int32_t L, int32_t S, double min_density, const int32_t *pre, int32_t *counts_out)
{
    int32_t n = ntlo_count_windows(L, S);
    w->n = n;
    w->ws = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2));
    w->we = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2));
    w->cls = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 2));
    w->den = (double *)malloc(sizeof(double) * ((size_t)n + 2));
    ntlo_split_telo(L, S, w->ws + 1, w->we + 1, n);
    for (int32_t i = 1; i <= n; i++) {
        int32_t c = covered(pre, L, w->ws[i], w->we[i]);
        double d = (double)c / (double)(w->we[i] - w->ws[i] + 1);       /* :467 */
        int32_t cl = CLS_TELO;                                           /* :751-758 */
        if (d < min_density) cl = (d < 0.1) ? CLS_SKIP : CLS_NONE;
        w->den[i] = d; w->cls[i] = cl;
        if (counts_out) counts_out[i - 1] = c;
    }
}
static void wtab_free(wtab *w) { free(w->ws); free(w->we); free(w->cls); free(w->den); }

/* ---------------------------------------------------------------- find_telo_position (NanoTel.R:973-1077) */
static void find_telo_position(const wtab *w, double min_in_a_row, double min_density_score,
                               int32_t *ps, int32_t *pe)
{
    int32_t n = w->n;
    double score = 0.0;
    int32_t start = -1, end = -1, in_a_row = 0;
    int32_t start_end_diff = n >= 1 ? w->we[1] - w->ws[1] : 0;           /* :997 */
    int32_t end_position = 0;
    for (int32_t i = 1; i <= n; i++) {                                    /* :1003-1025 */
        if (w->cls[i] != CLS_TELO) { score = 0; start = -1; in_a_row = 0; }
        else {
            in_a_row += 1;
            score = score + w->den[i];
            if (start == -1) start = w->ws[i];
        }
        if ((double)in_a_row >= min_in_a_row && score >= min_density_score) { end_position = i + 1; break; }
    }
    if (end_position == 0) { *ps = -1; *pe = -1; return; }                /* :1026-1028 */
    end = -1; score = 0.0; in_a_row = 0;
    if ((double)end_position >= (double)n - min_in_a_row + 1.0) {         /* :1037-1044 */
        int32_t i = n;
        while (w->cls[i] != CLS_TELO && i > end_position) i -= 1;
        end = w->we[i];
    } else {                                                              /* :1046-1068 */
        for (int32_t i = n; i >= end_position; i--) {
            if (w->cls[i] != CLS_TELO) { score = 0.0; end = -1; in_a_row = 0; }
            else {
                in_a_row += 1;
                score = score + w->den[i];
                if (end == -1) end = w->we[i];
            }
            if ((double)in_a_row >= min_in_a_row && score >= min_density_score) break;
        }
    }
    if (start > end) end = start + start_end_diff;                        /* :1072-1074 */
    *ps = start; *pe = end;
}

/* ---------------------------------------------------------------- find_left_telo (NanoTel.R:906-959) */
static void find_left_telo(const wtab *w, int32_t *ps, int32_t *pe)
{
    const int32_t max_diff = 200;
    int32_t n = w->n, start = 1, end = 1, last_i = 1;
    for (int32_t i = 1; i <= n; i++) {
        if (w->ws[i] > max_diff) { *ps = -1; *pe = -1; return; }
        if (w->cls[i] != CLS_TELO) continue;
        start = w->ws[i]; last_i = i; break;
    }
    int32_t last_i_start = last_i;
    /* for (i in last_i:nrow): with nrow == 0 this is 1:0 and the NA row at i = 1 breaks immediately */
    for (int32_t i = last_i; i <= n; i++) {
        if (w->cls[i] != CLS_TELO) break;
        end = w->we[i];
    }
    if (start > end) end = start + (w->we[last_i_start] - w->ws[last_i_start]);
    *ps = start; *pe = end;
}

/* ---------------------------------------------------------------- find_right_telo (NanoTel.R:843-899)
 * returns -1 if the reference would stop() (zero-row table: `if (logical(0))`, :859-861). */
static int find_right_telo(const wtab *w, int32_t L, int32_t *ps, int32_t *pe)
{
    const int32_t max_diff = 200;
    int32_t n = w->n, start = 1, end = 1, last_i = 1;
    if (n == 0) return -1;
    for (int32_t i = n; i >= 1; i--) {
        if (w->we[i] < L - max_diff) { *ps = -1; *pe = -1; return 0; }
        if (w->cls[i] != CLS_TELO) continue;
        end = w->we[i]; last_i = i; break;
    }
    for (int32_t i = last_i; i >= 1; i--) {
        if (w->cls[i] != CLS_TELO) break;
        start = w->ws[i]; last_i = i;
    }
    if (start > end) end = start + (w->we[last_i] - w->ws[last_i]);
    *ps = start; *pe = end;
    return 0;
}

/* ---------------------------------------------------------------- get_accurate_end (NanoTel.R:1692-1721) */
static int32_t get_accurate_end(int32_t telo_end, const rlist *r)
{
    if (telo_end == -1) return -1;
    int32_t e_index = telo_end;
    int found = 0; int32_t mx = 0;
    for (int32_t i = 0; i < r->n; i++)
        if (r->end[i] >= e_index - 99 && r->end[i] <= e_index) { if (!found || r->end[i] > mx) mx = r->end[i]; found = 1; }
    if (found) e_index = mx;
    found = 0;
    for (int32_t i = 0; i < r->n; i++)
        if (r->end[i] >= telo_end + 1 && r->end[i] <= telo_end + 50) { if (!found || r->end[i] > mx) mx = r->end[i]; found = 1; }
    if (found) e_index = mx;
    return e_index;
}

static int min_start_in(const rlist *r, int32_t lo, int32_t hi, int32_t *out)
{
    int found = 0; int32_t mn = 0;
    for (int32_t i = 0; i < r->n; i++)
        if (r->start[i] >= lo && r->start[i] <= hi) { if (!found || r->start[i] < mn) mn = r->start[i]; found = 1; }
    if (found) *out = mn;
    return found;
}

/* ---------------------------------------------------------------- get_accurate_start (NanoTel.R:1726-1764) */
static int32_t get_accurate_start(int32_t telo_start, const rlist *r, const int32_t *pre, int32_t L)
{
    if (telo_start == -1) return telo_start;
    int32_t s_index = telo_start;
    double first_50 = sub_density(pre, L, telo_start, telo_start + 49);   /* IRanges(start, width = 50) */
    if (first_50 < 0.3) {
        min_start_in(r, s_index + 48, s_index + 99, &telo_start);
        min_start_in(r, s_index + 33, s_index + 48, &telo_start);
    } else {
        min_start_in(r, s_index, s_index + 99, &telo_start);
        if (first_50 >= 0.72) min_start_in(r, s_index - 36, s_index - 1, &telo_start);
    }
    return telo_start;
}

/* ---------------------------------------------------------------- 18-bp re-match (NanoTel.R:496-697)
 * matchPattern on subseq(read, a, b) with the DEFAULT fixed = TRUE (byte equality), own out-of-bounds rule,
 * hits NOT trimmed.  Returns 1 and the min start / max end (absolute coordinates) if any pattern hits. */
static int step_window(const uint8_t *s, int32_t a, int32_t b, const patset_t *ps, int k, int use_tvr,
                       int32_t *min_start_abs, int32_t *max_end_abs)
{
    int32_t W = b - a + 1;
    int any = 0; int32_t mn = 0, mx = 0;
    const uint8_t *sub = s + (a - 1);            /* sub[1..W] */
    for (int pass = 0; pass < 2; pass++) {
        const pat_t *pl = pass == 0 ? ps->pat : ps->tvr;
        int32_t np = pass == 0 ? ps->n_pat : (use_tvr ? ps->n_tvr : 0);
        int32_t kk = pass == 0 ? k : 0;          /* TVR patterns: exact (:519, :566) */
        for (int32_t q = 0; q < np; q++) {
            rlist h; rl_init(&h);
            match_pattern(sub, W, pl[q].code, pl[q].m, kk, /*fixed=*/1, &h);
            for (int32_t i = 0; i < h.n; i++) {
                if (!any || h.start[i] < mn) mn = h.start[i];
                if (!any || h.end[i] > mx) mx = h.end[i];
                any = 1;
            }
            rl_free(&h);
        }
    }
    if (any) { *min_start_abs = mn + a - 1; *max_end_abs = mx + a - 1; }
    return any;
}

/* search_left_patterns (NanoTel.R:576-633): subseq_width 18, step_size 10, max_steps 4 */
static int32_t search_left_patterns(const uint8_t *s, int32_t L, int32_t start_index, const patset_t *ps, int k,
                                    int use_tvr)
{
    int32_t subseq_start = start_index - 18 > 1 ? start_index - 18 : 1;
    int32_t new_start = start_index;
    for (int i = 1; i <= 4; i++) {
        int32_t curr_end = subseq_start + 18 - 1 < L ? subseq_start + 18 - 1 : L;
        int32_t mn, mx;
        if (!step_window(s, subseq_start, curr_end, ps, k, use_tvr, &mn, &mx)) break;
        new_start = mn;
        int32_t nn = subseq_start - 10 + 1 > 1 ? subseq_start - 10 + 1 : 1;
        if (nn == subseq_start) break;
        subseq_start = nn;
    }
    return new_start;
}

/* search_right_patterns (NanoTel.R:635-697) */
static int32_t search_right_patterns(const uint8_t *s, int32_t L, int32_t end_index, const patset_t *ps, int k,
                                     int use_tvr)
{
    int32_t subseq_end = end_index + 18 < L ? end_index + 18 : L;
    int32_t new_end = end_index;
    for (int i = 1; i <= 4; i++) {
        int32_t curr_start = subseq_end - 18 + 1 > 1 ? subseq_end - 18 + 1 : 1;
        int32_t mn, mx;
        if (!step_window(s, curr_start, subseq_end, ps, k, use_tvr, &mn, &mx)) break;
        new_end = mx;
        int32_t nn = subseq_end + 10 + 1 < L ? subseq_end + 10 + 1 : L;
        if (nn == subseq_end) break;
        subseq_end = nn;
    }
    return new_end;
}

/* ---------------------------------------------------------------- find_telo_position_wraper (NanoTel.R:1080-1155)
 * returns 0, or -1 if the reference would stop() */
static int telo_position_wrapper(const uint8_t *s, int32_t L, const patset_t *ps, int with_mismatch, int use_tvr,
                                 int32_t S, int right_edge, const wtab *w, const rlist *ranges, const int32_t *pre,
                                 ntlo_track *t)
{
    int32_t ts, te;
    find_telo_position(w, 3.0, 2.0, &ts, &te);                             /* :1084-1086 */
    double telo_density = sub_density(pre, L, ts, te);                     /* :1099 */
    int32_t num_rows = (te - ts + 1) / S;                                  /* :1103 width %/% S */
    if (telo_density < 0.85 && num_rows > 5) {                             /* :1104-1110 */
        double min_rows = num_rows <= 7 ? (double)(num_rows - 2) : 7.0;
        double min_density = 0.6 * min_rows;
        find_telo_position(w, min_rows, min_density, &ts, &te);
    }
    t->coarse_start = ts; t->coarse_end = te;
    int32_t start_acc = get_accurate_start(ts, ranges, pre, L);            /* :1119 */
    int32_t end_acc = get_accurate_end(te, ranges);                        /* :1120 */
    if (start_acc > end_acc) end_acc = start_acc;                          /* :1122-1124 */
    ts = start_acc; te = end_acc;                                          /* :1126 */
    t->acc_start = ts; t->acc_end = te;
    t->acc_density = sub_density(pre, L, ts, te);
    if (te - ts + 1 < 100) {                                               /* :1129-1136 */
        if (right_edge) { if (find_right_telo(w, L, &ts, &te)) return -1; }
        else find_left_telo(w, &ts, &te);
    }
    t->edge_start = ts; t->edge_end = te;
    int k = with_mismatch ? 1 : 0;
    int32_t e2, s2;
    if (te < L) e2 = search_right_patterns(s, L, te + 1, ps, k, use_tvr);  /* :1140-1144 */
    else e2 = te;
    if (ts > 1) s2 = search_left_patterns(s, L, ts - 1, ps, k, use_tvr);   /* :1145-1149 */
    else s2 = ts;
    if (e2 < s2 - 1) return -1;                                            /* IRanges(start, end) would stop() */
    t->start = s2; t->end = e2;                                            /* :1152 */
    return 0;
}

/* ---------------------------------------------------------------- analyze_read (NanoTel.R:1774-1976) */
static int analyze_codes(const ntlo_params *p, const patset_t *ps, const uint8_t *s /*1-based*/, int32_t L,
                         ntlo_read *out, int32_t *win_counts, int32_t max_win)
{
    memset(out, 0, sizeof *out);
    out->length = L;
    int n_tracks = ps->n_tvr > 0 ? 3 : 2;
    int32_t S = p->subseq_length;
    out->n_win = ntlo_count_windows(L, S);
    if (out->n_win == 0) out->flags |= NTLO_FLAG_NO_WINDOWS;
    int err = 0;
    for (int tr = 0; tr < n_tracks; tr++) {
        int with_mismatch = tr >= 1, use_tvr = tr == 2;
        rlist ranges;
        get_density_iranges(s, L, ps, with_mismatch, use_tvr, &ranges);    /* :734 */
        int32_t *pre = build_cov_prefix(&ranges, L);
        wtab w;
        int32_t *co = NULL;
        if (win_counts && out->n_win <= max_win) co = win_counts + (size_t)tr * (size_t)max_win;
        wtab_build(&w, L, S, p->min_density, pre, co);                      /* :737-764 */
        ntlo_track *t = &out->t[tr];
        t->n_ranges = ranges.n;
        if (telo_position_wrapper(s, L, ps, with_mismatch, use_tvr, S, p->right_edge, &w, &ranges, pre, t)) err = 1;
        else t->density = sub_density(pre, L, t->start, t->end);            /* :1840-1844 */
        wtab_free(&w); free(pre); rl_free(&ranges);
        if (err) break;
    }
    if (err) { out->flags |= NTLO_FLAG_REF_ERROR; out->keep = 0; return 0; }
    int32_t mxw = 0;
    for (int tr = 0; tr < n_tracks; tr++) {                                 /* :1847, :1857 */
        int32_t wd = out->t[tr].end - out->t[tr].start + 1;
        if (wd > mxw) mxw = wd;
    }
    out->keep = mxw < 30 ? 0 : 1;
    return 0;
}

static uint8_t *to_codes(const char *seq, int32_t len)
{
    uint8_t *s = (uint8_t *)malloc((size_t)len + 2);
    s[0] = 0;
    for (int32_t i = 0; i < len; i++) {
        uint8_t c = g_code[(unsigned char)seq[i]];
        if (c == 0xFF) { free(s); return NULL; }
        s[i + 1] = c;
    }
    s[len + 1] = 0;
    return s;
}

int ntlo_analyze_read(const ntlo_params *p, const char *seq, int32_t len, ntlo_read *out, int32_t *win_counts,
                      int32_t max_win)
{
    patset_t ps;
    int rc = build_patset(p, &ps);
    if (rc) return rc;
    if (len < 1 || p->subseq_length < 1) return -3;          /* seq(1, 0, by = S) stops in R (:216) */
    uint8_t *s = to_codes(seq, len);
    if (!s) return -4;
    rc = analyze_codes(p, &ps, s, len, out, win_counts, max_win);
    free(s);
    return rc;
}

/* ---------------------------------------------------------------- filter (NanoTel.R:2083-2163) */
static int filter_codes(const ntlo_params *p, const patset_t *ps, const uint8_t *s, int32_t L)
{
    if (L < 1000) return 0;                                    /* :2124 */
    int32_t a, b;
    if (p->right_edge) { b = L - 70; a = b - 200 + 1; }        /* subseq(end = -(70+1), width = 200) :2131-2134 */
    else { a = 71; b = 270; }                                  /* subseq(start = 71, width = 200)   :2136 */
    const uint8_t *sub = s + (a - 1);
    int32_t W = b - a + 1;
    rlist all; rl_init(&all);
    int n = ps->is_list ? ps->n_upat : 1;
    for (int i = 0; i < n; i++) {                              /* :2088-2098, fixed = FALSE always */
        const pat_t *q = ps->is_list ? &ps->upat[i] : &ps->pat[0];
        rlist cur; rl_init(&cur);
        match_pattern(sub, W, q->code, q->m, 0, /*fixed=*/0, &cur);
        rl_union_into(&all, &cur);
        rl_free(&cur);
    }
    int32_t sum = 0;
    for (int32_t i = 0; i < all.n; i++) sum += all.end[i] - all.start[i] + 1;
    rl_free(&all);
    double total_density = (double)sum / (double)W;            /* :2100 nchar(sequence) = 200 */
    return total_density >= p->min_density * 0.8;              /* :2143, :2101 */
}

int ntlo_filter_read(const ntlo_params *p, const char *seq, int32_t len)
{
    patset_t ps;
    if (build_patset(p, &ps)) return -1;
    uint8_t *s = to_codes(seq, len);
    if (!s) return -4;
    int r = filter_codes(p, &ps, s, len);
    free(s);
    return r;
}

void ntlo_revcomp(const char *in, int32_t len, char *out)
{
    init_tables();
    for (int32_t i = 0; i < len; i++) out[i] = (char)g_comp[(unsigned char)in[len - 1 - i]];
}

int32_t ntlo_match_pattern(const char *seq, int32_t len, const char *pat, int32_t max_mismatch, int32_t fixed,
                           int32_t *starts, int32_t max_hits)
{
    init_tables();
    pat_t q;
    if (parse_pat(pat, &q)) return -1;
    uint8_t *s = to_codes(seq, len);
    if (!s) return -4;
    rlist h; rl_init(&h);
    match_pattern(s, len, q.code, q.m, max_mismatch, fixed, &h);
    int32_t n = h.n;
    for (int32_t i = 0; i < n && i < max_hits; i++) starts[i] = h.start[i];
    rl_free(&h); free(s);
    return n;
}

/* get_sub_density (NanoTel.R:449-468) on an explicit range list, for the worked example in its own comment
 * (NanoTel.R:459-464): the same reduce / coverage-prefix / division code analyze_read uses. */
double ntlo_sub_density_ranges(const int32_t *starts, const int32_t *ends, int32_t n, int32_t L, int32_t a, int32_t b)
{
    rlist r; rl_init(&r);
    for (int32_t i = 0; i < n; i++) rl_push(&r, starts[i], ends[i]);
    rl_reduce(&r);
    int32_t *pre = build_cov_prefix(&r, L);
    const double d = sub_density(pre, L, a, b);
    free(pre); rl_free(&r);
    return d;
}

/* ---------------------------------------------------------------- batch driver (CPU baseline), pthreads */
typedef struct {
    const ntlo_params *p; const patset_t *ps; const char *const *seqs; const int32_t *lens; int32_t n;
    int32_t do_rc, use_filter; ntlo_read *out; uint8_t *pass; const int64_t *win_off; int32_t *win_counts;
    int n_tracks; volatile int32_t *next; volatile int *fail;
} batch_job;

static void batch_one(const batch_job *J, int32_t i)
{
    int32_t L = J->lens[i];
    memset(&J->out[i], 0, sizeof J->out[i]);
    if (J->pass) J->pass[i] = 0;
    if (L < 1) { *J->fail = 1; return; }
    char *tmp = NULL;
    const char *sq = J->seqs[i];
    if (J->do_rc) { tmp = (char *)malloc((size_t)L); ntlo_revcomp(sq, L, tmp); sq = tmp; }
    uint8_t *s = to_codes(sq, L);
    free(tmp);
    if (!s) { *J->fail = 1; return; }
    int ok = 1;
    if (J->use_filter) ok = filter_codes(J->p, J->ps, s, L);
    if (J->pass) J->pass[i] = (uint8_t)ok;
    if (ok) {
        int32_t *wc = NULL; int32_t mw = 0;
        if (J->win_off && J->win_counts) {
            mw = (int32_t)((J->win_off[i + 1] - J->win_off[i]) / J->n_tracks);
            wc = J->win_counts + J->win_off[i];
        }
        analyze_codes(J->p, J->ps, s, L, &J->out[i], wc, mw);
    } else {
        J->out[i].length = L;
    }
    free(s);
}

static void *batch_worker(void *arg)
{
    const batch_job *J = (const batch_job *)arg;
    for (;;) {
        int32_t i0 = __sync_fetch_and_add(J->next, 4);
        if (i0 >= J->n) break;
        for (int32_t i = i0; i < i0 + 4 && i < J->n; i++) batch_one(J, i);
    }
    return NULL;
}

int ntlo_scan_batch(const ntlo_params *p, const char *const *seqs, const int32_t *lens, int32_t n, int32_t do_rc,
                    int32_t use_filter, ntlo_read *out, uint8_t *pass, const int64_t *win_off, int32_t *win_counts,
                    int32_t n_threads)
{
    patset_t ps;
    int rc = build_patset(p, &ps);
    if (rc) return rc;
    volatile int32_t next = 0; volatile int fail = 0;
    batch_job J = {p, &ps, seqs, lens, n, do_rc, use_filter, out, pass, win_off, win_counts,
                   ps.n_tvr > 0 ? 3 : 2, &next, &fail};
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < n_threads - 1; t++)
        if (pthread_create(&th[started], NULL, batch_worker, &J) == 0) started++;
    batch_worker(&J);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return fail ? -4 : 0;
}

/* ---------------------------------------------------------------- Serial numbering (NanoTel.R:2050-2069, 2234-2258) */
int32_t ntlo_assign_serials(const int32_t *keep, int32_t n, int32_t serial_start, int32_t prev_max_serial,
                            int32_t *serial, int32_t *row_order, int32_t *next_serial, int32_t *max_serial)
{
    int32_t rows = 0, mx = prev_max_serial;
    for (int32_t i = 0; i < n; i++) serial[i] = 0;
    if (n < 8) {                                               /* :2236-2239 */
        int32_t cur = serial_start;
        for (int32_t i = 0; i < n; i++)
            if (keep[i]) { serial[i] = cur; row_order[rows++] = i; if (cur > mx) mx = cur; cur++; }
    } else {                                                   /* :2242-2254: group g = reads g, g+8, ... */
        int32_t offset = 0;
        for (int32_t g = 0; g < 8; g++) {
            int32_t cur = serial_start + offset, size = 0;
            for (int32_t i = g; i < n; i += 8) {
                size++;
                if (keep[i]) { serial[i] = cur; row_order[rows++] = i; if (cur > mx) mx = cur; cur++; }
            }
            offset += size;
        }
    }
    *max_serial = mx;
    /* :2258 serial_start <- max(df_summary$Serial) + 1; with no row yet R yields -Inf (reference bug):
     * documented, not imitated -- serial_start is left unchanged in that case. */
    *next_serial = mx >= 1 ? mx + 1 : serial_start;
    return rows;
}
