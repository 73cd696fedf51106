/*
 * nanotel_oracle.h -- CPU restatement of NanoTel.R's per-read telomere detection.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.  The product
 * path (telomere-analyzer_b200/) never links, imports or calls it.
 *
 * Parity status: stages up to NanoTel.R:1126 (get_accurate_start/end) and the
 * density arithmetic are PINNED by the reference's own golden file
 * Example/Example_output/summary.csv (see tests/test_oracle_golden.py).
 * Everything after (NanoTel.R:1129-1152 edge fallback and 18-bp re-match), IUPAC
 * patterns, TVR track, filter and Serial numbering is "parity unpinned":
 * no artefact of the reference exercises it and R/Bioconductor cannot run here.
 *
 * Third-party semantics restated (not vendored in /root/reference):
 * Biostrings 2.66-2.68 matchPattern/trim/subseq/reverseComplement,
 * IRanges 2.32-2.34 union/intersect/reduce (README.md:83, run.log:8).
 */
#ifndef NANOTEL_ORACLE_H
#define NANOTEL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTLO_MAX_PATTERNS 16
#define NTLO_MAX_PATLEN   18      /* NanoTel.R:589,647 assert */

/* flags */
#define NTLO_FLAG_REF_ERROR   1   /* NanoTel.R would stop() on this read */
#define NTLO_FLAG_NO_WINDOWS  2   /* split_telo produced an empty table  */

typedef struct {
    int32_t n_patterns;                 /* --patterns tokens, in order, NOT de-duplicated   */
    const char *patterns[NTLO_MAX_PATTERNS];
    int32_t n_tvr;                      /* --tvr_patterns tokens (0 = NULL)                  */
    const char *tvr[NTLO_MAX_PATTERNS];
    double  min_density;                /* --min_density                                     */
    int32_t subseq_length;              /* --subseq_length                                   */
    int32_t right_edge;                 /* --check_right_edge                                */
} ntlo_params;

typedef struct {
    int32_t coarse_start, coarse_end;   /* find_telo_position, after the optional re-run (NanoTel.R:1084-1110) */
    int32_t acc_start, acc_end;         /* after get_accurate_start/end (NanoTel.R:1119-1126) -- golden stage  */
    int32_t edge_start, edge_end;       /* after the <100 bp edge fallback (NanoTel.R:1129-1136)               */
    int32_t start, end;                 /* final, after search_left/right_patterns (NanoTel.R:1140-1152)       */
    double  acc_density;                /* get_sub_density(acc interval)  -- golden stage                      */
    double  density;                    /* get_sub_density(final interval) (NanoTel.R:1840-1844)               */
    int32_t n_ranges;                   /* length(ranges) of this track (raw hits or reduced runs)             */
    int32_t pad;
} ntlo_track;

typedef struct {
    int32_t keep;                       /* 1 = a summary row is emitted (NanoTel.R:1847-1868)                  */
    int32_t flags;
    int32_t n_win;
    int32_t length;
    ntlo_track t[3];                    /* A (exact), B (1 mismatch), C (1 mismatch + TVR; only if n_tvr > 0)  */
} ntlo_read;

/* Analyse one read (ASCII, already reverse-complemented if --rc).
 * win_counts: optional, [n_tracks][max_win] covered-base counts per window (int32), row stride max_win.
 * Returns 0, or <0 on invalid parameters. */
int ntlo_analyze_read(const ntlo_params *p, const char *seq, int32_t len,
                      ntlo_read *out, int32_t *win_counts, int32_t max_win);

/* split_telo (NanoTel.R:199-227): writes up to max_win (start,end) pairs; returns n_win. */
int32_t ntlo_split_telo(int32_t len, int32_t S, int32_t *starts, int32_t *ends, int32_t max_win);
int32_t ntlo_count_windows(int32_t len, int32_t S);

/* filter_reads / filter_density (NanoTel.R:2083-2163) for one (already rc'd) read: 1 = keep. */
int ntlo_filter_read(const ntlo_params *p, const char *seq, int32_t len);

/* Biostrings::reverseComplement on ASCII IUPAC (NanoTel.R:2219-2221). */
void ntlo_revcomp(const char *in, int32_t len, char *out);

/* matchPattern restatement exposed for unit tests: returns number of hits, writes starts (1-based, may be 0). */
int32_t ntlo_match_pattern(const char *seq, int32_t len, const char *pat, int32_t max_mismatch, int32_t fixed,
                           int32_t *starts, int32_t max_hits);

/* get_sub_density (NanoTel.R:449-468) on an explicit range list (unit tests: the example at NanoTel.R:459-464). */
double ntlo_sub_density_ranges(const int32_t *starts, const int32_t *ends, int32_t n, int32_t L, int32_t a, int32_t b);

/* Batch driver: analyse n reads with n_threads OpenMP threads (CPU baseline).
 * do_rc / use_filter follow NanoTel.R:2219-2232.  pass[i] = 0 if the read was filtered out.
 * win_off (n+1 entries, may be NULL) gives the per-read offset into win_counts, laid out
 * [read][track][n_win] (tracks contiguous per read). */
int ntlo_scan_batch(const ntlo_params *p, const char *const *seqs, const int32_t *lens, int32_t n,
                    int32_t do_rc, int32_t use_filter, ntlo_read *out, uint8_t *pass,
                    const int64_t *win_off, int32_t *win_counts, int32_t n_threads);

/* Serial numbering of one chunk (NanoTel.R:2050-2069, 2234-2258).  keep[] over the post-filter reads of the
 * chunk; writes serial[i] (0 = no row) and the row order (indices into the chunk, group-major); returns the
 * number of rows; *next_serial = max(Serial so far)+1 as NanoTel.R:2258 (unchanged if no row exists yet). */
int32_t ntlo_assign_serials(const int32_t *keep, int32_t n, int32_t serial_start, int32_t prev_max_serial,
                            int32_t *serial, int32_t *row_order, int32_t *next_serial, int32_t *max_serial);

#ifdef __cplusplus
}
#endif
#endif
