/*
 * ntl_dev.h -- parameter blocks shared by the host side and the sm_100a kernels of libnanotel_b200.
 *
 * Packed read layout in HBM (our choice; nothing like it exists in the reference):
 *   position p of a read (1-based, after --rc) lives at bit index p of the read's bit stream; bit 0 is a pad so
 *   that an alignment starting one base before the read (Biostrings out-of-bounds hit) has a
 *   place for its hit bit.  The stream is cut into 32-bit words (LSB first) and stored as "quads" of 128 bit
 *   positions:
 *       2-bit reads (ACGT only):  quad = { lo[4], hi[4] }               32 bytes  -> one LDG.256 per lane
 *       4-bit reads (IUPAC):      quad = { A[4], C[4], G[4], T[4] }     64 bytes  (Biostrings code bits)
 *   2-bit code = (ASCII >> 1) & 3 :  A = 0, C = 1, T = 2, G = 3  (lo = bit 0, hi = bit 1); complement = flip hi.
 *   A warp streams a read in chunks of 32 quads = 4096 positions = 1 KiB (2-bit).
 */
#ifndef NTL_DEV_H
#define NTL_DEV_H

#ifdef __CUDACC_RTC__
typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#else
#include <stdint.h>
#endif

#define NTL_DEV_MAX_PAT   16
#define NTL_DEV_MAX_LEN   18
#define NTL_CHUNK_BITS    4096
#define NTL_LANE_BITS     128

/* One unique pattern, pre-digested on the host. */
typedef struct {
    int32_t  m;                        /* length                                                                  */
    int32_t  fixed;                    /* !grepl("[WSMKRYBDHVN]") (NanoTel.R:334): 1 = byte equality              */
    uint32_t mux2[NTL_DEV_MAX_LEN][4]; /* main scan, 2-bit reads: all-ones/zero word per 2-bit code c: letter j
                                          accepts code c under this pattern's own fixed flag                       */
    uint32_t mux4[NTL_DEV_MAX_LEN][4]; /* all-ones/zero word per nibble bit A, C, G, T of letter j                */
    uint32_t q4[4];                    /* bit j set <=> pattern letter j has the A / C / G / T bit (nibble planes) */
    uint8_t  nib[NTL_DEV_MAX_LEN];     /* Biostrings nibble of letter j                                           */
    uint8_t  pad[2];
} ntl_dev_pat;

typedef struct {
    int32_t n_main;                    /* unique --patterns, sorted by length (groups of equal length contiguous) */
    int32_t n_tvr;                     /* unique --tvr_patterns, sorted by length                                 */
    int32_t n_tracks;                  /* 2, or 3 with TVR                                                        */
    int32_t raw_hits_A;                /* track A keeps raw hits: scalar, fixed pattern (NanoTel.R:349-354)       */
    int32_t S;                         /* subseq_length                                                           */
    int32_t right_edge;
    int32_t use_filter;
    int32_t debug_stages;
    int32_t n_main_groups;             /* runs of equal pattern length among main_pat[]                           */
    int32_t n_tvr_groups;
    int32_t thr_reg;                   /* smallest covered count that makes a width-S window telomeric            */
    int32_t pad0;
    int32_t main_group_begin[NTL_DEV_MAX_PAT + 1];
    int32_t tvr_group_begin[NTL_DEV_MAX_PAT + 1];
    double  min_density;
    double  filter_threshold;          /* min_density * 0.8 (NanoTel.R:2143), multiplied in double on the host    */
    ntl_dev_pat main_pat[NTL_DEV_MAX_PAT];
    ntl_dev_pat tvr_pat[NTL_DEV_MAX_PAT];
} ntl_dev_params;

/* Arguments of the scan kernel (K2). */
typedef struct {
    const uint32_t *packed;            /* all reads, words                                                        */
    const int32_t  *len;               /* [n_reads]                                                               */
    const int64_t  *woff;              /* [n_reads] first word of the read                                        */
    const int64_t  *win_off;           /* [n_reads] first window of the read                                      */
    const uint8_t  *pass;              /* [n_reads] edge-filter verdict, or NULL                                  */
    const int32_t  *order;             /* [n_items] read indices, longest first                                   */
    int32_t         n_items;
    uint32_t       *counter;           /* work counter, zeroed before launch                                      */
    uint16_t       *cum[3];            /* per track: covered bases in [1, window end], mod 2^16                   */
} ntl_scan_args;

/* Arguments of the filter (K4) and locate (K3) kernels. */
typedef struct {
    const uint32_t *packed;
    const int32_t  *len;
    const int64_t  *woff;
    const int64_t  *win_off;
    const uint8_t  *fmt;               /* [n_reads] 0 = 2-bit, 1 = 4-bit                                          */
    uint8_t        *pass;              /* [n_reads] written by K4, read by K3; NULL when the filter is off        */
    const uint16_t *cum[3];
    const uint16_t *thr;               /* [2 S + 2] smallest covered count that makes a window of that width
                                          telomeric: !(count / width < min_density), NanoTel.R:751-758         */
    const double   *dens;              /* [S + 1] dens[c] = (double)c / (double)S (NanoTel.R:467 for width-S windows) */
    const int32_t  *order;             /* [n_reads] read indices, longest first (triage walks reads in this order)   */
    int32_t        *cand;              /* [n_reads] reads the triage kernel hands on to the locate kernel            */
    int32_t        *cand_state;        /* [n_reads][4] per candidate: tracks done, max interval width, error, pad     */
    uint32_t       *counters;          /* [0] number of entries in cand[], [1] locate work counter; zeroed per pass   */
    void           *results;           /* ntl_read_result[n_reads]                                                */
    void           *stages;            /* ntl_stage[n_reads][3] or NULL                                           */
    int32_t         n_reads;
} ntl_read_args;

#endif
