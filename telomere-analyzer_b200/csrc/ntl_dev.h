/*
 * ntl_dev.h -- parameter blocks shared by the host side and the sm_100a kernels of libnanotel_b200.
 *
 * Packed read layout in HBM (our choice; nothing like it exists in the reference):
 *   position p of a read (1-based, after --rc) is bit (p - 1) of the read's bit stream, cut into 32-bit "position
 *   words" (LSB first).  Word w of a read is stored as one record of planes:
 *       2-bit reads (ACGT only):  { lo, hi }          8 bytes   code = (ASCII >> 1) & 3 : A 0, C 1, T 2, G 3
 *       4-bit reads (IUPAC):      { A, C, G, T }     16 bytes   Biostrings code bits
 *   Reads are cut into SPANS of W position words (32 W positions).  W is chosen from --subseq_length so that a span
 *   holds a whole number of count blocks of SG positions (SG divides subseq_length; ntl_geometry): every block of
 *   every read then lies inside ONE span at a position known at compile time, and a span is an independent unit of
 *   work for the scan kernel -- one thread, no exchange with its neighbours beyond reading one word on either side.
 *   A read occupies ceil(L / 32 W) spans (zero padded), starts at a span index that is a multiple of span_align, and
 *   the 2-bit and 4-bit reads live in two arenas of such spans.
 *   The scan kernel writes, per track, one uint16 per block: the covered bases of [1 + j SG, (j + 1) SG] (clipped to
 *   the read).  Block j of span s lives at entry (s * blocks_per_span + j) of the track's plane, so a read's blocks
 *   are contiguous from cnt_off[r] = first_span[r] * blocks_per_span on.  A window of split_telo (NanoTel.R:199-227)
 *   is Q = subseq_length / SG consecutive blocks; the last window takes every remaining block.
 *   CLASS BITS.  Beside the counts every track has a plane of class bits: one bit per block, set iff the block's count
 *   reaches blk_thr = ceil(thr_reg / Q) -- for Q = 1 exactly "this (regular) window is telomeric" (NanoTel.R:751-758),
 *   for Q > 1 a necessary condition for the window that holds the block.  The triage and locate kernels find telomeric
 *   windows by scanning these bits (1/16 of the bytes of the counts).  Bit address of block g of the plane:
 *       (g / cls_bps) * 8 + g % cls_bps
 *   cls_bps = blocks per span when a span holds at most 8 blocks (the scan kernel stores ONE byte per span and
 *   track), else 8 (dense bits, written from the counts by ntl_cls_kernel).  cnt_off[r] is a multiple of 8 and of the
 *   blocks per span, so a read's bits start on a byte.  Bits outside a scanned read's blocks are undefined.
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
#define NTL_ITEM_SPANS    32          /* spans per work item of the scan kernel: one per lane                       */

/* span flags (one byte per span, built by the host packer; the edge filter rewrites SKIP) */
#define NTL_SPAN_FIRST    1           /* first span of its read: nothing of the previous word belongs to the read    */
#define NTL_SPAN_TAIL     2           /* handled by the validity-aware tail pass (last span of a read, and the one
                                         before it when the last holds fewer than NTL_DEV_MAX_LEN positions)        */
#define NTL_SPAN_SKIP     4           /* padding between reads, or a read the edge filter dropped                    */

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
    int32_t SG;                        /* positions per count block (divides S)                                   */
    int32_t Q;                         /* blocks per regular window = S / SG                                      */
    int32_t W;                         /* position words per span                                                 */
    int32_t BPS;                       /* blocks per span = 32 W / SG (0: generic geometry, no spans)             */
    int32_t cls_bps;                   /* class-bit layout: blocks per byte (see above)                           */
    int32_t main_group_begin[NTL_DEV_MAX_PAT + 1];
    int32_t tvr_group_begin[NTL_DEV_MAX_PAT + 1];
    double  min_density;
    double  filter_threshold;          /* min_density * 0.8 (NanoTel.R:2143), multiplied in double on the host    */
    ntl_dev_pat main_pat[NTL_DEV_MAX_PAT];
    ntl_dev_pat tvr_pat[NTL_DEV_MAX_PAT];
} ntl_dev_params;

/* Arguments of the span scan kernel (K2), one launch per arena. */
typedef struct {
    const uint32_t *arena;             /* first position word of span 0 of this arena (16 readable bytes before it) */
    const uint8_t  *flags;             /* [n_spans rounded up to 32] NTL_SPAN_*                                   */
    const int32_t  *items;             /* work items (groups of 32 spans) to scan, or NULL = all of them          */
    const uint32_t *n_items_dev;       /* number of entries of items[] (device memory), or NULL                   */
    int32_t         n_items;           /* ceil(n_spans / 32) when items == NULL                                   */
    uint32_t       *work_counter;      /* zeroed before the launch: hands out the tail units, then the items       */
    int32_t         n_reads;           /* tail pass: reads of the whole batch ...                                 */
    const int32_t  *len;               /* [n_reads]                                                               */
    const int64_t  *woff;              /* [n_reads] first position word of the read inside ITS arena              */
    const uint8_t  *fmt;               /* [n_reads] 0 = 2-bit, 1 = 4-bit: the launch handles the reads of its arena */
    const uint8_t  *pass;              /* [n_reads] edge-filter verdict, or NULL                                  */
    int64_t         cnt_base;          /* entry of span 0 of this arena inside a track's block-count plane        */
    uint16_t       *cnt[3];            /* per track: covered bases per block                                      */
    uint8_t        *cls[3];            /* per track: class byte of every span (blocks per span <= 8), or NULL       */
    int64_t         cls_base;          /* entry of span 0 of this arena inside a track's class plane              */
    int32_t         blk_thr;           /* a block's class bit: count >= blk_thr                                    */
    int32_t         pad;               /* always 0: the kernel derives a multiplier of 1 from it that the compiler
                                          cannot fold (ntl_mad1: additions moved to the FMA pipe)                    */
} ntl_scan_args;

/* Arguments of the filter (K4), generic scan and locate (K3) kernels. */
typedef struct {
    const uint32_t *arena2;            /* 2-bit arena: records {lo, hi}                                           */
    const uint32_t *arena4;            /* 4-bit arena: records {A, C, G, T}                                       */
    const int32_t  *len;
    const int64_t  *woff;              /* first position word of the read inside its arena                        */
    const int64_t  *cnt_off;           /* first block of the read inside a track's block-count plane (multiple of 8) */
    const uint8_t  *fmt;               /* [n_reads] 0 = 2-bit, 1 = 4-bit                                          */
    uint8_t        *pass;              /* [n_reads] written by K4, read by K2/K3; NULL when the filter is off     */
    uint16_t       *cnt[3];            /* per track: covered bases per block                                      */
    uint8_t        *cls[3];            /* per track: class bits of the blocks (layout: cls_bps)                    */
    const uint16_t *thr;               /* [2 S + 2] smallest covered count that makes a window of that width
                                          telomeric: !(count / width < min_density), NanoTel.R:751-758         */
    const double   *dens;              /* [S + 1] dens[c] = (double)c / (double)S (NanoTel.R:467 for width-S windows) */
    const int32_t  *order;             /* [n_reads] read indices, longest first (triage walks reads in this order)   */
    int32_t        *cand;              /* [n_reads] reads the triage kernel hands on to the locate kernel            */
    int32_t        *cand_state;        /* [n_reads][4] per candidate: tracks done, width of track 0, 1, 2             */
    uint32_t       *counters;          /* [0] entries in cand[], [1] locate work counter, [2], [3] entries of the item
                                          lists of the two arenas, [4] generic scan, [5], [6] span scan work counters  */
    void           *results;           /* ntl_read_result[n_reads]                                                */
    void           *stages;            /* ntl_stage[n_reads][3] or NULL                                           */
    uint8_t        *span_flags[2];     /* per arena: NTL_SPAN_* (the filter rewrites SKIP), or NULL                */
    uint8_t        *item_active[2];    /* per arena: [n_items] set by the filter for items holding a passing read  */
    int32_t         n_reads;
} ntl_read_args;

#endif
