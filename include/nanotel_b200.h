/*
 * nanotel_b200.h -- C ABI of libnanotel_b200.so, the B200-native (sm_100a) replacement for NanoTel's per-read
 * telomere-detection hot path.
 *
 * The reference (Tzfatilab/Telomere-Analyzer, NanoTel.R) has NO native/FFI boundary: it is one R script.  The seam
 * this library replaces is the body of the chunk loop of run_future_worker_chuncks():
 *
 *     NanoTel.R:2219-2221   reverseComplement(dna_reads)            (--rc)            -> ntl_params.rc
 *     NanoTel.R:2227-2232   filter_reads(...)                       (--use_filter)    -> ntl_params.use_filter
 *     NanoTel.R:2234-2254   8 x search_patterns(...) futures        (analyze_read)    -> ntl_scan_batch()
 *
 * One ntl_scan_batch() call takes the reads of one --nrec chunk (what readDNAStringSet returned at
 * NanoTel.R:2213, as plain ASCII) and returns, for every read in input order, what analyze_read()
 * (NanoTel.R:1774-1976) computes before it starts writing files: the keep decision (:1847-1868), and per track
 * (exact / 1-mismatch / 1-mismatch+TVR) the telomere start, end and density (:1840-1844, :1923-1961), plus the
 * per-window density tables that analyze_subtelos() (:717-766) hands to the plot functions (:1876-1918).
 * Serial numbering, CSV/FASTA/plot writing stay on the host side (R, or the Python mirror in
 * telomere-analyzer_b200/nanotel_b200).
 *
 * Plain C, no R or torch types: loadable with dyn.load/.Call glue (telomere-analyzer_b200/R/r_shim.c),
 * Python ctypes, or dlopen.  All functions return 0 (NTL_OK) or a negative ntl_status; the message is available
 * from ntl_last_error().  No exceptions or abort() cross the boundary; CUDA errors are translated.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with NTL_ERR_CUDA.
 *
 * Threading: a context is owned by one host thread at a time (R's main thread); calls block until the batch is
 * done.  Never call from a forked child (future multicore / mclapply): CUDA contexts do not survive fork().
 */
#ifndef NANOTEL_B200_H
#define NANOTEL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTL_VERSION 200          /* 0.2.0 */
#define NTL_MAX_PATTERNS 16      /* per list (--patterns, --tvr_patterns)                         */
#define NTL_MAX_PATLEN   18      /* NanoTel.R:589,647: assert(str_length(pattern) <= subseq_width) */
#define NTL_MAX_DEVICES  16      /* GPUs of one box a context may shard its batches over            */
#define NTL_MAX_SUBSEQ   43690   /* --subseq_length: the merged last window (up to S + ceil(S/2) - 1 wide) must fit 16 bits */

typedef enum {
    NTL_OK = 0,
    NTL_ERR_ARG = -1,            /* NULL pointer, bad count, subseq_length out of [1, NTL_MAX_SUBSEQ] */
    NTL_ERR_PATTERN = -2,        /* empty pattern, > 18 nt, letter outside the IUPAC DNA alphabet  */
    NTL_ERR_SEQUENCE = -3,       /* a letter outside the Biostrings DNA alphabet, read > 2^30 bases */
    NTL_ERR_CUDA = -4,           /* no device / CUDA runtime error (message has the CUDA string)   */
    NTL_ERR_NOMEM = -5,
    NTL_ERR_JIT = -6,            /* no specialised scan kernel (precompiled / cached / NVRTC) and NTL_OPT_REQUIRE_JIT was set */
    NTL_ERR_STATE = -7,          /* call order violated (e.g. ntl_batch_run before ntl_batch_pack) */
    NTL_ERR_IO = -8              /* an output file could not be written (ntl_write_*)              */
} ntl_status;

/* ntl_params.options */
#define NTL_OPT_NO_JIT        1u  /* use the generic (runtime-pattern) scan kernel only              */
#define NTL_OPT_REQUIRE_JIT   2u  /* fail ntl_create if no specialised scan kernel can be had        */
#define NTL_OPT_DEBUG_STAGES  4u  /* also record the intermediate intervals (ntl_get_stages)         */

/* ntl_scan_path(): where the scan kernel of a context comes from */
#define NTL_SCAN_GENERIC      0   /* generic kernel: any patterns / subseq_length, several times slower (see ntl_scan_path_note) */
#define NTL_SCAN_PRECOMPILED  1   /* specialised, cubin built ahead of time next to the library (no NVRTC needed) */
#define NTL_SCAN_CACHED       2   /* specialised, from the per-user cubin cache                                    */
#define NTL_SCAN_NVRTC        3   /* specialised, compiled by NVRTC in ntl_create                                  */

typedef struct {
    int32_t n_patterns;                 /* --patterns tokens in CLI order (NanoTel.R:2322-2326); 1 token = scalar */
    const char *const *patterns;        /* ASCII IUPAC, <= 18 nt each                                              */
    int32_t n_tvr;                      /* --tvr_patterns tokens (0 = NULL, NanoTel.R:2328-2334)                   */
    const char *const *tvr_patterns;
    double  min_density;                /* --min_density   (NanoTel.R:2337)                                        */
    int32_t subseq_length;              /* --subseq_length (NanoTel.R:2338), 1..NTL_MAX_SUBSEQ                     */
    int32_t rc;                         /* --rc: scan the reverse complement of every read (NanoTel.R:2219-2221)   */
    int32_t use_filter;                 /* --use_filter (NanoTel.R:2227-2232)                                      */
    int32_t right_edge;                 /* --check_right_edge (NanoTel.R:2394 right_edge =)                        */
    int32_t device;                     /* CUDA device ordinal (used when n_devices == 0)                          */
    uint32_t options;                   /* NTL_OPT_*                                                               */
    int32_t host_threads;               /* packer threads in total; 0 = number of online CPUs                      */
    int32_t n_devices;                  /* > 0: shard every batch over device_ids[0 .. n_devices) -- what replaces the
                                           8 forked workers of NanoTel.R:2207, :2234-2254: contiguous shards balanced by
                                           bases, one per GPU, records gathered in input order; no collective         */
    const int32_t *device_ids;
} ntl_params;

/* ntl_read_result.status bits */
#define NTL_READ_KEEP        1   /* a summary row is emitted for this read (max interval width >= 30, :1847-1868) */
#define NTL_READ_FILTERED    2   /* dropped by --use_filter (NanoTel.R:2123-2163); tracks are not computed        */
#define NTL_READ_REF_ERROR   4   /* NanoTel.R itself would stop() on this read (see DESIGN.md "degenerate inputs") */
#define NTL_READ_NO_WINDOWS  8   /* split_telo returned an empty table (length - 1 < subseq_length / 2)           */
#define NTL_READ_IUPAC       16  /* read holds non-ACGT letters and went through the 4-bit path                   */

typedef struct {
    int32_t start;      /* Telomere_start* ; -1 => the row prints NA for this track (NanoTel.R:1926-1961) */
    int32_t end;        /* Telomere_end*   ; length = end - start + 1                                      */
    double  density;    /* telo_density*   = covered bases of [start, end] / (end - start + 1)             */
} ntl_track;

typedef struct {        /* 64 bytes, one per read, input order */
    int32_t status;     /* NTL_READ_* bits */
    int32_t n_win;      /* rows of the window table (split_telo, NanoTel.R:199-227) */
    ntl_track track[3]; /* 0: exact, 1: one mismatch, 2: one mismatch + TVR (valid only if n_tvr > 0) */
    int64_t win_offset; /* first count block of this read inside its device's count planes (internal) */
} ntl_read_result;

typedef struct {        /* NTL_OPT_DEBUG_STAGES: one per read and track, mirrors the oracle's stages */
    int32_t coarse_start, coarse_end;   /* find_telo_position after the optional re-run (NanoTel.R:1084-1110) */
    int32_t acc_start, acc_end;         /* after get_accurate_start/end (NanoTel.R:1119-1126)                  */
    int32_t edge_start, edge_end;       /* after the < 100 bp edge fallback (NanoTel.R:1129-1136)              */
    double  acc_density;                /* density of the acc interval (what the 2023 golden summary.csv holds) */
} ntl_stage;

typedef struct {        /* wall/device times of the last batch, milliseconds */
    double pack_ms;         /* host: ASCII -> planar 2-bit/4-bit in pinned memory (wall clock)              */
    double h2d_ms;          /* device: packed reads + tables upload (CUDA events)                            */
    double filter_ms;       /* device: edge-filter kernel                                                    */
    double scan_ms;         /* device: match + coverage + window-prefix kernel(s)  (the dominant kernel)     */
    double locate_ms;       /* device: triage + per-read locator / refinement kernels                        */
    double triage_ms;       /* device: the triage kernel alone (part of locate_ms)                           */
    double d2h_ms;          /* device: results + window prefixes download                                    */
    double total_ms;        /* host wall clock of the whole ntl_scan_batch call                              */
    int64_t bases;          /* bases in the batch                                                            */
    int64_t packed_bytes;   /* bytes of packed reads resident on the device                                  */
    int64_t window_bytes;   /* bytes of per-block covered counts written by the scan kernel                  */
    int64_t h2d_bytes, d2h_bytes;
    int32_t kernel_launches;/* kernels launched by the passes the last ntl_batch_wait covered                */
    int32_t scan_is_jit;    /* 1 if the specialised span scan kernel ran (0: the generic kernel)             */
    int32_t steps;          /* passes covered by filter_ms / scan_ms / locate_ms (sums over those passes)    */
    int32_t candidates;     /* reads of the last pass that needed the full locate kernel (the rest ended in triage) */
} ntl_timings;

typedef struct ntl_ctx ntl_ctx;

/* -- lifecycle --------------------------------------------------------------------------------------------- */
int  ntl_version(void);
/* Replaces the argument handling of search_patterns()/filter_reads() (NanoTel.R:2001-2002, 2123). */
int  ntl_create(ntl_ctx **ctx, const ntl_params *params);
void ntl_destroy(ntl_ctx *ctx);
/* ctx may be NULL: returns the message of the last failed ntl_create on this thread. */
const char *ntl_last_error(const ntl_ctx *ctx);
/* Which scan kernel the context runs (NTL_SCAN_*), and why when it is the generic one ("" otherwise).  A context
 * that had to fall back to the generic kernel also prints one warning line on stderr in ntl_create. */
int  ntl_scan_path(const ntl_ctx *ctx);
const char *ntl_scan_path_note(const ntl_ctx *ctx);
int  ntl_device_count(const ntl_ctx *ctx);
/* Shards of the last batch: bounds[0 .. n + 1) (read indices) and devices[0 .. n); returns the number of shards. */
int  ntl_get_shards(const ntl_ctx *ctx, int32_t *bounds, int32_t *devices, int32_t cap);
/* The span geometry chosen for --subseq_length (telomere-analyzer_b200/csrc/ntl_dev.h): positions per count block,
 * blocks per window, position words per span, blocks per span (0 = generic layout). */
int  ntl_get_geometry(const ntl_ctx *ctx, int32_t *block, int32_t *blocks_per_window, int32_t *words_per_span,
                      int32_t *blocks_per_span);

/* -- one --nrec chunk, host buffers in, host results out (replaces NanoTel.R:2219-2254) --------------------- */
/* seq[i] points at len[i] ASCII letters (no terminator needed).  *results stays valid until the next batch call
 * or ntl_destroy().  Equivalent to ntl_batch_pack + ntl_batch_upload + ntl_batch_run + ntl_batch_download. */
int ntl_scan_batch(ntl_ctx *ctx, const char *const *seq, const int64_t *len, int32_t n_reads,
                   const ntl_read_result **results);
/* Same, reads given as one concatenated buffer: read i = buf[offsets[i] .. offsets[i+1]). */
int ntl_scan_batch_concat(ntl_ctx *ctx, const char *buf, const int64_t *offsets, int32_t n_reads,
                          const ntl_read_result **results);
/* Same, reads given the way an XStringSet holds them (the dna_reads of NanoTel.R:2213, no as.character() copy): one
 * pool of bytes + 1-based start + width per read.  biostrings_codes != 0: the pool holds Biostrings' DNA byte codes
 * (A 1, C 2, G 4, T 8, IUPAC letters = OR of those bits, '-' 16, '+' 32, '.' 64), else ASCII. */
int ntl_scan_batch_pool(ntl_ctx *ctx, const unsigned char *pool, const int32_t *start, const int32_t *width,
                        int32_t n_reads, int32_t biostrings_codes, const ntl_read_result **results);

/* -- the same path in stages (used by bench.py to time the device-resident part; same results) ------------- */
int ntl_batch_pack(ntl_ctx *ctx, const char *const *seq, const int64_t *len, int32_t n_reads);
int ntl_batch_upload(ntl_ctx *ctx);       /* pinned host -> HBM (async on the context stream, then sync)  */
int ntl_batch_run(ntl_ctx *ctx);          /* filter + scan + locate kernels on the resident batch, sync    */
int ntl_batch_enqueue(ntl_ctx *ctx);      /* the same pass, enqueued on the context stream without waiting  */
int ntl_batch_wait(ntl_ctx *ctx);         /* wait for the enqueued passes (<= 256); timings = sums over them */
/* HBM -> pinned host: the 64-byte records of all reads, then the window tables of the reads with NTL_READ_KEEP (the
 * only ones analyze_read plots, NanoTel.R:1876-1918), gathered on the device into one block.                      */
int ntl_batch_download(ntl_ctx *ctx, const ntl_read_result **results);
int ntl_get_timings(const ntl_ctx *ctx, ntl_timings *out);
void *ntl_stream(const ntl_ctx *ctx);     /* cudaStream_t the kernels are launched on */

/* -- per-window tables of the last batch (the data.frames analyze_subtelos returns, NanoTel.R:740-765) ------- */
/* Fills up to cap rows for read read_idx, track (0..2): start_index, end_index, covered bases and
 * density = covered / width (the plot vectors).  Any output pointer may be NULL.  Returns n_win or < 0.
 * Kept reads are served from the host copy; the table of any other read is still on the device and is fetched by
 * this call (one small synchronous copy), until the next batch replaces it. */
int ntl_get_windows(const ntl_ctx *ctx, int32_t read_idx, int32_t track, int32_t cap,
                    int32_t *start_index, int32_t *end_index, int32_t *covered, double *density);
/* Bulk form for whole-batch comparisons (parity checks at BASELINE sizes): the covered-base counts of `track` for
 * EVERY read of the last batch, read after read in input order, n_win entries per read (results[i].n_win; window
 * order; no padding), fetched from the device in one copy.  Entries of filtered reads are undefined.  Returns the
 * total number of entries (the sum of n_win) or a negative ntl_status; writes nothing beyond cap entries. */
int64_t ntl_get_window_counts(const ntl_ctx *ctx, int32_t track, uint16_t *out, int64_t cap);
/* NTL_OPT_DEBUG_STAGES only: intermediate intervals of read read_idx, track. */
int ntl_get_stages(const ntl_ctx *ctx, int32_t read_idx, int32_t track, ntl_stage *out);

/* -- record reader (host side; replaces open_input_files + readDNAStringSet(files, nrec, format), NanoTel.R:2180, 2213)
 * FASTA (multi-line) / FASTQ (4-line records), gzip transparent, records streamed nrec at a time across the file
 * list, names = full header line without '>' / '@', qualities skipped.  A chunk is ONE contiguous sequence buffer
 * + n+1 offsets (what ntl_scan_batch_concat takes) and one name buffer + n+1 offsets; both stay valid until the next
 * ntl_reader_next / ntl_reader_close.  The following chunk is assembled by a background thread meanwhile, and every
 * file of the list is inflated and parsed by a thread of its own, several files ahead of the one being handed out. */
typedef struct ntl_reader ntl_reader;
int  ntl_reader_open(ntl_reader **reader, const char *const *paths, int32_t n_paths, const char *format);
/* Returns the number of records of the chunk (0 = end of input) or a negative ntl_status. nrec <= 0: everything. */
int32_t ntl_reader_next(ntl_reader *reader, int32_t nrec, const char **seq_buf, const int64_t **seq_off,
                        const char **name_buf, const int64_t **name_off);
const char *ntl_reader_error(const ntl_reader *reader);
void ntl_reader_close(ntl_reader *reader);

/* -- per-read output files (host side; replaces writeXStringSet + the tables handed to the plot functions,
 *    NanoTel.R:1870-1918) ---------------------------------------------------------------------------------- */
/* For every summary row j of the last batch (read order[j], Serial serial[order[j]]; ntl_assign_serials) write
 *   <out_dir>/reads/<Serial>.fasta.gz            '>' + header line + the upper-case sequence, 80 letters per line
 *                                                (reverse-complemented back if rc_applied), gzip level 6
 *   <out_dir>/density_vectors/read<Serial>.csv   ID,start_index,end_index and density / class per track
 * with `threads` host threads.  buf/offsets: the chunk's sequences as given to ntl_scan_batch_concat; names /
 * name_off: the header lines, concatenated (n_reads + 1 offsets).  Returns NTL_OK or NTL_ERR_IO / NTL_ERR_ARG. */
int ntl_write_read_outputs(const ntl_ctx *ctx, const char *out_dir, const char *buf, const int64_t *offsets,
                           const char *names, const int64_t *name_off, const int32_t *serial, const int32_t *order,
                           int32_t n_rows, int32_t n_tracks, double min_density, int32_t rc_applied, int32_t threads);
/* One reads/<Serial>.fasta.gz as above (no context, no device). */
int ntl_write_fasta_gz(const char *path, const char *name, const char *seq, int64_t len, int32_t rc);

/* -- diagnostics ------------------------------------------------------------------------------------------ */
/* NVRTC-compile the pattern-specialised scan kernel for `arch` ("sm_100a") without touching a device; optionally
 * write the cubin to cubin_path (for cuobjdump).  Returns the cubin size in bytes or a negative ntl_status. */
long ntl_jit_compile_check(const ntl_params *params, const char *arch, char *log, int log_cap,
                           const char *cubin_path);
/* The same, stored as <dir>/<key>.cubin in the format ntl_create looks for in <library directory>/precompiled
 * (build.py calls this for the default pattern sets, so that they need neither NVRTC nor a CUDA toolkit at run time). */
long ntl_jit_precompile_to(const ntl_params *params, const char *arch, const char *dir, char *log, int log_cap);
/* The generated source of the specialised kernel (prologue of constants + #include "ntl_scan.cuh"); returns its
 * length.  Used by the CPU tests, which compile the same text as a host model. */
long ntl_jit_get_source(const ntl_params *params, char *buf, long cap);
/* Pack ONE read exactly as ntl_batch_pack does (no device needed): writes its position words ({lo, hi} per 32
 * positions, or {A, C, G, T} if the read holds a letter other than A/C/G/T: *four_bit = 1) to `words` (capacity in
 * 32-bit words) and returns the number of words written, negative on error.  Layout: csrc/ntl_dev.h. */
long ntl_pack_read(const char *seq, int64_t len, int32_t rc, uint32_t *words, int64_t capacity, int32_t *four_bit);

/* Rate (GB/s) at which `threads` host threads read `bytes` bytes at `buf` (AVX2 loads, 4 KiB software prefetch: the
 * packer's own access pattern without its arithmetic or its stores), best of `reps` passes.  The ceiling of
 * ntl_batch_pack on this host: bench.py prints it beside the packer's rate.  reps < 0: -reps passes that also write
 * a quarter of the bytes with non-temporal stores (the packer's whole memory traffic).  No device needed. */
double ntl_host_read_gbs(const void *buf, int64_t bytes, int32_t threads, int32_t reps);

/* -- host-side helpers of the same path ------------------------------------------------------------------- */
/* Serial numbers and row order of one chunk exactly as search_patterns + the 8-way split assign them
 * (NanoTel.R:2050-2069, 2234-2258).  results: the (post-filter) reads of the chunk; serial[i] = 0 for reads
 * without a row; row_order receives the read indices in summary-row order.  Returns the number of rows. */
int ntl_assign_serials(const ntl_read_result *results, int32_t n_reads, int32_t serial_start,
                       int32_t *serial, int32_t *row_order, int32_t *next_serial_start);
/* split_telo (NanoTel.R:199-227): number of windows of a read of that length. */
int32_t ntl_count_windows(int64_t length, int32_t subseq_length);

#ifdef __cplusplus
}
#endif
#endif /* NANOTEL_B200_H */
