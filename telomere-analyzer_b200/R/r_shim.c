/*
 * r_shim.c -- .Call glue between R and libnanotel_b200.so (include/nanotel_b200.h).
 *
 * NOT COMPILED OR TESTED IN THIS REPOSITORY'S IMAGE: R, Rinternals.h and Bioconductor are absent here (SURVEY.md
 * section 0), so this file is the binding a maintainer adds on a machine that has R:
 *
 *     R CMD SHLIB r_shim.c -I../../include -L../nanotel_b200 -lnanotel_b200 -o nanotel_r.so
 *     dyn.load("nanotel_r.so")      # see NanoTelGPU.R
 *
 * It replaces the body of NanoTel.R's chunk loop (NanoTel.R:2219-2258: reverseComplement, filter_reads and the eight
 * search_patterns() futures) with ONE call per --nrec chunk.  Threading: .Call runs on R's main thread and blocks;
 * the library must never be called from a forked future/mclapply child (CUDA contexts do not survive fork), which is
 * why the futures are replaced and not wrapped.  Ownership: R owns the input strings (read-only during the call);
 * the library owns its result buffers until the next batch; everything returned to R is copied into fresh R vectors.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>
#include "nanotel_b200.h"

static void ctx_finalizer(SEXP ptr)
{
    ntl_ctx *c = (ntl_ctx *)R_ExternalPtrAddr(ptr);
    if (c) { ntl_destroy(c); R_ClearExternalPtr(ptr); }
}

static ntl_ctx *get_ctx(SEXP ptr)
{
    ntl_ctx *c = (ntl_ctx *)R_ExternalPtrAddr(ptr);
    if (!c) Rf_error("nanotel_b200: context was destroyed");
    return c;
}

/* ntl_R_create(patterns chr, tvr_patterns chr or NULL, min_density dbl, subseq_length int, rc lgl, use_filter lgl,
 *              right_edge lgl, devices int vector) -> external pointer
 * patterns / tvr_patterns: the character vectors cur_patterns / cur_tvr_patterns of NanoTel.R:2322-2334 (unlist()ed).
 * devices: CUDA ordinals; with more than one, every chunk is sharded over them (contiguous shards balanced by bases,
 * records gathered in input order) -- the replacement for plan(multicore, workers = 8) (NanoTel.R:2207). */
SEXP ntl_R_create(SEXP patterns, SEXP tvr, SEXP min_density, SEXP subseq_length, SEXP rc, SEXP use_filter,
                  SEXP right_edge, SEXP devices)
{
    int32_t ids[NTL_MAX_DEVICES];
    const int nd = LENGTH(devices);
    if (nd < 1 || nd > NTL_MAX_DEVICES) Rf_error("nanotel_b200: 1..%d devices", NTL_MAX_DEVICES);
    for (int i = 0; i < nd; i++) ids[i] = INTEGER(devices)[i];
    const char *pp[NTL_MAX_PATTERNS], *tp[NTL_MAX_PATTERNS];
    int np = LENGTH(patterns), nt = Rf_isNull(tvr) ? 0 : LENGTH(tvr);
    if (np < 1 || np > NTL_MAX_PATTERNS || nt > NTL_MAX_PATTERNS) Rf_error("nanotel_b200: 1..%d patterns", NTL_MAX_PATTERNS);
    for (int i = 0; i < np; i++) pp[i] = CHAR(STRING_ELT(patterns, i));
    for (int i = 0; i < nt; i++) tp[i] = CHAR(STRING_ELT(tvr, i));
    ntl_params p;
    memset(&p, 0, sizeof p);
    p.n_patterns = np; p.patterns = pp;
    p.n_tvr = nt; p.tvr_patterns = nt ? tp : NULL;
    p.min_density = Rf_asReal(min_density);
    p.subseq_length = Rf_asInteger(subseq_length);
    p.rc = Rf_asLogical(rc) == TRUE;
    p.use_filter = Rf_asLogical(use_filter) == TRUE;
    p.right_edge = Rf_asLogical(right_edge) == TRUE;
    p.device = ids[0];
    p.n_devices = nd; p.device_ids = ids;
    ntl_ctx *c = NULL;
    int st = ntl_create(&c, &p);
    if (st != NTL_OK) Rf_error("nanotel_b200: ntl_create failed (%d): %s", st, ntl_last_error(NULL));
    if (ntl_scan_path(c) == NTL_SCAN_GENERIC)           /* never a silent slow path */
        Rf_warning("nanotel_b200: %s", ntl_scan_path_note(c));
    SEXP ptr = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* ntl_R_scan_batch(ctx, seqs chr)  with seqs = as.character(dna_reads) of one chunk (NOT reverse-complemented: the
 * library applies --rc itself).  Returns a named list of vectors of length(seqs):
 *   keep lgl, filtered lgl, ref_error lgl, n_win int,
 *   start / end (int, NA where the reference prints NA) and density (dbl) for tracks "", "_mismatch", "_mismatch_tvr". */
static SEXP results_to_list(const ntl_read_result *res, int n);

SEXP ntl_R_scan_batch(SEXP ctxp, SEXP seqs)
{
    ntl_ctx *c = get_ctx(ctxp);
    const int n = LENGTH(seqs);
    const char **sp = (const char **)R_alloc((size_t)n + 1, sizeof(char *));
    int64_t *len = (int64_t *)R_alloc((size_t)n + 1, sizeof(int64_t));
    for (int i = 0; i < n; i++) { SEXP s = STRING_ELT(seqs, i); sp[i] = CHAR(s); len[i] = (int64_t)LENGTH(s); }
    const ntl_read_result *res = NULL;
    int st = ntl_scan_batch(c, sp, len, n, &res);
    if (st != NTL_OK) Rf_error("nanotel_b200: ntl_scan_batch failed (%d): %s", st, ntl_last_error(c));
    return results_to_list(res, n);
}

/* ntl_R_scan_xstringset(ctx, dna_reads): the DNAStringSet of NanoTel.R:2213 handed over WITHOUT as.character(): an
 * XStringSet is a pool of shared raw vectors (x@pool@xp_list[[g]], an external pointer whose tag is the RAWSXP) plus
 * x@ranges@group / @start / @width; the bytes are Biostrings' DNA codes, which the library decodes itself
 * (ntl_scan_batch_pool).  Returns NULL if the reads do not all live in one pool element (the caller then falls back
 * to ntl_R_scan_batch(as.character(dna_reads))); readDNAStringSet() always yields a single one. */
SEXP ntl_R_scan_xstringset(SEXP ctxp, SEXP x)
{
    ntl_ctx *c = get_ctx(ctxp);
    SEXP ranges = R_do_slot(x, Rf_install("ranges"));
    SEXP group = R_do_slot(ranges, Rf_install("group"));
    SEXP start = R_do_slot(ranges, Rf_install("start"));
    SEXP width = R_do_slot(ranges, Rf_install("width"));
    const int n = LENGTH(start);
    if (n == 0) return results_to_list(NULL, 0);
    const int g0 = INTEGER(group)[0];
    for (int i = 1; i < n; i++) if (INTEGER(group)[i] != g0) return R_NilValue;
    SEXP xp_list = R_do_slot(R_do_slot(x, Rf_install("pool")), Rf_install("xp_list"));
    SEXP raw = R_ExternalPtrTag(VECTOR_ELT(xp_list, g0 - 1));
    if (TYPEOF(raw) != RAWSXP) return R_NilValue;
    const ntl_read_result *res = NULL;
    int st = ntl_scan_batch_pool(c, RAW(raw), INTEGER(start), INTEGER(width), n, /*biostrings_codes=*/1, &res);
    if (st != NTL_OK) Rf_error("nanotel_b200: ntl_scan_batch_pool failed (%d): %s", st, ntl_last_error(c));
    return results_to_list(res, n);
}

static SEXP results_to_list(const ntl_read_result *res, int n)
{

    static const char *names[] = {"keep", "filtered", "ref_error", "n_win",
                                  "start", "end", "density",
                                  "start_mismatch", "end_mismatch", "density_mismatch",
                                  "start_mismatch_tvr", "end_mismatch_tvr", "density_mismatch_tvr", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SEXP keep = PROTECT(Rf_allocVector(LGLSXP, n)), filt = PROTECT(Rf_allocVector(LGLSXP, n));
    SEXP err = PROTECT(Rf_allocVector(LGLSXP, n)), nwin = PROTECT(Rf_allocVector(INTSXP, n));
    SET_VECTOR_ELT(out, 0, keep); SET_VECTOR_ELT(out, 1, filt); SET_VECTOR_ELT(out, 2, err); SET_VECTOR_ELT(out, 3, nwin);
    for (int t = 0; t < 3; t++) {
        SEXP s = PROTECT(Rf_allocVector(INTSXP, n)), e = PROTECT(Rf_allocVector(INTSXP, n));
        SEXP d = PROTECT(Rf_allocVector(REALSXP, n));
        for (int i = 0; i < n; i++) {
            const ntl_track *tr = &res[i].track[t];
            const int na = tr->start == -1;                  /* NanoTel.R:1926-1961 */
            INTEGER(s)[i] = na ? NA_INTEGER : tr->start;
            INTEGER(e)[i] = na ? NA_INTEGER : tr->end;
            REAL(d)[i] = na ? NA_REAL : tr->density;
        }
        SET_VECTOR_ELT(out, 4 + 3 * t, s); SET_VECTOR_ELT(out, 5 + 3 * t, e); SET_VECTOR_ELT(out, 6 + 3 * t, d);
        UNPROTECT(3);
    }
    for (int i = 0; i < n; i++) {
        LOGICAL(keep)[i] = (res[i].status & NTL_READ_KEEP) != 0;
        LOGICAL(filt)[i] = (res[i].status & NTL_READ_FILTERED) != 0;
        LOGICAL(err)[i] = (res[i].status & NTL_READ_REF_ERROR) != 0;
        INTEGER(nwin)[i] = res[i].n_win;
    }
    UNPROTECT(5);
    return out;
}

/* ntl_R_windows(ctx, read_index (1-based), track (1..3), min_density) -> data.frame-ready list(ID, start_index,
 * end_index, density, class): the `subs` table analyze_subtelos returns (NanoTel.R:740-765; class = -5 CCCTAA /
 * 1 NONE / 0 SKIP by :749-758), consumed unchanged by plot_single_telo_with_*(). */
SEXP ntl_R_windows(SEXP ctxp, SEXP read_index, SEXP track, SEXP min_density)
{
    ntl_ctx *c = get_ctx(ctxp);
    const int i = Rf_asInteger(read_index) - 1, t = Rf_asInteger(track) - 1;
    const double md = Rf_asReal(min_density);
    int n = ntl_get_windows(c, i, t, 0, NULL, NULL, NULL, NULL);
    if (n < 0) Rf_error("nanotel_b200: ntl_get_windows failed (%d): %s", n, ntl_last_error(c));
    static const char *names[] = {"ID", "start_index", "end_index", "density", "class", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SEXP id = PROTECT(Rf_allocVector(INTSXP, n)), st = PROTECT(Rf_allocVector(INTSXP, n));
    SEXP en = PROTECT(Rf_allocVector(INTSXP, n)), de = PROTECT(Rf_allocVector(REALSXP, n));
    SEXP cl = PROTECT(Rf_allocVector(REALSXP, n));
    ntl_get_windows(c, i, t, n, INTEGER(st), INTEGER(en), NULL, REAL(de));
    for (int k = 0; k < n; k++) {
        INTEGER(id)[k] = k + 1;
        const double d = REAL(de)[k];
        REAL(cl)[k] = d < md ? (d < 0.1 ? 0.0 : 1.0) : -5.0;      /* :749-758 */
    }
    SET_VECTOR_ELT(out, 0, id); SET_VECTOR_ELT(out, 1, st); SET_VECTOR_ELT(out, 2, en); SET_VECTOR_ELT(out, 3, de);
    SET_VECTOR_ELT(out, 4, cl);
    UNPROTECT(6);
    return out;
}

/* ntl_R_assign_serials(keep lgl, filtered lgl, serial_start int) -> list(serial int (0 = no row), order int (1-based),
 * next_serial_start int): search_patterns' counter under the 8-way split (NanoTel.R:2050-2069, 2234-2258). */
SEXP ntl_R_assign_serials(SEXP keep, SEXP filtered, SEXP serial_start)
{
    const int n = LENGTH(keep);
    ntl_read_result *tmp = (ntl_read_result *)R_alloc((size_t)n + 1, sizeof(ntl_read_result));
    memset(tmp, 0, ((size_t)n + 1) * sizeof(ntl_read_result));
    for (int i = 0; i < n; i++)
        tmp[i].status = (LOGICAL(keep)[i] == TRUE ? NTL_READ_KEEP : 0) | (LOGICAL(filtered)[i] == TRUE ? NTL_READ_FILTERED : 0);
    SEXP serial = PROTECT(Rf_allocVector(INTSXP, n));
    int32_t *order = (int32_t *)R_alloc((size_t)n + 1, sizeof(int32_t));
    int32_t next = Rf_asInteger(serial_start);
    int rows = ntl_assign_serials(tmp, n, next, INTEGER(serial), order, &next);
    if (rows < 0) Rf_error("nanotel_b200: ntl_assign_serials failed (%d)", rows);
    SEXP ord = PROTECT(Rf_allocVector(INTSXP, rows));
    for (int k = 0; k < rows; k++) INTEGER(ord)[k] = order[k] + 1;
    static const char *names[] = {"serial", "order", "next_serial_start", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, serial); SET_VECTOR_ELT(out, 1, ord); SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(next));
    UNPROTECT(3);
    return out;
}

SEXP ntl_R_destroy(SEXP ctxp) { ctx_finalizer(ctxp); return R_NilValue; }

static const R_CallMethodDef call_methods[] = {
    {"ntl_R_create", (DL_FUNC)&ntl_R_create, 8},
    {"ntl_R_scan_batch", (DL_FUNC)&ntl_R_scan_batch, 2},
    {"ntl_R_scan_xstringset", (DL_FUNC)&ntl_R_scan_xstringset, 2},
    {"ntl_R_windows", (DL_FUNC)&ntl_R_windows, 4},
    {"ntl_R_assign_serials", (DL_FUNC)&ntl_R_assign_serials, 3},
    {"ntl_R_destroy", (DL_FUNC)&ntl_R_destroy, 1},
    {NULL, NULL, 0}};

void R_init_nanotel_r(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
