# NanoTelGPU.R -- drop-in replacement for the body of NanoTel.R's chunk loop (NanoTel.R:2209-2260).
#
# NOT RUN IN THIS REPOSITORY'S IMAGE (no R here).  Usage on a machine with R + Bioconductor + B200(s):
#   R CMD SHLIB r_shim.c -I../../include -L../nanotel_b200 -lnanotel_b200 -o nanotel_r.so
#   In NanoTel.R: source("NanoTelGPU.R") after the function definitions, and call run_gpu_worker_chuncks() where
#   NanoTel.R:2392 calls run_future_worker_chuncks() (same arguments, plus `devices`).
# It keeps NanoTel.R's CLI, log, summary CSV, reads_ids.txt, reads/<Serial>.fasta.gz and the three plots per telomeric
# read; only the per-read detection (reverseComplement + filter_reads + 8 x search_patterns futures) is replaced by one
# .Call per chunk.  plot_single_telo_with_gray_area / plot_single_telo_with_tvr (NanoTel.R:1271-1624) and create_dirs
# (:1979-1996) are the reference's own functions, used unchanged.

dyn.load("nanotel_r.so")

# The three plots analyze_read draws for one telomeric read (NanoTel.R:1876-1918), from the window tables and the
# intervals the library returned.  `iv` = list(start, end, start_mismatch, end_mismatch[, start_mismatch_tvr, ...]);
# a track the reference prints as NA (start == -1, :1926-1961) is drawn as the IRanges(-1, -1) it holds there.
plot_read_gpu <- function(serial, read_length, subs, subs_mm, subs_tvr, iv, max_length, title,
                          output_jpegs, output_jpegs_1) {
  na1 <- function(v) if (is.na(v)) -1L else v
  common <- list(seq_length = read_length, subs = subs, subs_mismatch = subs_mm, serial_num = serial,
                 seq_start = na1(iv$start), seq_end = na1(iv$end), gray_start = na1(iv$start_mismatch),
                 gray_end = na1(iv$end_mismatch), save_it = TRUE, main_title = title, w = 750, h = 300)
  if (is.null(subs_tvr)) {
    do.call(plot_single_telo_with_gray_area, c(common, list(x_length = max_length, output_jpegs = output_jpegs)))      # :1877-1882
    do.call(plot_single_telo_with_gray_area, c(common, list(x_length = read_length, output_jpegs = output_jpegs_1)))   # :1884-1889
    do.call(plot_single_telo_with_gray_area, c(common, list(x_length = read_length, output_jpegs = output_jpegs_1,
                                                             eps = TRUE)))                                             # :1892-1896
  } else {
    tv <- c(common, list(subs_tvr = subs_tvr, tvr_start = na1(iv$start_mismatch_tvr), tvr_end = na1(iv$end_mismatch_tvr)))
    do.call(plot_single_telo_with_tvr, c(tv, list(x_length = max_length, output_jpegs = output_jpegs)))                # :1898-1903
    do.call(plot_single_telo_with_tvr, c(tv, list(x_length = read_length, output_jpegs = output_jpegs_1)))             # :1905-1910
    do.call(plot_single_telo_with_tvr, c(tv, list(x_length = read_length, output_jpegs = output_jpegs_1, eps = TRUE))) # :1913-1917
  }
}

run_gpu_worker_chuncks <- function(input_path, output_path, format = c("fasta", "fastq"), nrec = 10000,
                                   patterns, do_rc, use_filter = FALSE, right_edge = TRUE, tvr_patterns,
                                   devices = 0L, max_length = 1e5, title = "") {
  filepath <- if (dir.exists(input_path)) dir(full.names = TRUE, path = input_path, recursive = TRUE,
                                              include.dirs = FALSE) else input_path
  files <- open_input_files(filepath)
  ctx <- .Call("ntl_R_create", unlist(patterns), if (is.null(tvr_patterns)) NULL else unlist(tvr_patterns),
               as.double(global_min_density), as.integer(global_subseq_length), isTRUE(do_rc), isTRUE(use_filter),
               isTRUE(right_edge), as.integer(devices))
  on.exit(.Call("ntl_R_destroy", ctx))
  has_tvr <- !is.null(tvr_patterns)
  df_summary <- NULL
  dna_length <- integer(0)
  serial_start <- 1L
  output_reads <- file.path(output_path, "reads")                       # create_dirs(), NanoTel.R:1979-1996
  output_jpegs <- file.path(output_path, "single_read_plots")
  output_jpegs_1 <- file.path(output_path, "single_read_plots_adj")
  repeat {
    dna_reads <- readDNAStringSet(files, nrec = nrec, format = format)      # NanoTel.R:2213
    if (length(dna_reads) == 0L) break
    dna_length <- c(dna_length, width(dna_reads))                            # :2225
    # replaces :2219-2254.  The XStringSet's own pool goes to the library (no as.character() copy through R's string
    # cache); the character route is the fallback for a set whose reads do not share one pool element.
    res <- .Call("ntl_R_scan_xstringset", ctx, dna_reads)
    if (is.null(res)) res <- .Call("ntl_R_scan_batch", ctx, as.character(dna_reads))
    if (any(res$ref_error)) stop("NanoTel.R would have stopped on read(s): ", paste(which(res$ref_error), collapse = " "))
    ser <- .Call("ntl_R_assign_serials", res$keep, res$filtered, serial_start)   # :2050-2069, :2234-2258
    serial_start <- ser$next_serial_start
    if (do_rc) dna_reads <- reverseComplement(dna_reads)                     # frame of the saved FASTA (:2219-2221)
    for (i in ser$order) {
      serial <- ser$serial[i]
      writeXStringSet(dna_reads[i], file.path(output_reads, paste0(serial, ".fasta.gz")), compress = TRUE)  # :1871-1873
      md <- as.double(global_min_density)
      subs <- as.data.frame(.Call("ntl_R_windows", ctx, i, 1L, md))          # analyze_list[[1]]  (:1781)
      subs_mm <- as.data.frame(.Call("ntl_R_windows", ctx, i, 2L, md))       # analyze_list2[[1]] (:1792)
      subs_tvr <- if (has_tvr) as.data.frame(.Call("ntl_R_windows", ctx, i, 3L, md)) else NULL   # :1809
      iv <- lapply(res[c("start", "end", "start_mismatch", "end_mismatch", "start_mismatch_tvr", "end_mismatch_tvr")],
                   function(v) v[i])
      plot_read_gpu(serial, width(dna_reads)[i], subs, subs_mm, subs_tvr, iv, max_length, title,
                    output_jpegs, output_jpegs_1)                             # :1876-1918
      row <- data.frame(Serial = serial, sequence_ID = names(dna_reads)[i], sequence_length = width(dna_reads)[i],
                        telo_density = res$density[i], Telomere_start = res$start[i], Telomere_end = res$end[i],
                        Telomere_length = res$end[i] - res$start[i] + 1L,
                        telo_density_mismatch = res$density_mismatch[i],
                        Telomere_start_mismatch = res$start_mismatch[i], Telomere_end_mismatch = res$end_mismatch[i],
                        Telomere_length_mismatch = res$end_mismatch[i] - res$start_mismatch[i] + 1L)
      if (has_tvr) {
        row$telo_density_mismatch_tvr <- res$density_mismatch_tvr[i]
        row$Telomere_start_mismatch_tvr <- res$start_mismatch_tvr[i]
        row$Telomere_end_mismatch_tvr <- res$end_mismatch_tvr[i]
        row$Telomere_length_mismatch_tvr <- res$end_mismatch_tvr[i] - res$start_mismatch_tvr[i] + 1L
      }
      df_summary <- if (is.null(df_summary)) row else dplyr::rows_append(df_summary, row)
    }
  }
  list(df_summary = df_summary, all_reads_length_vec = dna_length)
}
