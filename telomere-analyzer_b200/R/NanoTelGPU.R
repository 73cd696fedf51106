# NanoTelGPU.R -- drop-in replacement for the body of NanoTel.R's chunk loop (NanoTel.R:2209-2260).
#
# NOT RUN IN THIS REPOSITORY'S IMAGE (no R here).  Usage on a machine with R + Bioconductor + a B200:
#   R CMD SHLIB r_shim.c -I../../include -L../nanotel_b200 -lnanotel_b200 -o nanotel_r.so
#   source("NanoTel.R" functions you keep: create_dirs, plot_single_telo_with_gray_area, plot_single_telo_with_tvr)
#   source("NanoTelGPU.R"); then call run_gpu_worker_chuncks() where NanoTel.R:2392 calls run_future_worker_chuncks().
# It keeps NanoTel.R's CLI, log, summary CSV, reads_ids.txt, reads/<Serial>.fasta.gz and plots; only the per-read
# detection (reverseComplement + filter_reads + 8 x search_patterns futures) is replaced by one .Call per chunk.

dyn.load("nanotel_r.so")

run_gpu_worker_chuncks <- function(input_path, output_path, format = c("fasta", "fastq"), nrec = 10000,
                                   patterns, do_rc, use_filter = FALSE, right_edge = TRUE, tvr_patterns,
                                   device = 0L) {
  filepath <- if (dir.exists(input_path)) dir(full.names = TRUE, path = input_path, recursive = TRUE,
                                              include.dirs = FALSE) else input_path
  files <- open_input_files(filepath)
  ctx <- .Call("ntl_R_create", unlist(patterns), if (is.null(tvr_patterns)) NULL else unlist(tvr_patterns),
               as.double(global_min_density), as.integer(global_subseq_length), isTRUE(do_rc), isTRUE(use_filter),
               isTRUE(right_edge), as.integer(device))
  on.exit(.Call("ntl_R_destroy", ctx))
  has_tvr <- !is.null(tvr_patterns)
  df_summary <- NULL
  dna_length <- integer(0)
  serial_start <- 1L
  output_reads <- file.path(output_path, "reads")
  repeat {
    dna_reads <- readDNAStringSet(files, nrec = nrec, format = format)      # NanoTel.R:2213
    if (length(dna_reads) == 0L) break
    dna_length <- c(dna_length, width(dna_reads))                            # :2225
    res <- .Call("ntl_R_scan_batch", ctx, as.character(dna_reads))           # replaces :2219-2254
    if (any(res$ref_error)) stop("NanoTel.R would have stopped on read(s): ", paste(which(res$ref_error), collapse = " "))
    ser <- .Call("ntl_R_assign_serials", res$keep, res$filtered, serial_start)   # :2050-2069, :2234-2258
    serial_start <- ser$next_serial_start
    if (do_rc) dna_reads <- reverseComplement(dna_reads)                     # frame of the saved FASTA (:2219-2221)
    for (i in ser$order) {
      serial <- ser$serial[i]
      writeXStringSet(dna_reads[i], file.path(output_reads, paste0(serial, ".fasta.gz")), compress = TRUE)  # :1871-1873
      subs <- as.data.frame(.Call("ntl_R_windows", ctx, i, 1L))
      subs_mm <- as.data.frame(.Call("ntl_R_windows", ctx, i, 2L))
      # plot_single_telo_with_gray_area(...) / plot_single_telo_with_tvr(...) exactly as NanoTel.R:1876-1918,
      # with subs / subs_mm (/ subs_tvr = track 3) and res$start[i], res$end[i], res$start_mismatch[i], ...
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
