# plot_density_vectors.R -- the reference's per-read plots from the files the R-free driver wrote.
#
# `python -m nanotel_b200 ...` (telomere-analyzer_b200/nanotel_b200/nanotel.py) writes, next to the reference's own
# outputs, density_vectors/read<Serial>.csv: the window tables analyze_subtelos returns (ID, start_index, end_index,
# density, class per track; NanoTel.R:740-765).  This script turns them into the three plots analyze_read draws for
# every telomeric read (NanoTel.R:1876-1918), with the reference's OWN plot functions:
#
#   Rscript plot_density_vectors.R <path/to/NanoTel.R> <save_path> [max_length = 1e5] [title]
#
# NOT RUN IN THIS REPOSITORY'S IMAGE (no R here).  Only base R + readr are needed besides what NanoTel.R's plot
# functions use (grDevices); the function definitions are taken out of NanoTel.R without running its main part.
args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 2) stop("usage: Rscript plot_density_vectors.R <NanoTel.R> <save_path> [max_length] [title]")
nanotel <- args[1]; save_path <- args[2]
max_length <- if (length(args) >= 3) as.numeric(args[3]) else 1e5
title <- if (length(args) >= 4) args[4] else ""

# source only the two plot functions (NanoTel.R:1271-1624): evaluate the top-level `name <- function` assignments
exprs <- parse(nanotel, keep.source = FALSE)
for (e in exprs) {
  if (is.call(e) && identical(e[[1]], as.name("<-")) && is.name(e[[2]]) &&
      as.character(e[[2]]) %in% c("plot_single_telo_with_gray_area", "plot_single_telo_with_tvr")) eval(e, globalenv())
}

summary_csv <- list.files(save_path, pattern = "_summary\\.csv$", full.names = TRUE)[1]
df <- readr::read_csv(summary_csv, show_col_types = FALSE)
has_tvr <- "Telomere_start_mismatch_tvr" %in% names(df)
output_jpegs <- file.path(save_path, "single_read_plots")
output_jpegs_1 <- file.path(save_path, "single_read_plots_adj")
dir.create(output_jpegs, showWarnings = FALSE); dir.create(output_jpegs_1, showWarnings = FALSE)
na1 <- function(v) if (is.na(v)) -1L else as.integer(v)

for (r in seq_len(nrow(df))) {
  row <- df[r, ]
  tab <- readr::read_csv(file.path(save_path, "density_vectors", paste0("read", row$Serial, ".csv")), show_col_types = FALSE)
  track <- function(sfx) data.frame(ID = tab$ID, start_index = tab$start_index, end_index = tab$end_index,
                                    density = tab[[paste0("density", sfx)]], class = tab[[paste0("class", sfx)]])
  common <- list(seq_length = row$sequence_length, subs = track(""), subs_mismatch = track("_mismatch"),
                 serial_num = row$Serial, seq_start = na1(row$Telomere_start), seq_end = na1(row$Telomere_end),
                 gray_start = na1(row$Telomere_start_mismatch), gray_end = na1(row$Telomere_end_mismatch),
                 save_it = TRUE, main_title = title, w = 750, h = 300)
  if (!has_tvr) {
    f <- plot_single_telo_with_gray_area
  } else {
    f <- plot_single_telo_with_tvr
    common <- c(common, list(subs_tvr = track("_mismatch_tvr"), tvr_start = na1(row$Telomere_start_mismatch_tvr),
                             tvr_end = na1(row$Telomere_end_mismatch_tvr)))
  }
  do.call(f, c(common, list(x_length = max_length, output_jpegs = output_jpegs)))                          # :1877 / :1898
  do.call(f, c(common, list(x_length = row$sequence_length, output_jpegs = output_jpegs_1)))               # :1884 / :1905
  do.call(f, c(common, list(x_length = row$sequence_length, output_jpegs = output_jpegs_1, eps = TRUE)))   # :1892 / :1913
}
