#!/usr/bin/env Rscript
# Ground truth from the REAL reference package, for anyone with an R installation (R is not in the build image, so
# this script has never run there; it uses only LDWeaver's exported API, man/*.Rd).
#
#   Rscript baseline/run_reference.R <aln.fa[.gz]> <out_dir> [pos_file|-] [g|-] [method] [sr_dist] [lr_retain_links]
#                                    [max_blk_sz] [ncores] [threshold]
#
# Inputs of the same bytes as the GPU arm: `python tools/make_synthetic_fasta.py C2 out.fa.gz` writes the synthetic
# alignment bench.py scans (and its .pos file).  Outputs under <out_dir>, compared by tools/compare_with_reference.py:
#   POS.txt r.txt uqe.tsv seq_names.txt     fields of snp.dat            (R/extractSNPs.R:138-141)
#   hdw.txt                                 17 significant digits        (R/performPopulationStuctureCorrection.R:76)
#   lr_links.tsv sr_links.tsv               as the package writes them   (R/computePairwiseMI.R:140,362)
#   timings.json                            wall seconds per stage, ncores
# cds_var is not derived from an annotation (none for synthetic data): paint = three contiguous thirds, nclust = 3,
# exactly what ldweaver_b200.synth / bench.py use.
suppressPackageStartupMessages(library(LDWeaver))
args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 2) stop("usage: run_reference.R <aln> <out_dir> [pos_file|-] [g|-] [method] [sr_dist] [lr_retain_links] [max_blk_sz] [ncores] [threshold]")
arg <- function(i, default) if (length(args) >= i && args[i] != "-") args[i] else default
aln <- args[1]
out <- args[2]
pos_file <- arg(3, NA)
g <- as.numeric(arg(4, NA))
method <- arg(5, "default")
sr_dist <- as.numeric(arg(6, 20000))
lr_retain_links <- as.numeric(arg(7, 1e6))
max_blk_sz <- as.numeric(arg(8, 10000))
ncores <- as.integer(arg(9, parallel::detectCores()))
threshold <- as.numeric(arg(10, 0.1))
dir.create(out, showWarnings = FALSE, recursive = TRUE)
tm <- list(ncores = ncores)

t0 <- Sys.time()
if (!is.na(pos_file)) {
  pos <- as.numeric(readLines(pos_file))
  snp.dat <- LDWeaver::parse_fasta_SNP_alignment(aln, pos = pos, method = method)
  if (is.na(g)) stop("SNP-only input needs the genome length g (4th argument)")
  snp.dat$g <- g                                    # as LDWeaver() patches it from the annotation, R/BacGWES.R:338-345
} else {
  snp.dat <- LDWeaver::parse_fasta_alignment(aln, method = method)
  if (!is.na(g)) snp.dat$g <- g
}
tm$parse_s <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
writeLines(format(snp.dat$POS, scientific = FALSE, trim = TRUE), file.path(out, "POS.txt"))
writeLines(as.character(snp.dat$r), file.path(out, "r.txt"))
write.table(snp.dat$uqe, file.path(out, "uqe.tsv"), sep = "\t", row.names = FALSE, col.names = FALSE, quote = FALSE)
writeLines(snp.dat$seq.names, file.path(out, "seq_names.txt"))

t0 <- Sys.time()
hdw <- LDWeaver::estimate_Hamming_distance_weights(snp.dat, threshold = threshold)
tm$hdw_s <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
writeLines(sprintf("%.17g", hdw), file.path(out, "hdw.txt"))

n <- snp.dat$nsnp
paint <- rep(1L, n)
paint[(n %/% 3 + 1):n] <- 2L
paint[(2 * n %/% 3 + 1):n] <- 3L
cds_var <- list(paint = paint, nclust = 3)

lr_path <- file.path(out, "lr_links.tsv")
sr_path <- file.path(out, "sr_links.tsv")
unlink(c(lr_path, sr_path))                         # both are opened with append = TRUE
t0 <- Sys.time()
sr_links <- LDWeaver::perform_MI_computation(snp.dat = snp.dat, hdw = hdw, cds_var = cds_var, ncores = ncores,
                                             lr_save_path = lr_path, sr_save_path = sr_path,
                                             plt_folder = file.path(out, "plots"), sr_dist = sr_dist,
                                             lr_retain_links = lr_retain_links, max_blk_sz = max_blk_sz)
tm$perform_MI_computation_s <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
tm$nsnp <- n
tm$nseq <- snp.dat$nseq
tm$pairs_per_s <- (n * (n - 1) / 2) / tm$perform_MI_computation_s
writeLines(paste0("{", paste(sprintf('"%s": %s', names(tm), format(unlist(tm), digits = 10, scientific = FALSE, trim = TRUE)), collapse = ", "), "}"),
           file.path(out, "timings.json"))
cat("done:", out, "\n")
