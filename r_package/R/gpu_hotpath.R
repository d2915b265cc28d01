# Drop-in bodies for the three hot functions of LDWeaver (signatures and return values unchanged).
# They replace R/extractSNPs.R:23-142,168-281, R/performPopulationStuctureCorrection.R:20-81 and the block loop of
# R/computePairwiseMI.R:69-116.  What follows the scan (R/computePairwiseMI.R:118-143) runs through the library's host
# entry points: mergeNsort_sr_links -> ldw_sr_postprocess (no plots / .rds files), runARACNE -> ldw_run_aracne (also a
# drop-in for the exported runARACNE, used by analyse_long_range_links, R/lr_analyser.R:101-108).  Setting
# options(LDWeaver.native_post = FALSE) keeps the reference's own R code for those steps.

# Devices (SURVEY section 5): options(LDWeaver.gpus = c(0, 1, 2, 3)) or the environment variable LDW_GPUS ("0,1,2,3", or
# a count "4" meaning devices 0..3); default: device 0.  More than one device runs the weights and the scan on a device
# group (ldw_group_*): the class matrix is uploaded once and broadcast with NCCL, blocks are dealt across the GPUs and
# ONE link table comes back -- results are bit-identical to the single-device run.
.ldw_gpus <- function() {
  g <- getOption("LDWeaver.gpus", NULL)
  if (is.null(g)) {
    e <- Sys.getenv("LDW_GPUS", "")
    if (nzchar(e)) g <- if (grepl(",", e)) as.integer(strsplit(e, ",")[[1]]) else seq_len(as.integer(e)) - 1L
  }
  if (is.null(g) || length(g) == 0) g <- 0L
  as.integer(g)
}

# Drop-in for the Rcpp stub of the same name (R/RcppExports.R:4-6; caller R/estimateCDSDiversity.R:85): masks the
# reference allele of each SNP in the 5 x nsnp matrix `nv` IN PLACE and returns NULL (quirk Q11).
.ACGTN2num <- function(nv, cv, ncores) {
  invisible(.Call("_LDWeaver_ACGTN2num", nv, cv, ncores, PACKAGE = "LDWeaver"))
}

.codes_to_snpdat <- function(enc, pos = NULL) {
  if (enc$seq.length == -1) stop("Error! sequences are of different lengths!")
  if (enc$num.seqs == 0) stop("File does not contain any sequences!")
  if (enc$num.snps == 0) stop("File does not contain any SNPs")
  nsnp <- enc$num.snps; nseq <- enc$num.seqs
  codes <- matrix(as.integer(enc$codes), nrow = nseq, ncol = nsnp)   # codes[k*nseq + s] -> [s, k]
  seq.names <- gsub("^>", "", enc$seq.names)
  uqe <- apply(enc$ACGTN_table > 0, 1, function(x) as.numeric(x > 0))
  POS <- if (is.null(pos)) enc$pos else as.integer(pos[enc$pos])
  mk <- function(a) Matrix::t(Matrix::sparseMatrix(i = row(codes)[codes == a], j = col(codes)[codes == a], x = TRUE,
                                                   dims = c(nseq, nsnp), dimnames = list(seq.names, POS)))
  out <- list(snp.matrix_A = mk(0L), snp.matrix_C = mk(1L), snp.matrix_G = mk(2L), snp.matrix_T = mk(3L),
              snp.matrix_N = mk(4L), g = if (is.null(pos)) enc$seq.length else NULL, nsnp = nsnp, nseq = nseq,
              seq.names = seq.names, r = rowSums(uqe), uqe = uqe, POS = POS)
  attr(out, "codes") <- enc$codes   # raw nsnp x nseq class matrix: lets the next two stages skip the rebuild
  out
}

.snpdat_codes <- function(snp.dat) {
  cd <- attr(snp.dat, "codes")
  if (!is.null(cd)) return(cd)
  # snp.dat restored from RDS (R/BacGWES.R:281,301-302): rebuild the class matrix from the five lgCMatrix slots
  m <- matrix(4L, nrow = snp.dat$nseq, ncol = snp.dat$nsnp)
  for (a in 0:3) {
    M <- snp.dat[[paste0("snp.matrix_", c("A", "C", "G", "T")[a + 1])]]
    idx <- Matrix::which(M, arr.ind = TRUE)          # (snp, seq)
    m[cbind(idx[, 2], idx[, 1])] <- a
  }
  as.raw(m)
}

parse_fasta_alignment <- function(aln_path, gap_freq = 0.15, maf_freq = 0.01, method = "default", mega_dset = F) {
  aln_path <- normalizePath(aln_path)
  if (!file.exists(aln_path)) stop(paste("Can't locate file", aln_path))
  filter <- if (method == "relaxed") 1L else { if (method != "default") warning("Unkown filtering method, using default..."); 0L }
  enc <- .Call("_LDWeaver_gpu_encode", aln_path, filter, gap_freq, maf_freq, .ldw_gpus(), PACKAGE = "LDWeaver")
  .codes_to_snpdat(enc)
}

parse_fasta_SNP_alignment <- function(aln_path, pos, gap_freq = 0.15, maf_freq = 0.01, method = "default", mega_dset = F) {
  aln_path <- normalizePath(aln_path)
  if (!file.exists(aln_path)) stop(paste("Can't locate file", aln_path))
  filter <- if (method == "relaxed") 1L else { if (method != "default") warning("Unkown filtering method, using default..."); 0L }
  enc <- .Call("_LDWeaver_gpu_encode", aln_path, filter, gap_freq, maf_freq, .ldw_gpus(), PACKAGE = "LDWeaver")
  if (enc$seq.length > 0 && length(pos) != enc$seq.length) stop("Error! Number of positions do not match the fasta sequence length")
  .codes_to_snpdat(enc, pos)
}

estimate_Hamming_distance_weights <- function(snp.dat, threshold = 0.1, mega_dset = F) {
  t0 <- Sys.time()
  hdw <- .Call("_LDWeaver_gpu_hdw", .snpdat_codes(snp.dat), snp.dat$nsnp, snp.dat$nseq, threshold, .ldw_gpus(), PACKAGE = "LDWeaver")
  names(hdw) <- snp.dat$seq.names
  cat(paste("Done in", round(difftime(Sys.time(), t0, units = "secs"), 2), "s\n"))
  hdw
}

perform_MI_computation <- function(snp.dat, hdw, cds_var, ncores, lr_save_path = NULL, sr_save_path = NULL, plt_folder = NULL,
                                   sr_dist = 20000, lr_retain_links = 1e6, max_blk_sz = 10000, srp_cutoff = 3, runARACNE = TRUE,
                                   perform_SR_analysis_only = FALSE, order_links = T, mega_dset = F) {
  t000 <- Sys.time()
  if (is.null(lr_save_path)) lr_save_path <- file.path(getwd(), "lr_links.tsv")
  if (is.null(sr_save_path)) sr_save_path <- file.path(getwd(), "sr_links.tsv")
  if (is.null(plt_folder)) plt_folder <- file.path(getwd(), "PLOTS")
  if (!file.exists(plt_folder)) dir.create(plt_folder)
  cat("Begin MI computation... \n")
  max_blk_sz <- round(max_blk_sz, -3)
  lr_links_approx <- 1
  if (!perform_SR_analysis_only) {   # R/computePairwiseMI.R:94-97, kept in R because it uses R's RNG stream
    snp_subset <- min(snp.dat$nsnp, round(snp.dat$nsnp * 0.1))
    set.seed(1988)
    lr_link_count <- sapply(snp.dat$POS[sample(snp.dat$nsnp, snp_subset)], function(x)
      sum((0.5 * snp.dat$g - abs((x - snp.dat$POS) %% snp.dat$g - 0.5 * snp.dat$g)) > sr_dist))
    lr_links_approx <- sum(lr_link_count) / snp_subset * snp.dat$nsnp / 2
  }
  gpus <- .ldw_gpus()
  if (length(gpus) == 1 && isTRUE(getOption("LDWeaver.native_post", TRUE)) && isTRUE(getOption("LDWeaver.device_post", TRUE))) {
    # default on one GPU: the short-range table (9e7 rows at 616 x 100k) stays in device memory; the scan refines its MI in
    # fp64 and mergeNsort_sr_links runs on the device (ldw_sr_postprocess_dev) -- only the rows of sr_links_df come back
    res <- .Call("_LDWeaver_gpu_mi_scan_post", .snpdat_codes(snp.dat), snp.dat$nsnp, snp.dat$nseq, as.numeric(hdw),
                 as.integer(snp.dat$POS), as.integer(cds_var$paint), as.numeric(snp.dat$g), sr_dist, lr_retain_links,
                 lr_links_approx, max_blk_sz, perform_SR_analysis_only, as.integer(cds_var$nclust), srp_cutoff, gpus,
                 PACKAGE = "LDWeaver")
    if (length(res$lr$MI) > 0)
      write.table(x = as.data.frame(res$lr[1:6]), file = lr_save_path, append = T, quote = F, row.names = F, col.names = F, sep = '\t')
    post <- res$post
    frame <- function(idx) data.frame(clust_c = post$clust_c[idx], pos1 = post$rows$pos1[idx], pos2 = post$rows$pos2[idx],
                                      clust1 = as.numeric(post$rows$clust1[idx]), clust2 = as.numeric(post$rows$clust2[idx]),
                                      len = post$rows$len[idx], MI = post$rows$MI[idx], srp_max = post$srp_max[idx])
    sr_links_red <- frame(post$red)
    sr_links_ARACNE_check <- frame(post$chk)
  } else {
  res <- .Call("_LDWeaver_gpu_mi_scan", .snpdat_codes(snp.dat), snp.dat$nsnp, snp.dat$nseq, as.numeric(hdw),
               as.integer(snp.dat$POS), as.integer(cds_var$paint), as.numeric(snp.dat$g), sr_dist, lr_retain_links,
               lr_links_approx, max_blk_sz, perform_SR_analysis_only,
               isTRUE(getOption("LDWeaver.exact_sr", TRUE)),   # fp64 MI for the short-range links (the fp32 epilogue's 2e-7 is amplified by the beta fit below)
               gpus, PACKAGE = "LDWeaver")
  if (length(res$lr$MI) > 0)   # same rows, same order as the per-block appends of R/computePairwiseMI.R:362
    write.table(x = as.data.frame(res$lr[1:6]), file = lr_save_path, append = T, quote = F, row.names = F, col.names = F, sep = '\t')
  if (!isTRUE(getOption("LDWeaver.native_post", TRUE))) {
    sr_df <- as.data.frame(res$sr[1:6])
    sr_links <- lapply(1:cds_var$nclust, function(i) sr_df[sr_df$clust1 == i | sr_df$clust2 == i, ])   # :372-376
    sr_links_all <- mergeNsort_sr_links(cds_var = cds_var, sr_links = sr_links, sr_dist = sr_dist, plt_path = plt_folder, srp_cutoff = srp_cutoff)
    sr_links_red <- sr_links_all$sr_links_red
    sr_links_ARACNE_check <- sr_links_all$sr_links_ARACNE_check
  } else {
    post <- .Call("_LDWeaver_gpu_sr_post", as.integer(res$sr$pos1), as.integer(res$sr$pos2), res$sr$clust1, res$sr$clust2,
                  as.integer(res$sr$len), res$sr$MI,
                  as.integer(cds_var$nclust), sr_dist, srp_cutoff, PACKAGE = "LDWeaver")
    frame <- function(idx) {   # columns of sr_links_df (R/computePairwiseMI.R:470): clust_c + the six link columns + srp_max
      r <- post$row[idx]
      data.frame(clust_c = post$clust_c[idx], pos1 = res$sr$pos1[r], pos2 = res$sr$pos2[r], clust1 = as.numeric(res$sr$clust1[r]),
                 clust2 = as.numeric(res$sr$clust2[r]), len = as.numeric(res$sr$len[r]), MI = res$sr$MI[r], srp_max = post$srp_max[idx])
    }
    sr_links_red <- frame(post$red)
    sr_links_ARACNE_check <- frame(post$chk)
  }
  }
  if (runARACNE) {
    sr_links_red$ARACNE <- as.numeric(runARACNE(sr_links_red, sr_links_ARACNE_check))
  } else {
    warning('ARACNE not run, all values will be set to 1')
    sr_links_red$ARACNE <- 1
  }
  if (order_links) { sr_links_red <- sr_links_red[order(sr_links_red$srp_max, decreasing = T), ]; rownames(sr_links_red) <- NULL }
  write.table(x = sr_links_red, file = sr_save_path, append = T, quote = F, row.names = F, col.names = F, sep = '\t')
  cat(paste("All done in", round(difftime(Sys.time(), t000, units = "mins"), 2), "mins \n"))
  sr_links_red
}

# Drop-in for runARACNE (R/io_functions.R:101-164): same arguments, same logical vector
runARACNE <- function(links_to_check, links_full) {
  t0 <- Sys.time()
  cat(paste("Running ARACNE on", nrow(links_to_check), "links... \n"))
  out <- .Call("_LDWeaver_gpu_runARACNE", as.numeric(links_to_check$pos1), as.numeric(links_to_check$pos2), as.numeric(links_to_check$MI),
               as.numeric(links_full$pos1), as.numeric(links_full$pos2), as.numeric(links_full$MI), PACKAGE = "LDWeaver")
  cat(paste("\nDone in", round(difftime(Sys.time(), t0, units = "secs"), 2), "s\n"))
  out
}
