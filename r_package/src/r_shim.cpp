// R <-> libldwgpu glue (.Call entry points).  NOT compiled in the build container (no R headers there); kept
// deliberately thin so that review suffices.  Build inside the LDWeaver package with the Makevars next to it.
//
// Replaces, for the hot path only, the Rcpp-generated shims of the reference
// (src/RcppExports.cpp:16 _LDWeaver_ACGTN2num, :109 _LDWeaver_extractAlnParam, :123 _LDWeaver_extractSNPs)
// and adds the two entry points that the bodies of estimate_Hamming_distance_weights() and
// perform_MI_computation() call instead of Matrix/MatrixExtra + .fastHadamard.
//
// Conventions: all SEXP work happens on the calling (R main) thread; the library never prints; a nonzero
// return code becomes Rf_error(ldw_last_error()) after every native resource has been released.
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ldw.h"

namespace {

ldw_ctx* g_ctx = nullptr;

ldw_ctx* ctx() {
  if (!g_ctx) {
    int dev = 0;
    const char* e = getenv("LDW_DEVICE");
    if (e) dev = atoi(e);
    if (ldw_create(dev, &g_ctx) != 0) Rf_error("%s", ldw_last_error());
  }
  return g_ctx;
}

SEXP links_to_list(const ldw_links& L) {
  // data.frame-ready list of columns: pos1, pos2, clust1, clust2, len, MI (R/computePairwiseMI.R:326-331) + block
  const char* names[] = {"pos1", "pos2", "clust1", "clust2", "len", "MI", "block"};
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 7));
  const int32_t* icol[] = {L.pos1, L.pos2, L.clust1, L.clust2, L.len, nullptr, L.block};
  for (int k = 0; k < 7; k++) {
    SET_STRING_ELT(nm, k, Rf_mkChar(names[k]));
    SEXP col;
    if (k == 5) {
      col = PROTECT(Rf_allocVector(REALSXP, L.n));
      if (L.n) memcpy(REAL(col), L.MI, sizeof(double) * (size_t)L.n);
    } else if (k == 2 || k == 3 || k == 6) {
      col = PROTECT(Rf_allocVector(INTSXP, L.n));
      if (L.n) memcpy(INTEGER(col), icol[k], sizeof(int) * (size_t)L.n);
    } else {  // pos1, pos2, len are numeric in the reference's data.frame
      col = PROTECT(Rf_allocVector(REALSXP, L.n));
      double* d = REAL(col);
      for (int64_t i = 0; i < L.n; i++) d[i] = (double)icol[k][i];
    }
    SET_VECTOR_ELT(out, k, col);
    UNPROTECT(1);
  }
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}

}  // namespace

extern "C" {

// .Call("_LDWeaver_gpu_encode", path, filter, gap, maf) -> list(num.seqs, num.snps, seq.length, seq.names, pos,
//                                                              codes (raw nsnp x nseq), ACGTN_table (5 x nsnp))
// stands in for .extractAlnParam + .extractSNPs (R/extractSNPs.R:39,45)
SEXP LDWeaver_gpu_encode(SEXP path_, SEXP filter_, SEXP gap_, SEXP maf_) {
  const char* path = CHAR(STRING_ELT(path_, 0));
  int64_t nseq = 0, slen = 0, names_len = 0, nsnp = 0;
  uint8_t* aln = nullptr;   // library-allocated (ldw_read_fasta_alloc: one pass over the file), released before any Rf_error
  char* names = nullptr;
  int32_t* pos = nullptr;
  auto release = [&]() { ldw_buffer_free(aln); ldw_buffer_free(names); free(pos); aln = nullptr; names = nullptr; pos = nullptr; };
  if (ldw_read_fasta_alloc(path, &nseq, &slen, &aln, &names, &names_len) != 0) Rf_error("%s", ldw_last_error());
  if (slen == -1 || nseq == 0) {  // sentinel values the R wrapper turns into stop() (R/extractSNPs.R:41-42)
    release();
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, Rf_ScalarInteger((int)nseq));
    SET_VECTOR_ELT(out, 1, Rf_ScalarInteger((int)slen));
    UNPROTECT(1);
    return out;
  }
  pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)(slen > 0 ? slen : 1));
  if (!pos) { release(); Rf_error("out of memory"); }
  if (ldw_aln_param(ctx(), aln, nseq, slen, Rf_asInteger(filter_), Rf_asReal(gap_), Rf_asReal(maf_), pos, &nsnp, nullptr) != 0) {
    release();
    Rf_error("%s", ldw_last_error());
  }
  SEXP codes = PROTECT(Rf_allocVector(RAWSXP, nsnp * nseq));
  SEXP table = PROTECT(Rf_allocMatrix(REALSXP, 5, (int)nsnp));
  if (nsnp > 0 && ldw_extract_snps(ctx(), aln, nseq, slen, pos, nsnp, RAW(codes), REAL(table)) != 0) {
    release();
    UNPROTECT(2);
    Rf_error("%s", ldw_last_error());
  }
  SEXP rpos = PROTECT(Rf_allocVector(INTSXP, nsnp));
  if (nsnp) memcpy(INTEGER(rpos), pos, sizeof(int) * (size_t)nsnp);
  SEXP rnames = PROTECT(Rf_allocVector(STRSXP, nseq));
  const char* q = names;
  for (int64_t i = 0; i < nseq; i++) { SET_STRING_ELT(rnames, i, Rf_mkChar(q)); q += strlen(q) + 1; }
  release();
  const char* nm[] = {"num.seqs", "num.snps", "seq.length", "seq.names", "pos", "codes", "ACGTN_table"};
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
  SEXP onm = PROTECT(Rf_allocVector(STRSXP, 7));
  for (int k = 0; k < 7; k++) SET_STRING_ELT(onm, k, Rf_mkChar(nm[k]));
  SET_VECTOR_ELT(out, 0, Rf_ScalarInteger((int)nseq));
  SET_VECTOR_ELT(out, 1, Rf_ScalarInteger((int)nsnp));
  SET_VECTOR_ELT(out, 2, Rf_ScalarInteger((int)slen));
  SET_VECTOR_ELT(out, 3, rnames);
  SET_VECTOR_ELT(out, 4, rpos);
  SET_VECTOR_ELT(out, 5, codes);
  SET_VECTOR_ELT(out, 6, table);
  Rf_setAttrib(out, R_NamesSymbol, onm);
  UNPROTECT(6);
  return out;
}

// .Call("_LDWeaver_gpu_hdw", codes(raw nsnp*nseq), nsnp, nseq, threshold) -> numeric(nseq)
// body of estimate_Hamming_distance_weights (R/performPopulationStuctureCorrection.R:23-76)
SEXP LDWeaver_gpu_hdw(SEXP codes_, SEXP nsnp_, SEXP nseq_, SEXP thr_) {
  int64_t n = (int64_t)Rf_asReal(nsnp_), S = (int64_t)Rf_asReal(nseq_);
  SEXP w = PROTECT(Rf_allocVector(REALSXP, S));
  int rc = ldw_hdw(ctx(), RAW(codes_), n, S, Rf_asReal(thr_), nullptr, REAL(w), nullptr);
  UNPROTECT(1);
  if (rc != 0) Rf_error("%s", ldw_last_error());
  return w;
}

// .Call("_LDWeaver_gpu_mi_scan", codes, nsnp, nseq, hdw, POS, paint, g, sr_dist, lr_retain_links, lr_links_approx, blk, sr_only,
//       exact_sr) -> list(sr = <columns>, lr = <columns>, borderline = <columns>, thr = numeric(nblocks))
// scan part of perform_MI_computation (R/computePairwiseMI.R:69-116).  exact_sr = TRUE replaces the fp32-accurate MI of the
// short-range links by fp64 values (ldw_links_to_cells + ldw_mi_pairs_exact per block) before R derives statistics from them.
SEXP LDWeaver_gpu_mi_scan(SEXP codes_, SEXP nsnp_, SEXP nseq_, SEXP hdw_, SEXP pos_, SEXP paint_, SEXP g_, SEXP srd_, SEXP retain_,
                          SEXP approx_, SEXP blk_, SEXP sronly_, SEXP exact_) {
  int64_t n = (int64_t)Rf_asReal(nsnp_), S = (int64_t)Rf_asReal(nseq_), blk = (int64_t)Rf_asReal(blk_);
  ldw_mi_plan* plan = nullptr;
  if (ldw_mi_plan_create(ctx(), RAW(codes_), n, S, REAL(hdw_), INTEGER(pos_), INTEGER(paint_), blk, &plan) != 0)
    Rf_error("%s", ldw_last_error());
  int64_t nr = (n + blk - 1) / blk, nblk = nr * (nr + 1) / 2;
  SEXP thr = PROTECT(Rf_allocVector(REALSXP, nblk));
  ldw_links sr, lr, bd;
  ldw_scan_stats st;
  int flags = Rf_asLogical(sronly_) ? LDW_SCAN_SR_ONLY : 0;
  // SR-only scans index reduced SNP lists (quirk Q12), which ldw_links_to_cells cannot address: there the scan itself
  // refines the short-range MI (mi_sr_exact_kernel)
  if ((flags & LDW_SCAN_SR_ONLY) && Rf_asLogical(exact_)) flags |= LDW_SCAN_SR_EXACT;
  int rc = ldw_mi_scan(plan, Rf_asReal(g_), Rf_asReal(srd_), Rf_asReal(retain_), Rf_asReal(approx_), flags, 1, 0, &sr, &lr, &bd,
                       REAL(thr), nullptr, &st);
  if (rc != 0) {
    ldw_mi_plan_destroy(plan);
    UNPROTECT(1);
    Rf_error("%s", ldw_last_error());
  }
  // cells of the short-range links in their blocks' MI matrices (host only), while the library-owned columns are at hand
  const bool exact = Rf_asLogical(exact_) && !(flags & LDW_SCAN_SR_ONLY) && sr.n > 0;
  std::vector<int32_t> fl, tl, blk_of;
  if (exact) {
    fl.resize((size_t)sr.n); tl.resize((size_t)sr.n);
    if (ldw_links_to_cells(INTEGER(pos_), n, blk, &sr, fl.data(), tl.data()) != 0) {
      ldw_mi_plan_destroy(plan);
      UNPROTECT(1);
      Rf_error("%s", ldw_last_error());
    }
    blk_of.assign(sr.block, sr.block + sr.n);
  }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, links_to_list(sr));
  if (exact) {  // rows are in make_blocks order: one ldw_mi_pairs_exact call per run of equal block ids, written in place
    double* mi = REAL(VECTOR_ELT(VECTOR_ELT(out, 0), 5));
    for (int64_t lo = 0; lo < sr.n;) {
      int64_t hi = lo;
      while (hi < sr.n && blk_of[hi] == blk_of[lo]) hi++;
      if (ldw_mi_pairs_exact(plan, blk_of[lo], fl.data() + lo, tl.data() + lo, hi - lo, mi + lo) != 0) {
        ldw_mi_plan_destroy(plan);
        UNPROTECT(2);
        Rf_error("%s", ldw_last_error());
      }
      lo = hi;
    }
  }
  SET_VECTOR_ELT(out, 1, links_to_list(lr));
  SET_VECTOR_ELT(out, 2, links_to_list(bd));
  SET_VECTOR_ELT(out, 3, thr);
  const char* nm[] = {"sr", "lr", "borderline", "thr"};
  SEXP onm = PROTECT(Rf_allocVector(STRSXP, 4));
  for (int k = 0; k < 4; k++) SET_STRING_ELT(onm, k, Rf_mkChar(nm[k]));
  Rf_setAttrib(out, R_NamesSymbol, onm);
  ldw_mi_plan_destroy(plan);  // link columns were copied into R vectors above
  UNPROTECT(3);
  return out;
}

// .Call("_LDWeaver_ACGTN2num", nv, cv, ncores): same symbol and in-place semantics as the reference
// (src/ACGTN2num_parallel.cpp:10-43, quirk Q11)
SEXP LDWeaver_gpu_ACGTN2num(SEXP nv_, SEXP cv_, SEXP ncores_) {
  R_xlen_t n = XLENGTH(cv_);
  std::vector<char> ref((size_t)n);
  for (R_xlen_t i = 0; i < n; i++) {
    const char* s = CHAR(STRING_ELT(cv_, i));
    ref[i] = s[0];
  }
  if (ldw_acgtn2num(ctx(), REAL(nv_), ref.data(), (int64_t)n) != 0) Rf_error("%s", ldw_last_error());
  return R_NilValue;
}

// .Call("_LDWeaver_gpu_runARACNE", chk_pos1, chk_pos2, chk_MI, full_pos1, full_pos2, full_MI) -> logical(length(chk_MI))
// body of runARACNE(links_to_check, links_full) (R/io_functions.R:101-164); all arguments numeric vectors
SEXP LDWeaver_gpu_runARACNE(SEXP c1_, SEXP c2_, SEXP cm_, SEXP f1_, SEXP f2_, SEXP fm_) {
  const R_xlen_t nc = XLENGTH(cm_), nf = XLENGTH(fm_);
  std::vector<uint8_t> keep((size_t)nc, 1);
  if (ldw_run_aracne((int64_t)nc, REAL(c1_), REAL(c2_), REAL(cm_), (int64_t)nf, REAL(f1_), REAL(f2_), REAL(fm_), keep.data()) != 0)
    Rf_error("%s", ldw_last_error());
  SEXP out = PROTECT(Rf_allocVector(LGLSXP, nc));
  for (R_xlen_t i = 0; i < nc; i++) LOGICAL(out)[i] = keep[i];
  UNPROTECT(1);
  return out;
}

// .Call("_LDWeaver_gpu_sr_post", pos1, pos2, clust1, clust2, len, MI (the scan's short-range columns: integer x5, numeric),
//       nclust, sr_dist, srp_cutoff) -> list(clust_c, row (1-based), srp_max, red (1-based), chk (1-based), shape, coef)
// body of mergeNsort_sr_links (R/computePairwiseMI.R:400-495) on the un-split short-range table
SEXP LDWeaver_gpu_sr_post(SEXP p1_, SEXP p2_, SEXP c1_, SEXP c2_, SEXP len_, SEXP mi_, SEXP nclust_, SEXP srd_, SEXP cut_) {
  ldw_links sr;
  memset(&sr, 0, sizeof(sr));
  sr.n = (int64_t)XLENGTH(mi_);
  sr.pos1 = INTEGER(p1_); sr.pos2 = INTEGER(p2_); sr.clust1 = INTEGER(c1_); sr.clust2 = INTEGER(c2_); sr.len = INTEGER(len_);
  sr.MI = REAL(mi_);
  ldw_sr_post post;
  const int nclust = Rf_asInteger(nclust_);
  if (ldw_sr_postprocess(&sr, nclust, Rf_asReal(srd_), Rf_asReal(cut_), &post) != 0) Rf_error("%s", ldw_last_error());
  SEXP cc = PROTECT(Rf_allocVector(INTSXP, post.n_df)), row = PROTECT(Rf_allocVector(REALSXP, post.n_df));
  SEXP srp = PROTECT(Rf_allocVector(REALSXP, post.n_df));
  SEXP red = PROTECT(Rf_allocVector(REALSXP, post.n_red)), chk = PROTECT(Rf_allocVector(REALSXP, post.n_chk));
  SEXP shape = PROTECT(Rf_allocVector(REALSXP, 2 * nclust)), coef = PROTECT(Rf_allocVector(REALSXP, 2 * nclust));
  for (int64_t i = 0; i < post.n_df; i++) { INTEGER(cc)[i] = post.clust_c[i]; REAL(row)[i] = (double)(post.row[i] + 1); REAL(srp)[i] = post.srp_max[i]; }
  for (int64_t i = 0; i < post.n_red; i++) REAL(red)[i] = (double)(post.red[i] + 1);
  for (int64_t i = 0; i < post.n_chk; i++) REAL(chk)[i] = (double)(post.chk[i] + 1);
  for (int k = 0; k < 2 * nclust; k++) { REAL(shape)[k] = post.shape[k]; REAL(coef)[k] = post.coef[k]; }
  ldw_sr_post_free(&post);  // everything needed was copied into R vectors above
  const char* nm[] = {"clust_c", "row", "srp_max", "red", "chk", "shape", "coef"};
  SEXP vals[] = {cc, row, srp, red, chk, shape, coef};
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 7)), onm = PROTECT(Rf_allocVector(STRSXP, 7));
  for (int k = 0; k < 7; k++) { SET_VECTOR_ELT(out, k, vals[k]); SET_STRING_ELT(onm, k, Rf_mkChar(nm[k])); }
  Rf_setAttrib(out, R_NamesSymbol, onm);
  UNPROTECT(9);
  return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"_LDWeaver_gpu_encode", (DL_FUNC)&LDWeaver_gpu_encode, 4},
    {"_LDWeaver_gpu_hdw", (DL_FUNC)&LDWeaver_gpu_hdw, 4},
    {"_LDWeaver_gpu_mi_scan", (DL_FUNC)&LDWeaver_gpu_mi_scan, 13},
    {"_LDWeaver_gpu_ACGTN2num", (DL_FUNC)&LDWeaver_gpu_ACGTN2num, 3},
    {"_LDWeaver_gpu_runARACNE", (DL_FUNC)&LDWeaver_gpu_runARACNE, 6},
    {"_LDWeaver_gpu_sr_post", (DL_FUNC)&LDWeaver_gpu_sr_post, 9},
    {NULL, NULL, 0}};

// merged into R_init_LDWeaver (src/RcppExports.cpp:169-172) next to the Rcpp-generated table
void R_init_LDWeaver_gpu(DllInfo* dll) { R_registerRoutines(dll, NULL, CallEntries, NULL, NULL); }

}  // extern "C"
