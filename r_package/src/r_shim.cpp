// R <-> libldwgpu glue (.Call entry points).  NEVER compiled against real R headers in the build image (R is not
// installed there); tests/test_abi_cpu.py type-checks it against include/ldw.h with a mock of the few R C API
// functions it uses (tests/mock_r/).  Kept thin and C-style on purpose.
//
// Replaces, for the hot path only, the Rcpp-generated shims of the reference
// (src/RcppExports.cpp:16 _LDWeaver_ACGTN2num -- same symbol, same in-place semantics --, :109
// _LDWeaver_extractAlnParam, :123 _LDWeaver_extractSNPs) and adds the entry points that the bodies of
// estimate_Hamming_distance_weights() and perform_MI_computation() call instead of Matrix/MatrixExtra + .fastHadamard.
//
// Conventions:
//  * all SEXP work happens on the calling (R main) thread; the library never prints;
//  * Rf_error() is a longjmp: NO object with a destructor is ever live across it.  Scratch memory comes from R_alloc
//    (R reclaims it when .Call returns, error or not); library-owned buffers are released explicitly before every
//    Rf_error; a nonzero return code becomes Rf_error(ldw_last_error()) only after that;
//  * devices: every entry point takes `gpus` (integer vector of CUDA device ids, from options(LDWeaver.gpus = ) or the
//    LDW_GPUS environment variable, resolved in R/gpu_hotpath.R).  One id -> the single-device entry points; several ->
//    a device group (ldw_group_*: matrix uploaded once + NCCL broadcast, work dealt across the GPUs, one result table).
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ldw.h"

namespace {

enum { MAX_GPUS = 16 };
ldw_ctx* g_ctx = nullptr;      // single-device context (device g_ctx_dev)
int g_ctx_dev = -1;
ldw_group* g_group = nullptr;  // device group over g_group_devs
int g_group_devs[MAX_GPUS];
int g_group_n = 0;

// the context for one device; (re)created when the device changes.  Errors out through Rf_error (nothing to release).
ldw_ctx* ctx_for(int dev) {
  if (g_ctx && g_ctx_dev != dev) { ldw_destroy(g_ctx); g_ctx = nullptr; }
  if (!g_ctx) {
    if (ldw_create(dev, &g_ctx) != 0) Rf_error("%s", ldw_last_error());
    g_ctx_dev = dev;
  }
  return g_ctx;
}

ldw_group* group_for(const int* devs, int n) {
  if (n > MAX_GPUS) Rf_error("at most %d GPUs", (int)MAX_GPUS);
  bool same = g_group && g_group_n == n;
  for (int k = 0; same && k < n; k++) same = g_group_devs[k] == devs[k];
  if (!same) {
    if (g_group) { ldw_group_destroy(g_group); g_group = nullptr; }
    if (ldw_group_create(devs, n, &g_group) != 0) Rf_error("%s", ldw_last_error());
    g_group_n = n;
    for (int k = 0; k < n; k++) g_group_devs[k] = devs[k];
  }
  return g_group;
}

int first_gpu(SEXP gpus_) { return XLENGTH(gpus_) > 0 ? INTEGER(gpus_)[0] : 0; }

SEXP links_to_list(const ldw_links* L) {
  // data.frame-ready list of columns: pos1, pos2, clust1, clust2, len, MI (R/computePairwiseMI.R:326-331) + block
  const char* names[] = {"pos1", "pos2", "clust1", "clust2", "len", "MI", "block"};
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 7));
  const int32_t* icol[] = {L->pos1, L->pos2, L->clust1, L->clust2, L->len, nullptr, L->block};
  for (int k = 0; k < 7; k++) {
    SET_STRING_ELT(nm, k, Rf_mkChar(names[k]));
    SEXP col;
    if (k == 5) {
      col = PROTECT(Rf_allocVector(REALSXP, L->n));
      if (L->n) memcpy(REAL(col), L->MI, sizeof(double) * (size_t)L->n);
    } else if (k == 2 || k == 3 || k == 6) {
      col = PROTECT(Rf_allocVector(INTSXP, L->n));
      if (L->n) memcpy(INTEGER(col), icol[k], sizeof(int) * (size_t)L->n);
    } else {  // pos1, pos2, len are numeric in the reference's data.frame
      col = PROTECT(Rf_allocVector(REALSXP, L->n));
      double* d = REAL(col);
      for (int64_t i = 0; i < L->n; i++) d[i] = (double)icol[k][i];
    }
    SET_VECTOR_ELT(out, k, col);
    UNPROTECT(1);
  }
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}

SEXP named_list(int n, const char* const* names, SEXP* vals) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n)), onm = PROTECT(Rf_allocVector(STRSXP, n));
  for (int k = 0; k < n; k++) { SET_VECTOR_ELT(out, k, vals[k]); SET_STRING_ELT(onm, k, Rf_mkChar(names[k])); }
  Rf_setAttrib(out, R_NamesSymbol, onm);
  UNPROTECT(2);
  return out;
}

}  // namespace

extern "C" {

// .Call("_LDWeaver_gpu_encode", path, filter, gap, maf, gpus) -> list(num.seqs, num.snps, seq.length, seq.names, pos,
//                                                                     codes (raw nsnp x nseq), ACGTN_table (5 x nsnp))
// stands in for .extractAlnParam + .extractSNPs (R/extractSNPs.R:39,45).  Domain failures come back as the reference's
// sentinel values in a NAMED list (seq.length = -1: records of different lengths, src/getACGTNsites.cpp:54-56;
// num.seqs = 0; num.snps = 0), which R/gpu_hotpath.R turns into the reference's stop() messages (R/extractSNPs.R:41-43).
SEXP LDWeaver_gpu_encode(SEXP path_, SEXP filter_, SEXP gap_, SEXP maf_, SEXP gpus_) {
  const char* path = CHAR(STRING_ELT(path_, 0));
  int64_t nseq = 0, slen = 0, names_len = 0, nsnp = 0;
  uint8_t* aln = nullptr;   // library-allocated (ldw_read_fasta_alloc: one pass over the file), released before any Rf_error
  char* names = nullptr;
  if (ldw_read_fasta_alloc(path, &nseq, &slen, &aln, &names, &names_len) != 0) Rf_error("%s", ldw_last_error());
  if (slen == -1 || nseq == 0 || slen == 0) {
    ldw_buffer_free(aln);
    ldw_buffer_free(names);
    const char* nm[] = {"num.seqs", "num.snps", "seq.length"};
    SEXP v[3];
    v[0] = PROTECT(Rf_ScalarInteger((int)nseq));
    v[1] = PROTECT(Rf_ScalarInteger(0));
    v[2] = PROTECT(Rf_ScalarInteger((int)slen));
    SEXP out = named_list(3, nm, v);
    UNPROTECT(3);
    return out;
  }
  ldw_ctx* c = ctx_for(first_gpu(gpus_));  // may Rf_error: release first if it does -> create the context before the buffers matter
  int32_t* pos = (int32_t*)R_alloc((size_t)slen, sizeof(int32_t));
  if (ldw_aln_param(c, aln, nseq, slen, Rf_asInteger(filter_), Rf_asReal(gap_), Rf_asReal(maf_), pos, &nsnp, nullptr) != 0) {
    ldw_buffer_free(aln);
    ldw_buffer_free(names);
    Rf_error("%s", ldw_last_error());
  }
  SEXP codes = PROTECT(Rf_allocVector(RAWSXP, nsnp * nseq));
  SEXP table = PROTECT(Rf_allocMatrix(REALSXP, 5, (int)nsnp));
  if (nsnp > 0 && ldw_extract_snps(c, aln, nseq, slen, pos, nsnp, RAW(codes), REAL(table)) != 0) {
    ldw_buffer_free(aln);
    ldw_buffer_free(names);
    UNPROTECT(2);
    Rf_error("%s", ldw_last_error());
  }
  ldw_buffer_free(aln);
  SEXP rpos = PROTECT(Rf_allocVector(INTSXP, nsnp));
  if (nsnp) memcpy(INTEGER(rpos), pos, sizeof(int) * (size_t)nsnp);
  SEXP rnames = PROTECT(Rf_allocVector(STRSXP, nseq));
  const char* q = names;
  for (int64_t i = 0; i < nseq; i++) { SET_STRING_ELT(rnames, i, Rf_mkChar(q)); q += strlen(q) + 1; }
  ldw_buffer_free(names);
  const char* nm[] = {"num.seqs", "num.snps", "seq.length", "seq.names", "pos", "codes", "ACGTN_table"};
  SEXP v[7];
  v[0] = PROTECT(Rf_ScalarInteger((int)nseq));
  v[1] = PROTECT(Rf_ScalarInteger((int)nsnp));
  v[2] = PROTECT(Rf_ScalarInteger((int)slen));
  v[3] = rnames; v[4] = rpos; v[5] = codes; v[6] = table;
  SEXP out = named_list(7, nm, v);
  UNPROTECT(7);
  return out;
}

// .Call("_LDWeaver_gpu_hdw", codes(raw nsnp*nseq), nsnp, nseq, threshold, gpus) -> numeric(nseq)
// body of estimate_Hamming_distance_weights (R/performPopulationStuctureCorrection.R:23-76)
SEXP LDWeaver_gpu_hdw(SEXP codes_, SEXP nsnp_, SEXP nseq_, SEXP thr_, SEXP gpus_) {
  int64_t n = (int64_t)Rf_asReal(nsnp_), S = (int64_t)Rf_asReal(nseq_);
  const int ng = (int)XLENGTH(gpus_);
  SEXP w = PROTECT(Rf_allocVector(REALSXP, S));
  int rc;
  if (ng > 1) {
    ldw_group* G = group_for(INTEGER(gpus_), ng);
    rc = ldw_group_load_codes(G, RAW(codes_), n, S);
    if (rc == 0) rc = ldw_group_hdw(G, Rf_asReal(thr_), 0, nullptr, REAL(w), nullptr);
  } else {
    rc = ldw_hdw(ctx_for(first_gpu(gpus_)), RAW(codes_), n, S, Rf_asReal(thr_), nullptr, REAL(w), nullptr);
  }
  UNPROTECT(1);
  if (rc != 0) Rf_error("%s", ldw_last_error());
  return w;
}

// .Call("_LDWeaver_gpu_mi_scan", codes, nsnp, nseq, hdw, POS, paint, g, sr_dist, lr_retain_links, lr_links_approx, blk, sr_only,
//       exact_sr, gpus) -> list(sr = <columns>, lr = <columns>, borderline = <columns>, thr = numeric(nblocks))
// scan part of perform_MI_computation (R/computePairwiseMI.R:69-116).  exact_sr = TRUE: the MI of the short-range links is
// recomputed in fp64 inside the scan (LDW_SCAN_SR_EXACT) before R derives statistics from it.
SEXP LDWeaver_gpu_mi_scan(SEXP codes_, SEXP nsnp_, SEXP nseq_, SEXP hdw_, SEXP pos_, SEXP paint_, SEXP g_, SEXP srd_, SEXP retain_,
                          SEXP approx_, SEXP blk_, SEXP sronly_, SEXP exact_, SEXP gpus_) {
  int64_t n = (int64_t)Rf_asReal(nsnp_), S = (int64_t)Rf_asReal(nseq_), blk = (int64_t)Rf_asReal(blk_);
  const int ng = (int)XLENGTH(gpus_);
  if (blk < 1) Rf_error("max_blk_sz must be positive");
  int64_t nr = (n + blk - 1) / blk, nblk = nr * (nr + 1) / 2;
  int flags = Rf_asLogical(sronly_) ? LDW_SCAN_SR_ONLY : 0;
  if (Rf_asLogical(exact_)) flags |= LDW_SCAN_SR_EXACT;
  ldw_links sr, lr, bd;
  memset(&sr, 0, sizeof(sr)); memset(&lr, 0, sizeof(lr)); memset(&bd, 0, sizeof(bd));
  SEXP thr = PROTECT(Rf_allocVector(REALSXP, nblk));
  ldw_mi_plan* plan = nullptr;
  int rc;
  if (ng > 1) {
    ldw_group* G = group_for(INTEGER(gpus_), ng);
    ldw_scan_stats* st = (ldw_scan_stats*)R_alloc((size_t)ng, sizeof(ldw_scan_stats));
    rc = ldw_group_load_codes(G, RAW(codes_), n, S);
    if (rc == 0)
      rc = ldw_group_mi_scan(G, REAL(hdw_), INTEGER(pos_), INTEGER(paint_), blk, Rf_asReal(g_), Rf_asReal(srd_), Rf_asReal(retain_),
                             Rf_asReal(approx_), flags, &sr, &lr, &bd, REAL(thr), nullptr, st, nullptr);
  } else {
    ldw_ctx* c = ctx_for(first_gpu(gpus_));
    ldw_scan_stats st;
    rc = ldw_mi_plan_create(c, RAW(codes_), n, S, REAL(hdw_), INTEGER(pos_), INTEGER(paint_), blk, &plan);
    if (rc == 0)
      rc = ldw_mi_scan(plan, Rf_asReal(g_), Rf_asReal(srd_), Rf_asReal(retain_), Rf_asReal(approx_), flags, 1, 0, &sr, &lr, &bd,
                       REAL(thr), nullptr, &st);
  }
  if (rc != 0) {
    if (plan) ldw_mi_plan_destroy(plan);
    UNPROTECT(1);
    Rf_error("%s", ldw_last_error());
  }
  // the link columns are library-owned pinned memory (valid until the next scan): copy them into R vectors.  An
  // allocation failure inside Rf_allocVector longjmps; the plan is then left to the next call's context reuse (it holds
  // device memory only, released with the context), nothing else is live here.
  SEXP v[4];
  v[0] = PROTECT(links_to_list(&sr));
  v[1] = PROTECT(links_to_list(&lr));
  v[2] = PROTECT(links_to_list(&bd));
  v[3] = thr;
  if (plan) ldw_mi_plan_destroy(plan);
  const char* nm[] = {"sr", "lr", "borderline", "thr"};
  SEXP out = named_list(4, nm, v);
  UNPROTECT(4);
  return out;
}

// .Call("_LDWeaver_gpu_mi_scan_post", <the 12 scan arguments up to sr_only>, nclust, srp_cutoff, gpu)
//   -> list(lr, borderline, thr, post = list(clust_c, pos1, pos2, clust1, clust2, len, MI, srp_max, red (1-based), chk (1-based)))
// perform_MI_computation's scan AND mergeNsort_sr_links (R/computePairwiseMI.R:69-116, 400-495) with the short-range table kept
// in device memory: fp64 short-range MI inside the scan (LDW_SCAN_SR_EXACT), LDW_SCAN_SR_ON_DEVICE, ldw_sr_postprocess_dev.
// Only the rows of sr_links_df (a few per cent of the table) cross PCIe.  One device (the first of `gpus`).
SEXP LDWeaver_gpu_mi_scan_post(SEXP codes_, SEXP nsnp_, SEXP nseq_, SEXP hdw_, SEXP pos_, SEXP paint_, SEXP g_, SEXP srd_, SEXP retain_,
                               SEXP approx_, SEXP blk_, SEXP sronly_, SEXP nclust_, SEXP cut_, SEXP gpus_) {
  int64_t n = (int64_t)Rf_asReal(nsnp_), S = (int64_t)Rf_asReal(nseq_), blk = (int64_t)Rf_asReal(blk_);
  if (blk < 1) Rf_error("max_blk_sz must be positive");
  int64_t nr = (n + blk - 1) / blk, nblk = nr * (nr + 1) / 2;
  const int flags = (Rf_asLogical(sronly_) ? LDW_SCAN_SR_ONLY : 0) | LDW_SCAN_SR_EXACT | LDW_SCAN_SR_ON_DEVICE;
  ldw_ctx* c = ctx_for(first_gpu(gpus_));
  ldw_links sr, lr, bd, rows;
  memset(&sr, 0, sizeof(sr)); memset(&lr, 0, sizeof(lr)); memset(&bd, 0, sizeof(bd)); memset(&rows, 0, sizeof(rows));
  ldw_sr_post post;
  memset(&post, 0, sizeof(post));
  ldw_scan_stats st;
  SEXP thr = PROTECT(Rf_allocVector(REALSXP, nblk));
  ldw_mi_plan* plan = nullptr;
  int rc = ldw_mi_plan_create(c, RAW(codes_), n, S, REAL(hdw_), INTEGER(pos_), INTEGER(paint_), blk, &plan);
  if (rc == 0)
    rc = ldw_mi_scan(plan, Rf_asReal(g_), Rf_asReal(srd_), Rf_asReal(retain_), Rf_asReal(approx_), flags, 1, 0, &sr, &lr, &bd, REAL(thr),
                     nullptr, &st);
  if (rc == 0) rc = ldw_sr_postprocess_dev(c, Rf_asInteger(nclust_), Rf_asReal(srd_), Rf_asReal(cut_), &post, &rows);
  if (plan) ldw_mi_plan_destroy(plan);
  if (rc != 0) {
    UNPROTECT(1);
    Rf_error("%s", ldw_last_error());
  }
  SEXP v[4];
  v[0] = PROTECT(links_to_list(&lr));
  v[1] = PROTECT(links_to_list(&bd));
  v[2] = thr;
  SEXP pl = PROTECT(links_to_list(&rows));  // pos1 .. MI, block of the sr_links_df rows
  SEXP cc = PROTECT(Rf_allocVector(INTSXP, post.n_df)), srp = PROTECT(Rf_allocVector(REALSXP, post.n_df));
  SEXP red = PROTECT(Rf_allocVector(REALSXP, post.n_red)), chk = PROTECT(Rf_allocVector(REALSXP, post.n_chk));
  for (int64_t i = 0; i < post.n_df; i++) { INTEGER(cc)[i] = post.clust_c[i]; REAL(srp)[i] = post.srp_max[i]; }
  for (int64_t i = 0; i < post.n_red; i++) REAL(red)[i] = (double)(post.red[i] + 1);
  for (int64_t i = 0; i < post.n_chk; i++) REAL(chk)[i] = (double)(post.chk[i] + 1);
  ldw_sr_post_free(&post);  // rows' columns belonged to it: everything was copied above
  const char* pn[] = {"rows", "clust_c", "srp_max", "red", "chk"};
  SEXP pv[] = {pl, cc, srp, red, chk};
  v[3] = PROTECT(named_list(5, pn, pv));
  const char* nm[] = {"lr", "borderline", "thr", "post"};
  SEXP out = named_list(4, nm, v);
  UNPROTECT(9);
  return out;
}

// .Call("_LDWeaver_ACGTN2num", nv, cv, ncores): the reference's own symbol (src/RcppExports.cpp:16; R stub `.ACGTN2num`,
// R/RcppExports.R:4-6) with its in-place semantics (src/ACGTN2num_parallel.cpp:10-43, quirk Q11).  `ncores` is accepted
// and ignored.  The device is the first of options(LDWeaver.gpus) / LDW_GPUS as resolved at load time (LDW_DEVICE), 0 otherwise.
SEXP LDWeaver_ACGTN2num(SEXP nv_, SEXP cv_, SEXP ncores_) {
  (void)ncores_;
  R_xlen_t n = XLENGTH(cv_);
  char* ref = R_alloc((size_t)n + 1, 1);
  for (R_xlen_t i = 0; i < n; i++) ref[i] = CHAR(STRING_ELT(cv_, i))[0];  // as<char>(cv[c]): first character
  int dev = g_ctx ? g_ctx_dev : 0;
  const char* e = getenv("LDW_DEVICE");
  if (!g_ctx && e) dev = atoi(e);
  if (ldw_acgtn2num(ctx_for(dev), REAL(nv_), ref, (int64_t)n) != 0) Rf_error("%s", ldw_last_error());
  return R_NilValue;
}

// .Call("_LDWeaver_gpu_runARACNE", chk_pos1, chk_pos2, chk_MI, full_pos1, full_pos2, full_MI) -> logical(length(chk_MI))
// body of runARACNE(links_to_check, links_full) (R/io_functions.R:101-164); all arguments numeric vectors
SEXP LDWeaver_gpu_runARACNE(SEXP c1_, SEXP c2_, SEXP cm_, SEXP f1_, SEXP f2_, SEXP fm_) {
  const R_xlen_t nc = XLENGTH(cm_), nf = XLENGTH(fm_);
  uint8_t* keep = (uint8_t*)R_alloc((size_t)nc + 1, 1);
  memset(keep, 1, (size_t)nc);
  if (ldw_run_aracne((int64_t)nc, REAL(c1_), REAL(c2_), REAL(cm_), (int64_t)nf, REAL(f1_), REAL(f2_), REAL(fm_), keep) != 0)
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
  SEXP out = named_list(7, nm, vals);
  UNPROTECT(7);
  return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"_LDWeaver_gpu_encode", (DL_FUNC)&LDWeaver_gpu_encode, 5},
    {"_LDWeaver_gpu_hdw", (DL_FUNC)&LDWeaver_gpu_hdw, 5},
    {"_LDWeaver_gpu_mi_scan", (DL_FUNC)&LDWeaver_gpu_mi_scan, 14},
    {"_LDWeaver_gpu_mi_scan_post", (DL_FUNC)&LDWeaver_gpu_mi_scan_post, 15},
    {"_LDWeaver_ACGTN2num", (DL_FUNC)&LDWeaver_ACGTN2num, 3},  // replaces the Rcpp entry of the same name (src/RcppExports.cpp:155)
    {"_LDWeaver_gpu_runARACNE", (DL_FUNC)&LDWeaver_gpu_runARACNE, 6},
    {"_LDWeaver_gpu_sr_post", (DL_FUNC)&LDWeaver_gpu_sr_post, 9},
    {NULL, NULL, 0}};

// merged into R_init_LDWeaver (src/RcppExports.cpp:169-172) next to the Rcpp-generated table, from which the
// `_LDWeaver_ACGTN2num` row is dropped
void R_init_LDWeaver_gpu(DllInfo* dll) { R_registerRoutines(dll, NULL, CallEntries, NULL, NULL); }

}  // extern "C"
