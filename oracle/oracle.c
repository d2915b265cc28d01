/*
 * CPU ORACLE, C restatement (test infrastructure, NOT the product path).
 *
 * Independent restatement, in plain C + OpenMP, of the reference's genome-wide
 * pairwise-LD hot path (LDWeaver v1.5.2).  It is "reference-shaped": per block it
 * builds the ten sqrt(w)-weighted dense allele sub-matrices, runs the 25
 * [dense x sparse product, five rank-1 temporaries, fastHadamard loop] calls, then
 * enumerates upper/lower-triangle pairs, the circular distance, the type-7
 * quantile and the >= filter -- exactly the work the R code performs -- so it is also
 * the timed CPU baseline of bench.py ("kind": "port").
 *
 * PARITY STATUS: "parity unpinned" (the reference holds no golden vectors for this
 * path and R is absent from the image).  This file and oracle/ldw_oracle.py are two
 * independent restatements that tests/test_oracle.py requires to agree (integers
 * bit-exact, MI <= 1e-12).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * Matrices are column-major like R.  Citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ classification
 * src/getACGTNsites.cpp:59-69, :233-263  (quirk Q8: anything not [AaCcGgTt] -> 4) */
static inline int classify(unsigned char c) {
  if (c == 'a' || c == 'A') return 0;
  if (c == 'c' || c == 'C') return 1;
  if (c == 'g' || c == 'G') return 2;
  if (c == 't' || c == 'T') return 3;
  return 4;
}

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* a1: extractAlnParam, src/getACGTNsites.cpp:50-85 (counts) and :104-166 (filters).
 * aln: nseq x L bytes, row-major (one record per row).  counts: 5 x L doubles,
 * column-major (counts[a + 5*j]).  pos_out: 1-based retained columns (capacity L).
 * returns number of retained sites. */
EXPORT int64_t ldwo_aln_param(const uint8_t* aln, int64_t nseq, int64_t L, int filter, double gap_thresh,
                              double maf_thresh, double* counts, int32_t* pos_out) {
  memset(counts, 0, sizeof(double) * 5 * (size_t)L);
  for (int64_t s = 0; s < nseq; s++) {
    const uint8_t* row = aln + s * L;
    for (int64_t j = 0; j < L; j++) counts[classify(row[j]) + 5 * j] += 1; /* :58-70 */
  }
  int n = (int)nseq;
  int64_t n_snp = 0;
  if (filter == 0) {
    int min_maf = (int)(n * maf_thresh); /* :105 */
    for (int64_t j = 0; j < L; j++) {
      int chk = 0;
      for (int k = 0; k < 4; k++) {
        if (counts[k + 5 * j] > 0) { /* :113 */
          chk += 1;
          if (chk > 1) {                              /* :115 */
            if (counts[4 + 5 * j] / n < gap_thresh) { /* :118 */
              double snp[4] = {counts[5 * j], counts[1 + 5 * j], counts[2 + 5 * j], counts[3 + 5 * j]};
              qsort(snp, 4, sizeof(double), cmp_double); /* :121 */
              if (snp[2] > min_maf) pos_out[n_snp++] = (int32_t)(j + 1); /* :122-123 */
            }
            break; /* :130 */
          }
        }
      }
    }
  } else {
    int min_maf = (int)(n * (1 - maf_thresh)); /* :136 */
    for (int64_t j = 0; j < L; j++) {
      int chk = 0;
      for (int k = 0; k < 4; k++) {
        if (counts[k + 5 * j] > 0) {
          chk += 1;
          if (chk > 1) {
            if (counts[4 + 5 * j] / n < gap_thresh) { /* :148 */
              double mx = counts[5 * j];
              for (int a = 1; a < 5; a++)
                if (counts[a + 5 * j] > mx) mx = counts[a + 5 * j];
              if (mx <= min_maf) pos_out[n_snp++] = (int32_t)(j + 1); /* :153-154 */
            }
            break;
          }
        }
      }
    }
  }
  return n_snp;
}

/* a2: extractSNPs, src/getACGTNsites.cpp:222-267.  codes: nsnp x nseq (codes[k*nseq + s]);
 * table: 5 x nsnp doubles column-major. */
EXPORT void ldwo_extract_snps(const uint8_t* aln, int64_t nseq, int64_t L, const int32_t* pos, int64_t nsnp,
                              uint8_t* codes, double* table) {
  memset(table, 0, sizeof(double) * 5 * (size_t)nsnp);
  for (int64_t s = 0; s < nseq; s++) {
    const uint8_t* row = aln + s * L;
    for (int64_t k = 0; k < nsnp; k++) {
      int a = classify(row[pos[k] - 1]);
      codes[k * nseq + s] = (uint8_t)a;
      table[a + 5 * k] += 1; /* ++ACGTN_table(a, k-1) */
    }
  }
}

/* a4: ACGTN2num, src/ACGTN2num_parallel.cpp:18-41 (uppercase only; N or '-' -> row 4). */
EXPORT void ldwo_acgtn2num(double* nv, const char* cv, int64_t n) {
  for (int64_t c = 0; c < n; c++) {
    char cc = cv[c];
    if (cc == 'A') nv[c * 5] = 0;
    else if (cc == 'C') nv[c * 5 + 1] = 0;
    else if (cc == 'G') nv[c * 5 + 2] = 0;
    else if (cc == 'T') nv[c * 5 + 3] = 0;
    else if (cc == 'N') nv[c * 5 + 4] = 0;
    else if (cc == '-') nv[c * 5 + 4] = 0;
  }
}

/* a5: estimate_Hamming_distance_weights, R/performPopulationStuctureCorrection.R:23,49-76.
 * shared = sum_a crossprod(M_a) is restated as a per-pair match count (same integers).
 * dist_out may be NULL; else nseq x nseq int32. */
EXPORT void ldwo_hdw(const uint8_t* codes, int64_t nsnp, int64_t nseq, double threshold, int32_t* cnt_out,
                     double* hdw_out, int32_t* dist_out) {
  int thresh = (int)(nsnp * threshold); /* as.integer :23 */
  /* transpose to sequence-major for the inner loop */
  uint8_t* T = (uint8_t*)malloc((size_t)nsnp * nseq);
  for (int64_t k = 0; k < nsnp; k++)
    for (int64_t s = 0; s < nseq; s++) T[s * nsnp + k] = codes[k * nseq + s];
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t s = 0; s < nseq; s++) {
    int cnt = 0;
    for (int64_t t = 0; t < nseq; t++) {
      int64_t shared = 0;
      const uint8_t *a = T + s * nsnp, *b = T + t * nsnp;
      for (int64_t k = 0; k < nsnp; k++) shared += (a[k] == b[k]);
      int64_t d = nsnp - shared;
      if (dist_out) dist_out[s + t * nseq] = (int32_t)d;
      if (d < thresh) cnt++; /* strict '<', includes t == s :76 */
    }
    cnt_out[s] = cnt;
    hdw_out[s] = 1.0 / (cnt + 1);
  }
  free(T);
}

/* a9: fastHadamard, src/computeMI.cpp:11-21 (linear index over nf*nt; RXY indexed linearly too) */
static void fast_hadamard(double* MI, const double* den, const double* uq, const double* pxy, const double* pxpy,
                          const double* RXY, const double* pXrX, const double* pYrY, int64_t n, int ncores) {
#pragma omp parallel for num_threads(ncores)
  for (int64_t c = 0; c < n; c++)
    MI[c] += uq[c] * pxy[c] / den[c] * log(pxy[c] / (pxpy[c] + RXY[c] + pXrX[c] + pYrY[c]) * den[c]);
}

/* a7+a8: one block's MI matrix (nf x nt, column-major), R/computePairwiseMI.R:198-298,390-396.
 * from/to: 0-based global SNP ids.  r: rowSums(uqe); uqe: nsnp x 5 column-major. */
EXPORT int ldwo_block_mi(const uint8_t* codes, int64_t nsnp, int64_t nseq, const double* hdw, const double* r,
                         const double* uqe, const int32_t* from, int64_t nf, const int32_t* to, int64_t nt,
                         int ncores, double* MI) {
  if (ncores < 1) ncores = 1;
  int64_t cells = nf * nt;
  double neff = 0; /* :77 */
  for (int64_t s = 0; s < nseq; s++) neff += hdw[s];
  double* hsq = (double*)malloc(sizeof(double) * nseq); /* diag(sqrt(hdw)) :89 */
  for (int64_t s = 0; s < nseq; s++) hsq[s] = sqrt(hdw[s]);
  int fromISto = (nf == nt); /* :198-202 */
  if (fromISto)
    for (int64_t i = 0; i < nf; i++)
      if (from[i] != to[i]) { fromISto = 0; break; }
  /* weighted dense sub-matrices tXfh (nf x nseq) and CSR of tYth (nt rows), marginals :238-256 */
  double* tfh[5];
  double* pf[5];
  double* pt[5];
  int64_t* csr_ptr[5];
  int32_t* csr_idx[5];
  for (int a = 0; a < 5; a++) {
    tfh[a] = (double*)calloc((size_t)nf * nseq, sizeof(double));
    pf[a] = (double*)calloc(nf, sizeof(double));
    pt[a] = (double*)calloc(nt, sizeof(double));
    csr_ptr[a] = (int64_t*)calloc(nt + 1, sizeof(int64_t));
    for (int64_t i = 0; i < nf; i++) {
      const uint8_t* row = codes + (int64_t)from[i] * nseq;
      double acc = 0;
      for (int64_t s = 0; s < nseq; s++)
        if (row[s] == a) {
          tfh[a][i + s * nf] = hsq[s];
          acc += hsq[s] * hsq[s]; /* rowSums(tAfh^2) */
        }
      pf[a][i] = acc;
    }
    for (int64_t j = 0; j < nt; j++) {
      const uint8_t* row = codes + (int64_t)to[j] * nseq;
      int64_t c = 0;
      double acc = 0;
      for (int64_t s = 0; s < nseq; s++)
        if (row[s] == a) { c++; acc += hsq[s] * hsq[s]; }
      csr_ptr[a][j + 1] = csr_ptr[a][j] + c;
      pt[a][j] = acc;
    }
    csr_idx[a] = (int32_t*)malloc(sizeof(int32_t) * (csr_ptr[a][nt] > 0 ? csr_ptr[a][nt] : 1));
    for (int64_t j = 0; j < nt; j++) {
      const uint8_t* row = codes + (int64_t)to[j] * nseq;
      int64_t w = csr_ptr[a][j];
      for (int64_t s = 0; s < nseq; s++)
        if (row[s] == a) csr_idx[a][w++] = (int32_t)s;
    }
  }
  double* den = (double*)malloc(sizeof(double) * cells);
  double* rft = (double*)malloc(sizeof(double) * cells); /* nt x nf, column-major */
  double* rfh = (double*)malloc(sizeof(double) * nf);
  double* rth = (double*)malloc(sizeof(double) * nt);
  for (int64_t j = 0; j < nt; j++)
    for (int64_t i = 0; i < nf; i++) {
      den[i + j * nf] = neff + r[from[i]] * r[to[j]] * 0.5; /* :260 */
      rft[j + i * nt] = r[from[i]] * r[to[j]] * 0.25;       /* t(tcrossprod(rf, rt))*0.25  :261 */
    }
  for (int64_t i = 0; i < nf; i++) rfh[i] = 0.5 * r[from[i]]; /* :262 */
  for (int64_t j = 0; j < nt; j++) rth[j] = 0.5 * r[to[j]];   /* :263 */
  double* pxy = (double*)malloc(sizeof(double) * cells);
  double* uq = (double*)malloc(sizeof(double) * cells);
  double* pXrX = (double*)malloc(sizeof(double) * cells);
  double* pYrY = (double*)malloc(sizeof(double) * cells);
  double* pxpy = (double*)malloc(sizeof(double) * cells);
  memset(MI, 0, sizeof(double) * cells); /* :268 */
  for (int a = 0; a < 5; a++) {
    for (int b = 0; b < 5; b++) { /* :270-298 */
      /* computeMI_Sprase :390-396 */
#pragma omp parallel for num_threads(ncores) schedule(static)
      for (int64_t j = 0; j < nt; j++) {
        double* col = pxy + j * nf;
        for (int64_t i = 0; i < nf; i++) col[i] = 0;
        for (int64_t q = csr_ptr[b][j]; q < csr_ptr[b][j + 1]; q++) { /* dense x CSR^T  :391 */
          int32_t s = csr_idx[b][q];
          double v = hsq[s];
          const double* xs = tfh[a] + (int64_t)s * nf;
          for (int64_t i = 0; i < nf; i++) col[i] += xs[i] * v;
        }
        double uj = uqe[to[j] + b * nsnp], pyr = pt[b][j] * rth[j], pj = pt[b][j];
        for (int64_t i = 0; i < nf; i++) {
          col[i] += 0.5;
          uq[i + j * nf] = uqe[from[i] + a * nsnp] * uj; /* :392 */
          pXrX[i + j * nf] = pf[a][i] * rfh[i];           /* :393 */
          pYrY[i + j * nf] = pyr;                         /* :394 */
          pxpy[i + j * nf] = pf[a][i] * pj;               /* :395 */
        }
      }
      fast_hadamard(MI, den, uq, pxy, pxpy, rft, pXrX, pYrY, cells, ncores); /* :396 */
    }
  }
  for (int a = 0; a < 5; a++) { free(tfh[a]); free(pf[a]); free(pt[a]); free(csr_ptr[a]); free(csr_idx[a]); }
  free(den); free(rft); free(rfh); free(rth); free(pxy); free(uq); free(pXrX); free(pYrY); free(pxpy); free(hsq);
  (void)fromISto;
  return 0;
}

static double fmod_floor(double a, double g) { /* R's %% */
  double m = fmod(a, g);
  if (m != 0 && ((m < 0) != (g < 0))) m += g;
  return m;
}

/* stats::quantile type 7 on a scratch copy (quirk Q3) */
static double quantile7(double* x, int64_t n, double prob) {
  double index = 1 + (double)(n - 1 > 0 ? n - 1 : 0) * prob;
  int64_t lo = (int64_t)floor(index), hi = (int64_t)ceil(index);
  qsort(x, n, sizeof(double), cmp_double);
  double qs = x[lo - 1];
  if (index > lo && x[hi - 1] != qs) {
    double h = index - lo;
    qs = (1 - h) * qs + h * x[hi - 1];
  }
  return qs;
}

/* a10: one block's links, R/computePairwiseMI.R:306-364.  MI: nf x nt from ldwo_block_mi.
 * Outputs (capacity nf*nt each): row/col local 0-based indices in reference row order, len, mi,
 * is_sr flag, lr_keep flag.  Returns number of rows; *thr_out = disc_thresh (NaN if no LR branch). */
EXPORT int64_t ldwo_block_links(const double* MI, const double* POS, const int32_t* from, int64_t nf,
                                const int32_t* to, int64_t nt, double g, double sr_dist, double lr_retain_links,
                                double lr_links_approx, int sr_only, int32_t* row_out, int32_t* col_out,
                                double* len_out, double* mi_out, uint8_t* is_sr, uint8_t* lr_keep, double* thr_out,
                                double* prob_out) {
  int fromISto = (nf == nt);
  if (fromISto)
    for (int64_t i = 0; i < nf; i++)
      if (from[i] != to[i]) { fromISto = 0; break; }
  int64_t n = 0;
  if (fromISto) { /* :307 lower.tri(t(MI)): row > col, column-major */
    for (int64_t c = 0; c < nt; c++)
      for (int64_t rr = c + 1; rr < nf; rr++) { row_out[n] = (int32_t)rr; col_out[n] = (int32_t)c; n++; }
  } else { /* :309 upper.tri then lower.tri, each column-major */
    for (int64_t c = 0; c < nt; c++)
      for (int64_t rr = 0; rr < nf && rr < c; rr++) { row_out[n] = (int32_t)rr; col_out[n] = (int32_t)c; n++; }
    for (int64_t c = 0; c < nt; c++)
      for (int64_t rr = c + 1; rr < nf; rr++) { row_out[n] = (int32_t)rr; col_out[n] = (int32_t)c; n++; }
  }
  int64_t n_lr = 0;
  for (int64_t k = 0; k < n; k++) {
    double pos2 = POS[from[row_out[k]]], pos1 = POS[to[col_out[k]]]; /* :319-320 */
    double ln = 0.5 * g - fabs(fmod_floor(pos1 - pos2, g) - 0.5 * g); /* :330 */
    len_out[k] = ln;
    mi_out[k] = MI[row_out[k] + (int64_t)col_out[k] * nf];
    is_sr[k] = (ln <= sr_dist); /* :333 */
    lr_keep[k] = 0;
    if (!is_sr[k]) n_lr++;
  }
  *thr_out = NAN;
  *prob_out = NAN;
  if (n_lr > 0 && !sr_only) { /* :347 */
    double prob = 1 - ((lr_retain_links * ((double)n_lr / lr_links_approx)) / (double)n_lr); /* :352 */
    if (prob < 0) prob = 0;
    double* tmp = (double*)malloc(sizeof(double) * n_lr);
    int64_t w = 0;
    for (int64_t k = 0; k < n; k++)
      if (!is_sr[k]) tmp[w++] = mi_out[k];
    double thr = quantile7(tmp, n_lr, prob); /* :354 */
    free(tmp);
    for (int64_t k = 0; k < n; k++)
      if (!is_sr[k] && mi_out[k] >= thr) lr_keep[k] = 1; /* :358 */
    *thr_out = thr;
    *prob_out = prob;
  }
  return n;
}

EXPORT int ldwo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
