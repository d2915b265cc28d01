// Host-side writer of lr_links.tsv: the rows perform_MI_computation_ACGTN appends per block with
//   write.table(x = MI_df_lr, file = lr_save_path, append = T, quote = F, row.names = F, col.names = F, sep = '\t')
// (R/computePairwiseMI.R:362; columns pos1 pos2 clust1 clust2 len MI, :326-331; reader R/io_functions.R:34-35).
// write.table encodes every cell on its own (utils:::writetable -> EncodeElement0): integers as plain digits, doubles
// with up to 15 significant digits, in fixed notation unless scientific notation is strictly narrower (formatReal with
// R_print.digits = DBL_DIG, scipen = 0).  Every column of MI_df is a double: pos1 / pos2 come from
// POS_f = as.numeric(snp.dat$POS[from]) (R/computePairwiseMI.R:176-177), paint is built by rep(0, n)
// (R/estimateCDSDiversity.R:152) -- so a position or a len of 100000 is written "1e+05", exactly as R does.  Only clust_c
// of sr_links.tsv is an integer column (the loop index of :464-467).
// Rows are formatted on a few host threads and written in order.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <charconv>
#include <string>
#include <vector>

#include "host_util.h"
#include "../../include/ldw.h"

namespace {

// One double the way write.table prints it.  Returns the number of characters written (no terminator needed).
// The 15 correctly rounded significant digits come from std::to_chars (same rounding of the exact binary value as
// printf's "%.14e", several times faster); both of R's candidate forms are then laid out from those digits: the fixed form
// sprintf("%.*f", rgt, x) shows exactly the first nsig significant digits (digits nsig+1..15 are zeros after the 15-digit
// rounding, so rounding x directly to rgt decimals cannot differ), the scientific one "%.*e" with nsig - 1 decimals.
int format_r_real(double x, char* out) {
  if (isnan(x)) { memcpy(out, "NA", 2); return 2; }
  if (isinf(x)) { if (x > 0) { memcpy(out, "Inf", 3); return 3; } memcpy(out, "-Inf", 4); return 4; }
  if (x == 0.0) { out[0] = '0'; return 1; }
  const int neg = x < 0;
  const double ax = fabs(x);
  char e15[48];
  const std::to_chars_result tr = std::to_chars(e15, e15 + sizeof(e15), ax, std::chars_format::scientific, 14);
  *tr.ptr = 0;  // d.dddddddddddddde[+-]XX[X]
  char dig[16];
  dig[0] = e15[0];
  memcpy(dig + 1, e15 + 2, 14);
  int nsig = 15;
  while (nsig > 1 && dig[nsig - 1] == '0') nsig--;
  const int kpower = atoi(e15 + 17);
  int left, rgt;
  if (kpower >= 0) {
    left = kpower + 1;
    rgt = nsig - kpower - 1;
    if (rgt < 0) rgt = 0;
    if (kpower > 0 && kpower <= 22 && ax < pow(10.0, kpower)) left--;  // formatReal's `roundingwidens`
  }
  else { left = 1; rgt = nsig - kpower - 1; }
  const int wF = neg + left + (rgt ? rgt + 1 : 0);
  const int wE = neg + (nsig > 1 ? nsig + 1 : 1) + ((kpower >= 100 || kpower <= -100) ? 5 : 4);
  if (wF <= wE && kpower >= 15) return snprintf(out, 400, "%.*f", rgt, x);  // more integer digits than the 15 at hand: printf's exact expansion
  int o = 0;
  if (neg) out[o++] = '-';
  if (wF <= wE) {  // fixed notation
    if (kpower >= 0) {
      for (int k = 0; k <= kpower; k++) out[o++] = k < 15 ? dig[k] : '0';
      if (rgt > 0) {
        out[o++] = '.';
        for (int k = 0; k < rgt; k++) out[o++] = (kpower + 1 + k) < 15 ? dig[kpower + 1 + k] : '0';
      }
    } else {
      out[o++] = '0';
      out[o++] = '.';
      for (int k = 0; k < -kpower - 1; k++) out[o++] = '0';
      memcpy(out + o, dig, nsig);
      o += nsig;
    }
    return o;
  }
  out[o++] = dig[0];
  if (nsig > 1) { out[o++] = '.'; memcpy(out + o, dig + 1, nsig - 1); o += nsig - 1; }
  out[o++] = 'e';
  out[o++] = kpower < 0 ? '-' : '+';
  const int ae = kpower < 0 ? -kpower : kpower;
  if (ae >= 100) out[o++] = (char)('0' + ae / 100);
  out[o++] = (char)('0' + (ae / 10) % 10);
  out[o++] = (char)('0' + ae % 10);
  return o;
}

int format_int(int64_t v, char* out) {
  char tmp[24];
  int n = 0, o = 0;
  uint64_t u = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
  do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  if (v < 0) out[o++] = '-';
  while (n) out[o++] = tmp[--n];
  return o;
}

// Whole-valued doubles below 10^15 (clust, len): same rule as format_r_real without going through printf.
int format_r_whole(int64_t v, char* out) {
  if (v == 0) { out[0] = '0'; return 1; }
  char dig[24];
  const int neg = v < 0;
  const int nd = format_int(neg ? -v : v, dig);
  int nsig = nd;
  while (nsig > 1 && dig[nsig - 1] == '0') nsig--;
  const int kpower = nd - 1;
  const int wF = neg + nd, wE = neg + (nsig > 1 ? nsig + 1 : 1) + 4;
  int o = 0;
  if (neg) out[o++] = '-';
  if (wF <= wE) { memcpy(out + o, dig, nd); return o + nd; }
  out[o++] = dig[0];
  if (nsig > 1) { out[o++] = '.'; memcpy(out + o, dig + 1, nsig - 1); o += nsig - 1; }
  out[o++] = 'e'; out[o++] = '+';
  out[o++] = (char)('0' + kpower / 10); out[o++] = (char)('0' + kpower % 10);
  return o;
}

}  // namespace

extern "C" int ldw_format_r_real(double x, char* out, int cap) {
  return ldw::guarded("ldw_format_r_real", [&]() -> int {
  char buf[512];
  int n = format_r_real(x, buf);
  if (!out || cap <= n) return ldw::set_error(LDW_ERR_ARG, "ldw_format_r_real: buffer too small");
  memcpy(out, buf, n);
  out[n] = 0;
  return 0;
  });
}

namespace {
// Rows are formatted chunk by chunk on a few host threads and written in order.
template <class RowFn>
int write_rows(const char* who, const char* path, int append, int64_t n, RowFn row) {
  FILE* f = fopen(path, append ? "ab" : "wb");
  if (!f) return ldw::set_error(LDW_ERR_ARG, "%s: can't open %s", who, path);
  const int64_t chunk = 1 << 13;  // small enough that a few 10^5 rows (sr_links.tsv) already keep every thread busy
  const int64_t nchunks = (n + chunk - 1) / chunk;
  const int64_t wave = 256;  // chunks formatted concurrently, then written in order
  int rc = 0;
  for (int64_t c0 = 0; c0 < nchunks && rc == 0; c0 += wave) {
    const int64_t nc = std::min<int64_t>(wave, nchunks - c0);
    std::vector<std::string> bufs(nc);
    ldw::parallel_for(nc, 32, [&](int64_t k) {
      std::string& s = bufs[k];
      const int64_t lo = (c0 + k) * chunk, hi = std::min<int64_t>(n, lo + chunk);
      s.reserve((size_t)(hi - lo) * 96);
      for (int64_t i = lo; i < hi; i++) row(i, s);
    });
    for (int64_t k = 0; k < nc; k++)
      if (fwrite(bufs[k].data(), 1, bufs[k].size(), f) != bufs[k].size()) { rc = ldw::set_error(LDW_ERR_ARG, "%s: write to %s failed", who, path); break; }
  }
  if (fclose(f) != 0 && rc == 0) rc = ldw::set_error(LDW_ERR_ARG, "%s: closing %s failed", who, path);
  return rc;
}
}  // namespace

extern "C" int ldw_write_lr_tsv(const char* path, const ldw_links* lr, int append) {
  return ldw::guarded("ldw_write_lr_tsv", [&]() -> int {
  if (!path || !lr) return ldw::set_error(LDW_ERR_ARG, "ldw_write_lr_tsv: null argument");
  return write_rows("ldw_write_lr_tsv", path, append, lr->n, [&](int64_t i, std::string& s) {
    char tmp[512];
    s.append(tmp, format_r_whole(lr->pos1[i], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(lr->pos2[i], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(lr->clust1[i], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(lr->clust2[i], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(lr->len[i], tmp)); s.push_back('\t');
    s.append(tmp, format_r_real(lr->MI[i], tmp)); s.push_back('\n');
  });
  });
}

// sr_links.tsv (R/computePairwiseMI.R:140): clust_c is the integer loop index of mergeNsort_sr_links (:411,470), pos1 / pos2
// integer, clust1 / clust2 / len / MI / srp_max doubles, ARACNE = as.numeric(logical) (:126) or the constant 1 (:129).
extern "C" int ldw_write_sr_tsv(const char* path, const ldw_links* sr, int64_t n, const int64_t* rows, const int32_t* clust_c,
                                const double* srp_max, const double* aracne, int append) {
  return ldw::guarded("ldw_write_sr_tsv", [&]() -> int {
  if (!path || !sr || (n > 0 && (!rows || !clust_c || !srp_max || !aracne))) return ldw::set_error(LDW_ERR_ARG, "ldw_write_sr_tsv: null argument");
  for (int64_t i = 0; i < n; i++)
    if (rows[i] < 0 || rows[i] >= sr->n) return ldw::set_error(LDW_ERR_ARG, "ldw_write_sr_tsv: row %lld outside the link table", (long long)rows[i]);
  return write_rows("ldw_write_sr_tsv", path, append, n, [&](int64_t i, std::string& s) {
    char tmp[512];
    const int64_t r = rows[i];
    s.append(tmp, format_int(clust_c[i], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(sr->pos1[r], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(sr->pos2[r], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(sr->clust1[r], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(sr->clust2[r], tmp)); s.push_back('\t');
    s.append(tmp, format_r_whole(sr->len[r], tmp)); s.push_back('\t');
    s.append(tmp, format_r_real(sr->MI[r], tmp)); s.push_back('\t');
    s.append(tmp, format_r_real(srp_max[i], tmp)); s.push_back('\t');
    const double ar = aracne[i];  // 0 / 1 (as.numeric(logical)): no need to go through printf
    s.append(tmp, (ar == 0.0 || ar == 1.0) ? format_r_whole((int64_t)ar, tmp) : format_r_real(ar, tmp)); s.push_back('\n');
  });
  });
}

// ---------------------------------------------------------------------------------------------------------------------
// Reader of the two link files: read.table(path, sep = "\t", header = F, quote = "", comment.char = "") as
// read_LongRangeLinks / read_ShortRangeLinks call it (R/io_functions.R:34,62): every column of these files is numeric,
// so the table comes back as column-major doubles ("NA" -> NaN).  The file is cut at line ends into chunks that are
// counted and then parsed on host threads.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" void ldw_table_free(double* cols) { free(cols); }

extern "C" int ldw_read_numeric_tsv(const char* path, int ncols, int64_t* nrows_out, double** cols_out) {
  return ldw::guarded("ldw_read_numeric_tsv", [&]() -> int {
  if (!path || !nrows_out || !cols_out || ncols < 1) return ldw::set_error(LDW_ERR_ARG, "ldw_read_numeric_tsv: bad argument");
  *nrows_out = 0;
  *cols_out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return ldw::set_error(LDW_ERR_ARG, "ldw_read_numeric_tsv: can't open %s", path);
  std::string buf;
  {
    char tmp[1 << 16];
    size_t k;
    while ((k = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.append(tmp, k);
  }
  fclose(f);
  if (!buf.empty() && buf.back() != '\n') buf.push_back('\n');
  const size_t len = buf.size();
  if (len == 0) return 0;
  const size_t target = 1 << 22;
  std::vector<size_t> cut(1, 0);
  while (cut.back() < len) {
    size_t e = std::min(len, cut.back() + target);
    while (e < len && buf[e - 1] != '\n') e++;
    cut.push_back(e);
  }
  const int64_t nchunks = (int64_t)cut.size() - 1;
  std::vector<int64_t> first(nchunks + 1, 0);
  ldw::parallel_for(nchunks, 8, [&](int64_t c) {
    int64_t n = 0;
    for (size_t i = cut[c]; i < cut[c + 1]; i++) n += buf[i] == '\n';
    first[c + 1] = n;
  });
  for (int64_t c = 0; c < nchunks; c++) first[c + 1] += first[c];
  const int64_t nrows = first[nchunks];
  double* cols = (double*)malloc(sizeof(double) * (size_t)std::max<int64_t>(1, nrows) * ncols);
  if (!cols) return ldw::set_error(LDW_ERR_NOMEM, "ldw_read_numeric_tsv: out of memory (%lld rows)", (long long)nrows);
  std::vector<int64_t> bad(nchunks, -1);
  ldw::parallel_for(nchunks, 8, [&](int64_t c) {
    const char* p = buf.data() + cut[c];
    const char* end = buf.data() + cut[c + 1];
    int64_t row = first[c];
    while (p < end) {
      for (int k = 0; k < ncols; k++) {
        double v;
        if (p[0] == 'N' && p[1] == 'A' && (p[2] == '\t' || p[2] == '\n' || p[2] == '\r')) { v = NAN; p += 2; }
        else {
          char* q;
          v = strtod(p, &q);
          if (q == p) { if (bad[c] < 0) bad[c] = row; }
          p = q;
        }
        cols[(size_t)k * nrows + row] = v;
        const char sep = (k + 1 < ncols) ? '\t' : '\n';
        if (*p == '\r' && sep == '\n') p++;
        if (*p != sep) {  // wrong number of fields (or text in a field): skip to the end of the line, report it
          if (bad[c] < 0) bad[c] = row;
          while (*p != '\n') p++;
          for (int j = k + 1; j < ncols; j++) cols[(size_t)j * nrows + row] = NAN;
          k = ncols;
        }
        p++;
      }
      row++;
    }
  });
  for (int64_t c = 0; c < nchunks; c++)
    if (bad[c] >= 0) {
      free(cols);
      return ldw::set_error(LDW_ERR_ARG, "ldw_read_numeric_tsv: line %lld of %s did not have %d numeric fields", (long long)bad[c] + 1, path, ncols);
    }
  *nrows_out = nrows;
  *cols_out = cols;
  return 0;
  });
}
