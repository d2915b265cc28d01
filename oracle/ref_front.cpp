// C front of oracle/_ref/libldw_ref.so (test infrastructure, NOT the product path).
//
// The library is the reference's OWN C++ -- src/getACGTNsites.cpp, src/computeMI.cpp, src/ACGTN2num_parallel.cpp,
// src/fintersect.cpp and the vendored src/kseq2.h -- compiled unmodified from /root/reference against the stand-in
// oracle/mock_rcpp/Rcpp.h (recipe: oracle/Makefile, target `ref`).  This file only adds plain-C entry points so that
// ctypes (oracle/ref_lib.py) can call that object code; it restates nothing.
#include <Rcpp.h>
#include <zlib.h>

#include "kseq2.h"  // -I/root/reference/src : the reference's reader, used as is
KSEQ_INIT(gzFile, gzread)

using namespace Rcpp;

// the reference functions, as their definitions declare them
List extractAlnParam(std::string file, int filter, double gap_thresh, double maf_thresh);  // getACGTNsites.cpp:13
List extractSNPs(std::string file, int n_seq, int n_snp, std::vector<int> POS);             // getACGTNsites.cpp:179
List extractRef(std::string file);                                                          // getACGTNsites.cpp:295
void ACGTN2num(NumericMatrix nv, StringVector cv, int ncores);                              // ACGTN2num_parallel.cpp:10
void fastHadamard(NumericMatrix MIt, NumericMatrix den, NumericMatrix uq_t, NumericMatrix pxy_t, NumericMatrix pxpy_t,
                  NumericMatrix RXY, NumericMatrix pXrX, NumericMatrix pYrY, int ncores);   // computeMI.cpp:11
LogicalVector compareToRow(NumericMatrix x, NumericVector y);                               // computeMI.cpp:25
NumericVector vecPosMatch(NumericVector x, NumericVector y);                                // computeMI.cpp:44
bool compareTriplet(NumericVector MI0X, NumericVector MI0Z, double MI0);                    // computeMI.cpp:63
std::vector<int> fast_intersect(std::vector<int> A, std::vector<int> B);                    // fintersect.cpp:6

#define API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------- generic access to a returned List
API void ldwref_list_free(void* h) { delete static_cast<List*>(h); }
API int ldwref_list_has(void* h, const char* name) { return static_cast<List*>(h)->find(name) != nullptr; }
API int ldwref_list_int(void* h, const char* name, int* out) {
  const Value* v = static_cast<List*>(h)->find(name);
  if (!v || v->kind != Value::INT) return -1;
  *out = v->i;
  return 0;
}
API long ldwref_list_intvec(void* h, const char* name, const int** out) {
  const Value* v = static_cast<List*>(h)->find(name);
  if (!v || v->kind != Value::INTVEC) return -1;
  *out = v->iv.data();
  return (long)v->iv.size();
}
API long ldwref_list_matrix(void* h, const char* name, const double** out, int* nrow, int* ncol) {
  const Value* v = static_cast<List*>(h)->find(name);
  if (!v || v->kind != Value::MATRIX) return -1;
  *out = v->dv.data();
  *nrow = v->nrow;
  *ncol = v->ncol;
  return (long)v->dv.size();
}
API long ldwref_list_strvec_len(void* h, const char* name) {
  const Value* v = static_cast<List*>(h)->find(name);
  if (!v || v->kind != Value::STRVEC) return -1;
  return (long)v->sv.size();
}
API const char* ldwref_list_strvec_get(void* h, const char* name, long i) {
  const Value* v = static_cast<List*>(h)->find(name);
  if (!v || v->kind != Value::STRVEC || i < 0 || i >= (long)v->sv.size()) return nullptr;
  return v->sv[(size_t)i].c_str();
}
API const char* ldwref_list_str(void* h, const char* name, long* len) {
  const Value* v = static_cast<List*>(h)->find(name);
  if (!v || v->kind != Value::STR) return nullptr;
  *len = (long)v->s.size();
  return v->s.c_str();
}

// ---------------------------------------------------------------------------------- the exported reference functions
API void* ldwref_extractAlnParam(const char* file, int filter, double gap_thresh, double maf_thresh) {
  return new List(extractAlnParam(file, filter, gap_thresh, maf_thresh));
}
API void* ldwref_extractSNPs(const char* file, int n_seq, int n_snp, const int* pos, long npos) {
  return new List(extractSNPs(file, n_seq, n_snp, std::vector<int>(pos, pos + npos)));
}
API void* ldwref_extractRef(const char* file) { return new List(extractRef(file)); }

// nv: 5 x n column-major doubles, modified in place (quirk Q11); cv: n one-character strings given as n chars
API void ldwref_ACGTN2num(double* nv, const char* cv, long n, int ncores) {
  StringVector sv(n);
  for (long i = 0; i < n; ++i) sv[i] = std::string(1, cv[i]);
  ACGTN2num(NumericMatrix::view(nv, 5, (int)n), sv, ncores);
}

// every matrix column-major with its own shape; MIt is modified in place (quirk Q11).  RXY may be nt x nf (quirk Q1):
// the reference reads all eight by the same linear index, so only the element counts matter.
API void ldwref_fastHadamard(double* MIt, int nrow, int ncol, double* den, double* uq_t, double* pxy_t, double* pxpy_t,
                             double* RXY, int rxy_nrow, int rxy_ncol, double* pXrX, double* pYrY, int ncores) {
  fastHadamard(NumericMatrix::view(MIt, nrow, ncol), NumericMatrix::view(den, nrow, ncol),
               NumericMatrix::view(uq_t, nrow, ncol), NumericMatrix::view(pxy_t, nrow, ncol),
               NumericMatrix::view(pxpy_t, nrow, ncol), NumericMatrix::view(RXY, rxy_nrow, rxy_ncol),
               NumericMatrix::view(pXrX, nrow, ncol), NumericMatrix::view(pYrY, nrow, ncol), ncores);
}

API void ldwref_compareToRow(double* x, int nr, int nc, double* y, long ny, int* out) {
  LogicalVector r = compareToRow(NumericMatrix::view(x, nr, nc), NumericVector::view(y, ny));
  for (int j = 0; j < nr; ++j) out[j] = r[j];
}
API void ldwref_vecPosMatch(double* x, long nx, double* y, long ny, double* out) {
  NumericVector r = vecPosMatch(NumericVector::view(x, nx), NumericVector::view(y, ny));
  for (long i = 0; i < nx; ++i) out[i] = r[i];
}
API int ldwref_compareTriplet(double* a, double* b, long n, double mi0) {
  return compareTriplet(NumericVector::view(a, n), NumericVector::view(b, n), mi0) ? 1 : 0;
}
API long ldwref_fast_intersect(const int* a, long na, const int* b, long nb, int* out) {
  std::vector<int> r = fast_intersect(std::vector<int>(a, a + na), std::vector<int>(b, b + nb));
  for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
  return (long)r.size();
}

// ---------------------------------------------------------------------------------- kseq_read itself, record by record
// Runs the loop every reference reader runs (`while ((l = kseq_read(seq)) >= 0)`, getACGTNsites.cpp:50,222) and keeps
// what each call left in seq->name / seq->seq, so a tokeniser can be compared with src/kseq2.h:167-207 byte for byte.
struct KseqDump {
  std::vector<std::string> names, seqs;  // seqs hold seq.l bytes (embedded NULs included)
  int last_rc;                           // the return value that ended the loop (-1 EOF, -2 truncated quality)
};
API void* ldwref_kseq_read_all(const char* file) {
  gzFile fp = gzopen(file, "r");
  if (!fp) return nullptr;
  kseq_t* seq = kseq_init(fp);
  KseqDump* d = new KseqDump();
  int l;
  while ((l = kseq_read(seq)) >= 0) {
    d->names.push_back(std::string(seq->name.s ? seq->name.s : "", seq->name.l));
    d->seqs.push_back(std::string(seq->seq.s ? seq->seq.s : "", seq->seq.l));
  }
  d->last_rc = l;
  kseq_destroy(seq);
  gzclose(fp);
  return d;
}
API long ldwref_kseq_count(void* h) { return (long)static_cast<KseqDump*>(h)->names.size(); }
API int ldwref_kseq_last_rc(void* h) { return static_cast<KseqDump*>(h)->last_rc; }
API const char* ldwref_kseq_name(void* h, long i) { return static_cast<KseqDump*>(h)->names[(size_t)i].c_str(); }
API const char* ldwref_kseq_seq(void* h, long i, long* len) {
  const std::string& s = static_cast<KseqDump*>(h)->seqs[(size_t)i];
  *len = (long)s.size();
  return s.data();
}
API void ldwref_kseq_free(void* h) { delete static_cast<KseqDump*>(h); }
