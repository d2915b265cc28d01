// Minimal stand-in for <Rcpp.h> (test infrastructure, NOT the product path).
//
// Purpose: let the reference's own C++ sources
//     /root/reference/src/{getACGTNsites.cpp, computeMI.cpp, ACGTN2num_parallel.cpp, fintersect.cpp}
// compile UNMODIFIED, from where they lie, into oracle/_ref/libldw_ref.so (recipe: oracle/Makefile, target `ref`),
// so that the oracle restatements and the CUDA path can be checked against the reference's object code.
// Only the handful of Rcpp types/functions those four files use are provided, with R's semantics where the
// sources depend on them: column-major matrices, shallow (shared) copies of vectors, named lists.
// Nothing here is taken from Rcpp's sources.
#ifndef LDW_MOCK_RCPP_H
#define LDW_MOCK_RCPP_H

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace Rcpp {

// ---------------------------------------------------------------- Rcout: swallowed (the R console is not here)
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
};
static NullStream Rcout;

struct Placeholder {};
static const Placeholder _ = Placeholder();

// ---------------------------------------------------------------- vectors share storage on copy, like SEXP handles
template <typename T> class Vec {
 public:
  Vec() : own_(std::make_shared<std::vector<T>>()), ext_(nullptr), n_(0) {}
  explicit Vec(long n) : own_(std::make_shared<std::vector<T>>(static_cast<size_t>(n), T())), ext_(nullptr), n_(n) {}
  Vec(long n, T fill) : own_(std::make_shared<std::vector<T>>(static_cast<size_t>(n), fill)), ext_(nullptr), n_(n) {}
  Vec(const std::vector<T>& v) : own_(std::make_shared<std::vector<T>>(v)), ext_(nullptr), n_((long)v.size()) {}
  // view of caller memory (what an R vector passed through .Call is): no copy, writes go through
  static Vec view(T* p, long n) { Vec v; v.own_.reset(); v.ext_ = p; v.n_ = n; return v; }
  static Vec create(T a) { return Vec(std::vector<T>{a}); }
  static Vec create(T a, T b) { return Vec(std::vector<T>{a, b}); }
  static Vec create(T a, T b, T c) { return Vec(std::vector<T>{a, b, c}); }
  static Vec create(T a, T b, T c, T d) { return Vec(std::vector<T>{a, b, c, d}); }
  static Vec create(T a, T b, T c, T d, T e) { return Vec(std::vector<T>{a, b, c, d, e}); }

  T* data() { return ext_ ? ext_ : own_->data(); }
  const T* data() const { return ext_ ? ext_ : own_->data(); }
  long length() const { return n_; }
  long size() const { return n_; }
  T& operator[](long i) { return data()[i]; }
  const T& operator[](long i) const { return data()[i]; }
  T* begin() { return data(); }
  T* end() { return data() + n_; }
  const T* begin() const { return data(); }
  const T* end() const { return data() + n_; }
  // x[IntegerVector] -> new vector (0-based indices, as Rcpp subsetting)
  template <typename I> Vec operator[](const Vec<I>& idx) const {
    Vec out(idx.length());
    for (long k = 0; k < idx.length(); ++k) out[k] = (*this)[(long)idx[k]];
    return out;
  }
  void push_back(const T& x) {
    own_->push_back(x);
    n_ = (long)own_->size();
  }
  std::vector<T> to_std() const { return std::vector<T>(begin(), end()); }

 protected:
  std::shared_ptr<std::vector<T>> own_;
  T* ext_;
  long n_;
};

// LogicalVector needs addressable elements (std::vector<bool> has none): store int like R does
typedef Vec<double> NumericVector;
typedef Vec<int> IntegerVector;
typedef Vec<int> LogicalVector;

template <typename T> inline T max(const Vec<T>& v) {
  T m = v[0];
  for (long i = 1; i < v.length(); ++i)
    if (v[i] > m) m = v[i];
  return m;
}

// ---------------------------------------------------------------- column-major numeric matrix
class NumericMatrix {
 public:
  NumericMatrix() : nr_(0), nc_(0) {}
  NumericMatrix(int nr, int nc) : v_((long)nr * (long)nc), nr_(nr), nc_(nc) {}
  static NumericMatrix view(double* p, int nr, int nc) {
    NumericMatrix m;
    m.v_ = NumericVector::view(p, (long)nr * (long)nc);
    m.nr_ = nr;
    m.nc_ = nc;
    return m;
  }
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
  double& operator()(int i, int j) { return v_[(long)i + (long)j * nr_]; }
  const double& operator()(int i, int j) const { return v_[(long)i + (long)j * nr_]; }
  double& operator[](long k) { return v_[k]; }
  const double& operator[](long k) const { return v_[k]; }
  // m(_, j): the column as a vector (a copy is enough: the sources only read it)
  NumericVector operator()(Placeholder, int j) const {
    NumericVector c(nr_);
    for (int i = 0; i < nr_; ++i) c[i] = (*this)(i, j);
    return c;
  }
  const double* data() const { return v_.data(); }
  long length() const { return v_.length(); }

 private:
  NumericVector v_;
  int nr_, nc_;
};

// ---------------------------------------------------------------- strings
class String {
 public:
  String() {}
  String(const char* s) : s_(s ? s : "") {}
  String(const std::string& s) : s_(s) {}
  const std::string& str() const { return s_; }
  const char* get_cstring() const { return s_.c_str(); }

 private:
  std::string s_;
};

class StringVector {
 public:
  // element proxy: assignable from char*, readable by as<char>()
  struct Elem {
    std::string* p;
    Elem& operator=(const char* s) { *p = s ? s : ""; return *this; }
    Elem& operator=(const std::string& s) { *p = s; return *this; }
  };
  StringVector() : v_(std::make_shared<std::vector<std::string>>()) {}
  explicit StringVector(long n) : v_(std::make_shared<std::vector<std::string>>((size_t)n)) {}
  void push_back(const char* s) { v_->push_back(s ? s : ""); }
  void push_back(const std::string& s) { v_->push_back(s); }
  long length() const { return (long)v_->size(); }
  long size() const { return (long)v_->size(); }
  Elem operator[](long i) { return Elem{&(*v_)[(size_t)i]}; }
  const std::vector<std::string>& to_std() const { return *v_; }

 private:
  std::shared_ptr<std::vector<std::string>> v_;
};
typedef StringVector CharacterVector;

template <typename T> T as(const StringVector::Elem& e);
// as<char>(CHARSXP): the first character of the string ('\0' for an empty one)
template <> inline char as<char>(const StringVector::Elem& e) { return e.p->empty() ? '\0' : (*e.p)[0]; }

// ---------------------------------------------------------------- named list
struct Value {
  enum Kind { NONE, INT, REAL, INTVEC, REALVEC, STRVEC, MATRIX, STR } kind;
  int i;
  double d;
  std::vector<int> iv;
  std::vector<double> dv;
  std::vector<std::string> sv;
  std::string s;
  int nrow, ncol;
  Value() : kind(NONE), i(0), d(0), nrow(0), ncol(0) {}
  Value(int x) : kind(INT), i(x), d(0), nrow(0), ncol(0) {}
  Value(double x) : kind(REAL), i(0), d(x), nrow(0), ncol(0) {}
  Value(const std::vector<int>& x) : kind(INTVEC), i(0), d(0), iv(x), nrow(0), ncol(0) {}
  Value(const IntegerVector& x) : kind(INTVEC), i(0), d(0), iv(x.to_std()), nrow(0), ncol(0) {}
  Value(const NumericVector& x) : kind(REALVEC), i(0), d(0), dv(x.to_std()), nrow(0), ncol(0) {}
  Value(const StringVector& x) : kind(STRVEC), i(0), d(0), sv(x.to_std()), nrow(0), ncol(0) {}
  Value(const String& x) : kind(STR), i(0), d(0), s(x.str()), nrow(0), ncol(0) {}
  Value(const NumericMatrix& m)
      : kind(MATRIX), i(0), d(0), dv(m.data(), m.data() + m.length()), nrow(m.nrow()), ncol(m.ncol()) {}
};

template <typename T> inline Value wrap(const T& x) { return Value(x); }

struct NamedValue {
  std::string name;
  Value value;
};
struct Named {
  std::string name;
  explicit Named(const char* n) : name(n) {}
  template <typename T> NamedValue operator=(const T& x) const { return NamedValue{name, Value(x)}; }
  NamedValue operator=(const Value& x) const { return NamedValue{name, x}; }
};

class List {
 public:
  template <typename... A> static List create(const A&... a) {
    List l;
    (l.items_.push_back(a), ...);
    return l;
  }
  const Value* find(const char* name) const {
    for (const NamedValue& nv : items_)
      if (nv.name == name) return &nv.value;
    return nullptr;
  }
  size_t size() const { return items_.size(); }

 private:
  std::vector<NamedValue> items_;
};

}  // namespace Rcpp

#endif
