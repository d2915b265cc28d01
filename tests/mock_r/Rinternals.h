// Minimal declarations of the R C API used by r_package/src/r_shim.cpp -- a MOCK for a syntax / type check only (R is not
// installed in the build image); signatures follow R's public Rinternals.h.
#pragma once
#include <stddef.h>
typedef struct SEXPREC* SEXP; typedef ptrdiff_t R_xlen_t; typedef unsigned char Rbyte;
enum { INTSXP=13, REALSXP=14, LGLSXP=10, STRSXP=16, VECSXP=19, RAWSXP=24 };
extern "C" { SEXP Rf_allocVector(int, R_xlen_t); SEXP Rf_protect(SEXP); void Rf_unprotect(int); double Rf_asReal(SEXP); int Rf_asInteger(SEXP); int Rf_asLogical(SEXP);
int* INTEGER(SEXP); double* REAL(SEXP); int* LOGICAL(SEXP); Rbyte* RAW(SEXP); R_xlen_t XLENGTH(SEXP); SEXP STRING_ELT(SEXP, R_xlen_t); const char* CHAR(SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); SEXP VECTOR_ELT(SEXP, R_xlen_t); SEXP Rf_mkChar(const char*); SEXP Rf_setAttrib(SEXP, SEXP, SEXP); SEXP Rf_ScalarInteger(int); SEXP Rf_ScalarReal(double);
void Rf_error(const char*, ...); extern SEXP R_NamesSymbol; extern SEXP R_NilValue; SEXP Rf_allocMatrix(int,int,int); SEXP Rf_ScalarLogical(int); void R_CheckUserInterrupt(void); char* R_alloc(size_t, int);
__attribute__((noreturn)) void Rf_errorcall(SEXP, const char*, ...);}
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
