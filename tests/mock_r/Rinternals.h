/* Minimal stand-in for R's C API (Rinternals.h), TEST INFRASTRUCTURE ONLY: R is not installed in this image, so
 * r-package/TADpoleB200/src/r_shim.c is compiled against these declarations and driven by tests/mock_r/mock_r.c.
 * Only what the shim uses is declared; names, argument orders and semantics follow "Writing R Extensions". */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mock_sexp *SEXP;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22, RAWSXP = 24, CHARSXP = 9 };
extern SEXP R_NilValue;
extern double R_NaReal;
#define NA_REAL R_NaReal
#define ISNAN(x) ((x) != (x))
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(int type, ptrdiff_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
double *REAL(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
unsigned char *RAW(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
#define asInteger Rf_asInteger
#define asReal Rf_asReal
#define asLogical Rf_asLogical
SEXP Rf_ScalarInteger(int);
#define ScalarInteger Rf_ScalarInteger
SEXP SET_VECTOR_ELT(SEXP, ptrdiff_t, SEXP);
SEXP VECTOR_ELT(SEXP, ptrdiff_t);
SEXP STRING_ELT(SEXP, ptrdiff_t);
void SET_STRING_ELT(SEXP, ptrdiff_t, SEXP);
SEXP Rf_mkChar(const char *);
#define mkChar Rf_mkChar
const char *CHAR(SEXP);
int Rf_length(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_isReal(SEXP);
int Rf_isInteger(SEXP);
ptrdiff_t Rf_xlength(SEXP);
#define length Rf_length
#define nrows Rf_nrows
#define ncols Rf_ncols
#define isReal Rf_isReal
#define isInteger Rf_isInteger
#define XLENGTH Rf_xlength
char *R_alloc(size_t n, int size);
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
#define error Rf_error
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean onexit);
#ifdef __cplusplus
}
#endif
#endif
