/* Minimal stand-in for R's <Rinternals.h>: ONLY the declarations r/ccgp_shim.c uses, with the
 * signatures of R >= 4.0's public C API.  It exists so the CPU test suite can run
 * `gcc -fsyntax-only -Wall -Werror` over the shim in an image that has no R; it is never linked. */
#ifndef CCGP_STUB_RINTERNALS_H
#define CCGP_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue;
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
double* REAL(SEXP);
int* INTEGER(SEXP);
R_xlen_t XLENGTH(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isMatrix(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isNull(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_ExternalPtrProtected(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
void Rf_error(const char*, ...) __attribute__((noreturn, format(printf, 1, 2)));
char* R_alloc(size_t, int);
#endif
