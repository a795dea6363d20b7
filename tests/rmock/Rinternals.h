/* Stand-in for Rinternals.h: see R.h in this directory. */
#pragma once
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
extern SEXP R_NilValue;
#define NILSXP 0
#define SYMSXP 1
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22
#define S4SXP 25
SEXP Rf_allocVector(unsigned int, R_xlen_t); SEXP Rf_allocMatrix(unsigned int, int, int); SEXP Rf_alloc3DArray(unsigned int, int, int, int);
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
double *REAL(SEXP); int *INTEGER(SEXP); int *LOGICAL(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP R_do_slot(SEXP, SEXP); SEXP R_do_slot_assign(SEXP, SEXP, SEXP); SEXP Rf_install(const char *);
SEXP R_do_MAKE_CLASS(const char *); SEXP R_do_new_object(SEXP);
void Rf_error(const char *, ...) __attribute__((noreturn));
int Rf_asLogical(SEXP); int Rf_asInteger(SEXP); double Rf_asReal(SEXP);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP); void *R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP); void R_SetExternalPtrAddr(SEXP, void *);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
int Rf_isNull(SEXP); R_xlen_t Rf_xlength(SEXP); int Rf_length(SEXP); R_xlen_t XLENGTH(SEXP); int LENGTH(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP); SEXP Rf_setAttrib(SEXP, SEXP, SEXP); extern SEXP R_DimSymbol; extern SEXP R_NamesSymbol;
SEXP Rf_mkString(const char *); SEXP Rf_mkChar(const char *); SEXP Rf_ScalarReal(double); SEXP Rf_ScalarInteger(int);
int Rf_nrows(SEXP); int Rf_ncols(SEXP); int Rf_isReal(SEXP); int Rf_isInteger(SEXP); SEXP Rf_coerceVector(SEXP, unsigned int);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
