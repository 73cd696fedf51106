/* see R.h in this directory: prototypes only, test infrastructure */
#ifndef NTL_STUB_RINTERNALS_H
#define NTL_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef unsigned int SEXPTYPE;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24
typedef unsigned char Rbyte;
extern SEXP R_NilValue;
extern int R_NaInt;
extern double R_NaReal;
#define NA_INTEGER R_NaInt
#define NA_REAL R_NaReal
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_mkNamed(SEXPTYPE, const char **);
SEXP Rf_ScalarInteger(int);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isNull(SEXP);
void Rf_error(const char *, ...) __attribute__((noreturn));
void Rf_warning(const char *, ...);
SEXP Rf_install(const char *);
SEXP R_do_slot(SEXP, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP R_ExternalPtrTag(SEXP);
int TYPEOF(SEXP);
Rbyte *RAW(SEXP);
int LENGTH(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
double *REAL(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
const char *CHAR(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
