/* Declarations-only stand-ins for the R headers (no R in this image): just enough of the API that r_shim/ uses for a
 * `gcc -fsyntax-only` check of the shim against include/baynorm_b200.h.  Not the real headers, never linked. */
#pragma once
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo *, const void *, const R_CallMethodDef *, const void *, const void *);
int R_useDynamicSymbols(DllInfo *, Rboolean);
