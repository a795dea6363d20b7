/* Declarations-only stand-ins for the R headers (no R in this image): just enough of the API that r_shim/ uses for a
 * `gcc -fsyntax-only` check of the shim against include/baynorm_b200.h.  Not the real headers, never linked. */
#pragma once
#include <stddef.h>
void GetRNGstate(void); void PutRNGstate(void); double unif_rand(void);
