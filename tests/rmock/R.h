/* Stand-in for R.h (there is no R in this image): the declarations r_shim/ uses.  tests/rmock/rmock.c implements them
 * well enough to EXECUTE the shim against libbaynorm_b200.so from the Python tests -- vectors with attributes, S4 slots,
 * external pointers with finalizers, Rf_error as a longjmp, the .Call registration table.  Not the real R API. */
#pragma once
#include <stddef.h>
void GetRNGstate(void); void PutRNGstate(void); double unif_rand(void);
