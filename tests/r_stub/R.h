/* Minimal declarations of the R C API -- TEST INFRASTRUCTURE ONLY (tests/test_r_shim_syntax.py): R is not installed in
 * this image, so the .Call glue (telomere-analyzer_b200/R/r_shim.c) cannot be built; these prototypes, written from
 * the "Writing R Extensions" manual, let gcc at least parse and type-check it against include/nanotel_b200.h. */
#ifndef NTL_STUB_R_H
#define NTL_STUB_R_H
#include <stddef.h>
char *R_alloc(size_t n, int size);
#endif
