// stand-in: the CHOLMOD types the sparse-block-matrix headers name (SuiteSparse is not in this image); nothing here computes
#ifndef KB_SHIM_CHOLMOD_H
#define KB_SHIM_CHOLMOD_H
#include <cstddef>
#define CHOLMOD_REAL 1
#define CHOLMOD_INT 0
#define CHOLMOD_LONG 2
#define CHOLMOD_PATTERN 0
#define CHOLMOD_DOUBLE 0
#define CHOLMOD_A 0
struct cholmod_sparse { size_t nrow, ncol, nzmax; void *p, *i, *nz, *x, *z; int stype, itype, xtype, dtype, sorted, packed; };
struct cholmod_dense { size_t nrow, ncol, nzmax, d; void *x, *z; int xtype, dtype; };
struct cholmod_factor { size_t n; void* Perm; };
struct cholmod_common { int status, nmethods, supernodal; struct { int ordering; } method[10]; int postorder; };
#endif
