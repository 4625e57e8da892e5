// stand-in for SuiteSparse's <cholmod.h> on the include path of the LINALG pin (oracle/ref_pin_linalg.cpp; SuiteSparse is not in this image):
// the struct names IC/include/aslam/calibration/algorithms/linalg.h forward-declares and the few constants linalg.cpp names.  The sparse-
// matrix entry points linalg.cpp's QR-coupled functions call (cholmod_l_submatrix, _ssmult, _add, ...) are declared and, when reached, fail
// loudly: only the functions that work on plain arrays are exercised by the pin.
#ifndef KB_SHIM_LINALG_CHOLMOD_H
#define KB_SHIM_LINALG_CHOLMOD_H
#include <cstddef>
#include <cstdlib>
#include <stdexcept>
#define CHOLMOD_REAL 1
#define CHOLMOD_DOUBLE 0
#define CHOLMOD_LONG 2
#define CHOLMOD_OK 0
typedef struct cholmod_sparse_struct { size_t nrow, ncol, nzmax; void *p, *i, *nz, *x, *z; int stype, itype, xtype, dtype, sorted, packed; } cholmod_sparse;
typedef struct cholmod_dense_struct { size_t nrow, ncol, nzmax, d; void *x, *z; int xtype, dtype; } cholmod_dense;
typedef struct cholmod_common_struct { int status; } cholmod_common;
#define KB_LINALG_UNAVAILABLE(name) throw std::runtime_error(name ": SuiteSparse is not in this image (linalg pin)")
inline cholmod_sparse* cholmod_l_submatrix(cholmod_sparse*, void*, long, void*, long, int, int, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_submatrix"); }
inline cholmod_sparse* cholmod_l_allocate_sparse(size_t nrow, size_t ncol, size_t nzmax, int sorted, int packed, int stype, int xtype, cholmod_common*) {
  cholmod_sparse* A = static_cast<cholmod_sparse*>(std::calloc(1, sizeof(cholmod_sparse)));
  A->nrow = nrow; A->ncol = ncol; A->nzmax = nzmax; A->sorted = sorted; A->packed = packed; A->stype = stype; A->xtype = xtype; A->itype = CHOLMOD_LONG;
  A->p = std::calloc(ncol + 1, sizeof(std::ptrdiff_t));
  A->i = std::calloc(nzmax ? nzmax : 1, sizeof(std::ptrdiff_t));
  A->x = std::calloc(nzmax ? nzmax : 1, sizeof(double));
  return A;
}
inline int cholmod_l_free_sparse(cholmod_sparse** A, cholmod_common*) {
  if (A && *A) { std::free((*A)->p); std::free((*A)->i); std::free((*A)->x); std::free(*A); *A = NULL; }
  return 1;
}
inline cholmod_dense* cholmod_l_allocate_dense(size_t nrow, size_t ncol, size_t d, int xtype, cholmod_common*) {
  cholmod_dense* X = static_cast<cholmod_dense*>(std::calloc(1, sizeof(cholmod_dense)));
  X->nrow = nrow; X->ncol = ncol; X->d = d; X->nzmax = d * ncol; X->xtype = xtype; X->dtype = CHOLMOD_DOUBLE;
  X->x = std::calloc(X->nzmax ? X->nzmax : 1, sizeof(double));
  return X;
}
inline int cholmod_l_free_dense(cholmod_dense** X, cholmod_common*) { if (X && *X) { std::free((*X)->x); std::free(*X); *X = NULL; } return 1; }
inline cholmod_sparse* cholmod_l_transpose(cholmod_sparse*, int, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_transpose"); }
inline cholmod_sparse* cholmod_l_ssmult(cholmod_sparse*, cholmod_sparse*, int, int, int, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_ssmult"); }
inline cholmod_sparse* cholmod_l_add(cholmod_sparse*, cholmod_sparse*, double*, double*, int, int, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_add"); }
inline cholmod_sparse* cholmod_l_aat(cholmod_sparse*, void*, size_t, int, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_aat"); }
inline int cholmod_l_scale(cholmod_dense*, int, cholmod_sparse*, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_scale"); }
inline int cholmod_l_sdmult(cholmod_sparse*, int, double*, double*, cholmod_dense*, cholmod_dense*, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_sdmult"); }
inline cholmod_dense* cholmod_l_sparse_to_dense(cholmod_sparse*, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_sparse_to_dense"); }
inline cholmod_sparse* cholmod_l_dense_to_sparse(cholmod_dense*, int, cholmod_common*) { KB_LINALG_UNAVAILABLE("cholmod_l_dense_to_sparse"); }
#endif
