// stand-in for SuiteSparse's <cholmod.h> (SuiteSparse is not in this image) on the include path of the reference pin.  It carries the
// types and constants the reference's headers name and a FUNCTIONAL, dense replacement of the few entry points aslam_backend's own
// wrapper (BE/include/aslam/backend/implementation/Cholmod.hpp) calls on the SparseCholesky path: start / finish, analyze, factorize,
// solve, and the allocate / free family.  NOT reference code and NOT CHOLMOD: cholmod_factorize expands the compressed-column matrix A
// it is given (the reference hands it J^T with the damping block appended as extra columns), forms A A^T densely and factorises it with
// an unpivoted Cholesky in extended precision (long double, so that the stand-in adds as little rounding of its own as possible); a
// pivot that is not positive sets status = CHOLMOD_NOT_POSDEF exactly where CHOLMOD would report it.  Only the arithmetic of the
// factorisation differs from CHOLMOD's (ordering, supernodes), i.e. rounding; everything above it - the Jacobian transpose in
// compressed-column form, its row order, the appended diagonal, the right-hand side - is the reference's own code.
#ifndef KB_SHIM_CHOLMOD_H
#define KB_SHIM_CHOLMOD_H
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef long UF_long;
#define CHOLMOD_PATTERN 0
#define CHOLMOD_REAL 1
#define CHOLMOD_COMPLEX 2
#define CHOLMOD_ZOMPLEX 3
#define CHOLMOD_DOUBLE 0
#define CHOLMOD_SINGLE 1
#define CHOLMOD_INT 0
#define CHOLMOD_INTLONG 1
#define CHOLMOD_LONG 2
#define CHOLMOD_A 0
#define CHOLMOD_OK 0
#define CHOLMOD_NOT_INSTALLED (-1)
#define CHOLMOD_OUT_OF_MEMORY (-2)
#define CHOLMOD_TOO_LARGE (-3)
#define CHOLMOD_INVALID (-4)
#define CHOLMOD_NOT_POSDEF 1
#define CHOLMOD_DSMALL 2
#define CHOLMOD_NATURAL 0
#define CHOLMOD_AMD 2
#define CHOLMOD_SIMPLICIAL 0
#define CHOLMOD_AUTO 1
#define CHOLMOD_SUPERNODAL 2
#define CHOLMOD_SCALAR 0
#define CHOLMOD_ROW 1
#define CHOLMOD_COL 2
#define CHOLMOD_SYM 3

struct cholmod_sparse { size_t nrow, ncol, nzmax; void *p, *i, *nz, *x, *z; int stype, itype, xtype, dtype, sorted, packed; };
struct cholmod_dense { size_t nrow, ncol, nzmax, d; void *x, *z; int xtype, dtype; };
struct cholmod_factor { size_t n; void* Perm; long double* L; /* dense lower triangle, row-major n x n */ };
struct cholmod_common {
  int status, nmethods, supernodal;
  struct { int ordering; } method[10];
  int postorder, quick_return_if_not_posdef;
  size_t memory_inuse;
};

namespace kb_shim_cholmod {
template <typename I>
inline int factorize(cholmod_sparse* A, cholmod_factor* F, cholmod_common* c) {
  if (!A || !F || A->nrow != F->n) { c->status = CHOLMOD_INVALID; return 0; }
  const size_t n = A->nrow;
  const I* p = static_cast<const I*>(A->p);
  const I* idx = static_cast<const I*>(A->i);
  const double* x = static_cast<const double*>(A->x);
  std::vector<long double> M(n * n, 0.0L);  // extended precision: the stand-in adds as little rounding of its own as it can
  if (A->stype == 0) {  // unsymmetric: CHOLMOD factorises A A^T
    for (size_t col = 0; col < A->ncol; ++col)
      for (I a = p[col]; a < p[col + 1]; ++a)
        for (I b = p[col]; b < p[col + 1]; ++b) M[(size_t)idx[a] * n + idx[b]] += (long double)x[a] * (long double)x[b];
  } else {  // symmetric, one triangle stored
    for (size_t col = 0; col < A->ncol; ++col)
      for (I a = p[col]; a < p[col + 1]; ++a) M[(size_t)idx[a] * n + col] = M[col * n + idx[a]] = x[a];
  }
  long double* L = F->L;
  for (size_t i = 0; i < n * n; ++i) L[i] = 0.0L;
  c->status = CHOLMOD_OK;
  for (size_t j = 0; j < n; ++j)
    for (size_t i = j; i < n; ++i) {
      long double s = M[i * n + j];
      for (size_t k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0.0L)) { c->status = CHOLMOD_NOT_POSDEF; return 1; }  // CHOLMOD returns TRUE and flags the status
        L[j * n + j] = sqrtl(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  return 1;
}
inline cholmod_dense* allocate_dense(size_t nrow, size_t ncol, size_t d, int xtype, cholmod_common* c) {
  cholmod_dense* X = static_cast<cholmod_dense*>(std::calloc(1, sizeof(cholmod_dense)));
  X->nrow = nrow; X->ncol = ncol; X->d = d; X->nzmax = d * ncol; X->xtype = xtype; X->dtype = CHOLMOD_DOUBLE;
  X->x = std::calloc(X->nzmax ? X->nzmax : 1, sizeof(double));
  c->status = CHOLMOD_OK;
  return X;
}
inline cholmod_dense* solve(int sys, cholmod_factor* F, cholmod_dense* B, cholmod_common* c) {
  if (sys != CHOLMOD_A || !F || !B || B->nrow != F->n || B->ncol != 1) { c->status = CHOLMOD_INVALID; return NULL; }
  const size_t n = F->n;
  cholmod_dense* X = allocate_dense(n, 1, n, CHOLMOD_REAL, c);
  const long double* L = F->L;
  const double* b = static_cast<const double*>(B->x);
  std::vector<long double> x(n);
  for (size_t i = 0; i < n; ++i) {
    long double s = b[i];
    for (size_t k = 0; k < i; ++k) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (size_t i = n; i-- > 0;) {
    long double s = x[i];
    for (size_t k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (size_t i = 0; i < n; ++i) static_cast<double*>(X->x)[i] = (double)x[i];
  return X;
}
inline cholmod_factor* analyze(cholmod_sparse* A, cholmod_common* c) {
  cholmod_factor* F = static_cast<cholmod_factor*>(std::calloc(1, sizeof(cholmod_factor)));
  F->n = A->nrow;
  F->L = static_cast<long double*>(std::calloc(F->n * F->n ? F->n * F->n : 1, sizeof(long double)));
  c->status = CHOLMOD_OK;
  return F;
}
inline int free_dense(cholmod_dense** X) { if (X && *X) { std::free((*X)->x); std::free(*X); *X = NULL; } return 1; }
inline int free_factor(cholmod_factor** F) { if (F && *F) { std::free((*F)->L); std::free(*F); *F = NULL; } return 1; }
inline int free_sparse(cholmod_sparse** A) {
  if (A && *A) { std::free((*A)->p); std::free((*A)->i); std::free((*A)->x); std::free(*A); *A = NULL; }
  return 1;
}
inline int start(cholmod_common* c) { std::memset(c, 0, sizeof(*c)); return 1; }
}  // namespace kb_shim_cholmod

// the entry points the reference's CholmodIndexTraits<int> / <SuiteSparse_long> name (implementation/Cholmod.hpp:36-137)
#define KB_SHIM_CHOLMOD_API(PREFIX, I)                                                                                                  \
  inline int PREFIX##start(cholmod_common* c) { return kb_shim_cholmod::start(c); }                                                      \
  inline int PREFIX##finish(cholmod_common*) { return 1; }                                                                               \
  inline int PREFIX##print_sparse(cholmod_sparse*, const char*, cholmod_common*) { return 1; }                                           \
  inline cholmod_factor* PREFIX##analyze(cholmod_sparse* A, cholmod_common* c) { return kb_shim_cholmod::analyze(A, c); }                \
  inline int PREFIX##free_sparse(cholmod_sparse** A, cholmod_common*) { return kb_shim_cholmod::free_sparse(A); }                        \
  inline int PREFIX##free_dense(cholmod_dense** X, cholmod_common*) { return kb_shim_cholmod::free_dense(X); }                           \
  inline int PREFIX##free_factor(cholmod_factor** F, cholmod_common*) { return kb_shim_cholmod::free_factor(F); }                        \
  inline void* PREFIX##free(size_t, size_t, void* p, cholmod_common*) { std::free(p); return NULL; }                                     \
  inline int PREFIX##factorize(cholmod_sparse* A, cholmod_factor* F, cholmod_common* c) { return kb_shim_cholmod::factorize<I>(A, F, c); } \
  inline cholmod_dense* PREFIX##solve(int sys, cholmod_factor* F, cholmod_dense* B, cholmod_common* c) {                                 \
    return kb_shim_cholmod::solve(sys, F, B, c);                                                                                         \
  }                                                                                                                                      \
  inline cholmod_sparse* PREFIX##aat(cholmod_sparse*, I*, size_t, int, cholmod_common* c) { c->status = CHOLMOD_NOT_INSTALLED; return NULL; } \
  inline int PREFIX##scale(cholmod_dense*, int, cholmod_sparse*, cholmod_common* c) { c->status = CHOLMOD_NOT_INSTALLED; return 0; }     \
  inline cholmod_dense* PREFIX##allocate_dense(size_t nrow, size_t ncol, size_t d, int xtype, cholmod_common* c) {                       \
    return kb_shim_cholmod::allocate_dense(nrow, ncol, d, xtype, c);                                                                     \
  }
KB_SHIM_CHOLMOD_API(cholmod_, int)
KB_SHIM_CHOLMOD_API(cholmod_l_, UF_long)
#undef KB_SHIM_CHOLMOD_API
#endif
