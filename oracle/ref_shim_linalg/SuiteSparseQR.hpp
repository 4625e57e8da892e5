// stand-in for <SuiteSparseQR.hpp> on the include path of the LINALG pin: the names linalg.cpp's QR-coupled functions mention; they fail
// loudly when reached (SuiteSparse is not in this image) - the pin exercises only the functions that work on plain arrays.
#ifndef KB_SHIM_LINALG_SPQR_HPP
#define KB_SHIM_LINALG_SPQR_HPP
#include <cholmod.h>
#define SPQR_QTX 0
#define SPQR_QX 1
#define SPQR_XQT 2
#define SPQR_XQ 3
#define SPQR_RETX_EQUALS_B 1
struct spqr_symbolic { long m, n, anz; };
struct spqr_numeric_stub { int unused; };
template <typename T> struct SuiteSparseQR_factorization { spqr_symbolic* QRsym; spqr_numeric_stub* QRnum; long rank; };
template <typename T> inline cholmod_sparse* SuiteSparseQR_qmult(int, SuiteSparseQR_factorization<T>*, cholmod_sparse*, cholmod_common*) { KB_LINALG_UNAVAILABLE("SuiteSparseQR_qmult"); }
template <typename T> inline cholmod_dense* SuiteSparseQR_qmult(int, SuiteSparseQR_factorization<T>*, cholmod_dense*, cholmod_common*) { KB_LINALG_UNAVAILABLE("SuiteSparseQR_qmult"); }
template <typename T> inline cholmod_dense* SuiteSparseQR_solve(int, SuiteSparseQR_factorization<T>*, cholmod_dense*, cholmod_common*) { KB_LINALG_UNAVAILABLE("SuiteSparseQR_solve"); }
#endif
