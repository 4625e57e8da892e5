// TEST INFRASTRUCTURE: the decision numerics of the incremental estimator's solver from the REFERENCE's own source -
// IC/src/algorithms/linalg.cpp compiled where it lies (oracle/Makefile: _ref/libkalibr_ref_linalg.so) against the stand-in headers of
// oracle/ref_shim/ and oracle/ref_shim_linalg/.  Exercised here, as reference code: colNorm and columnScalingMatrix (the column scaling of
// LinearSolver::solve), rankTol, estimateNumericalRank, svGap (the numerical rank, its tolerance and gap), solveSVD (the truncated solve),
// qrTol's formula, and the dense <-> cholmod converters around them.  NOT reference code: Eigen::JacobiSVD behind analyzeSVD (one-sided
// Jacobi stand-in, ref_shim/Eigen/SVD: rounding and vector signs), spqr_maxcolnorm (SPQR's published helper: the largest column 2-norm,
// restated below), and every SuiteSparseQR / cholmod sparse-algebra entry point - the QR-coupled functions of the file (reduceLeftHandSide,
// reduceRightHandSide, solveQR, the submatrix helpers) compile but fail loudly if called.  So this pins what decides rank, truncation and
// scaling; the QR elimination itself stays unpinned.
#include <aslam/calibration/algorithms/linalg.h>

#include <cmath>
#include <cstdint>
#include <iostream>

template <typename Entry, typename Int>
double spqr_maxcolnorm(cholmod_sparse* A, cholmod_common*) {
  const Int* Ap = static_cast<const Int*>(A->p);
  const Entry* Ax = static_cast<const Entry*>(A->x);
  double maxnorm = 0.0;
  for (size_t j = 0; j < A->ncol; ++j) {
    double s = 0.0;
    for (Int p = Ap[j]; p < Ap[j + 1]; ++p) s += (double)Ax[p] * (double)Ax[p];
    maxnorm = std::max(maxnorm, std::sqrt(s));
  }
  return maxnorm;
}
template double spqr_maxcolnorm<double, int64_t>(cholmod_sparse*, cholmod_common*);

using namespace aslam::calibration;

namespace {
Eigen::MatrixXd fromRowMajor(const double* a, int rows, int cols) {
  Eigen::MatrixXd M(rows, cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) M(r, c) = a[(size_t)r * cols + c];
  return M;
}
}  // namespace

// out = [rankTol(sv, eps), estimateNumericalRank(sv, tol), svGap(sv, rank)] with tol = svd_tol unless svd_tol == -1 (then rankTol), the
// choice LinearSolver::solve makes (IC/src/core/LinearSolver.cpp:427-431)
extern "C" __attribute__((visibility("default"))) int32_t ref_linalg_rank(const double* sv, int32_t n, double eps, double svd_tol, double* out) {
  try {
    Eigen::VectorXd s(n);
    for (int i = 0; i < n; ++i) s(i) = sv[i];
    const double tol = svd_tol != -1.0 ? svd_tol : rankTol(s, eps);
    const std::ptrdiff_t rank = estimateNumericalRank(s, tol);
    out[0] = tol;
    out[1] = (double)rank;
    out[2] = svGap(s, rank);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_linalg_rank: " << e.what() << std::endl;
    return -1;
  }
}

// A: rows x cols, row-major.  G[cols] = columnScalingMatrix(A, eps); *qr_tol = qrTol(A, eps_qr)
extern "C" __attribute__((visibility("default"))) int32_t ref_linalg_column_scaling(const double* A, int32_t rows, int32_t cols, double eps, double eps_qr, double* G,
                                                                                    double* qr_tol) {
  try {
    cholmod_common cholmod;
    cholmod.status = CHOLMOD_OK;
    cholmod_sparse* As = eigenDenseToCholmodSparseCopy(fromRowMajor(A, rows, cols), &cholmod, 0.0);
    cholmod_dense* g = columnScalingMatrix(As, &cholmod, eps);
    for (int j = 0; j < cols; ++j) G[j] = static_cast<const double*>(g->x)[j];
    *qr_tol = qrTol(As, &cholmod, eps_qr);
    cholmod_l_free_dense(&g, &cholmod);
    cholmod_l_free_sparse(&As, &cholmod);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_linalg_column_scaling: " << e.what() << std::endl;
    return -1;
  }
}

// Omega: n x n row-major (symmetric), b[n].  analyzeSVD(Omega) -> sv[n]; tolerance / rank / gap as ref_linalg_rank; x = solveSVD(b, sv, U, V, rank).
// out = [tolerance, rank, gap]
extern "C" __attribute__((visibility("default"))) int32_t ref_linalg_svd_solve(const double* Omega, const double* b, int32_t n, double eps, double svd_tol, double* sv,
                                                                               double* x, double* out) {
  try {
    cholmod_common cholmod;
    cholmod.status = CHOLMOD_OK;
    cholmod_sparse* Os = eigenDenseToCholmodSparseCopy(fromRowMajor(Omega, n, n), &cholmod, 0.0);
    Eigen::VectorXd s, bv(n), xv;
    Eigen::MatrixXd U, V;
    analyzeSVD(Os, s, U, V);
    cholmod_l_free_sparse(&Os, &cholmod);
    const double tol = svd_tol != -1.0 ? svd_tol : rankTol(s, eps);
    const std::ptrdiff_t rank = estimateNumericalRank(s, tol);
    for (int i = 0; i < n; ++i) bv(i) = b[i];
    cholmod_dense bd;
    eigenDenseToCholmodDenseView(bv, &bd);
    solveSVD(&bd, s, U, V, rank, xv);
    for (int i = 0; i < n; ++i) { sv[i] = s(i); x[i] = xv(i); }
    out[0] = tol;
    out[1] = (double)rank;
    out[2] = svGap(s, rank);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_linalg_svd_solve: " << e.what() << std::endl;
    return -1;
  }
}
