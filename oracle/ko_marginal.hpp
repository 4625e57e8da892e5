// ORACLE — TEST INFRASTRUCTURE ONLY (see ko_math.hpp header: parity PINNED for the per-term part against the reference's own code, UNPINNED for the rest).
//
// Marginal analysis of the calibration block, the step the incremental estimator runs after every re-optimisation
// (IC = aslam_incremental_calibration/incremental_calibration):
//   IC/src/core/LinearSolver.cpp:466-528   analyzeMarginal: QR of the pose columns A_l, Omega = A_r^T A_r - (A_r^T Q)(A_r^T Q)^T,
//                                          SVD of Omega, numerical rank
//   IC/src/algorithms/linalg.cpp:244-282   estimateNumericalRank, rankTol (sv(0) * eps * n), svGap
//   IC/src/algorithms/linalg.cpp:284-335   reduceLeftHandSide
//   IC/src/algorithms/linalg.cpp:409-425   analyzeSVD (Eigen::JacobiSVD of the dense Omega)
//   IC/src/core/LinearSolver.cpp:196-200   getSingularValuesLog2Sum (over the first `rank` singular values)
// SuiteSparseQR and Eigen::JacobiSVD are not available here: the QR is a dense Householder factorisation that skips
// numerically zero columns, the SVD of the symmetric Omega a cyclic Jacobi eigenvalue iteration.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

namespace ko {

struct MarginalResult {
  int n = 0, rank = 0, rankDeficiency = 0;
  double tolerance = 0, svLog2Sum = 0, svGap = 0;
  std::vector<double> singularValues;  // descending
  std::vector<double> V;               // n x n row-major, column k = k-th right singular vector
  std::vector<double> Omega;           // n x n row-major
};

// A_l: m x nl, A_r: m x nr, both row-major; consumed.
inline MarginalResult analyzeMarginalDense(std::vector<double>& Al, int m, int nl, std::vector<double>& Ar, int nr, double epsSvd, double svdTol) {
  // A_r^T A_r before A_r is overwritten by Q^T A_r
  std::vector<double> Om((size_t)nr * nr, 0.0);
  for (int r = 0; r < m; ++r)
    for (int a = 0; a < nr; ++a) {
      const double va = Ar[(size_t)r * nr + a];
      if (va == 0.0) continue;
      for (int b = a; b < nr; ++b) Om[(size_t)a * nr + b] += va * Ar[(size_t)r * nr + b];
    }
  // Householder QR of A_l, the reflectors applied to A_r as they come; prow = rows consumed so far (= rank found)
  double maxNorm = 0.0;
  for (int j = 0; j < nl; ++j) {
    double s = 0.0;
    for (int r = 0; r < m; ++r) s += Al[(size_t)r * nl + j] * Al[(size_t)r * nl + j];
    maxNorm = std::max(maxNorm, std::sqrt(s));
  }
  const double qrTol = 20.0 * (m + nl) * std::numeric_limits<double>::epsilon() * maxNorm;  // SPQR's default: 20 (m+n) eps max column norm
  int prow = 0;
  std::vector<double> v(m);
  for (int j = 0; j < nl && prow < m; ++j) {
    double s = 0.0;
    for (int r = prow; r < m; ++r) s += Al[(size_t)r * nl + j] * Al[(size_t)r * nl + j];
    const double norm = std::sqrt(s);
    if (norm <= qrTol) continue;  // dependent column: no reflector, no row consumed
    const double x0 = Al[(size_t)prow * nl + j];
    const double alpha = x0 >= 0 ? -norm : norm;
    double vnorm2 = 0.0;
    for (int r = prow; r < m; ++r) {
      v[r] = Al[(size_t)r * nl + j] - (r == prow ? alpha : 0.0);
      vnorm2 += v[r] * v[r];
    }
    if (vnorm2 > 0.0) {
      const double beta = 2.0 / vnorm2;
      for (int c = j; c < nl; ++c) {
        double d = 0.0;
        for (int r = prow; r < m; ++r) d += v[r] * Al[(size_t)r * nl + c];
        d *= beta;
        if (d != 0.0)
          for (int r = prow; r < m; ++r) Al[(size_t)r * nl + c] -= d * v[r];
      }
      for (int c = 0; c < nr; ++c) {
        double d = 0.0;
        for (int r = prow; r < m; ++r) d += v[r] * Ar[(size_t)r * nr + c];
        d *= beta;
        if (d != 0.0)
          for (int r = prow; r < m; ++r) Ar[(size_t)r * nr + c] -= d * v[r];
      }
    }
    ++prow;
  }
  // Omega = A_r^T A_r - B^T B with B = the first prow rows of Q^T A_r
  for (int r = 0; r < prow; ++r)
    for (int a = 0; a < nr; ++a) {
      const double va = Ar[(size_t)r * nr + a];
      if (va == 0.0) continue;
      for (int b = a; b < nr; ++b) Om[(size_t)a * nr + b] -= va * Ar[(size_t)r * nr + b];
    }
  for (int a = 0; a < nr; ++a)
    for (int b = 0; b < a; ++b) Om[(size_t)a * nr + b] = Om[(size_t)b * nr + a];
  MarginalResult R;
  R.n = nr;
  R.Omega = Om;
  // cyclic Jacobi eigenvalue iteration on the symmetric Omega
  std::vector<double> A = Om, V((size_t)nr * nr, 0.0);
  for (int i = 0; i < nr; ++i) V[(size_t)i * nr + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < nr; ++j) (i == j ? diag : off) += A[(size_t)i * nr + j] * A[(size_t)i * nr + j];
    if (off <= 1e-32 * diag) break;
    for (int p = 0; p < nr - 1; ++p)
      for (int q = p + 1; q < nr; ++q) {
        const double apq = A[(size_t)p * nr + q];
        if (apq == 0.0) continue;
        const double theta = (A[(size_t)q * nr + q] - A[(size_t)p * nr + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < nr; ++k) {  // columns p, q
          const double akp = A[(size_t)k * nr + p], akq = A[(size_t)k * nr + q];
          A[(size_t)k * nr + p] = c * akp - s * akq;
          A[(size_t)k * nr + q] = s * akp + c * akq;
        }
        for (int k = 0; k < nr; ++k) {  // rows p, q
          const double apk = A[(size_t)p * nr + k], aqk = A[(size_t)q * nr + k];
          A[(size_t)p * nr + k] = c * apk - s * aqk;
          A[(size_t)q * nr + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < nr; ++k) {
          const double vkp = V[(size_t)k * nr + p], vkq = V[(size_t)k * nr + q];
          V[(size_t)k * nr + p] = c * vkp - s * vkq;
          V[(size_t)k * nr + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(nr);
  std::iota(order.begin(), order.end(), 0);
  std::vector<double> sv(nr);
  for (int i = 0; i < nr; ++i) sv[i] = std::fabs(A[(size_t)i * nr + i]);  // singular values of a symmetric matrix
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sv[a] > sv[b]; });
  R.singularValues.resize(nr);
  R.V.assign((size_t)nr * nr, 0.0);
  for (int k = 0; k < nr; ++k) {
    R.singularValues[k] = sv[order[k]];
    for (int r = 0; r < nr; ++r) R.V[(size_t)r * nr + k] = V[(size_t)r * nr + order[k]];
  }
  // linalg.cpp:244-282
  R.tolerance = svdTol != -1.0 ? svdTol : R.singularValues[0] * epsSvd * nr;
  R.rank = nr;
  for (int i = nr - 1; i > 0; --i) {
    if (R.singularValues[i] > R.tolerance) break;
    R.rank--;
  }
  R.rankDeficiency = nr - R.rank;
  R.svGap = R.rank < nr ? R.singularValues[R.rank - 1] / R.singularValues[R.rank] : std::numeric_limits<double>::infinity();
  R.svLog2Sum = 0.0;
  for (int i = 0; i < R.rank; ++i) R.svLog2Sum += std::log(R.singularValues[i]);
  R.svLog2Sum /= std::log(2.0);
  return R;
}

}  // namespace ko
