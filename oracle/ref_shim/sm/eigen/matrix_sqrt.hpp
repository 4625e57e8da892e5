// stand-in that SHADOWS sm_eigen's computeMatrixSqrt (Eigen::LDLT): A = S S^T by a pivoted LDL^T written out (pivot on the largest
// remaining diagonal entry, first one on ties - Eigen's rule).  NOT reference code; the reference pin runs with invR = identity.
#ifndef KB_SHIM_SM_EIGEN_MATRIX_SQRT
#define KB_SHIM_SM_EIGEN_MATRIX_SQRT
#include <Eigen/Core>
#include <cmath>
#include <vector>
namespace sm { namespace eigen {
template <typename DERIVED1, typename DERIVED2>
void computeMatrixSqrt(const Eigen::MatrixBase<DERIVED1>& inMatrix, const Eigen::MatrixBase<DERIVED2>& outMatrixSqrt) {
  DERIVED2& result = const_cast<DERIVED2&>(outMatrixSqrt.derived());
  const int n = inMatrix.rows();
  Eigen::MatrixXd A(inMatrix), L(n, n);
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  L.setZero();
  std::vector<double> d(n, 0.0);
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A(i, i)) > std::fabs(A(piv, piv))) piv = i;
    if (piv != k) {
      for (int c = 0; c < n; ++c) std::swap(A(k, c), A(piv, c));
      for (int r = 0; r < n; ++r) std::swap(A(r, k), A(r, piv));
      for (int c = 0; c < k; ++c) std::swap(L(k, c), L(piv, c));
      std::swap(perm[k], perm[piv]);
    }
    d[k] = A(k, k);
    L(k, k) = 1.0;
    for (int i = k + 1; i < n; ++i) {
      L(i, k) = d[k] != 0.0 ? A(i, k) / d[k] : 0.0;
      for (int j = k + 1; j <= i; ++j) { A(i, j) -= L(i, k) * d[k] * L(j, k); A(j, i) = A(i, j); }
    }
  }
  Eigen::MatrixXd S(n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) S(perm[i], j) = L(i, j) * std::sqrt(d[j] > 0.0 ? d[j] : 0.0);
  result = S;
}
} }
#endif
