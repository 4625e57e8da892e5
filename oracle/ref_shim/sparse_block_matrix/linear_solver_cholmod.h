// stand-in that SHADOWS the reference's CHOLMOD wrapper on the include path of the reference pin (SuiteSparse is not in this image):
// the same LinearSolver interface over the reference's own SparseBlockMatrix, the factorisation done densely - A is expanded with the
// matrix's own toDense(), its upper triangle mirrored, and factorised by an unpivoted Cholesky; solve() returns false when a pivot is
// not positive (CHOLMOD reports "not positive definite").  NOT reference code: only the arithmetic of the factorisation differs from
// CHOLMOD's supernodal one, i.e. rounding.
#ifndef KB_SHIM_LINEAR_SOLVER_CHOLMOD
#define KB_SHIM_LINEAR_SOLVER_CHOLMOD
#include <sparse_block_matrix/linear_solver.h>

#include <cmath>
#include <vector>
namespace sparse_block_matrix {
template <typename MatrixType>
class LinearSolverCholmod : public LinearSolver<MatrixType> {
 public:
  virtual ~LinearSolverCholmod() {}
  virtual bool init() { return true; }
  virtual bool solve(const SparseBlockMatrix<MatrixType>& A, double* x, double* b) {
    const Eigen::MatrixXd D = A.toDense();
    const int n = D.rows();
    std::vector<double> L((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) {
      for (int i = j; i < n; ++i) {
        double s = i >= j && D(j, i) != 0.0 ? D(j, i) : D(i, j);  // the upper block triangle is what the solvers fill
        if (i == j) s = D(j, j);
        for (int k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
        if (i == j) {
          if (!(s > 0.0)) return false;
          L[(size_t)j * n + j] = std::sqrt(s);
        } else {
          L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
        }
      }
    }
    for (int i = 0; i < n; ++i) {
      double s = b[i];
      for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * x[k];
      x[i] = s / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = x[i];
      for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * x[k];
      x[i] = s / L[(size_t)i * n + i];
    }
    return true;
  }
};
}  // namespace sparse_block_matrix
#endif
