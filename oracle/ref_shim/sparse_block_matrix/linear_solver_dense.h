// stand-in that SHADOWS the reference's dense block solver header: named by Optimizer2.cpp, never instantiated by the reference pin
#ifndef KB_SHIM_LINEAR_SOLVER_DENSE
#define KB_SHIM_LINEAR_SOLVER_DENSE
#include <sparse_block_matrix/linear_solver.h>
namespace sparse_block_matrix {
template <typename MatrixType>
class LinearSolverDense : public LinearSolver<MatrixType> {
 public:
  virtual bool init() { return false; }
  virtual bool solve(const SparseBlockMatrix<MatrixType>&, double*, double*) { return false; }
};
}  // namespace sparse_block_matrix
#endif
