// stand-in that SHADOWS the reference's SuiteSparseQR wrapper: named by the block solver, never instantiated by the reference pin
#ifndef KB_SHIM_LINEAR_SOLVER_SPQR
#define KB_SHIM_LINEAR_SOLVER_SPQR
#include <sparse_block_matrix/linear_solver.h>
namespace sparse_block_matrix {
template <typename MatrixType>
class LinearSolverQr : public LinearSolver<MatrixType> {
 public:
  virtual bool init() { return false; }
  virtual bool solve(const SparseBlockMatrix<MatrixType>&, double*, double*) { return false; }
};
}  // namespace sparse_block_matrix
#endif
