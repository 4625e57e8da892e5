// stand-in that SHADOWS the reference's sparse block matrix on the include path of the reference pin: the Jacobian container only
// needs it in evaluateHessian / asSparseMatrix, which the pin never calls (it reads the per-design-variable Jacobians directly).
#ifndef KB_SHIM_SPARSE_BLOCK_MATRIX
#define KB_SHIM_SPARSE_BLOCK_MATRIX
#include <Eigen/Core>
#include <vector>
namespace sparse_block_matrix {
template <typename MatrixType>
class SparseBlockMatrix {
 public:
  typedef MatrixType SparseMatrixBlock;
  SparseBlockMatrix() {}
  SparseBlockMatrix(const int*, const int*, int, int, bool = true) {}
  SparseBlockMatrix(const std::vector<int>&, const std::vector<int>&, bool = true) {}
  MatrixType* block(int, int, bool = false) { return &dummy_; }
  const MatrixType* block(int, int) const { return &dummy_; }
  int rows() const { return 0; }
  int cols() const { return 0; }
  int bRows() const { return 0; }
  int bCols() const { return 0; }
  int rowsOfBlock(int) const { return 0; }
  int colsOfBlock(int) const { return 0; }
  int rowBaseOfBlock(int) const { return 0; }
  int colBaseOfBlock(int) const { return 0; }
  void clear(bool = false) {}
  Eigen::MatrixXd toDense() const { return Eigen::MatrixXd(); }
 private:
  MatrixType dummy_;
};
}  // namespace sparse_block_matrix
#endif
