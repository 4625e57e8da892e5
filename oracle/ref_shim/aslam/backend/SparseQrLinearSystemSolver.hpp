// stand-in that SHADOWS the reference header of the same name on the include path of the reference pin: this solver needs CHOLMOD /
// SuiteSparseQR / CSparse; Optimizer2.cpp only names it when its options ask for it, which the pin (block_cholesky) never does.
#ifndef KB_SHIM_ASLAM_SparseQrLinearSystemSolver
#define KB_SHIM_ASLAM_SparseQrLinearSystemSolver
#include <aslam/backend/LinearSystemSolver.hpp>
#include <stdexcept>
namespace aslam { namespace backend {
class SparseQrLinearSystemSolver : public LinearSystemSolver {
 public:
  SparseQrLinearSystemSolver() {}
  SparseQrLinearSystemSolver(const sm::PropertyTree&) {}
  virtual void buildSystem(size_t, bool) { throw std::runtime_error("SparseQrLinearSystemSolver is not part of the reference pin"); }
  virtual bool solveSystem(Eigen::VectorXd&) { return false; }
  virtual std::string name() const { return "SparseQrLinearSystemSolver (stand-in)"; }
  virtual double rhsJtJrhs() { return 0.0; }
  virtual void initMatrixStructureImplementation(const std::vector<DesignVariable*>&, const std::vector<ErrorTerm*>&, bool) {}
};
} }
#endif
