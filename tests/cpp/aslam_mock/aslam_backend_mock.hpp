// TEST SCAFFOLDING - a stand-in for the part of the reference's plugin surface the B200 adapter binds to, so that
// include/kalibr_b200/reference_adapter.hpp can be COMPILED and RUN in an image that has neither Eigen nor Boost nor the
// reference's libraries.  Only the declarations the adapter touches, with the reference's exact names, virtual
// signatures and member semantics:
//   aslam::backend::LinearSystemSolver   BE/include/aslam/backend/LinearSystemSolver.hpp:16-109, BE/src/LinearSystemSolver.cpp:12-138
//   aslam::backend::DesignVariable       BE/include/aslam/backend/DesignVariable.hpp:18-145
//   aslam::backend::ErrorTerm            BE/include/aslam/backend/ErrorTerm.hpp:32-160
// (BE = aslam_optimizer/aslam_backend) plus just enough of Eigen::VectorXd / MatrixXd and SM_DEFINE_EXCEPTION.  The forwarding
// headers next to this file reproduce the reference's include paths, so the adapter's #include lines are the ones it uses in a real
// Kalibr2 tree.  Threads: the base class's fork/join over the error terms is replaced by one serial pass (same results, same order
// of the calls a plugin sees from one caller thread).  KB_REFERENCE_HAS_VIRTUAL_EVALUATE_ERROR selects INTEGRATION.md §2 option (i).
#pragma once
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace Eigen {
class VectorXd {
 public:
  VectorXd() {}
  explicit VectorXd(std::ptrdiff_t n) : v_((size_t)n, 0.0) {}
  static VectorXd Zero(std::ptrdiff_t n) { return VectorXd(n); }
  static VectorXd Constant(std::ptrdiff_t n, double c) { VectorXd r(n); for (auto& x : r.v_) x = c; return r; }
  std::ptrdiff_t size() const { return (std::ptrdiff_t)v_.size(); }
  void resize(std::ptrdiff_t n) { v_.assign((size_t)n, 0.0); }
  void conservativeResize(std::ptrdiff_t n) { v_.resize((size_t)n, 0.0); }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator[](std::ptrdiff_t i) { return v_[(size_t)i]; }
  double operator[](std::ptrdiff_t i) const { return v_[(size_t)i]; }
  double& operator()(std::ptrdiff_t i) { return v_[(size_t)i]; }
  double operator()(std::ptrdiff_t i) const { return v_[(size_t)i]; }

 private:
  std::vector<double> v_;
};
class MatrixXd {  // column-major, like Eigen's default
 public:
  MatrixXd() {}
  MatrixXd(std::ptrdiff_t r, std::ptrdiff_t c) : r_(r), c_(c), v_((size_t)(r * c), 0.0) {}
  void resize(std::ptrdiff_t r, std::ptrdiff_t c) { r_ = r; c_ = c; v_.assign((size_t)(r * c), 0.0); }
  std::ptrdiff_t rows() const { return r_; }
  std::ptrdiff_t cols() const { return c_; }
  std::ptrdiff_t size() const { return r_ * c_; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(std::ptrdiff_t r, std::ptrdiff_t c) { return v_[(size_t)(c * r_ + r)]; }
  double operator()(std::ptrdiff_t r, std::ptrdiff_t c) const { return v_[(size_t)(c * r_ + r)]; }

 private:
  std::ptrdiff_t r_ = 0, c_ = 0;
  std::vector<double> v_;
};
}  // namespace Eigen

#define SM_DEFINE_EXCEPTION(name, base) \
  struct name : public base {           \
    explicit name(const std::string& m) : base(m) {} \
  }
#define SM_ASSERT_TRUE(ex, cond, msg) do { if (!(cond)) throw ex(std::string(#cond) + " failed"); } while (0)
#define SM_ASSERT_EQ(ex, a, b, msg) do { if (!((a) == (b))) throw ex(std::string(#a " == " #b) + " failed"); } while (0)

namespace boost {
using std::make_shared;
using std::shared_ptr;
}  // namespace boost

namespace aslam {
SM_DEFINE_EXCEPTION(Exception, std::runtime_error);
namespace backend {

class JacobianContainer;
class SparseBlockMatrix;
class Matrix;

class DesignVariable {
 public:
  DesignVariable() {}
  virtual ~DesignVariable() {}
  virtual int minimalDimensions() const { return minimalDimensionsImplementation(); }
  void update(const double* update, int size) { updateImplementation(update, size); }
  void revertUpdate() { revertUpdateImplementation(); }
  bool isActive() const { return _isActive; }
  void setActive(bool active) { _isActive = active; }
  int blockIndex() const { return _blockIndex; }
  void setBlockIndex(int blockIndex) { _blockIndex = blockIndex; }
  void setScaling(double scaling) { _scaling = scaling; }
  double scaling() const { return _scaling; }
  int columnBase() const { return _columnBase; }
  void setColumnBase(int columnBase) { _columnBase = columnBase; }
  void getParameters(Eigen::MatrixXd& value) const { getParametersImplementation(value); }
  void setParameters(const Eigen::MatrixXd& value) { setParametersImplementation(value); }

 protected:
  virtual int minimalDimensionsImplementation() const = 0;
  virtual void updateImplementation(const double* dp, int size) = 0;
  virtual void revertUpdateImplementation() = 0;
  virtual void getParametersImplementation(Eigen::MatrixXd& value) const = 0;
  virtual void setParametersImplementation(const Eigen::MatrixXd& value) = 0;

 private:
  int _blockIndex = -1, _columnBase = -1;
  bool _isActive = false;
  double _scaling = 1.0;
};

class ErrorTerm {
 public:
  typedef boost::shared_ptr<aslam::backend::ErrorTerm> Ptr;
  ErrorTerm() {}
  virtual ~ErrorTerm() {}
  double evaluateError() { return _squaredError = evaluateErrorImplementation(); }
  void evaluateJacobians(JacobianContainer& outJacobians) const { evaluateJacobiansImplementation(outJacobians); }
  virtual void getWeightedJacobians(JacobianContainer& outJc, bool useMEstimator) = 0;
  virtual void getWeightedError(Eigen::VectorXd& e, bool useMEstimator) const = 0;
  virtual void getInvR(Eigen::MatrixXd& invR) const = 0;
  virtual Eigen::MatrixXd vsInvR() const = 0;
  virtual void vsSetInvR(const Eigen::MatrixXd& invR) = 0;
  void buildHessian(SparseBlockMatrix& outHessian, Eigen::VectorXd& outRhs, bool useMEstimator) { buildHessianImplementation(outHessian, outRhs, useMEstimator); }
  size_t numDesignVariables() const { return _designVariables.size(); }
  DesignVariable* designVariable(size_t i) { return _designVariables[i]; }
  const DesignVariable* designVariable(size_t i) const { return _designVariables[i]; }
  double getWeightedSquaredError() const { return _squaredError; }
  size_t dimension() const { return getDimensionImplementation(); }
  const std::vector<DesignVariable*>& designVariables() const { return _designVariables; }
  size_t rowBase() const { return _rowBase; }
  void setRowBase(size_t b) { _rowBase = b; }

 protected:
  virtual double evaluateErrorImplementation() = 0;
  virtual void evaluateJacobiansImplementation(JacobianContainer& outJacobians) const = 0;
  virtual size_t getDimensionImplementation() const = 0;
  virtual void buildHessianImplementation(SparseBlockMatrix& outHessian, Eigen::VectorXd& outRhs, bool useMEstimator) = 0;
  virtual Eigen::VectorXd vsErrorImplementation() const = 0;
  void setDesignVariables(const std::vector<DesignVariable*>& designVariables) { _designVariables = designVariables; }

 private:
  double _squaredError = 0.0;
  std::vector<DesignVariable*> _designVariables;
  size_t _rowBase = 0;
};

class LinearSystemSolver {
 public:
  SM_DEFINE_EXCEPTION(Exception, std::runtime_error);
  LinearSystemSolver() {}
  virtual ~LinearSystemSolver() {}
#ifdef KB_REFERENCE_HAS_VIRTUAL_EVALUATE_ERROR
  virtual
#endif
  double evaluateError(size_t nThreads, bool useMEstimator) {
    (void)nThreads;
    _threadLocalErrors.assign(1, 0.0);
    evaluateErrors(0, 0, _errorTerms.size(), useMEstimator);
    return _threadLocalErrors[0];
  }
  void initMatrixStructure(const std::vector<DesignVariable*>& dvs, const std::vector<ErrorTerm*>& errors, bool useDiagonalConditioner) {
    setOrdering(dvs, errors);
    _errorTerms = errors;
    _JRows = 0;
    for (ErrorTerm* e : errors) _JRows += e->dimension();
    _JCols = 0;
    for (DesignVariable* d : dvs) _JCols += (size_t)d->minimalDimensions();
    _e.resize((std::ptrdiff_t)_JRows);
    _rhs.resize((std::ptrdiff_t)_JCols);
    _diagonalConditioner = Eigen::VectorXd::Zero((std::ptrdiff_t)_JCols);
    initMatrixStructureImplementation(dvs, errors, useDiagonalConditioner);
  }
  virtual void buildSystem(size_t nThreads, bool useMEstimator) = 0;
  virtual void setConditioner(const Eigen::VectorXd& diag) {
    SM_ASSERT_EQ(Exception, (size_t)diag.size(), _JCols, "conditioner size");
    _diagonalConditioner = diag;
  }
  virtual void setConstantConditioner(double diag) { _diagonalConditioner = Eigen::VectorXd::Constant((std::ptrdiff_t)_JCols, diag); }
  virtual bool solveSystem(Eigen::VectorXd& outDx) = 0;
  virtual std::string name() const = 0;
  virtual const Eigen::VectorXd& rhs() const { return _rhs; }
  virtual const Matrix* Jacobian() const { return NULL; }
  virtual const Matrix* Hessian() const { return NULL; }
  virtual const Eigen::VectorXd& e() const { return _e; }
  size_t JRows() const { return _JRows; }
  size_t JCols() const { return _JCols; }
  virtual double rhsJtJrhs() = 0;

 protected:
  virtual void initMatrixStructureImplementation(const std::vector<DesignVariable*>& dvs, const std::vector<ErrorTerm*>& errors, bool useDiagonalConditioner) = 0;
  virtual void setOrdering(const std::vector<DesignVariable*>&, const std::vector<ErrorTerm*>&) {}
  void evaluateErrors(size_t threadId, size_t startIdx, size_t endIdx, bool useMEstimator) {
    Eigen::VectorXd e;
    for (size_t i = startIdx; i < endIdx; ++i) {
      _threadLocalErrors[threadId] += _errorTerms[i]->evaluateError();
      _errorTerms[i]->getWeightedError(e, useMEstimator);
      for (size_t r = 0; r < _errorTerms[i]->dimension(); ++r) _e[(std::ptrdiff_t)(_errorTerms[i]->rowBase() + r)] = -e[(std::ptrdiff_t)r];
    }
  }
  std::vector<ErrorTerm*> _errorTerms;
  std::vector<double> _threadLocalErrors;
  Eigen::VectorXd _e, _rhs;
  bool _useDiagonalConditioner = false;
  Eigen::VectorXd _diagonalConditioner;
  size_t _JRows = 0, _JCols = 0;
};

}  // namespace backend
}  // namespace aslam
