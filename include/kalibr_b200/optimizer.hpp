// Host-side C++ mirror of the reference's optimiser plugin surface for the batch-calibration hot path,
// written over the C ABI (kalibr_b200.h).  Same class / method names, argument meaning and error behaviour as
//   aslam::backend::LinearSystemSolver              BE/include/aslam/backend/LinearSystemSolver.hpp:16-109
//   aslam::backend::TrustRegionPolicy subclasses (policy selection only; the schedule lives in lm_state_machine.h)
//   aslam::backend::Optimizer2 / Optimizer2Options / SolutionReturnValue
//                                                   BE/src/Optimizer2.cpp:183-318, Optimizer2Options.hpp:9-41, backend.hpp:14-27
// (BE = aslam_optimizer/aslam_backend).  Eigen::VectorXd is replaced by std::vector<double>; design variables and
// error terms live on the device behind the handle, so initMatrixStructure takes the flattened problem.
// INTEGRATION.md shows the adapter a Kalibr2 maintainer would add on the reference side.
#pragma once
#include <cmath>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../kalibr_b200.h"
#include "lm_state_machine.h"

namespace kalibr_b200 {
namespace backend {

// A failed C-ABI call: carries the kb_status so that kb_optimize* can hand the original code back to the caller
struct KbError : std::runtime_error {
  KbError(kb_status c, const std::string& what) : std::runtime_error(what), code(c) {}
  kb_status code;
};

struct SolutionReturnValue {  // backend.hpp:14-27
  double JStart = 0, JFinal = 0;
  int iterations = 0, failedIterations = 0;
  double dXFinal = 0, dJFinal = 0;
  bool linearSolverFailure = false;
};

// The solver interface the optimiser drives.  evaluateError is virtual here (SURVEY.md §8b option (i)).
class LinearSystemSolver {
 public:
  virtual ~LinearSystemSolver() {}
  virtual double evaluateError(size_t nThreads, bool useMEstimator) = 0;
  virtual void buildSystem(size_t nThreads, bool useMEstimator) = 0;
  virtual bool solveSystem(std::vector<double>& outDx) = 0;
  virtual void setConstantConditioner(double diag) = 0;
  virtual const std::vector<double>& rhs() = 0;
  virtual const std::vector<double>& e() = 0;
  virtual size_t JRows() const = 0;
  virtual size_t JCols() const = 0;
  virtual std::string name() const = 0;
  // dx^T (lambda dx + rhs) for the last solution; the default follows getLmRho literally on the host
  virtual double lmRhoDenominator(double lambda, const std::vector<double>& dx) {
    const std::vector<double>& r = rhs();
    double d2 = 0;
    for (size_t i = 0; i < dx.size(); ++i) d2 += dx[i] * (lambda * dx[i] + r[i]);
    return d2;
  }
  // Optimizer2::applyStateUpdate / revertLastStateUpdate act on design variables the solver's device owns
  virtual double applyStateUpdate(const std::vector<double>& dx) = 0;
  virtual void revertLastStateUpdate() = 0;
};

// B200 replacement of BlockCholeskyLinearSystemSolver (BE/src/BlockCholeskyLinearSystemSolver.cpp:34-106):
// fused linearise+assemble, batched Schur elimination of the per-set poses, dense Cholesky of the camera system.
class B200SchurLinearSystemSolver : public LinearSystemSolver {
 public:
  explicit B200SchurLinearSystemSolver(kb_handle* h, bool keepDxOnDevice = true) : _h(h), _deviceDx(keepDxOnDevice) {}
  std::string name() const override { return "b200_schur"; }
  size_t JRows() const override { return (size_t)kb_jrows(_h); }
  size_t JCols() const override { return (size_t)kb_jcols(_h); }
  double evaluateError(size_t /*nThreads*/, bool useMEstimator) override {
    double J = 0;
    check(kb_evaluate_error(_h, useMEstimator ? 1 : 0, &J));
    return J;
  }
  void buildSystem(size_t /*nThreads*/, bool useMEstimator) override { check(kb_build_system(_h, useMEstimator ? 1 : 0)); }
  void setConstantConditioner(double diag) override { check(kb_set_constant_conditioner(_h, diag)); }
  bool solveSystem(std::vector<double>& outDx) override {
    int32_t posDef = 0;
    if (_deviceDx) {
      check(kb_solve_system(_h, nullptr, 0, &posDef));
    } else {
      outDx.resize(JCols());
      check(kb_solve_system(_h, outDx.data(), 1, &posDef));
    }
    return posDef != 0;
  }
  const std::vector<double>& rhs() override {
    _rhs.resize(JCols());
    check(kb_get_rhs(_h, _rhs.data()));
    return _rhs;
  }
  const std::vector<double>& e() override {
    _e.resize((size_t)kb_local_jrows(_h));
    check(kb_get_error_vector(_h, _e.data()));
    return _e;
  }
  double lmRhoDenominator(double lambda, const std::vector<double>& dx) override {
    if (!_deviceDx) return LinearSystemSolver::lmRhoDenominator(lambda, dx);
    double d = 0;
    check(kb_lm_rho_denominator(_h, lambda, &d));
    return d;
  }
  double applyStateUpdate(const std::vector<double>& /*dx: the device already holds the last solution*/) override {
    double m = 0;
    check(kb_apply_state_update(_h, &m));
    return m;
  }
  void revertLastStateUpdate() override { check(kb_revert_last_state_update(_h)); }

 private:
  void check(kb_status s) {
    if (s != KB_OK) throw KbError(s, std::string("kalibr_b200: ") + kb_last_error(_h));
  }
  kb_handle* _h;
  bool _deviceDx;
  std::vector<double> _rhs, _e;
};

// B200 replacement of aslam::calibration::LinearSolver (aslam_incremental_calibration/incremental_calibration/src/core/
// LinearSolver.cpp:241-463), the solver of the incremental estimator: same build, but an undamped solve whose calibration block goes
// through a truncated SVD (kb_solve_system_svd).
class B200SvdLinearSystemSolver : public B200SchurLinearSystemSolver {
 public:
  B200SvdLinearSystemSolver(kb_handle* h, const kb_svd_solver_options& options) : B200SchurLinearSystemSolver(h, true), _hs(h), _svdOptions(options) {}
  std::string name() const override { return "b200_marginal_svd"; }
  void setConstantConditioner(double /*diag*/) override {}  // the Gauss-Newton policy never asks for one
  bool solveSystem(std::vector<double>& /*outDx: stays on the device*/) override {
    const kb_status st = kb_solve_system_svd(_hs, &_svdOptions, nullptr, 0, &_last, nullptr);
    if (st == KB_OK) return true;
    // only the NUMERICAL outcomes (pose block not positive definite, SVD iteration not converged) are "the solve failed" (≙ the
    // exceptions LinearSolver::solveSystem turns into false, IC/src/core/LinearSolver.cpp:245-285); a device / NCCL / state error
    // must not be retried as a linear-solver failure
    if (st == KB_ERR_NUMERICAL) return false;
    throw KbError(st, std::string("kalibr_b200: ") + kb_last_error(_hs));
  }
  const kb_svd_solve_result& lastSolve() const { return _last; }

 private:
  kb_handle* _hs;
  kb_svd_solver_options _svdOptions;
  kb_svd_solve_result _last{};
};

// Trust-region policies.  In the reference these classes carry the lambda schedule themselves
// (BE/src/LevenbergMarquardtTrustRegionPolicy.cpp:50-113, BE/src/GaussNewtonTrustRegionPolicy.cpp:18-40, BE/src/TrustRegionPolicy.cpp:28-57);
// here they only SELECT a policy and hold its options: the transitions live once, in lm_state_machine.h, for host and device.
class TrustRegionPolicy {
 public:
  virtual ~TrustRegionPolicy() {}
  virtual std::string name() const = 0;
  virtual int kind() const = 0;  // KB_POLICY_*
  virtual bool requiresAugmentedDiagonal() const = 0;
  virtual bool revertOnFailure() const = 0;
  virtual double lambdaInit() const { return 0.0; }
};
class GaussNewtonTrustRegionPolicy : public TrustRegionPolicy {
 public:
  std::string name() const override { return "gauss_newton"; }
  int kind() const override { return KB_POLICY_GAUSS_NEWTON; }
  bool requiresAugmentedDiagonal() const override { return false; }
  bool revertOnFailure() const override { return false; }
};
class LevenbergMarquardtTrustRegionPolicy : public TrustRegionPolicy {
 public:
  explicit LevenbergMarquardtTrustRegionPolicy(double lambdaInit = 1e-3) : _lambdaInit(lambdaInit) {}
  std::string name() const override { return "levenberg_marquardt"; }
  int kind() const override { return KB_POLICY_LEVENBERG_MARQUARDT; }
  bool requiresAugmentedDiagonal() const override { return true; }
  bool revertOnFailure() const override { return true; }
  double lambdaInit() const override { return _lambdaInit; }

 private:
  double _lambdaInit;
};

struct Optimizer2Options {  // Optimizer2Options.hpp:9-41 with kalibr2's values (CalibrationTools.hpp:57-66) as defaults
  double convergenceDeltaJ = 1.0;
  double convergenceDeltaX = 1e-3;
  int maxIterations = 200;
  int nThreads = 4;
  bool verbose = false;
  std::shared_ptr<LinearSystemSolver> linearSystemSolver;
  std::shared_ptr<TrustRegionPolicy> trustRegionPolicy;
};

// Host driver of the state machine over a LinearSystemSolver: the same three events per iteration the device-resident loop runs
// in its control kernels (kb_kernels.cu), with the solver's virtual calls in between.  ≙ Optimizer2::optimize (BE/src/Optimizer2.cpp:183-273)
class Optimizer2 {
 public:
  explicit Optimizer2(const Optimizer2Options& options) : _options(options) {}
  Optimizer2Options& options() { return _options; }
  double J() const { return _state.J; }
  const std::vector<double>& dx() const { return _dx; }
  const std::vector<double>& trace() const { return _trace; }  // (J, deltaX, lambda) per iteration
  const LmState& state() const { return _state; }

  SolutionReturnValue optimize() {
    if (!_options.linearSystemSolver) throw std::runtime_error("kalibr_b200::Optimizer2: a B200 linear system solver must be set (no CPU fallback)");
    LinearSystemSolver& solver = *_options.linearSystemSolver;
    std::shared_ptr<TrustRegionPolicy> policy =
        _options.trustRegionPolicy ? _options.trustRegionPolicy : std::shared_ptr<TrustRegionPolicy>(std::make_shared<LevenbergMarquardtTrustRegionPolicy>());
    _trace.clear();
    const double J0 = solver.evaluateError(_options.nThreads, true);
    LmState& c = _state;
    // semantic 1: the host-driven solver keeps its own lambda^2 / lambda residual (kb_solve_system), the machine's copy is unused
    lm_start(&c, policy->kind(), J0, policy->lambdaInit(), _options.convergenceDeltaX, _options.convergenceDeltaJ, _options.maxIterations, 1);
    if (_options.verbose) std::printf("[0.0]: J: %.10g\n", J0);
    while (!c.done) {
      lm_before_solve(&c);
      if (c.need_build) solver.buildSystem(_options.nThreads, true);
      if (policy->requiresAugmentedDiagonal()) solver.setConstantConditioner(c.lambda);
      const bool ok = solver.solveSystem(_dx);
      // dx^T (lambda dx + rhs) of THIS solve: what getLmRho reads at the next iteration, before any rebuild
      const double rho_den = ok && policy->kind() == KB_POLICY_LEVENBERG_MARQUARDT ? solver.lmRhoDenominator(c.lambda, _dx) : 1.0;
      if (!ok) {
        if (_options.verbose) std::printf("[WARNING] System solution failed\n");
        lm_after_solve(&c, rho_den, 0.0, 0);
        continue;
      }
      const double max_dx = solver.applyStateUpdate(_dx);
      lm_after_solve(&c, rho_den, max_dx, 1);
      const double J = solver.evaluateError(_options.nThreads, true);
      lm_after_eval(&c, J);
      if (c.revert) {
        if (_options.verbose) std::printf("Last step was a regression. Reverting\n");
        solver.revertLastStateUpdate();
      }
      _trace.push_back(c.J);
      _trace.push_back(c.deltaX);
      _trace.push_back(c.lambda);
      if (_options.verbose)
        std::printf("[%d]: J: %.10g, dJ: %.6g, deltaX: %.6g, %s - lambda:%.6g mu:%.6g\n", c.iterations, c.J, c.deltaJ, c.deltaX, policy->name().c_str(), c.lambda, c.mu);
    }
    SolutionReturnValue srv;
    srv.JStart = c.JStart;
    srv.JFinal = c.pJ;
    srv.iterations = c.iterations;
    srv.failedIterations = c.failed;
    srv.dXFinal = c.deltaX;
    srv.dJFinal = c.deltaJ;
    srv.linearSolverFailure = c.solver_failure != 0;
    return srv;
  }

 private:
  Optimizer2Options _options;
  LmState _state{};
  std::vector<double> _dx, _trace;
};

}  // namespace backend
}  // namespace kalibr_b200
