// The reference-side binding: a drop-in aslam::backend::LinearSystemSolver over the C ABI (kalibr_b200.h), for an UNMODIFIED
// aslam::backend::Optimizer2 that owns the design variables on the host.  This is the file a Kalibr2 maintainer adds to the
// reference tree; it includes the reference's own headers
//   aslam/backend/LinearSystemSolver.hpp   (BE/include/aslam/backend/LinearSystemSolver.hpp:16-109: the virtuals implemented below)
//   aslam/backend/DesignVariable.hpp       (BE/include/aslam/backend/DesignVariable.hpp:18-145: getParameters, columnBase, ...)
//   aslam/backend/ErrorTerm.hpp            (BE/include/aslam/backend/ErrorTerm.hpp:32-160: the batched error-term provider)
// and nothing else of this repository but the C header.  In this repository it is compiled and run against a stand-in of those three
// headers (tests/cpp/aslam_mock/, same include paths and signatures) by tests/test_drivers_gpu.py::test_reference_adapter_*.
//
// How it sits under Optimizer2 (BE/src/Optimizer2.cpp:95-151, 183-318):
//   initialize()            -> initMatrixStructureImplementation: kb_create from the recorded problem; the design-variable layout the
//                              optimiser assigned (blockIndex / columnBase) is CHECKED against the library's (same ordering rules)
//   evaluateError()         -> NON-virtual in the reference (LinearSystemSolver.hpp:25): it loops over the error terms.  The terms the
//                              recorder hands to the problem are ReprojectionErrorProxy objects whose evaluateErrorImplementation() /
//                              getWeightedError() read one batched device evaluation (kb_evaluate_error + kb_get_error_vector), run
//                              by the first term asked after the state may have changed (INTEGRATION.md §2 option (ii)).  With
//                              KB_REFERENCE_HAS_VIRTUAL_EVALUATE_ERROR (option (i): `virtual` added to that one declaration) the
//                              solver overrides evaluateError itself and no residual leaves the device.
//   buildSystem()           -> kb_build_system (after pushing the host design variables if they changed: applyStateUpdate and
//                              revertLastStateUpdate act on the HOST variables, Optimizer2.cpp:290-318)
//   setConstantConditioner  -> base class (fills _diagonalConditioner); solveSystem hands it over with kb_set_conditioner
//   solveSystem(outDx)      -> kb_solve_system (dx gathered to the host) + kb_get_rhs into _rhs (getLmRho reads rhs():
//                              LevenbergMarquardtTrustRegionPolicy.cpp:107-113)
#pragma once
#include <aslam/backend/DesignVariable.hpp>
#include <aslam/backend/ErrorTerm.hpp>
#include <aslam/backend/LinearSystemSolver.hpp>

#include <cstring>
#include <string>
#include <vector>

#include "../kalibr_b200.h"

namespace kalibr2 {
namespace b200 {

// Filled by the driver while it adds design variables and error terms (the loops of CalibrationTools.hpp:93-144, 183-300, 376-428,
// 460-521 already visit exactly these objects): the flattened problem kb_create wants, and which host design variable holds which
// part of the state.
struct ProblemRecorder {
  int32_t driver_order = KB_ORDER_RIG;  // which driver built the problem: fixes the design-variable order
  std::vector<int32_t> cam_model;       // per camera, from CreateCalibrator's model string (CameraCalibrator.hpp:421-441)
  std::vector<double> target_points;    // [n][3]
  std::vector<double> y_u, y_v;
  std::vector<int32_t> corner_id, view_set, view_cam;
  std::vector<int64_t> view_begin{0};
  // host design variables, by role
  std::vector<aslam::backend::DesignVariable*> projection, distortion;  // per camera (CameraCalibrator.hpp:116-122)
  std::vector<aslam::backend::DesignVariable*> baseline_q, baseline_t;  // per baseline (CalibrationTools.hpp:32-45)
  std::vector<aslam::backend::DesignVariable*> set_q, set_t;            // per synced set
  // AddReprojectionErrorsForView: one call per valid corner, then endView (CameraCalibrator.hpp:245-264)
  void addTerm(int corner, double yu, double yv) { corner_id.push_back(corner); y_u.push_back(yu); y_v.push_back(yv); }
  void endView(int set, int cam) { view_set.push_back(set); view_cam.push_back(cam); view_begin.push_back((int64_t)y_u.size()); }
  int n_cams() const { return (int)cam_model.size(); }
  int n_sets() const { return (int)set_q.size(); }
};

class SchurLinearSystemSolver : public aslam::backend::LinearSystemSolver {
 public:
  explicit SchurLinearSystemSolver(const ProblemRecorder& recorder, int device = 0) : _rec(recorder), _device(device) {}
  ~SchurLinearSystemSolver() override { kb_destroy(_h); }
  std::string name() const override { return "b200_schur"; }
  kb_handle* handle() { return _h; }

  // ---- LinearSystemSolver ----
  void buildSystem(size_t /*nThreads*/, bool useMEstimator) override {
    pushStateIfChanged();
    check(kb_build_system(_h, useMEstimator ? 1 : 0));
  }
  bool solveSystem(Eigen::VectorXd& outDx) override {
    check(kb_set_conditioner(_h, _diagonalConditioner.data()));  // constant for every policy of the reference (LinearSystemSolver.cpp:111-114)
    outDx.resize((std::ptrdiff_t)_JCols);
    int32_t posDef = 0;
    check(kb_solve_system(_h, outDx.data(), /*gather_dx=*/1, &posDef));
    check(kb_get_rhs(_h, _rhs.data()));
    _residualsStale = true;  // Optimizer2 applies outDx to the host variables next, then evaluates
    return posDef != 0;
  }
  double rhsJtJrhs() override { throw Exception("b200_schur: rhsJtJrhs (DogLeg) is not provided"); }
#ifdef KB_REFERENCE_HAS_VIRTUAL_EVALUATE_ERROR
  double evaluateError(size_t /*nThreads*/, bool useMEstimator) override {
    pushStateIfChanged();
    double J = 0.0;
    check(kb_evaluate_error(_h, useMEstimator ? 1 : 0, &J));
    _eFetched = false;
    return J;
  }
  const Eigen::VectorXd& e() const override {  // on demand: nothing in the LM loop reads it
    if (!_eFetched) {
      const_cast<SchurLinearSystemSolver*>(this)->check(kb_get_error_vector(_h, const_cast<double*>(_e.data())));
      _eFetched = true;
    }
    return _e;
  }
#endif

  // ---- the batched residuals behind the error-term proxies (option (ii)) ----
  // -e() of term i as the device evaluated it at the host's current state; the first call after a solve runs the batch
  const double* weightedError(size_t term, bool useMEstimator) {
    if (_residualsStale || _batchUseM != useMEstimator) {
      pushStateIfChanged();
      double J = 0.0;
      check(kb_evaluate_error(_h, useMEstimator ? 1 : 0, &J));
      _batch.resize(_JRows);
      check(kb_get_error_vector(_h, _batch.data()));
      _residualsStale = false;
      _batchUseM = useMEstimator;
    }
    return &_batch[2 * term];
  }

 protected:
  void initMatrixStructureImplementation(const std::vector<aslam::backend::DesignVariable*>& dvs, const std::vector<aslam::backend::ErrorTerm*>& errors,
                                         bool /*useDiagonalConditioner*/) override {
    const ProblemRecorder& r = _rec;
    gatherState();
    kb_problem_desc d;
    std::memset(&d, 0, sizeof(d));
    d.driver_order = r.driver_order;
    d.n_cams = r.n_cams();
    d.cam_model = r.cam_model.data();
    d.cam_params = _cam.data();
    d.baselines = _base.data();
    d.n_sets = r.n_sets();
    d.set_poses = _sets.data();
    d.n_target_points = (int32_t)(r.target_points.size() / 3);
    d.target_points = r.target_points.data();
    d.n_views = (int32_t)r.view_set.size();
    d.view_set = r.view_set.data();
    d.view_cam = r.view_cam.data();
    d.view_begin = r.view_begin.data();
    d.n_terms = (int64_t)r.y_u.size();
    d.y_u = r.y_u.data();
    d.y_v = r.y_v.data();
    d.corner_id = r.corner_id.data();
    d.n_ranks = 1;
    d.device = _device;
    if (_h) { kb_destroy(_h); _h = nullptr; }
    if (kb_create(&d, &_h) != KB_OK) throw Exception(std::string("b200_schur: ") + kb_last_error(nullptr));
    _pushed_cam = _cam; _pushed_base = _base; _pushed_sets = _sets;
    // the optimiser's layout must be the library's: same rules (Optimizer2.cpp:110-135), checked variable by variable
    SM_ASSERT_EQ(Exception, (size_t)kb_jcols(_h), _JCols, "design-variable layout mismatch");
    SM_ASSERT_EQ(Exception, (size_t)kb_jrows(_h), _JRows, "error-term layout mismatch");
    SM_ASSERT_EQ(Exception, (size_t)kb_num_design_variables(_h), dvs.size(), "number of active design variables");
    std::vector<int32_t> col(dvs.size()), dim(dvs.size());
    check(kb_get_dv_layout(_h, col.data(), dim.data()));
    for (size_t i = 0; i < dvs.size(); ++i)
      if (dvs[i]->columnBase() != col[i] || dvs[i]->minimalDimensions() != dim[i] || dvs[i]->blockIndex() != (int)i)
        throw Exception("b200_schur: design variable " + std::to_string(i) + " is not where driver order " + std::to_string(r.driver_order) + " puts it");
    auto at = [&](aslam::backend::DesignVariable* v, const char* what) {
      if (!v->isActive() || v->blockIndex() < 0 || (size_t)v->blockIndex() >= dvs.size() || dvs[(size_t)v->blockIndex()] != v)
        throw Exception(std::string("b200_schur: recorded ") + what + " design variable is not an active variable of the problem");
      return (size_t)v->blockIndex();
    };
    // roles: the recorded variables sit at the block indices the library derives for them
    std::vector<int32_t> expect = roleBlocks();
    size_t q = 0;
    for (int k = 0; k < r.n_cams(); ++k) {
      if ((int32_t)at(r.projection[(size_t)k], "projection") != expect[q++]) throw Exception("b200_schur: projection variable order");
      if ((int32_t)at(r.distortion[(size_t)k], "distortion") != expect[q++]) throw Exception("b200_schur: distortion variable order");
    }
    for (size_t j = 0; j < r.baseline_q.size(); ++j) {
      if ((int32_t)at(r.baseline_q[j], "baseline") != expect[q++]) throw Exception("b200_schur: baseline variable order");
      if ((int32_t)at(r.baseline_t[j], "baseline") != expect[q++]) throw Exception("b200_schur: baseline variable order");
    }
    for (size_t v = 0; v < r.set_q.size(); ++v) {
      if ((int32_t)at(r.set_q[v], "set pose") != expect[q++]) throw Exception("b200_schur: set pose variable order");
      if ((int32_t)at(r.set_t[v], "set pose") != expect[q++]) throw Exception("b200_schur: set pose variable order");
    }
    (void)errors;
    _residualsStale = true;
  }

 private:
  void check(kb_status s) {
    if (s != KB_OK) throw Exception(std::string("b200_schur: ") + kb_last_error(_h));
  }
  // block index of every recorded variable, in the order (proj, dist per camera; q, t per baseline; q, t per set), for the four
  // driver orders of kb_driver_order
  std::vector<int32_t> roleBlocks() const {
    const int C = _rec.n_cams(), S = _rec.n_sets(), B = C - 1;
    int intr0 = 0, base0 = 0, set0 = 0;
    switch (_rec.driver_order) {
      case KB_ORDER_SINGLE: intr0 = 0; set0 = 2; break;
      case KB_ORDER_STEREO: base0 = 0; set0 = 2 * B; intr0 = 2 * B + 2 * S; break;
      case KB_ORDER_BATCH: set0 = 0; base0 = 2 * S; intr0 = 2 * S + 2 * B; break;
      default: intr0 = 0; base0 = 2 * C; set0 = 2 * C + 2 * B; break;
    }
    std::vector<int32_t> b;
    for (int k = 0; k < C; ++k) { b.push_back(intr0 + 2 * k); b.push_back(intr0 + 2 * k + 1); }
    for (int j = 0; j < B; ++j) { b.push_back(base0 + 2 * j); b.push_back(base0 + 2 * j + 1); }
    for (int v = 0; v < S; ++v) { b.push_back(set0 + 2 * v); b.push_back(set0 + 2 * v + 1); }
    return b;
  }
  static void readInto(const aslam::backend::DesignVariable* v, double* dst, size_t n) {
    Eigen::MatrixXd m;
    v->getParameters(m);
    const size_t have = (size_t)m.size();
    for (size_t i = 0; i < n; ++i) dst[i] = i < have ? m.data()[i] : 0.0;
  }
  // host design variables -> the library's packed state arrays (DesignVariable::getParameters)
  void gatherState() {
    const ProblemRecorder& r = _rec;
    const size_t C = (size_t)r.n_cams(), S = (size_t)r.n_sets();
    _cam.assign(C * KB_CAM_PARAM_STRIDE, 0.0);
    _base.assign((C > 0 ? C - 1 : 0) * KB_POSE_STRIDE, 0.0);
    _sets.assign(S * KB_POSE_STRIDE, 0.0);
    for (size_t k = 0; k < C; ++k) {
      Eigen::MatrixXd p, dd;
      r.projection[k]->getParameters(p);
      r.distortion[k]->getParameters(dd);
      double* o = &_cam[k * KB_CAM_PARAM_STRIDE];
      for (std::ptrdiff_t i = 0; i < p.size(); ++i) o[i] = p.data()[i];
      for (std::ptrdiff_t i = 0; i < dd.size(); ++i) o[p.size() + i] = dd.data()[i];
    }
    for (size_t j = 0; j + 1 < C; ++j) {
      readInto(r.baseline_q[j], &_base[j * KB_POSE_STRIDE], 4);
      readInto(r.baseline_t[j], &_base[j * KB_POSE_STRIDE + 4], 3);
    }
    for (size_t v = 0; v < S; ++v) {
      readInto(r.set_q[v], &_sets[v * KB_POSE_STRIDE], 4);
      readInto(r.set_t[v], &_sets[v * KB_POSE_STRIDE + 4], 3);
    }
  }
  void pushStateIfChanged() {
    gatherState();
    const bool c = _cam != _pushed_cam, b = _base != _pushed_base, s = _sets != _pushed_sets;
    if (!c && !b && !s) return;
    check(kb_set_state(_h, c ? _cam.data() : nullptr, b ? _base.data() : nullptr, s ? _sets.data() : nullptr));
    if (c) _pushed_cam = _cam;
    if (b) _pushed_base = _base;
    if (s) _pushed_sets = _sets;
  }

  const ProblemRecorder& _rec;
  int _device;
  kb_handle* _h = nullptr;
  std::vector<double> _cam, _base, _sets, _pushed_cam, _pushed_base, _pushed_sets, _batch;
  bool _residualsStale = true, _batchUseM = true;
  mutable bool _eFetched = false;
};

// One reprojection term of the problem as an unmodified Optimizer2 needs to see it: it knows its design variables (for the
// optimiser's bookkeeping) and returns the residual the device computed for it.  Jacobians are never asked for: buildSystem is
// the solver's.  ≙ the role of aslam::ReprojectionError (aslam_cv/aslam_cv_error_terms/.../ReprojectionError.hpp:27-75) on this path
class ReprojectionErrorProxy : public aslam::backend::ErrorTerm {
 public:
  ReprojectionErrorProxy(SchurLinearSystemSolver* solver, size_t term, const std::vector<aslam::backend::DesignVariable*>& dvs)
      : _solver(solver), _term(term) {
    setDesignVariables(dvs);
  }
  void getWeightedJacobians(aslam::backend::JacobianContainer&, bool) override { throw aslam::Exception("b200: Jacobians stay on the device"); }
  void getWeightedError(Eigen::VectorXd& e, bool useMEstimator) const override {
    const double* w = _solver->weightedError(_term, useMEstimator);  // the device holds -e (LinearSystemSolver.cpp:21)
    e.resize(2);
    e[0] = -w[0];
    e[1] = -w[1];
  }
  void getInvR(Eigen::MatrixXd& invR) const override { invR = vsInvR(); }
  Eigen::MatrixXd vsInvR() const override {
    Eigen::MatrixXd m(2, 2);
    double s[4];
    kb_get_sqrt_inv_r(_solver->handle(), s);  // row-major S with invR = S S^T
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) m(i, j) = s[i * 2] * s[j * 2] + s[i * 2 + 1] * s[j * 2 + 1];
    return m;
  }
  void vsSetInvR(const Eigen::MatrixXd&) override { throw aslam::Exception("b200: invR is a property of the whole problem (kb_set_inv_r)"); }

 protected:
  double evaluateErrorImplementation() override {  // w e^T invR e = |weighted error|^2 (Optimizer2 always evaluates with the M-estimator on)
    const double* w = _solver->weightedError(_term, true);
    return w[0] * w[0] + w[1] * w[1];
  }
  void evaluateJacobiansImplementation(aslam::backend::JacobianContainer&) const override { throw aslam::Exception("b200: Jacobians stay on the device"); }
  size_t getDimensionImplementation() const override { return 2; }
  void buildHessianImplementation(aslam::backend::SparseBlockMatrix&, Eigen::VectorXd&, bool) override { throw aslam::Exception("b200: the Hessian is built on the device"); }
  Eigen::VectorXd vsErrorImplementation() const override {
    Eigen::VectorXd e;
    getWeightedError(e, false);
    return e;
  }

 private:
  SchurLinearSystemSolver* _solver;
  size_t _term;
};

}  // namespace b200
}  // namespace kalibr2
