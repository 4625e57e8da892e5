// Runs include/kalibr_b200/reference_adapter.hpp the way a Kalibr2 tree would: a HOST optimiser in the role of
// aslam::backend::Optimizer2 owns the design variables (rotation quaternions, Euclidean points, camera parameter blocks), assigns
// block indices / column bases, applies and reverts state updates ON THE HOST (BE/src/Optimizer2.cpp:95-151, 290-318), and reaches
// the GPU only through the aslam::backend::LinearSystemSolver virtuals of the adapter.  The reference's interfaces come from the
// stand-in under tests/cpp/aslam_mock (same include paths / signatures).  Built twice by tests/test_drivers_gpu.py: as is
// (evaluateError non-virtual: the error-term proxies feed the base class's loop) and with -DKB_REFERENCE_HAS_VIRTUAL_EVALUATE_ERROR.
//   reference_adapter_main <problem.bin> <driver order 0..3>
// prints "solution iterations failed solver_failure J_start J_final", the final camera parameters and baselines.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>

#include "kalibr_b200/lm_state_machine.h"
#include "kalibr_b200/reference_adapter.hpp"

using aslam::backend::DesignVariable;
using aslam::backend::ErrorTerm;

namespace {

template <typename T>
std::vector<T> readArray(std::ifstream& f) {
  int64_t n = 0;
  f.read(reinterpret_cast<char*>(&n), sizeof(n));
  std::vector<T> v((size_t)n);
  f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * (size_t)n);
  return v;
}

// ---- host design variables, in the roles of the reference's classes ----
// a parameter block updated by addition: EuclideanPoint (BX/src/EuclideanPoint.cpp:23-32) and the projection / distortion blocks behind
// DesignVariableAdapter (BE/include/aslam/backend/implementation/DesignVariableAdapter.hpp:42-55)
class AdditiveBlock : public DesignVariable {
 public:
  AdditiveBlock(const double* p, int n) : _p(p, p + n), _backup(_p) {}
  const std::vector<double>& values() const { return _p; }

 protected:
  int minimalDimensionsImplementation() const override { return (int)_p.size(); }
  void updateImplementation(const double* dp, int size) override {
    _backup = _p;
    for (int i = 0; i < size && i < (int)_p.size(); ++i) _p[(size_t)i] += dp[i];
  }
  void revertUpdateImplementation() override { _p = _backup; }
  void getParametersImplementation(Eigen::MatrixXd& v) const override {
    v.resize((std::ptrdiff_t)_p.size(), 1);
    for (size_t i = 0; i < _p.size(); ++i) v((std::ptrdiff_t)i, 0) = _p[i];
  }
  void setParametersImplementation(const Eigen::MatrixXd& v) override {
    _backup = _p;
    for (size_t i = 0; i < _p.size(); ++i) _p[i] = v((std::ptrdiff_t)i, 0);
  }

 private:
  std::vector<double> _p, _backup;
};
// RotationQuaternion (BX/src/RotationQuaternion.cpp:22-36): q <- updateQuat(q, dq), scalar-last JPL convention
// (Schweizer-Messer/sm_kinematics/src/quaternion_algebra.cpp:200-219, 302-317)
class QuaternionBlock : public DesignVariable {
 public:
  explicit QuaternionBlock(const double* q) { for (int i = 0; i < 4; ++i) _q[i] = _b[i] = q[i]; }
  const double* q() const { return _q; }

 protected:
  int minimalDimensionsImplementation() const override { return 3; }
  void updateImplementation(const double* dq, int) override {
    for (int i = 0; i < 4; ++i) _b[i] = _q[i];
    const double theta = std::sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
    const double na = theta < std::sqrt(std::sqrt(2.220446049250313e-16)) ? 0.5 + theta * theta * (1.0 / 48.0) : std::sin(theta * 0.5) / theta;
    const double d0 = dq[0] * na, d1 = dq[1] * na, d2 = dq[2] * na, ca = std::cos(theta * 0.5);
    const double q0 = _q[0], q1 = _q[1], q2 = _q[2], q3 = _q[3];
    _q[0] = q0 * ca + d0 * q3 - d1 * q2 + d2 * q1;
    _q[1] = q1 * ca + d0 * q2 + d1 * q3 - d2 * q0;
    _q[2] = q2 * ca - d0 * q1 + d1 * q0 + d2 * q3;
    _q[3] = q3 * ca - d0 * q0 - d1 * q1 - d2 * q2;
  }
  void revertUpdateImplementation() override { for (int i = 0; i < 4; ++i) _q[i] = _b[i]; }
  void getParametersImplementation(Eigen::MatrixXd& v) const override {
    v.resize(4, 1);
    for (int i = 0; i < 4; ++i) v(i, 0) = _q[i];
  }
  void setParametersImplementation(const Eigen::MatrixXd& v) override {
    for (int i = 0; i < 4; ++i) { _b[i] = _q[i]; _q[i] = v(i, 0); }
  }

 private:
  double _q[4], _b[4];
};

// ---- the host optimiser: Optimizer2's role.  The scalar policy is the shared state machine; everything that touches design
// variables happens here, on the host. ----
class HostOptimizer {
 public:
  std::vector<DesignVariable*> problemDvs;  // insertion order of the driver
  std::vector<ErrorTerm*> problemTerms;
  std::shared_ptr<aslam::backend::LinearSystemSolver> solver;
  double lambdaInit = 10.0, convDx = 1e-3, convDj = 1.0;  // kalibr2::tools::CreateDefaultOptimizer (CalibrationTools.hpp:57-66)
  int maxIterations = 200, nThreads = 4;
  kalibr_b200::LmState c{};

  void initialize() {  // Optimizer2.cpp:95-151
    _dvs.clear();
    int col = 0;
    for (DesignVariable* dv : problemDvs) {
      if (!dv->isActive()) continue;
      dv->setBlockIndex((int)_dvs.size());
      dv->setColumnBase(col);
      col += dv->minimalDimensions();
      _dvs.push_back(dv);
    }
    size_t row = 0;
    for (ErrorTerm* t : problemTerms) { t->setRowBase(row); row += t->dimension(); }
    solver->initMatrixStructure(_dvs, problemTerms, true);
  }
  double applyStateUpdate(const Eigen::VectorXd& dx) {  // Optimizer2.cpp:290-307
    double m = 0.0;
    for (DesignVariable* dv : _dvs) {
      const int n = dv->minimalDimensions();
      dv->update(dx.data() + dv->columnBase(), n);
      for (int i = 0; i < n; ++i) m = std::fmax(m, std::fabs(dx[dv->columnBase() + i]));
    }
    return m;
  }
  void revertLastStateUpdate() { for (DesignVariable* dv : _dvs) dv->revertUpdate(); }
  void optimize() {
    using namespace kalibr_b200;
    initialize();
    const double J0 = solver->evaluateError((size_t)nThreads, true);
    lm_start(&c, KB_POLICY_LEVENBERG_MARQUARDT, J0, lambdaInit, convDx, convDj, maxIterations, 1);
    Eigen::VectorXd dx;
    while (!c.done) {
      lm_before_solve(&c);
      if (c.need_build) solver->buildSystem((size_t)nThreads, true);
      solver->setConstantConditioner(c.lambda);
      const bool ok = solver->solveSystem(dx);
      if (!ok) { lm_after_solve(&c, 1.0, 0.0, 0); continue; }
      // getLmRho's denominator, the reference's way: from dx and rhs() on the host (LevenbergMarquardtTrustRegionPolicy.cpp:107-113)
      const Eigen::VectorXd& rhs = solver->rhs();
      double den = 0.0;
      for (std::ptrdiff_t i = 0; i < dx.size(); ++i) den += dx[i] * (c.lambda * dx[i] + rhs[i]);
      const double maxDx = applyStateUpdate(dx);
      lm_after_solve(&c, den, maxDx, 1);
      const double J = solver->evaluateError((size_t)nThreads, true);
      lm_after_eval(&c, J);
      if (c.revert) revertLastStateUpdate();
    }
  }

 private:
  std::vector<DesignVariable*> _dvs;
};

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  if (!f) return 3;
  const int order = std::atoi(argv[2]);
  auto dims = readArray<int32_t>(f);  // n_cams, n_sets, rows, cols
  auto models = readArray<int32_t>(f);
  auto res = readArray<int32_t>(f);
  auto params = readArray<double>(f);
  auto baselines = readArray<double>(f);
  auto points = readArray<double>(f);
  auto view_set = readArray<int32_t>(f);
  auto view_cam = readArray<int32_t>(f);
  auto view_begin = readArray<int64_t>(f);
  auto corner = readArray<int32_t>(f);
  auto yu = readArray<double>(f);
  auto yv = readArray<double>(f);
  auto set_poses = readArray<double>(f);
  const int C = dims[0], S = dims[1];
  static const int P_of[7] = {4, 4, 5, 6, 6, 4, 5}, D_of[7] = {4, 4, 4, 0, 0, 1, 0};
  try {
    // ---- the driver's part: design variables and error terms, in the driver's insertion order ----
    std::vector<std::unique_ptr<AdditiveBlock>> proj, dist, base_t, set_t;
    std::vector<std::unique_ptr<QuaternionBlock>> base_q, set_q;
    for (int k = 0; k < C; ++k) {
      const double* p = &params[(size_t)k * KB_CAM_PARAM_STRIDE];
      proj.emplace_back(new AdditiveBlock(p, P_of[models[(size_t)k]]));
      dist.emplace_back(new AdditiveBlock(p + P_of[models[(size_t)k]], D_of[models[(size_t)k]]));  // 0-dim but active for NoDistortion
    }
    for (int j = 0; j + 1 < C; ++j) {
      base_q.emplace_back(new QuaternionBlock(&baselines[(size_t)j * 7]));
      base_t.emplace_back(new AdditiveBlock(&baselines[(size_t)j * 7 + 4], 3));
    }
    for (int v = 0; v < S; ++v) {
      set_q.emplace_back(new QuaternionBlock(&set_poses[(size_t)v * 7]));
      set_t.emplace_back(new AdditiveBlock(&set_poses[(size_t)v * 7 + 4], 3));
    }
    HostOptimizer opt;
    kalibr2::b200::ProblemRecorder rec;
    rec.driver_order = order;
    rec.cam_model.assign(models.begin(), models.end());
    rec.target_points = points;
    auto addIntr = [&](int k) { opt.problemDvs.push_back(proj[(size_t)k].get()); opt.problemDvs.push_back(dist[(size_t)k].get()); };
    auto addBase = [&]() { for (int j = 0; j + 1 < C; ++j) { opt.problemDvs.push_back(base_q[(size_t)j].get()); opt.problemDvs.push_back(base_t[(size_t)j].get()); } };
    auto addSets = [&]() { for (int v = 0; v < S; ++v) { opt.problemDvs.push_back(set_q[(size_t)v].get()); opt.problemDvs.push_back(set_t[(size_t)v].get()); } };
    switch (order) {
      case KB_ORDER_SINGLE: addIntr(0); addSets(); break;
      case KB_ORDER_STEREO: addBase(); addSets(); addIntr(0); addIntr(1); break;
      case KB_ORDER_BATCH: addSets(); addBase(); for (int k = 0; k < C; ++k) addIntr(k); break;
      default: for (int k = 0; k < C; ++k) addIntr(k); addBase(); addSets(); break;
    }
    for (DesignVariable* dv : opt.problemDvs) dv->setActive(true);
    for (int k = 0; k < C; ++k) { rec.projection.push_back(proj[(size_t)k].get()); rec.distortion.push_back(dist[(size_t)k].get()); }
    for (int j = 0; j + 1 < C; ++j) { rec.baseline_q.push_back(base_q[(size_t)j].get()); rec.baseline_t.push_back(base_t[(size_t)j].get()); }
    for (int v = 0; v < S; ++v) { rec.set_q.push_back(set_q[(size_t)v].get()); rec.set_t.push_back(set_t[(size_t)v].get()); }
    auto solver = std::make_shared<kalibr2::b200::SchurLinearSystemSolver>(rec);
    opt.solver = solver;
    std::vector<std::unique_ptr<kalibr2::b200::ReprojectionErrorProxy>> terms;
    for (size_t w = 0; w < view_set.size(); ++w) {
      const int v = view_set[w], k = view_cam[w];
      std::vector<DesignVariable*> dvs = {set_q[(size_t)v].get(), set_t[(size_t)v].get()};
      for (int j = 0; j < k; ++j) { dvs.push_back(base_q[(size_t)j].get()); dvs.push_back(base_t[(size_t)j].get()); }
      dvs.push_back(proj[(size_t)k].get());
      dvs.push_back(dist[(size_t)k].get());
      for (int64_t i = view_begin[w]; i < view_begin[w + 1]; ++i) {
        rec.addTerm(corner[(size_t)i], yu[(size_t)i], yv[(size_t)i]);
        terms.emplace_back(new kalibr2::b200::ReprojectionErrorProxy(solver.get(), terms.size(), dvs));
        opt.problemTerms.push_back(terms.back().get());
      }
      rec.endView(v, k);
    }
    // ---- the optimiser's part ----
    opt.optimize();
    const kalibr_b200::LmState& c = opt.c;
    std::printf("solution %d %d %d %.17g %.17g\n", c.iterations, c.failed, c.solver_failure, c.JStart, c.pJ);
    for (int k = 0; k < C; ++k) {
      std::printf("camera%d", k);
      for (double x : proj[(size_t)k]->values()) std::printf(" %.17g", x);
      for (double x : dist[(size_t)k]->values()) std::printf(" %.17g", x);
      std::printf("\n");
    }
    for (int j = 0; j + 1 < C; ++j) {
      const double* q = base_q[(size_t)j]->q();
      const std::vector<double>& t = base_t[(size_t)j]->values();
      std::printf("baseline%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", j, q[0], q[1], q[2], q[3], t[0], t[1], t[2]);
    }
    std::printf("launches %lld\n", (long long)kb_kernel_launches(solver->handle()));
  } catch (const std::exception& e) {
    std::printf("error %s\n", e.what());
    return 1;
  }
  return 0;
}
