// Host-side C++ mirror of aslam::calibration::IncrementalEstimator over the C ABI (kalibr_b200.h): the loop kalibr2_ros actually
// runs (aslam_offline_calibration/kalibr2_ros/src/CalibrateCameras.cpp:279-304) — every synced set is offered as a batch, the
// whole problem is re-optimised with Gauss-Newton over the truncated-SVD solver, and the batch is kept only if it adds information
// about the calibration parameters.  Same names, argument meaning and decisions as
//   IncrementalEstimator::addBatch / Options / ReturnValue   IC/src/core/IncrementalEstimator.cpp:338-540,
//                                                            IC/include/aslam/calibration/core/IncrementalEstimator.h:80-140
//   kalibr2::tools::CreateBatchProblem                       K2/include/kalibr2/CalibrationTools.hpp:460-521
// (IC = aslam_incremental_calibration/incremental_calibration).  The reference mutates one growing OptimizationProblem; here the
// accepted batches are kept flattened and the device problem is rebuilt per batch (the reference re-initialises its matrix
// structure per batch as well), all numerics — optimisation, marginal analysis — run on the device.
// Two behaviours of the reference are kept on purpose: analyzeMarginal() sees the Jacobian of the LAST Gauss-Newton iteration (one
// update behind the final state), and after a solve it keeps that solve's rank — the rank of the column-SCALED system — for the
// log2 sum over the UNSCALED singular values (IC/src/core/LinearSolver.cpp:517-523, 196-200).
#pragma once
#include <cmath>
#include <limits>

#include "calibration_tools.hpp"

namespace kalibr_b200 {
namespace calibration {

class IncrementalEstimator {
 public:
  struct Options {  // IncrementalEstimator.h:80-95
    double infoGainDelta = 0.2;
    bool checkValidity = false;
    bool verbose = false;
  };
  struct ReturnValue {  // IncrementalEstimator.h:97-140 (the matrices stay on request: singular values only)
    bool batchAccepted = false;
    double informationGain = 0.0;
    std::ptrdiff_t rankTheta = -1, rankThetaDeficiency = -1;
    double svdTolerance = 0.0;
    std::vector<double> singularValues;
    size_t numIterations = 0;
    double JStart = 0.0, JFinal = 0.0;
  };

  // kalibr2_ros' settings by default: column scaling with epsSVD = 1e-6, at most 20 Gauss-Newton iterations (CalibrateCameras.cpp:259-272)
  IncrementalEstimator(std::vector<tools::Camera> cameras, std::vector<tools::Transformation> baselines, tools::Target target)
      : IncrementalEstimator(std::move(cameras), std::move(baselines), std::move(target), Options{}) {}
  IncrementalEstimator(std::vector<tools::Camera> cameras, std::vector<tools::Transformation> baselines, tools::Target target, Options options)
      : _cameras(std::move(cameras)), _baselines(std::move(baselines)), _target(std::move(target)), _options(options) {
    kb_default_svd_solver_options(&_solverOptions);
    _solverOptions.column_scaling = 1;
    _solverOptions.eps_svd = 1e-6;
    kb_default_optimizer_options(&_optimizerOptions);
    _optimizerOptions.convergence_delta_x = 1e-3;
    _optimizerOptions.convergence_delta_j = 1e-3;
    _optimizerOptions.max_iterations = 20;
  }
  kb_svd_solver_options& linearSolverOptions() { return _solverOptions; }
  kb_optimizer_options& optimizerOptions() { return _optimizerOptions; }
  Options& getOptions() { return _options; }

  // ≙ CreateBatchProblem + IncrementalEstimator::addBatch: one synced set with its target-pose guess (getTargetPoseGuess)
  ReturnValue addBatch(const tools::SyncedSet& batch, const tools::Transformation& T_tc_guess, bool force = false) {
    using namespace tools::detail;
    Problem p;
    for (const tools::Camera& c : _cameras) p.addCamera(c);
    for (const tools::Transformation& b : _baselines) p.addPose(p.baselines, b);
    p.n_sets = (int32_t)_batches.size() + 1;
    for (size_t s = 0; s <= _batches.size(); ++s) {
      const tools::SyncedSet& set = s < _batches.size() ? _batches[s] : batch;
      for (size_t k = 0; k < set.size(); ++k)
        if (set[k]) p.addView((int)s, (int)k, *set[k]);
      p.addPose(p.set_poses, s < _batches.size() ? _poses[s] : T_tc_guess);
    }
    Handle h(p, KB_ORDER_RIG, _target);
    kb_solution sol;
    h.check(kb_optimize_gauss_newton(h.get(), &_optimizerOptions, &_solverOptions, &sol));
    // analyzeMarginal(): unscaled system of the last iteration's Jacobian; rank / tolerance of the last (scaled) solve
    kb_marginal_options mo;
    kb_default_marginal_options(&mo);
    mo.eps_svd = _solverOptions.eps_svd;
    mo.svd_tol = _solverOptions.svd_tol;
    kb_marginal_result mres;
    ReturnValue ret;
    ret.singularValues.resize((size_t)numCalibrationParameters(p));
    h.check(kb_analyze_marginal_last_build(h.get(), &mo, &mres, ret.singularValues.data(), nullptr, nullptr));
    kb_svd_solve_result last;
    h.check(kb_get_last_svd_solve(h.get(), &last));
    const std::ptrdiff_t rank = last.rank >= 0 ? last.rank : mres.rank;
    ret.rankTheta = rank;
    ret.rankThetaDeficiency = (std::ptrdiff_t)ret.singularValues.size() - rank;
    ret.svdTolerance = last.rank >= 0 ? last.tolerance : mres.tolerance;
    ret.numIterations = (size_t)sol.iterations;
    ret.JStart = sol.j_start;
    ret.JFinal = sol.j_final;
    double svLog2Sum = 0.0;  // LinearSolver::getSingularValuesLog2Sum
    for (std::ptrdiff_t i = 0; i < rank; ++i) svLog2Sum += std::log2(ret.singularValues[(size_t)i]);
    bool solutionValid = true;
    if (_options.checkValidity && (sol.iterations == _optimizerOptions.max_iterations || sol.j_final >= sol.j_start)) solutionValid = false;
    ret.informationGain = 0.5 * (svLog2Sum - _svLog2Sum);
    const bool keepBatch = ((ret.informationGain > _options.infoGainDelta || ret.rankTheta > _rankTheta) && solutionValid) || force;
    if (keepBatch) {
      _informationGain = ret.informationGain;
      _svLog2Sum = svLog2Sum;
      _rankTheta = ret.rankTheta;
      _rankThetaDeficiency = ret.rankThetaDeficiency;
      _singularValues = ret.singularValues;
      _initialCost = sol.j_start;
      _finalCost = sol.j_final;
      // the optimised design variables stay; a rejected batch leaves everything as it was (restoreDesignVariables)
      std::vector<tools::Camera*> cams;
      for (tools::Camera& c : _cameras) cams.push_back(&c);
      h.readCameras(cams);
      std::vector<double> b(_baselines.size() * KB_POSE_STRIDE), sp((size_t)p.n_sets * KB_POSE_STRIDE);
      if (!b.empty()) h.check(kb_get_baselines(h.get(), b.data()));
      h.check(kb_get_set_poses(h.get(), sp.data()));
      for (size_t j = 0; j < _baselines.size(); ++j) unpack(&b[j * KB_POSE_STRIDE], _baselines[j]);
      _batches.push_back(batch);
      _poses.resize((size_t)p.n_sets);
      for (size_t s = 0; s < _poses.size(); ++s) unpack(&sp[s * KB_POSE_STRIDE], _poses[s]);
    }
    ret.batchAccepted = keepBatch;
    return ret;
  }

  size_t getNumBatches() const { return _batches.size(); }
  const std::vector<tools::Camera>& cameras() const { return _cameras; }
  const std::vector<tools::Transformation>& baselines() const { return _baselines; }
  const std::vector<tools::Transformation>& targetPoses() const { return _poses; }
  double getInformationGain() const { return _informationGain; }
  std::ptrdiff_t getRankTheta() const { return _rankTheta; }
  std::ptrdiff_t getRankThetaDeficiency() const { return _rankThetaDeficiency; }
  const std::vector<double>& getSingularValues() const { return _singularValues; }
  double getInitialCost() const { return _initialCost; }
  double getFinalCost() const { return _finalCost; }

 private:
  static void unpack(const double* p7, tools::Transformation& T) {
    for (int i = 0; i < 4; ++i) T.q[i] = p7[i];
    for (int i = 0; i < 3; ++i) T.t[i] = p7[4 + i];
  }
  static int numCalibrationParameters(const tools::detail::Problem& p) {
    static const int P[KB_NUM_MODELS] = {4, 4, 5, 6, 6, 4, 5}, D[KB_NUM_MODELS] = {4, 4, 4, 0, 0, 1, 0};
    int n = 6 * ((int)p.cam_model.size() - 1);
    for (int32_t m : p.cam_model) n += P[m] + D[m];
    return n;
  }
  std::vector<tools::Camera> _cameras;
  std::vector<tools::Transformation> _baselines;
  tools::Target _target;
  Options _options;
  kb_svd_solver_options _solverOptions;
  kb_optimizer_options _optimizerOptions;
  std::vector<tools::SyncedSet> _batches;
  std::vector<tools::Transformation> _poses;
  double _informationGain = 0.0, _svLog2Sum = 0.0, _initialCost = 0.0, _finalCost = 0.0;
  std::ptrdiff_t _rankTheta = -1, _rankThetaDeficiency = -1;
  std::vector<double> _singularValues;
};

}  // namespace calibration
}  // namespace kalibr_b200
