// Host-side C++ mirror of aslam::calibration::IncrementalEstimator over the C ABI (kalibr_b200.h): the loop kalibr2_ros actually
// runs (aslam_offline_calibration/kalibr2_ros/src/CalibrateCameras.cpp:279-304) — every synced set is offered as a batch, the
// whole problem is re-optimised with Gauss-Newton over the truncated-SVD solver, and the batch is kept only if it adds information
// about the calibration parameters.  Same names, argument meaning and decisions as
//   IncrementalEstimator::addBatch / Options / ReturnValue   IC/src/core/IncrementalEstimator.cpp:338-540,
//                                                            IC/include/aslam/calibration/core/IncrementalEstimator.h:80-140
//   kalibr2::tools::CreateBatchProblem                       K2/include/kalibr2/CalibrationTools.hpp:460-521
// (IC = aslam_incremental_calibration/incremental_calibration).  Like the reference, the estimator keeps ONE growing problem: a live
// device handle in the design-variable order of the merged incremental problem (KB_ORDER_BATCH: set poses, baselines, intrinsics -
// the calibration block is the last n_c columns) to which a batch is appended (kb_append_set: only the new observations travel) and
// from which a rejected batch is removed again (kb_remove_last_set + kb_restore_design_variables ≙ restoreDesignVariables +
// IncrementalOptimizationProblem::remove); all numerics - optimisation, marginal analysis - run on the device.
// Two behaviours of the reference are kept on purpose: analyzeMarginal() sees the Jacobian of the LAST Gauss-Newton iteration (one
// update behind the final state), and after a solve it keeps that solve's rank — the rank of the column-SCALED system — for the
// log2 sum over the UNSCALED singular values (IC/src/core/LinearSolver.cpp:517-523, 196-200).
#pragma once
#include <cmath>
#include <limits>
#include <memory>

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
  // row-major dense matrix (Eigen::MatrixXd of the reference's ReturnValue)
  struct Matrix {
    std::ptrdiff_t rows = 0, cols = 0;
    std::vector<double> data;
    double operator()(std::ptrdiff_t r, std::ptrdiff_t c) const { return data[(size_t)(r * cols + c)]; }
  };
  struct ReturnValue {  // IncrementalEstimator.h:97-140.  Rows of the bases / covariances: the calibration block in the reference's
                        // column order (baselines q,t ..., then per camera projection, distortion)
    bool batchAccepted = false;
    double informationGain = 0.0;
    std::ptrdiff_t rankPsi = -1, rankPsiDeficiency = -1;      // the eliminated pose block (6 per synced set; full rank or the solve fails)
    std::ptrdiff_t rankTheta = -1, rankThetaDeficiency = -1;
    double svdTolerance = 0.0, qrTolerance = -1.0;            // no QR here: the poses are eliminated by the Schur complement
    Matrix nobsBasis, obsBasis;                               // ≙ getNullSpace / getRowSpace of the unscaled marginal system
    Matrix sigma2Theta;                                       // ≙ getCovariance: V_r S_r^-1 V_r^T
    std::vector<double> sigma2ThetaObs;                       // ≙ getRowSpaceCovariance (diagonal): 1 / singular value
    std::vector<double> singularValues;
    Matrix nobsBasisScaled, obsBasisScaled, sigma2ThetaScaled;  // the same of the column-scaled system of the last solve (columnScaling on)
    std::vector<double> sigma2ThetaObsScaled, singularValuesScaled;
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
    // ---- _problem->add(batch): the live problem grows by one synced set ----
    Problem one;  // the batch, flattened (also the whole problem when it is the first one)
    for (const tools::Camera& c : _cameras) one.addCamera(c);
    for (const tools::Transformation& b : _baselines) one.addPose(one.baselines, b);
    one.n_sets = 1;
    for (size_t k = 0; k < batch.size(); ++k)
      if (batch[k]) one.addView(0, (int)k, *batch[k]);
    one.addPose(one.set_poses, T_tc_guess);
    if (!_handle) {
      _handle = std::make_unique<Handle>(one, KB_ORDER_BATCH, _target);
      _handle->check(kb_save_design_variables(_handle->get()));
    } else {
      _handle->check(kb_save_design_variables(_handle->get()));  // saveDesignVariables (restored when the batch is rejected)
      _handle->check(kb_append_set(_handle->get(), (int32_t)one.view_cam.size(), one.view_cam.data(), one.view_begin.data(), one.y_u.data(),
                                   one.y_v.data(), one.corner_id.data(), one.set_poses.data()));
    }
    Handle& h = *_handle;
    const size_t n_sets = _batches.size() + 1;
    kb_solution sol;
    h.check(kb_optimize_gauss_newton(h.get(), &_optimizerOptions, &_solverOptions, &sol));
    const int n = numCalibrationParameters(one);
    const std::ptrdiff_t margStart = (std::ptrdiff_t)kb_jcols(h.get()) - n;  // setMargStartIndex(JCols - dim)
    ReturnValue ret;
    std::vector<double> V((size_t)n * n);
    std::vector<int32_t> cols((size_t)n);
    kb_svd_solve_result last;
    h.check(kb_get_last_svd_solve(h.get(), &last));
    if (_solverOptions.column_scaling && last.rank >= 0) {  // "grep the scaled singular values if scaling enabled"
      ret.singularValuesScaled.resize((size_t)n);
      h.check(kb_get_last_svd_decomposition(h.get(), ret.singularValuesScaled.data(), V.data(), cols.data()));
      fillSpaces(V, cols, margStart, ret.singularValuesScaled, last.rank, ret.nobsBasisScaled, ret.obsBasisScaled, ret.sigma2ThetaScaled,
                 ret.sigma2ThetaObsScaled);
    }
    // analyzeMarginal(): unscaled system of the last iteration's Jacobian; rank / tolerance of the last (scaled) solve
    kb_marginal_options mo;
    kb_default_marginal_options(&mo);
    mo.eps_svd = _solverOptions.eps_svd;
    mo.svd_tol = _solverOptions.svd_tol;
    kb_marginal_result mres;
    ret.singularValues.resize((size_t)n);
    h.check(kb_analyze_marginal_last_build(h.get(), &mo, &mres, ret.singularValues.data(), V.data(), cols.data()));
    const std::ptrdiff_t rank = last.rank >= 0 ? last.rank : mres.rank;
    ret.rankTheta = rank;
    ret.rankThetaDeficiency = (std::ptrdiff_t)n - rank;
    ret.svdTolerance = last.rank >= 0 ? last.tolerance : mres.tolerance;
    ret.rankPsi = (std::ptrdiff_t)(6 * n_sets);
    ret.rankPsiDeficiency = 0;
    fillSpaces(V, cols, margStart, ret.singularValues, rank, ret.nobsBasis, ret.obsBasis, ret.sigma2Theta, ret.sigma2ThetaObs);
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
      // the optimised design variables stay (they live on the device; the host copies are for the accessors)
      std::vector<tools::Camera*> cams;
      for (tools::Camera& c : _cameras) cams.push_back(&c);
      h.readCameras(cams);
      std::vector<double> b(_baselines.size() * KB_POSE_STRIDE), sp(n_sets * KB_POSE_STRIDE);
      if (!b.empty()) h.check(kb_get_baselines(h.get(), b.data()));
      h.check(kb_get_set_poses(h.get(), sp.data()));
      for (size_t j = 0; j < _baselines.size(); ++j) unpack(&b[j * KB_POSE_STRIDE], _baselines[j]);
      _batches.push_back(batch);
      _poses.resize(n_sets);
      for (size_t s = 0; s < _poses.size(); ++s) unpack(&sp[s * KB_POSE_STRIDE], _poses[s]);
    } else {
      // restoreDesignVariables + _problem->remove(batch): everything as it was before the batch
      h.check(kb_remove_last_set(h.get()));
      h.check(kb_restore_design_variables(h.get()));
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
  // null / row space, covariance and row-space covariance of a decomposition (V rows in reduced-system order; `columns` gives their
  // design-variable columns, which the calibration block's start turns into the reference's row order)
  static void fillSpaces(const std::vector<double>& V, const std::vector<int32_t>& columns, std::ptrdiff_t margStart, const std::vector<double>& sv,
                         std::ptrdiff_t rank, Matrix& nullSpace, Matrix& rowSpace, Matrix& covariance, std::vector<double>& rowSpaceCovariance) {
    const std::ptrdiff_t n = (std::ptrdiff_t)sv.size();
    if (rank < 0 || rank > n) return;
    auto cut = [&](std::ptrdiff_t c0, std::ptrdiff_t c1, Matrix& M) {
      M.rows = n;
      M.cols = c1 - c0;
      M.data.assign((size_t)(n * M.cols), 0.0);
      for (std::ptrdiff_t i = 0; i < n; ++i)
        for (std::ptrdiff_t c = c0; c < c1; ++c) M.data[(size_t)((columns[(size_t)i] - margStart) * M.cols + (c - c0))] = V[(size_t)(i * n + c)];
    };
    cut(0, rank, rowSpace);
    cut(rank, n, nullSpace);
    rowSpaceCovariance.resize((size_t)rank);
    for (std::ptrdiff_t k = 0; k < rank; ++k) rowSpaceCovariance[(size_t)k] = 1.0 / sv[(size_t)k];
    covariance.rows = covariance.cols = n;
    covariance.data.assign((size_t)(n * n), 0.0);
    for (std::ptrdiff_t r = 0; r < n; ++r)
      for (std::ptrdiff_t c = 0; c < n; ++c) {
        double a = 0.0;
        for (std::ptrdiff_t k = 0; k < rank; ++k) a += rowSpace(r, k) * rowSpaceCovariance[(size_t)k] * rowSpace(c, k);
        covariance.data[(size_t)(r * n + c)] = a;
      }
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
  std::unique_ptr<tools::detail::Handle> _handle;  // the live problem on the device
  std::vector<tools::SyncedSet> _batches;
  std::vector<tools::Transformation> _poses;
  double _informationGain = 0.0, _svLog2Sum = 0.0, _initialCost = 0.0, _finalCost = 0.0;
  std::ptrdiff_t _rankTheta = -1, _rankThetaDeficiency = -1;
  std::vector<double> _singularValues;
};

}  // namespace calibration
}  // namespace kalibr_b200
