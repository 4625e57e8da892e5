// Host-side C++ mirror of kalibr2's batch drivers over the C ABI (kalibr_b200.h): the callers on either side of the hot path.
// Same function names, argument meaning and return values as
//   kalibr2::tools::CalibrateSingleCamera    K2/include/kalibr2/CalibrationTools.hpp:93-152
//   kalibr2::tools::CalibrateStereoPair      K2/include/kalibr2/CalibrationTools.hpp:183-300
//   kalibr2::tools::getTargetPoseGuess       K2/include/kalibr2/CalibrationTools.hpp:316-356
//   kalibr2::tools::CalibrateMultiCameraRig  K2/include/kalibr2/CalibrationTools.hpp:376-428
// (K2 = aslam_offline_calibration/kalibr2) with the reference's object graph flattened: a GridCalibrationTargetObservation is the
// list of observed corners of one image, a CameraCalibratorBase is (model, image size, parameter vector), sm::kinematics::
// Transformation is (q xyzw, t).  Every numeric step — initializeIntrinsics, estimateTransformation (PnP), the baseline median,
// Optimizer2 with the Levenberg-Marquardt policy — runs on the device behind one kb_handle; this header only assembles the
// problem in the reference's design-variable and error-term order and carries the results back.
#pragma once
#include <cmath>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../kalibr_b200.h"

namespace kalibr_b200 {
namespace tools {

struct Transformation {  // sm::kinematics::Transformation: T_a_b as (q_a_b xyzw scalar last, t_a_b_a)
  double q[4] = {0.0, 0.0, 0.0, 1.0};
  double t[3] = {0.0, 0.0, 0.0};
};

struct Observation {  // the part of aslam::cameras::GridCalibrationTargetObservation this path reads
  std::vector<int32_t> corner_id;  // target point index (row * cols + col) of every observed corner
  std::vector<double> u, v;        // its image coordinates
};
using SyncedSet = std::vector<std::optional<Observation>>;  // one slot per camera (CalibrationTools.hpp: SyncedSet)

struct Camera {  // kalibr2::CameraCalibratorBase: camera_geometry() as model + parameters, calibrated in place
  int32_t model = KB_PINHOLE_RADTAN;
  int32_t ru = 0, rv = 0;
  double params[KB_CAM_PARAM_STRIDE] = {0};
};

struct Target {  // aslam::cameras::GridCalibrationTargetBase
  int32_t rows = 0, cols = 0;
  std::vector<double> points;  // [rows * cols][3]
};

struct ReprojectionStatistics { double n, mean_u, mean_v, std_u, std_v, rmse; };  // CameraCalibrator::PrintReprojectionErrorStatistics

namespace detail {

inline void quat2r(const double* q, double R[9]) {  // Schweizer-Messer/sm_kinematics/src/quaternion_algebra.cpp:77-101, row-major
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = x * x - y * y - z * z + w * w; R[1] = 2 * x * y + 2 * z * w; R[2] = 2 * x * z - 2 * y * w;
  R[3] = 2 * x * y - 2 * z * w; R[4] = -x * x + y * y - z * z + w * w; R[5] = 2 * x * w + 2 * y * z;
  R[6] = 2 * x * z + 2 * y * w; R[7] = -2 * x * w + 2 * y * z; R[8] = -x * x - y * y + z * z + w * w;
}
inline void r2quat(const double R[9], double q[4]) {  // quaternion_algebra.cpp:16-75
  const double c1 = R[0], c2 = R[3], c3 = R[6], c4 = R[1], c5 = R[4], c6 = R[7], c7 = R[2], c8 = R[5], c9 = R[8];
  const double dc[4] = {std::fabs(1.0 + c1 - c5 - c9), std::fabs(1.0 - c1 + c5 - c9), std::fabs(1.0 - c1 - c5 + c9), std::fabs(1.0 + c1 + c5 + c9)};
  int m = 0;
  for (int i = 1; i < 4; ++i) if (dc[i] > dc[m]) m = i;
  double c;
  if (m == 0) { q[0] = 0.5 * std::sqrt(dc[0]); c = 0.25 / q[0]; q[1] = c * (c4 + c2); q[2] = c * (c7 + c3); q[3] = c * (c8 - c6); }
  else if (m == 1) { q[1] = 0.5 * std::sqrt(dc[1]); c = 0.25 / q[1]; q[0] = c * (c4 + c2); q[2] = c * (c6 + c8); q[3] = c * (c3 - c7); }
  else if (m == 2) { q[2] = 0.5 * std::sqrt(dc[2]); c = 0.25 / q[2]; q[0] = c * (c3 + c7); q[1] = c * (c6 + c8); q[3] = c * (c4 - c2); }
  else { q[3] = 0.5 * std::sqrt(dc[3]); c = 0.25 / q[3]; q[0] = c * (c8 - c6); q[1] = c * (c3 - c7); q[2] = c * (c4 - c2); }
  if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
}
// a * inverse(b)
inline Transformation mulInverse(const Transformation& a, const Transformation& b) {
  double Ra[9], Rb[9], R[9];
  quat2r(a.q, Ra);
  quat2r(b.q, Rb);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = Ra[i * 3] * Rb[j * 3] + Ra[i * 3 + 1] * Rb[j * 3 + 1] + Ra[i * 3 + 2] * Rb[j * 3 + 2];  // Ra Rb^T
  Transformation o;
  r2quat(R, o.q);
  for (int i = 0; i < 3; ++i) o.t[i] = a.t[i] - (R[i * 3] * b.t[0] + R[i * 3 + 1] * b.t[1] + R[i * 3 + 2] * b.t[2]);
  return o;
}

// The flattened OptimizationProblem (kb_problem_desc) with the arrays it points into.
struct Problem {
  std::vector<int32_t> cam_model, view_set, view_cam, corner_id, resolution;
  std::vector<double> cam_params, baselines, set_poses, y_u, y_v;
  std::vector<int64_t> view_begin{0};
  int32_t n_sets = 0;
  void addView(int set, int cam, const Observation& o) {
    corner_id.insert(corner_id.end(), o.corner_id.begin(), o.corner_id.end());
    y_u.insert(y_u.end(), o.u.begin(), o.u.end());
    y_v.insert(y_v.end(), o.v.begin(), o.v.end());
    view_set.push_back(set);
    view_cam.push_back(cam);
    view_begin.push_back((int64_t)y_u.size());
  }
  void addCamera(const Camera& c) {
    cam_model.push_back(c.model);
    cam_params.insert(cam_params.end(), c.params, c.params + KB_CAM_PARAM_STRIDE);
    resolution.push_back(c.ru);
    resolution.push_back(c.rv);
  }
  void addPose(std::vector<double>& dst, const Transformation& T) {
    dst.insert(dst.end(), T.q, T.q + 4);
    dst.insert(dst.end(), T.t, T.t + 3);
  }
};

class Handle {  // owns a kb_handle
 public:
  Handle(Problem& p, int driver_order, const Target& target, int device = 0) {
    kb_problem_desc d;
    std::memset(&d, 0, sizeof(d));
    d.driver_order = driver_order;
    d.n_cams = (int32_t)p.cam_model.size();
    d.cam_model = p.cam_model.data();
    d.cam_params = p.cam_params.data();
    d.baselines = p.baselines.data();
    d.n_sets = p.n_sets;
    d.set_poses = p.set_poses.data();
    d.n_target_points = target.rows * target.cols;
    d.target_points = target.points.data();
    d.n_views = (int32_t)p.view_set.size();
    d.view_set = p.view_set.data();
    d.view_cam = p.view_cam.data();
    d.view_begin = p.view_begin.data();
    d.n_terms = (int64_t)p.y_u.size();
    d.y_u = p.y_u.data();
    d.y_v = p.y_v.data();
    d.corner_id = p.corner_id.data();
    d.n_ranks = 1;
    d.device = device;
    if (kb_create(&d, &_h) != KB_OK) throw std::runtime_error(std::string("kb_create: ") + kb_last_error(nullptr));
  }
  ~Handle() { kb_destroy(_h); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
  kb_handle* get() const { return _h; }
  void check(kb_status s) const {
    if (s != KB_OK) throw std::runtime_error(std::string("kalibr_b200: ") + kb_last_error(_h));
  }
  // CreateDefaultOptimizer (CalibrationTools.hpp:57-67) + optimizer.optimize(): returns !linearSolverFailure
  bool optimize(kb_solution* out = nullptr) const {
    kb_optimizer_options o;
    kb_default_optimizer_options(&o);
    kb_solution sol;
    check(kb_optimize(_h, &o, &sol));
    if (out) *out = sol;
    return !sol.linear_solver_failure;
  }
  void readCameras(std::vector<Camera*> cams) const {
    std::vector<double> prm(cams.size() * KB_CAM_PARAM_STRIDE);
    check(kb_get_camera_params(_h, prm.data()));
    for (size_t k = 0; k < cams.size(); ++k) std::memcpy(cams[k]->params, &prm[k * KB_CAM_PARAM_STRIDE], sizeof(cams[k]->params));
  }

 private:
  kb_handle* _h = nullptr;
};

inline void identityPoses(std::vector<double>& dst, int n) {
  dst.clear();
  for (int i = 0; i < n; ++i) { const double p[7] = {0, 0, 0, 1, 0, 0, 0}; dst.insert(dst.end(), p, p + 7); }
}

}  // namespace detail

// ≙ kalibr2::tools::CalibrateSingleCamera: initializeIntrinsics, one estimateTransformation per observation (observations whose
// PnP fails are left out, :117-121), full-batch optimisation.  `camera` is updated in place.  Returns !linearSolverFailure.
inline bool CalibrateSingleCamera(const std::vector<Observation>& observations, Camera& camera, const Target& target,
                                  std::optional<double> fallback_focal_length = std::nullopt, kb_solution* solution = nullptr,
                                  ReprojectionStatistics* statistics = nullptr) {
  using namespace detail;
  auto build = [&](const std::vector<int>& keep, const Camera& cam, const std::vector<double>* poses) {
    Problem p;
    p.addCamera(cam);
    p.n_sets = (int32_t)keep.size();
    for (size_t i = 0; i < keep.size(); ++i) p.addView((int)i, 0, observations[keep[i]]);
    if (poses) p.set_poses = *poses; else identityPoses(p.set_poses, p.n_sets);
    return p;
  };
  std::vector<int> all(observations.size());
  for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
  Problem p0 = build(all, camera, nullptr);
  auto h = std::make_unique<Handle>(p0, KB_ORDER_SINGLE, target);
  int32_t ok_init = 0;
  h->check(kb_initialize_intrinsics(h->get(), 0, target.rows, target.cols, p0.resolution.data(), fallback_focal_length.value_or(0.0), camera.params, &ok_init));
  std::vector<double> T(observations.size() * KB_POSE_STRIDE);
  std::vector<int32_t> ok(observations.size());
  h->check(kb_estimate_transformations(h->get(), p0.resolution.data(), T.data(), ok.data()));
  std::vector<int> keep;
  std::vector<double> poses;
  for (size_t i = 0; i < ok.size(); ++i)
    if (ok[i]) { keep.push_back((int)i); poses.insert(poses.end(), &T[i * KB_POSE_STRIDE], &T[(i + 1) * KB_POSE_STRIDE]); }
  Problem p1 = build(keep, camera, &poses);  // the reference only adds the observations whose PnP succeeded
  h.reset();
  Handle h1(p1, KB_ORDER_SINGLE, target);
  const bool good = h1.optimize(solution);
  h1.readCameras({&camera});
  if (statistics) h1.check(kb_reprojection_statistics(h1.get(), &statistics->n));
  return good;
}

// ≙ kalibr2::tools::CalibrateStereoPair: baseline guess = median of the PnP baselines over the sets both cameras saw, one pose per
// synced set (PnP in L, else PnP in H chained through the guess), joint optimisation in the stereo design-variable order.  Both
// cameras are updated in place; returns T_camH_camL.
inline Transformation CalibrateStereoPair(Camera& camera_L, Camera& camera_H, const std::vector<std::optional<Observation>>& observations_L,
                                          const std::vector<std::optional<Observation>>& observations_H, const Target& target,
                                          kb_solution* solution = nullptr) {
  using namespace detail;
  if (observations_L.size() != observations_H.size()) throw std::runtime_error("The number of observations for both cameras must be the same.");
  // sets without any observation get no pose design variable (:246-258)
  std::vector<int> set_of(observations_L.size(), -1);
  int n_sets = 0;
  for (size_t i = 0; i < observations_L.size(); ++i)
    if (observations_L[i] || observations_H[i]) set_of[i] = n_sets++;
  Problem p;
  p.addCamera(camera_L);
  p.addCamera(camera_H);
  p.n_sets = n_sets;
  for (size_t i = 0; i < observations_L.size(); ++i)  // error terms: all of camera L, then all of camera H (:275-292)
    if (observations_L[i]) p.addView(set_of[i], 0, *observations_L[i]);
  for (size_t i = 0; i < observations_H.size(); ++i)
    if (observations_H[i]) p.addView(set_of[i], 1, *observations_H[i]);
  identityPoses(p.set_poses, n_sets);
  identityPoses(p.baselines, 1);
  Handle h0(p, KB_ORDER_STEREO, target);
  Transformation baseline;
  double b7[7];
  h0.check(kb_estimate_stereo_baseline(h0.get(), p.resolution.data(), 0, 1, b7, nullptr));
  std::memcpy(baseline.q, b7, sizeof(baseline.q));
  std::memcpy(baseline.t, b7 + 4, sizeof(baseline.t));
  std::vector<double> T(p.view_set.size() * KB_POSE_STRIDE);
  std::vector<int32_t> ok(p.view_set.size());
  h0.check(kb_estimate_transformations(h0.get(), p.resolution.data(), T.data(), ok.data()));
  std::vector<int> view_L(n_sets, -1), view_H(n_sets, -1);
  for (size_t w = 0; w < p.view_set.size(); ++w) (p.view_cam[w] == 0 ? view_L : view_H)[p.view_set[w]] = (int)w;
  auto pose_of = [&](int w) {
    Transformation t;
    std::memcpy(t.q, &T[(size_t)w * KB_POSE_STRIDE], sizeof(t.q));
    std::memcpy(t.t, &T[(size_t)w * KB_POSE_STRIDE + 4], sizeof(t.t));
    return t;
  };
  // The reference reuses its T_L / T_H variables across the loop and ignores estimateTransformation's return value (:249-255): a
  // failed PnP leaves the previous iteration's value in place.
  Transformation T_L, T_H;
  p.set_poses.clear();
  for (int s = 0; s < n_sets; ++s) {
    if (view_L[s] >= 0) {
      if (ok[view_L[s]]) T_L = pose_of(view_L[s]);
    } else {
      if (ok[view_H[s]]) T_H = pose_of(view_H[s]);
      T_L = mulInverse(T_H, baseline);
    }
    p.addPose(p.set_poses, T_L);
  }
  p.baselines.clear();
  p.addPose(p.baselines, baseline);
  Handle h(p, KB_ORDER_STEREO, target);
  if (!h.optimize(solution)) { /* the reference constructs a runtime_error here without throwing it (:296-298) */ }
  h.readCameras({&camera_L, &camera_H});
  h.check(kb_get_baselines(h.get(), b7));
  std::memcpy(baseline.q, b7, sizeof(baseline.q));
  std::memcpy(baseline.t, b7 + 4, sizeof(baseline.t));
  return baseline;
}

// ≙ kalibr2::tools::CalibrateMultiCameraRig: intrinsics of every camera, the baseline chain and one target pose per synced set
// (getTargetPoseGuess) optimised jointly.  Cameras are updated in place; returns the optimised baselines.  Throws when the linear
// solver fails, as the reference does (:417-419).
inline std::vector<Transformation> CalibrateMultiCameraRig(std::vector<Camera>& cameras, const std::vector<SyncedSet>& synced_observations,
                                                           const Target& target, const std::vector<Transformation>& baseline_guesses,
                                                           kb_solution* solution = nullptr) {
  using namespace detail;
  Problem p;
  for (const Camera& c : cameras) p.addCamera(c);
  for (const Transformation& b : baseline_guesses) p.addPose(p.baselines, b);
  p.n_sets = (int32_t)synced_observations.size();
  for (size_t s = 0; s < synced_observations.size(); ++s)
    for (size_t k = 0; k < synced_observations[s].size(); ++k)
      if (synced_observations[s][k]) p.addView((int)s, (int)k, *synced_observations[s][k]);
  identityPoses(p.set_poses, p.n_sets);
  Handle h(p, KB_ORDER_RIG, target);
  h.check(kb_initialize_set_poses(h.get(), p.resolution.data(), nullptr));  // getTargetPoseGuess for every synced set
  if (!h.optimize(solution)) throw std::runtime_error("Linear solver failed during optimization.");
  std::vector<Camera*> cams;
  for (Camera& c : cameras) cams.push_back(&c);
  h.readCameras(cams);
  std::vector<double> b(baseline_guesses.size() * KB_POSE_STRIDE);
  if (!b.empty()) h.check(kb_get_baselines(h.get(), b.data()));
  std::vector<Transformation> out(baseline_guesses.size());
  for (size_t j = 0; j < out.size(); ++j) {
    std::memcpy(out[j].q, &b[j * KB_POSE_STRIDE], sizeof(out[j].q));
    std::memcpy(out[j].t, &b[j * KB_POSE_STRIDE + 4], sizeof(out[j].t));
  }
  return out;
}

}  // namespace tools
}  // namespace kalibr_b200
