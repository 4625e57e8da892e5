// Test driver for include/kalibr_b200/calibration_tools.hpp: reads a flattened set of observations written by
// tests/test_drivers_gpu.py, runs one of the three kalibr2 drivers on the device and prints the results as "key value..." lines.
//   driver_main <single|stereo|rig> <problem.bin>
//   driver_main estimator <problem.bin> <infoGainDelta> [check]   every synced set offered to the incremental estimator, in order
//                                                                 (check: also verifies the returned null / row spaces and covariance)
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <fstream>

#include "kalibr_b200/incremental_estimator.hpp"

using namespace kalibr_b200::tools;

template <typename T>
static std::vector<T> readArray(std::ifstream& f) {
  int64_t n = 0;
  f.read(reinterpret_cast<char*>(&n), sizeof(n));
  std::vector<T> v((size_t)n);
  f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * (size_t)n);
  return v;
}

static void printCamera(const char* key, const Camera& c) {
  std::printf("%s", key);
  for (double p : c.params) std::printf(" %.17g", p);
  std::printf("\n");
}
static void printPose(const char* key, const Transformation& T) {
  std::printf("%s %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", key, T.q[0], T.q[1], T.q[2], T.q[3], T.t[0], T.t[1], T.t[2]);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string mode = argv[1];
  std::ifstream f(argv[2], std::ios::binary);
  if (!f) return 3;
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
  const int n_cams = dims[0], n_sets = dims[1];
  Target target;
  target.rows = dims[2];
  target.cols = dims[3];
  target.points = points;
  std::vector<Camera> cams((size_t)n_cams);
  for (int k = 0; k < n_cams; ++k) {
    cams[k].model = models[k];
    cams[k].ru = res[2 * k];
    cams[k].rv = res[2 * k + 1];
    for (int i = 0; i < KB_CAM_PARAM_STRIDE; ++i) cams[k].params[i] = params[(size_t)k * KB_CAM_PARAM_STRIDE + i];
  }
  std::vector<SyncedSet> synced((size_t)n_sets, SyncedSet((size_t)n_cams));
  for (size_t w = 0; w < view_set.size(); ++w) {
    Observation o;
    for (int64_t i = view_begin[w]; i < view_begin[w + 1]; ++i) {
      o.corner_id.push_back(corner[(size_t)i]);
      o.u.push_back(yu[(size_t)i]);
      o.v.push_back(yv[(size_t)i]);
    }
    synced[(size_t)view_set[w]][(size_t)view_cam[w]] = std::move(o);
  }
  kb_solution sol;
  try {
    if (mode == "single") {
      std::vector<Observation> obs;
      for (auto& s : synced) if (s[0]) obs.push_back(*s[0]);
      ReprojectionStatistics st;
      const bool ok = CalibrateSingleCamera(obs, cams[0], target, std::nullopt, &sol, &st);
      std::printf("ok %d\n", ok ? 1 : 0);
      std::printf("stats %.17g %.17g %.17g %.17g %.17g %.17g\n", st.n, st.mean_u, st.mean_v, st.std_u, st.std_v, st.rmse);
    } else if (mode == "stereo") {
      std::vector<std::optional<Observation>> L, H;
      for (auto& s : synced) { L.push_back(s[0]); H.push_back(s[1]); }
      printPose("baseline0", CalibrateStereoPair(cams[0], cams[1], L, H, target, &sol));
    } else if (mode == "rig") {
      std::vector<Transformation> guesses((size_t)n_cams - 1);
      for (int j = 0; j + 1 < n_cams; ++j) {
        std::memcpy(guesses[j].q, &baselines[(size_t)j * 7], sizeof(guesses[j].q));
        std::memcpy(guesses[j].t, &baselines[(size_t)j * 7 + 4], sizeof(guesses[j].t));
      }
      auto out = CalibrateMultiCameraRig(cams, synced, target, guesses, &sol);
      for (size_t j = 0; j < out.size(); ++j) printPose(("baseline" + std::to_string(j)).c_str(), out[j]);
    } else if (mode == "estimator") {
      std::vector<Transformation> guesses((size_t)n_cams - 1);
      for (int j = 0; j + 1 < n_cams; ++j) {
        std::memcpy(guesses[j].q, &baselines[(size_t)j * 7], sizeof(guesses[j].q));
        std::memcpy(guesses[j].t, &baselines[(size_t)j * 7 + 4], sizeof(guesses[j].t));
      }
      kalibr_b200::calibration::IncrementalEstimator::Options eo;
      eo.infoGainDelta = argc > 3 ? std::atof(argv[3]) : 0.2;
      eo.checkValidity = true;  // CalibrateCameras.cpp:261
      kalibr_b200::calibration::IncrementalEstimator estimator(cams, guesses, target, eo);
      const auto t0 = std::chrono::steady_clock::now();
      std::vector<double> batch_ms;
      for (int s = 0; s < n_sets; ++s) {
        Transformation T;
        std::memcpy(T.q, &set_poses[(size_t)s * 7], sizeof(T.q));
        std::memcpy(T.t, &set_poses[(size_t)s * 7 + 4], sizeof(T.t));
        const auto tb = std::chrono::steady_clock::now();
        auto r = estimator.addBatch(synced[(size_t)s], T);
        batch_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb).count());
        std::printf("batch%d %d %.17g %td %zu %.17g %.17g\n", s, r.batchAccepted ? 1 : 0, r.informationGain, r.rankTheta, r.numIterations, r.JStart, r.JFinal);
        if (argc > 4) {  // "check": ReturnValue's bases: [obsBasis | nobsBasis] must be an orthonormal basis of the calibration block; sigma2Theta = V_r S_r^-1 V_r^T
          const auto& O = r.obsBasis; const auto& N = r.nobsBasis;
          const std::ptrdiff_t n = O.rows;
          double orth = 0.0, cov = 0.0;
          auto col = [&](std::ptrdiff_t c, std::ptrdiff_t i) { return c < O.cols ? O(i, c) : N(i, c - O.cols); };
          for (std::ptrdiff_t a1 = 0; a1 < n; ++a1)
            for (std::ptrdiff_t b1 = a1; b1 < n; ++b1) {
              double d = 0.0;
              for (std::ptrdiff_t i = 0; i < n; ++i) d += col(a1, i) * col(b1, i);
              orth = std::max(orth, std::fabs(d - (a1 == b1 ? 1.0 : 0.0)));
            }
          {  // trace(sigma2Theta) = sum of the reciprocal retained singular values (V_r has orthonormal columns)
            double tr = 0.0, sum = 0.0;
            for (std::ptrdiff_t i = 0; i < n; ++i) tr += r.sigma2Theta(i, i);
            for (std::ptrdiff_t k = 0; k < O.cols; ++k) sum += 1.0 / r.singularValues[(size_t)k];
            cov = std::fabs(tr - sum) / sum;
          }
          std::printf("spaces%d %td %td %td %.3e %.3e %zu %td\n", s, n, O.cols, N.cols, orth, cov, r.singularValuesScaled.size(), r.obsBasisScaled.cols);
        }
      }
      std::printf("loop_ms %.3f\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      std::printf("batch_ms");  // per addBatch; the first one pays the CUDA context / module load of the process
      for (double m : batch_ms) std::printf(" %.3f", m);
      std::printf("\n");
      std::printf("accepted %zu\n", estimator.getNumBatches());
      for (size_t j = 0; j < estimator.baselines().size(); ++j) printPose(("baseline" + std::to_string(j)).c_str(), estimator.baselines()[j]);
      for (int k = 0; k < n_cams; ++k) printCamera(("camera" + std::to_string(k)).c_str(), estimator.cameras()[(size_t)k]);
      return 0;
    } else {
      return 2;
    }
  } catch (const std::exception& e) {
    std::printf("error %s\n", e.what());
    return 1;
  }
  std::printf("solution %d %d %d %.17g %.17g\n", sol.iterations, sol.failed_iterations, sol.linear_solver_failure, sol.j_start, sol.j_final);
  for (int k = 0; k < n_cams; ++k) printCamera(("camera" + std::to_string(k)).c_str(), cams[k]);
  return 0;
}
