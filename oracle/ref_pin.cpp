// TEST INFRASTRUCTURE: the REFERENCE's own camera models - aslam_cv/aslam_cameras' PinholeProjection / OmniProjection /
// ExtendedUnifiedProjection / DoubleSphereProjection with RadialTangential / Equidistant / Fov / No distortion - compiled from the
// sources where they lie under /root/reference (headers + src/*Distortion.cpp) against the stand-in headers of oracle/ref_shim/
// (Eigen, Boost, OpenCV and sm_* are not in this image), behind a C entry point with the signature of the oracle's ko_camera_project.
// Built by `make -C oracle _ref/libkalibr_ref_cameras.so` into the git-ignored oracle/_ref/; used by tests/test_reference_pin_cpu.py and
// tests/golden/make_reference_camera_golden.py to PIN the oracle's restatement of rows a11-a17 of SURVEY.md §8 against reference code:
// homogeneousToKeypoint (value, validity flag, point Jacobian), homogeneousToKeypointIntrinsicsJacobian and
// homogeneousToKeypointDistortionJacobian - the calls ReprojectionError and CameraDesignVariable make
// (CVE/.../ReprojectionError.hpp:50-77, CVB/.../CameraDesignVariable.hpp).  No reference source is copied into this repository.
// (OmniProjection first: the EUCM and double-sphere headers use it without including it, as the reference's own umbrella header does)
#include <aslam/cameras/EquidistantDistortion.hpp>
#include <aslam/cameras/FovDistortion.hpp>
#include <aslam/cameras/NoDistortion.hpp>
#include <aslam/cameras/RadialTangentialDistortion.hpp>
#include <aslam/cameras/OmniProjection.hpp>
#include <aslam/cameras/PinholeProjection.hpp>
#include <aslam/cameras/DoubleSphereProjection.hpp>
#include <aslam/cameras/ExtendedUnifiedProjection.hpp>

#include <cstdint>

using namespace aslam::cameras;

namespace {
template <typename CAMERA>
int run(const CAMERA& cam, const double* ph_, double* y, double* Jp, double* Ji, double* Jd) {
  Eigen::Vector4d ph(ph_[0], ph_[1], ph_[2], ph_[3]);
  Eigen::VectorXd yk(2);
  yk.setZero();
  Eigen::MatrixXd J(2, 4), I, D;
  J.setZero();
  Eigen::VectorXd y2(2);
  y2.setZero();
  cam.homogeneousToKeypoint(ph, y2);                      // the evaluation of the error term
  const bool ok = cam.homogeneousToKeypoint(ph, yk, J);   // the evaluation of its Jacobians
  y[0] = y2(0);
  y[1] = y2(1);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 4; ++j) Jp[i * 4 + j] = J(i, j);
  cam.homogeneousToKeypointIntrinsicsJacobian(ph, I);
  cam.homogeneousToKeypointDistortionJacobian(ph, D);
  for (int i = 0; i < 12; ++i) Ji[i] = 0.0;
  for (int i = 0; i < 8; ++i) Jd[i] = 0.0;
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < I.cols() && j < 6; ++j) Ji[i * 6 + j] = I(i, j);
    for (int j = 0; j < D.cols() && j < 4; ++j) Jd[i * 4 + j] = D(i, j);
  }
  return ok ? 1 : 0;
}
}  // namespace

// model ids and parameter order as in include/kalibr_b200.h (kb_camera_model): projection parameters, then distortion parameters
extern "C" __attribute__((visibility("default"))) int32_t ref_camera_project(int32_t model, const double* p, const double* ph, double* y, double* Jp,
                                                                             double* Ji, double* Jd) {
  const int ru = 1 << 20, rv = 1 << 20;  // image size: only isValid() looks at it, the calls below do not
  switch (model) {
    case 0: return run(PinholeProjection<RadialTangentialDistortion>(p[0], p[1], p[2], p[3], ru, rv, RadialTangentialDistortion(p[4], p[5], p[6], p[7])), ph, y, Jp, Ji, Jd);
    case 1: return run(PinholeProjection<EquidistantDistortion>(p[0], p[1], p[2], p[3], ru, rv, EquidistantDistortion(p[4], p[5], p[6], p[7])), ph, y, Jp, Ji, Jd);
    case 2: return run(OmniProjection<RadialTangentialDistortion>(p[0], p[1], p[2], p[3], p[4], ru, rv, RadialTangentialDistortion(p[5], p[6], p[7], p[8])), ph, y, Jp, Ji, Jd);
    case 3: return run(ExtendedUnifiedProjection<NoDistortion>(p[0], p[1], p[2], p[3], p[4], p[5], ru, rv), ph, y, Jp, Ji, Jd);
    case 4: return run(DoubleSphereProjection<NoDistortion>(p[0], p[1], p[2], p[3], p[4], p[5], ru, rv), ph, y, Jp, Ji, Jd);
    case 5: return run(PinholeProjection<FovDistortion>(p[0], p[1], p[2], p[3], ru, rv, FovDistortion(p[4])), ph, y, Jp, Ji, Jd);
    case 6: return run(OmniProjection<NoDistortion>(p[0], p[1], p[2], p[3], p[4], ru, rv), ph, y, Jp, Ji, Jd);
  }
  return -1;
}
