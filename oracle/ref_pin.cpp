// TEST INFRASTRUCTURE: the REFERENCE's own camera models - aslam_cv/aslam_cameras' PinholeProjection / OmniProjection /
// ExtendedUnifiedProjection / DoubleSphereProjection with RadialTangential / Equidistant / Fov / No distortion - compiled from the
// sources where they lie under /root/reference (headers + src/*Distortion.cpp) against the stand-in headers of oracle/ref_shim/
// (Eigen, Boost, OpenCV and sm_* are not in this image), behind a C entry point with the signature of the oracle's ko_camera_project.
// Built by `make -C oracle _ref/libkalibr_ref.so` into the git-ignored oracle/_ref/; used by tests/test_reference_pin_cpu.py and
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

#include <sm/kinematics/quaternion_algebra.hpp>
#include <sm/kinematics/transformations.hpp>

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

// ---- sm_kinematics: the SE(3) helpers the expression nodes and the quaternion design variable call, from the
// reference's own quaternion_algebra.cpp / rotations.cpp / transformations.cpp --------------------------------------------------------
extern "C" __attribute__((visibility("default"))) void ref_quat2r(const double* q, double* R9) {
  const Eigen::Matrix3d R = sm::kinematics::quat2r(Eigen::Vector4d(q[0], q[1], q[2], q[3]));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R(i, j);
}
extern "C" __attribute__((visibility("default"))) void ref_update_quat(const double* q, const double* dq, double* out) {
  const Eigen::Vector4d r = sm::kinematics::updateQuat(Eigen::Vector4d(q[0], q[1], q[2], q[3]), Eigen::Vector3d(dq[0], dq[1], dq[2]));
  for (int i = 0; i < 4; ++i) out[i] = r(i);
}
extern "C" __attribute__((visibility("default"))) void ref_box_minus(const double* p4, double* out24_rowmajor) {
  const Eigen::Matrix<double, 4, 6> B = sm::kinematics::boxMinus(Eigen::Vector4d(p4[0], p4[1], p4[2], p4[3]));
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 6; ++j) out24_rowmajor[i * 6 + j] = B(i, j);
}
extern "C" __attribute__((visibility("default"))) void ref_box_times(const double* T16_rowmajor, double* out36_rowmajor) {
  Eigen::Matrix4d T;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) T(i, j) = T16_rowmajor[i * 4 + j];
  const Eigen::Matrix<double, 6, 6> B = sm::kinematics::boxTimes(T);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) out36_rowmajor[i * 6 + j] = B(i, j);
}

// ---- the expression tree of a reprojection term: T_cam_w = B_{k-1} ... B_0 inverse(T_target_cam0), p_c = T_cam_w p_t,
// built from the reference's own design variables and expression nodes exactly as K2/CalibrationTools.hpp:32-45, 405-408 and
// K2/CameraCalibrator.hpp:213-222 build it, evaluated with HomogeneousExpression::toHomogeneous / evaluateJacobians(container, chain rule)
// - the calls CVE/.../ReprojectionError.hpp:50-77 makes ------------------------------------------------------------------------------
#include <aslam/backend/EuclideanPoint.hpp>
#include <aslam/backend/HomogeneousExpression.hpp>
#include <aslam/backend/HomogeneousPoint.hpp>
#include <aslam/backend/JacobianContainer.hpp>
#include <aslam/backend/RotationQuaternion.hpp>
#include <aslam/backend/TransformationBasic.hpp>
#include <aslam/backend/TransformationExpression.hpp>

// poses7: [1 + n_base][7] = (q xyzw, t) of the set pose T_target_cam0, then of the baselines; chain (optional): rows x 4 row-major matrix the
// Jacobians are multiplied with (the 2x4 point Jacobian of the camera); out_J: [(1 + n_base) * 2][rows][3] row-major in the order
// (set q, set t, baseline 0 q, baseline 0 t, ...), rows = chain_rows > 0 ? chain_rows : 4.  Returns the number of Jacobian blocks written.
extern "C" __attribute__((visibility("default"))) int32_t ref_point_chain(int32_t n_base, const double* poses7, const double* p4, int32_t chain_rows,
                                                                          const double* chain, double* out_pc4, double* out_J) {
  using namespace aslam::backend;
  const int n_pose = 1 + n_base;
  std::vector<boost::shared_ptr<RotationQuaternion>> qs;
  std::vector<boost::shared_ptr<EuclideanPoint>> ts;
  std::vector<boost::shared_ptr<TransformationBasic>> Ts;
  for (int i = 0; i < n_pose; ++i) {
    const double* p = poses7 + 7 * i;
    qs.push_back(boost::make_shared<RotationQuaternion>(Eigen::Vector4d(p[0], p[1], p[2], p[3])));
    qs.back()->setActive(true);
    qs.back()->setBlockIndex(2 * i);
    ts.push_back(boost::make_shared<EuclideanPoint>(Eigen::Vector3d(p[4], p[5], p[6])));
    ts.back()->setActive(true);
    ts.back()->setBlockIndex(2 * i + 1);
    Ts.push_back(boost::make_shared<TransformationBasic>(qs.back()->toExpression(), ts.back()->toExpression()));
  }
  TransformationExpression T_cam_w = Ts[0]->toExpression().inverse();
  for (int j = 0; j < n_base; ++j) T_cam_w = Ts[1 + j]->toExpression() * T_cam_w;
  HomogeneousPoint target_point(Eigen::Vector4d(p4[0], p4[1], p4[2], p4[3]));  // not active: the target is fixed
  HomogeneousExpression point = T_cam_w * target_point.toExpression();
  const Eigen::Vector4d pc = point.toHomogeneous();
  for (int i = 0; i < 4; ++i) out_pc4[i] = pc(i);
  const int rows = chain_rows > 0 ? chain_rows : 4;
  JacobianContainer jc(rows);
  if (chain_rows > 0) {
    Eigen::MatrixXd C(rows, 4);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < 4; ++c) C(r, c) = chain[r * 4 + c];
    point.evaluateJacobians(jc, C);
  } else {
    point.evaluateJacobians(jc);
  }
  for (int i = 0; i < 2 * n_pose * rows * 3; ++i) out_J[i] = 0.0;
  int n = 0;
  for (JacobianContainer::map_t::iterator it = jc.begin(); it != jc.end(); ++it, ++n) {
    const int b = it->first->blockIndex();
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < 3; ++c) out_J[(b * rows + r) * 3 + c] = it->second(r, c);
  }
  return n;
}

// ---- M-estimator policies: the reference's weight functions, BE/src/MEstimatorPolicies.cpp ---------------------
// kind as kb_m_estimator: 0 none, 1 Huber(k = p0), 2 Cauchy(sigma2 = p0), 3 Geman-McClure(sigma2 = p0), 4 Blake-Zisserman(df = p0, pCut = p1,
// wCut = p2).  (Blake-Zisserman's epsilon needs a chi-squared quantile: Boost.Math in the reference, a stand-in here - ref_shim/boost/math.)
#include <aslam/backend/MEstimatorPolicies.hpp>
extern "C" __attribute__((visibility("default"))) double ref_m_estimator_weight(int32_t kind, double p0, double p1, double p2, double squared_error) {
  using namespace aslam::backend;
  switch (kind) {
    case 0: return NoMEstimator().getWeight(squared_error);
    case 1: return HuberMEstimator(p0).getWeight(squared_error);
    case 2: return CauchyMEstimator(p0).getWeight(squared_error);
    case 3: return GemanMcClureMEstimator(p0).getWeight(squared_error);
    case 4: return BlakeZissermanMEstimator((size_t)p0, p1, p2).getWeight(squared_error);
  }
  return -1.0;
}
