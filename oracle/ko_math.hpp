// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
// (kalibr_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.
//
// PARITY: PINNED for the per-term part - camera models, SE(3) helpers, the expression tree of a reprojection term (residual and complete
// Jacobian rows) and the M-estimator weights - against the REFERENCE'S OWN CODE: oracle/ref_pin.cpp compiles the reference's headers and
// sources where they lie under /root/reference against stand-in headers for Eigen / Boost / OpenCV / sm_* (oracle/ref_shim/; none of those
// libraries is in this image) into the git-ignored oracle/_ref/, its outputs are committed as tests/golden/reference_golden.npz and
// tests/test_reference_pin_cpu.py holds this restatement to them (bit-identical in the build container).  PINNED for the LOOP too - design-
// variable order, ErrorTermFs::buildHessian, JacobianContainer::evaluateHessian, SparseBlockMatrix accumulation, the BlockCholesky solver's
// damping quirk, the LM policy, Optimizer2::optimize with update / revert / sticky solver failure: oracle/ref_pin_optimizer.cpp compiles
// those reference sources the same way and runs eleven small calibrations; counts, cost per iteration and final design variables are in the
// same fixture (keys opt*) and the same test file holds ko_optimize to them.  PINNED for the SparseCholesky regime (Kalibr2's default) as
// well: the reference's SparseCholeskyLinearSystemSolver.cpp, CompressedColumnJacobianTransposeBuilder, CompressedColumnMatrix and Cholmod
// wrapper compile the same way; the compressed-column J^T (layout bit for bit), e, rhs, a damped step and the eleven calibrations over that
// solver are in tests/golden/reference_sparse_golden.npz (tests/test_reference_sparse_pin_cpu.py).  UNPINNED: the CHOLMOD factorisation
// itself (a dense Cholesky stands in for the library in the pin; here an exact block-arrow Cholesky) and the few lines of glue in
// ReprojectionError / CameraDesignVariable around the pinned pieces (they need SuiteSparse / OpenCV): held by the reference's own property tests re-expressed in tests/ (H == J^T J, Schur == dense,
// solver-vs-solver) and by an independent derivation (tests/independent_model.py).  Each function cites the reference file:line it follows.
//
// ko_math.hpp: a small heap-backed dense matrix (stands in for Eigen::MatrixXd so the CPU baseline keeps the
// reference's per-term heap-allocation structure) and the sm_kinematics helpers.
#pragma once
#include <cassert>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace ko {

// Column-major like Eigen's default.
struct Mat {
  int r = 0, c = 0;
  std::vector<double> d;
  Mat() {}
  Mat(int rows, int cols) : r(rows), c(cols), d((size_t)rows * cols, 0.0) {}
  double& operator()(int i, int j) { return d[(size_t)j * r + i]; }
  double operator()(int i, int j) const { return d[(size_t)j * r + i]; }
  void setZero() { std::fill(d.begin(), d.end(), 0.0); }
  static Mat Identity(int n) {
    Mat m(n, n);
    for (int i = 0; i < n; ++i) m(i, i) = 1.0;
    return m;
  }
};

inline Mat operator*(const Mat& a, const Mat& b) {
  assert(a.c == b.r);
  Mat o(a.r, b.c);
  for (int j = 0; j < b.c; ++j)
    for (int k = 0; k < a.c; ++k) {
      const double bkj = b(k, j);
      for (int i = 0; i < a.r; ++i) o(i, j) += a(i, k) * bkj;
    }
  return o;
}
inline Mat operator*(double s, const Mat& a) {
  Mat o = a;
  for (double& v : o.d) v *= s;
  return o;
}
inline Mat operator-(const Mat& a) {
  Mat o = a;
  for (double& v : o.d) v = -v;
  return o;
}
inline Mat& operator+=(Mat& a, const Mat& b) {
  assert(a.r == b.r && a.c == b.c);
  for (size_t i = 0; i < a.d.size(); ++i) a.d[i] += b.d[i];
  return a;
}
inline Mat transpose(const Mat& a) {
  Mat o(a.c, a.r);
  for (int i = 0; i < a.r; ++i)
    for (int j = 0; j < a.c; ++j) o(j, i) = a(i, j);
  return o;
}

// ---- sm_kinematics -------------------------------------------------------------------------------

// Schweizer-Messer/sm_kinematics/src/rotations.cpp:78-84
inline Mat crossMx(double x, double y, double z) {
  Mat C(3, 3);
  C(0, 1) = -z; C(0, 2) = y;
  C(1, 0) = z;  C(1, 2) = -x;
  C(2, 0) = -y; C(2, 1) = x;
  return C;
}

// Schweizer-Messer/sm_kinematics/src/transformations.cpp:45-53
inline Mat boxMinus(const double p[4]) {
  Mat B(4, 6);
  B(0, 0) = p[3]; B(0, 4) = -p[2]; B(0, 5) = p[1];
  B(1, 1) = p[3]; B(1, 3) = p[2];  B(1, 5) = -p[0];
  B(2, 2) = p[3]; B(2, 3) = -p[1]; B(2, 4) = p[0];
  return B;
}

// Schweizer-Messer/sm_kinematics/src/transformations.cpp:132-142
inline Mat boxTimes(const Mat& T) {
  Mat o(6, 6);
  Mat C(3, 3);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C(i, j) = T(i, j);
  Mat tC = -(crossMx(T(0, 3), T(1, 3), T(2, 3)) * C);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      o(i, j) = C(i, j);
      o(i + 3, j + 3) = C(i, j);
      o(i, j + 3) = tC(i, j);
    }
  return o;
}

// Schweizer-Messer/sm_kinematics/src/quaternion_algebra.cpp:77-101 (scalar-last quaternion)
inline Mat quat2r(const double q[4]) {
  Mat R(3, 3);
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R(0, 0) = x * x - y * y - z * z + w * w;
  R(0, 1) = x * y * 2.0 + z * w * 2.0;
  R(0, 2) = x * z * 2.0 - y * w * 2.0;
  R(1, 0) = x * y * 2.0 - z * w * 2.0;
  R(1, 1) = -x * x + y * y - z * z + w * w;
  R(1, 2) = x * w * 2.0 + y * z * 2.0;
  R(2, 0) = x * z * 2.0 + y * w * 2.0;
  R(2, 1) = x * w * (-2.0) + y * z * 2.0;
  R(2, 2) = -x * x - y * y + z * z + w * w;
  return R;
}

// Schweizer-Messer/sm_kinematics/src/quaternion_algebra.cpp:10-12, 200-219
inline void axisAngle2quat(const double a[3], double out[4]) {
  const double theta = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  double na;
  if (theta < std::pow(std::numeric_limits<double>::epsilon(), 0.25)) {
    na = 0.5 + (theta * theta) * (1.0 / 48.0);
  } else {
    na = std::sin(theta * 0.5) / theta;
  }
  out[0] = a[0] * na; out[1] = a[1] * na; out[2] = a[2] * na;
  out[3] = std::cos(theta * 0.5);
}

// Schweizer-Messer/sm_kinematics/src/quaternion_algebra.cpp:302-317
inline void updateQuat(const double q[4], const double dq[3], double out[4]) {
  double d[4];
  axisAngle2quat(dq, d);
  const double ca = d[3];
  out[0] = q[0] * ca + d[0] * q[3] - d[1] * q[2] + d[2] * q[1];
  out[1] = q[1] * ca + d[0] * q[2] + d[1] * q[3] - d[2] * q[0];
  out[2] = q[2] * ca - d[0] * q[1] + d[1] * q[0] + d[2] * q[3];
  out[3] = q[3] * ca - d[0] * q[0] - d[1] * q[1] - d[2] * q[2];
}

// General 4x4 inverse by cofactors: stands in for Eigen's Matrix4d::inverse(), which
// aslam_backend_expressions/src/TransformationExpressionNode.cpp:81,88 applies to a rigid transform.
inline Mat inverse4(const Mat& M) {
  const double* m = M.d.data();  // column-major: m[col*4+row]
  auto a = [&](int i, int j) { return m[j * 4 + i]; };
  double s0 = a(0, 0) * a(1, 1) - a(1, 0) * a(0, 1);
  double s1 = a(0, 0) * a(1, 2) - a(1, 0) * a(0, 2);
  double s2 = a(0, 0) * a(1, 3) - a(1, 0) * a(0, 3);
  double s3 = a(0, 1) * a(1, 2) - a(1, 1) * a(0, 2);
  double s4 = a(0, 1) * a(1, 3) - a(1, 1) * a(0, 3);
  double s5 = a(0, 2) * a(1, 3) - a(1, 2) * a(0, 3);
  double c5 = a(2, 2) * a(3, 3) - a(3, 2) * a(2, 3);
  double c4 = a(2, 1) * a(3, 3) - a(3, 1) * a(2, 3);
  double c3 = a(2, 1) * a(3, 2) - a(3, 1) * a(2, 2);
  double c2 = a(2, 0) * a(3, 3) - a(3, 0) * a(2, 3);
  double c1 = a(2, 0) * a(3, 2) - a(3, 0) * a(2, 2);
  double c0 = a(2, 0) * a(3, 1) - a(3, 0) * a(2, 1);
  double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  double id = 1.0 / det;
  Mat o(4, 4);
  o(0, 0) = (a(1, 1) * c5 - a(1, 2) * c4 + a(1, 3) * c3) * id;
  o(0, 1) = (-a(0, 1) * c5 + a(0, 2) * c4 - a(0, 3) * c3) * id;
  o(0, 2) = (a(3, 1) * s5 - a(3, 2) * s4 + a(3, 3) * s3) * id;
  o(0, 3) = (-a(2, 1) * s5 + a(2, 2) * s4 - a(2, 3) * s3) * id;
  o(1, 0) = (-a(1, 0) * c5 + a(1, 2) * c2 - a(1, 3) * c1) * id;
  o(1, 1) = (a(0, 0) * c5 - a(0, 2) * c2 + a(0, 3) * c1) * id;
  o(1, 2) = (-a(3, 0) * s5 + a(3, 2) * s2 - a(3, 3) * s1) * id;
  o(1, 3) = (a(2, 0) * s5 - a(2, 2) * s2 + a(2, 3) * s1) * id;
  o(2, 0) = (a(1, 0) * c4 - a(1, 1) * c2 + a(1, 3) * c0) * id;
  o(2, 1) = (-a(0, 0) * c4 + a(0, 1) * c2 - a(0, 3) * c0) * id;
  o(2, 2) = (a(3, 0) * s4 - a(3, 1) * s2 + a(3, 3) * s0) * id;
  o(2, 3) = (-a(2, 0) * s4 + a(2, 1) * s2 - a(2, 3) * s0) * id;
  o(3, 0) = (-a(1, 0) * c3 + a(1, 1) * c1 - a(1, 2) * c0) * id;
  o(3, 1) = (a(0, 0) * c3 - a(0, 1) * c1 + a(0, 2) * c0) * id;
  o(3, 2) = (-a(3, 0) * s3 + a(3, 1) * s1 - a(3, 2) * s0) * id;
  o(3, 3) = (a(2, 0) * s3 - a(2, 1) * s1 + a(2, 2) * s0) * id;
  return o;
}

}  // namespace ko
