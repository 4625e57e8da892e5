// Initial-guess stage on the device: what Kalibr2's drivers run BEFORE the batch solve.
//
//   pnp_kernel<MODEL>      CameraGeometry::estimateTransformation for a batch of views, one warp per view:
//                          corners rounded to float (cv::Point2f / Point3f) -> keypointToEuclidean -> 80 degree cone filter ->
//                          normalised coordinates rounded to float -> planar PnP (cv::solvePnP, SOLVEPNP_ITERATIVE: homography
//                          initialisation, then Levenberg-Marquardt on the reprojection error) -> T_target_camera
//                            ≙ CAM/.../implementation/PinholeProjection.hpp:831-891, OmniProjection.hpp:882-955,
//                              ExtendedUnifiedProjection.hpp:791-860, DoubleSphereProjection.hpp:842-910
//   set_pose_guess_kernel  getTargetPoseGuess for every synced set: the view of the camera that saw most corners, chained through the
//                          baseline guesses                           ≙ K2/include/kalibr2/CalibrationTools.hpp:316-356
//
// The PnP minimises the same cost as OpenCV from the same kind of start (plane-to-image homography), on the manifold
// (R <- exp(d) R) instead of the Rodrigues vector; both end in the same minimum (tests pin it against cv2.solvePnP outputs).
// Everything a warp needs lives in shared memory (6 doubles per corner) and registers; sums are warp butterflies in a fixed order.
#include "kb_device.cuh"
#include "kb_models.cuh"

namespace kb {
namespace {

constexpr int PNP_WARPS = 4;
constexpr double COS80 = 0.17364817766693041;       // std::cos(80.0 * M_PI / 180.0)
constexpr double FOV_MAX_VALID_ANGLE = 1.5533430342749532;  // 89 degrees: CAM/include/aslam/cameras/FovDistortion.hpp:146

// A view is worked on by a group of PNP_G adjacent lanes (4 views per warp): the serial parts of the algorithm (small
// factorisations, the LM control flow) are latency-bound and identical in every lane of a group, so narrower groups keep more
// views in flight per warp.  Sums are butterflies inside the group, in a fixed order.
#ifndef KB_PNP_G
#define KB_PNP_G 8
#endif
constexpr int PNP_G = KB_PNP_G;
#ifndef KB_PNP_MINB
#define KB_PNP_MINB 4
#endif
#ifndef KB_PNP_STEP_TOL
#define KB_PNP_STEP_TOL 1e-10
#endif
constexpr int PNP_GROUPS = PNP_WARPS * 32 / PNP_G;  // views in flight per CTA
__device__ __forceinline__ double wsum(double v, unsigned mask) {
#pragma unroll
  for (int o = PNP_G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// ---- undistort: CAM/.../implementation/RadialTangentialDistortion.hpp:68-100, EquidistantDistortion.hpp:186-211, FovDistortion.hpp:89-116
__device__ __forceinline__ void undistort_radtan(const double* __restrict__ k, double& x, double& y) {
  double bx = x, by = y;
  for (int i = 0; i < 5; ++i) {
    double tx = bx, ty = by, F[2][2], Jk[2][4];
    radtan<true>(k, tx, ty, F, Jk);
    const double ex = x - tx, ey = y - ty;
    // du = (F^T F)^-1 F^T e
    const double a00 = F[0][0] * F[0][0] + F[1][0] * F[1][0], a01 = F[0][0] * F[0][1] + F[1][0] * F[1][1], a11 = F[0][1] * F[0][1] + F[1][1] * F[1][1];
    const double g0 = F[0][0] * ex + F[1][0] * ey, g1 = F[0][1] * ex + F[1][1] * ey;
    const double id = 1.0 / (a00 * a11 - a01 * a01);
    bx += (a11 * g0 - a01 * g1) * id;
    by += (a00 * g1 - a01 * g0) * id;
    if (ex * ex + ey * ey < 1e-15) break;
  }
  x = bx;
  y = by;
}
__device__ __forceinline__ void undistort_equi(const double* __restrict__ k, double& x, double& y) {
  const double thetad = sqrt(x * x + y * y);
  double theta = thetad;
  for (int i = 0; i < 20; ++i) {  // the reference always runs 20 rounds; once the value repeats, the remaining ones reproduce it
    const double t2 = theta * theta, t4 = t2 * t2;
    const double tn = thetad / (1.0 + k[0] * t2 + k[1] * t4 + k[2] * t4 * t2 + k[3] * t4 * t4);
    if (tn == theta) break;
    theta = tn;
  }
  const double scaling = tan(theta) / thetad;  // 0 / 0 at the image centre, as in the reference: the corner drops out below
  x *= scaling;
  y *= scaling;
}
__device__ __forceinline__ void undistort_fov(const double* __restrict__ k, double& x, double& y) {
  const double w = k[0];
  const double mul2tanwby2 = tan(w / 2.0) * 2.0;
  const double r_d = sqrt(x * x + y * y);
  if (mul2tanwby2 == 0.0 || r_d == 0.0) return;
  if (fabs(r_d * w) <= FOV_MAX_VALID_ANGLE) {
    const double r_u = tan(r_d * w) / (r_d * mul2tanwby2);
    x *= r_u;
    y *= r_u;
  }
}

// keypointToEuclidean: PinholeProjection.hpp:200-229, OmniProjection.hpp:232-263, ExtendedUnifiedProjection.hpp:248-283,
// DoubleSphereProjection.hpp:271-307.  res = (ru, rv) or null: the pinhole's isValid(keypoint) image-bounds test.
template <int MODEL>
__device__ __forceinline__ bool keypoint_to_euclidean(const double* __restrict__ prm, const int* __restrict__ res, double u, double v, double out[3]) {
  constexpr int P = model_P(MODEL);
  const double* k = prm + P;
  if (MODEL == PINHOLE_RADTAN || MODEL == PINHOLE_EQUI || MODEL == PINHOLE_FOV) {
    double x = (u - prm[2]) / prm[0], y = (v - prm[3]) / prm[1];
    if (MODEL == PINHOLE_RADTAN) undistort_radtan(k, x, y);
    else if (MODEL == PINHOLE_EQUI) undistort_equi(k, x, y);
    else undistort_fov(k, x, y);
    out[0] = x; out[1] = y; out[2] = 1.0;
    return res == nullptr || (u >= 0.0 && u < (double)res[0] && v >= 0.0 && v < (double)res[1]);
  } else if (MODEL == OMNI_RADTAN || MODEL == OMNI_NONE) {
    const double xi = prm[0];
    double x = (1.0 / prm[1]) * (u - prm[3]), y = (1.0 / prm[2]) * (v - prm[4]);
    if (MODEL == OMNI_RADTAN) undistort_radtan(k, x, y);
    const double rho2 = x * x + y * y;
    if (!(xi <= 1.0 || rho2 <= 1.0 / (xi * xi - 1.0))) return false;
    out[0] = x; out[1] = y;
    out[2] = 1.0 - xi * (rho2 + 1.0) / (xi + sqrt(1.0 + (1.0 - xi * xi) * rho2));
    return true;
  } else if (MODEL == EUCM_NONE) {
    const double alpha = prm[0], beta = prm[1];
    const double mx = (1.0 / prm[2]) * (u - prm[4]), my = (1.0 / prm[3]) * (v - prm[5]);
    const double r2 = mx * mx + my * my;
    if (!(alpha <= 0.5 || r2 <= 1.0 / (beta * (2.0 * alpha - 1.0)))) return false;
    const double gamma = 1.0 - alpha;
    const double kk = (1.0 - alpha * alpha * beta * r2) / (alpha * sqrt(1.0 - (alpha - gamma) * beta * r2) + gamma);
    const double ni = 1.0 / sqrt(r2 + kk * kk);
    out[0] = mx * ni; out[1] = my * ni; out[2] = kk * ni;
    return true;
  } else {  // DS_NONE
    const double xi = prm[0], alpha = prm[1];
    const double mx = (1.0 / prm[2]) * (u - prm[4]), my = (1.0 / prm[3]) * (v - prm[5]);
    const double r2 = mx * mx + my * my;
    if (!(alpha <= 0.5 || r2 <= 1.0 / (2.0 * alpha - 1.0))) return false;
    const double mz = (1.0 - alpha * alpha * r2) / (alpha * sqrt(1.0 - (2.0 * alpha - 1.0) * r2) + 1.0 - alpha);
    const double mz2 = mz * mz;
    const double kk = (mz * xi + sqrt(mz2 + (1.0 - xi * xi) * r2)) / (mz2 + r2);
    out[0] = kk * mx; out[1] = kk * my; out[2] = kk * mz - xi;
    return true;
  }
}

// ---- small dense helpers (uniform across the warp: every lane computes the same values) ----------------------------------
__device__ __forceinline__ void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_inv_t(const double A[9], double B[9]) {  // B = A^-T
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double id = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
  B[0] = c00 * id; B[1] = c01 * id; B[2] = c02 * id;
  B[3] = (A[2] * A[7] - A[1] * A[8]) * id; B[4] = (A[0] * A[8] - A[2] * A[6]) * id; B[5] = (A[1] * A[6] - A[0] * A[7]) * id;
  B[6] = (A[1] * A[5] - A[2] * A[4]) * id; B[7] = (A[2] * A[3] - A[0] * A[5]) * id; B[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}
// nearest rotation (polar factor = U V^T of the SVD, what cv::Rodrigues(matrix) projects onto): Newton iteration X <- (X + X^-T) / 2
__device__ __forceinline__ void nearest_rotation(double R[9]) {
  for (int it = 0; it < 12; ++it) {
    double Y[9];
    mat3_inv_t(R, Y);
    double d = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double n = 0.5 * (R[i] + Y[i]);
      d = fmax(d, fabs(n - R[i]));
      R[i] = n;
    }
    if (d < 1e-15) break;
  }
}
// R <- exp([w]x) R (Rodrigues formula)
__device__ __forceinline__ void rotate_left(const double w[3], const double R[9], double out[9]) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double a, b;  // exp = I + a K + b K^2
  if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
  else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
  const double K[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
  double K2[9], E[9];
  mat3_mul(K, K, K2);
#pragma unroll
  for (int i = 0; i < 9; ++i) E[i] = a * K[i] + b * K2[i] + ((i % 4 == 0) ? 1.0 : 0.0);
  mat3_mul(E, R, out);
}
// symmetric 3x3 eigen-decomposition by cyclic Jacobi; eigenvalues in w, eigenvectors in the columns of V
__device__ __forceinline__ void eig3(double A[9], double w[3], double V[9]) {
#pragma unroll
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
    if (off < 1e-300) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const double apq = A[p * 3 + q];
      if (fabs(apq) < 1e-300) continue;
      const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // A <- A J
        const double akp = A[k * 3 + p], akq = A[k * 3 + q];
        A[k * 3 + p] = c * akp - s * akq;
        A[k * 3 + q] = s * akp + c * akq;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // A <- J^T A
        const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
        A[p * 3 + k] = c * apk - s * aqk;
        A[q * 3 + k] = s * apk + c * aqk;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
        V[k * 3 + p] = c * vkp - s * vkq;
        V[k * 3 + q] = s * vkp + c * vkq;
      }
    }
  }
  w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
}
// Cholesky solve of an N x N SPD system held as a full row-major array (in place); returns false when not positive definite
template <int N>
__device__ __forceinline__ bool chol_solve(double A[N * N], double x[N]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double d = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[j * N + k] * A[j * N + k];
    if (!(d > 0.0)) ok = false;
    d = sqrt(d);
    A[j * N + j] = d;
    const double id = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double s = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[i * N + k] * A[j * N + k];
      A[i * N + j] = s * id;
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= A[i * N + k] * x[k];
    x[i] = s / A[i * N + i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double s = x[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s -= A[k * N + i] * x[k];
    x[i] = s / A[i * N + i];
  }
  return ok;
}
// sm::kinematics::r2quat (Schweizer-Messer/sm_kinematics/src/quaternion_algebra.cpp:16-75), R row-major
__device__ __forceinline__ void r2quat(const double R[9], double q[4]) {
  const double c1 = R[0], c2 = R[3], c3 = R[6], c4 = R[1], c5 = R[4], c6 = R[7], c7 = R[2], c8 = R[5], c9 = R[8];
  const double dc[4] = {fabs(1.0 + c1 - c5 - c9), fabs(1.0 - c1 + c5 - c9), fabs(1.0 - c1 - c5 + c9), fabs(1.0 + c1 + c5 + c9)};
  int m = 0;
  double mv = dc[0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (dc[i] > mv) { m = i; mv = dc[i]; }
  double c;
  if (m == 0) { q[0] = 0.5 * sqrt(dc[0]); c = 0.25 / q[0]; q[1] = c * (c4 + c2); q[2] = c * (c7 + c3); q[3] = c * (c8 - c6); }
  else if (m == 1) { q[1] = 0.5 * sqrt(dc[1]); c = 0.25 / q[1]; q[0] = c * (c4 + c2); q[2] = c * (c6 + c8); q[3] = c * (c3 - c7); }
  else if (m == 2) { q[2] = 0.5 * sqrt(dc[2]); c = 0.25 / q[2]; q[0] = c * (c3 + c7); q[1] = c * (c6 + c8); q[3] = c * (c4 - c2); }
  else { q[3] = 0.5 * sqrt(dc[3]); c = 0.25 / q[3]; q[0] = c * (c8 - c6); q[1] = c * (c3 - c7); q[2] = c * (c4 - c2); }
  if (q[3] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
}
// sm quat2r (quaternion_algebra.cpp:77-101), row-major
__device__ __forceinline__ void quat2r_rm(const double* __restrict__ q, double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = x * x - y * y - z * z + w * w; R[1] = 2.0 * x * y + 2.0 * z * w;        R[2] = 2.0 * x * z - 2.0 * y * w;
  R[3] = 2.0 * x * y - 2.0 * z * w;     R[4] = -x * x + y * y - z * z + w * w;   R[5] = 2.0 * x * w + 2.0 * y * z;
  R[6] = 2.0 * x * z + 2.0 * y * w;     R[7] = -2.0 * x * w + 2.0 * y * z;       R[8] = -x * x - y * y + z * z + w * w;
}

// reprojection cost of the pose (R, t) over the warp's points (weights 0 / 1), summed over the warp
__device__ __forceinline__ double pnp_cost(const double R[9], const double t[3], const float* sX, const float* sY, const float* sZ,
                                           const float* sx, const float* sy, const float* sw, int n, int lane, unsigned mask) {
  double c = 0.0;
  for (int i = lane; i < n; i += PNP_G) {
    const double X = sX[i], Y = sY[i], Z = sZ[i];
    const double px = R[0] * X + R[1] * Y + R[2] * Z + t[0], py = R[3] * X + R[4] * Y + R[5] * Z + t[1], pz = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    const double iz = 1.0 / pz;
    const double ru = px * iz - sx[i], rv = py * iz - sy[i];
    c += sw[i] * (ru * ru + rv * rv);
  }
  return wsum(c, mask);
}

// ---- planar PnP of one view by a group of PNP_G lanes ------------------------------------------------------------------------
// Terms [b, b + n) of the problem, intrinsics prm (model MODEL), res = (ru, rv) or null.  On success (R, t) = T_camera_target.
// sX .. sw: the group's staging arrays (n_max floats each).  Every lane of the group returns the same values.
template <int MODEL>
__device__ __forceinline__ bool pnp_group(const DevProblem& p, const double* __restrict__ prm, const int* __restrict__ res, int b, int n, float* sX,
                                          float* sY, float* sZ, float* sx, float* sy, float* sw, int lane, unsigned mask, double R[9], double t[3]) {
  __syncwarp(mask);
  // -- back-projection and the 80 degree cone; everything passes through float as cv::Point2f / cv::Point3f do
  double cnt = 0.0, cX = 0.0, cY = 0.0, cZ = 0.0, cx = 0.0, cy = 0.0;
  for (int i = lane; i < n; i += PNP_G) {
    const double u = (double)(float)p.y_u[b + i], v = (double)(float)p.y_v[b + i];
    const double* tp = p.target + 3 * p.corner[b + i];
    double bp[3] = {0.0, 0.0, 0.0};
    bool ok = keypoint_to_euclidean<MODEL>(prm, res, u, v, bp);
    ok = ok && (bp[2] / sqrt(bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2]) > COS80);  // false for NaN as well
    const double X = (double)(float)tp[0], Y = (double)(float)tp[1], Z = (double)(float)tp[2];
    const double mx = ok ? (double)(float)(bp[0] / bp[2]) : 0.0, my = ok ? (double)(float)(bp[1] / bp[2]) : 0.0;
    const double w = ok ? 1.0 : 0.0;
    sX[i] = (float)X; sY[i] = (float)Y; sZ[i] = (float)Z; sx[i] = (float)mx; sy[i] = (float)my; sw[i] = (float)w;
    cnt += w; cX += w * X; cY += w * Y; cZ += w * Z; cx += w * mx; cy += w * my;
  }
  __syncwarp(mask);
  cnt = wsum(cnt, mask);
  if (cnt < 4.0) {  // "if (Ps.size() < 4) return false"
    return false;
  }
  const double icnt = 1.0 / cnt;
  cX = wsum(cX, mask) * icnt; cY = wsum(cY, mask) * icnt; cZ = wsum(cZ, mask) * icnt; cx = wsum(cx, mask) * icnt; cy = wsum(cy, mask) * icnt;
  // -- the object plane: principal axes of the corners (cvFindExtrinsicCameraParams2's planar branch)
  double Rt[9], Tt[3];
  {
    double m[6] = {0, 0, 0, 0, 0, 0};
    for (int i = lane; i < n; i += PNP_G) {
      const double w = sw[i], dx = sX[i] - cX, dy = sY[i] - cY, dz = sZ[i] - cZ;
      m[0] += w * dx * dx; m[1] += w * dx * dy; m[2] += w * dx * dz; m[3] += w * dy * dy; m[4] += w * dy * dz; m[5] += w * dz * dz;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) m[i] = wsum(m[i], mask);
    double A[9] = {m[0], m[1], m[2], m[1], m[3], m[4], m[2], m[4], m[5]}, ev[3], V[9];
    eig3(A, ev, V);
    // sort descending: the plane normal is the eigenvector of the smallest eigenvalue
    int i0 = 0, i2 = 0;
#pragma unroll
    for (int i = 1; i < 3; ++i) { if (ev[i] > ev[i0]) i0 = i; if (ev[i] < ev[i2]) i2 = i; }
    if (i0 == i2) { i0 = 0; i2 = 2; }
    const int i1 = 3 - i0 - i2;
    const bool planar = ev[i2] < 1e-3 * ev[i1];
    if (!planar) {  // kalibr's grid targets are planar; a general 3-D object would need the 12-parameter DLT start
      return false;
    }
    if (V[0 * 3 + i2] * V[0 * 3 + i2] + V[1 * 3 + i2] * V[1 * 3 + i2] < 1e-10) {  // normal along z already
#pragma unroll
      for (int i = 0; i < 9; ++i) Rt[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) { Rt[0 * 3 + c] = V[c * 3 + i0]; Rt[1 * 3 + c] = V[c * 3 + i1]; Rt[2 * 3 + c] = V[c * 3 + i2]; }
      const double det = Rt[0] * (Rt[4] * Rt[8] - Rt[5] * Rt[7]) - Rt[1] * (Rt[3] * Rt[8] - Rt[5] * Rt[6]) + Rt[2] * (Rt[3] * Rt[7] - Rt[4] * Rt[6]);
      if (det < 0.0)
#pragma unroll
        for (int i = 0; i < 9; ++i) Rt[i] = -Rt[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) Tt[i] = -(Rt[i * 3] * cX + Rt[i * 3 + 1] * cY + Rt[i * 3 + 2] * cZ);
  }
  // -- plane-to-image homography: normalised DLT with h33 = 1 (8 x 8 normal equations)
  {
    // per-axis scales: count / sum |deviation| (the normalisation of OpenCV's homography kernel)
    double aX = 0.0, aY = 0.0, ax = 0.0, ay = 0.0;
    for (int i = lane; i < n; i += PNP_G) {
      const double w = sw[i];
      const double X = Rt[0] * sX[i] + Rt[1] * sY[i] + Rt[2] * sZ[i] + Tt[0], Y = Rt[3] * sX[i] + Rt[4] * sY[i] + Rt[5] * sZ[i] + Tt[1];
      aX += w * fabs(X); aY += w * fabs(Y); ax += w * fabs(sx[i] - cx); ay += w * fabs(sy[i] - cy);
    }
    const double sMx = cnt / wsum(aX, mask), sMy = cnt / wsum(aY, mask), smx = cnt / wsum(ax, mask), smy = cnt / wsum(ay, mask);
    // L^T L of the rows [X Y 1 0 0 0 -xX -xY -x], [0 0 0 X Y 1 -yX -yY -y] in blocks: A = sum q q^T, B = sum x q q^T,
    // C = sum y q q^T, D = sum (x^2 + y^2) q q^T with q = (X, Y, 1)
    double Aq[6] = {0, 0, 0, 0, 0, 0}, Bq[6] = {0, 0, 0, 0, 0, 0}, Cq[6] = {0, 0, 0, 0, 0, 0}, Dq[6] = {0, 0, 0, 0, 0, 0};
    for (int i = lane; i < n; i += PNP_G) {
      const double w = sw[i];
      const double X = (Rt[0] * sX[i] + Rt[1] * sY[i] + Rt[2] * sZ[i] + Tt[0]) * sMx, Y = (Rt[3] * sX[i] + Rt[4] * sY[i] + Rt[5] * sZ[i] + Tt[1]) * sMy;
      const double x = (sx[i] - cx) * smx, y = (sy[i] - cy) * smy;
      const double qq[6] = {w * X * X, w * X * Y, w * X, w * Y * Y, w * Y, w};
      const double r2 = x * x + y * y;
#pragma unroll
      for (int k = 0; k < 6; ++k) { Aq[k] += qq[k]; Bq[k] += x * qq[k]; Cq[k] += y * qq[k]; Dq[k] += r2 * qq[k]; }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { Aq[k] = wsum(Aq[k], mask); Bq[k] = wsum(Bq[k], mask); Cq[k] = wsum(Cq[k], mask); Dq[k] = wsum(Dq[k], mask); }
    // symmetric 3x3 blocks from the 6 unique sums: index (r, c) -> {0: XX, 1: XY, 2: X, 3: YY, 4: Y, 5: 1}
    const int sidx[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    double N[64], rhs[8];
#pragma unroll
    for (int i = 0; i < 64; ++i) N[i] = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int s = sidx[r * 3 + c];
        N[r * 8 + c] = Aq[s];
        N[(3 + r) * 8 + 3 + c] = Aq[s];
        if (c < 2) {
          N[r * 8 + 6 + c] = -Bq[s];
          N[(6 + c) * 8 + r] = -Bq[s];
          N[(3 + r) * 8 + 6 + c] = -Cq[s];
          N[(6 + c) * 8 + 3 + r] = -Cq[s];
        }
        if (r < 2 && c < 2) N[(6 + r) * 8 + 6 + c] = Dq[s];
      }
    // right-hand side: -L^T L[0:8][8] = -(column of -x q, -y q, (x^2+y^2) q) at q-index 2 (the "1" entry)
#pragma unroll
    for (int r = 0; r < 3; ++r) { rhs[r] = Bq[sidx[r * 3 + 2]]; rhs[3 + r] = Cq[sidx[r * 3 + 2]]; }
    rhs[6] = -Dq[sidx[0 * 3 + 2]];
    rhs[7] = -Dq[sidx[1 * 3 + 2]];
    const bool spd = chol_solve<8>(N, rhs);
    if (!spd) {
      return false;
    }
    // de-normalise: H = Tm^-1 Hn TM, Tm^-1 = [1/smx 0 cx; 0 1/smy cy; 0 0 1], TM = diag(sMx, sMy, 1) (the plane points are centred)
    const double Hn[9] = {rhs[0], rhs[1], rhs[2], rhs[3], rhs[4], rhs[5], rhs[6], rhs[7], 1.0};
    double H[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double sc = c == 0 ? sMx : c == 1 ? sMy : 1.0;
      H[0 * 3 + c] = (Hn[0 * 3 + c] / smx + cx * Hn[2 * 3 + c]) * sc;
      H[1 * 3 + c] = (Hn[1 * 3 + c] / smy + cy * Hn[2 * 3 + c]) * sc;
      H[2 * 3 + c] = Hn[2 * 3 + c] * sc;
    }
    if (H[8] < 0.0)
#pragma unroll
      for (int i = 0; i < 9; ++i) H[i] = -H[i];
    const double n1 = sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]), n2 = sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
    const double i1 = 1.0 / fmax(n1, 2.220446049250313e-16), i2 = 1.0 / fmax(n2, 2.220446049250313e-16), i3 = 2.0 / fmax(n1 + n2, 2.220446049250313e-16);
    const double r1[3] = {H[0] * i1, H[3] * i1, H[6] * i1}, r2[3] = {H[1] * i2, H[4] * i2, H[7] * i2};
    double R0[9] = {r1[0], r2[0], r1[1] * r2[2] - r1[2] * r2[1], r1[1], r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[2], r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    nearest_rotation(R0);
    mat3_mul(R0, Rt, R);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = R0[i * 3] * Tt[0] + R0[i * 3 + 1] * Tt[1] + R0[i * 3 + 2] * Tt[2] + H[i * 3 + 2] * i3;
  }
  // -- Levenberg-Marquardt on the reprojection error in normalised coordinates, update R <- exp(d_theta) R, t <- t + d_t
  double cost = pnp_cost(R, t, sX, sY, sZ, sx, sy, sw, n, lane, mask);
  double lambda = 1e-3;
  bool good = isfinite(cost);
  for (int it = 0; it < 60 && good; ++it) {
    double Hs[21], g[6];
#pragma unroll
    for (int i = 0; i < 21; ++i) Hs[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = 0.0;
    for (int i = lane; i < n; i += PNP_G) {
      const double w = sw[i], X = sX[i], Y = sY[i], Z = sZ[i];
      const double qx = R[0] * X + R[1] * Y + R[2] * Z, qy = R[3] * X + R[4] * Y + R[5] * Z, qz = R[6] * X + R[7] * Y + R[8] * Z;
      const double px = qx + t[0], py = qy + t[1], pz = qz + t[2];
      const double iz = 1.0 / pz, u = px * iz, v = py * iz;
      const double ru = u - sx[i], rv = v - sy[i];
      // d(u, v)/dp = [iz 0 -u iz; 0 iz -v iz];  dp/d(theta) = -[q]x, dp/dt = I
      const double a = iz, bu = -u * iz, bv = -v * iz;
      const double Ju[6] = {bu * qy, a * qz - bu * qx, -a * qy, a, 0.0, bu};
      const double Jv[6] = {-a * qz + bv * qy, -bv * qx, a * qx, 0.0, a, bv};
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
#pragma unroll
        for (int c = r; c < 6; ++c) Hs[k++] += w * (Ju[r] * Ju[c] + Jv[r] * Jv[c]);
        g[r] += w * (Ju[r] * ru + Jv[r] * rv);
      }
    }
#pragma unroll
    for (int i = 0; i < 21; ++i) Hs[i] = wsum(Hs[i], mask);
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = wsum(g[i], mask);
    bool accepted = false, flat = false;
    double step_max = 0.0;
    for (int tries = 0; tries < 12 && !accepted; ++tries) {
      double A[36], d[6];
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) {
          const double v = Hs[k++];
          A[r * 6 + c] = v;
          A[c * 6 + r] = v;
        }
#pragma unroll
      for (int r = 0; r < 6; ++r) { A[r * 6 + r] *= 1.0 + lambda; d[r] = -g[r]; }
      if (!chol_solve<6>(A, d)) { lambda *= 10.0; continue; }
      double Rn[9];
      rotate_left(d, R, Rn);
      const double tn[3] = {t[0] + d[3], t[1] + d[4], t[2] + d[5]};
      const double cn = pnp_cost(Rn, tn, sX, sY, sZ, sx, sy, sw, n, lane, mask);
      if (cn <= cost) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = Rn[i];
        t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
        flat = (cost - cn) <= 1e-15 * cost;  // the cost has stopped moving at double precision
        cost = cn;
        lambda = fmax(lambda * 0.1, 1e-15);
        accepted = true;
#pragma unroll
        for (int i = 0; i < 6; ++i) step_max = fmax(step_max, fabs(d[i]));
      } else {
        lambda *= 10.0;
      }
    }
    if (!accepted || flat) break;  // no descent left at this precision: converged
    // the iteration contracts at least as fast as lambda (<= 1e-4 by now) near the minimum: after a step this small the next one
    // is below rounding
    if (step_max < 1e-10 * fmax(1.0, fmax(fabs(t[0]), fmax(fabs(t[1]), fabs(t[2]))))) break;
  }
  nearest_rotation(R);  // rounding drift of the multiplicative updates
  return good && isfinite(cost);
}

// ---- estimateTransformation for a list of views ---------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(PNP_WARPS * 32, KB_PNP_MINB) pnp_kernel(DevProblem p, const int* __restrict__ view_list, int n_list,
                                                              const unsigned char* __restrict__ view_mask, const int* __restrict__ resolution,
                                                              int n_max, double* __restrict__ T_out, int* __restrict__ ok_out) {
  // the staged values have all passed through float (cv::Point2f / Point3f), so float storage is exact: 24 bytes per corner
  extern __shared__ float smem_pnp[];
  const int lane = threadIdx.x & (PNP_G - 1), gib = threadIdx.x / PNP_G;
  const unsigned mask = ((1u << PNP_G) - 1u) << ((threadIdx.x & 31) & ~(PNP_G - 1));
  float* sX = smem_pnp + (size_t)gib * 6 * n_max;
  float* sY = sX + n_max;
  float* sZ = sY + n_max;
  float* sx = sZ + n_max;
  float* sy = sx + n_max;
  float* sw = sy + n_max;
  const int warp = blockIdx.x * PNP_GROUPS + gib, n_warps = gridDim.x * PNP_GROUPS;
  for (int vi = warp; vi < n_list; vi += n_warps) {
    const int view = view_list[vi];
    if (view_mask && !view_mask[view]) continue;
    const int cam = p.view_cam[view];
    const int b = p.view_begin[view];
    const int n = min(p.view_begin[view + 1] - b, n_max);
    double prm[CAM_PARAM_STRIDE];
#pragma unroll
    for (int i = 0; i < CAM_PARAM_STRIDE; ++i) prm[i] = p.cam_params[cam * CAM_PARAM_STRIDE + i];
    const int* res = resolution ? resolution + 2 * cam : nullptr;
    double R[9], t[3];
    const bool ok = pnp_group<MODEL>(p, prm, res, b, n, sX, sY, sZ, sx, sy, sw, lane, mask, R, t);
    // -- T_target_camera = inverse([R t]) -> (q, t) as sm::kinematics::Transformation::set stores it
    if (lane == 0) {
      double* To = T_out + (size_t)view * POSE_STRIDE;
      if (!ok) {  // estimateTransformation returned false
        To[0] = To[1] = To[2] = 0.0; To[3] = 1.0; To[4] = To[5] = To[6] = 0.0;
        ok_out[view] = 0;
      } else {
        const double Ri[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
        double q[4];
        r2quat(Ri, q);
        To[0] = q[0]; To[1] = q[1]; To[2] = q[2]; To[3] = q[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) To[4 + i] = -(Ri[i * 3] * t[0] + Ri[i * 3 + 1] * t[1] + Ri[i * 3 + 2] * t[2]);
        ok_out[view] = 1;
      }
    }
  }
}

// ---- initializeIntrinsics, pinhole family: focal-length guesses from the vanishing points of the grid rows ----------------
// ≙ PinholeProjection::initializeIntrinsics (CAM/.../implementation/PinholeProjection.hpp:713-803) with PinholeHelpers::fitCircle
// (:649-696) and intersectCircles (:612-647).  One warp per view of the camera: a lane per grid row fits its circle, a lane per row
// pair intersects two circles; fg[view][pair] = |v1 - v2| / pi or NaN.  The reference's pair loop lets k run to cols while its
// arrays have `rows` entries (undefined behaviour when cols > rows); pairs are taken among the rows that exist (quirk Q11).
constexpr int FG_WARPS = 4;
__global__ void __launch_bounds__(FG_WARPS * 32) focal_guess_kernel(DevProblem p, const int* __restrict__ cam_views, int n_views, int rows, int cols,
                                                                    int n_pairs, double* __restrict__ fg) {
  extern __shared__ double smem_fg[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int n_grid = rows * cols;
  double* ccx = smem_fg + (size_t)wib * (3 * rows + (n_grid + 1) / 2);
  double* ccy = ccx + rows;
  double* crr = ccy + rows;
  int* idx = reinterpret_cast<int*>(crr + rows);
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const int rr = min(rows, cols);
  for (int vi = blockIdx.x * FG_WARPS + wib; vi < n_views; vi += gridDim.x * FG_WARPS) {
    const int view = cam_views[vi];
    const int b = p.view_begin[view], e = p.view_begin[view + 1];
    __syncwarp();
    for (int i = lane; i < n_grid; i += 32) idx[i] = -1;
    __syncwarp();
    for (int i = b + lane; i < e; i += 32) {
      const int c = p.corner[i];
      if (c < n_grid) idx[c] = i;
    }
    __syncwarp();
    bool complete = true;
    for (int i = lane; i < n_grid; i += 32) complete = complete && idx[i] >= 0;
    complete = __all_sync(0xffffffffu, complete);
    double* out = fg + (size_t)vi * n_pairs;
    if (!complete) {  // "skip this image if the board view is not complete"
      for (int q = lane; q < n_pairs; q += 32) out[q] = nan;
      continue;
    }
    for (int r = lane; r < rows; r += 32) {
      double sx = 0, sy = 0, sxx = 0, sxy = 0, syy = 0, sxxx = 0, sxxy = 0, sxyy = 0, syyy = 0;
      for (int c = 0; c < cols; ++c) {
        const int i = idx[r * cols + c];
        const double x = (double)(float)p.y_u[i], y = (double)(float)p.y_v[i];  // cv::Point2f
        sx += x; sy += y; sxx += x * x; sxy += x * y; syy += y * y;
        sxxx += x * x * x; sxxy += x * x * y; sxyy += x * y * y; syyy += y * y * y;
      }
      const double n = (double)cols;
      const double A = n * sxx - sx * sx, B = n * sxy - sx * sy, C = n * syy - sy * sy;
      const double D = 0.5 * (n * sxyy - sx * syy + n * sxxx - sx * sxx), E = 0.5 * (n * sxxy - sy * sxx + n * syyy - sy * syy);
      const double cx = (D * C - B * E) / (A * C - B * B), cy = (A * E - B * D) / (A * C - B * B);
      double sr = 0.0;
      for (int c = 0; c < cols; ++c) {
        const int i = idx[r * cols + c];
        const double dx = (double)(float)p.y_u[i] - cx, dy = (double)(float)p.y_v[i] - cy;
        sr += sqrt(dx * dx + dy * dy);
      }
      ccx[r] = cx; ccy[r] = cy; crr[r] = sr / n;
    }
    __syncwarp();
    for (int q = lane; q < n_pairs; q += 32) {
      int j = 0, rem = q;
      while (rem >= rr - 1 - j) { rem -= rr - 1 - j; ++j; }
      const int k = j + 1 + rem;
      const double x1 = ccx[j], y1 = ccy[j], r1 = crr[j], x2 = ccx[k], y2 = ccy[k], r2 = crr[k];
      double f = nan;
      const double d = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
      if (!(d > r1 + r2) && !(d < fabs(r1 - r2))) {
        const double a = (r1 * r1 - r2 * r2 + d * d) / (2.0 * d);
        const double h = sqrt(r1 * r1 - a * a);
        if (!(h < 1e-10)) {  // two intersection points (a NaN h falls through to a NaN guess, which isfinite drops)
          // the two points are (x3 +- h (y2 - y1) / d, y3 -+ h (x2 - x1) / d): their distance
          const double px = (x1 + a * (x2 - x1) / d + h * (y2 - y1) / d) - (x1 + a * (x2 - x1) / d - h * (y2 - y1) / d);
          const double py = (y1 + a * (y2 - y1) / d - h * (x2 - x1) / d) - (y1 + a * (y2 - y1) / d + h * (x2 - x1) / d);
          f = sqrt(px * px + py * py) / 3.14159265358979323846;
        }
      }
      out[q] = f;
    }
  }
}

// Median of the finite entries of v as PinholeHelpers::medianOfVectorElements (:698-707: the mean of the two middle values for an
// even count), by radix selection on the IEEE bit patterns (positive doubles order like their bits): one CTA, 8 digit passes per
// order statistic, integer histograms only (exact and order-independent).  out[0] = median, out[1] = number of finite entries.
__global__ void __launch_bounds__(1024) select_median_kernel(const double* __restrict__ v, long long m, double* __restrict__ out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix, s_k, s_count;
  if (threadIdx.x == 0) s_count = 0ull;
  __syncthreads();
  unsigned long long local = 0;
  for (long long i = threadIdx.x; i < m; i += blockDim.x) {
    const double x = v[i];
    if (isfinite(x) && x > 0.0) ++local;
  }
  atomicAdd(&s_count, local);
  __syncthreads();
  const unsigned long long n = s_count;
  if (n == 0ull) {
    if (threadIdx.x == 0) { out[0] = 0.0; out[1] = 0.0; }
    return;
  }
  double result[2] = {0.0, 0.0};
  const int n_sel = (n % 2ull == 0ull) ? 2 : 1;
  for (int sel = 0; sel < n_sel; ++sel) {
    if (threadIdx.x == 0) { s_prefix = 0ull; s_k = (n_sel == 2) ? (n / 2ull - 1ull + (unsigned long long)sel) : n / 2ull; }
    for (int digit = 7; digit >= 0; --digit) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (long long i = threadIdx.x; i < m; i += blockDim.x) {
        const double x = v[i];
        if (!(isfinite(x) && x > 0.0)) continue;
        const unsigned long long key = (unsigned long long)__double_as_longlong(x);
        if (digit == 7 || (key >> (8 * (digit + 1))) == prefix) atomicAdd(&hist[(key >> (8 * digit)) & 255ull], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long k = s_k, cum = 0ull;
        int bin = 0;
        for (; bin < 256; ++bin) {
          if (cum + hist[bin] > k) break;
          cum += hist[bin];
        }
        s_k = k - cum;
        s_prefix = (prefix << 8) | (unsigned long long)bin;
      }
      __syncthreads();
    }
    result[sel] = __longlong_as_double((long long)s_prefix);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = n_sel == 2 ? (result[0] + result[1]) / 2.0 : result[0];
    out[1] = (double)n;
  }
}

// ---- initializeIntrinsics, omni family: one candidate per (view, grid row) ----------------------------------------------------
// ≙ OmniProjection::initializeIntrinsics (CAM/.../implementation/OmniProjection.hpp:721-846), which the EUCM and double-sphere
// projections delegate to (ExtendedUnifiedProjection.hpp:734-760, DoubleSphereProjection.hpp:785-811).  A group of PNP_G lanes per
// candidate: the row's corners as a line image -> null vector of the n x 4 system (cv::SVD::solveZ: one-sided Jacobi on the
// columns, rows spread over the lanes) -> gamma; then estimateTransformation with (xi = 1, gamma, image centre, no distortion) and
// the mean reprojection error of the view (computeReprojectionError, :848-866).  cand[item] = (gamma, mean error or +inf).
__global__ void __launch_bounds__(PNP_WARPS * 32, 3) omni_candidate_kernel(DevProblem p, const int* __restrict__ cam_views, int n_views, int rows, int cols,
                                                                           int ru, int rv, int n_max, double* __restrict__ cand) {
  extern __shared__ float smem_pnp[];
  const int lane = threadIdx.x & (PNP_G - 1), gib = threadIdx.x / PNP_G;
  const unsigned mask = ((1u << PNP_G) - 1u) << ((threadIdx.x & 31) & ~(PNP_G - 1));
  const int per_group = 6 * n_max + 2 * 2 * cols + 2 * cols;  // PnP staging | (u, v) of the row as doubles | presence flags (+ padding)
  float* sX = smem_pnp + (size_t)gib * per_group;
  float* sY = sX + n_max;
  float* sZ = sY + n_max;
  float* sx = sZ + n_max;
  float* sy = sx + n_max;
  float* sw = sy + n_max;
  double* ru_row = reinterpret_cast<double*>(sw + n_max);
  double* rv_row = ru_row + cols;
  int* present = reinterpret_cast<int*>(rv_row + cols);
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double cu = ((double)ru - 1.0) / 2.0, cv = ((double)rv - 1.0) / 2.0;
  const int n_items = n_views * rows;
  for (int item = blockIdx.x * PNP_GROUPS + gib; item < n_items; item += gridDim.x * PNP_GROUPS) {
    const int vi = item / rows, r = item % rows;
    const int view = cam_views[vi];
    const int b = p.view_begin[view];
    const int n = min(p.view_begin[view + 1] - b, n_max);
    double gamma = 0.0, avg = inf;
    __syncwarp(mask);
    for (int c = lane; c < cols; c += PNP_G) present[c] = 0;
    __syncwarp(mask);
    for (int i = lane; i < n; i += PNP_G) {
      const int cid = p.corner[b + i];
      if (cid / cols == r) {
        const int c = cid % cols;
        ru_row[c] = p.y_u[b + i] - cu;
        rv_row[c] = p.y_v[b + i] - cv;
        present[c] = 1;
      }
    }
    __syncwarp(mask);
    // rows of P = [u, v, 0.5, -0.5 (u^2 + v^2)]: lane l holds the present corners c = l, l + G, ... (at most 4: cols <= 4 G)
    double A[4][4];
    int count = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + q * PNP_G;
      const bool on = c < cols && present[c];
      const double u = on ? ru_row[c] : 0.0, v = on ? rv_row[c] : 0.0;
      A[q][0] = u; A[q][1] = v; A[q][2] = on ? 0.5 : 0.0; A[q][3] = on ? -0.5 * (u * u + v * v) : 0.0;
      count += on ? 1 : 0;
    }
#pragma unroll
    for (int o = PNP_G / 2; o > 0; o >>= 1) count += __shfl_xor_sync(mask, count, o);
    bool have = count > 4;  // MIN_CORNERS
    if (have) {
      double V[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
      for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int c2 = a + 1; c2 < 4; ++c2) {
            double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { al += A[q][a] * A[q][a]; be += A[q][c2] * A[q][c2]; ga += A[q][a] * A[q][c2]; }
            al = wsum(al, mask); be = wsum(be, mask); ga = wsum(ga, mask);
            if (fabs(ga) > 1e-16 * sqrt(al * be)) {
              rotated = true;
              const double zeta = (be - al) / (2.0 * ga);
              const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
              const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const double xa = A[q][a], xb = A[q][c2];
                A[q][a] = cs * xa - sn * xb;
                A[q][c2] = sn * xa + cs * xb;
                const double va = V[q][a], vb = V[q][c2];
                V[q][a] = cs * va - sn * vb;
                V[q][c2] = sn * va + cs * vb;
              }
            }
          }
        if (!rotated) break;
      }
      double nrm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double s2 = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) s2 += A[q][j] * A[q][j];
        nrm[j] = wsum(s2, mask);
      }
      int jm = 0;
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (nrm[j] < nrm[jm]) jm = j;
      double C[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) C[i] = jm == 0 ? V[i][0] : jm == 1 ? V[i][1] : jm == 2 ? V[i][2] : V[i][3];
      const double tq = C[0] * C[0] + C[1] * C[1] + C[2] * C[3];
      have = !(tq < 0.0);
      if (have) {
        const double d = sqrt(1.0 / tq);
        const double nx = C[0] * d, ny = C[1] * d;
        have = !(sqrt(nx * nx + ny * ny) > 0.95);  // radial line
        if (have) gamma = fabs(C[2] * d / sqrt(1.0 - nx * nx - ny * ny));
      }
    }
    if (have) {
      double prm[CAM_PARAM_STRIDE] = {1.0, gamma, gamma, cu, cv, 0.0, 0.0, 0.0, 0.0, 0.0};
      double R[9], t[3];
      if (pnp_group<OMNI_NONE>(p, prm, nullptr, b, n, sX, sY, sZ, sx, sy, sw, lane, mask, R, t)) {
        double err = 0.0, cnt = 0.0;
        for (int i = lane; i < n; i += PNP_G) {
          const double* tp = p.target + 3 * p.corner[b + i];
          const double px = R[0] * tp[0] + R[1] * tp[1] + R[2] * tp[2] + t[0], py = R[3] * tp[0] + R[4] * tp[1] + R[5] * tp[2] + t[1],
                       pz = R[6] * tp[0] + R[7] * tp[1] + R[8] * tp[2] + t[2];
          const double d = sqrt(px * px + py * py + pz * pz);
          if (pz <= -d) continue;  // _fov_parameter = 1 for xi = 1
          const double rz = 1.0 / (pz + d);
          const double yu = gamma * (px * rz) + cu, yv = gamma * (py * rz) + cv;
          if (!(yu >= 0.0 && yu < (double)ru && yv >= 0.0 && yv < (double)rv)) continue;
          const double du = p.y_u[b + i] - yu, dv = p.y_v[b + i] - yv;
          err += sqrt(du * du + dv * dv);
          cnt += 1.0;
        }
        err = wsum(err, mask);
        cnt = wsum(cnt, mask);
        if (cnt > 4.0) avg = err / cnt;
      }
    }
    if (lane == 0) {
      cand[2 * (size_t)item] = gamma;
      cand[2 * (size_t)item + 1] = avg;
    }
  }
}

// first minimum of the candidates' mean reprojection error in (view, row) order: out[0] = its gamma, out[1] = 1 if there is one
__global__ void __launch_bounds__(1024) best_candidate_kernel(const double* __restrict__ cand, int n_items, double* __restrict__ out) {
  __shared__ double s_err[1024];
  __shared__ int s_idx[1024];
  double best = __longlong_as_double(0x7ff0000000000000ll);
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
    const double e = cand[2 * (size_t)i + 1];
    if (e < best) { best = e; bi = i; }  // strided scan: indices ascend, so the first minimum of this thread's subsequence is kept
  }
  s_err[threadIdx.x] = best;
  s_idx[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double e2 = s_err[threadIdx.x + o];
      const int i2 = s_idx[threadIdx.x + o];
      if (e2 < s_err[threadIdx.x] || (e2 == s_err[threadIdx.x] && i2 < s_idx[threadIdx.x])) { s_err[threadIdx.x] = e2; s_idx[threadIdx.x] = i2; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const bool ok = s_idx[0] != 0x7fffffff && s_err[0] < 1.7976931348623157e308;
    out[0] = ok ? cand[2 * (size_t)s_idx[0]] : 0.0;
    out[1] = ok ? 1.0 : 0.0;
    out[2] = ok ? s_err[0] : 0.0;
  }
}

// ---- getTargetPoseGuess: one thread per synced set -------------------------------------------------------------------------
__global__ void __launch_bounds__(128) set_pose_guess_kernel(DevProblem p, const double* __restrict__ T_views, const int* __restrict__ ok_views,
                                                             double* __restrict__ set_poses_out, int* __restrict__ set_ok) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_sets) return;
  // camera with most corners, the first one on ties (std::max_element)
  int best = -1, best_n = 0, best_view = -1;
  for (int k = 0; k < p.n_cams; ++k) {
    const int w = p.set_view[(size_t)s * p.n_cams + k];
    const int n = w >= 0 ? p.view_begin[w + 1] - p.view_begin[w] : 0;
    if (best < 0 || n > best_n) { best = k; best_n = n; best_view = w; }
  }
  double* o = set_poses_out + (size_t)s * POSE_STRIDE;
  if (best_view < 0) {  // nobody saw this set: keep the pose that is there
    set_ok[s] = 0;
    return;
  }
  const double* Tv = T_views + (size_t)best_view * POSE_STRIDE;
  double R[9], t[3] = {Tv[4], Tv[5], Tv[6]};
  quat2r_rm(Tv, R);
  // T_t_c0 = ((T_t_cN B_0) B_1) ... B_{N-1}: std::accumulate over baseline_guesses[0 .. N) in that order (CalibrationTools.hpp:352-353)
  for (int j = 0; j < best; ++j) {
    const double* b = p.baselines + (size_t)j * POSE_STRIDE;
    double Rb[9], Rn[9];
    quat2r_rm(b, Rb);
    mat3_mul(R, Rb, Rn);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] += R[i * 3] * b[4] + R[i * 3 + 1] * b[5] + R[i * 3 + 2] * b[6];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  }
  double q[4];
  r2quat(R, q);
  o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; o[4] = t[0]; o[5] = t[1]; o[6] = t[2];
  set_ok[s] = ok_views[best_view];
}

// per view: 1 when the view is the one getTargetPoseGuess picks for its set
__global__ void __launch_bounds__(128) best_view_mask_kernel(DevProblem p, unsigned char* __restrict__ mask) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_sets) return;
  int best_n = 0, best_view = -1;
  bool first = true;
  for (int k = 0; k < p.n_cams; ++k) {
    const int w = p.set_view[(size_t)s * p.n_cams + k];
    const int n = w >= 0 ? p.view_begin[w + 1] - p.view_begin[w] : 0;
    if (w >= 0) mask[w] = 0;
    if (first || n > best_n) { best_n = n; best_view = w; first = false; }
  }
  if (best_view >= 0) mask[best_view] = 1;
}

template <int MODEL>
cudaError_t launch_pnp_model(const DevProblem& p, const int* list, int n, const unsigned char* mask, const int* resolution, int n_max, double* T_out,
                             int* ok_out, StreamCtx& s) {
  if (n <= 0) return cudaSuccess;
  const size_t smem = sizeof(float) * 6 * (size_t)n_max * PNP_GROUPS;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pnp_kernel<MODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = min((n + PNP_GROUPS - 1) / PNP_GROUPS, sms * 8);
  pnp_kernel<MODEL><<<grid, PNP_WARPS * 32, smem, s.stream>>>(p, list, n, mask, resolution, n_max, T_out, ok_out);
  ++*s.launches;
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_estimate_transformations(const DevProblem& p, const int* view_list, const int* mb, const unsigned char* view_mask,
                                            const int* resolution, double* T_out, int* ok_out, StreamCtx& s) {
  const int n_max = p.n_target;
  for (int m = 0; m < NUM_MODELS; ++m) {
    cudaError_t e = cudaSuccess;
    const int n = mb[m + 1] - mb[m];
    switch (m) {
      case 0: e = launch_pnp_model<0>(p, view_list + mb[0], n, view_mask, resolution, n_max, T_out, ok_out, s); break;
      case 1: e = launch_pnp_model<1>(p, view_list + mb[1], n, view_mask, resolution, n_max, T_out, ok_out, s); break;
      case 2: e = launch_pnp_model<2>(p, view_list + mb[2], n, view_mask, resolution, n_max, T_out, ok_out, s); break;
      case 3: e = launch_pnp_model<3>(p, view_list + mb[3], n, view_mask, resolution, n_max, T_out, ok_out, s); break;
      case 4: e = launch_pnp_model<4>(p, view_list + mb[4], n, view_mask, resolution, n_max, T_out, ok_out, s); break;
      case 5: e = launch_pnp_model<5>(p, view_list + mb[5], n, view_mask, resolution, n_max, T_out, ok_out, s); break;
      case 6: e = launch_pnp_model<6>(p, view_list + mb[6], n, view_mask, resolution, n_max, T_out, ok_out, s); break;
    }
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

cudaError_t launch_focal_guesses(const DevProblem& p, const int* cam_views, int n_views, int rows, int cols, double* fg, double* out2, StreamCtx& s) {
  const int rr = rows < cols ? rows : cols;
  const int n_pairs = rr * (rr - 1) / 2;
  if (n_views <= 0 || n_pairs <= 0) {
    cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(double), s.stream);
    return e;
  }
  const size_t smem = sizeof(double) * FG_WARPS * (size_t)(3 * rows + (rows * cols + 1) / 2);
  focal_guess_kernel<<<min((n_views + FG_WARPS - 1) / FG_WARPS, 148 * 8), FG_WARPS * 32, smem, s.stream>>>(p, cam_views, n_views, rows, cols, n_pairs, fg);
  ++*s.launches;
  select_median_kernel<<<1, 1024, 0, s.stream>>>(fg, (long long)n_views * n_pairs, out2);
  ++*s.launches;
  return cudaGetLastError();
}

cudaError_t launch_omni_candidates(const DevProblem& p, const int* cam_views, int n_views, int rows, int cols, int ru, int rv, double* cand, double* out3,
                                   StreamCtx& s) {
  if (n_views <= 0) return cudaMemsetAsync(out3, 0, 3 * sizeof(double), s.stream);
  const int n_max = p.n_target;
  const size_t smem = sizeof(float) * (size_t)(6 * n_max + 6 * cols) * PNP_GROUPS;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(omni_candidate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int n_items = n_views * rows;
  omni_candidate_kernel<<<min((n_items + PNP_GROUPS - 1) / PNP_GROUPS, 148 * 8), PNP_WARPS * 32, smem, s.stream>>>(p, cam_views, n_views, rows, cols, ru, rv,
                                                                                                                 n_max, cand);
  ++*s.launches;
  best_candidate_kernel<<<1, 1024, 0, s.stream>>>(cand, n_items, out3);
  ++*s.launches;
  return cudaGetLastError();
}

cudaError_t launch_best_view_mask(const DevProblem& p, unsigned char* mask, StreamCtx& s) {
  if (p.n_sets <= 0) return cudaSuccess;
  best_view_mask_kernel<<<(p.n_sets + 127) / 128, 128, 0, s.stream>>>(p, mask);
  ++*s.launches;
  return cudaGetLastError();
}

cudaError_t launch_set_pose_guess(const DevProblem& p, const double* T_views, const int* ok_views, double* set_poses_out, int* set_ok, StreamCtx& s) {
  if (p.n_sets <= 0) return cudaSuccess;
  set_pose_guess_kernel<<<(p.n_sets + 127) / 128, 128, 0, s.stream>>>(p, T_views, ok_views, set_poses_out, set_ok);
  ++*s.launches;
  return cudaGetLastError();
}

}  // namespace kb
