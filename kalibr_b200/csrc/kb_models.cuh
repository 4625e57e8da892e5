// Device camera models: forward projection of a point in the camera frame plus the three Jacobians the
// reprojection term needs (w.r.t. the point, the projection parameters and the distortion parameters), each
// evaluated ONCE per term with shared sub-expressions (the reference re-runs the distortion three times).
//
// Follows aslam_cv/aslam_cameras/include/aslam/cameras/implementation/ (CAM):
//   PinholeProjection.hpp:99-145, 326-378     OmniProjection.hpp:118-180, 383-445
//   ExtendedUnifiedProjection.hpp:133-198, 399-455 (incl. the fu-for-fv quirk Q4)
//   DoubleSphereProjection.hpp:142-221, 445-503
//   RadialTangentialDistortion.hpp:28-65, 153-182   EquidistantDistortion.hpp:32-183, 244-273   FovDistortion.hpp:23-92, 133-170
// Parameter order per model = kb_camera_model in include/kalibr_b200.h.
#pragma once
#include <cuda_runtime.h>

#include "kb_device.cuh"

namespace kb {


// Result of linearising one term in the camera frame.  Jp: d(y_hat)/d(p) 2x3 (the 4th homogeneous column is
// always zero in the reference), Ji: 2xP, Jd: 2xD.  valid = 0 reproduces "projection returned false before
// writing y_hat" (SURVEY.md Q6): the caller zero-weights the term.
template <int P, int D>
struct Linearisation {
  double y[2];
  double Jp[2][3];
  double Ji[2][P > 0 ? P : 1];
  double Jd[2][D > 0 ? D : 1];
  bool valid;
};

// ---- distortions ---------------------------------------------------------------------------------------
// radtan: in/out m (normalised point), Jm = d(distorted)/d(m) (symmetric 2x2), Jk = d(distorted)/d(k1,k2,p1,p2) at m
// sign flip on the integer pipe: the FP64 units are the bottleneck of every kernel that calls this
__device__ __forceinline__ double neg_int(double v) { return __hiloint2double(__double2hiint(v) ^ (int)0x80000000, __double2loint(v)); }

template <bool WITH_JAC>
__device__ __forceinline__ void radtan(const double* __restrict__ k, double& mx, double& my, double Jm[2][2], double Jk[2][4]) {
  // CAM/RadialTangentialDistortion.hpp:24-101 with the common subexpressions shared: 1 + rad = 1 + r^2 (k1 + k2 r^2),
  // d(rad)/d(r^2) = k1 + 2 k2 r^2; the reference's sums re-associated (differences at the 1e-16 level)
  const double k1 = k[0], k2 = k[1], p1 = k[2], p2 = k[3];
  const double tp1 = p1 + p1, tp2 = p2 + p2;
  const double x = mx, y = my;
  const double x2 = x * x, y2 = y * y, xy = x * y;
  const double rho2 = x2 + y2;
  const double k2r2 = k2 * rho2;
  const double t = k1 + k2r2;
  const double a = fma(t, rho2, 1.0);          // 1 + rad
  const double s3x = fma(2.0, x2, rho2);       // r^2 + 2 x^2
  const double s3y = fma(2.0, y2, rho2);
  if (WITH_JAC) {
    const double dq = t + k2r2;
    const double q = dq + dq;                  // 2 d(rad)/d(r^2)
    Jm[0][0] = fma(3.0 * tp2, x, fma(tp1, y, fma(q, x2, a)));
    Jm[1][0] = fma(q, xy, fma(tp1, x, tp2 * y));
    Jm[0][1] = Jm[1][0];
    Jm[1][1] = fma(tp2, x, fma(3.0 * tp1, y, fma(q, y2, a)));
    const double r4 = rho2 * rho2;
    const double txy = xy + xy;
    Jk[0][0] = x * rho2; Jk[0][1] = x * r4; Jk[0][2] = txy; Jk[0][3] = s3x;
    Jk[1][0] = y * rho2; Jk[1][1] = y * r4; Jk[1][2] = s3y; Jk[1][3] = txy;
  }
  mx = fma(x, a, fma(tp1, xy, p2 * s3x));
  my = fma(y, a, fma(tp2, xy, p1 * s3y));
}

// equidistant (Kannala-Brandt).  The point Jacobian has no guard at r = 0 (NaN there, as in the reference: Q5).
template <bool WITH_JAC>
__device__ __forceinline__ void equidistant(const double* __restrict__ k, double& mx, double& my, double Jm[2][2], double Jk[2][4]) {
  const double x = mx, y = my;
  const double r2 = x * x + y * y;
  const double r = sqrt(r2);
  const double th = atan(r);
  const double th2 = th * th, th4 = th2 * th2, th6 = th4 * th2, th8 = th4 * th4;
  const double poly = k[0] * th2 + k[1] * th4 + k[2] * th6 + k[3] * th8 + 1.0;
  if (WITH_JAC) {
    const double inv_r = 1.0 / r;
    const double th_r = th * inv_r;
    const double r2p1 = r2 + 1.0;
    const double th3 = th2 * th, th5 = th4 * th, th7 = th6 * th;
    // d(poly)/d(v) / v  (v = x or y)
    const double dpoly = ((k[1] * th3 * 4.0 + k[2] * th5 * 6.0 + k[3] * th7 * 8.0 + k[0] * th * 2.0) * inv_r) / r2p1;
    const double a = poly / (r2 * r2p1);         // coefficient of v*w
    const double b = th_r / r2 * poly;           // th / r^3 * poly
    const double common = th_r * dpoly + a - b;  // multiplies x*x, x*y, y*y
    Jm[0][0] = th_r * poly + x * x * common;
    Jm[0][1] = x * y * common;
    Jm[1][0] = Jm[0][1];
    Jm[1][1] = th_r * poly + y * y * common;
    const double pw3 = th3 * inv_r, pw5 = th5 * inv_r, pw7 = th7 * inv_r, pw9 = th8 * th * inv_r;
    Jk[0][0] = x * pw3; Jk[0][1] = x * pw5; Jk[0][2] = x * pw7; Jk[0][3] = x * pw9;
    Jk[1][0] = y * pw3; Jk[1][1] = y * pw5; Jk[1][2] = y * pw7; Jk[1][3] = y * pw9;
  }
  const double s = (r > 1e-8) ? (th * poly) / r : 1.0;
  mx = x * s;
  my = y * s;
}

// field-of-view distortion (Devernay-Faugeras), one parameter w.  Same three regimes as the reference: w^2 < 1e-5 (identity),
// r^2 < 1e-5 (constant factor 2 tan(w/2) / w), general.  CAM/FovDistortion.hpp:23-92 (value + point Jacobian), :133-170 (d/dw)
template <bool WITH_JAC>
__device__ __forceinline__ void fov_distortion(const double* __restrict__ k, double& mx, double& my, double Jm[2][2], double Jk[2][4]) {
  const double w = k[0];
  const double u = mx, v = my;
  const double r_u = sqrt(u * u + v * v);
  const double r2 = r_u * r_u;
  const double tanwhalf = tan(0.5 * w);
  const double tanwhalfsq = tanwhalf * tanwhalf;
  const double atan_wrd = atan(2.0 * tanwhalf * r_u);
  const bool w_small = w * w < 1e-5, r_small = r2 < 1e-5;
  double r_rd;
  if (w_small) r_rd = 1.0;
  else if (r_small) r_rd = 2.0 * tanwhalf / w;
  else r_rd = atan_wrd / (r_u * w);
  if (WITH_JAC) {
    if (w_small) {
      Jm[0][0] = 1.0; Jm[0][1] = 0.0; Jm[1][0] = 0.0; Jm[1][1] = 1.0;
      Jk[0][0] = 0.0; Jk[1][0] = 0.0;
    } else if (r_small) {
      const double f = 2.0 * tanwhalf / w;
      Jm[0][0] = f; Jm[0][1] = 0.0; Jm[1][0] = 0.0; Jm[1][1] = f;
      const double ch = cos(0.5 * w);
      const double g = (w - sin(w)) / (w * w * ch * ch);
      Jk[0][0] = g; Jk[1][0] = g;
    } else {
      const double uv2 = u * u + v * v;
      const double a = atan_wrd / (w * r_u);                                  // atan / (w r)
      const double b = atan_wrd / (w * (r_u * r_u * r_u));                    // atan / (w r^3)
      const double c = 2.0 * tanwhalf / (w * uv2 * (4.0 * tanwhalfsq * uv2 + 1.0));
      Jm[0][0] = a - u * u * b + u * u * c;
      Jm[0][1] = u * v * c - u * v * b;
      Jm[1][0] = Jm[0][1];
      Jm[1][1] = a - v * v * b + v * v * c;
      const double d = 2.0 * (0.5 * tanwhalfsq + 0.5) / (w * (4.0 * tanwhalfsq * r2 + 1.0));
      const double e = atan_wrd / (w * w * r_u);
      Jk[0][0] = u * d - u * e;
      Jk[1][0] = v * d - v * e;
    }
  }
  mx = u * r_rd;
  my = v * r_rd;
}

// ---- projections ----------------------------------------------------------------------------------------
// NEG = true returns the NEGATED Jacobians (-Jp, -Ji, -Jd): what the error term e = y - y_hat needs, with the sign folded
// into the focal-length factors instead of a separate pass over the rows.
template <int MODEL, bool WITH_JAC, bool NEG = false>
struct Camera;

// pinhole + (radtan | equi | fov): params fu,fv,cu,cv,d0..
template <int MODEL, bool WITH_JAC, bool NEG>
struct PinholeCamera {
  static constexpr int P = 4, D = model_D(MODEL);
  __device__ __forceinline__ static void eval(const double* __restrict__ prm, const double p[3], Linearisation<P, D>& L) {
    const double fu = prm[0], fv = prm[1], cu = prm[2], cv = prm[3];
    const double rz = 1.0 / p[2];
    double mx = p[0] * rz, my = p[1] * rz;
    double Jm[2][2], Jk[2][4];
    if (MODEL == PINHOLE_RADTAN)
      radtan<WITH_JAC>(prm + 4, mx, my, Jm, Jk);
    else if (MODEL == PINHOLE_EQUI)
      equidistant<WITH_JAC>(prm + 4, mx, my, Jm, Jk);
    else
      fov_distortion<WITH_JAC>(prm + 4, mx, my, Jm, Jk);
    if (WITH_JAC) {
      const double ju = NEG ? -fu : fu, jv = NEG ? -fv : fv;
      const double one = NEG ? -1.0 : 1.0;
      const double jur = ju * rz, jvr = jv * rz;
      const double ux = p[0] * rz, uy = p[1] * rz;  // the undistorted normalised point
      L.Jp[0][0] = jur * Jm[0][0];
      L.Jp[0][1] = jur * Jm[0][1];
      L.Jp[0][2] = fma(-uy, L.Jp[0][1], -ux * L.Jp[0][0]);  // -ju (x Jm00 + y Jm01) / z^2
      L.Jp[1][0] = jvr * Jm[1][0];
      L.Jp[1][1] = jvr * Jm[1][1];
      L.Jp[1][2] = fma(-uy, L.Jp[1][1], -ux * L.Jp[1][0]);
      L.Ji[0][0] = NEG ? neg_int(mx) : mx; L.Ji[0][1] = 0.0; L.Ji[0][2] = one; L.Ji[0][3] = 0.0;
      L.Ji[1][0] = 0.0; L.Ji[1][1] = NEG ? neg_int(my) : my; L.Ji[1][2] = 0.0; L.Ji[1][3] = one;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        L.Jd[0][j] = ju * Jk[0][j];
        L.Jd[1][j] = jv * Jk[1][j];
      }
    }
    L.y[0] = fu * mx + cu;
    L.y[1] = fv * my + cv;
    L.valid = true;  // the reference's pinhole never bails out before writing y_hat
  }
};
template <bool WITH_JAC, bool NEG>
struct Camera<PINHOLE_RADTAN, WITH_JAC, NEG> : PinholeCamera<PINHOLE_RADTAN, WITH_JAC, NEG> {};
template <bool WITH_JAC, bool NEG>
struct Camera<PINHOLE_EQUI, WITH_JAC, NEG> : PinholeCamera<PINHOLE_EQUI, WITH_JAC, NEG> {};
template <bool WITH_JAC, bool NEG>
struct Camera<PINHOLE_FOV, WITH_JAC, NEG> : PinholeCamera<PINHOLE_FOV, WITH_JAC, NEG> {};

// omni + (radtan | none): params xi,fu,fv,cu,cv[,k1,k2,p1,p2]
template <int MODEL, bool WITH_JAC, bool NEG>
struct OmniCamera {
  static constexpr int P = 5, D = model_D(MODEL);
  __device__ __forceinline__ static void eval(const double* __restrict__ prm, const double p[3], Linearisation<P, D>& L) {
    const double xi = prm[0], fu = prm[1], fv = prm[2], cu = prm[3], cv = prm[4];
    const double d = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    const double fov = (xi <= 1.0) ? xi : 1.0 / xi;
    L.valid = !(p[2] <= -(fov * d));
    const double rz = 1.0 / (p[2] + xi * d);
    double mx = p[0] * rz, my = p[1] * rz;
    double Jn[2][3];
    if (WITH_JAC) {
      const double s = rz * rz / d;
      Jn[0][0] = s * (d * p[2] + xi * (p[1] * p[1] + p[2] * p[2]));
      Jn[1][0] = -s * xi * p[0] * p[1];
      Jn[0][1] = Jn[1][0];
      Jn[1][1] = s * (d * p[2] + xi * (p[0] * p[0] + p[2] * p[2]));
      const double s2 = s * (-xi * p[2] - d);
      Jn[0][2] = p[0] * s2;
      Jn[1][2] = p[1] * s2;
    }
    const double jxi0 = -mx * d * rz, jxi1 = -my * d * rz;  // d(m)/d(xi) at the undistorted point
    double Jm[2][2] = {{1.0, 0.0}, {0.0, 1.0}}, Jk[2][4];
    if (MODEL == OMNI_RADTAN) radtan<WITH_JAC>(prm + 5, mx, my, Jm, Jk);
    if (WITH_JAC) {
      const double ju = NEG ? -fu : fu, jv = NEG ? -fv : fv;
      const double one = NEG ? -1.0 : 1.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        L.Jp[0][c] = ju * (Jn[0][c] * Jm[0][0] + Jn[1][c] * Jm[0][1]);
        L.Jp[1][c] = jv * (Jn[0][c] * Jm[1][0] + Jn[1][c] * Jm[1][1]);
      }
      L.Ji[0][0] = ju * Jm[0][0] * jxi0 + ju * Jm[0][1] * jxi1;
      L.Ji[1][0] = jv * Jm[1][0] * jxi0 + jv * Jm[1][1] * jxi1;
      L.Ji[0][1] = NEG ? neg_int(mx) : mx; L.Ji[0][2] = 0.0; L.Ji[0][3] = one; L.Ji[0][4] = 0.0;
      L.Ji[1][1] = 0.0; L.Ji[1][2] = NEG ? neg_int(my) : my; L.Ji[1][3] = 0.0; L.Ji[1][4] = one;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        L.Jd[0][j] = ju * Jk[0][j];
        L.Jd[1][j] = jv * Jk[1][j];
      }
    }
    L.y[0] = fu * mx + cu;
    L.y[1] = fv * my + cv;
  }
};
template <bool WITH_JAC, bool NEG>
struct Camera<OMNI_RADTAN, WITH_JAC, NEG> : OmniCamera<OMNI_RADTAN, WITH_JAC, NEG> {};
template <bool WITH_JAC, bool NEG>
struct Camera<OMNI_NONE, WITH_JAC, NEG> : OmniCamera<OMNI_NONE, WITH_JAC, NEG> {};

// EUCM (no distortion): params alpha,beta,fu,fv,cu,cv
template <bool WITH_JAC, bool NEG>
struct Camera<EUCM_NONE, WITH_JAC, NEG> {
  static constexpr int P = 6, D = 0;
  __device__ __forceinline__ static void eval(const double* __restrict__ prm, const double p[3], Linearisation<P, D>& L) {
    const double al = prm[0], be = prm[1], fu = prm[2], fv = prm[3], cu = prm[4], cv = prm[5];
    const double x = p[0], y = p[1], z = p[2];
    const double r2 = x * x + y * y;
    const double d = sqrt(be * r2 + z * z);
    const double fov = (al <= 0.5) ? al / (1.0 - al) : (1.0 - al) / al;
    L.valid = !(z <= -(fov * d));
    const double norm = al * d + (1.0 - al) * z;
    const double ni = 1.0 / norm;
    if (WITH_JAC) {
      const double ju = NEG ? -fu : fu, jv = NEG ? -fv : fv;
      const double one = NEG ? -1.0 : 1.0;
      const double di = 1.0 / d;
      const double denom = ni * ni * di;
      const double mid = -(al * be * x * y) * denom;
      const double add = norm * d;
      const double addz = al * z + (1.0 - al) * d;
      L.Jp[0][0] = ju * (add - x * x * al * be) * denom;
      L.Jp[1][0] = jv * mid;
      L.Jp[0][1] = ju * mid;
      L.Jp[1][1] = jv * (add - y * y * al * be) * denom;
      L.Jp[0][2] = -ju * x * addz * denom;
      L.Jp[1][2] = -jv * y * addz * denom;
      const double ni2 = ni * ni;
      const double tx = -ju * x * ni2;
      const double ty = -ju * y * ni2;  // Q4: the reference scales row 1 with fu as well
      const double t4 = d - z;
      const double t5 = 0.5 * al * r2 * di;
      L.Ji[0][0] = tx * t4; L.Ji[1][0] = ty * t4;
      L.Ji[0][1] = tx * t5; L.Ji[1][1] = ty * t5;
      L.Ji[0][2] = (NEG ? -x : x) * ni; L.Ji[0][3] = 0.0; L.Ji[0][4] = one; L.Ji[0][5] = 0.0;
      L.Ji[1][2] = 0.0; L.Ji[1][3] = (NEG ? -y : y) * ni; L.Ji[1][4] = 0.0; L.Ji[1][5] = one;
    }
    L.y[0] = fu * (x * ni) + cu;
    L.y[1] = fv * (y * ni) + cv;
  }
};

// double sphere (no distortion): params xi,alpha,fu,fv,cu,cv
template <bool WITH_JAC, bool NEG>
struct Camera<DS_NONE, WITH_JAC, NEG> {
  static constexpr int P = 6, D = 0;
  __device__ __forceinline__ static void eval(const double* __restrict__ prm, const double p[3], Linearisation<P, D>& L) {
    const double xi = prm[0], al = prm[1], fu = prm[2], fv = prm[3], cu = prm[4], cv = prm[5];
    const double x = p[0], y = p[1], z = p[2];
    const double xx = x * x, yy = y * y;
    const double r2 = xx + yy;
    const double d1 = sqrt(r2 + z * z);
    const double t = (al <= 0.5) ? al / (1.0 - al) : (1.0 - al) / al;
    const double fov = (t + xi) / sqrt(2.0 * t * xi + xi * xi + 1.0);
    L.valid = !(z <= -(fov * d1));
    const double k = xi * d1 + z;
    const double d2 = sqrt(r2 + k * k);
    const double norm = al * d2 + (1.0 - al) * k;
    const double ni = 1.0 / norm;
    if (WITH_JAC) {
      const double ju = NEG ? -fu : fu, jv = NEG ? -fv : fv;
      const double one = NEG ? -1.0 : 1.0;
      const double d1i = 1.0 / d1, d2i = 1.0 / d2;
      const double ni2 = ni * ni;
      const double xy = x * y;
      const double tt2 = xi * z * d1i + 1.0;
      const double dn = (xi * (1.0 - al) * d1i + al * (xi * k * d1i + 1.0) * d2i) * ni2;
      const double tmp2 = ((1.0 - al) * tt2 + al * k * tt2 * d2i) * ni2;
      L.Jp[0][0] = ju * (ni - xx * dn);
      L.Jp[1][0] = -jv * xy * dn;
      L.Jp[0][1] = -ju * xy * dn;
      L.Jp[1][1] = jv * (ni - yy * dn);
      L.Jp[0][2] = -ju * x * tmp2;
      L.Jp[1][2] = -jv * y * tmp2;
      const double t4 = (al - 1.0 - al * k * d2i) * d1 * ni2;
      const double t5 = (k - d2) * ni2;
      L.Ji[0][0] = ju * x * t4; L.Ji[1][0] = jv * y * t4;
      L.Ji[0][1] = ju * x * t5; L.Ji[1][1] = jv * y * t5;
      L.Ji[0][2] = (NEG ? -x : x) * ni; L.Ji[0][3] = 0.0; L.Ji[0][4] = one; L.Ji[0][5] = 0.0;
      L.Ji[1][2] = 0.0; L.Ji[1][3] = (NEG ? -y : y) * ni; L.Ji[1][4] = 0.0; L.Ji[1][5] = one;
    }
    L.y[0] = fu * (x * ni) + cu;
    L.y[1] = fv * (y * ni) + cv;
  }
};

}  // namespace kb
