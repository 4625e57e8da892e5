// ORACLE — TEST INFRASTRUCTURE ONLY (see ko_math.hpp header: parity PINNED against the reference's own compiled code for the per-term part, the optimiser loop and both solver regimes, UNPINNED for the CHOLMOD factorisation itself).
//
// Camera models: CPU restatement of aslam_cv/aslam_cameras (CAM below =
// aslam_cv/aslam_cameras/include/aslam/cameras/implementation).  Only what the batch reprojection path
// uses: forward projection of a homogeneous point, its 2x4 point Jacobian, the intrinsics Jacobian, the
// distortion-parameter Jacobian, and update/get/set of the parameters.
#pragma once
#include <memory>

#include "ko_math.hpp"

namespace ko {

// ---------------------------------------------------------------------------------------------------
struct Distortion {
  virtual ~Distortion() {}
  virtual int dims() const = 0;
  virtual void distort(double y[2]) const = 0;
  virtual void distort(double y[2], Mat& J) const = 0;               // J: 2x2 d(distorted)/d(y)
  virtual void distortParameterJacobian(const double y[2], Mat& J) const = 0;  // 2 x dims
  virtual void update(const double* v) = 0;
  virtual void getParameters(std::vector<double>& p) const = 0;
  virtual void setParameters(const std::vector<double>& p) = 0;
};

// CAM/RadialTangentialDistortion.hpp:5-65, 153-182; aslam_cv/aslam_cameras/src/RadialTangentialDistortion.cpp:38-43
struct RadialTangentialDistortion : Distortion {
  double k1 = 0, k2 = 0, p1 = 0, p2 = 0;
  int dims() const override { return 4; }
  void distort(double y[2]) const override {
    const double x2 = y[0] * y[0], y2 = y[1] * y[1], xy = y[0] * y[1];
    const double rho2 = x2 + y2;
    const double rad = k1 * rho2 + k2 * rho2 * rho2;
    const double dx = y[0] * rad + 2.0 * p1 * xy + p2 * (rho2 + 2.0 * x2);
    const double dy = y[1] * rad + 2.0 * p2 * xy + p1 * (rho2 + 2.0 * y2);
    y[0] += dx;
    y[1] += dy;
  }
  void distort(double y[2], Mat& J) const override {
    J = Mat(2, 2);
    const double x2 = y[0] * y[0], y2 = y[1] * y[1];
    const double rho2 = x2 + y2;
    const double rad = k1 * rho2 + k2 * rho2 * rho2;
    J(0, 0) = 1 + rad + k1 * 2.0 * x2 + k2 * rho2 * 4 * x2 + 2.0 * p1 * y[1] + 6 * p2 * y[0];
    J(1, 0) = k1 * 2.0 * y[0] * y[1] + k2 * 4 * rho2 * y[0] * y[1] + p1 * 2.0 * y[0] + 2.0 * p2 * y[1];
    J(0, 1) = J(1, 0);
    J(1, 1) = 1 + rad + k1 * 2.0 * y2 + k2 * rho2 * 4 * y2 + 6 * p1 * y[1] + 2.0 * p2 * y[0];
    distort(y);
  }
  void distortParameterJacobian(const double y[2], Mat& J) const override {
    J = Mat(2, 4);
    const double y0 = y[0], y1 = y[1];
    const double r2 = y0 * y0 + y1 * y1, r4 = r2 * r2;
    J(0, 0) = y0 * r2; J(0, 1) = y0 * r4; J(0, 2) = 2.0 * y0 * y1;      J(0, 3) = r2 + 2.0 * y0 * y0;
    J(1, 0) = y1 * r2; J(1, 1) = y1 * r4; J(1, 2) = r2 + 2.0 * y1 * y1; J(1, 3) = 2.0 * y0 * y1;
  }
  void update(const double* v) override { k1 += v[0]; k2 += v[1]; p1 += v[2]; p2 += v[3]; }
  void getParameters(std::vector<double>& p) const override { p = {k1, k2, p1, p2}; }
  void setParameters(const std::vector<double>& p) override { k1 = p[0]; k2 = p[1]; p1 = p[2]; p2 = p[3]; }
};

// CAM/EquidistantDistortion.hpp:5-29 (value), :32-183 (Jacobian: symbolic output, unguarded at r = 0 — quirk Q5),
// :244-273 (parameter Jacobian); aslam_cv/aslam_cameras/src/EquidistantDistortion.cpp:35-40
struct EquidistantDistortion : Distortion {
  double k1 = 0, k2 = 0, k3 = 0, k4 = 0;
  int dims() const override { return 4; }
  void distort(double y[2]) const override {
    const double r = std::sqrt(y[0] * y[0] + y[1] * y[1]);
    const double th = std::atan(r);
    const double th2 = th * th, th4 = th2 * th2, th6 = th4 * th2, th8 = th4 * th4;
    const double thd = th * (1 + k1 * th2 + k2 * th4 + k3 * th6 + k4 * th8);
    const double s = (r > 1e-8) ? thd / r : 1.0;
    y[0] *= s;
    y[1] *= s;
  }
  void distort(double y[2], Mat& J) const override {
    // The reference evaluates one long generated expression per entry; written here with named
    // sub-terms in the same operation order class (sum of the same four groups).  No guard at r = 0:
    // the entries are NaN there exactly as in the reference.
    J = Mat(2, 2);
    const double x = y[0], w = y[1];
    const double r2 = x * x + w * w;
    const double r = std::sqrt(r2);
    const double th = std::atan(r);
    const double poly = k1 * std::pow(th, 2.0) + k2 * std::pow(th, 4.0) + k3 * std::pow(th, 6.0) + k4 * std::pow(th, 8.0) + 1.0;
    const double r2p1 = r2 + 1.0;
    const double r32 = std::pow(r2, 3.0 / 2.0);
    // d(poly)/d(x) and d(poly)/d(w)
    auto dpoly = [&](double v) {
      return (k2 * v * std::pow(th, 3.0) * 1.0 / r * 4.0) / r2p1 + (k3 * v * std::pow(th, 5.0) * 1.0 / r * 6.0) / r2p1 +
             (k4 * v * std::pow(th, 7.0) * 1.0 / r * 8.0) / r2p1 + (k1 * v * th * 1.0 / r * 2.0) / r2p1;
    };
    const double dpx = dpoly(x), dpw = dpoly(w);
    J(0, 0) = th * 1.0 / r * poly + x * th * 1.0 / r * dpx + ((x * x) * poly) / (r2 * r2p1) - (x * x) * th * 1.0 / r32 * poly;
    J(0, 1) = x * th * 1.0 / r * dpw + (x * w * poly) / (r2 * r2p1) - x * w * th * 1.0 / r32 * poly;
    J(1, 0) = w * th * 1.0 / r * dpx + (x * w * poly) / (r2 * r2p1) - x * w * th * 1.0 / r32 * poly;
    J(1, 1) = th * 1.0 / r * poly + w * th * 1.0 / r * dpw + ((w * w) * poly) / (r2 * r2p1) - (w * w) * th * 1.0 / r32 * poly;
    distort(y);
  }
  void distortParameterJacobian(const double y[2], Mat& J) const override {
    J = Mat(2, 4);
    const double r = std::sqrt(y[0] * y[0] + y[1] * y[1]);
    const double th = std::atan(r);
    const double pw[4] = {std::pow(th, 3.0), std::pow(th, 5.0), std::pow(th, 7.0), std::pow(th, 9.0)};
    for (int j = 0; j < 4; ++j) {
      J(0, j) = y[0] * pw[j] * 1.0 / r;
      J(1, j) = y[1] * pw[j] * 1.0 / r;
    }
  }
  void update(const double* v) override { k1 += v[0]; k2 += v[1]; k3 += v[2]; k4 += v[3]; }
  void getParameters(std::vector<double>& p) const override { p = {k1, k2, k3, k4}; }
  void setParameters(const std::vector<double>& p) override { k1 = p[0]; k2 = p[1]; k3 = p[2]; k4 = p[3]; }
};

// CAM/NoDistortion.hpp:10-55; aslam_cv/aslam_cameras/src/NoDistortion.cpp:17-19 (0 dimensions, still a DV: quirk Q7)
struct NoDistortion : Distortion {
  int dims() const override { return 0; }
  void distort(double*) const override {}
  void distort(double*, Mat& J) const override { J = Mat::Identity(2); }
  void distortParameterJacobian(const double*, Mat& J) const override { J = Mat(2, 0); }
  void update(const double*) override {}
  void getParameters(std::vector<double>& p) const override { p.clear(); }
  void setParameters(const std::vector<double>&) override {}
};

// CAM/FovDistortion.hpp:23-92 (value + point Jacobian, three regimes), :133-170 (parameter Jacobian);
// aslam_cv/aslam_cameras/src/FovDistortion.cpp:24-27 (update), :52-54 (test value w = 1)
struct FovDistortion : Distortion {
  double w = 1.0;
  int dims() const override { return 1; }
  void distort(double y[2]) const override {
    Mat J;
    distort(y, J);
  }
  void distort(double y[2], Mat& J) const override {
    J = Mat(2, 2);
    const double r_u = std::sqrt(y[0] * y[0] + y[1] * y[1]);
    const double r_u_cubed = r_u * r_u * r_u;
    const double tanwhalf = std::tan(w / 2.);
    const double tanwhalfsq = tanwhalf * tanwhalf;
    const double atan_wrd = std::atan(2. * tanwhalf * r_u);
    double r_rd;
    if (w * w < 1e-5) {
      r_rd = 1.0;
    } else {
      if (r_u * r_u < 1e-5) r_rd = 2. * tanwhalf / w;
      else r_rd = atan_wrd / (r_u * w);
    }
    const double u = y[0], v = y[1];
    if (w * w < 1e-5) {
      J = Mat::Identity(2);
    } else if (r_u * r_u < 1e-5) {
      J = Mat::Identity(2);
      J(0, 0) *= (2. * tanwhalf / w);
      J(1, 1) *= (2. * tanwhalf / w);
    } else {
      const double duf_du = (atan_wrd) / (w * r_u) - (u * u * atan_wrd) / (w * r_u_cubed) +
                            (2 * u * u * tanwhalf) / (w * (u * u + v * v) * (4 * tanwhalfsq * (u * u + v * v) + 1));
      const double duf_dv = (2 * u * v * tanwhalf) / (w * (u * u + v * v) * (4 * tanwhalfsq * (u * u + v * v) + 1)) -
                            (u * v * atan_wrd) / (w * r_u_cubed);
      const double dvf_du = duf_dv;
      const double dvf_dv = (atan_wrd) / (w * r_u) - (v * v * atan_wrd) / (w * r_u_cubed) +
                            (2 * v * v * tanwhalf) / (w * (u * u + v * v) * (4 * tanwhalfsq * (u * u + v * v) + 1));
      J(0, 0) = duf_du; J(0, 1) = duf_dv;
      J(1, 0) = dvf_du; J(1, 1) = dvf_dv;
    }
    y[0] *= r_rd;
    y[1] *= r_rd;
  }
  void distortParameterJacobian(const double y[2], Mat& J) const override {
    J = Mat(2, 1);
    const double tanwhalf = std::tan(w / 2.);
    const double tanwhalfsq = tanwhalf * tanwhalf;
    const double r_u = std::sqrt(y[0] * y[0] + y[1] * y[1]);
    const double atan_wrd = std::atan(2. * tanwhalf * r_u);
    const double u = y[0], v = y[1];
    if (w * w < 1e-5) {
      J(0, 0) = 0.0;
      J(1, 0) = 0.0;
    } else if (r_u * r_u < 1e-5) {
      const double g = (w - std::sin(w)) / (w * w * std::cos(w / 2) * std::cos(w / 2));
      J(0, 0) = g;
      J(1, 0) = g;
    } else {
      J(0, 0) = (2 * u * (tanwhalfsq / 2 + 0.5)) / (w * (4 * tanwhalfsq * r_u * r_u + 1)) - (u * atan_wrd) / (w * w * r_u);
      J(1, 0) = (2 * v * (tanwhalfsq / 2 + 0.5)) / (w * (4 * tanwhalfsq * r_u * r_u + 1)) - (v * atan_wrd) / (w * w * r_u);
    }
  }
  void update(const double* v) override { w += v[0]; }
  void getParameters(std::vector<double>& p) const override { p = {w}; }
  void setParameters(const std::vector<double>& p) override { w = p[0]; }
};

// ---------------------------------------------------------------------------------------------------
// CameraGeometry<Projection<Distortion>, GlobalShutter, NoMask> — CAM/CameraGeometry.hpp:232-247 forwards to the projection.
struct Projection {
  std::unique_ptr<Distortion> distortion;
  virtual ~Projection() {}
  virtual int dims() const = 0;
  // returns the reference's validity bool; the caller (ReprojectionError) ignores it (quirk Q6)
  virtual bool euclideanToKeypoint(const double p[3], double y[2]) const = 0;
  virtual bool euclideanToKeypoint(const double p[3], double y[2], Mat& J /*2x3 block of the 2x4*/) const = 0;
  virtual void euclideanToKeypointIntrinsicsJacobian(const double p[3], Mat& J) const = 0;
  virtual void euclideanToKeypointDistortionJacobian(const double p[3], Mat& J) const = 0;
  virtual bool jacobianIgnoresHomogeneousSign() const { return false; }
  virtual void update(const double* v) = 0;
  virtual void getParameters(std::vector<double>& p) const = 0;
  virtual void setParameters(const std::vector<double>& p) = 0;

  // homogeneous wrappers: e.g. CAM/OmniProjection.hpp:185-229, 449-483
  bool homogeneousToKeypoint(const double ph[4], double y[2]) const {
    if (ph[3] < 0) {
      const double n[3] = {-ph[0], -ph[1], -ph[2]};
      return euclideanToKeypoint(n, y);
    }
    return euclideanToKeypoint(ph, y);
  }
  bool homogeneousToKeypoint(const double ph[4], double y[2], Mat& J) const {
    J = Mat(2, 4);
    Mat J3;
    bool ok;
    if (ph[3] < 0 && !jacobianIgnoresHomogeneousSign()) {
      const double n[3] = {-ph[0], -ph[1], -ph[2]};
      ok = euclideanToKeypoint(n, y, J3);
      J3 = -J3;
    } else {
      ok = euclideanToKeypoint(ph, y, J3);
    }
    for (int i = 0; i < J3.r; ++i)
      for (int j = 0; j < J3.c; ++j) J(i, j) = J3(i, j);
    return ok;
  }
  void homogeneousToKeypointIntrinsicsJacobian(const double ph[4], Mat& J) const {
    if (ph[3] < 0.0) {
      const double n[3] = {-ph[0], -ph[1], -ph[2]};
      euclideanToKeypointIntrinsicsJacobian(n, J);
    } else {
      euclideanToKeypointIntrinsicsJacobian(ph, J);
    }
  }
  void homogeneousToKeypointDistortionJacobian(const double ph[4], Mat& J) const {
    if (ph[3] < 0.0) {
      const double n[3] = {-ph[0], -ph[1], -ph[2]};
      euclideanToKeypointDistortionJacobian(n, J);
    } else {
      euclideanToKeypointDistortionJacobian(ph, J);
    }
  }
};

// CAM/PinholeProjection.hpp:73-145 (value, Jp), :165-198 (homogeneous; the Jacobian overload returns before
// looking at the sign of ph[3] — quirk Q3), :324-378 (Ji, Jd), :510-554 (update/get/set)
struct PinholeProjection : Projection {
  double fu = 0, fv = 0, cu = 0, cv = 0;
  int dims() const override { return 4; }
  bool jacobianIgnoresHomogeneousSign() const override { return true; }
  bool euclideanToKeypoint(const double p[3], double y[2]) const override {
    const double rz = 1.0 / p[2];
    y[0] = p[0] * rz;
    y[1] = p[1] * rz;
    distortion->distort(y);
    y[0] = fu * y[0] + cu;
    y[1] = fv * y[1] + cv;
    return p[2] > 0;
  }
  bool euclideanToKeypoint(const double p[3], double y[2], Mat& J) const override {
    J = Mat(2, 3);
    const double rz = 1.0 / p[2];
    const double rz2 = rz * rz;
    y[0] = p[0] * rz;
    y[1] = p[1] * rz;
    Mat Jd;
    distortion->distort(y, Jd);
    J(0, 0) = fu * Jd(0, 0) * rz;
    J(0, 1) = fu * Jd(0, 1) * rz;
    J(0, 2) = -fu * (p[0] * Jd(0, 0) + p[1] * Jd(0, 1)) * rz2;
    J(1, 0) = fv * Jd(1, 0) * rz;
    J(1, 1) = fv * Jd(1, 1) * rz;
    J(1, 2) = -fv * (p[0] * Jd(1, 0) + p[1] * Jd(1, 1)) * rz2;
    y[0] = fu * y[0] + cu;
    y[1] = fv * y[1] + cv;
    return p[2] > 0;
  }
  void euclideanToKeypointIntrinsicsJacobian(const double p[3], Mat& J) const override {
    J = Mat(2, 4);
    const double rz = 1.0 / p[2];
    double kp[2] = {p[0] * rz, p[1] * rz};
    distortion->distort(kp);
    J(0, 0) = kp[0]; J(0, 2) = 1;
    J(1, 1) = kp[1]; J(1, 3) = 1;
  }
  void euclideanToKeypointDistortionJacobian(const double p[3], Mat& J) const override {
    const double rz = 1.0 / p[2];
    const double kp[2] = {p[0] * rz, p[1] * rz};
    distortion->distortParameterJacobian(kp, J);
    for (int j = 0; j < J.c; ++j) { J(0, j) *= fu; J(1, j) *= fv; }
  }
  void update(const double* v) override { fu += v[0]; fv += v[1]; cu += v[2]; cv += v[3]; }
  void getParameters(std::vector<double>& p) const override { p = {fu, fv, cu, cv}; }
  void setParameters(const std::vector<double>& p) override { fu = p[0]; fv = p[1]; cu = p[2]; cv = p[3]; }
};

// CAM/OmniProjection.hpp:72-180 (value, Jp; validity cone :89-90, :143-144), :381-445 (Ji incl. d/dxi, Jd),
// :626-673 (temporaries, update/get/set)
struct OmniProjection : Projection {
  double xi = 0, fu = 0, fv = 0, cu = 0, cv = 0;
  double fov_parameter = 0;
  void updateTemporaries() { fov_parameter = (xi <= 1.0) ? xi : 1 / xi; }
  int dims() const override { return 5; }
  bool euclideanToKeypoint(const double p[3], double y[2]) const override {
    const double d = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    if (p[2] <= -(fov_parameter * d)) return false;  // y left untouched, as in the reference
    const double rz = 1.0 / (p[2] + xi * d);
    y[0] = p[0] * rz;
    y[1] = p[1] * rz;
    distortion->distort(y);
    y[0] = fu * y[0] + cu;
    y[1] = fv * y[1] + cv;
    return true;
  }
  bool euclideanToKeypoint(const double p[3], double y[2], Mat& J) const override {
    J = Mat(2, 3);
    const double d = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    if (p[2] <= -(fov_parameter * d)) return false;
    double rz = 1.0 / (p[2] + xi * d);
    y[0] = p[0] * rz;
    y[1] = p[1] * rz;
    rz = rz * rz / d;
    J(0, 0) = rz * (d * p[2] + xi * (p[1] * p[1] + p[2] * p[2]));
    J(1, 0) = -rz * xi * p[0] * p[1];
    J(0, 1) = J(1, 0);
    J(1, 1) = rz * (d * p[2] + xi * (p[0] * p[0] + p[2] * p[2]));
    rz = rz * (-xi * p[2] - d);
    J(0, 2) = p[0] * rz;
    J(1, 2) = p[1] * rz;
    Mat Jd;
    distortion->distort(y, Jd);
    for (int c = 0; c < 3; ++c) {
      const double a = fu * (J(0, c) * Jd(0, 0) + J(1, c) * Jd(0, 1));
      J(1, c) = fv * (J(0, c) * Jd(1, 0) + J(1, c) * Jd(1, 1));
      J(0, c) = a;
    }
    y[0] = fu * y[0] + cu;
    y[1] = fv * y[1] + cv;
    return true;
  }
  void euclideanToKeypointIntrinsicsJacobian(const double p[3], Mat& J) const override {
    J = Mat(2, 5);
    const double d = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    const double rz = 1.0 / (p[2] + xi * d);
    double kp[2] = {p[0] * rz, p[1] * rz};
    const double Jxi[2] = {-kp[0] * d * rz, -kp[1] * d * rz};
    Mat Jd;
    distortion->distort(kp, Jd);
    for (int j = 0; j < 2; ++j) { Jd(0, j) *= fu; Jd(1, j) *= fv; }
    J(0, 0) = Jd(0, 0) * Jxi[0] + Jd(0, 1) * Jxi[1];
    J(1, 0) = Jd(1, 0) * Jxi[0] + Jd(1, 1) * Jxi[1];
    J(0, 1) = kp[0]; J(0, 3) = 1;
    J(1, 2) = kp[1]; J(1, 4) = 1;
  }
  void euclideanToKeypointDistortionJacobian(const double p[3], Mat& J) const override {
    const double d = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    const double rz = 1.0 / (p[2] + xi * d);
    const double kp[2] = {p[0] * rz, p[1] * rz};
    distortion->distortParameterJacobian(kp, J);
    for (int j = 0; j < J.c; ++j) { J(0, j) *= fu; J(1, j) *= fv; }
  }
  void update(const double* v) override {
    xi += v[0]; fu += v[1]; fv += v[2]; cu += v[3]; cv += v[4];
    updateTemporaries();
  }
  void getParameters(std::vector<double>& p) const override { p = {xi, fu, fv, cu, cv}; }
  void setParameters(const std::vector<double>& p) override {
    xi = p[0]; fu = p[1]; fv = p[2]; cu = p[3]; cv = p[4];
    updateTemporaries();
  }
};

// CAM/ExtendedUnifiedProjection.hpp:89-198 (value, Jp; distortion object never applied :118-125),
// :397-455 (Ji; row 1 of the alpha/beta columns scales with fu, not fv — quirk Q4), :459-471 (Jd = 2x0), :650-699
struct ExtendedUnifiedProjection : Projection {
  double alpha = 0, beta = 0, fu = 0, fv = 0, cu = 0, cv = 0;
  double fov_parameter = 0;
  void updateTemporaries() { fov_parameter = (alpha <= 0.5) ? alpha / (1 - alpha) : (1 - alpha) / alpha; }
  int dims() const override { return 6; }
  bool euclideanToKeypoint(const double p[3], double y[2]) const override {
    const double x = p[0], w = p[1], z = p[2];
    const double d = std::sqrt(beta * (x * x + w * w) + z * z);
    if (z <= -(fov_parameter * d)) return false;
    const double norm = alpha * d + (1 - alpha) * z;
    const double norm_inv = 1.0 / norm;
    y[0] = fu * (x * norm_inv) + cu;
    y[1] = fv * (w * norm_inv) + cv;
    return true;
  }
  bool euclideanToKeypoint(const double p[3], double y[2], Mat& J) const override {
    J = Mat(2, 3);
    const double x = p[0], w = p[1], z = p[2];
    const double d = std::sqrt(beta * (x * x + w * w) + z * z);
    const double d_inv = 1.0 / d;
    if (z <= -(fov_parameter * d)) return false;
    const double norm = alpha * d + (1 - alpha) * z;
    const double norm_inv = 1.0 / norm;
    y[0] = fu * (x * norm_inv) + cu;
    y[1] = fv * (w * norm_inv) + cv;
    const double denom = norm_inv * norm_inv * d_inv;
    const double mid = -(alpha * beta * x * w) * denom;
    const double add = norm * d;
    const double addz = (alpha * z + (1 - alpha) * d);
    J(0, 0) = fu * (add - x * x * alpha * beta) * denom;
    J(1, 0) = fv * mid;
    J(0, 1) = fu * mid;
    J(1, 1) = fv * (add - w * w * alpha * beta) * denom;
    J(0, 2) = -fu * x * addz * denom;
    J(1, 2) = -fv * w * addz * denom;
    return true;
  }
  void euclideanToKeypointIntrinsicsJacobian(const double p[3], Mat& J) const override {
    J = Mat(2, 6);
    const double x = p[0], w = p[1], z = p[2];
    const double r2 = x * x + w * w;
    const double d = std::sqrt(beta * r2 + z * z);
    const double d_inv = 1.0 / d;
    const double norm = alpha * d + (1 - alpha) * z;
    const double norm_inv = 1.0 / norm;
    const double norm_inv2 = norm_inv * norm_inv;
    const double tmp_x = -fu * x * norm_inv2;
    const double tmp_y = -fu * w * norm_inv2;  // Q4: the reference uses fu here
    const double tmp4 = (d - z);
    const double tmp5 = 0.5 * alpha * r2 * d_inv;
    J(0, 0) = tmp_x * tmp4; J(1, 0) = tmp_y * tmp4;
    J(0, 1) = tmp_x * tmp5; J(1, 1) = tmp_y * tmp5;
    J(0, 2) = x * norm_inv; J(0, 4) = 1;
    J(1, 3) = w * norm_inv; J(1, 5) = 1;
  }
  void euclideanToKeypointDistortionJacobian(const double*, Mat& J) const override { J = Mat(2, 0); }
  void update(const double* v) override {
    alpha += v[0]; beta += v[1]; fu += v[2]; fv += v[3]; cu += v[4]; cv += v[5];
    updateTemporaries();
  }
  void getParameters(std::vector<double>& p) const override { p = {alpha, beta, fu, fv, cu, cv}; }
  void setParameters(const std::vector<double>& p) override {
    alpha = p[0]; beta = p[1]; fu = p[2]; fv = p[3]; cu = p[4]; cv = p[5];
    updateTemporaries();
  }
};

// CAM/DoubleSphereProjection.hpp:90-221 (value, Jp), :443-503 (Ji), :507-519 (Jd = 2x0), :698-748
struct DoubleSphereProjection : Projection {
  double xi = 0, alpha = 0, fu = 0, fv = 0, cu = 0, cv = 0;
  double fov_parameter = 0;
  void updateTemporaries() {
    const double temp = alpha <= 0.5 ? alpha / (1 - alpha) : (1 - alpha) / alpha;
    fov_parameter = (temp + xi) / std::sqrt(2 * temp * xi + xi * xi + 1);
  }
  int dims() const override { return 6; }
  bool euclideanToKeypoint(const double p[3], double y[2]) const override {
    const double x = p[0], w = p[1], z = p[2];
    const double r2 = x * x + w * w;
    const double d1 = std::sqrt(r2 + z * z);
    if (z <= -(fov_parameter * d1)) return false;
    const double k = xi * d1 + z;
    const double d2 = std::sqrt(r2 + k * k);
    const double norm = alpha * d2 + (1 - alpha) * k;
    const double norm_inv = 1.0 / norm;
    y[0] = fu * (x * norm_inv) + cu;
    y[1] = fv * (w * norm_inv) + cv;
    return true;
  }
  bool euclideanToKeypoint(const double p[3], double y[2], Mat& J) const override {
    J = Mat(2, 3);
    const double x = p[0], w = p[1], z = p[2];
    const double xx = x * x, yy = w * w;
    const double r2 = xx + yy;
    const double d1 = std::sqrt(r2 + z * z);
    const double d1_inv = 1.0 / d1;
    if (z <= -(fov_parameter * d1)) return false;
    const double k = xi * d1 + z;
    const double d2 = std::sqrt(r2 + k * k);
    const double d2_inv = 1.0 / d2;
    const double norm = alpha * d2 + (1 - alpha) * k;
    const double norm_inv = 1.0 / norm;
    const double norm_inv2 = norm_inv * norm_inv;
    y[0] = fu * (x * norm_inv) + cu;
    y[1] = fv * (w * norm_inv) + cv;
    const double xy = x * w;
    const double tt2 = xi * z * d1_inv + 1;
    const double d_norm_d_r2 = (xi * (1 - alpha) * d1_inv + alpha * (xi * k * d1_inv + 1) * d2_inv) * norm_inv2;
    const double tmp2 = ((1 - alpha) * tt2 + alpha * k * tt2 * d2_inv) * norm_inv2;
    J(0, 0) = fu * (norm_inv - xx * d_norm_d_r2);
    J(1, 0) = -fv * xy * d_norm_d_r2;
    J(0, 1) = -fu * xy * d_norm_d_r2;
    J(1, 1) = fv * (norm_inv - yy * d_norm_d_r2);
    J(0, 2) = -fu * x * tmp2;
    J(1, 2) = -fv * w * tmp2;
    return true;
  }
  void euclideanToKeypointIntrinsicsJacobian(const double p[3], Mat& J) const override {
    J = Mat(2, 6);
    const double x = p[0], w = p[1], z = p[2];
    const double r2 = x * x + w * w;
    const double d1 = std::sqrt(r2 + z * z);
    const double k = xi * d1 + z;
    const double d2 = std::sqrt(r2 + k * k);
    const double d2_inv = 1.0 / d2;
    const double norm = alpha * d2 + (1 - alpha) * k;
    const double norm_inv = 1.0 / norm;
    const double norm_inv2 = norm_inv * norm_inv;
    const double tmp4 = (alpha - 1 - alpha * k * d2_inv) * d1 * norm_inv2;
    const double tmp5 = (k - d2) * norm_inv2;
    J(0, 0) = fu * x * tmp4; J(1, 0) = fv * w * tmp4;
    J(0, 1) = fu * x * tmp5; J(1, 1) = fv * w * tmp5;
    J(0, 2) = x * norm_inv;  J(0, 4) = 1;
    J(1, 3) = w * norm_inv;  J(1, 5) = 1;
  }
  void euclideanToKeypointDistortionJacobian(const double*, Mat& J) const override { J = Mat(2, 0); }
  void update(const double* v) override {
    xi += v[0]; alpha += v[1]; fu += v[2]; fv += v[3]; cu += v[4]; cv += v[5];
    updateTemporaries();
  }
  void getParameters(std::vector<double>& p) const override { p = {xi, alpha, fu, fv, cu, cv}; }
  void setParameters(const std::vector<double>& p) override {
    xi = p[0]; alpha = p[1]; fu = p[2]; fv = p[3]; cu = p[4]; cv = p[5];
    updateTemporaries();
  }
};

// model ids follow include/kalibr_b200.h (kb_camera_model); K2/include/kalibr2/CameraCalibrator.hpp:421-441
inline std::unique_ptr<Projection> makeCamera(int model, const double* params /*P then D*/) {
  std::unique_ptr<Projection> proj;
  std::unique_ptr<Distortion> dist;
  switch (model) {
    case 0: proj.reset(new PinholeProjection()); dist.reset(new RadialTangentialDistortion()); break;
    case 1: proj.reset(new PinholeProjection()); dist.reset(new EquidistantDistortion()); break;
    case 2: proj.reset(new OmniProjection()); dist.reset(new RadialTangentialDistortion()); break;
    case 3: proj.reset(new ExtendedUnifiedProjection()); dist.reset(new NoDistortion()); break;
    case 4: proj.reset(new DoubleSphereProjection()); dist.reset(new NoDistortion()); break;
    case 5: proj.reset(new PinholeProjection()); dist.reset(new FovDistortion()); break;
    case 6: proj.reset(new OmniProjection()); dist.reset(new NoDistortion()); break;
    default: return nullptr;
  }
  const int P = proj->dims(), D = dist->dims();
  proj->setParameters(std::vector<double>(params, params + P));
  dist->setParameters(std::vector<double>(params + P, params + P + D));
  proj->distortion = std::move(dist);
  return proj;
}

}  // namespace ko
