// ORACLE — TEST INFRASTRUCTURE ONLY (see ko_math.hpp header: parity PINNED against the reference's own compiled code for the per-term part, the optimiser loop and both solver regimes, UNPINNED for the CHOLMOD factorisation itself).
//
// aslam_backend / aslam_backend_expressions / aslam_cv_error_terms restated on the CPU, object per term,
// heap-backed small matrices and std::map containers like the reference so that it is a fair CPU baseline.
// BE = aslam_optimizer/aslam_backend, BX = aslam_optimizer/aslam_backend_expressions,
// SBM = aslam_optimizer/sparse_block_matrix, CVE = aslam_cv/aslam_cv_error_terms, CVB = aslam_cv/aslam_cv_backend.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <thread>

#include "ko_cameras.hpp"

namespace ko {

// ---- BE/include/aslam/backend/DesignVariable.hpp:18-145 ------------------------------------------
struct DesignVariable {
  int blockIndex = -1;
  int columnBase = -1;
  bool active = false;
  double scaling = 1.0;
  virtual ~DesignVariable() {}
  virtual int minimalDimensions() const = 0;
  virtual void update(const double* dp, int size) = 0;
  virtual void revertUpdate() = 0;
  virtual void getParameters(std::vector<double>& p) const = 0;
};
struct BlockIndexOrdering {  // DesignVariable.hpp:25-31
  bool operator()(const DesignVariable* a, const DesignVariable* b) const { return a->blockIndex < b->blockIndex; }
};

// ---- BE/include/aslam/backend/JacobianContainer.hpp:119-135 --------------------------------------
struct JacobianContainer {
  typedef std::map<DesignVariable*, Mat, BlockIndexOrdering> map_t;
  int rows;
  map_t jacobianMap;
  explicit JacobianContainer(int r) : rows(r) {}
  void add(DesignVariable* dv, const Mat& J) {
    if (!dv->active) return;
    map_t::iterator it = jacobianMap.find(dv);
    if (it == jacobianMap.end())
      jacobianMap.insert(jacobianMap.end(), std::make_pair(dv, J));
    else
      it->second += J;
  }
};

// ---- BX/src/RotationQuaternion.cpp:10-56 ----------------------------------------------------------
struct RotationQuaternion : DesignVariable {
  double q[4], p_q[4];
  Mat C;
  explicit RotationQuaternion(const double* q_) {
    std::memcpy(q, q_, sizeof(q));
    std::memcpy(p_q, q_, sizeof(q));
    C = quat2r(q);
  }
  int minimalDimensions() const override { return 3; }
  void update(const double* dp, int) override {
    std::memcpy(p_q, q, sizeof(q));
    double nq[4];
    updateQuat(q, dp, nq);
    std::memcpy(q, nq, sizeof(q));
    C = quat2r(q);
  }
  void revertUpdate() override {
    std::memcpy(q, p_q, sizeof(q));
    C = quat2r(q);
  }
  void getParameters(std::vector<double>& p) const override { p.assign(q, q + 4); }
  void evaluateJacobians(JacobianContainer& out, const Mat& chain) const { out.add(const_cast<RotationQuaternion*>(this), chain); }
};

// ---- BX/src/EuclideanPoint.cpp:17-53 --------------------------------------------------------------
struct EuclideanPoint : DesignVariable {
  double p[3], p_p[3];
  explicit EuclideanPoint(const double* p_) {
    std::memcpy(p, p_, sizeof(p));
    std::memcpy(p_p, p_, sizeof(p));
  }
  int minimalDimensions() const override { return 3; }
  void update(const double* dp, int) override {
    std::memcpy(p_p, p, sizeof(p));
    for (int i = 0; i < 3; ++i) p[i] += dp[i];
  }
  void revertUpdate() override { std::memcpy(p, p_p, sizeof(p)); }
  void getParameters(std::vector<double>& out) const override { out.assign(p, p + 3); }
  void evaluateJacobians(JacobianContainer& out, const Mat& chain) const { out.add(const_cast<EuclideanPoint*>(this), chain); }
};

// ---- BE/include/aslam/backend/implementation/DesignVariableAdapter.hpp:22-55 over a projection or a distortion -----
struct ProjectionDv : DesignVariable {
  Projection* proj;
  std::vector<double> backup;
  explicit ProjectionDv(Projection* p) : proj(p) {}
  int minimalDimensions() const override { return proj->dims(); }
  void update(const double* dp, int) override {
    proj->getParameters(backup);
    proj->update(dp);
  }
  void revertUpdate() override { proj->setParameters(backup); }
  void getParameters(std::vector<double>& p) const override { proj->getParameters(p); }
};
struct DistortionDv : DesignVariable {
  Distortion* dist;
  std::vector<double> backup;
  explicit DistortionDv(Distortion* d) : dist(d) {}
  int minimalDimensions() const override { return dist->dims(); }
  void update(const double* dp, int) override {
    dist->getParameters(backup);
    dist->update(dp);
  }
  void revertUpdate() override { dist->setParameters(backup); }
  void getParameters(std::vector<double>& p) const override { dist->getParameters(p); }
};

// ---- CVB/include/aslam/backend/implementation/CameraDesignVariable.hpp:5-54 -----------------------
struct CameraDesignVariable {
  std::unique_ptr<Projection> camera;
  std::unique_ptr<ProjectionDv> projectionDv;
  std::unique_ptr<DistortionDv> distortionDv;
  // (the shutter DV is a 0-dim inactive GlobalShutter adapter: never enters the optimiser)
  explicit CameraDesignVariable(std::unique_ptr<Projection> cam) : camera(std::move(cam)) {
    projectionDv.reset(new ProjectionDv(camera.get()));
    distortionDv.reset(new DistortionDv(camera->distortion.get()));
  }
  void setActive(bool p, bool d) {
    projectionDv->active = p;
    distortionDv->active = d;
  }
  void evaluateJacobians(JacobianContainer& out, const double ph[4]) const {
    if (projectionDv->active) {
      Mat Jp;
      camera->homogeneousToKeypointIntrinsicsJacobian(ph, Jp);
      out.add(projectionDv.get(), -Jp);
    }
    if (distortionDv->active) {
      Mat Jd;
      camera->homogeneousToKeypointDistortionJacobian(ph, Jd);
      out.add(distortionDv.get(), -Jd);
    }
  }
};

// ---- BX transformation expression nodes -----------------------------------------------------------
// The reference's nodes cache fixed-size Eigen::Matrix4d members that worker threads of evaluateError overwrite
// concurrently with identical values (terms of one view share their transformation nodes).  The caches here are
// heap-backed, so they are overwritten element-wise in place: never re-allocated after construction.
inline void storeCache(Mat& dst, const Mat& src) {
  if (dst.r != src.r || dst.c != src.c) {
    dst = src;
    return;
  }
  for (size_t i = 0; i < src.d.size(); ++i) dst.d[i] = src.d[i];
}

struct TransformationExpressionNode {
  virtual ~TransformationExpressionNode() {}
  virtual Mat toTransformationMatrix() = 0;
  virtual void evaluateJacobians(JacobianContainer& out, const Mat& chain) const = 0;
};

// BX/src/TransformationBasic.cpp:19-67
struct TransformationBasic : TransformationExpressionNode {
  std::shared_ptr<RotationQuaternion> rotation;
  std::shared_ptr<EuclideanPoint> translation;
  TransformationBasic(std::shared_ptr<RotationQuaternion> r, std::shared_ptr<EuclideanPoint> t) : rotation(r), translation(t) {}
  Mat toTransformationMatrix() override {
    Mat T = Mat::Identity(4);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) T(i, j) = rotation->C(i, j);
      T(i, 3) = translation->p[i];
    }
    return T;
  }
  void evaluateJacobians(JacobianContainer& out, const Mat& chain) const override {
    const double* r = translation->p;
    Mat crRotation(6, 3);
    Mat mcx = -crossMx(r[0], r[1], r[2]);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) crRotation(i, j) = mcx(i, j);
    for (int i = 0; i < 3; ++i) crRotation(3 + i, i) = 1.0;
    rotation->evaluateJacobians(out, chain * crRotation);
    Mat crTranslation(6, 3);
    for (int i = 0; i < 3; ++i) crTranslation(i, i) = 1.0;
    translation->evaluateJacobians(out, chain * crTranslation);
  }
};

// BX/src/TransformationExpressionNode.cpp:40-72
struct TransformationExpressionNodeMultiply : TransformationExpressionNode {
  std::shared_ptr<TransformationExpressionNode> lhs, rhs;
  Mat T_lhs, T_rhs;
  TransformationExpressionNodeMultiply(std::shared_ptr<TransformationExpressionNode> l, std::shared_ptr<TransformationExpressionNode> r)
      : lhs(l), rhs(r) {
    T_lhs = lhs->toTransformationMatrix();
    T_rhs = rhs->toTransformationMatrix();
  }
  Mat toTransformationMatrix() override {
    const Mat l = lhs->toTransformationMatrix();
    const Mat r = rhs->toTransformationMatrix();
    storeCache(T_lhs, l);
    storeCache(T_rhs, r);
    return l * r;
  }
  void evaluateJacobians(JacobianContainer& out, const Mat& chain) const override {
    rhs->evaluateJacobians(out, chain * boxTimes(T_lhs));
    lhs->evaluateJacobians(out, chain);
  }
};

// BX/src/TransformationExpressionNode.cpp:79-101
struct TransformationExpressionNodeInverse : TransformationExpressionNode {
  std::shared_ptr<TransformationExpressionNode> dvTransformation;
  Mat T;
  explicit TransformationExpressionNodeInverse(std::shared_ptr<TransformationExpressionNode> t) : dvTransformation(t) {
    T = inverse4(dvTransformation->toTransformationMatrix());
  }
  Mat toTransformationMatrix() override {
    const Mat inv = inverse4(dvTransformation->toTransformationMatrix());
    storeCache(T, inv);
    return inv;
  }
  void evaluateJacobians(JacobianContainer& out, const Mat& chain) const override {
    dvTransformation->evaluateJacobians(out, chain * (-boxTimes(T)));
  }
};

// BX/src/HomogeneousExpressionNode.cpp:48-99 (Multiply over a Constant point)
struct HomogeneousExpressionNodeMultiply {
  std::shared_ptr<TransformationExpressionNode> lhs;
  double p_rhs[4];
  mutable Mat T_lhs;
  HomogeneousExpressionNodeMultiply(std::shared_ptr<TransformationExpressionNode> l, const double p[4]) : lhs(l) {
    std::memcpy(p_rhs, p, sizeof(p_rhs));
    T_lhs = lhs->toTransformationMatrix();
  }
  void toHomogeneous(double out[4]) const {
    const Mat Tl = lhs->toTransformationMatrix();
    storeCache(T_lhs, Tl);
    for (int i = 0; i < 4; ++i) {
      double s = 0;
      for (int j = 0; j < 4; ++j) s += Tl(i, j) * p_rhs[j];
      out[i] = s;
    }
  }
  void evaluateJacobians(JacobianContainer& out, const Mat& chain) const {
    double tp[4];
    for (int i = 0; i < 4; ++i) {
      double s = 0;
      for (int j = 0; j < 4; ++j) s += T_lhs(i, j) * p_rhs[j];
      tp[i] = s;
    }
    lhs->evaluateJacobians(out, chain * boxMinus(tp));
    // rhs is a HomogeneousExpressionNodeConstant: no design variables (HomogeneousExpressionNode.cpp:90-99)
  }
};

// ---- SBM/include/sparse_block_matrix/implementation/sparse_block_matrix.hpp:97-143 -----------------
struct SparseBlockMatrix {
  std::vector<int> blockIndices;  // partial sums of the DV dimensions (same for rows and cols)
  std::vector<std::map<int, Mat*>> blockCols;
  SparseBlockMatrix() {}
  explicit SparseBlockMatrix(const std::vector<int>& partial) : blockIndices(partial), blockCols(partial.size()) {}
  ~SparseBlockMatrix() { clear(true); }
  SparseBlockMatrix(const SparseBlockMatrix&) = delete;
  SparseBlockMatrix& operator=(const SparseBlockMatrix&) = delete;
  void reset(const std::vector<int>& partial) {
    clear(true);
    blockIndices = partial;
    blockCols.assign(partial.size(), std::map<int, Mat*>());
  }
  int bRows() const { return (int)blockIndices.size(); }
  int rows() const { return blockIndices.empty() ? 0 : blockIndices.back(); }
  int dimOfBlock(int i) const { return i ? blockIndices[i] - blockIndices[i - 1] : blockIndices[0]; }
  int baseOfBlock(int i) const { return i ? blockIndices[i - 1] : 0; }
  void clear(bool dealloc) {
    for (auto& col : blockCols) {
      for (auto& kv : col) {
        if (dealloc) delete kv.second; else kv.second->setZero();
      }
      if (dealloc) col.clear();
    }
  }
  Mat* block(int r, int c, bool alloc) {
    auto it = blockCols[c].find(r);
    if (it == blockCols[c].end()) {
      if (!alloc) return nullptr;
      Mat* b = new Mat(dimOfBlock(r), dimOfBlock(c));
      blockCols[c].insert(std::make_pair(r, b));
      return b;
    }
    return it->second;
  }
  const Mat* block(int r, int c) const {
    auto it = blockCols[c].find(r);
    return it == blockCols[c].end() ? nullptr : it->second;
  }
};

// ---- BE/src/JacobianContainer.cpp:103-167 ----------------------------------------------------------
inline void evaluateHessian(const JacobianContainer& jc, const double e[2], const Mat& sqrtInvR, SparseBlockMatrix& H,
                            std::vector<double>& rhs) {
  JacobianContainer::map_t mapCopy = jc.jacobianMap;
  Mat sT = transpose(sqrtInvR);
  for (auto& kv : mapCopy) kv.second = kv.first->scaling * (sT * kv.second);
  Mat ev(2, 1);
  ev(0, 0) = e[0];
  ev(1, 0) = e[1];
  Mat we = sT * ev;
  for (auto it = mapCopy.begin(); it != mapCopy.end(); ++it) {
    const int j1 = it->first->blockIndex;
    const Mat& J1 = it->second;
    Mat J1t = transpose(J1);
    Mat g = J1t * we;
    const int base = H.baseOfBlock(j1);
    for (int i = 0; i < g.r; ++i) rhs[base + i] -= g(i, 0);
    for (auto it2 = it; it2 != mapCopy.end(); ++it2) {
      const int j2 = it2->first->blockIndex;
      Mat* blk = H.block(j1, j2, true);
      *blk += J1t * it2->second;
    }
  }
}

// ---- BE/src/MEstimatorPolicies.cpp:13-125, BE/include/aslam/backend/MEstimatorPolicies.hpp ---------
struct MEstimator {
  virtual ~MEstimator() {}
  virtual double getWeight(double squaredError) const = 0;
};
struct NoMEstimator : MEstimator {  // MEstimatorPolicies.cpp:16-21
  double getWeight(double) const override { return 1.0; }
};
struct GemanMcClureMEstimator : MEstimator {  // :23-39
  double sigma2;
  explicit GemanMcClureMEstimator(double s2) : sigma2(s2) {}
  double getWeight(double error) const override {
    const double se = sigma2 + error;
    return sigma2 / (se * se);
  }
};
struct CauchyMEstimator : MEstimator {  // :41-57
  double sigma2;
  explicit CauchyMEstimator(double s2) : sigma2(s2) {}
  double getWeight(double error) const override {
    const double se = error / sigma2;
    return 1.0 / (1.0 + se);
  }
};
struct HuberMEstimator : MEstimator {  // :61-75
  double k, k2;
  explicit HuberMEstimator(double k_) : k(k_), k2(k_ * k_) {}
  double getWeight(double error) const override { return error < k2 ? 1.0 : k / std::sqrt(error); }
};
// Quantile of the chi-squared distribution.  The reference calls boost::math::quantile(chi_squared_distribution<>(df), p)
// (MEstimatorPolicies.cpp:116-119; Boost.Math is a third-party dependency that is absent here, apt libboost-all-dev, unpinned:
// packages-ubuntu-24.04.txt).  Restated from the definition: the x with P(df/2, x/2) = p, P the regularised lower incomplete
// gamma function (series for x < a + 1, Lentz continued fraction otherwise), solved by bisection then Newton; checked against
// scipy.stats.chi2.ppf in tests/test_oracle_cpu.py.
inline double regularizedGammaP(double a, double x) {
  if (x <= 0.0) return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double sum = 1.0 / a, term = sum;
    for (int n = 1; n < 10000; ++n) {
      term *= x / (a + n);
      sum += term;
      if (std::fabs(term) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}
inline double chi2InvCDF(double p, size_t df) {
  const double a = 0.5 * (double)df;
  double lo = 0.0, hi = std::max(1.0, 2.0 * a);
  while (regularizedGammaP(a, 0.5 * hi) < p) hi *= 2.0;
  for (int i = 0; i < 200; ++i) {
    const double mid = 0.5 * (lo + hi);
    if (regularizedGammaP(a, 0.5 * mid) < p) lo = mid; else hi = mid;
  }
  double x = 0.5 * (lo + hi);
  for (int i = 0; i < 3; ++i) {  // Newton polish: pdf of chi2(df)
    const double pdf = std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - std::lgamma(a)) * 0.5;
    if (!(pdf > 0.0)) break;
    const double nx = x - (regularizedGammaP(a, 0.5 * x) - p) / pdf;
    if (nx > lo && nx < hi) x = nx;
  }
  return x;
}
struct BlakeZissermanMEstimator : MEstimator {  // :77-123
  size_t df;
  double pCut, wCut, epsilon;
  explicit BlakeZissermanMEstimator(size_t df_, double pCut_ = 0.999, double wCut_ = 0.1)
      : df(df_), pCut(pCut_), wCut(wCut_), epsilon(computeEpsilon(df_, pCut_, wCut_)) {}
  double getWeight(double mahalanobis2) const override { return std::exp(-mahalanobis2) / (std::exp(-mahalanobis2) + epsilon); }
  static double computeEpsilon(size_t df, double pCut, double wCut) { return (1 - wCut) / wCut * std::exp(-chi2InvCDF(pCut, df)); }
};

// Schweizer-Messer/sm_eigen/include/sm/eigen/matrix_sqrt.hpp:21-40 for a 2x2 matrix: S = P^T L sqrt(D) of Eigen::LDLT, which
// pivots on the larger diagonal entry (first one on ties); A = S S^T.  Eigen (libeigen3-dev, unpinned; noble ships 3.4.0) is absent:
// restated from Eigen/src/Cholesky/LDLT.h (ldlt_inplace<Lower>::unblocked).
inline Mat computeMatrixSqrt2(const Mat& A) {
  const bool swap = std::fabs(A(1, 1)) > std::fabs(A(0, 0));
  const double a = swap ? A(1, 1) : A(0, 0), d = swap ? A(0, 0) : A(1, 1), b = A(1, 0);
  const double l10 = (std::fabs(a) > 0.0) ? b / a : b;
  const double d0 = a, d1 = d - l10 * (d0 * l10);
  Mat L(2, 2);
  L(0, 0) = 1.0;
  L(1, 0) = l10;
  L(1, 1) = 1.0;
  Mat R(2, 2);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) R(i, j) = L(swap ? 1 - i : i, j);
  const double s0 = std::sqrt(d0), s1 = std::sqrt(d1);
  for (int i = 0; i < 2; ++i) {
    R(i, 0) *= s0;
    R(i, 1) *= s1;
  }
  return R;
}

// ---- CVE/include/aslam/backend/implementation/ReprojectionError.hpp:27-75 over BE ErrorTermFs<2> -----
struct ReprojectionError {
  double y[2];
  HomogeneousExpressionNodeMultiply point;
  CameraDesignVariable* camera;
  double error[2] = {0, 0};
  double squaredError = 0.0;  // ErrorTerm::_squaredError: the raw e^T invR e of the last evaluateError()
  Mat sqrtInvR;  // computeMatrixSqrt(I) = I  (Schweizer-Messer/sm_eigen/include/sm/eigen/matrix_sqrt.hpp:21-40)
  std::shared_ptr<MEstimator> mEstimatorPolicy;  // BE/src/ErrorTerm.cpp:8-12: NoMEstimator by default
  int rowBase = 0;
  ReprojectionError(const double y_[2], const HomogeneousExpressionNodeMultiply& pt, CameraDesignVariable* cam)
      : point(pt), camera(cam), sqrtInvR(Mat::Identity(2)), mEstimatorPolicy(std::make_shared<NoMEstimator>()) {
    y[0] = y_[0];
    y[1] = y_[1];
  }
  // BE/include/aslam/backend/implementation/ErrorTerm.hpp:118-127
  void setInvR(const Mat& invR) { sqrtInvR = computeMatrixSqrt2(invR); }
  // BE/src/ErrorTerm.cpp:46-56
  void setMEstimatorPolicy(const std::shared_ptr<MEstimator>& m) { mEstimatorPolicy = m; }
  // ReprojectionError.hpp:50-60 ; the projection's validity bool is ignored (Q6)
  double evaluateErrorImplementation() {
    double p[4];
    point.toHomogeneous(p);
    double hat_y[2] = {0.0, 0.0};  // the reference leaves this uninitialised when the projection bails out
    camera->camera->homogeneousToKeypoint(p, hat_y);
    error[0] = y[0] - hat_y[0];
    error[1] = y[1] - hat_y[1];
    // e^T invR e with invR = sqrtInvR sqrtInvR^T
    Mat invR = sqrtInvR * transpose(sqrtInvR);
    return error[0] * (invR(0, 0) * error[0] + invR(0, 1) * error[1]) + error[1] * (invR(1, 0) * error[0] + invR(1, 1) * error[1]);
  }
  // BE/src/ErrorTerm.cpp:19-24: the returned cost is ALWAYS weighted by the policy (useMEstimator only gates e and J)
  double evaluateError() {
    squaredError = evaluateErrorImplementation();
    return mEstimatorPolicy->getWeight(squaredError) * squaredError;
  }
  // ReprojectionError.hpp:63-77
  void evaluateJacobians(JacobianContainer& out) const {
    double p[4];
    point.toHomogeneous(p);
    Mat J;
    double hat_y[2];
    camera->camera->homogeneousToKeypoint(p, hat_y, J);
    point.evaluateJacobians(out, -J);
    camera->evaluateJacobians(out, p);
  }
  double sqrtWeight(bool useMEstimator) const { return useMEstimator ? std::sqrt(mEstimatorPolicy->getWeight(squaredError)) : 1.0; }
  // BE/include/aslam/backend/implementation/ErrorTerm.hpp:97-109
  void buildHessian(SparseBlockMatrix& H, std::vector<double>& rhs, bool useMEstimator = true) {
    JacobianContainer J(2);
    evaluateJacobians(J);
    evaluateHessian(J, error, sqrtWeight(useMEstimator) * sqrtInvR, H, rhs);
  }
  // ErrorTerm.hpp:183-192
  void getWeightedError(double e[2], bool useMEstimator = true) const {
    Mat sT = transpose(sqrtInvR);
    const double sw = sqrtWeight(useMEstimator);
    e[0] = (sT(0, 0) * error[0] + sT(0, 1) * error[1]) * sw;
    e[1] = (sT(1, 0) * error[0] + sT(1, 1) * error[1]) * sw;
  }
  // ErrorTerm.hpp:170-181
  void getWeightedJacobians(JacobianContainer& out, bool useMEstimator = true) const {
    evaluateJacobians(out);
    Mat sT = transpose(sqrtInvR);
    const double sw = sqrtWeight(useMEstimator);
    for (auto& kv : out.jacobianMap) kv.second = (sw * kv.first->scaling) * (sT * kv.second);
  }
};

// fork/join over contiguous term ranges: BE/src/LinearSystemSolver.cpp:50-78
inline void setupThreadedJob(const std::function<void(size_t, size_t, size_t)>& job, size_t nThreads, size_t nTerms) {
  if (nThreads <= 1) {
    job(0, 0, nTerms);
    return;
  }
  nThreads = std::min(nThreads, nTerms);
  std::vector<size_t> indices(nThreads + 1, 0);
  const size_t per = std::max<size_t>(1, nTerms / nThreads);
  for (size_t i = 0; i < nThreads; ++i) indices[i + 1] = indices[i] + per;
  indices.back() = nTerms;
  std::vector<std::thread> threads;
  for (size_t i = 0; i < nThreads; ++i) threads.emplace_back(job, i, indices[i], indices[i + 1]);
  for (auto& t : threads) t.join();
}

}  // namespace ko
