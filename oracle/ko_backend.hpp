// ORACLE — TEST INFRASTRUCTURE ONLY (see ko_math.hpp header).  PARITY UNPINNED.
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

// ---- CVE/include/aslam/backend/implementation/ReprojectionError.hpp:27-75 over BE ErrorTermFs<2> -----
struct ReprojectionError {
  double y[2];
  HomogeneousExpressionNodeMultiply point;
  CameraDesignVariable* camera;
  double error[2] = {0, 0};
  Mat sqrtInvR;  // computeMatrixSqrt(I) = I  (Schweizer-Messer/sm_eigen/include/sm/eigen/matrix_sqrt.hpp:21-40)
  int rowBase = 0;
  ReprojectionError(const double y_[2], const HomogeneousExpressionNodeMultiply& pt, CameraDesignVariable* cam)
      : point(pt), camera(cam), sqrtInvR(Mat::Identity(2)) {
    y[0] = y_[0];
    y[1] = y_[1];
  }
  // ReprojectionError.hpp:50-60 ; the projection's validity bool is ignored (Q6)
  double evaluateError() {
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
  // BE/include/aslam/backend/implementation/ErrorTerm.hpp:97-109 (no M-estimator: weight 1)
  void buildHessian(SparseBlockMatrix& H, std::vector<double>& rhs) {
    JacobianContainer J(2);
    evaluateJacobians(J);
    evaluateHessian(J, error, sqrtInvR, H, rhs);
  }
  // ErrorTerm.hpp:183-192
  void getWeightedError(double e[2]) const {
    Mat sT = transpose(sqrtInvR);
    e[0] = sT(0, 0) * error[0] + sT(0, 1) * error[1];
    e[1] = sT(1, 0) * error[0] + sT(1, 1) * error[1];
  }
  // ErrorTerm.hpp:170-181
  void getWeightedJacobians(JacobianContainer& out) const {
    evaluateJacobians(out);
    Mat sT = transpose(sqrtInvR);
    for (auto& kv : out.jacobianMap) kv.second = kv.first->scaling * (sT * kv.second);
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
