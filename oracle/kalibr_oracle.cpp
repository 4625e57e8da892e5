// ORACLE — TEST INFRASTRUCTURE ONLY (see ko_math.hpp header: parity PINNED against the reference's own compiled code for the per-term part, the optimiser loop and both solver regimes, UNPINNED for the CHOLMOD factorisation itself).
//
// Problem construction (kalibr2 drivers), LinearSystemSolver / BlockCholesky / SparseCholesky semantics,
// LevenbergMarquardtTrustRegionPolicy and Optimizer2, restated on the CPU behind a small C API (ko_*)
// that tests/ and bench.py's CPU legs call through ctypes.  CHOLMOD (SuiteSparse, absent) is replaced by
// an exact block-arrow Cholesky: elimination of the per-set pose blocks (the semantic of
// BE/src/sparse_matrix_functions.cpp:8-83) followed by a dense Cholesky of the reduced system; a dense
// Cholesky of the whole matrix is kept as a cross-check.
//
// K2 = aslam_offline_calibration/kalibr2/include/kalibr2, BE = aslam_optimizer/aslam_backend.
#include <array>
#include <chrono>
#include <cstdio>

#include "../include/kalibr_b200.h"
#include "ko_backend.hpp"
#include "ko_marginal.hpp"

namespace ko {

// ---- dense helpers (stand-ins for CHOLMOD / Eigen LDLT) ---------------------------------------------
// In-place lower Cholesky of a row-major n x n SPD matrix; false if a pivot is not positive.
static bool choleskyInPlace(std::vector<double>& A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}
static void choleskySolve(const std::vector<double>& L, int n, double* x /*in: b, out: x*/) {
  for (int i = 0; i < n; ++i) {
    double s = x[i];
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * x[k];
    x[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * x[k];
    x[i] = s / L[(size_t)i * n + i];
  }
}

// ---- CCS J^T: BE/include/aslam/backend/implementation/CompressedColumnJacobianTransposeBuilder.hpp:19-101,
//      BE/include/aslam/backend/CompressedColumnMatrix.hpp:236-304 ---------------------------------------
struct CompressedColumnJt {
  std::vector<int64_t> col_ptr;  // 2 columns per term
  std::vector<int32_t> row_idx;
  std::vector<double> values;
};

struct Problem;

// ---- BE/include/aslam/backend/LinearSystemSolver.hpp:16-109, BE/src/LinearSystemSolver.cpp ------------
struct LinearSystemSolver {
  Problem* problem = nullptr;
  std::vector<ReprojectionError*> errorTerms;
  std::vector<double> e, rhs, diagonalConditioner;
  size_t JRows = 0, JCols = 0;
  bool useDiagonalConditioner = true;
  bool useMEstimator = true;  // Optimizer2 always passes true (Optimizer2.cpp:198, 237; LevenbergMarquardtTrustRegionPolicy.cpp:56, 72)
  virtual ~LinearSystemSolver() {}
  // LinearSystemSolver.cpp:12-23, 81-92
  double evaluateError(size_t nThreads) {
    nThreads = std::max<size_t>(1, nThreads);
    std::vector<double> threadLocalErrors(nThreads, 0.0);
    setupThreadedJob(
        [&](size_t tid, size_t a, size_t b) {
          for (size_t i = a; i < b; ++i) {
            threadLocalErrors[tid] += errorTerms[i]->evaluateError();
            double we[2];
            errorTerms[i]->getWeightedError(we, useMEstimator);
            e[errorTerms[i]->rowBase] = -we[0];
            e[errorTerms[i]->rowBase + 1] = -we[1];
          }
        },
        nThreads, errorTerms.size());
    double err = 0.0;
    for (double v : threadLocalErrors) err += v;
    return err;
  }
  void setConstantConditioner(double d) { diagonalConditioner.assign(JCols, d); }  // LinearSystemSolver.cpp:111-114
  void initMatrixStructure(const std::vector<DesignVariable*>& dvs, const std::vector<ReprojectionError*>& errs, bool useDiag) {
    errorTerms = errs;
    JRows = 2 * errs.size();
    JCols = 0;
    for (auto* dv : dvs) JCols += dv->minimalDimensions();
    e.assign(JRows, 0.0);
    rhs.assign(JCols, 0.0);
    diagonalConditioner.assign(JCols, 0.0);
    initMatrixStructureImplementation(dvs, useDiag);
  }
  virtual void initMatrixStructureImplementation(const std::vector<DesignVariable*>& dvs, bool useDiag) = 0;
  virtual void buildSystem(size_t nThreads) = 0;
  virtual bool solveSystem(std::vector<double>& dx) = 0;
  virtual const char* name() const = 0;
};

// ---- the problem: K2/CalibrationTools.hpp:32-45, 93-144, 183-300, 376-428; K2/CameraCalibrator.hpp:116-122, 238-265 ----
struct Problem {
  int driverOrder = 0, nCams = 0, nSets = 0;
  std::vector<std::unique_ptr<CameraDesignVariable>> cams;
  std::vector<std::shared_ptr<RotationQuaternion>> baseQ, setQ;
  std::vector<std::shared_ptr<EuclideanPoint>> baseT, setT;
  std::vector<std::shared_ptr<TransformationBasic>> baseTf, setTf;
  std::vector<DesignVariable*> problemDvs;      // OptimizationProblem insertion order
  std::vector<DesignVariable*> designVariables; // active ones, Optimizer2::initialize
  std::vector<std::unique_ptr<ReprojectionError>> terms;
  std::vector<ReprojectionError*> errorTerms;
  std::vector<int> termCam, termSet;
  std::vector<std::vector<int>> setCams;  // cameras observing each set (for the arrow solve)
  std::unique_ptr<LinearSystemSolver> solver;
  std::vector<double> dx;
  double J = 0, p_J = 0;

  void addPose(const double* pose7, std::vector<std::shared_ptr<RotationQuaternion>>& qs, std::vector<std::shared_ptr<EuclideanPoint>>& ts,
               std::vector<std::shared_ptr<TransformationBasic>>& tfs) {
    auto q = std::make_shared<RotationQuaternion>(pose7);
    q->active = true;
    problemDvs.push_back(q.get());
    auto t = std::make_shared<EuclideanPoint>(pose7 + 4);
    t->active = true;
    problemDvs.push_back(t.get());
    qs.push_back(q);
    ts.push_back(t);
    tfs.push_back(std::make_shared<TransformationBasic>(q, t));
  }
  void addIntrinsics(int k) {  // CameraCalibrator.hpp:116-122: setActive(true, true, false)
    cams[k]->setActive(true, true);
    problemDvs.push_back(cams[k]->projectionDv.get());
    problemDvs.push_back(cams[k]->distortionDv.get());
  }

  explicit Problem(const kb_problem_desc* d) {
    driverOrder = d->driver_order;
    nCams = d->n_cams;
    nSets = d->n_sets;
    for (int k = 0; k < nCams; ++k) {
      auto cam = makeCamera(d->cam_model[k], d->cam_params + (size_t)k * KB_CAM_PARAM_STRIDE);
      if (!cam) throw std::runtime_error("unknown camera model");
      cams.emplace_back(new CameraDesignVariable(std::move(cam)));
    }
    auto addBaselines = [&]() {
      for (int k = 0; k + 1 < nCams; ++k) addPose(d->baselines + (size_t)k * KB_POSE_STRIDE, baseQ, baseT, baseTf);
    };
    auto addSets = [&]() {
      for (int v = 0; v < nSets; ++v) addPose(d->set_poses + (size_t)v * KB_POSE_STRIDE, setQ, setT, setTf);
    };
    switch (driverOrder) {
      case KB_ORDER_SINGLE:  // CalibrationTools.hpp:102-134
        if (nCams != 1) throw std::runtime_error("single-camera order needs one camera");
        addIntrinsics(0);
        addSets();
        break;
      case KB_ORDER_STEREO:  // CalibrationTools.hpp:236-262
        if (nCams != 2) throw std::runtime_error("stereo order needs two cameras");
        addBaselines();
        addSets();
        addIntrinsics(0);
        addIntrinsics(1);
        break;
      case KB_ORDER_RIG:  // CalibrationTools.hpp:380-399
        for (int k = 0; k < nCams; ++k) addIntrinsics(k);
        addBaselines();
        addSets();
        break;
      case KB_ORDER_BATCH:  // the incremental estimator's merged problem: CalibrationTools.hpp:460-491 (groups 1, 0, 2 by first appearance,
                            // IncrementalOptimizationProblem.cpp:186-223), marginalised group 0 moved last (IncrementalEstimator.cpp:550-565);
                            // inside group 0 the baselines come before the intrinsics (CalibrationTools.hpp:470-486)
        addSets();
        addBaselines();
        for (int k = 0; k < nCams; ++k) addIntrinsics(k);
        break;
      default: throw std::runtime_error("unknown driver order");
    }
    // error terms in the caller's (reference insertion) order
    setCams.assign(nSets, std::vector<int>());
    for (int w = 0; w < d->n_views; ++w) {
      const int v = d->view_set[w], k = d->view_cam[w];
      if (v < 0 || v >= nSets || k < 0 || k >= nCams) throw std::runtime_error("view index out of range");
      if (d->view_begin[w + 1] > d->view_begin[w]) setCams[v].push_back(k);
      // T_cam_w = B_{k-1} * ... * B_0 * inverse(T_v): CalibrationTools.hpp:405-408
      std::shared_ptr<TransformationExpressionNode> T = std::make_shared<TransformationExpressionNodeInverse>(setTf[v]);
      for (int j = 0; j < k; ++j) T = std::make_shared<TransformationExpressionNodeMultiply>(baseTf[j], T);
      for (int64_t i = d->view_begin[w]; i < d->view_begin[w + 1]; ++i) {
        const double* tp = d->target_points + 3 * (size_t)d->corner_id[i];
        const double ph[4] = {tp[0], tp[1], tp[2], 1.0};  // toHomogeneous(target->point(i)): CameraCalibrator.hpp:246
        const double y[2] = {d->y_u[i], d->y_v[i]};
        terms.emplace_back(new ReprojectionError(y, HomogeneousExpressionNodeMultiply(T, ph), cams[k].get()));
        termCam.push_back(k);
        termSet.push_back(v);
      }
    }
  }

  // BE/src/Optimizer2.cpp:95-151
  void initialize(std::unique_ptr<LinearSystemSolver> s) {
    solver = std::move(s);
    solver->problem = this;
    designVariables.clear();
    for (auto* dv : problemDvs)
      if (dv->active) designVariables.push_back(dv);
    int columnBase = 0;
    for (size_t i = 0; i < designVariables.size(); ++i) {
      designVariables[i]->blockIndex = (int)i;
      designVariables[i]->columnBase = columnBase;
      columnBase += designVariables[i]->minimalDimensions();
    }
    errorTerms.clear();
    int dim = 0;
    for (auto& t : terms) {
      errorTerms.push_back(t.get());
      t->rowBase = dim;
      dim += 2;
    }
    solver->initMatrixStructure(designVariables, errorTerms, true /* LM requiresAugmentedDiagonal */);
  }
  double evaluateError(size_t nThreads) {
    J = solver->evaluateError(nThreads);
    return J;
  }
  // Optimizer2.cpp:290-307
  double applyStateUpdate() {
    int startIdx = 0;
    double maxAbs = 0;
    for (auto* d : designVariables) {
      const int dbd = d->minimalDimensions();
      std::vector<double> dxS(dx.begin() + startIdx, dx.begin() + startIdx + dbd);
      for (double& v : dxS) v *= d->scaling;
      if (dbd > 0) d->update(dxS.data(), dbd);
      startIdx += dbd;
    }
    for (double v : dx) maxAbs = std::max(maxAbs, std::fabs(v));
    return maxAbs;
  }
  void revertLastStateUpdate() {  // Optimizer2.cpp:313-318
    for (auto* d : designVariables) d->revertUpdate();
  }
};

// ---- exact block-arrow Cholesky standing in for LinearSolverCholmod::solve
//      (SBM/include/sparse_block_matrix/linear_solver_cholmod.h:70-112) ---------------------------------
static bool solveArrow(Problem& P, const SparseBlockMatrix& H, const std::vector<double>& rhs, std::vector<double>& dx) {
  const int n = H.rows();
  dx.assign(n, 0.0);
  // reduced (camera-side) design variables = everything that is not a per-set pose
  std::vector<char> isPose(H.bRows(), 0);
  for (int v = 0; v < P.nSets; ++v) {
    isPose[P.setQ[v]->blockIndex] = 1;
    isPose[P.setT[v]->blockIndex] = 1;
  }
  std::vector<int> redBase(H.bRows(), -1);
  int nc = 0;
  std::vector<int> redBlocks;
  for (int b = 0; b < H.bRows(); ++b)
    if (!isPose[b]) {
      redBase[b] = nc;
      nc += H.dimOfBlock(b);
      redBlocks.push_back(b);
    }
  auto getBlock = [&](int r, int c, Mat& out) -> bool {  // full symmetric access to the upper-stored blocks
    if (r <= c) {
      const Mat* b = H.block(r, c);
      if (!b) return false;
      out = *b;
      return true;
    }
    const Mat* b = H.block(c, r);
    if (!b) return false;
    out = transpose(*b);
    return true;
  };
  std::vector<double> A((size_t)nc * nc, 0.0), b(nc, 0.0);
  for (int bi : redBlocks)
    for (int bj : redBlocks) {
      if (bi > bj) continue;
      const Mat* blk = H.block(bi, bj);
      if (!blk) continue;
      for (int i = 0; i < blk->r; ++i)
        for (int j = 0; j < blk->c; ++j) {
          A[(size_t)(redBase[bi] + i) * nc + redBase[bj] + j] = (*blk)(i, j);
          A[(size_t)(redBase[bj] + j) * nc + redBase[bi] + i] = (*blk)(i, j);
        }
    }
  for (int bi : redBlocks)
    for (int i = 0; i < H.dimOfBlock(bi); ++i) b[redBase[bi] + i] = rhs[H.baseOfBlock(bi) + i];

  struct SetFactor {
    std::vector<double> L;       // 6x6 Cholesky factor of V_v
    std::vector<int> blocks;     // coupled reduced blocks
    std::vector<Mat> W;          // dim x 6 each
  };
  std::vector<SetFactor> factors(P.nSets);
  bool ok = true;
  for (int v = 0; v < P.nSets; ++v) {
    const int bq = P.setQ[v]->blockIndex, bt = P.setT[v]->blockIndex;
    SetFactor& f = factors[v];
    f.L.assign(36, 0.0);
    Mat Vqq, Vqt, Vtt;
    if (!getBlock(bq, bq, Vqq)) Vqq = Mat(3, 3);
    if (!getBlock(bq, bt, Vqt)) Vqt = Mat(3, 3);
    if (!getBlock(bt, bt, Vtt)) Vtt = Mat(3, 3);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        f.L[(size_t)i * 6 + j] = Vqq(i, j);
        f.L[(size_t)i * 6 + 3 + j] = Vqt(i, j);
        f.L[(size_t)(3 + j) * 6 + i] = Vqt(i, j);
        f.L[(size_t)(3 + i) * 6 + 3 + j] = Vtt(i, j);
      }
    if (!choleskyInPlace(f.L, 6)) {
      ok = false;
      break;
    }
    // coupled camera-side DVs of this set
    std::vector<char> seen(H.bRows(), 0);
    for (int k : P.setCams[v]) {
      std::vector<int> cand = {P.cams[k]->projectionDv->blockIndex, P.cams[k]->distortionDv->blockIndex};
      for (int j = 0; j < k; ++j) {
        cand.push_back(P.baseQ[j]->blockIndex);
        cand.push_back(P.baseT[j]->blockIndex);
      }
      for (int c : cand)
        if (c >= 0 && !seen[c]) {
          seen[c] = 1;
          f.blocks.push_back(c);
        }
    }
    std::sort(f.blocks.begin(), f.blocks.end());
    const double* rv_q = &rhs[H.baseOfBlock(bq)];
    const double* rv_t = &rhs[H.baseOfBlock(bt)];
    double viRhs[6] = {rv_q[0], rv_q[1], rv_q[2], rv_t[0], rv_t[1], rv_t[2]};
    choleskySolve(f.L, 6, viRhs);  // V^-1 rhs_v
    for (int c : f.blocks) {
      Mat Wq, Wt;
      const int dc = H.dimOfBlock(c);
      if (!getBlock(c, bq, Wq)) Wq = Mat(dc, 3);
      if (!getBlock(c, bt, Wt)) Wt = Mat(dc, 3);
      Mat W(dc, 6);
      for (int i = 0; i < dc; ++i)
        for (int j = 0; j < 3; ++j) {
          W(i, j) = Wq(i, j);
          W(i, 3 + j) = Wt(i, j);
        }
      f.W.push_back(W);
    }
    // Y = W V^-1 ;  A -= Y W^T ; b -= W V^-1 rhs_v   (sparse_matrix_functions.cpp:22-52)
    for (size_t a = 0; a < f.blocks.size(); ++a) {
      const Mat& Wa = f.W[a];
      Mat Y(Wa.r, 6);
      for (int i = 0; i < Wa.r; ++i) {
        double row[6];
        for (int j = 0; j < 6; ++j) row[j] = Wa(i, j);
        choleskySolve(f.L, 6, row);
        for (int j = 0; j < 6; ++j) Y(i, j) = row[j];
        double s = 0;
        for (int j = 0; j < 6; ++j) s += Wa(i, j) * viRhs[j];
        b[redBase[f.blocks[a]] + i] -= s;
      }
      for (size_t c2 = 0; c2 < f.blocks.size(); ++c2) {
        const Mat& Wb = f.W[c2];
        for (int i = 0; i < Wa.r; ++i)
          for (int j = 0; j < Wb.r; ++j) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += Y(i, k) * Wb(j, k);
            A[(size_t)(redBase[f.blocks[a]] + i) * nc + redBase[f.blocks[c2]] + j] -= s;
          }
      }
    }
  }
  if (!ok) return false;
  if (nc > 0) {
    if (!choleskyInPlace(A, nc)) return false;
    choleskySolve(A, nc, b.data());
  }
  for (int bi : redBlocks)
    for (int i = 0; i < H.dimOfBlock(bi); ++i) dx[H.baseOfBlock(bi) + i] = b[redBase[bi] + i];
  // back-substitution: ds_v = V^-1 (rhs_v - W^T dx_c)   (sparse_matrix_functions.cpp:64-83)
  for (int v = 0; v < P.nSets; ++v) {
    const int bq = P.setQ[v]->blockIndex, bt = P.setT[v]->blockIndex;
    SetFactor& f = factors[v];
    double r[6];
    for (int i = 0; i < 3; ++i) {
      r[i] = rhs[H.baseOfBlock(bq) + i];
      r[3 + i] = rhs[H.baseOfBlock(bt) + i];
    }
    for (size_t a = 0; a < f.blocks.size(); ++a) {
      const Mat& W = f.W[a];
      for (int i = 0; i < W.r; ++i) {
        const double x = b[redBase[f.blocks[a]] + i];
        for (int j = 0; j < 6; ++j) r[j] -= W(i, j) * x;
      }
    }
    choleskySolve(f.L, 6, r);
    for (int i = 0; i < 3; ++i) {
      dx[H.baseOfBlock(bq) + i] = r[i];
      dx[H.baseOfBlock(bt) + i] = r[3 + i];
    }
  }
  return true;
}

static bool solveDense(const SparseBlockMatrix& H, const std::vector<double>& rhs, std::vector<double>& dx) {
  const int n = H.rows();
  std::vector<double> A((size_t)n * n, 0.0);
  for (int c = 0; c < H.bRows(); ++c)
    for (auto& kv : H.blockCols[c]) {
      const Mat& blk = *kv.second;
      const int rb = H.baseOfBlock(kv.first), cb = H.baseOfBlock(c);
      for (int i = 0; i < blk.r; ++i)
        for (int j = 0; j < blk.c; ++j) {
          A[(size_t)(rb + i) * n + cb + j] = blk(i, j);
          A[(size_t)(cb + j) * n + rb + i] = blk(i, j);
        }
    }
  dx = rhs;
  if (!choleskyInPlace(A, n)) return false;
  choleskySolve(A, n, dx.data());
  return true;
}

// ---- BE/src/BlockCholeskyLinearSystemSolver.cpp:34-106 --------------------------------------------------
struct BlockCholeskyLinearSystemSolver : LinearSystemSolver {
  SparseBlockMatrix H;
  bool denseCrossCheck = false;
  const char* name() const override { return "block_cholesky"; }
  void initMatrixStructureImplementation(const std::vector<DesignVariable*>& dvs, bool useDiag) override {
    useDiagonalConditioner = useDiag;
    std::vector<int> blocks;
    for (size_t i = 0; i < dvs.size(); ++i) {
      dvs[i]->blockIndex = (int)i;
      blocks.push_back(dvs[i]->minimalDimensions());
    }
    for (size_t i = 1; i < blocks.size(); ++i) blocks[i] += blocks[i - 1];
    H.reset(blocks);
  }
  void buildSystem(size_t /*nThreads: ignored, serial by design*/) override {
    H.clear(false);
    std::fill(rhs.begin(), rhs.end(), 0.0);
    for (auto* t : errorTerms) t->buildHessian(H, rhs, useMEstimator);
  }
  bool solveSystem(std::vector<double>& dx) override {
    if (useDiagonalConditioner) {
      int rowBase = 0;
      for (int i = 0; i < H.bRows(); ++i) {
        Mat& blk = *H.block(i, i, true);
        for (int k = 0; k < blk.r; ++k) blk(k, k) += diagonalConditioner[rowBase + k] * diagonalConditioner[rowBase + k];
        rowBase += blk.r;
      }
    }
    bool ok = denseCrossCheck ? solveDense(H, rhs, dx) : solveArrow(*problem, H, rhs, dx);
    if (useDiagonalConditioner) {
      int rowBase = 0;
      for (int i = 0; i < H.bRows(); ++i) {
        Mat& blk = *H.block(i, i, true);
        for (int k = 0; k < blk.r; ++k) blk(k, k) -= diagonalConditioner[rowBase + k];  // sic: lambda, not lambda^2 (Q2)
        rowBase += blk.r;
      }
    }
    return ok;
  }
};

// ---- BE/src/SparseCholeskyLinearSystemSolver.cpp:14-89 ---------------------------------------------------
struct SparseCholeskyLinearSystemSolver : LinearSystemSolver {
  CompressedColumnJt Jt;
  SparseBlockMatrix H;  // J^T J + diag^2, formed for the stand-in factorisation
  const char* name() const override { return "sparse_cholesky"; }
  void initMatrixStructureImplementation(const std::vector<DesignVariable*>& dvs, bool useDiag) override {
    useDiagonalConditioner = useDiag;
    std::vector<int> blocks;
    for (auto* dv : dvs) blocks.push_back(dv->minimalDimensions());
    for (size_t i = 1; i < blocks.size(); ++i) blocks[i] += blocks[i - 1];
    H.reset(blocks);
    // structure: every term contributes two columns whose rows are its DV columns sorted by block index
    Jt.col_ptr.assign(2 * errorTerms.size() + 1, 0);
    Jt.row_idx.clear();
    for (size_t i = 0; i < errorTerms.size(); ++i) {
      JacobianContainer jc(2);
      errorTerms[i]->evaluateJacobians(jc);
      std::vector<int32_t> rows;
      for (auto& kv : jc.jacobianMap)
        for (int c = 0; c < kv.second.c; ++c) rows.push_back(kv.first->columnBase + c);
      for (int r = 0; r < 2; ++r) {
        Jt.col_ptr[2 * i + r + 1] = Jt.col_ptr[2 * i + r] + (int64_t)rows.size();
        Jt.row_idx.insert(Jt.row_idx.end(), rows.begin(), rows.end());
      }
    }
    Jt.values.assign(Jt.row_idx.size(), 0.0);
  }
  void buildSystem(size_t nThreads) override {
    // CompressedColumnJacobianTransposeBuilder.hpp:59-100: threaded getWeightedJacobians + writeJacobians
    setupThreadedJob(
        [&](size_t, size_t a, size_t b) {
          for (size_t i = a; i < b; ++i) {
            JacobianContainer jc(2);
            errorTerms[i]->getWeightedJacobians(jc, useMEstimator);
            for (int r = 0; r < 2; ++r) {
              double* out = &Jt.values[Jt.col_ptr[2 * i + r]];
              for (auto& kv : jc.jacobianMap)
                for (int c = 0; c < kv.second.c; ++c) *out++ = kv.second(r, c);
            }
          }
        },
        std::max<size_t>(1, nThreads), errorTerms.size());
    // rhs = J^T * e  with e = -(weighted error): SparseCholeskyLinearSystemSolver.cpp:39-46
    std::fill(rhs.begin(), rhs.end(), 0.0);
    for (size_t c = 0; c + 1 < Jt.col_ptr.size(); ++c)
      for (int64_t k = Jt.col_ptr[c]; k < Jt.col_ptr[c + 1]; ++k) rhs[Jt.row_idx[k]] += Jt.values[k] * e[c];
  }
  bool solveSystem(std::vector<double>& dx) override {
    // CHOLMOD factorises [J^T | diag] [J^T | diag]^T = J^T J + diag^2 (SparseCholeskyLinearSystemSolver.cpp:48-66).
    H.clear(false);
    std::vector<int> blockOfCol(JCols);
    for (int b = 0; b < H.bRows(); ++b)
      for (int i = 0; i < H.dimOfBlock(b); ++i) blockOfCol[H.baseOfBlock(b) + i] = b;
    for (size_t c = 0; c + 1 < Jt.col_ptr.size(); ++c) {
      const int64_t a = Jt.col_ptr[c], bnd = Jt.col_ptr[c + 1];
      for (int64_t i = a; i < bnd; ++i)
        for (int64_t j = i; j < bnd; ++j) {
          const int ri = Jt.row_idx[i], rj = Jt.row_idx[j];
          const int bi = blockOfCol[ri], bj = blockOfCol[rj];
          (*H.block(bi, bj, true))(ri - H.baseOfBlock(bi), rj - H.baseOfBlock(bj)) += Jt.values[i] * Jt.values[j];
        }
    }
    // mirror the in-block lower triangles of diagonal blocks
    for (int b = 0; b < H.bRows(); ++b) {
      Mat& blk = *H.block(b, b, true);
      for (int i = 0; i < blk.r; ++i)
        for (int j = 0; j < i; ++j) blk(i, j) = blk(j, i);
      if (useDiagonalConditioner)
        for (int k = 0; k < blk.r; ++k) blk(k, k) += diagonalConditioner[H.baseOfBlock(b) + k] * diagonalConditioner[H.baseOfBlock(b) + k];
    }
    return solveArrow(*problem, H, rhs, dx);
  }
};

// ---- BE/src/TrustRegionPolicy.cpp:28-57 + BE/src/LevenbergMarquardtTrustRegionPolicy.cpp:37-113 -----------
struct LevenbergMarquardtTrustRegionPolicy {
  double lambdaInit, gammaInit = 3, betaInit = 2, muInit = 2;
  int pInit = 3;
  double lambda = 0, gamma = 0, beta = 0, mu = 0;
  int p = 0;
  double J = 0, p_J = 0, last_successful_J = 0;
  bool isFirstIteration = true;
  LinearSystemSolver* solver = nullptr;
  std::vector<double> dx;
  explicit LevenbergMarquardtTrustRegionPolicy(double l0) : lambdaInit(l0) {}
  void optimizationStarting(double J0) {
    J = p_J = last_successful_J = J0;
    isFirstIteration = true;
    lambda = lambdaInit; gamma = gammaInit; beta = betaInit; p = pInit; mu = muInit;
  }
  double getLmRho() {
    const double d1 = p_J - J;
    double d2 = 0;
    for (size_t i = 0; i < dx.size(); ++i) d2 += dx[i] * (lambda * dx[i] + solver->rhs[i]);
    return d1 / d2;
  }
  bool solveSystem(double Jnow, bool previousIterationFailed, int nThreads, std::vector<double>& outDx) {
    if (previousIterationFailed) {
      J = Jnow;
    } else {
      p_J = last_successful_J;
      last_successful_J = Jnow;
      J = Jnow;
    }
    if (isFirstIteration) {
      solver->buildSystem(nThreads);
    } else {
      const double rho = getLmRho();
      if (previousIterationFailed) {
        mu *= 2;
        lambda *= mu;
      } else if (rho <= 0) {
        mu *= 10;
        lambda *= mu;
      } else {
        solver->buildSystem(nThreads);
        if (lambda > 1e-16) {
          const double u1 = 1 / gamma;
          const double u2 = 1 - (beta - 1) * std::pow((2 * rho - 1), p);
          lambda *= (u1 > u2) ? u1 : u2;
          mu = beta;
        } else {
          lambda = 1e-15;
        }
      }
    }
    solver->setConstantConditioner(lambda);
    const bool success = solver->solveSystem(dx);
    outDx = dx;
    isFirstIteration = false;
    return success;
  }
};

// ---- BE/src/Optimizer2.cpp:183-273 -------------------------------------------------------------------------
static void optimize(Problem& P, const kb_optimizer_options& opt, int nThreads, kb_solution* out, std::vector<double>* trace) {
  kb_solution srv;
  std::memset(&srv, 0, sizeof(srv));
  P.p_J = 0.0;
  P.evaluateError(nThreads);
  P.p_J = P.J;
  srv.j_start = P.p_J;
  double deltaX = opt.convergence_delta_x + 1.0;
  double deltaJ = opt.convergence_delta_j + 1.0;
  bool previousIterationFailed = false;
  bool linearSolverFailure = false;
  LevenbergMarquardtTrustRegionPolicy policy(opt.lm_lambda_init);
  policy.solver = P.solver.get();
  policy.optimizationStarting(P.J);
  while (srv.iterations < opt.max_iterations && srv.failed_iterations < opt.max_iterations &&
         ((deltaX > opt.convergence_delta_x && std::fabs(deltaJ) > opt.convergence_delta_j) || linearSolverFailure)) {
    const bool solutionSuccess = policy.solveSystem(P.J, previousIterationFailed, nThreads, P.dx);
    if (!solutionSuccess) {
      previousIterationFailed = true;
      linearSolverFailure = true;
      srv.failed_iterations++;
    } else {
      deltaX = P.applyStateUpdate();
      P.evaluateError(nThreads);
      deltaJ = P.p_J - P.J;
      if (deltaJ < 0.0) {
        P.revertLastStateUpdate();
        srv.failed_iterations++;
        previousIterationFailed = true;
      } else {
        P.p_J = P.J;
        previousIterationFailed = false;
      }
      srv.iterations++;
      if (trace) {
        trace->push_back(P.J);
        trace->push_back(deltaX);
        trace->push_back(policy.lambda);
      }
      if (opt.verbose) std::printf("[oracle %d]: J: %.10g, dJ: %.6g, deltaX: %.6g, lambda: %.6g\n", srv.iterations, P.J, deltaJ, deltaX, policy.lambda);
    }
  }
  srv.j_final = P.p_J;
  srv.dx_final = deltaX;
  srv.dj_final = deltaJ;
  srv.linear_solver_failure = linearSolverFailure;
  *out = srv;
}

}  // namespace ko

// =============================================================================================================
// C API (ctypes)
// =============================================================================================================
using namespace ko;

struct ko_problem {
  std::unique_ptr<Problem> P;
  std::vector<double> trace;
  std::string error;
};

static thread_local std::string g_ko_error;

extern "C" {

__attribute__((visibility("default"))) const char* ko_last_error() { return g_ko_error.c_str(); }

// solver_kind: 0 = BlockCholesky semantic, 1 = SparseCholesky semantic, 2 = BlockCholesky with a dense full-matrix solve
__attribute__((visibility("default"))) ko_problem* ko_create(const kb_problem_desc* d, int solver_kind) {
  try {
    std::unique_ptr<ko_problem> h(new ko_problem());
    h->P.reset(new Problem(d));
    std::unique_ptr<LinearSystemSolver> s;
    if (solver_kind == 1) {
      s.reset(new SparseCholeskyLinearSystemSolver());
    } else {
      auto* b = new BlockCholeskyLinearSystemSolver();
      b->denseCrossCheck = (solver_kind == 2);
      s.reset(b);
    }
    h->P->initialize(std::move(s));
    return h.release();
  } catch (const std::exception& e) {
    g_ko_error = e.what();
    return nullptr;
  }
}
__attribute__((visibility("default"))) void ko_destroy(ko_problem* h) { delete h; }
__attribute__((visibility("default"))) int64_t ko_jrows(ko_problem* h) { return (int64_t)h->P->solver->JRows; }
__attribute__((visibility("default"))) int64_t ko_jcols(ko_problem* h) { return (int64_t)h->P->solver->JCols; }
__attribute__((visibility("default"))) int32_t ko_num_design_variables(ko_problem* h) { return (int32_t)h->P->designVariables.size(); }
__attribute__((visibility("default"))) void ko_get_dv_layout(ko_problem* h, int32_t* column_base, int32_t* dims) {
  for (size_t i = 0; i < h->P->designVariables.size(); ++i) {
    column_base[i] = h->P->designVariables[i]->columnBase;
    dims[i] = h->P->designVariables[i]->minimalDimensions();
  }
}
__attribute__((visibility("default"))) double ko_evaluate_error(ko_problem* h, int nThreads) { return h->P->evaluateError(nThreads); }
__attribute__((visibility("default"))) void ko_get_error_vector(ko_problem* h, double* e) {
  std::memcpy(e, h->P->solver->e.data(), sizeof(double) * h->P->solver->e.size());
}
__attribute__((visibility("default"))) void ko_build_system(ko_problem* h, int nThreads) { h->P->solver->buildSystem(nThreads); }
__attribute__((visibility("default"))) void ko_set_constant_conditioner(ko_problem* h, double l) { h->P->solver->setConstantConditioner(l); }
__attribute__((visibility("default"))) int32_t ko_solve_system(ko_problem* h, double* dx) {
  const bool ok = h->P->solver->solveSystem(h->P->dx);
  if (ok && dx) std::memcpy(dx, h->P->dx.data(), sizeof(double) * h->P->dx.size());
  return ok ? 1 : 0;
}
__attribute__((visibility("default"))) void ko_get_rhs(ko_problem* h, double* rhs) {
  std::memcpy(rhs, h->P->solver->rhs.data(), sizeof(double) * h->P->solver->rhs.size());
}
__attribute__((visibility("default"))) double ko_apply_state_update(ko_problem* h) { return h->P->applyStateUpdate(); }
__attribute__((visibility("default"))) void ko_revert_last_state_update(ko_problem* h) { h->P->revertLastStateUpdate(); }
// Optimizer2::applyStateUpdate with a step computed outside (the estimator's linear solver restated in oracle/ko_estimator.py)
__attribute__((visibility("default"))) double ko_apply_dx(ko_problem* h, const double* dx) {
  h->P->dx.assign(dx, dx + h->P->solver->JCols);
  return h->P->applyStateUpdate();
}
__attribute__((visibility("default"))) void ko_optimize(ko_problem* h, const kb_optimizer_options* o, int nThreads, kb_solution* out) {
  h->trace.clear();
  optimize(*h->P, *o, nThreads, out, &h->trace);
}
// per-iteration trace of the last ko_optimize: triples (J, deltaX, lambda)
__attribute__((visibility("default"))) int32_t ko_get_trace(ko_problem* h, double* out, int32_t max_triples) {
  const int32_t n = (int32_t)(h->trace.size() / 3);
  if (out)
    for (int32_t i = 0; i < std::min(n, max_triples) * 3; ++i) out[i] = h->trace[i];
  return n;
}

// J^T in compressed-column form, as CompressedColumnJacobianTransposeBuilder would hold it (unweighted scaling 1, invR = I).
// Call with NULL arrays to obtain nnz.
__attribute__((visibility("default"))) int64_t ko_get_jacobian_ccs(ko_problem* h, int nThreads, int64_t* col_ptr, int32_t* row_idx, double* values) {
  Problem& P = *h->P;
  const size_t n = P.errorTerms.size();
  std::vector<int64_t> cp(2 * n + 1, 0);
  for (size_t i = 0; i < n; ++i) {
    int w = 0;
    const int k = P.termCam[i];
    w += 6 + 6 * k + P.cams[k]->projectionDv->minimalDimensions() + P.cams[k]->distortionDv->minimalDimensions();
    cp[2 * i + 1] = cp[2 * i] + w;
    cp[2 * i + 2] = cp[2 * i + 1] + w;
  }
  if (!col_ptr || !row_idx || !values) return cp.back();
  std::memcpy(col_ptr, cp.data(), sizeof(int64_t) * cp.size());
  setupThreadedJob(
      [&](size_t, size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
          JacobianContainer jc(2);
          P.errorTerms[i]->getWeightedJacobians(jc, P.solver->useMEstimator);
          for (int r = 0; r < 2; ++r) {
            int64_t o = cp[2 * i + r];
            for (auto& kv : jc.jacobianMap)
              for (int c = 0; c < kv.second.c; ++c) {
                row_idx[o] = kv.first->columnBase + c;
                values[o++] = kv.second(r, c);
              }
          }
        }
      },
      std::max(1, nThreads), n);
  return cp.back();
}

// Marginal analysis of the camera-side block at the current state (ko_marginal.hpp).  The set-pose columns are A_l, the
// remaining columns, in the order [camera 0 projection | distortion, camera 1 ..., baseline 0 q | t, ...], are A_r.
// V, columns, omega may be NULL.
__attribute__((visibility("default"))) int32_t ko_analyze_marginal(ko_problem* h, int nThreads, const kb_marginal_options* o, kb_marginal_result* out,
                                                                   double* sv, double* V, int32_t* columns, double* omega) {
  Problem& P = *h->P;
  P.evaluateError(std::max(1, nThreads));  // Jacobians are taken at the state of the last evaluation (quirk Q8)
  const size_t m = 2 * P.errorTerms.size(), nAll = P.solver->JCols;
  std::vector<char> isPose(nAll, 0);
  for (size_t v = 0; v < P.setQ.size(); ++v) {
    for (int c = 0; c < 3; ++c) {
      isPose[P.setQ[v]->columnBase + c] = 1;
      isPose[P.setT[v]->columnBase + c] = 1;
    }
  }
  std::vector<int> colR, colL;
  for (int k = 0; k < P.nCams; ++k) {
    for (int c = 0; c < P.cams[k]->projectionDv->minimalDimensions(); ++c) colR.push_back(P.cams[k]->projectionDv->columnBase + c);
    for (int c = 0; c < P.cams[k]->distortionDv->minimalDimensions(); ++c) colR.push_back(P.cams[k]->distortionDv->columnBase + c);
  }
  for (size_t j = 0; j < P.baseQ.size(); ++j) {
    for (int c = 0; c < 3; ++c) colR.push_back(P.baseQ[j]->columnBase + c);
    for (int c = 0; c < 3; ++c) colR.push_back(P.baseT[j]->columnBase + c);
  }
  std::vector<int> slot(nAll, -1);
  for (size_t i = 0; i < colR.size(); ++i) slot[colR[i]] = (int)i;
  for (size_t c = 0; c < nAll; ++c)
    if (isPose[c]) { slot[c] = (int)colL.size(); colL.push_back((int)c); }
  const int nl = (int)colL.size(), nr = (int)colR.size();
  std::vector<double> Al(m * (size_t)nl, 0.0), Ar(m * (size_t)nr, 0.0);
  for (size_t i = 0; i < P.errorTerms.size(); ++i) {
    JacobianContainer jc(2);
    P.errorTerms[i]->getWeightedJacobians(jc, P.solver->useMEstimator);
    for (int r = 0; r < 2; ++r)
      for (auto& kv : jc.jacobianMap)
        for (int c = 0; c < kv.second.c; ++c) {
          const int col = kv.first->columnBase + c;
          if (isPose[col]) Al[(2 * i + r) * (size_t)nl + slot[col]] = kv.second(r, c);
          else Ar[(2 * i + r) * (size_t)nr + slot[col]] = kv.second(r, c);
        }
  }
  MarginalResult R = analyzeMarginalDense(Al, (int)m, nl, Ar, nr, o->eps_svd, o->svd_tol);
  out->n = R.n;
  out->rank = R.rank;
  out->rank_deficiency = R.rankDeficiency;
  out->tolerance = R.tolerance;
  out->sv_log2_sum = R.svLog2Sum;
  out->sv_gap = R.svGap;
  std::memcpy(sv, R.singularValues.data(), sizeof(double) * nr);
  if (V) std::memcpy(V, R.V.data(), sizeof(double) * nr * nr);
  if (omega) std::memcpy(omega, R.Omega.data(), sizeof(double) * nr * nr);
  if (columns)
    for (int i = 0; i < nr; ++i) columns[i] = colR[i];
  return nr;
}

// Upper-triangular block pattern + values of H as SparseBlockMatrix holds them (BlockCholesky semantic only).
__attribute__((visibility("default"))) int32_t ko_get_hessian_blocks(ko_problem* h, int64_t* n_blocks, int64_t* n_values, int64_t* col_ptr,
                                                                     int32_t* block_row, int64_t* value_ptr, double* values) {
  SparseBlockMatrix* H = nullptr;
  if (auto* b = dynamic_cast<BlockCholeskyLinearSystemSolver*>(h->P->solver.get())) H = &b->H;
  if (auto* s = dynamic_cast<SparseCholeskyLinearSystemSolver*>(h->P->solver.get())) H = &s->H;
  if (!H) return -1;
  int64_t nb = 0, nv = 0;
  for (int c = 0; c < H->bRows(); ++c) {
    if (col_ptr) col_ptr[c] = nb;
    for (auto& kv : H->blockCols[c]) {
      if (block_row) block_row[nb] = kv.first;
      if (value_ptr) value_ptr[nb] = nv;
      if (values) std::memcpy(values + nv, kv.second->d.data(), sizeof(double) * kv.second->d.size());
      nv += (int64_t)kv.second->d.size();
      ++nb;
    }
  }
  if (col_ptr) col_ptr[H->bRows()] = nb;
  if (n_blocks) *n_blocks = nb;
  if (n_values) *n_values = nv;
  return 0;
}

__attribute__((visibility("default"))) void ko_get_camera_params(ko_problem* h, double* out) {
  for (int k = 0; k < h->P->nCams; ++k) {
    std::vector<double> p, d;
    h->P->cams[k]->projectionDv->getParameters(p);
    h->P->cams[k]->distortionDv->getParameters(d);
    double* o = out + (size_t)k * KB_CAM_PARAM_STRIDE;
    std::fill(o, o + KB_CAM_PARAM_STRIDE, 0.0);
    std::copy(p.begin(), p.end(), o);
    std::copy(d.begin(), d.end(), o + p.size());
  }
}
__attribute__((visibility("default"))) void ko_get_baselines(ko_problem* h, double* out) {
  for (size_t k = 0; k < h->P->baseQ.size(); ++k) {
    std::memcpy(out + 7 * k, h->P->baseQ[k]->q, 4 * sizeof(double));
    std::memcpy(out + 7 * k + 4, h->P->baseT[k]->p, 3 * sizeof(double));
  }
}
__attribute__((visibility("default"))) void ko_get_set_poses(ko_problem* h, double* out) {
  for (size_t k = 0; k < h->P->setQ.size(); ++k) {
    std::memcpy(out + 7 * k, h->P->setQ[k]->q, 4 * sizeof(double));
    std::memcpy(out + 7 * k + 4, h->P->setT[k]->p, 3 * sizeof(double));
  }
}

// One camera-model evaluation at a homogeneous point (for the CameraGeometryTestHarness-style property tests).
// Jp: 2x4 row-major, Ji: 2x6 row-major (first P columns used), Jd: 2x4 row-major (first D columns used).
__attribute__((visibility("default"))) int32_t ko_camera_project(int32_t model, const double* params, const double* ph, double* y, double* Jp,
                                                                  double* Ji, double* Jd) {
  auto cam = makeCamera(model, params);
  if (!cam) return -1;
  double yy[2] = {0, 0};
  Mat J;
  const bool ok = cam->homogeneousToKeypoint(ph, yy, J);
  double y2[2] = {0, 0};
  cam->homogeneousToKeypoint(ph, y2);
  y[0] = y2[0];
  y[1] = y2[1];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 4; ++j) Jp[i * 4 + j] = J(i, j);
  Mat I, D;
  cam->homogeneousToKeypointIntrinsicsJacobian(ph, I);
  cam->homogeneousToKeypointDistortionJacobian(ph, D);
  std::fill(Ji, Ji + 12, 0.0);
  std::fill(Jd, Jd + 8, 0.0);
  for (int i = 0; i < 2; ++i) {
    for (int j = 0; j < I.c; ++j) Ji[i * 6 + j] = I(i, j);
    for (int j = 0; j < D.c; ++j) Jd[i * 4 + j] = D(i, j);
  }
  return ok ? 1 : 0;
}

// sm_kinematics spot checks
__attribute__((visibility("default"))) void ko_quat2r(const double* q, double* R9) {
  Mat R = quat2r(q);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R(i, j);
}
__attribute__((visibility("default"))) void ko_update_quat(const double* q, const double* dq, double* out) { updateQuat(q, dq, out); }
__attribute__((visibility("default"))) void ko_box_minus(const double* p4, double* out24_rowmajor) {
  Mat B = boxMinus(p4);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 6; ++j) out24_rowmajor[i * 6 + j] = B(i, j);
}
__attribute__((visibility("default"))) void ko_box_times(const double* T16_rowmajor, double* out36_rowmajor) {
  Mat T(4, 4);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) T(i, j) = T16_rowmajor[i * 4 + j];
  Mat B = boxTimes(T);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) out36_rowmajor[i * 6 + j] = B(i, j);
}
__attribute__((visibility("default"))) void ko_inverse4(const double* M16_rowmajor, double* out16_rowmajor) {
  Mat M(4, 4);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) M(i, j) = M16_rowmajor[i * 4 + j];
  Mat I = inverse4(M);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out16_rowmajor[i * 4 + j] = I(i, j);
}

// ---- weighting: ErrorTermFs<2>::setInvR / ErrorTerm::setMEstimatorPolicy on every term -------------------------------------
__attribute__((visibility("default"))) void ko_set_inv_r(ko_problem* h, const double* invR_rowmajor) {
  Mat A(2, 2);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) A(i, j) = invR_rowmajor[i * 2 + j];
  for (auto* t : h->P->errorTerms) t->setInvR(A);
}
// kind = kb_m_estimator; parameters as kb_set_m_estimator.  Returns the policy's derived parameter (epsilon for Blake-Zisserman).
__attribute__((visibility("default"))) double ko_set_m_estimator(ko_problem* h, int kind, double p0, double p1, double p2) {
  std::shared_ptr<MEstimator> m;
  double prm = p0;
  switch (kind) {
    case 1: m = std::make_shared<HuberMEstimator>(p0); break;
    case 2: m = std::make_shared<CauchyMEstimator>(p0); break;
    case 3: m = std::make_shared<GemanMcClureMEstimator>(p0); break;
    case 4: {
      auto bz = std::make_shared<BlakeZissermanMEstimator>((size_t)p0, p1, p2);
      prm = bz->epsilon;
      m = bz;
      break;
    }
    default: m = std::make_shared<NoMEstimator>(); prm = 0.0;
  }
  for (auto* t : h->P->errorTerms) t->setMEstimatorPolicy(m);  // one shared policy object, as kalibr's Python tooling did
  return prm;
}
__attribute__((visibility("default"))) void ko_set_use_m_estimator(ko_problem* h, int on) { h->P->solver->useMEstimator = on != 0; }
__attribute__((visibility("default"))) double ko_chi2_inv_cdf(double p, int df) { return chi2InvCDF(p, (size_t)df); }
__attribute__((visibility("default"))) double ko_m_estimator_weight(int kind, double p0, double p1, double p2, double squaredError) {
  switch (kind) {
    case 1: return HuberMEstimator(p0).getWeight(squaredError);
    case 2: return CauchyMEstimator(p0).getWeight(squaredError);
    case 3: return GemanMcClureMEstimator(p0).getWeight(squaredError);
    case 4: return BlakeZissermanMEstimator((size_t)p0, p1, p2).getWeight(squaredError);
  }
  return NoMEstimator().getWeight(squaredError);
}
__attribute__((visibility("default"))) void ko_matrix_sqrt2(const double* A_rowmajor, double* S_rowmajor) {
  Mat A(2, 2);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) A(i, j) = A_rowmajor[i * 2 + j];
  Mat S = computeMatrixSqrt2(A);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) S_rowmajor[i * 2 + j] = S(i, j);
}

// ---- K2/include/kalibr2/CameraCalibrator.hpp:267-286, 368-405: PrintReprojectionErrorStatistics per camera ------------------
// out[cam] = {n, mean_u, mean_v, std_u, std_v, rmse}; error values = getMeasurement() - getPredictedMeasurement() at the current state
__attribute__((visibility("default"))) void ko_reprojection_statistics(ko_problem* h, double* out) {
  Problem& P = *h->P;
  for (int k = 0; k < P.nCams; ++k) {
    std::vector<std::array<double, 2>> error_values;
    for (size_t i = 0; i < P.errorTerms.size(); ++i) {
      if (P.termCam[i] != k) continue;
      ReprojectionError* t = P.errorTerms[i];
      double p4[4], hat_y[2] = {0.0, 0.0};
      t->point.toHomogeneous(p4);
      t->camera->camera->homogeneousToKeypoint(p4, hat_y);
      error_values.push_back({t->y[0] - hat_y[0], t->y[1] - hat_y[1]});
    }
    double* o = out + 6 * k;
    for (int i = 0; i < 6; ++i) o[i] = 0.0;
    const double n = (double)error_values.size();
    o[0] = n;
    if (error_values.empty()) continue;
    double sum[2] = {0.0, 0.0};
    for (auto& e : error_values) { sum[0] += e[0]; sum[1] += e[1]; }
    const double mean[2] = {sum[0] / n, sum[1] / n};
    double ssd[2] = {0.0, 0.0};
    if (error_values.size() > 1) {
      for (auto& e : error_values) {
        ssd[0] += (e[0] - mean[0]) * (e[0] - mean[0]);
        ssd[1] += (e[1] - mean[1]) * (e[1] - mean[1]);
      }
      o[3] = std::sqrt(ssd[0] / (n - 1.0));
      o[4] = std::sqrt(ssd[1] / (n - 1.0));
    }
    o[1] = mean[0];
    o[2] = mean[1];
    o[5] = std::sqrt(sum[0] * sum[0] + sum[1] * sum[1]) / std::sqrt(n);  // sum_of_errors.norm() / sqrt(size), as printed
  }
}

// CPU baseline timing: one LM-iteration's worth of hot path (evaluate + build + solve) at the current state,
// wall-clock seconds per stage.  out[0]=evaluate, out[1]=build(linearise+assemble), out[2]=solve.
__attribute__((visibility("default"))) int32_t ko_time_iteration(ko_problem* h, int nThreads, double lambda, double* out) {
  using clk = std::chrono::steady_clock;
  auto t0 = clk::now();
  h->P->evaluateError(nThreads);
  auto t1 = clk::now();
  h->P->solver->buildSystem(nThreads);
  auto t2 = clk::now();
  h->P->solver->setConstantConditioner(lambda);
  const bool ok = h->P->solver->solveSystem(h->P->dx);
  auto t3 = clk::now();
  out[0] = std::chrono::duration<double>(t1 - t0).count();
  out[1] = std::chrono::duration<double>(t2 - t1).count();
  out[2] = std::chrono::duration<double>(t3 - t2).count();
  return ok ? 1 : 0;
}

}  // extern "C"
