// TEST INFRASTRUCTURE: the REFERENCE's optimiser on a small calibration problem - aslam_backend's Optimizer2, LevenbergMarquardt /
// TrustRegionPolicy, LinearSystemSolver, BlockCholeskyLinearSystemSolver (build, conditioner, the lambda^2 / lambda damping of SURVEY.md
// Q2), ErrorTermFs (weighting, buildHessian), JacobianContainer::evaluateHessian, OptimizationProblem and the sparse_block_matrix
// container, all compiled from the sources where they lie under /root/reference against the stand-in headers of oracle/ref_shim/.
// NOT reference code in this translation unit, and said so where it stands: (1) the factorisation behind LinearSolverCholmod (CHOLMOD is
// not in the image: a dense Cholesky over the reference's own SparseBlockMatrix, ref_shim/sparse_block_matrix/linear_solver_cholmod.h);
// (2) the glue below - the error term, which restates CVE/.../implementation/ReprojectionError.hpp:50-77 and the two active-block lines of
// CVB/.../implementation/CameraDesignVariable.hpp:39-54 line by line over the reference's projection classes and expression tree (the
// reference's own templates need the CameraGeometry / Frame / Image headers, i.e. OpenCV).  The camera design variables themselves ARE the
// reference's: its DesignVariableAdapter over its projection and distortion objects, created as CameraDesignVariable creates them.
// The per-term arithmetic the glue stands for is pinned separately against the real classes (ref_pin.cpp).  What this file pins is the
// LOOP: iteration and failed-iteration counts, the cost per iteration, the lambda schedule, the damping quirk, the final parameters.
// The SparseCholesky regime (Kalibr2's default: Optimizer2.cpp:83-86) is here as well: SparseCholeskyLinearSystemSolver.cpp,
// CompressedColumnJacobianTransposeBuilder, CompressedColumnMatrix and the reference's Cholmod wrapper compile from their sources; the
// CHOLMOD entry points that wrapper calls are the dense extended-precision stand-in of ref_shim/cholmod.h (ref_sparse_system,
// ref_optimize_rig_solver, ref_time_evaluate_build_solver below).
#include <aslam/cameras/EquidistantDistortion.hpp>
#include <aslam/cameras/FovDistortion.hpp>
#include <aslam/cameras/NoDistortion.hpp>
#include <aslam/cameras/RadialTangentialDistortion.hpp>
#include <aslam/cameras/OmniProjection.hpp>
#include <aslam/cameras/PinholeProjection.hpp>
#include <aslam/cameras/DoubleSphereProjection.hpp>
#include <aslam/cameras/ExtendedUnifiedProjection.hpp>

#include <aslam/backend/BlockCholeskyLinearSystemSolver.hpp>
#include <aslam/backend/DesignVariableAdapter.hpp>
#include <aslam/backend/CompressedColumnJacobianTransposeBuilder.hpp>
#include <aslam/backend/SparseCholeskyLinearSystemSolver.hpp>
#include <aslam/backend/ErrorTerm.hpp>
#include <aslam/backend/EuclideanPoint.hpp>
#include <aslam/backend/HomogeneousExpression.hpp>
#include <aslam/backend/HomogeneousPoint.hpp>
#include <aslam/backend/GaussNewtonTrustRegionPolicy.hpp>
#include <aslam/backend/LevenbergMarquardtTrustRegionPolicy.hpp>
#include <aslam/backend/MEstimatorPolicies.hpp>
#include <aslam/backend/OptimizationProblem.hpp>
#include <aslam/backend/Optimizer2.hpp>
#include <aslam/backend/RotationQuaternion.hpp>
#include <aslam/backend/TransformationBasic.hpp>
#include <aslam/backend/TransformationExpression.hpp>

#include <aslam/calibration/core/IncrementalOptimizationProblem.h>
#include <aslam/calibration/core/OptimizationProblem.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <cstdint>
#include <vector>

using namespace aslam::backend;
using namespace aslam::cameras;

namespace {
const int N_P[7] = {4, 4, 5, 6, 6, 4, 5}, N_D[7] = {4, 4, 4, 0, 0, 1, 0};

// One camera: the reference's own projection object (its distortion object inside it) and the reference's own DesignVariableAdapter over
// each of the two, created as CVB/.../implementation/CameraDesignVariable.hpp:8-10 creates them (adapters that do not own the objects) -
// so update (backup + Projection::update / Distortion::update), revert (setParameters(backup)), getParameters and the 0-dimensional
// active distortion block of the models without distortion (quirk Q7) are the reference's code (BE/.../implementation/DesignVariableAdapter.hpp).
struct CameraModel {
  int model, P, D;
  boost::shared_ptr<DesignVariable> proj, dist;
  virtual ~CameraModel() {}
  virtual bool project(const Eigen::Vector4d& ph, Eigen::VectorXd& y, Eigen::MatrixXd* J, Eigen::MatrixXd* Ji, Eigen::MatrixXd* Jd) const = 0;
  virtual void parameters(double* prm) const = 0;  // [P projection parameters, D distortion parameters]
};

template <typename PROJECTION>
struct CameraModelOf : public CameraModel {
  typedef typename PROJECTION::distortion_t distortion_t;
  PROJECTION projection;
  CameraModelOf(int model_, const PROJECTION& p) : projection(p) {
    model = model_;
    P = N_P[model];
    D = N_D[model];
    proj.reset(new DesignVariableAdapter<PROJECTION>(&projection, false));
    dist.reset(new DesignVariableAdapter<distortion_t>(&projection.distortion(), false));
  }
  virtual bool project(const Eigen::Vector4d& ph, Eigen::VectorXd& y, Eigen::MatrixXd* J, Eigen::MatrixXd* Ji, Eigen::MatrixXd* Jd) const {
    if (!J) return projection.homogeneousToKeypoint(ph, y);
    const bool ok = projection.homogeneousToKeypoint(ph, y, *J);
    projection.homogeneousToKeypointIntrinsicsJacobian(ph, *Ji);
    projection.homogeneousToKeypointDistortionJacobian(ph, *Jd);
    return ok;
  }
  virtual void parameters(double* prm) const {
    Eigen::MatrixXd a, b;
    projection.getParameters(a);
    projection.distortion().getParameters(b);
    for (int i = 0; i < P; ++i) prm[i] = a(i, 0);
    for (int i = 0; i < D; ++i) prm[P + i] = b(i, 0);
  }
};

template <typename PROJECTION>
boost::shared_ptr<CameraModel> makeCameraOf(int model, const PROJECTION& p) { return boost::shared_ptr<CameraModel>(new CameraModelOf<PROJECTION>(model, p)); }

boost::shared_ptr<CameraModel> makeCamera(int model, const double* p) {
  const int ru = 1 << 20, rv = 1 << 20;
  switch (model) {
    case 0: return makeCameraOf(model, PinholeProjection<RadialTangentialDistortion>(p[0], p[1], p[2], p[3], ru, rv, RadialTangentialDistortion(p[4], p[5], p[6], p[7])));
    case 1: return makeCameraOf(model, PinholeProjection<EquidistantDistortion>(p[0], p[1], p[2], p[3], ru, rv, EquidistantDistortion(p[4], p[5], p[6], p[7])));
    case 2: return makeCameraOf(model, OmniProjection<RadialTangentialDistortion>(p[0], p[1], p[2], p[3], p[4], ru, rv, RadialTangentialDistortion(p[5], p[6], p[7], p[8])));
    case 3: return makeCameraOf(model, ExtendedUnifiedProjection<NoDistortion>(p[0], p[1], p[2], p[3], p[4], p[5], ru, rv));
    case 4: return makeCameraOf(model, DoubleSphereProjection<NoDistortion>(p[0], p[1], p[2], p[3], p[4], p[5], ru, rv));
    case 5: return makeCameraOf(model, PinholeProjection<FovDistortion>(p[0], p[1], p[2], p[3], ru, rv, FovDistortion(p[4])));
    case 6: return makeCameraOf(model, OmniProjection<NoDistortion>(p[0], p[1], p[2], p[3], p[4], ru, rv));
  }
  throw std::runtime_error("unknown camera model");
}

// The weighting every term of the problems built after ref_set_weighting gets: invR through ErrorTermFs<2>::setInvR (K2 passes I in the
// batch drivers, CalibrationTools.hpp:105-108, and I / sigma^2 in CreateBatchProblem, :495-496) and an M-estimator policy through
// ErrorTerm::setMEstimatorPolicy (BE/src/MEstimatorPolicies.cpp; kind as kb_m_estimator, see ref_m_estimator_weight in ref_pin.cpp).
// setInvR takes the matrix square root through sm::eigen::computeMatrixSqrt = Eigen::LDLT, a STAND-IN in this build
// (ref_shim/sm/eigen/matrix_sqrt.hpp); for invR = c I - the only form Kalibr2 passes - that square root is sqrt(c) I with no pivoting.
struct Weighting {
  double inv_r[4] = {1.0, 0.0, 0.0, 1.0};
  int kind = 0;
  double p0 = 0.0, p1 = 0.999, p2 = 0.1;
} g_weighting;

boost::shared_ptr<MEstimator> makePolicy(const Weighting& w) {
  switch (w.kind) {
    case 1: return boost::shared_ptr<MEstimator>(new HuberMEstimator(w.p0));
    case 2: return boost::shared_ptr<MEstimator>(new CauchyMEstimator(w.p0));
    case 3: return boost::shared_ptr<MEstimator>(new GemanMcClureMEstimator(w.p0));
    case 4: return boost::shared_ptr<MEstimator>(new BlakeZissermanMEstimator((size_t)w.p0, w.p1, w.p2));
  }
  return boost::shared_ptr<MEstimator>();
}

// CVE/.../implementation/ReprojectionError.hpp:50-77 over the classes above
class ReprojectionTerm : public ErrorTermFs<2> {
 public:
  ReprojectionTerm(const Eigen::Vector2d& y, const HomogeneousExpression& point, CameraModel* cam, const boost::shared_ptr<MEstimator>& policy)
      : y_(y), point_(point), cam_(cam) {
    Eigen::Matrix2d invR;
    invR(0, 0) = g_weighting.inv_r[0]; invR(0, 1) = g_weighting.inv_r[1]; invR(1, 0) = g_weighting.inv_r[2]; invR(1, 1) = g_weighting.inv_r[3];
    setInvR(invR);
    if (policy) setMEstimatorPolicy(policy);
    DesignVariable::set_t dvs;
    point_.getDesignVariables(dvs);
    dvs.insert(cam_->proj.get());
    dvs.insert(cam_->dist.get());
    setDesignVariablesIterator(dvs.begin(), dvs.end());
  }
 protected:
  virtual double evaluateErrorImplementation() {
    const Eigen::Vector4d p = point_.toHomogeneous();
    Eigen::VectorXd hat_y(2);
    hat_y.setZero();
    cam_->project(p, hat_y, 0, 0, 0);
    setError(y_ - hat_y);
    return error().dot(invR() * error());
  }
  virtual void evaluateJacobiansImplementation(JacobianContainer& jacobians) const {
    const Eigen::Vector4d p = point_.toHomogeneous();
    Eigen::VectorXd hat_y(2);
    Eigen::MatrixXd J(2, 4), Ji, Jd;
    J.setZero();
    cam_->project(p, hat_y, &J, &Ji, &Jd);
    point_.evaluateJacobians(jacobians, -J);
    // CVB/.../implementation/CameraDesignVariable.hpp: the negated intrinsics / distortion Jacobians, for the active blocks
    if (cam_->proj->isActive()) jacobians.add(cam_->proj.get(), Eigen::MatrixXd(-Ji));
    if (cam_->dist->isActive() && cam_->D > 0) jacobians.add(cam_->dist.get(), Eigen::MatrixXd(-Jd));
  }
 private:
  Eigen::Vector2d y_;
  HomogeneousExpression point_;
  CameraModel* cam_;
};
}  // namespace

// Problem in the layout of include/kalibr_b200.h (kb_problem_desc); design variables are added in the order of the driver named by
// driver_order (kb_driver_order: 0/2 = cameras, baselines, then one pose per synced set, K2/CalibrationTools.hpp:183-300, 375-408;
// 1 = baselines, poses, cameras, :222-262; 3 = poses, baselines, cameras, :460-491); error terms in term order.
namespace {
struct RigProblem {
  boost::shared_ptr<OptimizationProblem> problem;
  std::vector<boost::shared_ptr<CameraModel>> cams;
  std::vector<boost::shared_ptr<RotationQuaternion>> bq, sq;
  std::vector<boost::shared_ptr<EuclideanPoint>> bt, st;
  std::vector<boost::shared_ptr<TransformationBasic>> B, S;
  std::vector<boost::shared_ptr<HomogeneousPoint>> points;

  boost::shared_ptr<TransformationBasic> addPose(const double* p, std::vector<boost::shared_ptr<RotationQuaternion>>& qs, std::vector<boost::shared_ptr<EuclideanPoint>>& ts) {
    qs.push_back(boost::make_shared<RotationQuaternion>(Eigen::Vector4d(p[0], p[1], p[2], p[3])));
    qs.back()->setActive(true);
    problem->addDesignVariable(qs.back());
    ts.push_back(boost::make_shared<EuclideanPoint>(Eigen::Vector3d(p[4], p[5], p[6])));
    ts.back()->setActive(true);
    problem->addDesignVariable(ts.back());
    return boost::make_shared<TransformationBasic>(qs.back()->toExpression(), ts.back()->toExpression());
  }

  RigProblem(int32_t n_cams, const int32_t* cam_model, const double* cam_params, const double* baselines, int32_t n_sets, const double* set_poses, int32_t n_target,
             const double* target, int32_t n_views, const int32_t* view_set, const int32_t* view_cam, const int64_t* view_begin, const double* y_u, const double* y_v,
             const int32_t* corner_id, int32_t driver_order)
      : problem(new OptimizationProblem()), cams(n_cams) {
    auto addCameras = [&]() {
      for (int k = 0; k < n_cams; ++k) {
        cams[k] = makeCamera(cam_model[k], cam_params + k * 10);
        cams[k]->proj->setActive(true);  // K2/CameraCalibrator.hpp:116-122: setActive(true, true, false)
        cams[k]->dist->setActive(true);
        problem->addDesignVariable(cams[k]->proj);
        problem->addDesignVariable(cams[k]->dist);
      }
    };
    auto addBaselines = [&]() { for (int j = 0; j + 1 < n_cams; ++j) B.push_back(addPose(baselines + 7 * j, bq, bt)); };
    auto addSets = [&]() { for (int v = 0; v < n_sets; ++v) S.push_back(addPose(set_poses + 7 * v, sq, st)); };
    if (driver_order == 1) { addBaselines(); addSets(); addCameras(); }
    else if (driver_order == 3) { addSets(); addBaselines(); addCameras(); }
    else { addCameras(); addBaselines(); addSets(); }
    for (int i = 0; i < n_target; ++i) points.push_back(boost::make_shared<HomogeneousPoint>(Eigen::Vector4d(target[3 * i], target[3 * i + 1], target[3 * i + 2], 1.0)));
    const boost::shared_ptr<MEstimator> policy = makePolicy(g_weighting);  // one policy object shared by all terms, as a driver would set it
    for (int w = 0; w < n_views; ++w) {
      TransformationExpression T_cam_w = S[view_set[w]]->toExpression().inverse();
      for (int j = 0; j < view_cam[w]; ++j) T_cam_w = B[j]->toExpression() * T_cam_w;
      for (int64_t i = view_begin[w]; i < view_begin[w + 1]; ++i)
        problem->addErrorTerm(boost::make_shared<ReprojectionTerm>(Eigen::Vector2d(y_u[i], y_v[i]), T_cam_w * points[corner_id[i]]->toExpression(), cams[view_cam[w]].get(), policy));
    }
  }
};
}  // namespace

// 0 = LevenbergMarquardtTrustRegionPolicy(lambda_init) (the batch drivers, K2/CalibrationTools.hpp:57-66), 1 = GaussNewtonTrustRegionPolicy
// (what the incremental estimator's Optimizer2 runs: build + undamped solve every iteration, never revert); applies to the optimisations
// started afterwards
int g_trust_region_policy = 0;
extern "C" __attribute__((visibility("default"))) void ref_set_trust_region_policy(int32_t kind) { g_trust_region_policy = kind; }

// inv_r: 2x2 row-major or NULL for the identity; kind 0 removes the policy.  Applies to every problem built afterwards.
extern "C" __attribute__((visibility("default"))) void ref_set_weighting(const double* inv_r, int32_t kind, double p0, double p1, double p2) {
  const double identity[4] = {1.0, 0.0, 0.0, 1.0};
  for (int i = 0; i < 4; ++i) g_weighting.inv_r[i] = inv_r ? inv_r[i] : identity[i];
  g_weighting.kind = kind;
  g_weighting.p0 = p0;
  g_weighting.p1 = p1;
  g_weighting.p2 = p2;
}

// out_scalars: [iterations, failedIterations, JStart, JFinal, linearSolverFailure]; state arrays are updated in place.
// solver_kind: 0 = BlockCholeskyLinearSystemSolver, 1 = SparseCholeskyLinearSystemSolver (Kalibr2's default: Optimizer2.cpp:83-86; the
// reference's own SparseCholeskyLinearSystemSolver.cpp, CompressedColumnJacobianTransposeBuilder, CompressedColumnMatrix and Cholmod
// wrapper over the dense stand-in of ref_shim/cholmod.h).  n_threads: Optimizer2Options::nThreads (error evaluation and, for the sparse
// solver, the Jacobian materialisation are threaded over it).
static int32_t optimize_rig(int32_t n_cams, const int32_t* cam_model, double* cam_params, double* baselines, int32_t n_sets, double* set_poses, int32_t n_target,
                            const double* target, int32_t n_views, const int32_t* view_set, const int32_t* view_cam, const int64_t* view_begin, const double* y_u,
                            const double* y_v, const int32_t* corner_id, int32_t driver_order, int32_t max_iterations, double conv_dx, double conv_dj,
                            double lambda_init, int32_t solver_kind, int32_t n_threads, double* out_scalars) {
  try {
    RigProblem rp(n_cams, cam_model, cam_params, baselines, n_sets, set_poses, n_target, target, n_views, view_set, view_cam, view_begin, y_u, y_v, corner_id, driver_order);
    Optimizer2Options options;  // K2/CalibrationTools.hpp:57-66
    options.nThreads = n_threads;
    options.convergenceDeltaX = conv_dx;
    options.convergenceDeltaJ = conv_dj;
    options.maxIterations = max_iterations;
    if (g_trust_region_policy == 1) options.trustRegionPolicy = boost::make_shared<GaussNewtonTrustRegionPolicy>();
    else options.trustRegionPolicy = boost::make_shared<LevenbergMarquardtTrustRegionPolicy>(lambda_init);
    if (solver_kind == 1) options.linearSystemSolver = boost::make_shared<SparseCholeskyLinearSystemSolver>();
    else options.linearSystemSolver = boost::make_shared<BlockCholeskyLinearSystemSolver>();
    Optimizer2 optimizer(options);
    optimizer.setProblem(rp.problem);
    SolutionReturnValue r = optimizer.optimize();
    out_scalars[0] = r.iterations;
    out_scalars[1] = r.failedIterations;
    out_scalars[2] = r.JStart;
    out_scalars[3] = r.JFinal;
    out_scalars[4] = r.linearSolverFailure ? 1.0 : 0.0;
    for (int k = 0; k < n_cams; ++k) rp.cams[k]->parameters(cam_params + k * 10);
    auto store = [](double* p, const boost::shared_ptr<RotationQuaternion>& q, const boost::shared_ptr<EuclideanPoint>& t) {
      const Eigen::Vector4d qv = q->getQuaternion();
      const Eigen::Vector3d tv = t->toEuclidean();
      for (int i = 0; i < 4; ++i) p[i] = qv(i);
      for (int i = 0; i < 3; ++i) p[4 + i] = tv(i);
    };
    for (size_t j = 0; j < rp.bq.size(); ++j) store(baselines + 7 * j, rp.bq[j], rp.bt[j]);
    for (size_t v = 0; v < rp.sq.size(); ++v) store(set_poses + 7 * v, rp.sq[v], rp.st[v]);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_optimize_rig: " << e.what() << std::endl;
    return -1;
  }
}

extern "C" __attribute__((visibility("default"))) int32_t ref_optimize_rig(int32_t n_cams, const int32_t* cam_model, double* cam_params /*[n_cams][10]*/,
                                                                           double* baselines /*[n_cams-1][7]*/, int32_t n_sets, double* set_poses /*[n_sets][7]*/,
                                                                           int32_t n_target, const double* target /*[n_target][3]*/, int32_t n_views,
                                                                           const int32_t* view_set, const int32_t* view_cam, const int64_t* view_begin,
                                                                           const double* y_u, const double* y_v, const int32_t* corner_id, int32_t driver_order, int32_t max_iterations,
                                                                           double conv_dx, double conv_dj, double lambda_init, double* out_scalars) {
  return optimize_rig(n_cams, cam_model, cam_params, baselines, n_sets, set_poses, n_target, target, n_views, view_set, view_cam, view_begin, y_u, y_v, corner_id,
                      driver_order, max_iterations, conv_dx, conv_dj, lambda_init, 0, 1, out_scalars);
}

extern "C" __attribute__((visibility("default"))) int32_t ref_optimize_rig_solver(int32_t n_cams, const int32_t* cam_model, double* cam_params, double* baselines,
                                                                                  int32_t n_sets, double* set_poses, int32_t n_target, const double* target,
                                                                                  int32_t n_views, const int32_t* view_set, const int32_t* view_cam,
                                                                                  const int64_t* view_begin, const double* y_u, const double* y_v,
                                                                                  const int32_t* corner_id, int32_t driver_order, int32_t max_iterations, double conv_dx,
                                                                                  double conv_dj, double lambda_init, int32_t solver_kind, int32_t n_threads,
                                                                                  double* out_scalars) {
  return optimize_rig(n_cams, cam_model, cam_params, baselines, n_sets, set_poses, n_target, target, n_views, view_set, view_cam, view_begin, y_u, y_v, corner_id,
                      driver_order, max_iterations, conv_dx, conv_dj, lambda_init, solver_kind, n_threads, out_scalars);
}

// The SparseCholesky regime's linear system as the reference's own classes hold it: J^T in compressed-column form
// (CompressedColumnJacobianTransposeBuilder<int> -> CompressedColumnMatrix<int>: one column per residual row, the row indices of a
// column in the order of the term's design variables, BE/.../implementation/CompressedColumnMatrix.hpp), the weighted error vector e
// (LinearSystemSolver::evaluateError), rhs = J^T e exactly as SparseCholeskyLinearSystemSolver::buildSystem forms it
// (CompressedColumnMatrix::rightMultiply), and dx of one solveSystem with the constant conditioner `lambda` (the damping block is pushed
// as extra columns of J^T, so CHOLMOD - here the dense stand-in - factorises J^T J + lambda^2 I).
// out_sizes: [rows of J^T (= columns of J), columns of J^T (= residual rows), nnz, solve succeeded]; with col_ptr == NULL only the sizes
// are written.  col_ptr has columns + 1 entries.
extern "C" __attribute__((visibility("default"))) int32_t ref_sparse_system(int32_t n_cams, const int32_t* cam_model, const double* cam_params, const double* baselines,
                                                                            int32_t n_sets, const double* set_poses, int32_t n_target, const double* target, int32_t n_views,
                                                                            const int32_t* view_set, const int32_t* view_cam, const int64_t* view_begin, const double* y_u,
                                                                            const double* y_v, const int32_t* corner_id, int32_t driver_order, int32_t n_threads,
                                                                            double lambda, int64_t* out_sizes, int64_t* col_ptr, int32_t* row_ind, double* values,
                                                                            double* e, double* rhs, double* dx, double* cost) {
  try {
    RigProblem rp(n_cams, cam_model, cam_params, baselines, n_sets, set_poses, n_target, target, n_views, view_set, view_cam, view_begin, y_u, y_v, corner_id, driver_order);
    Optimizer2Options options;
    options.nThreads = 1;
    options.trustRegionPolicy = boost::make_shared<LevenbergMarquardtTrustRegionPolicy>(lambda);
    options.linearSystemSolver = boost::make_shared<SparseCholeskyLinearSystemSolver>();
    Optimizer2 optimizer(options);
    optimizer.setProblem(rp.problem);
    optimizer.initialize();  // block indices, column bases, row bases; the solver's initMatrixStructure
    optimizer.evaluateError(true);  // one serial pass first (see time_evaluate_build below), then the threaded ones
    optimizer.getSolver<SparseCholeskyLinearSystemSolver>()->buildSystem(1, true);
    optimizer.options().nThreads = n_threads;
    const double J = optimizer.evaluateError(true);
    SparseCholeskyLinearSystemSolver* solver = optimizer.getSolver<SparseCholeskyLinearSystemSolver>();
    solver->buildSystem(n_threads, true);
    // the same builder class the solver holds privately, on the optimiser's own design-variable and error-term lists
    std::vector<DesignVariable*> dvs;
    for (size_t i = 0; i < optimizer.numDesignVariables(); ++i) dvs.push_back(optimizer.designVariable(i));
    std::vector<ErrorTerm*> errors;
    for (size_t i = 0; i < rp.problem->numErrorTerms(); ++i) errors.push_back(rp.problem->errorTerm(i));
    CompressedColumnJacobianTransposeBuilder<int> builder;
    builder.initMatrixStructure(dvs, errors);
    builder.buildSystem(n_threads, true);
    const CompressedColumnMatrix<int>& Jt = builder.J_transpose();
    out_sizes[0] = Jt.rows();
    out_sizes[1] = Jt.cols();
    out_sizes[2] = Jt.nnz();
    out_sizes[3] = 0;
    if (!col_ptr) return 0;
    for (size_t c = 0; c <= Jt.cols(); ++c) col_ptr[c] = Jt.col_ptr()[c];
    for (size_t k = 0; k < Jt.nnz(); ++k) { row_ind[k] = Jt.row_ind()[k]; values[k] = Jt.values()[k]; }
    const Eigen::VectorXd& ev = solver->e();
    for (int i = 0; i < ev.size(); ++i) e[i] = ev[i];
    const Eigen::VectorXd& rv = solver->rhs();
    for (int i = 0; i < rv.size(); ++i) rhs[i] = rv[i];
    solver->setConstantConditioner(lambda);
    Eigen::VectorXd x;
    const bool ok = solver->solveSystem(x);
    out_sizes[3] = ok ? 1 : 0;
    if (ok) for (int i = 0; i < x.size(); ++i) dx[i] = x[i];
    *cost = J;
    return 0;
  } catch (const std::exception& ex) {
    std::cerr << "ref_sparse_system: " << ex.what() << std::endl;
    return -1;
  }
}

// Timing of the reference's own evaluate + build on the same problem (bench.py's cpu_baseline, kind "reference"): Optimizer2::initialize,
// then `repeats` x { Optimizer2::evaluateError(true) (threaded over nThreads: Optimizer2.cpp:300-345), the solver's buildSystem }.
// solver_kind 0: BlockCholeskyLinearSystemSolver::buildSystem (serial in the reference: BlockCholeskyLinearSystemSolver.cpp:57-72);
// solver_kind 1: SparseCholeskyLinearSystemSolver::buildSystem - Kalibr2's default - the Jacobian materialisation into the
// compressed-column J^T threaded over nThreads (CompressedColumnJacobianTransposeBuilder.hpp:59-79) + rhs = J^T e.
// out_seconds: [problem construction + initialize, evaluate per repeat, build per repeat, cost].  The SOLVE is not timed here: the
// factorisation behind the reference's CHOLMOD wrappers is a stand-in in this build (CHOLMOD is not in the image).
// One serial evaluate + build runs first (inside the set-up time): the expression nodes cache fixed-size matrices in members while they
// evaluate, and the threaded passes store identical values into nodes that neighbouring terms share.
static int32_t time_evaluate_build(int32_t n_cams, const int32_t* cam_model, const double* cam_params, const double* baselines, int32_t n_sets, const double* set_poses,
                                   int32_t n_target, const double* target, int32_t n_views, const int32_t* view_set, const int32_t* view_cam, const int64_t* view_begin,
                                   const double* y_u, const double* y_v, const int32_t* corner_id, int32_t driver_order, int32_t solver_kind, int32_t n_threads,
                                   int32_t repeats, double* out_seconds) {
  try {
    typedef std::chrono::steady_clock clock;
    auto seconds = [](clock::time_point a, clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const clock::time_point t0 = clock::now();
    RigProblem rp(n_cams, cam_model, cam_params, baselines, n_sets, set_poses, n_target, target, n_views, view_set, view_cam, view_begin, y_u, y_v, corner_id, driver_order);
    Optimizer2Options options;
    options.nThreads = 1;
    options.trustRegionPolicy = boost::make_shared<LevenbergMarquardtTrustRegionPolicy>(10.0);
    if (solver_kind == 1) options.linearSystemSolver = boost::make_shared<SparseCholeskyLinearSystemSolver>();
    else options.linearSystemSolver = boost::make_shared<BlockCholeskyLinearSystemSolver>();
    Optimizer2 optimizer(options);
    optimizer.setProblem(rp.problem);
    optimizer.initialize();
    LinearSystemSolver* solver = solver_kind == 1 ? static_cast<LinearSystemSolver*>(optimizer.getSolver<SparseCholeskyLinearSystemSolver>())
                                                  : static_cast<LinearSystemSolver*>(optimizer.getSolver<BlockCholeskyLinearSystemSolver>());
    optimizer.evaluateError(true);
    solver->buildSystem(1, true);
    optimizer.options().nThreads = n_threads;
    const clock::time_point t1 = clock::now();
    double te = 0.0, tb = 0.0, J = 0.0;
    for (int r = 0; r < repeats; ++r) {
      const clock::time_point a = clock::now();
      J = optimizer.evaluateError(true);
      const clock::time_point b = clock::now();
      solver->buildSystem(n_threads, true);  // what the policy calls (LevenbergMarquardtTrustRegionPolicy.cpp:72)
      const clock::time_point c = clock::now();
      te += seconds(a, b);
      tb += seconds(b, c);
    }
    out_seconds[0] = seconds(t0, t1);
    out_seconds[1] = te / repeats;
    out_seconds[2] = tb / repeats;
    out_seconds[3] = J;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_time_evaluate_build: " << e.what() << std::endl;
    return -1;
  }
}

extern "C" __attribute__((visibility("default"))) int32_t ref_time_evaluate_build(int32_t n_cams, const int32_t* cam_model, const double* cam_params, const double* baselines,
                                                                                  int32_t n_sets, const double* set_poses, int32_t n_target, const double* target,
                                                                                  int32_t n_views, const int32_t* view_set, const int32_t* view_cam, const int64_t* view_begin,
                                                                                  const double* y_u, const double* y_v, const int32_t* corner_id, int32_t driver_order,
                                                                                  int32_t n_threads, int32_t repeats, double* out_seconds) {
  return time_evaluate_build(n_cams, cam_model, cam_params, baselines, n_sets, set_poses, n_target, target, n_views, view_set, view_cam, view_begin, y_u, y_v, corner_id,
                             driver_order, 0, n_threads, repeats, out_seconds);
}

extern "C" __attribute__((visibility("default"))) int32_t ref_time_evaluate_build_solver(int32_t n_cams, const int32_t* cam_model, const double* cam_params,
                                                                                         const double* baselines, int32_t n_sets, const double* set_poses, int32_t n_target,
                                                                                         const double* target, int32_t n_views, const int32_t* view_set,
                                                                                         const int32_t* view_cam, const int64_t* view_begin, const double* y_u,
                                                                                         const double* y_v, const int32_t* corner_id, int32_t driver_order,
                                                                                         int32_t solver_kind, int32_t n_threads, int32_t repeats, double* out_seconds) {
  return time_evaluate_build(n_cams, cam_model, cam_params, baselines, n_sets, set_poses, n_target, target, n_views, view_set, view_cam, view_begin, y_u, y_v, corner_id,
                             driver_order, solver_kind, n_threads, repeats, out_seconds);
}

// The incremental estimator's MERGED problem through the reference's own containers (§8f rank 2): one
// aslam::calibration::OptimizationProblem per synced set, filled in the order of kalibr2::tools::CreateBatchProblem
// (K2/CalibrationTools.hpp:460-521: the set's pose q, t in group 1; the shared baselines q, t in group 0; the shared landmarks - inactive -
// in group 2; the shared intrinsics in group 0; then the set's reprojection terms camera by camera, invR as set by ref_set_weighting),
// merged batch by batch with IncrementalOptimizationProblem::add (IC/src/core/IncrementalOptimizationProblem.cpp:186-223) and ordered as
// IncrementalEstimator::orderMarginalizedDesignVariables does (IC/src/core/IncrementalEstimator.cpp:550-565: the marginalised group - the
// calibration group 0 - swapped to the end).  Optimizer2 (Gauss-Newton policy, as the estimator runs it) then enumerates the active design
// variables of that container: out_order gets one row [kind, index, column base, dimension] per active design variable in the optimiser's
// order (kind 0 / 1 = pose q / t of set `index`, 2 / 3 = baseline q / t, 4 / 5 = projection / distortion of camera `index`), out_counts =
// [active design variables, groups ordering as a decimal number, e.g. 120 for {1, 2, 0}].  The optimisation's scalars and final state as
// ref_optimize_rig.  The container classes and their ordering logic are the reference's; the fill order restates the K2 header (OpenCV /
// ROS types in its signature keep it from compiling here).
extern "C" __attribute__((visibility("default"))) int32_t ref_estimator_problem(int32_t n_cams, const int32_t* cam_model, double* cam_params, double* baselines,
                                                                                int32_t n_sets, double* set_poses, int32_t n_target, const double* target, int32_t n_views,
                                                                                const int32_t* view_set, const int32_t* view_cam, const int64_t* view_begin,
                                                                                const double* y_u, const double* y_v, const int32_t* corner_id, int32_t max_iterations,
                                                                                double conv_dx, double conv_dj, int32_t restore_after, int32_t* out_order /*[<= 2 n_sets + 4 n_cams][4]*/,
                                                                                int32_t* out_counts /*[2]*/, double* out_scalars) {
  try {
    namespace ic = aslam::calibration;
    const size_t CALIBRATION_GROUP_ID = 0, TRANSFORMATION_GROUP_ID = 1, LANDMARK_GROUP_ID = 2;
    std::vector<boost::shared_ptr<CameraModel>> cams(n_cams);
    for (int k = 0; k < n_cams; ++k) {
      cams[k] = makeCamera(cam_model[k], cam_params + k * 10);
      cams[k]->proj->setActive(true);
      cams[k]->dist->setActive(true);
    }
    std::vector<boost::shared_ptr<RotationQuaternion>> bq, sq;
    std::vector<boost::shared_ptr<EuclideanPoint>> bt, st;
    std::vector<boost::shared_ptr<TransformationBasic>> B, poses;  // kept alive: toExpression() does not own them (K2's BatchProblemStruct holds them too)
    for (int j = 0; j + 1 < n_cams; ++j) {  // the shared baseline design variables exist before the batches (PoseDesignVariables)
      const double* p = baselines + 7 * j;
      bq.push_back(boost::make_shared<RotationQuaternion>(Eigen::Vector4d(p[0], p[1], p[2], p[3])));
      bt.push_back(boost::make_shared<EuclideanPoint>(Eigen::Vector3d(p[4], p[5], p[6])));
      bq.back()->setActive(true);
      bt.back()->setActive(true);
      B.push_back(boost::make_shared<TransformationBasic>(bq.back()->toExpression(), bt.back()->toExpression()));
    }
    std::vector<boost::shared_ptr<HomogeneousPoint>> points;  // landmarks: never set active in kalibr2
    for (int i = 0; i < n_target; ++i) points.push_back(boost::make_shared<HomogeneousPoint>(Eigen::Vector4d(target[3 * i], target[3 * i + 1], target[3 * i + 2], 1.0)));
    const boost::shared_ptr<MEstimator> policy = makePolicy(g_weighting);

    boost::shared_ptr<ic::IncrementalOptimizationProblem> merged(new ic::IncrementalOptimizationProblem());
    for (int v = 0; v < n_sets; ++v) {
      boost::shared_ptr<ic::OptimizationProblem> problem(new ic::OptimizationProblem());
      const double* p = set_poses + 7 * v;
      sq.push_back(boost::make_shared<RotationQuaternion>(Eigen::Vector4d(p[0], p[1], p[2], p[3])));  // AddPoseDesignVariable, :31-45
      sq.back()->setActive(true);
      problem->addDesignVariable(sq.back(), TRANSFORMATION_GROUP_ID);
      st.push_back(boost::make_shared<EuclideanPoint>(Eigen::Vector3d(p[4], p[5], p[6])));
      st.back()->setActive(true);
      problem->addDesignVariable(st.back(), TRANSFORMATION_GROUP_ID);
      poses.push_back(boost::make_shared<TransformationBasic>(sq.back()->toExpression(), st.back()->toExpression()));
      const boost::shared_ptr<TransformationBasic>& S = poses.back();
      for (size_t j = 0; j < B.size(); ++j) {
        problem->addDesignVariable(bq[j], CALIBRATION_GROUP_ID);
        problem->addDesignVariable(bt[j], CALIBRATION_GROUP_ID);
      }
      for (size_t i = 0; i < points.size(); ++i) problem->addDesignVariable(points[i], LANDMARK_GROUP_ID);
      for (int k = 0; k < n_cams; ++k) {  // AddIntrinsicDesignVariables: projection, distortion (the inactive 0-dimensional shutter block is left out)
        problem->addDesignVariable(cams[k]->proj, CALIBRATION_GROUP_ID);
        problem->addDesignVariable(cams[k]->dist, CALIBRATION_GROUP_ID);
      }
      for (int k = 0; k < n_cams; ++k)  // the set's views camera by camera
        for (int w = 0; w < n_views; ++w) {
          if (view_set[w] != v || view_cam[w] != k) continue;
          TransformationExpression T_cam_w = S->toExpression().inverse();
          for (int j = 0; j < k; ++j) T_cam_w = B[j]->toExpression() * T_cam_w;
          for (int64_t i = view_begin[w]; i < view_begin[w + 1]; ++i)
            problem->addErrorTerm(boost::make_shared<ReprojectionTerm>(Eigen::Vector2d(y_u[i], y_v[i]), T_cam_w * points[corner_id[i]]->toExpression(), cams[k].get(), policy));
        }
      merged->add(problem);
    }
    {  // IncrementalEstimator::orderMarginalizedDesignVariables with _margGroupId = the calibration group
      std::vector<size_t> ordering = merged->getGroupsOrdering();
      std::vector<size_t>::iterator it = std::find(ordering.begin(), ordering.end(), CALIBRATION_GROUP_ID);
      if (it == ordering.end()) throw std::runtime_error("the calibration group is not in the problem");
      if (*it != ordering.back()) {
        std::swap(*it, ordering.back());
        merged->setGroupsOrdering(ordering);
      }
    }
    out_counts[1] = 0;
    for (size_t g : merged->getGroupsOrdering()) out_counts[1] = out_counts[1] * 10 + (int32_t)g;

    Optimizer2Options options;
    options.nThreads = 1;
    options.convergenceDeltaX = conv_dx;
    options.convergenceDeltaJ = conv_dj;
    options.maxIterations = max_iterations;
    options.trustRegionPolicy = boost::make_shared<GaussNewtonTrustRegionPolicy>();
    options.linearSystemSolver = boost::make_shared<SparseCholeskyLinearSystemSolver>();
    Optimizer2 optimizer(options);
    optimizer.setProblem(merged);
    // restore_after: the reject path of IncrementalEstimator::addBatch (IC/src/core/IncrementalEstimator.cpp:350, 515) - the container's own
    // saveDesignVariables before the optimisation and restoreDesignVariables after it (IncrementalOptimizationProblem.cpp:415-426)
    if (restore_after) merged->saveDesignVariables();
    SolutionReturnValue r = optimizer.optimize();  // initialises: block indices and column bases stay as the optimiser assigned them
    if (restore_after) merged->restoreDesignVariables();
    std::map<const DesignVariable*, std::pair<int, int>> label;
    for (int v = 0; v < n_sets; ++v) { label[sq[v].get()] = std::make_pair(0, v); label[st[v].get()] = std::make_pair(1, v); }
    for (size_t j = 0; j < B.size(); ++j) { label[bq[j].get()] = std::make_pair(2, (int)j); label[bt[j].get()] = std::make_pair(3, (int)j); }
    for (int k = 0; k < n_cams; ++k) { label[cams[k]->proj.get()] = std::make_pair(4, k); label[cams[k]->dist.get()] = std::make_pair(5, k); }
    out_counts[0] = (int32_t)optimizer.numDesignVariables();
    for (size_t i = 0; i < optimizer.numDesignVariables(); ++i) {
      const DesignVariable* dv = optimizer.designVariable(i);
      const std::pair<int, int> l = label.count(dv) ? label[dv] : std::make_pair(-1, -1);
      out_order[4 * i] = l.first;
      out_order[4 * i + 1] = l.second;
      out_order[4 * i + 2] = dv->columnBase();
      out_order[4 * i + 3] = dv->minimalDimensions();
    }
    out_scalars[0] = r.iterations;
    out_scalars[1] = r.failedIterations;
    out_scalars[2] = r.JStart;
    out_scalars[3] = r.JFinal;
    out_scalars[4] = r.linearSolverFailure ? 1.0 : 0.0;
    for (int k = 0; k < n_cams; ++k) cams[k]->parameters(cam_params + k * 10);
    auto store = [](double* p, const boost::shared_ptr<RotationQuaternion>& q, const boost::shared_ptr<EuclideanPoint>& t) {
      const Eigen::Vector4d qv = q->getQuaternion();
      const Eigen::Vector3d tv = t->toEuclidean();
      for (int i = 0; i < 4; ++i) p[i] = qv(i);
      for (int i = 0; i < 3; ++i) p[4 + i] = tv(i);
    };
    for (size_t j = 0; j < bq.size(); ++j) store(baselines + 7 * j, bq[j], bt[j]);
    for (size_t v = 0; v < sq.size(); ++v) store(set_poses + 7 * v, sq[v], st[v]);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ref_estimator_problem: " << e.what() << std::endl;
    return -1;
  }
}
