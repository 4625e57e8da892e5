/*
 * kalibr_b200.h — C ABI of the B200-native batch-calibration hot path.
 *
 * This is the drop-in boundary.  Every entry point names the
 * reference interface it replaces; file:line citations are relative to the
 * reference tree (ToyotaResearchInstitute/kalibr), abbreviations:
 *   BE  = aslam_optimizer/aslam_backend
 *   BX  = aslam_optimizer/aslam_backend_expressions
 *   CAM = aslam_cv/aslam_cameras
 *   K2  = aslam_offline_calibration/kalibr2
 *
 * Conventions: opaque handle, int status (0 = OK, negative = error, message
 * via kb_last_error), no exceptions cross the boundary.
 * Thread safety: a handle is driven by ONE caller thread at a time
 * (BE/src/Optimizer2.cpp:183-273 drives a solver from one thread); different
 * handles - on the same GPU or on different GPUs - may be driven from different
 * threads concurrently: every handle owns its stream, its scratch buffers and
 * its CUDA graph, and the library's only process-wide state (per-device launch
 * attribute caches, the dlopen'ed NCCL entry points) is guarded.
 * All pointers in the signatures are caller-owned HOST pointers (FP64 /
 * int32 / int64, SoA); the library owns every device buffer.  There is no CPU
 * fallback: kb_create fails with KB_ERR_NO_DEVICE when no sm_100 device is
 * usable.
 */
#ifndef KALIBR_B200_H_
#define KALIBR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB_API __attribute__((visibility("default")))

/* ---- status codes ------------------------------------------------------- */
typedef int32_t kb_status;
#define KB_OK 0
#define KB_ERR_INVALID_ARGUMENT (-1)
#define KB_ERR_NO_DEVICE (-2)
#define KB_ERR_CUDA (-3)
#define KB_ERR_NCCL (-4)
#define KB_ERR_STATE (-5)      /* call sequence violated (e.g. solve before build) */
#define KB_ERR_ALLOC (-6)
#define KB_ERR_NUMERICAL (-7)  /* a numerical outcome, not a fault: pose block not positive definite, SVD iteration not converged */

/* ---- camera models: K2/include/kalibr2/CameraCalibrator.hpp:421-441 ------ */
typedef enum {
  KB_PINHOLE_RADTAN = 0, /* PinholeProjection<RadialTangentialDistortion>  P=4 (fu,fv,cu,cv)            D=4 (k1,k2,p1,p2) */
  KB_PINHOLE_EQUI = 1,   /* PinholeProjection<EquidistantDistortion>       P=4                          D=4 (k1..k4)      */
  KB_OMNI_RADTAN = 2,    /* OmniProjection<RadialTangentialDistortion>     P=5 (xi,fu,fv,cu,cv)         D=4               */
  KB_EUCM_NONE = 3,      /* ExtendedUnifiedProjection<NoDistortion>        P=6 (alpha,beta,fu,fv,cu,cv) D=0 (active, 0-dim) */
  KB_DS_NONE = 4,        /* DoubleSphereProjection<NoDistortion>           P=6 (xi,alpha,fu,fv,cu,cv)   D=0 (active, 0-dim) */
  KB_PINHOLE_FOV = 5,    /* PinholeProjection<FovDistortion>               P=4 (fu,fv,cu,cv)            D=1 (w) */
  KB_OMNI_NONE = 6,      /* OmniProjection<NoDistortion>                   P=5 (xi,fu,fv,cu,cv)         D=0 (active, 0-dim) */
  KB_NUM_MODELS = 7      /* = the model table of kalibr2::CreateCalibrator (K2/include/kalibr2/CameraCalibrator.hpp:421-441) */
} kb_camera_model;

/* Design-variable insertion order of the three batch drivers
 * (K2/include/kalibr2/CalibrationTools.hpp:93-144, 183-300, 376-428).  It
 * fixes blockIndex / columnBase (BE/src/Optimizer2.cpp:110-124). */
typedef enum {
  KB_ORDER_SINGLE = 0, /* proj, dist ; per set q_v, t_v                                  */
  KB_ORDER_STEREO = 1, /* baseline q,t ; per set q_v,t_v ; proj0,dist0, proj1,dist1      */
  KB_ORDER_RIG = 2,    /* per cam proj,dist ; per baseline q,t ; per set q_v,t_v         */
  KB_ORDER_BATCH = 3   /* per set q_v,t_v ; per baseline q,t ; per cam proj,dist: the incremental estimator's merged problem -
                        * CreateBatchProblem puts the target pose in group 1, baselines then intrinsics in group 0 and the (inactive)
                        * landmarks in group 2 (K2/include/kalibr2/CalibrationTools.hpp:460-491); groups are ordered by first
                        * appearance, [1, 0, 2] (IC/src/core/IncrementalOptimizationProblem.cpp:186-223), and
                        * IncrementalEstimator::orderMarginalizedDesignVariables moves the marginalised group 0 last
                        * (IC/src/core/IncrementalEstimator.cpp:550-565): the calibration block is the LAST n_c columns */
} kb_driver_order;

#define KB_CAM_PARAM_STRIDE 10 /* per camera: P projection params then D distortion params, zero padded */
#define KB_POSE_STRIDE 7       /* q[4] (x,y,z,w; sm_kinematics JPL, scalar last) then t[3] */

/*
 * Problem description = what K2's drivers hand to Optimizer2 as an
 * OptimizationProblem, flattened.  A "view" is one (synced set, camera) image:
 * the run of ReprojectionError terms added by one AddReprojectionErrorsForView
 * call (K2/include/kalibr2/CameraCalibrator.hpp:238-265).  Terms (observed
 * corners) are listed in the reference's error-term insertion order, so term
 * index i has rowBase 2*i (BE/src/Optimizer2.cpp:130-135); views index
 * contiguous runs of it.
 */
typedef struct {
  int32_t driver_order;        /* kb_driver_order */
  int32_t n_cams;
  const int32_t* cam_model;    /* [n_cams] kb_camera_model */
  const double* cam_params;    /* [n_cams][KB_CAM_PARAM_STRIDE] initial intrinsics */
  const double* baselines;     /* [n_cams-1][KB_POSE_STRIDE]  T_cam(k+1)_cam(k) initial guess */
  int32_t n_sets;
  const double* set_poses;     /* [n_sets][KB_POSE_STRIDE]    T_target_cam0 initial guess (pose DV; the term uses its inverse) */
  int32_t n_target_points;
  const double* target_points; /* [n_target_points][3] */
  int32_t n_views;
  const int32_t* view_set;     /* [n_views] */
  const int32_t* view_cam;     /* [n_views] */
  const int64_t* view_begin;   /* [n_views+1] term range of each view */
  int64_t n_terms;
  const double* y_u;           /* [n_terms] measured corner, pixels */
  const double* y_v;           /* [n_terms] */
  const int32_t* corner_id;    /* [n_terms] index into target_points */
  /* multi-GPU: sets are sharded by contiguous ranges over n_ranks; NCCL id from kb_nccl_unique_id */
  int32_t n_ranks;
  int32_t rank;
  const char* nccl_id;         /* 128 bytes, NULL when n_ranks == 1 */
  int32_t device;              /* CUDA device ordinal */
  /* pre-sharded input (optional): when n_sets_total > 0 this description holds ONLY this rank's sets, which are the
   * global sets [set_offset, set_offset + n_sets) of a problem with n_sets_total sets; view_set is local (0-based). */
  int32_t n_sets_total;
  int32_t set_offset;
  int64_t n_terms_total;       /* global number of terms when pre-sharded (for kb_jrows) */
} kb_problem_desc;

typedef struct kb_handle kb_handle;

/* ---- life cycle ---------------------------------------------------------- */
/* ≙ Optimizer2::initialize + LinearSystemSolver::initMatrixStructure
 *   (BE/src/Optimizer2.cpp:95-151, BE/src/LinearSystemSolver.cpp:117-138,
 *    BE/src/BlockCholeskyLinearSystemSolver.cpp:34-55).  Uploads the problem. */
KB_API kb_status kb_create(const kb_problem_desc* desc, kb_handle** out);
KB_API void kb_destroy(kb_handle* h);
KB_API const char* kb_last_error(const kb_handle* h); /* h may be NULL: last create error */
KB_API kb_status kb_nccl_unique_id(char out[128]);     /* rank 0 calls, caller broadcasts */

/* sizes: ≙ LinearSystemSolver::JRows/JCols (BE/include/aslam/backend/LinearSystemSolver.hpp:62-66) */
KB_API int64_t kb_jrows(const kb_handle* h);  /* 2 * GLOBAL number of terms */
KB_API int64_t kb_jcols(const kb_handle* h);  /* sum of active DV dimensions */
KB_API int64_t kb_local_jrows(const kb_handle* h); /* 2 * number of terms owned by this rank (== kb_jrows when n_ranks == 1) */
KB_API int32_t kb_num_design_variables(const kb_handle* h); /* active DVs incl. 0-dim ones */
/* ≙ DesignVariable::blockIndex/columnBase/minimalDimensions (BE/include/aslam/backend/DesignVariable.hpp:18-145) */
KB_API kb_status kb_get_dv_layout(const kb_handle* h, int32_t* column_base /*[n_dv]*/, int32_t* dims /*[n_dv]*/);

/* Peer exchange over NVLink (optional, 2..8 ranks on one node, one process per GPU).  After kb_create every rank publishes the
 * CUDA IPC handle of its exchange buffer; the caller all-gathers the 64-byte handles (rank order) and hands them to
 * kb_attach_peers.  From then on the three exchange steps of an LM iteration (the reduced camera system, the solve scalars,
 * the cost) no longer go through NCCL: the producing kernels store straight into every rank's buffer and raise epoch flags,
 * the consuming kernels wait for the flags and sum in rank order.  Collective: every rank must attach. */
KB_API kb_status kb_peer_exchange_handle(kb_handle* h, char out[64]);
KB_API kb_status kb_attach_peers(kb_handle* h, const char* handles /*[n_ranks][64]*/);

/* ---- per-iteration hot path ---------------------------------------------- */
/* ≙ LinearSystemSolver::evaluateError (BE/src/LinearSystemSolver.cpp:81-92): returns J = sum w e^T invR e
 *   over ALL ranks (w = weight of the installed M-estimator policy, 1 without one: BE/src/ErrorTerm.cpp:19-24) and fills the
 *   device copy of e() = -sqrt(w) sqrtInvR^T e; use_m_estimator = 0 leaves sqrt(w) out of e() exactly as
 *   ErrorTermFs::getWeightedError does (BE/include/aslam/backend/implementation/ErrorTerm.hpp:183-192). */
KB_API kb_status kb_evaluate_error(kb_handle* h, int32_t use_m_estimator, double* out_cost);
/* ≙ BlockCholeskyLinearSystemSolver::buildSystem (BE/src/BlockCholeskyLinearSystemSolver.cpp:58-72):
 *   linearise every term at the current state and assemble H, rhs. */
KB_API kb_status kb_build_system(kb_handle* h, int32_t use_m_estimator);
/* ≙ LinearSystemSolver::setConstantConditioner (BE/src/LinearSystemSolver.cpp:111-114); squared when applied */
KB_API kb_status kb_set_constant_conditioner(kb_handle* h, double lambda);
/* ≙ BlockCholeskyLinearSystemSolver::solveSystem (BE/src/BlockCholeskyLinearSystemSolver.cpp:74-106), including
 *   the λ² augment / λ un-augment asymmetry.  dx has kb_jcols entries (this rank's poses + shared block; poses of
 *   other ranks' sets are returned as 0 unless gather_dx != 0).  *pos_def = 0 ≙ solveSystem returning false. */
KB_API kb_status kb_solve_system(kb_handle* h, double* dx, int32_t gather_dx, int32_t* pos_def);
/* dx^T (lambda dx + rhs) on the device: the denominator of LevenbergMarquardtTrustRegionPolicy::getLmRho
 *   (BE/src/LevenbergMarquardtTrustRegionPolicy.cpp:107-113) for the dx of the last kb_solve_system. */
KB_API kb_status kb_lm_rho_denominator(kb_handle* h, double lambda, double* out);
/* ≙ Optimizer2::applyStateUpdate (BE/src/Optimizer2.cpp:290-307) with the dx of the last solve; returns max|dx| */
KB_API kb_status kb_apply_state_update(kb_handle* h, double* out_max_abs_dx);
/* ≙ Optimizer2::revertLastStateUpdate (BE/src/Optimizer2.cpp:313-318) */
KB_API kb_status kb_revert_last_state_update(kb_handle* h);

/* One iteration's worth of the calls above in one go - ≙ evaluateError, buildSystem, setConstantConditioner(lambda), solveSystem,
 * applyStateUpdate and, if `revert`, revertLastStateUpdate (the sequence of BE/src/Optimizer2.cpp:237-249 with
 * LevenbergMarquardtTrustRegionPolicy.cpp:72-88) - enqueued back to back with a single host synchronisation at the end, for host
 * optimisers that only need the scalars of the iteration.  Same kernels, same results as the six separate calls.
 * out == NULL: the iteration is only ENQUEUED (no synchronisation; several iterations can be in flight); kb_wait() then waits for
 * everything enqueued on the handle and returns the scalars of the LAST iteration. */
typedef struct {
  double cost;            /* J of the state the iteration started from */
  double rho_denominator; /* dx^T (lambda dx + rhs) of its solution */
  double max_abs_dx;
  int32_t pos_def;
} kb_iteration_result;
KB_API kb_status kb_iterate(kb_handle* h, double lambda, int32_t use_m_estimator, int32_t revert, kb_iteration_result* out);
KB_API kb_status kb_wait(kb_handle* h, kb_iteration_result* out);

/* One whole Optimizer2::optimize() (BE/src/Optimizer2.cpp:183-273) with LevenbergMarquardtTrustRegionPolicy
 * (BE/src/LevenbergMarquardtTrustRegionPolicy.cpp:50-113) driven by the host C++ mirror; state stays on the device. */
typedef struct {
  double convergence_delta_x; /* 1e-3  K2 CalibrationTools.hpp:57-66 */
  double convergence_delta_j; /* 1.0   */
  int32_t max_iterations;     /* 200   */
  double lm_lambda_init;      /* 10.0  */
  int32_t verbose;
  int32_t device_loop;        /* 1 (default): the whole LM loop runs on the device - trust-region logic in single-thread control
                               * kernels, iterations enqueued ahead, the host only polls the "done" flag; 0: the host-side
                               * Optimizer2 mirror drives the call-by-call entry points (two synchronisations per iteration).
                               * Same iterations and results.  verbose = 1 or speculative linearisation off imply 0. */
} kb_optimizer_options;
typedef struct { /* ≙ SolutionReturnValue, BE/include/aslam/backend/backend.hpp:14-27 */
  double j_start, j_final, dx_final, dj_final;
  int32_t iterations, failed_iterations, linear_solver_failure;
} kb_solution;
KB_API void kb_default_optimizer_options(kb_optimizer_options* o);
KB_API kb_status kb_optimize(kb_handle* h, const kb_optimizer_options* o, kb_solution* out);
/* per-iteration trace of the last kb_optimize: triples (J, deltaX, lambda); returns the number of triples */
KB_API int32_t kb_get_trace(const kb_handle* h, double* out, int32_t max_triples);
/* Solver semantic: 0 = BlockCholesky (default; un-augments with lambda instead of lambda^2),
 * 1 = SparseCholesky (damping appended as columns, no residual: BE/src/SparseCholeskyLinearSystemSolver.cpp:48-66) */
KB_API kb_status kb_set_solver_semantic(kb_handle* h, int32_t semantic);
/* Speculative linearisation (default on): kb_evaluate_error runs the fused linearise+assemble kernel (its Gram block yields
 * the cost, e() is written too), so that a kb_build_system at the same state — the accepted-step case of the LM loop —
 * only has to reduce the view blocks.  Off: kb_evaluate_error runs the residual-only kernel. Results are identical. */
KB_API kb_status kb_set_speculative_linearise(kb_handle* h, int32_t on);

/* ---- weighting of the terms ---------------------------------------------------------
 * ≙ ErrorTermFs<2>::setInvR on every ReprojectionError (BE/include/aslam/backend/implementation/ErrorTerm.hpp:118-127;
 *   K2 passes invR = I in the batch drivers, CalibrationTools.hpp:105-108, and I / sigma^2 in CreateBatchProblem, :495-496).
 *   inv_r: symmetric positive definite 2x2, row-major.  Its square root follows sm::eigen::computeMatrixSqrt
 *   (Schweizer-Messer/sm_eigen/include/sm/eigen/matrix_sqrt.hpp:21-40: S = P^T L sqrt(D) of Eigen's pivoted LDL^T), so that
 *   e() = -S^T e and the exported Jacobian rows S^T J match the reference value for value.  Default: identity. */
KB_API kb_status kb_set_inv_r(kb_handle* h, const double inv_r[4]);
KB_API kb_status kb_get_sqrt_inv_r(const kb_handle* h, double sqrt_inv_r[4] /* row-major S */);
/* ≙ ErrorTerm::setMEstimatorPolicy on every term (BE/src/ErrorTerm.cpp:46-56) with the policies of
 *   BE/src/MEstimatorPolicies.cpp.  Parameters: HUBER p0 = k; CAUCHY, GEMAN_MCCLURE p0 = sigma^2;
 *   BLAKE_ZISSERMAN p0 = df, p1 = pCut (0.999), p2 = wCut (0.1) -> epsilon = (1 - wCut) / wCut exp(-chi2inv(pCut, df)).
 *   The weight w = policy(e^T invR e) always enters the cost; it scales e() and the Jacobian rows when the
 *   use_m_estimator argument of kb_evaluate_error / kb_build_system is non-zero (Optimizer2 always passes true:
 *   BE/src/Optimizer2.cpp:198, 237; BE/src/LevenbergMarquardtTrustRegionPolicy.cpp:56, 72). */
typedef enum {
  KB_MEST_NONE = 0,            /* NoMEstimator            w = 1 */
  KB_MEST_HUBER = 1,           /* HuberMEstimator         w = s < k^2 ? 1 : k / sqrt(s) */
  KB_MEST_CAUCHY = 2,          /* CauchyMEstimator        w = 1 / (1 + s / sigma^2) */
  KB_MEST_GEMAN_MCCLURE = 3,   /* GemanMcClureMEstimator  w = sigma^2 / (sigma^2 + s)^2 */
  KB_MEST_BLAKE_ZISSERMAN = 4  /* BlakeZissermanMEstimator w = exp(-s) / (exp(-s) + epsilon) */
} kb_m_estimator;
KB_API kb_status kb_set_m_estimator(kb_handle* h, int32_t kind, double p0, double p1, double p2);
KB_API double kb_m_estimator_parameter(const kb_handle* h); /* k, sigma^2 or the derived epsilon */

/* ---- reprojection statistics ---------------------------------------------------------
 * ≙ CameraCalibrator::PrintReprojectionErrorStatistics (K2/include/kalibr2/CameraCalibrator.hpp:368-405) for every camera, at
 *   the current state, over ALL ranks: per camera [n, mean_u, mean_v, std_u, std_v, rmse] with the raw errors y - y_hat
 *   (:267-286), the two-pass sample standard deviation (N - 1; 0 below two samples) and "RMSE" exactly as printed there:
 *   |sum of the error vectors| / sqrt(n).  Collective when n_ranks > 1.  e() is left untouched. */
#define KB_REPROJ_STAT_STRIDE 6
KB_API kb_status kb_reprojection_statistics(kb_handle* h, double* out /*[n_cams][KB_REPROJ_STAT_STRIDE]*/);

/* ---- initial-guess stage (what the drivers run before Optimizer2) -----------------------
 * ≙ CameraGeometry::estimateTransformation for EVERY local view with the handle's current intrinsics
 *   (CAM/include/aslam/cameras/implementation/PinholeProjection.hpp:831-891, OmniProjection.hpp:882-955,
 *    ExtendedUnifiedProjection.hpp:791-860, DoubleSphereProjection.hpp:842-910): corners through cv::Point2f / Point3f
 *   (float), keypointToEuclidean, the 80 degree cone, then cv::solvePnP (planar target: homography start + Levenberg-
 *   Marquardt on the reprojection error), one warp per view.  T_t_c: pose (q xyzw, t) of T_target_camera as
 *   sm::kinematics::Transformation::set stores it; ok[w] = 0 ≙ estimateTransformation returning false (fewer than 4 usable
 *   corners).  resolution: [n_cams][2] (ru, rv) for the pinhole models' isValid(keypoint) test, or NULL to skip it. */
KB_API kb_status kb_estimate_transformations(kb_handle* h, const int32_t* resolution, double* T_t_c /*[n_views][KB_POSE_STRIDE]*/,
                                             int32_t* ok /*[n_views]*/);
/* ≙ getTargetPoseGuess for every synced set (K2/include/kalibr2/CalibrationTools.hpp:316-356): PnP in the camera that saw most
 *   corners (first one on ties), chained through the handle's current baselines exactly as the reference's std::accumulate does
 *   (((T_t_cN B_0) B_1) ... B_{N-1}).  The guesses replace the set-pose state and what kb_reset_state returns to.
 *   n_failed (may be NULL): local sets whose PnP failed (they get the identity chained through the baselines, as the reference,
 *   which ignores estimateTransformation's return value there) or that no camera saw (pose left as it was). */
KB_API kb_status kb_initialize_set_poses(kb_handle* h, const int32_t* resolution, int32_t* n_failed);
/* ≙ the baseline guess of CalibrateStereoPair (CalibrationTools.hpp:195-234): per-component upper median (K2/src/
 *   BasicMathUtils.cpp:10-33) of the translation and the rotation vector (sm::kinematics::RotationVector) of T_H^-1 T_L over
 *   the synced sets both cameras saw.  baseline: pose (q, t) of T_camH_camL; n_used (may be NULL): sets that contributed. */
KB_API kb_status kb_estimate_stereo_baseline(kb_handle* h, const int32_t* resolution, int32_t cam_l, int32_t cam_h, double* baseline /*[7]*/,
                                             int32_t* n_used);

/* ≙ CameraGeometry::initializeIntrinsics for camera `cam` from all of its views, on a target_rows x target_cols grid target
 *   (corner id = row * cols + col, CAM/src/GridCalibrationTargetBase.cpp:28-30):
 *   pinhole-*: focal length = median of |v1 - v2| / pi over the vanishing-point pairs of the grid rows seen as circles, complete views
 *              only (CAM/.../implementation/PinholeProjection.hpp:599-803); fallback_focal_length > 0 is used when no guess exists;
 *   omni-*, eucm-none, ds-none: per (view, row) a focal guess from the row's line image, kept if it gives the lowest mean reprojection
 *              error after a PnP of that view (OmniProjection.hpp:721-866; ExtendedUnifiedProjection.hpp:734-760 and
 *              DoubleSphereProjection.hpp:785-811 halve it and set alpha = 0.5, beta = 1 / xi = 0, alpha = 0.5).
 *   Principal point = image centre, distortion cleared.  The guess replaces the camera's parameters (state and the reset point) and is
 *   returned in params; *success ≙ the reference's return value.  Single rank.  The reference's pinhole pair loop reads past its
 *   circle arrays when cols > rows (undefined behaviour); here the pairs are taken among the rows that exist. */
KB_API kb_status kb_initialize_intrinsics(kb_handle* h, int32_t cam, int32_t target_rows, int32_t target_cols,
                                          const int32_t* resolution /*[n_cams][2]*/, double fallback_focal_length /* <= 0: none */,
                                          double* params /*[KB_CAM_PARAM_STRIDE], may be NULL*/, int32_t* success);

/* ---- the incremental estimator's linear solver ------------------------------------------
 * ≙ aslam::calibration::LinearSolver::solveSystem -> solve (aslam_incremental_calibration/incremental_calibration/src/core/
 *   LinearSolver.cpp:245-285, 299-463), the solver IncrementalEstimator gives Optimizer2 together with the Gauss-Newton policy
 *   (src/core/IncrementalEstimator.cpp:66-71): an UNDAMPED least-squares step in which the set-pose columns are eliminated (QR
 *   there, the Schur complement here) and the calibration block (intrinsics + baselines) is solved through the SVD of
 *   Omega = A_r^T A_r - (A_r^T Q)(A_r^T Q)^T cut at the numerical rank (tolerance sv[0] * eps_svd * n unless svd_tol is given:
 *   src/algorithms/linalg.cpp:244-261), so that unobservable directions get a zero update; column_scaling scales the calibration
 *   columns to unit norm first (linalg.cpp:128-152, kalibr2_ros turns it on with eps_svd = 1e-6: CalibrateCameras.cpp:264-267).
 *   Call after kb_build_system; kb_apply_state_update then applies the step. */
typedef struct {
  int32_t column_scaling; /* 0 */
  double eps_norm;        /* std::numeric_limits<double>::epsilon() */
  double eps_svd;         /* std::numeric_limits<double>::epsilon() */
  double svd_tol;         /* -1: derive from eps_svd */
} kb_svd_solver_options;
typedef struct {
  int32_t n, rank, rank_deficiency; /* ≙ getSVDRank / getSVDRankDeficiency of the (scaled) system */
  double tolerance, sv_gap;         /* ≙ getSVDTolerance, getSvGap */
} kb_svd_solve_result;
KB_API void kb_default_svd_solver_options(kb_svd_solver_options* o);
KB_API kb_status kb_solve_system_svd(kb_handle* h, const kb_svd_solver_options* o, double* dx /*[jcols] or NULL*/, int32_t gather_dx,
                                     kb_svd_solve_result* out /*may be NULL*/, double* singular_values /*[n] or NULL: of the scaled system*/);

/* One Optimizer2::optimize() with GaussNewtonTrustRegionPolicy (BE/src/GaussNewtonTrustRegionPolicy.cpp:18-40: rebuild and solve
 * every iteration, no conditioner, never revert) over kb_solve_system_svd — the optimisation IncrementalEstimator::addBatch runs
 * (IC/src/core/IncrementalEstimator.cpp:66-71, 377; kalibr2_ros: max_iterations = 20, both convergence deltas 1e-3). */
/* rank / tolerance / gap of the most recent kb_solve_system_svd on this handle (rank = -1 before the first one): what
 * LinearSolver::getSVDRank() keeps returning after analyzeMarginal(), which only recomputes them when no solve has run
 * (LinearSolver.cpp:517-523) */
KB_API kb_status kb_get_last_svd_solve(const kb_handle* h, kb_svd_solve_result* out);
/* |eigenvalues| (descending) and eigenvectors of the (column-scaled) reduced system of the most recent kb_solve_system_svd, as they
 * sit on the device right after it: ≙ LinearSolver::getSingularValues / getNullSpace / getRowSpace after optimize() and before
 * analyzeMarginal() (IC/src/core/IncrementalEstimator.cpp:387-401).  V: [n][n] row-major, column k = k-th vector; columns: [n]
 * design-variable column of each row of V.  Any pointer may be NULL.  A later kb_analyze_marginal* overwrites the buffers. */
KB_API kb_status kb_get_last_svd_decomposition(kb_handle* h, double* singular_values, double* V, int32_t* columns);
KB_API kb_status kb_optimize_gauss_newton(kb_handle* h, const kb_optimizer_options* o, const kb_svd_solver_options* so, kb_solution* out);

/* ---- marginal analysis of the calibration block --------------------------------
 * ≙ aslam::calibration::LinearSolver::analyzeMarginal (aslam_incremental_calibration/incremental_calibration/src/core/
 *   LinearSolver.cpp:466-528) — what IncrementalEstimator::addBatch asks of its solver after every re-optimisation
 *   (src/core/IncrementalEstimator.cpp:404-433): the information matrix of the camera-side variables with the set poses
 *   marginalised out, Omega = A_r^T A_r - (A_r^T Q)(A_r^T Q)^T, is exactly the undamped Schur-reduced camera system of this
 *   path; its singular values, numerical rank (tolerance sv[0] * eps_svd * n unless svd_tol is given: linalg.cpp:244-261),
 *   gap and log2-sum over the first `rank` values (LinearSolver.cpp:196-200; information gain = half the difference of two
 *   such sums) come from a one-sided Jacobi iteration on the device.  Linearises at the current state and rebuilds the system
 *   (like kb_build_system); a following kb_solve_system needs no further build. */
typedef struct {
  double eps_svd; /* std::numeric_limits<double>::epsilon() (LinearSolverOptions.cpp:33) */
  double svd_tol; /* -1: derive from eps_svd (LinearSolverOptions.cpp:35) */
} kb_marginal_options;
typedef struct {
  int32_t n;               /* dimension of the calibration block (= camera-side columns) */
  int32_t rank;            /* ≙ getSVDRank */
  int32_t rank_deficiency; /* ≙ getSVDRankDeficiency */
  double tolerance;        /* ≙ getSVDTolerance */
  double sv_log2_sum;      /* ≙ getSingularValuesLog2Sum */
  double sv_gap;           /* ≙ getSvGap (inf when full rank) */
} kb_marginal_result;
KB_API void kb_default_marginal_options(kb_marginal_options* o);
/* singular_values: [n] descending; V (may be NULL): [n][n] row-major, column k = k-th right singular vector (≙ getRowSpace /
 * getNullSpace = its first rank / remaining columns); columns (may be NULL): [n] design-variable column of each row of V */
KB_API kb_status kb_analyze_marginal(kb_handle* h, const kb_marginal_options* o, kb_marginal_result* out, double* singular_values,
                                     double* V, int32_t* columns);
/* The same analysis on the system of the LAST kb_build_system instead of a fresh linearisation: IncrementalEstimator::addBatch calls
 * analyzeMarginal() after optimize(), i.e. on the Jacobian of the last Gauss-Newton iteration, one update behind the final state
 * (IC/src/core/IncrementalEstimator.cpp:377, 404; LinearSolver.cpp:530-537). */
KB_API kb_status kb_analyze_marginal_last_build(kb_handle* h, const kb_marginal_options* o, kb_marginal_result* out, double* singular_values,
                                                double* V, int32_t* columns);

/* ---- read-back (parity / results) ----------------------------------------- */
KB_API kb_status kb_get_error_vector(kb_handle* h, double* e /*[2*local terms]*/);      /* ≙ LinearSystemSolver::e() : -sqrtInvR^T e */
KB_API kb_status kb_get_rhs(kb_handle* h, double* rhs /*[jcols]*/);                     /* ≙ LinearSystemSolver::rhs() */
/* Materialising linearise (K1): e and J of every local term at the current state, J^T in the compressed-column
 * layout CompressedColumnJacobianTransposeBuilder holds (BE/include/aslam/backend/implementation/
 * CompressedColumnJacobianTransposeBuilder.hpp:59-100; CompressedColumnMatrix.hpp:236-304): two columns per term, rows =
 * design-variable columns ordered by block index.  kb_linearise runs the kernel (values stay on the device);
 * kb_get_jacobian_ccs copies them out.  Call with NULL arrays to obtain nnz (also the return of kb_jacobian_nnz). */
KB_API kb_status kb_linearise(kb_handle* h);
KB_API int64_t kb_jacobian_nnz(const kb_handle* h);
KB_API kb_status kb_get_jacobian_ccs(kb_handle* h, int64_t* col_ptr /*[2*local terms+1]*/, int32_t* row_idx /*[nnz]*/, double* values /*[nnz]*/);
/* Upper-triangular block pattern of H exactly as SparseBlockMatrix holds it after buildSystem+solveSystem
 * (sparse_block_matrix.hpp:121-143; JacobianContainer.cpp:112-126): for block column c, block_row[col_ptr[c]..col_ptr[c+1])
 * ascending.  values are the blocks column-major one after another (value_ptr[k] offsets).  Call with NULL arrays to query counts. */
KB_API kb_status kb_get_hessian_blocks(kb_handle* h, int64_t* n_blocks, int64_t* n_values, int64_t* col_ptr /*[n_dv+1]*/,
                                       int32_t* block_row, int64_t* value_ptr, double* values);
/* ---- a live handle grows by synced sets ------------------------------------------------------------------------
 * ≙ IncrementalOptimizationProblem::add / remove of one batch (IC/src/core/IncrementalOptimizationProblem.cpp:186-260; a kalibr2 batch
 *   is one synced set: CalibrationTools.hpp:460-521) on the problem a handle holds, instead of re-creating the handle per batch:
 *   kb_append_set adds synced set number kb "n_sets" with its views (terms appended behind the existing ones: only the new
 *   observations are uploaded; device arrays grow by capacity doubling; streams, events, pinned buffers and every buffer that is
 *   already large enough are kept) and its pose guess; kb_remove_last_set drops the last synced set again (a rejected batch:
 *   IncrementalEstimator.cpp:520-530).  The design-variable layout follows the handle's driver order with the new number of sets
 *   (kb_jcols / kb_get_dv_layout change).  Single rank.  view_cam: [n_views] distinct cameras; view_begin: [n_views + 1] term
 *   range of each view inside y_u / y_v / corner_id (as in kb_problem_desc); set_pose: [7]. */
KB_API kb_status kb_append_set(kb_handle* h, int32_t n_views, const int32_t* view_cam, const int64_t* view_begin, const double* y_u, const double* y_v,
                               const int32_t* corner_id, const double* set_pose);
KB_API kb_status kb_remove_last_set(kb_handle* h);
/* ≙ OptimizationProblem::saveDesignVariables / restoreDesignVariables (IC/src/core/OptimizationProblem.cpp:260-272), the bracket
 *   IncrementalEstimator::addBatch puts around a batch (IncrementalEstimator.cpp:348-350, 520-524): a snapshot of every design variable
 *   on the device, and back.  A set appended after the snapshot keeps its own pose on restore. */
KB_API kb_status kb_save_design_variables(kb_handle* h);
KB_API kb_status kb_restore_design_variables(kb_handle* h);

/* Replace the current state ≙ DesignVariable::setParameters / what Optimizer2::applyStateUpdate and revertLastStateUpdate leave in
 * the HOST design variables (BE/src/Optimizer2.cpp:290-318; BE/include/aslam/backend/DesignVariable.hpp:18-145): a
 * LinearSystemSolver plugged into an unmodified Optimizer2 does not own the design variables, so its adapter pushes their values
 * before every evaluateError / buildSystem.  Any of the three pointers may be NULL (left as is).  The normal
 * equations of the last kb_build_system stay valid for further solves (the LM policy re-solves a built system after a revert);
 * the cached speculative linearisation is invalidated.  set_poses: [n_sets][7] in the layout of kb_problem_desc (global sets,
 * or this rank's sets when pre-sharded); quaternions are taken as given (the reference's updates keep them normalised). */
KB_API kb_status kb_set_state(kb_handle* h, const double* cam_params /*[n_cams][KB_CAM_PARAM_STRIDE]*/, const double* baselines /*[n_cams-1][7]*/,
                              const double* set_poses /*[n_sets][7]*/);
KB_API kb_status kb_set_camera_params(kb_handle* h, const double* cam_params);
KB_API kb_status kb_set_baselines(kb_handle* h, const double* baselines);
KB_API kb_status kb_set_set_poses(kb_handle* h, const double* set_poses);
/* ≙ LinearSystemSolver::setConditioner(const Eigen::VectorXd&) (BE/include/aslam/backend/LinearSystemSolver.hpp:35-36): the per-column
 * conditioner.  The trust-region policies of the reference only ever install a constant one (setConstantConditioner,
 * LevenbergMarquardtTrustRegionPolicy.cpp:87), which is what the Schur path implements: a vector whose jcols entries are all equal
 * is accepted (= kb_set_constant_conditioner), anything else returns KB_ERR_INVALID_ARGUMENT. */
KB_API kb_status kb_set_conditioner(kb_handle* h, const double* diag /*[jcols]*/);
/* current state ≙ DesignVariable::getParameters */
KB_API kb_status kb_get_camera_params(kb_handle* h, double* cam_params /*[n_cams][KB_CAM_PARAM_STRIDE]*/);
KB_API kb_status kb_get_baselines(kb_handle* h, double* baselines /*[n_cams-1][7]*/);
KB_API kb_status kb_get_set_poses(kb_handle* h, double* set_poses /*[n_sets][7]; other ranks' sets untouched*/);
/* replace the measurements (same structure) — used by the end-to-end bench leg to time host→device per step */
KB_API kb_status kb_set_observations(kb_handle* h, const double* y_u, const double* y_v);
/* kb_set_observations + kb_evaluate_error in one call, with the host->device copy of the measurements (pinned host memory
 * for real overlap) pipelined against the evaluation: the terms travel in a few chunks and the kernel starts on each chunk
 * as it lands.  Same results as the two separate calls up to the summation order of the cost. */
KB_API kb_status kb_evaluate_error_streamed(kb_handle* h, const double* y_u, const double* y_v, int32_t use_m_estimator, double* out_cost);
/* Double buffering for callers that feed a new batch of measurements per step (e.g. an incremental estimator): prefetch starts
 * the host->device copy of the NEXT batch into a second set of device buffers on a copy stream and returns at once (the host
 * arrays must stay valid and pinned until the matching commit); commit makes that batch the current one (stream-ordered, no
 * host synchronisation).  The copy overlaps with whatever runs between the two calls. */
KB_API kb_status kb_prefetch_observations(kb_handle* h, const double* y_u, const double* y_v);
KB_API kb_status kb_commit_observations(kb_handle* h);
/* The same three entry points for measurements held in SINGLE precision - the type a corner detector delivers (the reference keeps
 * image points as cv::Point2f: aslam_cv/aslam_cameras/include/aslam/cameras/GridCalibrationTargetObservation.hpp) and widens to double
 * when it builds the error terms (K2/include/kalibr2/CameraCalibrator.hpp:245-264).  Here the widening - exact - happens on the device,
 * so half the bytes cross PCIe; results are bit-identical to passing the widened doubles. */
KB_API kb_status kb_set_observations_f32(kb_handle* h, const float* y_u, const float* y_v);
KB_API kb_status kb_evaluate_error_streamed_f32(kb_handle* h, const float* y_u, const float* y_v, int32_t use_m_estimator, double* out_cost);
KB_API kb_status kb_prefetch_observations_f32(kb_handle* h, const float* y_u, const float* y_v);
/* terms whose projection bailed out before writing y_hat (zero-weighted here) since creation */
KB_API int64_t kb_num_invalid_terms(kb_handle* h);
/* reset state to the initial guess given at kb_create */
KB_API kb_status kb_reset_state(kb_handle* h);

/* ---- instrumentation ------------------------------------------------------ */
/* number of this library's kernels launched since creation (bench.py's gpu_launches) */
KB_API int64_t kb_kernel_launches(const kb_handle* h);
/* device time in ms of the last call's stages measured with CUDA events on the library's stream:
 *  [0] evaluate  [1] linearise+assemble  [2] expand  [3] schur  [4] reduced solve  [5] backsub  [6] update
 *  [7] materialising linearise */
#define KB_NUM_STAGES 8
KB_API kb_status kb_get_stage_ms(kb_handle* h, double* ms /*[KB_NUM_STAGES]*/);
/* accumulated since the last kb_enable_stage_timing(h, 1): total ms and number of timed calls per stage */
KB_API kb_status kb_get_stage_totals(kb_handle* h, double* total_ms /*[KB_NUM_STAGES]*/, int64_t* calls /*[KB_NUM_STAGES]*/);
/* on = 0: off; 1: every stage; > 1: only the stages s whose bit (s + 1) is set (an event pair per timed stage sits between the
 * kernels of an iteration, so timing fewer stages perturbs the iteration less).  Up to 64 measurements per stage may wait for the
 * next synchronisation (kb_iterate without a result pointer). */
KB_API kb_status kb_enable_stage_timing(kb_handle* h, int32_t on);
KB_API void* kb_cuda_stream(kb_handle* h); /* cudaStream_t the library launches on */

#ifdef __cplusplus
}
#endif
#endif /* KALIBR_B200_H_ */
