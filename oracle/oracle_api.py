"""ctypes binding of the CPU oracle (oracle/libkalibr_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by
kalibr_b200/.  PARITY: the per-term part (camera models, SE(3) helpers, the expression tree of a reprojection term, M-estimator weights) is
PINNED against the reference's own code (oracle/ref_pin.cpp compiled from /root/reference against the stand-in headers of oracle/ref_shim/;
tests/golden/reference_golden.npz; tests/test_reference_pin_cpu.py), and so are the design-variable order, the Hessian assembly, the
BlockCholesky solver's damping, the LM policy and the optimiser loop (oracle/ref_pin_optimizer.cpp) and the SparseCholesky regime - the
compressed-column J^T, its right-hand side and the loop over SparseCholeskyLinearSystemSolver (tests/golden/reference_sparse_golden.npz).
Only the CHOLMOD factorisation itself is a stand-in there (see oracle/ko_math.hpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkalibr_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("kalibr_oracle.cpp", "ko_backend.hpp", "ko_cameras.hpp", "ko_marginal.hpp", "ko_math.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "kalibr_b200.h"))
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "libkalibr_oracle.so"], check=True)
    return _LIB_PATH


BLOCK_CHOLESKY_KIND, SPARSE_CHOLESKY_KIND = 0, 1  # solver_kind of the reference-pin entry points

REFERENCE_DIR = os.environ.get("KALIBR_REFERENCE", "/root/reference")
_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libkalibr_ref.so")
_ref_lib = None


def build_reference_cameras(force: bool = False):
    """oracle/_ref/libkalibr_ref.so: the REFERENCE's camera models compiled from the sources where they lie (oracle/ref_pin.cpp).
    Returns the path, or None when neither the reference tree nor a built library is there (the GPU box: only prebuilt files travel)."""
    have_ref = os.path.isdir(os.path.join(REFERENCE_DIR, "aslam_cv", "aslam_cameras"))
    if have_ref:  # make decides what is stale (the pin sources and every stand-in header under ref_shim/ are prerequisites)
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "REFERENCE=" + REFERENCE_DIR, "_ref/libkalibr_ref.so"], check=True)
    return _REF_LIB_PATH if os.path.exists(_REF_LIB_PATH) else None


def reference_camera_project(model: int, params, ph):
    """the reference's homogeneousToKeypoint (+ point Jacobian), ...IntrinsicsJacobian, ...DistortionJacobian: same layout as camera_project"""
    global _ref_lib
    if _ref_lib is None:
        path = build_reference_cameras()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libkalibr_ref.so is not built and the reference tree is absent")
        _ref_lib = C.CDLL(path)
        _ref_lib.ref_camera_project.restype = C.c_int32
        _ref_lib.ref_camera_project.argtypes = [C.c_int32] + [C.c_void_p] * 6
    params = np.ascontiguousarray(np.pad(np.asarray(params, np.float64), (0, 10 - len(params))))
    ph = np.ascontiguousarray(ph, np.float64)
    y = np.zeros(2)
    Jp = np.zeros((2, 4))
    Ji = np.zeros((2, 6))
    Jd = np.zeros((2, 4))
    ok = _ref_lib.ref_camera_project(model, _p(params), _p(ph), _p(y), _p(Jp), _p(Ji), _p(Jd))
    return y, Jp, Ji, Jd, ok


def reference_kinematics(name: str, *args):
    """the reference's sm_kinematics helpers: 'quat2r'(q) -> 3x3, 'update_quat'(q, dq) -> 4, 'box_minus'(p4) -> 4x6, 'box_times'(T 4x4) -> 6x6"""
    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    a = [np.ascontiguousarray(x, np.float64) for x in args]
    shape = {"quat2r": (3, 3), "update_quat": (4,), "box_minus": (4, 6), "box_times": (6, 6)}[name]
    out = np.zeros(shape)
    fn = getattr(_ref_lib, "ref_" + name)
    fn.restype = None
    fn.argtypes = [C.c_void_p] * (len(a) + 1)
    fn(*[_p(x) for x in a], _p(out))
    return out


def reference_point_chain(set_pose, baselines, p4, chain=None):
    """the reference's expression tree of a reprojection term (oracle/ref_pin.cpp: ref_point_chain): p_c = B_{k-1} ... B_0 inverse(T_set) p_t and the
    Jacobians of (chain @ p_c) with respect to (set q, set t, baseline 0 q, baseline 0 t, ...): returns (p_c [4], J [(1 + k) * 2, rows, 3])"""
    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    baselines = np.asarray(baselines, np.float64).reshape(-1, 7)
    poses = np.ascontiguousarray(np.vstack([np.asarray(set_pose, np.float64).reshape(1, 7), baselines]))
    p4 = np.ascontiguousarray(p4, np.float64)
    rows = 4 if chain is None else int(np.asarray(chain).shape[0])
    ch = np.ascontiguousarray(chain, np.float64) if chain is not None else np.zeros((4, 4))
    pc = np.zeros(4)
    J = np.zeros((2 * len(poses), rows, 3))
    fn = _ref_lib.ref_point_chain
    fn.restype = C.c_int32
    fn.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    fn(len(baselines), _p(poses), _p(p4), 0 if chain is None else rows, _p(ch), _p(pc), _p(J))
    return pc, J


def reference_m_estimator_weight(kind: int, squared_error: float, p0: float = 0.0, p1: float = 0.999, p2: float = 0.1) -> float:
    """the reference's MEstimator::getWeight (BE/src/MEstimatorPolicies.cpp compiled from its own source); kinds as m_estimator_weight"""
    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    fn = _ref_lib.ref_m_estimator_weight
    fn.restype = C.c_double
    fn.argtypes = [C.c_int32] + [C.c_double] * 4
    return float(fn(kind, p0, p1, p2, squared_error))


def reference_set_weighting(inv_r=None, policy=None):
    """weighting of every term of the problems the reference-pin entry points build from now on (oracle/ref_pin_optimizer.cpp:
    ref_set_weighting): invR through the reference's ErrorTermFs<2>::setInvR, an M-estimator policy (kind, p0, p1, p2 as kb_m_estimator)
    through ErrorTerm::setMEstimatorPolicy.  Call without arguments to go back to invR = I and no policy."""
    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    fn = _ref_lib.ref_set_weighting
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double]
    r = None if inv_r is None else np.ascontiguousarray(inv_r, np.float64).reshape(4)
    kind, p0, p1, p2 = (tuple(policy) + (0.0, 0.999, 0.1)[len(policy) - 1:])[:4] if policy else (0, 0.0, 0.999, 0.1)
    fn(None if r is None else _p(r), int(kind), float(p0), float(p1), float(p2))


def reference_set_trust_region_policy(gauss_newton: bool = False):
    """which trust-region policy reference_optimize runs from now on: the LM policy of the batch drivers (default) or
    GaussNewtonTrustRegionPolicy, the incremental estimator's (oracle/ref_pin_optimizer.cpp: ref_set_trust_region_policy)"""
    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    fn = _ref_lib.ref_set_trust_region_policy
    fn.restype = None
    fn.argtypes = [C.c_int32]
    fn(1 if gauss_newton else 0)


def _reference_problem_arrays(p):
    """the arrays of a kalibr_b200.problem.Problem as the ref_* entry points of oracle/ref_pin_optimizer.cpp take them"""
    cp = np.array(p.cam_params, np.float64, order="C")
    bl = np.array(p.baselines, np.float64, order="C").reshape(-1, 7)
    sp = np.array(p.set_poses, np.float64, order="C")
    cm, vs, vc = (np.ascontiguousarray(a, np.int32) for a in (p.cam_model, p.view_set, p.view_cam))
    vb = np.ascontiguousarray(p.view_begin, np.int64)
    yu, yv, tp = (np.ascontiguousarray(a, np.float64) for a in (p.y_u, p.y_v, p.target_points))
    ci = np.ascontiguousarray(p.corner_id, np.int32)
    keep = (cp, bl, sp, cm, vs, vc, vb, yu, yv, tp, ci)
    args = [len(cm), _p(cm), _p(cp), _p(bl), len(sp), _p(sp), len(tp), _p(tp), len(vs), _p(vs), _p(vc), _p(vb), _p(yu), _p(yv), _p(ci), int(p.driver_order)]
    types = [C.c_int32] + [C.c_void_p] * 3 + [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 6 + [C.c_int32]
    return keep, args, types


def reference_optimize(problem, options=None, solver_kind: int = BLOCK_CHOLESKY_KIND, n_threads: int = 1):
    """the REFERENCE's Optimizer2::optimize (BE/src/Optimizer2.cpp, LevenbergMarquardtTrustRegionPolicy.cpp, BlockCholeskyLinearSystemSolver.cpp
    or - solver_kind 1 - SparseCholeskyLinearSystemSolver.cpp with CompressedColumnJacobianTransposeBuilder / CompressedColumnMatrix / the
    Cholmod wrapper, ErrorTerm / JacobianContainer / OptimizationProblem / SparseBlockMatrix, the expression tree and the camera models, all
    compiled from the reference's sources: oracle/ref_pin_optimizer.cpp) on a kalibr_b200.problem.Problem, design variables in the problem's
    driver order.  Returns (dict(iterations, failed_iterations, j_start, j_final, linear_solver_failure), cam_params, baselines, set_poses)."""
    from kalibr_b200.problem import KbOptimizerOptions

    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    o = options or KbOptimizerOptions.kalibr2_default()
    keep, args, types = _reference_problem_arrays(problem)
    cp, bl, sp = keep[:3]
    out = np.zeros(8)
    fn = _ref_lib.ref_optimize_rig_solver
    fn.restype = C.c_int32
    fn.argtypes = types + [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_void_p]
    rc = fn(*args, int(o.max_iterations), float(o.convergence_delta_x), float(o.convergence_delta_j), float(o.lm_lambda_init), int(solver_kind), int(n_threads), _p(out))
    if rc != 0:
        raise RuntimeError("ref_optimize_rig_solver failed")
    res = dict(iterations=int(out[0]), failed_iterations=int(out[1]), j_start=float(out[2]), j_final=float(out[3]), linear_solver_failure=int(out[4]))
    return res, cp, bl, sp


def reference_sparse_system(problem, lam: float = 10.0, n_threads: int = 1):
    """the SparseCholesky regime's linear system from the REFERENCE's own classes (oracle/ref_pin_optimizer.cpp: ref_sparse_system): J^T in
    compressed-column form as CompressedColumnJacobianTransposeBuilder<int> lays it out, the error vector, rhs = J^T e as
    SparseCholeskyLinearSystemSolver::buildSystem forms it, and dx of one solveSystem under the constant conditioner `lam`.
    dict(col_ptr [rows + 1] int64, row_ind int32, values, e, rhs, dx (None when the solve failed), cost, jcols)"""
    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    keep, args, types = _reference_problem_arrays(problem)
    fn = _ref_lib.ref_sparse_system
    fn.restype = C.c_int32
    fn.argtypes = types + [C.c_int32, C.c_double] + [C.c_void_p] * 8
    sizes = np.zeros(4, np.int64)
    cost = np.zeros(1)
    if fn(*args, int(n_threads), float(lam), _p(sizes), None, None, None, None, None, None, _p(cost)) != 0:
        raise RuntimeError("ref_sparse_system failed")
    jcols, jrows, nnz = (int(v) for v in sizes[:3])
    col_ptr, row_ind, values = np.zeros(jrows + 1, np.int64), np.zeros(nnz, np.int32), np.zeros(nnz)
    e, rhs, dx = np.zeros(jrows), np.zeros(jcols), np.zeros(jcols)
    if fn(*args, int(n_threads), float(lam), _p(sizes), _p(col_ptr), _p(row_ind), _p(values), _p(e), _p(rhs), _p(dx), _p(cost)) != 0:
        raise RuntimeError("ref_sparse_system failed")
    return dict(col_ptr=col_ptr, row_ind=row_ind, values=values, e=e, rhs=rhs, dx=dx if sizes[3] else None, cost=float(cost[0]), jcols=jcols)


def reference_estimator_problem(problem, options=None, restore_after: bool = False):
    """the incremental estimator's MERGED problem through the REFERENCE's own containers (oracle/ref_pin_optimizer.cpp: ref_estimator_problem):
    one aslam::calibration::OptimizationProblem per synced set filled as kalibr2's CreateBatchProblem fills it, merged by
    IncrementalOptimizationProblem::add, ordered as IncrementalEstimator::orderMarginalizedDesignVariables orders it, then Optimizer2 with the
    Gauss-Newton policy.  Returns (order [n_active, 4] = kind, index, column base, dimension per active design variable in the optimiser's
    order; groups ordering as a tuple; dict of the optimisation's scalars; cam_params, baselines, set_poses).  restore_after: the container's
    saveDesignVariables before and restoreDesignVariables after the optimisation (the estimator's reject path): the returned state is then
    what the reference restored."""
    from kalibr_b200.problem import KbOptimizerOptions

    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    o = options or KbOptimizerOptions.estimator_default()
    keep, args, types = _reference_problem_arrays(problem)
    cp, bl, sp = keep[:3]
    order = np.full((2 * len(sp) + 4 * len(cp) + 8, 4), -7, np.int32)
    counts = np.zeros(2, np.int32)
    out = np.zeros(8)
    fn = _ref_lib.ref_estimator_problem
    fn.restype = C.c_int32
    fn.argtypes = types[:-1] + [C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    if fn(*args[:-1], int(o.max_iterations), float(o.convergence_delta_x), float(o.convergence_delta_j), 1 if restore_after else 0, _p(order), _p(counts), _p(out)) != 0:
        raise RuntimeError("ref_estimator_problem failed")
    res = dict(iterations=int(out[0]), failed_iterations=int(out[1]), j_start=float(out[2]), j_final=float(out[3]), linear_solver_failure=int(out[4]))
    return order[:counts[0]].copy(), tuple(int(c) for c in str(int(counts[1])).zfill(3)), res, cp, bl, sp


_REF_LINALG_PATH = os.path.join(_HERE, "_ref", "libkalibr_ref_linalg.so")
_ref_linalg = None


def reference_linalg():
    """oracle/_ref/libkalibr_ref_linalg.so: IC/src/algorithms/linalg.cpp compiled from the reference's source (oracle/ref_pin_linalg.cpp says what
    in it is reference code and what a stand-in); None when neither the reference tree nor a built library is there"""
    global _ref_linalg
    if _ref_linalg is None:
        if os.path.isdir(os.path.join(REFERENCE_DIR, "aslam_incremental_calibration")):
            subprocess.run(["make", "-C", _HERE, "-s", "REFERENCE=" + REFERENCE_DIR, "_ref/libkalibr_ref_linalg.so"], check=True)
        if not os.path.exists(_REF_LINALG_PATH):
            return None
        _ref_linalg = C.CDLL(_REF_LINALG_PATH)
    return _ref_linalg


def reference_linalg_rank(sv, eps, svd_tol=-1.0):
    """(tolerance, rank, gap) from the reference's rankTol / estimateNumericalRank / svGap"""
    sv = np.ascontiguousarray(sv, np.float64)
    out = np.zeros(3)
    fn = reference_linalg().ref_linalg_rank
    fn.restype = C.c_int32
    fn.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_void_p]
    if fn(_p(sv), len(sv), float(eps), float(svd_tol), _p(out)) != 0:
        raise RuntimeError("ref_linalg_rank failed")
    return float(out[0]), int(out[1]), float(out[2])


def reference_linalg_column_scaling(A, eps, eps_qr):
    """(G, qr_tol): the reference's columnScalingMatrix (1 / column norm, 0 below sqrt(rows * eps)) and qrTol of the dense matrix A"""
    A = np.ascontiguousarray(A, np.float64)
    G, qr = np.zeros(A.shape[1]), np.zeros(1)
    fn = reference_linalg().ref_linalg_column_scaling
    fn.restype = C.c_int32
    fn.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    if fn(_p(A), A.shape[0], A.shape[1], float(eps), float(eps_qr), _p(G), _p(qr)) != 0:
        raise RuntimeError("ref_linalg_column_scaling failed")
    return G, float(qr[0])


def reference_linalg_svd_solve(Omega, b, eps, svd_tol=-1.0):
    """the reference's analyzeSVD (SVD itself: stand-in) + rank decision + solveSVD on a dense symmetric Omega: (x, sv, tolerance, rank, gap)"""
    Omega, b = np.ascontiguousarray(Omega, np.float64), np.ascontiguousarray(b, np.float64)
    n = len(b)
    sv, x, out = np.zeros(n), np.zeros(n), np.zeros(3)
    fn = reference_linalg().ref_linalg_svd_solve
    fn.restype = C.c_int32
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    if fn(_p(Omega), _p(b), n, float(eps), float(svd_tol), _p(sv), _p(x), _p(out)) != 0:
        raise RuntimeError("ref_linalg_svd_solve failed")
    return x, sv, float(out[0]), int(out[1]), float(out[2])


def reference_time_evaluate_build(problem, n_threads: int = 4, repeats: int = 1, solver_kind: int = BLOCK_CHOLESKY_KIND):
    """seconds the REFERENCE's own code (oracle/ref_pin_optimizer.cpp: ref_time_evaluate_build_solver) spends on one Optimizer2::evaluateError and
    one buildSystem of `problem` - BlockCholeskyLinearSystemSolver's (serial Hessian assembly) or, solver_kind 1, SparseCholeskyLinearSystemSolver's
    (Kalibr2's default: threaded materialisation of the compressed-column J^T + rhs): dict(setup_s, evaluate_s, build_s, cost)"""
    reference_camera_project(0, [1, 1, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1])  # loads the library
    keep, args, types = _reference_problem_arrays(problem)
    out = np.zeros(4)
    fn = _ref_lib.ref_time_evaluate_build_solver
    fn.restype = C.c_int32
    fn.argtypes = types + [C.c_int32] * 3 + [C.c_void_p]
    if fn(*args, int(solver_kind), int(n_threads), int(repeats), _p(out)) != 0:
        raise RuntimeError("ref_time_evaluate_build_solver failed")
    return dict(setup_s=float(out[0]), evaluate_s=float(out[1]), build_s=float(out[2]), cost=float(out[3]))


def kinematics(name: str, *args):
    """the oracle's restatement of the same helpers (ko_math.hpp)"""
    a = [np.ascontiguousarray(x, np.float64) for x in args]
    shape = {"quat2r": (3, 3), "update_quat": (4,), "box_minus": (4, 6), "box_times": (6, 6)}[name]
    out = np.zeros(shape)
    fn = getattr(lib(), "ko_" + name)
    fn.restype = None
    fn.argtypes = [C.c_void_p] * (len(a) + 1)
    fn(*[_p(x) for x in a], _p(out))
    return out


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.ko_last_error.restype = C.c_char_p
        L.ko_create.restype = C.c_void_p
        L.ko_create.argtypes = [C.c_void_p, C.c_int]
        L.ko_destroy.argtypes = [C.c_void_p]
        for name in ("ko_jrows", "ko_jcols"):
            getattr(L, name).restype = C.c_int64
            getattr(L, name).argtypes = [C.c_void_p]
        L.ko_num_design_variables.restype = C.c_int32
        L.ko_num_design_variables.argtypes = [C.c_void_p]
        L.ko_get_dv_layout.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ko_evaluate_error.restype = C.c_double
        L.ko_evaluate_error.argtypes = [C.c_void_p, C.c_int]
        L.ko_get_error_vector.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_build_system.argtypes = [C.c_void_p, C.c_int]
        L.ko_set_constant_conditioner.argtypes = [C.c_void_p, C.c_double]
        L.ko_solve_system.restype = C.c_int32
        L.ko_solve_system.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_get_rhs.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_apply_state_update.restype = C.c_double
        L.ko_apply_state_update.argtypes = [C.c_void_p]
        L.ko_revert_last_state_update.argtypes = [C.c_void_p]
        L.ko_apply_dx.restype = C.c_double
        L.ko_apply_dx.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_optimize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ko_get_trace.restype = C.c_int32
        L.ko_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.ko_get_jacobian_ccs.restype = C.c_int64
        L.ko_get_jacobian_ccs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ko_get_hessian_blocks.restype = C.c_int32
        L.ko_get_hessian_blocks.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.ko_get_camera_params.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_get_baselines.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_get_set_poses.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_camera_project.restype = C.c_int32
        L.ko_camera_project.argtypes = [C.c_int32] + [C.c_void_p] * 6
        L.ko_quat2r.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_update_quat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ko_inverse4.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_analyze_marginal.restype = C.c_int32
        L.ko_analyze_marginal.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.ko_set_inv_r.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_set_m_estimator.restype = C.c_double
        L.ko_set_m_estimator.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.ko_set_use_m_estimator.argtypes = [C.c_void_p, C.c_int]
        L.ko_chi2_inv_cdf.restype = C.c_double
        L.ko_chi2_inv_cdf.argtypes = [C.c_double, C.c_int]
        L.ko_m_estimator_weight.restype = C.c_double
        L.ko_m_estimator_weight.argtypes = [C.c_int] + [C.c_double] * 4
        L.ko_matrix_sqrt2.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_reprojection_statistics.argtypes = [C.c_void_p, C.c_void_p]
        L.ko_time_iteration.restype = C.c_int32
        L.ko_time_iteration.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


BLOCK_CHOLESKY, SPARSE_CHOLESKY, BLOCK_CHOLESKY_DENSE = 0, 1, 2


class OracleProblem:
    """CPU restatement of Optimizer2 + LinearSystemSolver on a kalibr_b200.problem.Problem."""

    def __init__(self, problem, solver_kind: int = BLOCK_CHOLESKY, n_threads: int = 4):
        self._problem = problem  # keep arrays alive
        self._desc = problem.desc()
        self.n_threads = n_threads
        self._h = lib().ko_create(C.byref(self._desc), solver_kind)
        if not self._h:
            raise RuntimeError("ko_create failed: " + lib().ko_last_error().decode())
        self.jrows = lib().ko_jrows(self._h)
        self.jcols = lib().ko_jcols(self._h)
        self.n_dv = lib().ko_num_design_variables(self._h)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ko_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dv_layout(self):
        col = np.zeros(self.n_dv, np.int32)
        dims = np.zeros(self.n_dv, np.int32)
        lib().ko_get_dv_layout(self._h, _p(col), _p(dims))
        return col, dims

    def evaluate_error(self) -> float:
        return lib().ko_evaluate_error(self._h, self.n_threads)

    def error_vector(self) -> np.ndarray:
        e = np.zeros(self.jrows)
        lib().ko_get_error_vector(self._h, _p(e))
        return e

    def build_system(self):
        lib().ko_build_system(self._h, self.n_threads)

    def set_constant_conditioner(self, lam: float):
        lib().ko_set_constant_conditioner(self._h, lam)

    def solve_system(self):
        dx = np.zeros(self.jcols)
        ok = lib().ko_solve_system(self._h, _p(dx))
        return dx, bool(ok)

    def rhs(self) -> np.ndarray:
        r = np.zeros(self.jcols)
        lib().ko_get_rhs(self._h, _p(r))
        return r

    def apply_state_update(self) -> float:
        return lib().ko_apply_state_update(self._h)

    def apply_dx(self, dx) -> float:
        """applyStateUpdate with a step computed by the caller; returns max|dx|."""
        dx = np.ascontiguousarray(dx, np.float64)
        assert dx.size == self.jcols
        return lib().ko_apply_dx(self._h, _p(dx))

    def revert_last_state_update(self):
        lib().ko_revert_last_state_update(self._h)

    def optimize(self, options=None):
        from kalibr_b200.problem import KbOptimizerOptions, KbSolution

        options = options or KbOptimizerOptions.kalibr2_default()
        sol = KbSolution()
        lib().ko_optimize(self._h, C.byref(options), self.n_threads, C.byref(sol))
        n = lib().ko_get_trace(self._h, None, 0)
        tr = np.zeros((n, 3))
        if n:
            lib().ko_get_trace(self._h, _p(tr), n)
        return sol, tr

    def jacobian_ccs(self):
        nnz = lib().ko_get_jacobian_ccs(self._h, self.n_threads, None, None, None)
        col_ptr = np.zeros(self.jrows + 1, np.int64)
        row_idx = np.zeros(nnz, np.int32)
        vals = np.zeros(nnz)
        lib().ko_get_jacobian_ccs(self._h, self.n_threads, _p(col_ptr), _p(row_idx), _p(vals))
        return col_ptr, row_idx, vals

    def analyze_marginal(self, options=None):
        """LinearSolver::analyzeMarginal restated (oracle/ko_marginal.hpp): (result, singular values, V, DV columns, Omega)."""
        from kalibr_b200.problem import KbMarginalOptions, KbMarginalResult

        o = options or KbMarginalOptions.default()
        n = self._problem.n_c
        res = KbMarginalResult()
        sv, V, cols, om = np.zeros(n), np.zeros((n, n)), np.zeros(n, np.int32), np.zeros((n, n))
        got = lib().ko_analyze_marginal(self._h, self.n_threads, C.byref(o), C.byref(res), _p(sv), _p(V), _p(cols), _p(om))
        assert got == n
        return res, sv, V, cols, om

    def hessian_blocks(self):
        nb = C.c_int64()
        nv = C.c_int64()
        lib().ko_get_hessian_blocks(self._h, C.byref(nb), C.byref(nv), None, None, None, None)
        col_ptr = np.zeros(self.n_dv + 1, np.int64)
        block_row = np.zeros(nb.value, np.int32)
        value_ptr = np.zeros(nb.value, np.int64)
        values = np.zeros(nv.value)
        lib().ko_get_hessian_blocks(self._h, C.byref(nb), C.byref(nv), _p(col_ptr), _p(block_row), _p(value_ptr), _p(values))
        return col_ptr, block_row, value_ptr, values

    def camera_params(self) -> np.ndarray:
        out = np.zeros((self._problem.n_cams, 10))
        lib().ko_get_camera_params(self._h, _p(out))
        return out

    def baselines(self) -> np.ndarray:
        out = np.zeros((max(self._problem.n_cams - 1, 0), 7))
        if out.size:
            lib().ko_get_baselines(self._h, _p(out))
        return out

    def set_poses(self) -> np.ndarray:
        out = np.zeros((self._problem.n_sets, 7))
        lib().ko_get_set_poses(self._h, _p(out))
        return out

    def set_inv_r(self, inv_r):
        """ErrorTermFs<2>::setInvR on every term."""
        a = np.ascontiguousarray(inv_r, np.float64).reshape(2, 2)
        lib().ko_set_inv_r(self._h, _p(a))

    def set_m_estimator(self, kind: int, p0: float = 0.0, p1: float = 0.999, p2: float = 0.1) -> float:
        """ErrorTerm::setMEstimatorPolicy on every term; returns the policy's parameter (epsilon for Blake-Zisserman)."""
        return lib().ko_set_m_estimator(self._h, kind, p0, p1, p2)

    def set_use_m_estimator(self, on: bool):
        lib().ko_set_use_m_estimator(self._h, 1 if on else 0)

    def reprojection_statistics(self) -> np.ndarray:
        out = np.zeros((self._problem.n_cams, 6))
        lib().ko_reprojection_statistics(self._h, _p(out))
        return out

    def time_iteration(self, lam: float = 10.0):
        t = np.zeros(3)
        ok = lib().ko_time_iteration(self._h, self.n_threads, lam, _p(t))
        return t, bool(ok)


def camera_project(model: int, params, ph):
    params = np.ascontiguousarray(np.pad(np.asarray(params, np.float64), (0, 10 - len(params))))
    ph = np.ascontiguousarray(ph, np.float64)
    y = np.zeros(2)
    Jp = np.zeros((2, 4))
    Ji = np.zeros((2, 6))
    Jd = np.zeros((2, 4))
    ok = lib().ko_camera_project(model, _p(params), _p(ph), _p(y), _p(Jp), _p(Ji), _p(Jd))
    return y, Jp, Ji, Jd, ok


def quat2r(q):
    q = np.ascontiguousarray(q, np.float64)
    R = np.zeros((3, 3))
    lib().ko_quat2r(_p(q), _p(R))
    return R


def update_quat(q, dq):
    q = np.ascontiguousarray(q, np.float64)
    dq = np.ascontiguousarray(dq, np.float64)
    out = np.zeros(4)
    lib().ko_update_quat(_p(q), _p(dq), _p(out))
    return out


def inverse4(M):
    M = np.ascontiguousarray(M, np.float64)
    out = np.zeros((4, 4))
    lib().ko_inverse4(_p(M), _p(out))
    return out


def chi2_inv_cdf(p: float, df: int) -> float:
    return lib().ko_chi2_inv_cdf(p, df)


def m_estimator_weight(kind: int, squared_error: float, p0: float = 0.0, p1: float = 0.999, p2: float = 0.1) -> float:
    return lib().ko_m_estimator_weight(kind, p0, p1, p2, squared_error)


def matrix_sqrt2(A):
    A = np.ascontiguousarray(A, np.float64).reshape(2, 2)
    S = np.zeros((2, 2))
    lib().ko_matrix_sqrt2(_p(A), _p(S))
    return S
