"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by kalibr_b200/).

CPU restatement (numpy, dense) of the incremental estimator's numerical core (IC = aslam_incremental_calibration/
incremental_calibration):

  aslam::calibration::LinearSolver::solve            IC/src/core/LinearSolver.cpp:299-463
  columnScalingMatrix, rankTol, estimateNumericalRank, svGap, solveSVD   IC/src/algorithms/linalg.cpp:128-152, 244-282, 426-443
  Optimizer2::optimize with GaussNewtonTrustRegionPolicy   BE/src/Optimizer2.cpp:183-273, BE/src/GaussNewtonTrustRegionPolicy.cpp:18-40
  IncrementalEstimator::addBatch (accept / reject)   IC/src/core/IncrementalEstimator.cpp:338-540

Third-party arithmetic: the reference factorises the pose columns with SuiteSparseQR (libsuitesparse-dev, unpinned) and takes the
SVD of the reduced matrix with Eigen; both are replaced by numpy's dense QR / SVD, which compute the same mathematical objects
(the orthogonal complement of the pose columns, the singular triplets of Omega).  PARITY, part by part: PINNED against the reference's own
compiled code are the Gauss-Newton policy and loop (tests/test_reference_gauss_newton_pin_cpu.py: Optimizer2 + GaussNewtonTrustRegionPolicy
from their sources), the merged problem's design-variable order (same file: the reference's OptimizationProblem / IncrementalOptimizationProblem
containers) and the decision numerics - rankTol, estimateNumericalRank, svGap, columnScalingMatrix, solveSVD
(tests/test_reference_linalg_pin_cpu.py: IC/src/algorithms/linalg.cpp from its source).  UNPINNED: the sparse QR elimination that produces
Omega and addBatch's accept / reject bookkeeping (they need SuiteSparseQR); those are held through properties (tests/test_estimator_cpu.py):
the truncated solve equals the plain least-squares solution when the system has full rank, and the minimum-norm solution on the
observable subspace when not.
"""
from __future__ import annotations

import numpy as np

EPS = float(np.finfo(float).eps)


def dense_from_ccs(col_ptr, row_idx, vals, n_cols):
    J = np.zeros((col_ptr.size - 1, n_cols))
    for r in range(col_ptr.size - 1):
        J[r, row_idx[col_ptr[r]:col_ptr[r + 1]]] = vals[col_ptr[r]:col_ptr[r + 1]]
    return J


def calibration_columns(problem):
    """Columns of the calibration group (intrinsics + baselines: what IncrementalEstimator marginalises onto) and of the rest."""
    col, dims, labels = problem.dv_layout()
    cal = [c + i for c, d, l in zip(col, dims, labels) if l[0] in ("proj", "dist", "baseline_q", "baseline_t") for i in range(d)]
    rest = [c + i for c, d, l in zip(col, dims, labels) if l[0] in ("set_q", "set_t") for i in range(d)]
    return np.array(cal, int), np.array(rest, int)


def column_scaling(A, eps):
    """columnScalingMatrix: 1 / column norm, 0 for columns below sqrt(rows * eps)."""
    norm = np.sqrt((A * A).sum(0))
    tol = np.sqrt(A.shape[0] * eps)
    with np.errstate(divide="ignore"):
        return np.where(norm < tol, 0.0, 1.0 / norm)


def numerical_rank(sv, eps_svd=EPS, svd_tol=-1.0):
    """(tolerance, rank, gap) of a spectrum as LinearSolver::solve / analyzeMarginal decide them (IC/src/core/LinearSolver.cpp:427-431 over
    IC/src/algorithms/linalg.cpp:244-282: rankTol = sv[0] * eps * n unless a tolerance is given; estimateNumericalRank counts down from the
    smallest singular value and never goes below 1; svGap = sv[rank - 1] / sv[rank], infinite at full rank).  Pinned against the
    reference's own functions: tests/test_reference_linalg_pin_cpu.py."""
    sv = np.asarray(sv, float)
    tol = svd_tol if svd_tol != -1.0 else sv[0] * eps_svd * len(sv)
    rank = len(sv)
    for i in range(len(sv) - 1, 0, -1):
        if sv[i] > tol:
            break
        rank -= 1
    with np.errstate(divide="ignore", invalid="ignore"):
        gap = sv[rank - 1] / sv[rank] if rank < len(sv) else np.inf
    return tol, rank, gap


def svd_truncated_solve(Omega, b_r, eps_svd=EPS, svd_tol=-1.0):
    """analyzeSVD + solveSVD (linalg.cpp:412-444): x = V_r diag(1 / sv_r) U_r^T b over the numerical rank r.  (x, sv, tolerance, rank, gap)"""
    U, sv, Vt = np.linalg.svd(Omega)
    tol, rank, gap = numerical_rank(sv, eps_svd, svd_tol)
    x_r = Vt[:rank].T @ ((U[:, :rank].T @ b_r) / sv[:rank])
    return x_r, sv, tol, rank, gap


def linear_solver_solve(J, b, cal, rest, column_scaling_on=False, eps_norm=EPS, eps_svd=EPS, svd_tol=-1.0):
    """LinearSolver::solve: x minimising |J x - b| with the calibration block cut at the numerical rank of Omega.
    Returns (x, info) with info = dict(rank, tolerance, sv_gap, singular_values)."""
    A_l, A_r = J[:, rest].copy(), J[:, cal].copy()
    G_l = column_scaling(A_l, eps_norm) if column_scaling_on else np.ones(A_l.shape[1])
    G_r = column_scaling(A_r, eps_norm) if column_scaling_on else np.ones(A_r.shape[1])
    A_l *= G_l
    A_r *= G_r
    Q, R = np.linalg.qr(A_l)  # thin QR of the pose columns (full column rank: every pose is observed)
    ArtQ = A_r.T @ Q
    Omega = A_r.T @ A_r - ArtQ @ ArtQ.T
    b_r = A_r.T @ b - ArtQ @ (Q.T @ b)
    x_r, sv, tol, rank, gap = svd_truncated_solve(Omega, b_r, eps_svd, svd_tol)
    x_l = np.linalg.solve(R, Q.T @ (b - A_r @ x_r))
    x = np.zeros(J.shape[1])
    x[rest] = G_l * x_l
    x[cal] = G_r * x_r
    return x, dict(rank=rank, tolerance=tol, sv_gap=gap, singular_values=sv)


def system_of(o, problem):
    """(J, b) of the oracle problem at its current state: J dx ~ b with b = -(weighted error), as LinearSolver::solveSystem hands them over."""
    o.evaluate_error()
    J = dense_from_ccs(*o.jacobian_ccs(), o.jcols)
    return J, o.error_vector().copy()


def gauss_newton_optimize(o, problem, solver_kw, max_iterations=20, conv_dx=1e-3, conv_dj=1e-3):
    """Optimizer2::optimize with the Gauss-Newton policy (build + solve every iteration, never revert) and the estimator's solver.
    Returns dict(j_start, j_final, iterations, dx_final, dj_final, trace)."""
    cal, rest = calibration_columns(problem)
    J_cost = o.evaluate_error()
    p_J = J_cost
    out = dict(j_start=J_cost, iterations=0, trace=[])
    delta_x, delta_j = conv_dx + 1.0, conv_dj + 1.0
    while out["iterations"] < max_iterations and delta_x > conv_dx and abs(delta_j) > conv_dj:
        J, b = system_of(o, problem)
        dx, info = linear_solver_solve(J, b, cal, rest, **solver_kw)
        out["last_jacobian"], out["last_solve"] = J, info  # what analyzeMarginal() and getSVDRank() see after optimize()
        ratio = info["singular_values"] / info["tolerance"]
        out["rank_margin"] = min(out.get("rank_margin", np.inf), float(np.abs(np.log(ratio[ratio > 0])).min()))  # distance of the closest singular value to the cut
        delta_x = o.apply_dx(dx)
        J_cost = o.evaluate_error()
        delta_j = p_J - J_cost
        p_J = J_cost
        out["iterations"] += 1
        out["trace"].append((J_cost, delta_x))
    out.update(j_final=p_J, dx_final=delta_x, dj_final=delta_j)
    return out


def analyze_marginal(o, problem, eps_svd=EPS, svd_tol=-1.0):
    """LinearSolver::analyzeMarginal on the UNSCALED system (LinearSolver.cpp:466-528): singular values of Omega, rank, log2 sum."""
    cal, rest = calibration_columns(problem)
    J, _ = system_of(o, problem)
    Q, _ = np.linalg.qr(J[:, rest])
    A_r = J[:, cal]
    ArtQ = A_r.T @ Q
    sv = np.linalg.svd(A_r.T @ A_r - ArtQ @ ArtQ.T, compute_uv=False)
    tol = svd_tol if svd_tol != -1.0 else sv[0] * eps_svd * len(sv)
    rank = len(sv)
    for i in range(len(sv) - 1, 0, -1):
        if sv[i] > tol:
            break
        rank -= 1
    return dict(rank=rank, tolerance=tol, singular_values=sv, sv_log2_sum=float(np.log2(sv[:rank]).sum()))


def marginal_singular_values(J, cal, rest):
    """Singular values of Omega for the UNSCALED Jacobian J (LinearSolver::analyzeMarginal, LinearSolver.cpp:466-516)."""
    Q, _ = np.linalg.qr(J[:, rest])
    A_r = J[:, cal]
    ArtQ = A_r.T @ Q
    return np.linalg.svd(A_r.T @ A_r - ArtQ @ ArtQ.T, compute_uv=False)


class OracleIncrementalEstimator:
    """IncrementalEstimator::addBatch (IC/src/core/IncrementalEstimator.cpp:338-540) on flattened batches: a batch is one synced set
    {camera: (corner_id, y_u, y_v)} with a target-pose guess.  Quirks kept: analyzeMarginal() runs on the Jacobian of the LAST
    Gauss-Newton iteration (one update behind the final state) and, because a solve has run, keeps that solve's rank — the rank of the
    SCALED system — for the log2 sum over the UNSCALED singular values (LinearSolver.cpp:517-523, 196-200)."""

    def __init__(self, oracle_api, cam_model, cam_params, baselines, target_points, info_gain_delta=0.2, check_validity=False,
                 solver_kw=None, max_iterations=20):
        self.oa = oracle_api
        self.cam_model = np.asarray(cam_model, np.int32)
        self.cam_params = np.array(cam_params, float)
        self.baselines = np.array(baselines, float).reshape(-1, 7)
        self.target_points = np.asarray(target_points, float)
        self.info_gain_delta, self.check_validity, self.max_iterations = info_gain_delta, check_validity, max_iterations
        self.solver_kw = solver_kw or dict(column_scaling_on=True, eps_svd=1e-6)
        self.batches, self.poses = [], []
        self.sv_log2_sum, self.rank_theta, self.information_gain = 0.0, -1, 0.0

    def _problem(self, batches, poses):
        from kalibr_b200.problem import ORDER_BATCH, Problem  # the merged incremental problem: poses, baselines, intrinsics

        vs, vc, vb, cid, yu, yv = [], [], [0], [], [], []
        for s, b in enumerate(batches):
            for k in sorted(b):
                c, u, v = b[k]
                vs.append(s); vc.append(k)
                cid.extend(c); yu.extend(u); yv.extend(v)
                vb.append(len(cid))
        return Problem(ORDER_BATCH, self.cam_model, self.cam_params, self.baselines, np.array(poses, float).reshape(-1, 7), self.target_points,
                       np.array(vs, np.int32), np.array(vc, np.int32), np.array(vb, np.int64), np.array(yu, float), np.array(yv, float),
                       np.array(cid, np.int32))

    def add_batch(self, batch, pose_guess, force=False):
        p = self._problem(self.batches + [batch], self.poses + [np.asarray(pose_guess, float)])
        o = self.oa.OracleProblem(p)
        r = gauss_newton_optimize(o, p, self.solver_kw, max_iterations=self.max_iterations)
        cal, rest = calibration_columns(p)
        sv = marginal_singular_values(r["last_jacobian"], cal, rest)
        rank = r["last_solve"]["rank"]
        sv_log2_sum = float(np.log2(sv[:rank]).sum())
        gain = 0.5 * (sv_log2_sum - self.sv_log2_sum)
        valid = not (self.check_validity and (r["iterations"] == self.max_iterations or r["j_final"] >= r["j_start"]))
        keep = ((gain > self.info_gain_delta or rank > self.rank_theta) and valid) or force
        ret = dict(batch_accepted=bool(keep), information_gain=gain, rank_theta=rank, rank_theta_deficiency=len(sv) - rank,
                   svd_tolerance=r["last_solve"]["tolerance"], singular_values=sv, rank_margin=r["rank_margin"], num_iterations=r["iterations"], j_start=r["j_start"],
                   j_final=r["j_final"])
        if keep:
            self.information_gain, self.sv_log2_sum, self.rank_theta = gain, sv_log2_sum, rank
            self.cam_params, self.baselines = o.camera_params(), o.baselines()
            self.batches.append(batch)
            self.poses = list(o.set_poses())
        return ret
