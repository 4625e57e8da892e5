"""CPU tests of the oracle: the reference's own property tests for this path, re-expressed.

  CameraGeometryTestHarness.hpp:172-332  analytic point / intrinsics / distortion Jacobians vs finite differences on the
                                         getTestGeometry() parameter sets
  BE/test/TestOptimizer.cpp:101-120      H == J^T J, rhs == -J^T e
  BE/test/test_sparse_matrix_functions.cpp:49-95   Schur elimination == dense solve
  BE/test/LinearSolverTests.cpp:18-141   two LinearSystemSolver implementations agree (J, e, rhs, dx)
"""
import os

import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import (DS_NONE, EUCM_NONE, MODEL_D, MODEL_P, OMNI_NONE, OMNI_RADTAN, PINHOLE_EQUI, PINHOLE_FOV, PINHOLE_RADTAN,
                                 KbOptimizerOptions)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = [PINHOLE_RADTAN, PINHOLE_EQUI, OMNI_RADTAN, EUCM_NONE, DS_NONE, PINHOLE_FOV, OMNI_NONE]


def test_kinematics_helpers(oracle_lib):
    from scipy.spatial.transform import Rotation as R

    rng = np.random.default_rng(1)
    for _ in range(20):
        a = rng.normal(size=3)
        q = synthetic.axis_angle_to_quat(a)
        Ro = oracle_lib.quat2r(q)
        assert np.abs(Ro - R.from_rotvec(a).as_matrix()).max() < 1e-14
        assert np.abs(Ro - synthetic.quat2r(q)).max() < 1e-15
        assert np.abs(synthetic.r_to_quat(Ro) - q * np.sign(q[3])).max() < 1e-12
        # updateQuat(q, d): q' = dq (x) q in the JPL convention, i.e. C(q') = C(dq) C(q) with C(dq) = exp(-[d]x)
        d = 0.1 * rng.normal(size=3)
        q2 = oracle_lib.update_quat(q, d)
        assert abs(np.linalg.norm(q2) - 1) < 1e-14
        assert np.abs(oracle_lib.quat2r(q2) - R.from_rotvec(-d).as_matrix() @ Ro).max() < 1e-13
        assert np.abs(q2 - synthetic.quat_mul_update(q[None], d[None])[0]).max() < 1e-15
        T = np.eye(4)
        T[:3, :3] = Ro
        T[:3, 3] = rng.normal(size=3)
        assert np.abs(oracle_lib.inverse4(T) @ T - np.eye(4)).max() < 1e-13
    # Taylor branch of axisAngle2quat
    q = np.array([0.0, 0.0, 0.0, 1.0])
    assert np.abs(oracle_lib.update_quat(q, [1e-6, 0, 0]) - [5e-7, 0, 0, 1.0]).max() < 1e-12


def _points(rng, n=60):
    p = np.concatenate([rng.uniform(-0.5, 0.5, size=(n, 2)), rng.uniform(0.5, 2.0, size=(n, 1))], axis=1)
    return np.concatenate([p, np.ones((n, 1))], axis=1)


@pytest.mark.parametrize("model", MODELS)
def test_camera_point_jacobian_vs_finite_differences(oracle_lib, model):
    params = np.asarray(synthetic.TRUTH_PARAMS[model][0])
    rng = np.random.default_rng(model)
    for ph in _points(rng):
        y, Jp, _, _, ok = oracle_lib.camera_project(model, params, ph)
        assert ok == 1
        u, v, _ = synthetic.project(model, params, ph[None, :3])
        assert abs(y[0] - u[0]) < 1e-9 and abs(y[1] - v[0]) < 1e-9  # generator and oracle agree on the forward model
        assert np.all(Jp[:, 3] == 0.0)  # 4th homogeneous column is always zero
        h = 1e-6
        for j in range(3):
            dp = np.zeros(4)
            dp[j] = h
            yp = oracle_lib.camera_project(model, params, ph + dp)[0]
            ym = oracle_lib.camera_project(model, params, ph - dp)[0]
            fd = (yp - ym) / (2 * h)
            assert np.abs(fd - Jp[:, j]).max() < 1e-5 * max(1.0, np.abs(Jp[:, j]).max())


@pytest.mark.parametrize("model", MODELS)
def test_camera_parameter_jacobians_vs_finite_differences(oracle_lib, model):
    P, D = MODEL_P[model], MODEL_D[model]
    params = np.asarray(synthetic.TRUTH_PARAMS[model][0], float).copy()
    if model == EUCM_NONE:
        params[3] = 395.0  # fu != fv exposes quirk Q4
    rng = np.random.default_rng(10 + model)
    for ph in _points(rng, 30):
        _, _, Ji, Jd, _ = oracle_lib.camera_project(model, params, ph)
        for j in range(P + D):
            h = 1e-6 * max(1.0, abs(params[j]))
            pp, pm = params.copy(), params.copy()
            pp[j] += h
            pm[j] -= h
            fd = (oracle_lib.camera_project(model, pp, ph)[0] - oracle_lib.camera_project(model, pm, ph)[0]) / (2 * h)
            an = Ji[:, j] if j < P else Jd[:, j - P]
            if model == EUCM_NONE and j < 2:
                # Q4: the reference scales row 1 of the alpha/beta columns with fu instead of fv
                fd = fd * np.array([1.0, params[2] / params[3]])
            assert np.abs(fd - an).max() < 2e-5 * max(1.0, np.abs(an).max()), (model, j, fd, an)


def test_equidistant_jacobian_is_nan_on_axis(oracle_lib):
    """Q5: the value path guards r > 1e-8, the generated Jacobian does not."""
    params = np.asarray(synthetic.TRUTH_PARAMS[PINHOLE_EQUI][0])
    y, Jp, _, Jd, _ = oracle_lib.camera_project(PINHOLE_EQUI, params, np.array([0.0, 0.0, 1.0, 1.0]))
    assert np.allclose(y, params[2:4])
    assert np.isnan(Jp[:, :3]).any()


def _dense_from_ccs(col_ptr, row_idx, vals, n_cols):
    J = np.zeros((col_ptr.size - 1, n_cols))
    for r in range(col_ptr.size - 1):
        J[r, row_idx[col_ptr[r]:col_ptr[r + 1]]] = vals[col_ptr[r]:col_ptr[r + 1]]
    return J


def _dense_from_blocks(col_ptr, block_row, value_ptr, values, col, dims):
    n = int(col[-1] + dims[-1])
    H = np.zeros((n, n))
    for c in range(dims.size):
        for b in range(col_ptr[c], col_ptr[c + 1]):
            r = block_row[b]
            blk = values[value_ptr[b]:value_ptr[b] + dims[r] * dims[c]].reshape(dims[c], dims[r]).T  # column-major
            H[col[r]:col[r] + dims[r], col[c]:col[c] + dims[c]] = blk
            H[col[c]:col[c] + dims[c], col[r]:col[r] + dims[r]] = blk.T
    return H


@pytest.mark.parametrize("cfg,n_sets", [(1, 5), (2, 4), (3, 3), (4, 2), (6, 3), (7, 4)])
def test_hessian_is_JtJ_and_rhs_is_minus_Jte(oracle_lib, cfg, n_sets):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    o = oracle_lib.OracleProblem(p, n_threads=2)
    o.evaluate_error()
    cp, ri, jv = o.jacobian_ccs()
    J = _dense_from_ccs(cp, ri, jv, o.jcols)
    e = -o.error_vector()  # _e holds -e
    o.build_system()
    col, dims = o.dv_layout()
    H = _dense_from_blocks(*o.hessian_blocks(), col, dims)
    assert np.abs(H - J.T @ J).max() <= 1e-12 * np.abs(H).max()
    assert np.abs(o.rhs() + J.T @ e).max() <= 1e-12 * np.abs(o.rhs()).max()
    # analytic Jacobians vs a finite difference of the whole term chain through the public update path
    # (ErrorTermTestHarness.hpp:20-35): e(x [+] dx) - e(x) ~= J dx for a tiny step along the LM direction
    base = o.error_vector().copy()
    o2 = oracle_lib.OracleProblem(p, n_threads=1)
    o2.evaluate_error()
    o2.build_system()
    o2.set_constant_conditioner(1e6)  # heavy damping => tiny step
    dxs, ok = o2.solve_system()
    assert ok
    o2.apply_state_update()
    o2.evaluate_error()
    de = (-o2.error_vector()) - (-base)
    lin = J @ dxs
    assert np.abs(de - lin).max() <= 1e-3 * np.abs(lin).max() + 1e-9


@pytest.mark.parametrize("cfg,n_sets", [(1, 6), (2, 5), (3, 4), (5, 2)])
def test_arrow_solve_matches_dense_solve(oracle_lib, cfg, n_sets):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    a = oracle_lib.OracleProblem(p, oracle_lib.BLOCK_CHOLESKY)
    d = oracle_lib.OracleProblem(p, oracle_lib.BLOCK_CHOLESKY_DENSE)
    for o in (a, d):
        o.evaluate_error()
        o.build_system()
        o.set_constant_conditioner(10.0)
    dxa, oka = a.solve_system()
    dxd, okd = d.solve_system()
    assert oka and okd
    assert np.abs(dxa - dxd).max() <= 1e-9 * np.abs(dxd).max()
    # and against numpy on the exported blocks (damping lambda^2 - lambda is what un-augmenting leaves behind: Q2)
    col, dims = a.dv_layout()
    H = _dense_from_blocks(*a.hessian_blocks(), col, dims)
    # the exported H carries the residual lambda^2 - lambda; rebuild the matrix that was actually solved (H0 + lambda^2 I)
    H_solve = H - (100.0 - 10.0) * np.eye(H.shape[0]) + 100.0 * np.eye(H.shape[0])
    ref = np.linalg.solve(H_solve, a.rhs())
    assert np.abs(dxa - ref).max() <= 1e-8 * np.abs(ref).max()


@pytest.mark.parametrize("cfg,n_sets", [(1, 5), (2, 4), (3, 3)])
def test_sparse_and_block_solvers_agree(oracle_lib, cfg, n_sets):
    """compareSolvers<SparseCholesky, BlockCholesky>: J, e, rhs, dx."""
    p = synthetic.make_config(cfg, n_sets=n_sets)
    for threads in (1, 3):
        s = oracle_lib.OracleProblem(p, oracle_lib.SPARSE_CHOLESKY, n_threads=threads)
        b = oracle_lib.OracleProblem(p, oracle_lib.BLOCK_CHOLESKY, n_threads=threads)
        assert abs(s.evaluate_error() - b.evaluate_error()) <= 1e-12 * b.evaluate_error()
        assert np.array_equal(s.error_vector(), b.error_vector())
        s.build_system()
        b.build_system()
        assert np.abs(s.rhs() - b.rhs()).max() <= 1e-10 * np.abs(b.rhs()).max()
        s.set_constant_conditioner(10.0)
        b.set_constant_conditioner(10.0)
        dxs, oks = s.solve_system()
        dxb, okb = b.solve_system()
        assert oks and okb
        assert np.abs(dxs - dxb).max() <= 1e-6 * np.abs(dxb).max()


def test_block_cholesky_leaves_lambda_residual_on_diagonal(oracle_lib):
    """Q2: diag += lambda^2 before the solve, diag -= lambda after it."""
    p = synthetic.make_config(1, n_sets=4)
    o = oracle_lib.OracleProblem(p)
    o.evaluate_error()
    o.build_system()
    col, dims = o.dv_layout()
    H0 = _dense_from_blocks(*o.hessian_blocks(), col, dims)
    o.set_constant_conditioner(10.0)
    o.solve_system()
    H1 = _dense_from_blocks(*o.hessian_blocks(), col, dims)
    assert np.allclose(np.diag(H1) - np.diag(H0), 90.0, rtol=0, atol=1e-6)
    o.set_constant_conditioner(20.0)
    o.solve_system()
    H2 = _dense_from_blocks(*o.hessian_blocks(), col, dims)
    assert np.allclose(np.diag(H2) - np.diag(H0), 90.0 + 380.0, rtol=0, atol=1e-6)
    o.build_system()  # clear(false)
    H3 = _dense_from_blocks(*o.hessian_blocks(), col, dims)
    assert np.abs(H3 - H0).max() <= 1e-9 * np.abs(H0).max()


def test_zero_dim_distortion_blocks_are_in_the_pattern(oracle_lib):
    """Q7: eucm-none / ds-none keep an active 0-dimensional distortion design variable that owns a block index."""
    p = synthetic.make_config(3, n_sets=2)
    o = oracle_lib.OracleProblem(p)
    col, dims = o.dv_layout()
    assert list(dims[:8]) == [5, 4, 6, 0, 6, 0, 4, 4]
    o.evaluate_error()
    o.build_system()
    cp, br, vp, vals = o.hessian_blocks()
    n_zero = sum(1 for c in range(dims.size) for b in range(cp[c], cp[c + 1]) if dims[c] == 0 or dims[br[b]] == 0)
    assert n_zero > 0
    assert vals.size == sum(dims[br[b]] * dims[c] for c in range(dims.size) for b in range(cp[c], cp[c + 1]))


@pytest.mark.parametrize("cfg,n_sets", [(1, 30), (2, 20), (3, 12), (6, 14), (7, 20)])
def test_lm_converges_to_ground_truth(oracle_lib, cfg, n_sets):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    o = oracle_lib.OracleProblem(p, n_threads=4)
    sol, tr = o.optimize(KbOptimizerOptions.kalibr2_default())
    assert sol.linear_solver_failure == 0 and 2 <= sol.iterations < 30
    dof = 2 * p.n_terms - o.jcols
    assert 0.7 < sol.j_final / (dof * 0.3**2) < 1.3  # chi^2 at the noise level
    truth = p.truth["cam_params"]
    est = o.camera_params()
    for k, m in enumerate(p.cam_model):
        f = slice(MODEL_P[m] - 4, MODEL_P[m])  # fu, fv, cu, cv
        assert np.abs(est[k, f] - truth[k, f]).max() < 8.0  # 2 % of f; omni trades xi against f
    # the trace is monotone for accepted steps
    assert np.all(np.diff(tr[:, 0]) <= 1e-9 * tr[0, 0])


def test_empty_and_ragged_views(oracle_lib):
    p = synthetic.make_config(2, n_sets=6, dropout=0.3)
    assert np.unique(np.diff(p.view_begin)).size > 1  # ragged
    # make one view empty
    vb = p.view_begin.copy()
    n0 = vb[2] - vb[1]
    keep = np.ones(p.n_terms, bool)
    keep[vb[1]:vb[2]] = False
    vb[2:] -= n0
    from kalibr_b200.problem import Problem

    q = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, p.target_points, p.view_set, p.view_cam, vb,
                p.y_u[keep], p.y_v[keep], p.corner_id[keep])
    o = oracle_lib.OracleProblem(q)
    sol, _ = o.optimize(KbOptimizerOptions.kalibr2_default())
    assert sol.linear_solver_failure == 0


@pytest.mark.parametrize("name", ["cfg1_S3", "cfg2_S2", "cfg3_S2", "cfg4_S1", "cfg6_S2", "cfg7_S2"])
def test_oracle_reproduces_golden_fixtures(oracle_lib, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = synthetic.make_config(int(g["cfg"]), n_sets=int(g["n_sets"]))
    assert np.array_equal(p.y_u, g["y_u"]) and np.array_equal(p.init_cam if hasattr(p, "init_cam") else p.cam_params, g["init_cam"])
    o = oracle_lib.OracleProblem(p, n_threads=3)
    assert abs(o.evaluate_error() - float(g["J0"])) <= 1e-12 * float(g["J0"])
    assert np.abs(o.error_vector() - g["e"]).max() <= 1e-12 * np.abs(g["e"]).max()
    cp, ri, jv = o.jacobian_ccs()
    assert np.array_equal(cp, g["jt_col_ptr"]) and np.array_equal(ri, g["jt_row_idx"])
    assert np.abs(jv - g["jt_values"]).max() <= 1e-12 * np.abs(g["jt_values"]).max()
    o.build_system()
    o.set_constant_conditioner(10.0)
    dx, ok = o.solve_system()
    assert ok and np.abs(dx - g["dx"]).max() <= 1e-9 * np.abs(g["dx"]).max()
    hcp, hbr, hvp, hval = o.hessian_blocks()
    assert np.array_equal(hcp, g["h_col_ptr"]) and np.array_equal(hbr, g["h_block_row"]) and np.array_equal(hvp, g["h_value_ptr"])
    o2 = oracle_lib.OracleProblem(p, n_threads=3)
    sol, _ = o2.optimize(KbOptimizerOptions.kalibr2_default())
    assert sol.iterations == int(g["iterations"]) and sol.failed_iterations == int(g["failed_iterations"])
    assert np.abs(o2.camera_params() - g["cam_params"]).max() <= 1e-8 * np.abs(g["cam_params"]).max()


def _schur_marginal_numpy(o, cols):
    cp, ri, jv = o.jacobian_ccs()
    J = _dense_from_ccs(cp, ri, jv, o.jcols)
    H = J.T @ J
    pose = np.setdiff1d(np.arange(o.jcols), cols)
    return H[np.ix_(cols, cols)] - H[np.ix_(cols, pose)] @ np.linalg.pinv(H[np.ix_(pose, pose)]) @ H[np.ix_(pose, cols)]


@pytest.mark.parametrize("cfg,n_sets", [(1, 12), (2, 9), (3, 6), (6, 7)])
def test_marginal_analysis_matches_dense_schur_complement(oracle_lib, cfg, n_sets):
    """analyzeMarginal restated (QR of the pose columns, Omega, SVD) == the Schur complement of the dense normal equations;
    rank, tolerance and log2-sum follow aslam_incremental_calibration's linalg.cpp:244-282 / LinearSolver.cpp:196-200."""
    p = synthetic.make_config(cfg, n_sets=n_sets)
    o = oracle_lib.OracleProblem(p)
    res, sv, V, cols, om = o.analyze_marginal()
    S = _schur_marginal_numpy(o, cols)
    assert np.abs(S - om).max() <= 1e-11 * np.abs(S).max()
    svn = np.linalg.svd(S, compute_uv=False)
    assert np.abs(svn - sv).max() <= 1e-12 * svn[0]
    tol = svn[0] * np.finfo(float).eps * p.n_c
    rank = p.n_c
    for i in range(p.n_c - 1, 0, -1):  # estimateNumericalRank
        if svn[i] > tol:
            break
        rank -= 1
    assert res.n == p.n_c and res.rank == rank and res.rank_deficiency == p.n_c - rank
    assert np.isinf(res.sv_gap) if rank == p.n_c else abs(res.sv_gap - svn[rank - 1] / svn[rank]) <= 0.3 * res.sv_gap  # the tiny trailing value is only known absolutely
    svn = svn[:rank]
    assert abs(res.tolerance - sv[0] * np.finfo(float).eps * res.n) <= 1e-15 * res.tolerance
    assert abs(res.sv_log2_sum - np.log2(svn).sum()) <= 1e-9 * abs(res.sv_log2_sum)
    assert np.abs(V.T @ V - np.eye(res.n)).max() < 1e-12
    assert np.abs(V @ np.diag(sv) @ V.T - om).max() <= 1e-12 * sv[0]


def test_marginal_analysis_finds_the_rank_deficiency(oracle_lib):
    """A camera nobody observes leaves its intrinsics and its baseline unobservable: 8 + 6 zero singular values."""
    from kalibr_b200.problem import Problem

    p = synthetic.make_config(2, n_sets=8)
    keep_view = p.view_cam == 0
    keep_term = np.repeat(keep_view, np.diff(p.view_begin))
    vb = np.concatenate([[0], np.cumsum(np.diff(p.view_begin)[keep_view])]).astype(np.int64)
    q = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, p.target_points, p.view_set[keep_view],
                p.view_cam[keep_view], vb, p.y_u[keep_term], p.y_v[keep_term], p.corner_id[keep_term])
    res, sv, V, cols, om = oracle_lib.OracleProblem(q).analyze_marginal()
    assert res.rank == 8 and res.rank_deficiency == 14
    assert sv[8:].max() <= res.tolerance and sv[7] > 1e3 * res.tolerance
    assert res.sv_gap > 1e6
    assert abs(res.sv_log2_sum - np.log2(sv[:8]).sum()) < 1e-9
