"""The oracle's camera models and SE(3) helpers against the REFERENCE's own code.

tests/golden/reference_golden.npz holds what the reference's PinholeProjection / OmniProjection / ExtendedUnifiedProjection /
DoubleSphereProjection and *Distortion classes return - compiled from /root/reference by oracle/ref_pin.cpp (stand-in headers for Eigen /
Boost / OpenCV / sm_* in oracle/ref_shim/; generator: tests/golden/make_reference_golden.py).  The oracle must reproduce every
value: keypoint, point Jacobian, intrinsics Jacobian, distortion Jacobian, including the points where a model returns before writing its
outputs (quirk Q6), the NaNs of the equidistant model on the optical axis (Q5), the EUCM fu-for-fv entry (Q4) and the pinhole models'
indifference to the sign of the homogeneous scale in the Jacobian (Q3).  The same for sm_kinematics' quat2r, updateQuat, boxMinus and
boxTimes (compiled from the reference's quaternion_algebra.cpp / transformations.cpp), and for the residual and the complete Jacobian rows
of every reprojection term of two small problems as the reference's expression tree (RotationQuaternion, EuclideanPoint, TransformationBasic,
the Transformation / Homogeneous expression nodes, JacobianContainer: aslam_backend_expressions / aslam_backend compiled from their own
sources) and camera models produce them, combined as ReprojectionError does.  This pins rows a3-a17 of SURVEY.md §8.  Rows a1, a2, a19-a25 -
the LOOP - are pinned by the optimiser fixture: the reference's own Optimizer2, LevenbergMarquardtTrustRegionPolicy,
BlockCholeskyLinearSystemSolver, ErrorTermFs::buildHessian, JacobianContainer::evaluateHessian, OptimizationProblem and SparseBlockMatrix,
compiled from their sources (oracle/ref_pin_optimizer.cpp), run on eleven small problems in all four design-variable orders; only the
factorisation behind LinearSolverCholmod is a stand-in there (dense Cholesky; CHOLMOD is not in the image), i.e. rounding.  The GPU path is
compared with the oracle in the -m gpu suites and with this fixture in tests/test_reference_pin_gpu.py."""
import os

import numpy as np
import pytest

from oracle import oracle_api as oa

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz")
MODEL_NAMES = ["pinhole-radtan", "pinhole-equi", "omni-radtan", "eucm-none", "ds-none", "pinhole-fov", "omni-none"]
# the oracle is built with the same compiler as the reference pin and follows the reference's expression order: the values agree to the
# last bit in the build container; the bar leaves room for another compiler's choice of fused multiply-adds
RTOL = 1e-13


def same(a, b):
    """equal where both are finite (relative to max(|b|, 1)), NaN / Inf in the same places"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    if not np.array_equal(np.isfinite(a), fin) or not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    return bool(np.all(np.abs(a[fin] - b[fin]) <= RTOL * np.maximum(np.abs(b[fin]), 1.0)))


@pytest.mark.parametrize("model", range(7), ids=MODEL_NAMES)
def test_oracle_camera_models_reproduce_the_reference(oracle_lib, model):
    g = np.load(GOLD)
    P, H = g[f"m{model}_params"], g[f"m{model}_ph"]
    n_bailed = 0
    for i in range(len(P)):
        y, Jp, Ji, Jd, _ = oa.camera_project(model, P[i], H[i])
        assert same(y, g[f"m{model}_y"][i]), (model, i, H[i], y, g[f"m{model}_y"][i])
        assert same(Jp, g[f"m{model}_Jp"][i]), (model, i, H[i])
        assert same(Ji, g[f"m{model}_Ji"][i]), (model, i, H[i])
        assert same(Jd, g[f"m{model}_Jd"][i]), (model, i, H[i])
        n_bailed += int(not g[f"m{model}_y"][i].any())
    if model in (2, 3, 4, 6):  # the models with a validity cone: the fixture does contain points they refuse (outputs left untouched)
        assert n_bailed > 0


@pytest.mark.parametrize("name,args", [("quat2r", ("kin_q",)), ("update_quat", ("kin_q", "kin_dq")), ("box_minus", ("kin_p4",)), ("box_times", ("kin_T",))])
def test_oracle_se3_helpers_reproduce_the_reference(oracle_lib, name, args):
    g = np.load(GOLD)
    for i in range(len(g["kin_q"])):
        assert same(oa.kinematics(name, *[g[a][i] for a in args]), g["kin_" + name][i]), (name, i)


def term_problem(g, tag):
    from kalibr_b200.problem import Problem

    f = lambda n: g[f"term_{tag}_{n}"]  # noqa: E731
    return Problem(driver_order=int(f("order")), cam_model=f("cam_model"), cam_params=f("cam_params"), baselines=f("baselines"), set_poses=f("set_poses"),
                   target_points=f("target_points"), view_set=f("view_set"), view_cam=f("view_cam"), view_begin=f("view_begin"), y_u=f("y_u"), y_v=f("y_v"),
                   corner_id=f("corner_id"))


def dense_from_ccs(col_ptr, row_idx, values, jcols):
    J = np.zeros((len(col_ptr) - 1, jcols))
    for r in range(len(col_ptr) - 1):
        J[r, row_idx[col_ptr[r]:col_ptr[r + 1]]] = values[col_ptr[r]:col_ptr[r + 1]]
    return J


@pytest.mark.parametrize("tag", ["rig", "batch"])
def test_oracle_terms_reproduce_the_reference_expression_tree(oracle_lib, tag):
    """residual and Jacobian rows (pose chain through the expression nodes + camera part) of every term: the reference's own code vs the oracle"""
    g = np.load(GOLD)
    p = term_problem(g, tag)
    o = oa.OracleProblem(p)
    o.evaluate_error()
    res, rows = g[f"term_{tag}_residuals"], g[f"term_{tag}_jacobian"]
    assert np.abs(-o.error_vector() - res).max() <= 1e-12 * np.abs(p.y_u).max()  # e() = -(y - y_hat), pixels
    J = dense_from_ccs(*o.jacobian_ccs(), o.jcols)
    assert J.shape == rows.shape
    assert np.array_equal(J != 0, rows != 0)  # same sparsity
    assert (np.abs(J - rows) / np.maximum(np.abs(rows).max(axis=1, keepdims=True), 1.0)).max() <= 1e-13
    o.build_system()
    rhs_ref = -(rows.T @ res)  # rhs = -J^T e with the reference's J and e = y - y_hat
    assert np.abs(o.rhs() - rhs_ref).max() <= 1e-12 * np.abs(rhs_ref).max()


def test_oracle_m_estimator_weights_reproduce_the_reference(oracle_lib):
    """Huber / Cauchy / Geman-McClure / Blake-Zisserman weights of BE/src/MEstimatorPolicies.cpp (row a19); the chi-squared quantile behind
    Blake-Zisserman's epsilon is Boost.Math in the reference and a stand-in in the pin, so that one number is also checked against scipy"""
    g = np.load(GOLD)
    for kind, prm, sq, w in zip(g["mest_kind"], g["mest_params"], g["mest_s"], g["mest_w"]):
        assert abs(oa.m_estimator_weight(int(kind), float(sq), *[float(x) for x in prm]) - w) <= 1e-14 * max(abs(w), 1e-300), (kind, prm, sq)
    from scipy.stats import chi2

    eps = (1 - 0.1) / 0.1 * np.exp(-chi2.ppf(0.999, 2))
    w = np.exp(-3.0) / (np.exp(-3.0) + eps)
    assert abs(oa.m_estimator_weight(4, 3.0, 2.0, 0.999, 0.1) - w) <= 1e-12 * w


def opt_problem(g, n):
    from kalibr_b200.problem import KbOptimizerOptions, Problem

    f = lambda name: g[f"opt{n}_{name}"]  # noqa: E731
    p = Problem(driver_order=int(f("order")), cam_model=f("cam_model"), cam_params=f("cam_params"), baselines=f("baselines"), set_poses=f("set_poses"),
                target_points=f("target_points"), view_set=f("view_set"), view_cam=f("view_cam"), view_begin=f("view_begin"), y_u=f("y_u"), y_v=f("y_v"),
                corner_id=f("corner_id"))
    opt = KbOptimizerOptions.kalibr2_default(device_loop=0)
    opt.lm_lambda_init = float(f("lambda_init"))
    return p, opt


def check_against_reference_optimizer(g, n, solve, cost_rtol=1e-9):
    """solve(problem, options) -> (KbSolution, cam_params, baselines, set_poses): the whole run and the runs cut after 1, 2, ... iterations
    must walk the reference's iterations: same counts, same cost (1e-9), same final design variables (1e-6 of their scale; the
    factorisations differ in rounding and the last accepted steps are ~1e-3 of it)"""
    p, opt = opt_problem(g, n)
    it, failed, j_start, j_final, lsf = g[f"opt{n}_result"]
    sol, cp, bl, sp = solve(p, opt)
    assert (sol.iterations, sol.failed_iterations, sol.linear_solver_failure) == (int(it), int(failed), int(lsf))
    assert abs(sol.j_start - j_start) <= 1e-11 * j_start and abs(sol.j_final - j_final) <= cost_rtol * j_final
    for mine, ref in ((cp, g[f"opt{n}_final_cam_params"]), (np.reshape(bl, (-1, 7)), g[f"opt{n}_final_baselines"].reshape(-1, 7)), (sp, g[f"opt{n}_final_set_poses"])):
        if ref.size:
            assert np.abs(np.asarray(mine) - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1.0)
    for k, (itk, failedk, _, jk, lsfk) in enumerate(g[f"opt{n}_truncated"], start=1):
        opt.max_iterations = k
        s = solve(p, opt)[0]
        assert (s.iterations, s.failed_iterations, s.linear_solver_failure) == (int(itk), int(failedk), int(lsfk)), k
        assert abs(s.j_final - jk) <= cost_rtol * jk, k


N_OPT = 11


@pytest.mark.parametrize("n", range(N_OPT))
def test_oracle_optimizer_walks_the_reference_optimizer(oracle_lib, n):
    """rows a1, a2, a19-a25: Optimizer2::optimize with the LM policy and the BlockCholesky solver (Q1 / Q2 damping, rejected steps and
    reverts, the sticky linear-solver failure of problem 3) - the reference's compiled loop against the oracle's restatement"""
    g = np.load(GOLD)
    assert int(g["opt_count"]) == N_OPT

    def solve(p, opt):
        o = oa.OracleProblem(p)
        sol, _ = o.optimize(opt)
        return sol, o.camera_params(), o.baselines(), o.set_poses()

    check_against_reference_optimizer(g, n, solve)
    assert sum(int(g[f"opt{i}_result"][1]) > 0 for i in range(N_OPT)) >= 4 and sum(int(g[f"opt{i}_result"][4]) for i in range(N_OPT)) >= 1


def test_optimizer_fixture_is_what_the_reference_returns_now(oracle_lib):
    """build container only: the reference's compiled optimiser, run again, returns the committed numbers"""
    if oa.build_reference_cameras() is None:
        pytest.skip("no reference tree and no prebuilt oracle/_ref here")
    g = np.load(GOLD)
    for n in (0, 1, 2, 8, 9):
        p, opt = opt_problem(g, n)
        r, cp, bl, sp = oa.reference_optimize(p, opt)
        assert [r["iterations"], r["failed_iterations"], r["j_start"], r["j_final"], r["linear_solver_failure"]] == list(g[f"opt{n}_result"])
        assert np.array_equal(cp, g[f"opt{n}_final_cam_params"]) and np.array_equal(sp, g[f"opt{n}_final_set_poses"])


def test_fixture_is_what_the_reference_returns_now(oracle_lib):
    """In the build container (reference tree present): rebuild oracle/_ref from the reference sources and check the committed fixture
    and fresh random points against it; elsewhere only the fixture test above runs."""
    if oa.build_reference_cameras() is None:
        pytest.skip("no reference tree and no prebuilt oracle/_ref here")
    g = np.load(GOLD)
    rng = np.random.default_rng(5)
    for model in range(7):
        P, H = g[f"m{model}_params"], g[f"m{model}_ph"]
        for i in range(0, len(P), 7):
            y, Jp, Ji, Jd, ok = oa.reference_camera_project(model, P[i], H[i])
            assert np.array_equal(y, g[f"m{model}_y"][i], equal_nan=True) and np.array_equal(Jp, g[f"m{model}_Jp"][i], equal_nan=True)
            assert np.array_equal(Ji, g[f"m{model}_Ji"][i], equal_nan=True) and np.array_equal(Jd, g[f"m{model}_Jd"][i], equal_nan=True)
            assert ok == g[f"m{model}_ok"][i]
        for _ in range(200):
            prm = P[0] * (1.0 + 0.03 * rng.standard_normal(P[0].shape))
            ph = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-0.5, 3.0), rng.choice([1.0, -1.0, 2.0])])
            a, b = oa.camera_project(model, prm, ph), oa.reference_camera_project(model, prm, ph)
            for x, r in zip(a[:4], b[:4]):
                assert same(x, r), (model, ph)
    for i in range(0, len(g["kin_q"]), 5):
        assert np.array_equal(oa.reference_kinematics("update_quat", g["kin_q"][i], g["kin_dq"][i]), g["kin_update_quat"][i])
        assert np.array_equal(oa.reference_kinematics("box_times", g["kin_T"][i]), g["kin_box_times"][i])
