"""CPU tests of the oracle of the incremental estimator's numerical core (oracle/ko_estimator.py) through properties - the part of it that
has no reference pin (the QR elimination and the truncated solve as a whole; its Gauss-Newton loop, design-variable order and rank / scaling
numerics are pinned against the reference's compiled code in tests/test_reference_gauss_newton_pin_cpu.py and test_reference_linalg_pin_cpu.py).
  IC/src/core/LinearSolver.cpp:299-463, IC/src/algorithms/linalg.cpp:128-152, 244-282, 426-443
  IC/test/algorithms/*: the reference tests its marginalisation against a dense solve in the same way
"""
import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import Problem
from oracle import ko_estimator as ke


def without_camera(p, cam):
    keep_view = p.view_cam != cam
    keep_term = np.repeat(keep_view, np.diff(p.view_begin))
    vb = np.concatenate([[0], np.cumsum(np.diff(p.view_begin)[keep_view])]).astype(np.int64)
    return Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, p.target_points, p.view_set[keep_view], p.view_cam[keep_view], vb,
                   p.y_u[keep_term], p.y_v[keep_term], p.corner_id[keep_term])


@pytest.mark.parametrize("cfg,n_sets", [(1, 6), (2, 5), (7, 6)])
@pytest.mark.parametrize("scaling", [False, True])
def test_truncated_solve_equals_least_squares_at_full_rank(oracle_lib, cfg, n_sets, scaling):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    o = oracle_lib.OracleProblem(p)
    J, b = ke.system_of(o, p)
    cal, rest = ke.calibration_columns(p)
    assert len(cal) == p.n_c and len(cal) + len(rest) == o.jcols
    x, info = ke.linear_solver_solve(J, b, cal, rest, column_scaling_on=scaling, eps_svd=1e-6 if scaling else ke.EPS)
    xl = np.linalg.lstsq(J, b, rcond=None)[0]
    assert info["rank"] == p.n_c and info["sv_gap"] == np.inf
    assert np.abs(x - xl).max() <= 1e-8 * np.abs(xl).max()
    # and it is the undamped normal-equation step of the block solver
    o.build_system(); o.set_constant_conditioner(0.0)
    dx, ok = o.solve_system()
    assert ok and np.abs(dx - x).max() <= 1e-8 * np.abs(x).max()


def test_truncated_solve_leaves_unobservable_directions_alone(oracle_lib):
    """A camera nobody observes: its intrinsics and baseline get a zero update (minimum-norm solution), the rest the least-squares one."""
    p = without_camera(synthetic.make_config(2, n_sets=6), 1)
    o = oracle_lib.OracleProblem(p)
    J, b = ke.system_of(o, p)
    cal, rest = ke.calibration_columns(p)
    x, info = ke.linear_solver_solve(J, b, cal, rest, column_scaling_on=True, eps_svd=1e-6)
    assert info["rank"] == 8 and info["sv_gap"] > 1e6
    col, dims, labels = p.dv_layout()
    dead = [c + i for c, d, l in zip(col, dims, labels) if l in (("proj", 1), ("dist", 1), ("baseline_q", 0), ("baseline_t", 0)) for i in range(d)]
    assert np.all(x[dead] == 0.0)
    live = np.setdiff1d(np.arange(o.jcols), dead)
    xl = np.linalg.lstsq(J[:, live], b, rcond=None)[0]
    assert np.abs(x[live] - xl).max() <= 1e-8 * np.abs(xl).max()


def test_column_scaling_matrix():
    A = np.array([[3.0, 0.0, 1e-12], [4.0, 0.0, 0.0]])
    g = ke.column_scaling(A, ke.EPS)
    assert g[0] == 0.2 and g[1] == 0.0 and g[2] == 0.0  # columns below sqrt(rows * eps) are dropped


@pytest.mark.parametrize("cfg,n_sets", [(1, 10), (2, 8)])
def test_gauss_newton_converges_like_lm(oracle_lib, cfg, n_sets):
    from kalibr_b200.problem import KbOptimizerOptions

    p = synthetic.make_config(cfg, n_sets=n_sets)
    r = ke.gauss_newton_optimize(oracle_lib.OracleProblem(p), p, dict(column_scaling_on=True, eps_svd=1e-6))
    lm, _ = oracle_lib.OracleProblem(p).optimize(KbOptimizerOptions.kalibr2_default())
    assert r["iterations"] <= 20 and abs(r["j_final"] - lm.j_final) <= 1e-3 * lm.j_final
    m = ke.analyze_marginal(oracle_lib.OracleProblem(p), p)
    assert m["rank"] == p.n_c


def test_mixed_rig_with_few_sets_is_numerically_rank_deficient(oracle_lib):
    """What the truncation is for: four sets do not pin down every parameter of the mixed rig (EUCM's alpha / beta), the solver
    with kalibr2_ros' settings (column scaling, epsSVD = 1e-6) drops those directions instead of taking a wild step."""
    p = synthetic.make_config(3, n_sets=4)
    o = oracle_lib.OracleProblem(p)
    J, b = ke.system_of(o, p)
    cal, rest = ke.calibration_columns(p)
    x, info = ke.linear_solver_solve(J, b, cal, rest, column_scaling_on=True, eps_svd=1e-6)
    xl = np.linalg.lstsq(J, b, rcond=None)[0]
    assert info["rank"] < p.n_c
    assert np.abs(x).max() < 0.2 * np.abs(xl).max()
