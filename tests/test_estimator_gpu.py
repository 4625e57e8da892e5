"""GPU parity of the incremental estimator's numerical core through the C ABI: kb_solve_system_svd (≙ aslam::calibration::
LinearSolver::solve) and kb_optimize_gauss_newton (≙ Optimizer2 with GaussNewtonTrustRegionPolicy over it) against the dense
numpy oracle (oracle/ko_estimator.py).  Bars: dx within 1e-7 relative (as for the damped solve), same iteration count, cost within
1e-9, parameters within 1e-6.
"""
import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions, KbSvdSolverOptions
from oracle import ko_estimator as ke

from test_estimator_cpu import without_camera

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()
    return m


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def clear_rank(info):
    """no singular value within 1 % of the tolerance: the rank decision does not hinge on rounding (device and oracle agree on the
    singular values to ~1e-8 of the largest one, i.e. ~1e-4 of one that sits at the tolerance)"""
    r = info["singular_values"] / info["tolerance"]
    return not np.any((r > 0.99) & (r < 1.01))


@pytest.mark.parametrize("cfg,n_sets,scaling", [(1, 12, False), (2, 9, False), (7, 7, False), (1, 12, True), (2, 9, True), (3, 6, True), (3, 30, True),
                                                (4, 5, True), (6, 5, True), (7, 7, True)])
def test_svd_solve_matches_oracle(capi, oracle_lib, cfg, n_sets, scaling):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    g.evaluate_error()
    g.build_system()
    opt = KbSvdSolverOptions.kalibr2() if scaling else KbSvdSolverOptions.default()
    dx, res, sv = g.solve_system_svd(opt)
    J, b = ke.system_of(o, p)
    cal, rest = ke.calibration_columns(p)
    x, info = ke.linear_solver_solve(J, b, cal, rest, column_scaling_on=scaling, eps_svd=opt.eps_svd)
    assert rel(sv, info["singular_values"]) < 1e-8  # relative to the largest one
    assert abs(res.tolerance - info["tolerance"]) <= 1e-8 * info["tolerance"]
    # the cases are chosen (on the CPU, with the oracle) so that no singular value sits within 1 % of the rank tolerance: no skip
    assert clear_rank(info), "this case sits on the rank tolerance: pick another size / seed instead of skipping"
    assert res.n == p.n_c and res.rank == info["rank"] and res.rank_deficiency == p.n_c - info["rank"]
    if info["rank"] < p.n_c:
        assert abs(res.sv_gap - info["sv_gap"]) <= 1e-4 * info["sv_gap"]
    else:
        assert res.sv_gap == np.inf
    assert rel(dx, x) < 1e-7
    # the step can be applied like any other
    m = g.apply_state_update()
    assert abs(m - np.abs(x).max()) <= 1e-7 * np.abs(x).max()


def test_svd_solve_rank_deficient(capi, oracle_lib):
    p = without_camera(synthetic.make_config(2, n_sets=8), 1)
    g = capi.B200SchurLinearSystemSolver(p)
    g.evaluate_error(); g.build_system()
    dx, res, sv = g.solve_system_svd(KbSvdSolverOptions.kalibr2())
    o = oracle_lib.OracleProblem(p)
    J, b = ke.system_of(o, p)
    cal, rest = ke.calibration_columns(p)
    x, info = ke.linear_solver_solve(J, b, cal, rest, column_scaling_on=True, eps_svd=1e-6)
    assert res.rank == info["rank"] == 8 and res.rank_deficiency == 14 and res.sv_gap > 1e6
    assert rel(dx, x) < 1e-7
    col, dims, labels = p.dv_layout()
    dead = [c + i for c, d, l in zip(col, dims, labels) if l in (("proj", 1), ("dist", 1), ("baseline_q", 0), ("baseline_t", 0)) for i in range(d)]
    assert np.abs(dx[dead]).max() <= 1e-12 * np.abs(dx).max()   # unobservable directions are left alone


@pytest.mark.parametrize("cfg,n_sets", [(1, 20), (2, 12), (7, 10)])
def test_gauss_newton_optimize_matches_oracle(capi, oracle_lib, cfg, n_sets):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    sol, tr = g.optimize_gauss_newton(KbOptimizerOptions.estimator_default(), KbSvdSolverOptions.kalibr2())
    o = oracle_lib.OracleProblem(p)
    r = ke.gauss_newton_optimize(o, p, dict(column_scaling_on=True, eps_svd=1e-6))
    assert sol.iterations == r["iterations"] and sol.failed_iterations == 0 and not sol.linear_solver_failure
    assert abs(sol.j_start - r["j_start"]) <= 1e-11 * r["j_start"] and abs(sol.j_final - r["j_final"]) <= 1e-9 * r["j_final"]
    assert rel(tr[:, 0], [t[0] for t in r["trace"]]) < 1e-9
    oc = o.camera_params()
    assert (np.abs(g.camera_params() - oc) / np.maximum(np.abs(oc), 1e-3)).max() < 1e-6
    assert rel(g.set_poses(), o.set_poses()) < 1e-6
    # the marginal analysis that follows in addBatch (unscaled system)
    mres, msv, _, _ = g.analyze_marginal()
    m = ke.analyze_marginal(o, p)
    assert mres.rank == m["rank"] and abs(mres.sv_log2_sum - m["sv_log2_sum"]) < 1e-6
