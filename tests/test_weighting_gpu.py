"""GPU parity of the term weighting (kb_set_inv_r, kb_set_m_estimator) and of kb_reprojection_statistics against the
CPU oracle, through the C ABI.  Same bars as tests/test_parity_gpu.py: residuals / Jacobians / normal equations within
1e-9 relative, converged parameters within 1e-6 relative with the same iteration counts.
"""
import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions

pytestmark = pytest.mark.gpu

REL_J = 1e-9
REL_X = 1e-6
NONE, HUBER, CAUCHY, GEMAN, BLAKE = range(5)
INV_R = np.array([[3.0, 0.4], [0.4, 5.0]])        # general SPD, pivoted square root
INV_R_ISO = np.eye(2) / (0.3 * 0.3)               # I / sigma^2 as CreateBatchProblem passes (K2/CalibrationTools.hpp:495-496)
# (kind, p0, p1, p2): parameters in the units of the weighted squared error
POLICIES = [(HUBER, 1.5, 0.0, 0.0), (CAUCHY, 4.0, 0.0, 0.0), (GEMAN, 9.0, 0.0, 0.0), (BLAKE, 2, 0.999, 0.1)]
CASES = [(1, 12), (2, 8), (3, 6), (4, 3), (6, 5)]


def rel_err(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.abs(a - b).max(initial=0.0) / max(np.abs(b).max(initial=0.0), 1e-300)


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()
    return m


def pair(capi, oracle_lib, p, inv_r=None, policy=None):
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    if inv_r is not None:
        g.set_inv_r(inv_r)
        o.set_inv_r(inv_r)
    if policy is not None:
        pg = g.set_m_estimator(*policy)
        po = o.set_m_estimator(*policy)
        assert abs(pg - po) <= 1e-12 * abs(po)  # Blake-Zisserman: epsilon from the library's own chi-squared quantile
    return g, o


def test_sqrt_inv_r_matches_oracle(capi, oracle_lib):
    p = synthetic.make_config(1, n_sets=2)
    g = capi.B200SchurLinearSystemSolver(p)
    assert np.array_equal(g.sqrt_inv_r(), np.eye(2))
    for A in (INV_R, INV_R[::-1, ::-1].copy(), INV_R_ISO, np.array([[2.0, -0.7], [-0.7, 2.0]])):
        g.set_inv_r(A)
        assert rel_err(g.sqrt_inv_r(), oracle_lib.matrix_sqrt2(A)) < 1e-15
    with pytest.raises(capi.KalibrB200Error):
        g.set_inv_r(np.array([[1.0, 2.0], [2.0, 1.0]]))  # not positive definite
    with pytest.raises(capi.KalibrB200Error):
        g.set_m_estimator(HUBER, -1.0)


@pytest.mark.parametrize("policy", [None] + POLICIES)
@pytest.mark.parametrize("cfg,n_sets", CASES)
def test_weighted_evaluate_build_solve_match_oracle(capi, oracle_lib, cfg, n_sets, policy):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    g, o = pair(capi, oracle_lib, p, INV_R, policy)
    Jg, Jo = g.evaluate_error(), o.evaluate_error()
    assert abs(Jg - Jo) <= 1e-11 * abs(Jo)
    assert rel_err(g.error_vector(), o.error_vector()) < REL_J
    g.build_system(); o.build_system()
    assert rel_err(g.rhs(), o.rhs()) < REL_J
    g.set_constant_conditioner(10.0); o.set_constant_conditioner(10.0)
    gdx, gok = g.solve_system()
    odx, ook = o.solve_system()
    assert gok and ook
    gcp, gbr, gvp, gval = g.hessian_blocks()
    ocp, obr, ovp, oval = o.hessian_blocks()
    assert np.array_equal(gcp, ocp) and np.array_equal(gbr, obr) and np.array_equal(gvp, ovp)
    assert rel_err(gval, oval) < REL_J
    assert rel_err(gdx, odx) < 1e-7


@pytest.mark.parametrize("policy", [None, POLICIES[0], POLICIES[3]])
@pytest.mark.parametrize("cfg,n_sets", [(2, 6), (3, 4)])
def test_weighted_jacobian_export_matches_oracle(capi, oracle_lib, cfg, n_sets, policy):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    g, o = pair(capi, oracle_lib, p, INV_R, policy)
    g.evaluate_error(); o.evaluate_error()
    gp, gi, gv = g.jacobian_ccs()
    op, oi, ov = o.jacobian_ccs()
    assert np.array_equal(gp, op) and np.array_equal(gi, oi)
    row_of = np.repeat(np.arange(gp.size - 1), np.diff(gp))
    scale = np.maximum(np.maximum.reduceat(np.abs(ov), op[:-1])[row_of], 1e-300)  # a redescending policy can zero a row
    assert (np.abs(gv - ov) / scale).max() < REL_J
    assert rel_err(g.error_vector(), o.error_vector()) < REL_J


@pytest.mark.parametrize("speculative", [True, False])
def test_use_m_estimator_flag_only_gates_rows(capi, oracle_lib, speculative):
    """useMEstimator = false: the cost keeps the policy weight, e(), H and rhs lose it (ErrorTerm.hpp:97-109, 183-192)."""
    p = synthetic.make_config(2, n_sets=7)
    g, o = pair(capi, oracle_lib, p, INV_R_ISO, POLICIES[1])
    g.set_speculative_linearise(speculative)
    J_on = g.evaluate_error(True)
    e_on = g.error_vector()
    o.set_use_m_estimator(False)
    Jg, Jo = g.evaluate_error(False), o.evaluate_error()
    assert abs(Jg - Jo) <= 1e-11 * abs(Jo) and abs(Jg - J_on) <= 1e-12 * J_on
    assert rel_err(g.error_vector(), o.error_vector()) < REL_J
    assert rel_err(g.error_vector(), e_on) > 1e-3  # the two modes really differ
    g.build_system(False); o.build_system()
    assert rel_err(g.rhs(), o.rhs()) < REL_J
    # and back on: the cached linearisation of the other mode must not be reused
    o.set_use_m_estimator(True)
    g.evaluate_error(True); o.evaluate_error()
    g.build_system(True); o.build_system()
    assert rel_err(g.rhs(), o.rhs()) < REL_J
    g.build_system(False)
    o.set_use_m_estimator(False); o.build_system()
    assert rel_err(g.rhs(), o.rhs()) < REL_J


@pytest.mark.parametrize("device_loop", [1, 0])
@pytest.mark.parametrize("inv_r,policy", [(INV_R_ISO, None), (INV_R, POLICIES[0]), (None, POLICIES[1]), (INV_R_ISO, POLICIES[3])])
@pytest.mark.parametrize("cfg,n_sets", [(1, 30), (2, 16), (3, 10)])
def test_weighted_optimize_matches_oracle(capi, oracle_lib, cfg, n_sets, inv_r, policy, device_loop):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    rng = np.random.default_rng(11 + cfg)
    bad = rng.choice(p.n_terms, p.n_terms // 40, replace=False)  # a few gross outliers for the policies to act on
    p.y_u[bad] += rng.normal(0, 15.0, bad.size)
    p.y_v[bad] += rng.normal(0, 15.0, bad.size)
    g, o = pair(capi, oracle_lib, p, inv_r, None)
    opt = KbOptimizerOptions.kalibr2_default()
    opt.device_loop = device_loop
    if policy is not None:
        if policy[0] == BLAKE:  # a redescending policy is installed on a converged state (it cuts everything far from the model)
            g.optimize(opt)
            o.optimize(KbOptimizerOptions.kalibr2_default())
        assert abs(g.set_m_estimator(*policy) - o.set_m_estimator(*policy)) <= 1e-12
    gs, gt = g.optimize(opt)
    os_, ot = o.optimize(KbOptimizerOptions.kalibr2_default())
    assert gs.iterations == os_.iterations and gs.failed_iterations == os_.failed_iterations
    assert abs(gs.j_final - os_.j_final) <= 1e-9 * os_.j_final
    assert rel_err(gt[:, 0], ot[:, 0]) < 1e-9  # cost trace
    oc = o.camera_params()
    assert (np.abs(g.camera_params() - oc) / np.maximum(np.abs(oc), 1e-3)).max() < REL_X
    if p.n_cams > 1:
        assert rel_err(g.baselines(), o.baselines()) < REL_X
    assert rel_err(g.set_poses(), o.set_poses()) < REL_X
    # a second optimisation on the same handle after switching the weighting off: the captured graph must not be reused
    g.set_m_estimator(NONE); o.set_m_estimator(NONE)
    g.set_inv_r(np.eye(2)); o.set_inv_r(np.eye(2))
    g.reset_state()
    o2 = oracle_lib.OracleProblem(p)
    gs2, _ = g.optimize(opt)
    os2, _ = o2.optimize(KbOptimizerOptions.kalibr2_default())
    assert gs2.iterations == os2.iterations and abs(gs2.j_final - os2.j_final) <= 1e-9 * os2.j_final


def test_unweighted_results_do_not_change_with_identity_weighting(capi):
    """invR = I and NoMEstimator run the unweighted kernels: bit-identical to a handle that never set anything."""
    p = synthetic.make_config(3, n_sets=5)
    a = capi.B200SchurLinearSystemSolver(p)
    b = capi.B200SchurLinearSystemSolver(p)
    b.set_inv_r(np.eye(2))
    b.set_m_estimator(NONE)
    assert a.evaluate_error() == b.evaluate_error()
    assert np.array_equal(a.error_vector(), b.error_vector())
    a.build_system(); b.build_system()
    assert np.array_equal(a.rhs(), b.rhs())


@pytest.mark.parametrize("cfg,n_sets", [(1, 9), (2, 1), (3, 5), (4, 3), (7, 6)])
def test_reprojection_statistics_match_oracle(capi, oracle_lib, cfg, n_sets):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    g, o = pair(capi, oracle_lib, p, INV_R, POLICIES[0])  # statistics use raw errors whatever the weighting
    g.evaluate_error(); o.evaluate_error()
    e_before = g.error_vector()
    sg, so = g.reprojection_statistics(), o.reprojection_statistics()
    assert np.array_equal(sg[:, 0], so[:, 0])
    assert np.abs(sg[:, 1:] - so[:, 1:]).max() <= 1e-11 * max(np.abs(so[:, 1:]).max(), 1.0)
    assert np.array_equal(g.error_vector(), e_before)  # e() is left untouched
    # after an optimisation the statistics describe the converged state
    g.optimize(); o.optimize()
    sg, so = g.reprojection_statistics(), o.reprojection_statistics()
    assert np.abs(sg[:, 3:5] - so[:, 3:5]).max() <= 1e-7 * so[:, 3:5].max()


def test_reprojection_statistics_with_a_camera_without_views(capi, oracle_lib):
    """n = 0 for a camera nobody observes: every statistic of it is zero, the others are unaffected."""
    from kalibr_b200.problem import Problem

    p = synthetic.make_config(2, n_sets=5)
    keep_view = p.view_cam == 0
    keep_term = np.repeat(keep_view, np.diff(p.view_begin))
    vb = np.concatenate([[0], np.cumsum(np.diff(p.view_begin)[keep_view])]).astype(np.int64)
    q = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, p.target_points, p.view_set[keep_view],
                p.view_cam[keep_view], vb, p.y_u[keep_term], p.y_v[keep_term], p.corner_id[keep_term])
    g, o = pair(capi, oracle_lib, q)
    sg, so = g.reprojection_statistics(), o.reprojection_statistics()
    assert np.all(sg[1] == 0.0) and np.all(so[1] == 0.0)
    assert np.abs(sg[0] - so[0]).max() <= 1e-11 * max(np.abs(so[0]).max(), 1.0)
