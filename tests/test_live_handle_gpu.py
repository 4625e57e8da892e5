"""A live handle grows and shrinks by synced sets (kb_append_set / kb_remove_last_set ≙ IncrementalOptimizationProblem::add / remove of a
batch, IC/src/core/IncrementalOptimizationProblem.cpp:186-260) and snapshots its design variables (kb_save / kb_restore_design_variables ≙
OptimizationProblem::saveDesignVariables / restoreDesignVariables, IC/src/core/OptimizationProblem.cpp:260-272): a handle that was
grown set by set must be indistinguishable from one created with the whole problem, and the oracle agrees with both."""
import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions, Problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()
    return m


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max(initial=0.0) / max(np.abs(b).max(initial=0.0), 1e-300)


def first_sets(p, n):
    """the sub-problem of the first n synced sets (term order per set, as the rig / batch drivers produce it)"""
    keep = p.view_set < n
    assert np.all(np.diff(p.view_set) >= 0), "views are listed set by set"
    nv = int(keep.sum())
    nt = int(p.view_begin[nv])
    return Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses[:n], p.target_points, p.view_set[:nv], p.view_cam[:nv],
                   p.view_begin[:nv + 1], p.y_u[:nt], p.y_v[:nt], p.corner_id[:nt])


def append(g, p, s):
    w = np.flatnonzero(p.view_set == s)
    b, e = int(p.view_begin[w[0]]), int(p.view_begin[w[-1] + 1])
    g.append_set(p.view_cam[w], p.view_begin[w[0]:w[-1] + 2] - b, p.y_u[b:e], p.y_v[b:e], p.corner_id[b:e], p.set_poses[s])


def one_iteration(g):
    J = g.evaluate_error()
    e = g.error_vector()
    g.build_system()
    rhs = g.rhs()
    g.set_constant_conditioner(10.0)
    dx, ok = g.solve_system()
    return J, e, rhs, dx, ok


@pytest.mark.parametrize("cfg", [8, 3, 2])  # batch order (poses first: the camera columns move when a set arrives), rig order, stereo order
def test_grown_handle_equals_created_handle(capi, oracle_lib, cfg):
    full = synthetic.make_config(cfg, n_sets=9)
    if cfg == 2:  # the stereo driver lists all camera-0 views first: re-list the views set by set (the layout does not depend on the term order)
        order = np.lexsort((full.view_cam, full.view_set))
        vb, yu, yv, cid = [0], [], [], []
        for w in order:
            b, e = int(full.view_begin[w]), int(full.view_begin[w + 1])
            yu += list(full.y_u[b:e]); yv += list(full.y_v[b:e]); cid += list(full.corner_id[b:e]); vb.append(len(yu))
        full = Problem(full.driver_order, full.cam_model, full.cam_params, full.baselines, full.set_poses, full.target_points, full.view_set[order],
                       full.view_cam[order], np.array(vb, np.int64), np.array(yu), np.array(yv), np.array(cid, np.int32))
    part = first_sets(full, 5)
    g = capi.B200SchurLinearSystemSolver(part)
    ref_part = one_iteration(capi.B200SchurLinearSystemSolver(part))
    for s in range(5, 9):
        append(g, full, s)
    gf = capi.B200SchurLinearSystemSolver(full)
    assert g.jcols == gf.jcols and g.jrows == gf.jrows and g.n_dv == gf.n_dv
    assert all(np.array_equal(a, b) for a, b in zip(g.dv_layout(), gf.dv_layout()))
    a, b = one_iteration(g), one_iteration(gf)
    assert abs(a[0] - b[0]) <= 1e-13 * b[0] and np.array_equal(a[1], b[1])
    assert rel(a[2], b[2]) < 1e-12 and a[4] and b[4] and rel(a[3], b[3]) < 1e-9
    gp, gq = g.hessian_blocks(), gf.hessian_blocks()
    assert all(np.array_equal(x, y) for x, y in zip(gp[:3], gq[:3])) and rel(gp[3], gq[3]) < 1e-12
    # ... and equals the oracle on the whole problem, through a full optimisation
    g.reset_state()
    gs, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
    o = oracle_lib.OracleProblem(full, solver_kind=oracle_lib.SPARSE_CHOLESKY if cfg in (3, 8) else oracle_lib.BLOCK_CHOLESKY)
    if cfg in (3, 8):
        g2 = capi.B200SchurLinearSystemSolver(full)
        g2.set_solver_semantic(1)
        gs, _ = g2.optimize(KbOptimizerOptions.kalibr2_default())
        g.set_solver_semantic(1)
        g.reset_state()
        gs2, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
        assert (gs2.iterations, gs2.failed_iterations) == (gs.iterations, gs.failed_iterations) and abs(gs2.j_final - gs.j_final) <= 1e-12 * gs.j_final
    os_, _ = o.optimize(KbOptimizerOptions.kalibr2_default())
    assert gs.iterations == os_.iterations and gs.failed_iterations == os_.failed_iterations
    assert abs(gs.j_final - os_.j_final) <= 1e-9 * os_.j_final
    assert np.abs(g.set_poses() - o.set_poses()).max() < 1e-6
    # shrink again: back to the five-set problem, bit for bit in the residuals
    g.reset_state()
    for _ in range(4):
        g.remove_last_set()
    assert g.jcols == part.n_c + 6 * 5
    c = one_iteration(g)
    assert abs(c[0] - ref_part[0]) <= 1e-13 * ref_part[0] and np.array_equal(c[1], ref_part[1])
    assert rel(c[2], ref_part[2]) < 1e-12 and rel(c[3], ref_part[3]) < 1e-9


def test_handle_can_shrink_to_nothing_and_grow_again(capi):
    p = synthetic.make_config(8, n_sets=3)
    g = capi.B200SchurLinearSystemSolver(first_sets(p, 1))
    g.remove_last_set()
    assert g.jcols == p.n_c and g.jrows == 0
    assert g.evaluate_error() == 0.0
    for s in range(3):
        append(g, p, s)
    a, b = one_iteration(g), one_iteration(capi.B200SchurLinearSystemSolver(p))
    assert abs(a[0] - b[0]) <= 1e-13 * b[0] and rel(a[3], b[3]) < 1e-9


def test_save_and_restore_design_variables(capi):
    p = synthetic.make_config(8, n_sets=6)
    g = capi.B200SchurLinearSystemSolver(first_sets(p, 5))
    g.save_design_variables()
    cam0, base0, sets0 = g.camera_params(), g.baselines(), g.set_poses()
    append(g, p, 5)
    g.optimize(KbOptimizerOptions.kalibr2_default())
    assert np.abs(g.camera_params() - cam0).max() > 0
    # a rejected batch: the set goes, every design variable returns (IncrementalEstimator.cpp:520-530)
    g.remove_last_set()
    g.restore_design_variables()
    assert np.array_equal(g.camera_params(), cam0) and np.array_equal(g.baselines(), base0) and np.array_equal(g.set_poses(), sets0)
    with pytest.raises(capi.KalibrB200Error):
        capi.B200SchurLinearSystemSolver(first_sets(p, 2)).restore_design_variables()
