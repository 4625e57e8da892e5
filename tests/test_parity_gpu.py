"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): block pattern bit-exact; residuals and Jacobians within 1e-9 relative;
converged intrinsics/extrinsics within 1e-6 relative with the same iteration count.
"""
import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions

pytestmark = pytest.mark.gpu

REL_J = 1e-9   # residuals / Jacobians
REL_X = 1e-6   # converged parameters

# (config, n_sets): scaled-down versions of the five BASELINE configs that the oracle finishes in seconds
CASES = [(1, 40), (2, 30), (3, 24), (4, 12), (5, 6), (6, 20), (7, 25), (8, 18)]  # 6, 7: pinhole-fov and omni-none (rig / stereo order); 8: batch order


def rel_err(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    scale = max(np.abs(b).max(initial=0.0), 1e-300)
    return np.abs(a - b).max(initial=0.0) / scale


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()  # fails loudly if the extension is missing
    return m


def make(cfg, n_sets, **kw):
    return synthetic.make_config(cfg, n_sets=n_sets, **kw)


@pytest.mark.parametrize("cfg,n_sets", CASES)
def test_dv_layout_matches_oracle(capi, oracle_lib, cfg, n_sets):
    p = make(cfg, n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    gc, gd = g.dv_layout()
    oc, od = o.dv_layout()
    assert np.array_equal(gc, oc) and np.array_equal(gd, od)
    pc, pd, _ = p.dv_layout()
    assert np.array_equal(gc, pc) and np.array_equal(gd, pd)
    assert g.jrows == o.jrows and g.jcols == o.jcols


@pytest.mark.parametrize("cfg,n_sets", CASES)
def test_evaluate_error_matches_oracle(capi, oracle_lib, cfg, n_sets):
    p = make(cfg, n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    Jg, Jo = g.evaluate_error(), o.evaluate_error()
    assert abs(Jg - Jo) <= 1e-11 * abs(Jo)
    assert rel_err(g.error_vector(), o.error_vector()) < REL_J
    assert g.num_invalid_terms() == 0


@pytest.mark.parametrize("cfg,n_sets", CASES)
def test_jacobians_match_oracle(capi, oracle_lib, cfg, n_sets):
    p = make(cfg, n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    g.evaluate_error()
    o.evaluate_error()
    gp, gi, gv = g.jacobian_ccs()
    op, oi, ov = o.jacobian_ccs()
    assert np.array_equal(gp, op), "CCS column pointers differ"
    assert np.array_equal(gi, oi), "CCS row indices differ"
    # relative to the magnitude of each term's row
    n_rows = gp.size - 1
    row_of = np.repeat(np.arange(n_rows), np.diff(gp))
    scale = np.maximum.reduceat(np.abs(ov), op[:-1])[row_of]
    assert (np.abs(gv - ov) / scale).max() < REL_J
    # the materialising kernel also writes e
    assert rel_err(g.error_vector(), o.error_vector()) < REL_J


@pytest.mark.parametrize("cfg,n_sets", CASES)
def test_normal_equations_match_oracle(capi, oracle_lib, cfg, n_sets):
    p = make(cfg, n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    g.evaluate_error(); o.evaluate_error()
    g.build_system(); o.build_system()
    assert rel_err(g.rhs(), o.rhs()) < REL_J
    g.set_constant_conditioner(10.0); o.set_constant_conditioner(10.0)
    gdx, gok = g.solve_system()
    odx, ook = o.solve_system()
    assert gok and ook
    # block pattern: bit exact (after build + solve, as SparseBlockMatrix holds it)
    gcp, gbr, gvp, gval = g.hessian_blocks()
    ocp, obr, ovp, oval = o.hessian_blocks()
    assert np.array_equal(gcp, ocp), "block column pointers differ"
    assert np.array_equal(gbr, obr), "block row indices differ"
    assert np.array_equal(gvp, ovp), "block value offsets differ"
    assert rel_err(gval, oval) < REL_J
    assert rel_err(gdx, odx) < 1e-7


@pytest.mark.parametrize("cfg,n_sets", CASES)
def test_lm_step_sequence_matches_oracle(capi, oracle_lib, cfg, n_sets):
    """Two LM steps driven by hand, the second without a rebuild: exercises the lambda^2 / lambda residual (Q2),
    rho denominator, update and revert."""
    p = make(cfg, n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    g.evaluate_error(); o.evaluate_error()
    g.build_system(); o.build_system()
    for lam in (10.0, 20.0):
        g.set_constant_conditioner(lam); o.set_constant_conditioner(lam)
        gdx, gok = g.solve_system()
        odx, ook = o.solve_system()
        assert gok == ook
        assert rel_err(gdx, odx) < 1e-7
        rho_o = float(odx @ (lam * odx + o.rhs()))
        assert abs(g.lm_rho_denominator(lam) - rho_o) <= 1e-7 * abs(rho_o)
    mg = g.apply_state_update()
    mo = o.apply_state_update()
    assert abs(mg - mo) <= 1e-7 * mo
    Jg, Jo = g.evaluate_error(), o.evaluate_error()
    assert abs(Jg - Jo) <= 1e-7 * Jo
    assert rel_err(g.camera_params(), o.camera_params()) < 1e-9
    assert rel_err(g.set_poses(), o.set_poses()) < 1e-9
    if p.n_cams > 1:
        assert rel_err(g.baselines(), o.baselines()) < 1e-9
    g.revert_last_state_update(); o.revert_last_state_update()
    assert rel_err(g.camera_params(), o.camera_params()) == 0.0
    Jg2, Jo2 = g.evaluate_error(), o.evaluate_error()
    assert abs(Jg2 - Jo2) <= 1e-11 * Jo2


@pytest.mark.parametrize("cfg,n_sets,lams,pos_def", [(1, 40, (1e-3, 2e-3), True), (2, 30, (0.05, 0.02), True), (4, 12, (0.3, 0.4), True), (1, 40, (0.3, 0.4), True),
                                                     (3, 24, (0.3, 0.4), False), (6, 8, (0.3, 0.4), False), (7, 8, (0.3, 0.4), False), (3, 24, (1e-3, 2e-3), False)])
def test_negative_net_damping_after_a_solve_without_rebuild(capi, oracle_lib, cfg, n_sets, lams, pos_def):
    """Q2 with lambda < 1: BlockCholesky un-augments with lambda, not lambda^2 (BlockCholeskyLinearSystemSolver.cpp:77-97), so the second solve
    on the same built system sees diag(H) + l1^2 - l1 + l2^2, a NEGATIVE shift.  The solve must use exactly that: the same dx as the oracle
    where the shifted system is still positive definite, "not positive definite" where it is not (the reference's sticky
    linearSolverFailure; found by the reference pin, tests/test_reference_pin_gpu.py problem 3)."""
    p = make(cfg, n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    g.evaluate_error(); o.evaluate_error()
    g.build_system(); o.build_system()
    for i, lam in enumerate(lams):
        g.set_constant_conditioner(lam); o.set_constant_conditioner(lam)
        gdx, gok = g.solve_system()
        odx, ook = o.solve_system()
        assert gok == ook == (True if i == 0 else pos_def)
        if ook:
            assert rel_err(gdx, odx) < 1e-6
    assert lams[0] ** 2 - lams[0] + lams[1] ** 2 < 0.0


@pytest.mark.parametrize("cfg,n_sets", CASES)
def test_optimize_converges_like_oracle(capi, oracle_lib, cfg, n_sets):
    p = make(cfg, n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    gs, gtr = g.optimize(KbOptimizerOptions.kalibr2_default())
    os_, otr = o.optimize(KbOptimizerOptions.kalibr2_default())
    assert gs.iterations == os_.iterations
    assert gs.failed_iterations == os_.failed_iterations
    assert gs.linear_solver_failure == os_.linear_solver_failure == 0
    assert abs(gs.j_final - os_.j_final) <= 1e-9 * os_.j_final
    assert rel_err(gtr[:, 0], otr[:, 0]) < 1e-8
    assert rel_err(g.camera_params(), o.camera_params()) < REL_X
    assert rel_err(g.set_poses(), o.set_poses()) < REL_X
    if p.n_cams > 1:
        assert rel_err(g.baselines(), o.baselines()) < REL_X
    # and both recover the ground truth to the noise level
    truth = p.truth["cam_params"]
    assert np.abs(g.camera_params()[:, :4] - truth[:, :4]).max() < 5.0


@pytest.mark.parametrize("cfg,n_sets", [(2, 30), (3, 24)])
def test_speculative_and_plain_evaluate_agree(capi, cfg, n_sets):
    """kb_evaluate_error through the fused kernel (default) and through the residual-only kernel give the same cost, e()
    and the same LM trajectory."""
    p = make(cfg, n_sets)
    a = capi.B200SchurLinearSystemSolver(p)
    b = capi.B200SchurLinearSystemSolver(p)
    b.set_speculative_linearise(False)
    Ja, Jb = a.evaluate_error(), b.evaluate_error()
    assert abs(Ja - Jb) <= 1e-12 * Jb
    assert rel_err(a.error_vector(), b.error_vector()) < 1e-13
    sa, ta = a.optimize(KbOptimizerOptions.kalibr2_default())
    sb, tb = b.optimize(KbOptimizerOptions.kalibr2_default())
    assert sa.iterations == sb.iterations and sa.failed_iterations == sb.failed_iterations
    assert rel_err(ta[:, 0], tb[:, 0]) < 1e-12
    assert rel_err(a.camera_params(), b.camera_params()) < 1e-12


def test_rejected_step_keeps_the_built_system(capi, oracle_lib):
    """After a rejected step (update, evaluate, revert) the next solve must use the system built BEFORE the step, although the
    speculative evaluate has linearised at the trial state in between."""
    p = make(2, 20)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    for s in (g, o):
        s.evaluate_error()
        s.build_system()
        s.set_constant_conditioner(0.5)
        s.solve_system()
        s.apply_state_update()
        s.evaluate_error()            # trial state (speculative linearisation happens here on the GPU)
        s.revert_last_state_update()
        s.set_constant_conditioner(5.0)
    gdx, gok = g.solve_system()
    odx, ook = o.solve_system()
    assert gok == ook
    assert rel_err(gdx, odx) < 1e-7
    assert rel_err(g.rhs(), o.rhs()) < 1e-9


@pytest.mark.parametrize("cfg,n_sets", [(1, 40), (3, 24), (4, 37)])
def test_streamed_evaluate_matches_separate_upload_and_oracle(capi, oracle_lib, cfg, n_sets):
    """kb_evaluate_error_streamed (chunked upload overlapped with the fused kernel, second slice table) == kb_set_observations +
    kb_evaluate_error, and the system built from it matches the oracle on the new observations."""
    p = make(cfg, n_sets)
    rng = np.random.default_rng(5)
    yu = np.ascontiguousarray(p.y_u + rng.normal(0.0, 0.2, p.y_u.shape))
    yv = np.ascontiguousarray(p.y_v + rng.normal(0.0, 0.2, p.y_v.shape))
    a = capi.B200SchurLinearSystemSolver(p)
    b = capi.B200SchurLinearSystemSolver(p)
    a.evaluate_error()  # something queued before the streamed call (write-after-read ordering of the observation buffers)
    Ja = a.evaluate_error_streamed(yu, yv)
    b.set_observations(yu, yv)
    Jb = b.evaluate_error()
    assert abs(Ja - Jb) <= 1e-13 * Jb
    assert np.array_equal(a.error_vector(), b.error_vector())
    p2 = make(cfg, n_sets)
    p2.y_u[:] = yu
    p2.y_v[:] = yv
    o = oracle_lib.OracleProblem(p2)
    Jo = o.evaluate_error()
    assert abs(Ja - Jo) <= 1e-10 * Jo
    for s in (a, b, o):
        s.build_system()
        s.set_constant_conditioner(10.0)
    dxa, oka = a.solve_system()
    dxb, okb = b.solve_system()
    dxo, oko = o.solve_system()
    assert oka and okb and oko
    assert rel_err(dxa, dxb) < 1e-11
    assert rel_err(dxa, dxo) < 1e-7
    assert rel_err(a.rhs(), o.rhs()) < 1e-9


@pytest.mark.parametrize("cfg,n_sets,lam0", [(1, 40, 10.0), (2, 30, 10.0), (3, 24, 10.0), (4, 12, 1e-4), (2, 30, 1e-6), (3, 24, 1e3)])
def test_device_resident_loop_matches_host_loop_and_oracle(capi, oracle_lib, cfg, n_sets, lam0):
    """kb_optimize with the LM loop on the device (control kernels, no host round trip inside an iteration) walks exactly the
    same iterations as the host-side Optimizer2 mirror over the call-by-call entry points, and as the oracle."""
    p = make(cfg, n_sets)
    opt_dev = KbOptimizerOptions.kalibr2_default(device_loop=1)
    opt_host = KbOptimizerOptions.kalibr2_default(device_loop=0)
    opt_dev.lm_lambda_init = opt_host.lm_lambda_init = lam0
    a = capi.B200SchurLinearSystemSolver(p)
    b = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    sa, ta = a.optimize(opt_dev)
    sb, tb = b.optimize(opt_host)
    so, to = o.optimize(opt_host)
    for s in (sb, so):
        assert (sa.iterations, sa.failed_iterations, sa.linear_solver_failure) == (s.iterations, s.failed_iterations, s.linear_solver_failure)
    assert ta.shape == tb.shape == to.shape
    assert rel_err(ta[:, 2], tb[:, 2]) < 1e-9 and rel_err(ta[:, 2], to[:, 2]) < 1e-6   # lambda schedule
    assert rel_err(ta[:, 0], tb[:, 0]) < 1e-12 and rel_err(ta[:, 0], to[:, 0]) < 1e-9  # cost per iteration
    assert abs(sa.j_final - sb.j_final) <= 1e-12 * sb.j_final and abs(sa.j_final - so.j_final) <= 1e-9 * so.j_final
    assert abs(sa.dx_final - sb.dx_final) <= 1e-6 * max(sb.dx_final, 1e-12)
    assert rel_err(a.camera_params(), b.camera_params()) < 1e-10
    assert rel_err(a.camera_params(), o.camera_params()) < REL_X
    assert rel_err(a.set_poses(), b.set_poses()) < 1e-10
    # the call-by-call API keeps working on the same handle afterwards (neutral control flags, fresh linearisation)
    Ja, Jb = a.evaluate_error(), b.evaluate_error()
    assert abs(Ja - Jb) <= 1e-12 * Jb
    for s in (a, b):
        s.build_system()
        s.set_constant_conditioner(3.0)
    dxa, oka = a.solve_system()
    dxb, okb = b.solve_system()
    assert oka == okb and rel_err(dxa, dxb) < 1e-8


def _edit_views(p, empty=(), remove=()):
    """Copy of problem p with the views in `empty` emptied (zero terms, the view stays) and the views in `remove` dropped
    entirely (the set then has no image of that camera)."""
    from kalibr_b200.problem import Problem

    keep_term = np.ones(p.n_terms, bool)
    vs, vc, counts = [], [], []
    for w in range(p.n_views):
        b, e = int(p.view_begin[w]), int(p.view_begin[w + 1])
        if w in remove:
            keep_term[b:e] = False
            continue
        if w in empty:
            keep_term[b:e] = False
            counts.append(0)
        else:
            counts.append(e - b)
        vs.append(p.view_set[w])
        vc.append(p.view_cam[w])
    vb = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, p.target_points, np.array(vs, np.int32),
                   np.array(vc, np.int32), vb, p.y_u[keep_term], p.y_v[keep_term], p.corner_id[keep_term])


@pytest.mark.parametrize("cfg,n_sets,dropout,empty,remove", [
    (2, 12, 0.3, (), ()),            # ragged views: every view a different number of corners
    (3, 10, 0.45, (5,), ()),         # mixed models, ragged, one empty view
    (4, 6, 0.2, (3, 17), (9, 20, 21)),  # 8-camera chain with empty views and sets that miss cameras
    (2, 9, 0.9, (), (4,)),           # almost everything dropped: views of one to a few corners
])
def test_ragged_empty_and_missing_views_match_oracle(capi, oracle_lib, cfg, n_sets, dropout, empty, remove):
    """Edge cases of the observation structure: ragged views, empty views, sets without an image of some
    camera — residuals, CCS Jacobian, block pattern (bit exact), normal equations, solution and the LM run against the oracle."""
    p = _edit_views(make(cfg, n_sets, dropout=dropout), set(empty), set(remove))
    assert np.unique(np.diff(p.view_begin)).size > 2
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    Jg, Jo = g.evaluate_error(), o.evaluate_error()
    assert abs(Jg - Jo) <= 1e-11 * Jo
    assert rel_err(g.error_vector(), o.error_vector()) < REL_J
    g.linearise()
    gp, gi, gv = g.jacobian_ccs()
    op, oi, ov = o.jacobian_ccs()
    assert np.array_equal(gp, op) and np.array_equal(gi, oi)
    assert rel_err(gv, ov) < REL_J
    g.build_system(); o.build_system()
    assert rel_err(g.rhs(), o.rhs()) < REL_J
    g.set_constant_conditioner(10.0); o.set_constant_conditioner(10.0)
    gdx, gok = g.solve_system()
    odx, ook = o.solve_system()
    assert gok == ook
    gcp, gbr, gvp, gval = g.hessian_blocks()
    ocp, obr, ovp, oval = o.hessian_blocks()
    assert np.array_equal(gcp, ocp) and np.array_equal(gbr, obr) and np.array_equal(gvp, ovp)
    assert rel_err(gval, oval) < REL_J
    assert rel_err(gdx, odx) < 1e-7
    g.reset_state()
    sg, tg = g.optimize(KbOptimizerOptions.kalibr2_default())
    so, to = oracle_lib.OracleProblem(p).optimize(KbOptimizerOptions.kalibr2_default())
    assert (sg.iterations, sg.failed_iterations, sg.linear_solver_failure) == (so.iterations, so.failed_iterations, so.linear_solver_failure)
    assert abs(sg.j_final - so.j_final) <= 1e-9 * so.j_final


def test_double_buffered_observations(capi):
    """kb_prefetch_observations / kb_commit_observations: the committed batch is the one evaluated, batches alternate between the
    two device buffers, and a commit without a prefetch is a state error."""
    p = make(2, 20)
    rng = np.random.default_rng(11)
    batches = [(np.ascontiguousarray(p.y_u + rng.normal(0, 0.3, p.y_u.shape)), np.ascontiguousarray(p.y_v + rng.normal(0, 0.3, p.y_v.shape)))
               for _ in range(3)]
    a = capi.B200SchurLinearSystemSolver(p)
    b = capi.B200SchurLinearSystemSolver(p)
    with pytest.raises(capi.KalibrB200Error):
        a.commit_observations()
    J0 = a.evaluate_error()
    a.prefetch_observations(*batches[0])
    assert a.evaluate_error() == J0          # not committed yet: still the original observations
    for i, (yu, yv) in enumerate(batches):
        a.commit_observations()
        if i + 1 < len(batches):
            a.prefetch_observations(*batches[i + 1])
        b.set_observations(yu, yv)
        Ja, Jb = a.evaluate_error(), b.evaluate_error()
        assert Ja == Jb
        assert np.array_equal(a.error_vector(), b.error_vector())
        sa, _ = a.optimize(KbOptimizerOptions.kalibr2_default())
        sb, _ = b.optimize(KbOptimizerOptions.kalibr2_default())
        assert (sa.iterations, sa.j_final) == (sb.iterations, sb.j_final)
        a.reset_state(); b.reset_state()


@pytest.mark.parametrize("cfg,n_sets", [(1, 40), (2, 30), (3, 24), (4, 12), (5, 6), (6, 20)])
def test_marginal_analysis_matches_oracle(capi, oracle_lib, cfg, n_sets):
    """kb_analyze_marginal (undamped Schur-reduced camera system + one-sided Jacobi on the device) against the oracle's
    restatement of LinearSolver::analyzeMarginal: singular values, rank, tolerance, log2-sum, singular subspaces."""
    p = make(cfg, n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    g.optimize(KbOptimizerOptions.kalibr2_default())
    o.optimize(KbOptimizerOptions.kalibr2_default())
    rg, svg, Vg, cg = g.analyze_marginal()
    ro, svo, Vo, co, om = o.analyze_marginal()
    assert np.array_equal(cg, co)
    assert (rg.n, rg.rank, rg.rank_deficiency) == (ro.n, ro.rank, ro.rank_deficiency)
    assert np.abs(svg - svo).max() <= 1e-9 * svo[0]
    assert abs(rg.tolerance - ro.tolerance) <= 1e-9 * ro.tolerance
    assert abs(rg.sv_log2_sum - ro.sv_log2_sum) <= 1e-6 * abs(ro.sv_log2_sum)
    assert np.abs(Vg.T @ Vg - np.eye(rg.n)).max() < 1e-10
    assert np.abs(Vg @ np.diag(svg) @ Vg.T - om).max() <= 1e-9 * svo[0]
    # a solve right after the analysis works on the system it has built
    g.set_constant_conditioner(10.0)
    o.build_system()
    o.set_constant_conditioner(10.0)
    dxg, okg = g.solve_system()
    dxo, oko = o.solve_system()
    assert okg and oko and rel_err(dxg, dxo) < 1e-6


def test_marginal_analysis_rank_deficient_camera(capi, oracle_lib):
    """Stereo rig in which camera 1 has no image at all: its intrinsics (8) and the baseline (6) are unobservable."""
    p0 = make(2, 16)
    p = _edit_views(p0, remove={w for w in range(p0.n_views) if p0.view_cam[w] == 1})
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    rg, svg, Vg, cg = g.analyze_marginal()
    ro, svo, Vo, co, om = o.analyze_marginal()
    assert (rg.rank, rg.rank_deficiency) == (ro.rank, ro.rank_deficiency) == (8, 14)
    assert np.abs(svg[:8] - svo[:8]).max() <= 1e-9 * svo[0]
    assert svg[8:].max() <= rg.tolerance
    assert abs(rg.sv_log2_sum - ro.sv_log2_sum) <= 1e-6 * abs(ro.sv_log2_sum)
    # the null space is spanned by the unit vectors of camera 1's intrinsics and the baseline
    null_cols = np.abs(Vg[:, 8:]).sum(axis=1) > 1e-6
    expected = np.isin(cg, np.setdiff1d(cg, cg[:8])) if False else ~np.isin(np.arange(rg.n), np.arange(8))
    assert np.array_equal(null_cols, expected)


def test_single_precision_observations_are_widened_exactly(capi):
    """kb_*_observations_f32: measurements in the detector's precision (cv::Point2f in the reference) give bit-identical results to the
    same values passed as doubles - set, streamed and prefetched."""
    p = synthetic.make_config(3, n_sets=12, float_corners=True)
    yu32, yv32 = p.y_u.astype(np.float32), p.y_v.astype(np.float32)
    assert np.array_equal(yu32.astype(np.float64), p.y_u)
    g = capi.B200SchurLinearSystemSolver(p)
    J0, e0 = g.evaluate_error(), g.error_vector()
    rng = np.random.default_rng(5)
    shift = (rng.normal(0, 0.2, p.n_terms)).astype(np.float32)
    au32, av32 = (yu32 + shift).astype(np.float32), (yv32 - shift).astype(np.float32)
    au64, av64 = au32.astype(np.float64), av32.astype(np.float64)
    g.set_observations(au64, av64)
    J1, e1 = g.evaluate_error(), g.error_vector()
    assert J1 != J0
    g.set_observations(yu32, yv32)          # back to the original values, through the f32 path
    assert g.evaluate_error() == J0 and np.array_equal(g.error_vector(), e0)
    g.set_observations(au32, av32)
    assert g.evaluate_error() == J1 and np.array_equal(g.error_vector(), e1)
    Js = g.evaluate_error_streamed(yu32, yv32)
    assert abs(Js - J0) <= 1e-13 * J0 and np.array_equal(g.error_vector(), e0)   # the chunked evaluation sums the cost in another order
    g.prefetch_observations(au32, av32)
    g.commit_observations()
    assert g.evaluate_error() == J1 and np.array_equal(g.error_vector(), e1)
    with pytest.raises(capi.KalibrB200Error):
        g.set_observations(au32, av64)


def test_iterate_equals_the_six_separate_calls(capi):
    """kb_iterate = evaluateError, buildSystem, setConstantConditioner, solveSystem, applyStateUpdate [, revert] with one synchronisation:
    the same scalars and the same state as the separate calls."""
    p = make(3, 14)
    a, b = capi.B200SchurLinearSystemSolver(p), capi.B200SchurLinearSystemSolver(p)
    for lam, revert in [(10.0, False), (3.0, True), (1.0, False)]:
        J = a.evaluate_error()
        a.build_system()
        a.set_constant_conditioner(lam)
        _, ok = a.solve_system(fetch_dx=False)
        rho = a.lm_rho_denominator(lam)
        m = a.apply_state_update()
        if revert:
            a.revert_last_state_update()
        cost, rho2, m2, ok2 = b.iterate(lam, revert=revert)
        assert (cost, rho2, m2, ok2) == (J, rho, m, ok)
        assert np.array_equal(a.camera_params(), b.camera_params()) and np.array_equal(a.set_poses(), b.set_poses())
    # enqueue-only iterations (out == NULL) followed by one kb_wait: the scalars of the LAST one, the state after all of them
    c = capi.B200SchurLinearSystemSolver(p)
    for lam, revert in [(10.0, False), (3.0, True)]:
        assert c.iterate(lam, revert=revert, wait=False) is None
    c.iterate(1.0, revert=False, wait=False)
    assert c.wait_iterations() == (cost, rho2, m2, ok2)
    assert np.array_equal(c.camera_params(), b.camera_params()) and np.array_equal(c.set_poses(), b.set_poses())
