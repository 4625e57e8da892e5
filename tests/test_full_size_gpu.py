"""Parity at the BASELINE configs' STATED sizes (BASELINE.json configs[0..2]) against the CPU oracle, and the reference's edge cases
on the device.  The scaled-down suites (test_parity_gpu.py) exercise every code path; these runs exercise what only size shows:
thousands of slices / CTA partials, the 5 000- and 2 000-set summation orders, several views per warp slice.
cfg 4 / cfg 5 (minutes of oracle time) are compared once per round by tools/full_size_check.py -> profiles/.
"""
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


@pytest.mark.parametrize("cfg", [1, 2])
def test_full_size_optimize_matches_oracle(capi, oracle_lib, cfg):
    """BASELINE configs[0] (300 views, 36 000 terms) and configs[1] (2 000 synced sets, 480 000 terms): the whole LM run."""
    p = synthetic.make_config(cfg)
    assert p.n_sets == synthetic.CONFIGS[cfg][2]
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p, n_threads=16)
    gs, gt = g.optimize(KbOptimizerOptions.kalibr2_default())
    os_, ot = o.optimize(KbOptimizerOptions.kalibr2_default())
    assert gs.iterations == os_.iterations and gs.failed_iterations == os_.failed_iterations
    assert gs.linear_solver_failure == os_.linear_solver_failure == 0
    assert abs(gs.j_start - os_.j_start) <= 1e-11 * os_.j_start
    assert abs(gs.j_final - os_.j_final) <= 1e-9 * os_.j_final
    assert gt.shape == ot.shape
    assert rel(gt[:, 0], ot[:, 0]) < 1e-9 and rel(gt[:, 2], ot[:, 2]) < 1e-6  # cost and lambda per iteration
    oc = o.camera_params()
    assert (np.abs(g.camera_params() - oc) / np.maximum(np.abs(oc), 1e-3)).max() < 1e-6
    if p.n_cams > 1:
        assert np.abs(g.baselines() - o.baselines()).max() < 1e-6
    assert np.abs(g.set_poses() - o.set_poses()).max() < 1e-6
    assert g.num_invalid_terms() == 0


def test_cfg3_full_size_evaluate_build_solve_matches_oracle(capi, oracle_lib):
    """BASELINE configs[2] (4 mixed cameras, 5 000 synced sets, 2.4 M terms): one evaluate + build + solve(lambda = 10) against the
    SparseCholesky-semantic oracle (the reference's solver for 0-dim distortion variables), threaded."""
    p = synthetic.make_config(3)
    assert p.n_sets == 5000 and p.n_terms == 2_400_000
    g = capi.B200SchurLinearSystemSolver(p)
    g.set_solver_semantic(1)
    o = oracle_lib.OracleProblem(p, solver_kind=oracle_lib.SPARSE_CHOLESKY, n_threads=16)
    Jg, Jo = g.evaluate_error(), o.evaluate_error()
    assert abs(Jg - Jo) <= 1e-11 * abs(Jo)
    assert rel(g.error_vector(), o.error_vector()) < 1e-9
    g.build_system()
    o.build_system()
    assert rel(g.rhs(), o.rhs()) < 1e-9
    g.set_constant_conditioner(10.0)
    o.set_constant_conditioner(10.0)
    gdx, gok = g.solve_system()
    odx, ook = o.solve_system()
    assert gok and ook
    assert rel(gdx, odx) < 1e-7
    assert g.num_invalid_terms() == 0


# ---- the reference's edge cases, on the device ---------------------------------------------------------------------------------
def _behind_camera_problem(p, set_idx, cams):
    """p plus one extra target point that lies 2 m BEHIND camera 0 of synced set `set_idx` (on its optical axis), observed - with
    an arbitrary measurement - by the cameras in `cams`: for the omni / EUCM / double-sphere models z <= -fov * |p| there, the
    projection returns false before writing y_hat (OmniProjection.hpp:143-144, ExtendedUnifiedProjection.hpp:170-171,
    DoubleSphereProjection.hpp:181-182).  Returns (problem with the extra terms, the same problem without them,
    indices of the extra terms)."""
    from oracle import oracle_api

    q, t = p.set_poses[set_idx, :4], p.set_poses[set_idx, 4:]
    C = oracle_api.quat2r(q)  # T_target_cam0: p_t = C p_c + t
    extra = C @ np.array([0.0, 0.0, -2.0]) + t
    target = np.vstack([p.target_points, extra[None]])
    new_id = len(p.target_points)
    yu, yv, cid, vb, bad = [], [], [], [0], []
    for w in range(len(p.view_set)):
        b, e = int(p.view_begin[w]), int(p.view_begin[w + 1])
        yu += list(p.y_u[b:e]); yv += list(p.y_v[b:e]); cid += list(p.corner_id[b:e])
        if p.view_set[w] == set_idx and p.view_cam[w] in cams:
            # not at the end of the view: lanes after it in the chunk must be unaffected
            pos = len(yu) - 7
            yu.insert(pos, 311.5); yv.insert(pos, 207.25); cid.insert(pos, new_id)
            bad.append(pos)
        vb.append(len(yu))
    withp = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, target, p.view_set, p.view_cam, np.array(vb, np.int64),
                    np.array(yu), np.array(yv), np.array(cid, np.int32))
    clean = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, target, p.view_set, p.view_cam, p.view_begin, p.y_u, p.y_v,
                    p.corner_id)
    return withp, clean, np.array(bad)


@pytest.mark.parametrize("speculative", [True, False])
def test_projection_that_bails_out_is_zero_weighted_and_counted(capi, oracle_lib, speculative):
    """Q6: a term whose projection returns false contributes nothing (e = 0, zero rows) and is counted; everything else equals the
    oracle on the problem WITHOUT those terms - cost, e(), rhs, H blocks (pattern bit-exact), dx."""
    base = synthetic.make_config(3, n_sets=10)  # omni-radtan, eucm-none, ds-none, pinhole-equi
    p, clean, bad = _behind_camera_problem(base, set_idx=4, cams=(0, 1, 2))
    assert len(bad) == 3
    g = capi.B200SchurLinearSystemSolver(p)
    g.set_speculative_linearise(speculative)
    g.set_solver_semantic(1)
    o = oracle_lib.OracleProblem(clean, solver_kind=oracle_lib.SPARSE_CHOLESKY)
    Jg, Jo = g.evaluate_error(), o.evaluate_error()
    assert g.num_invalid_terms() == 3
    assert abs(Jg - Jo) <= 1e-11 * abs(Jo)
    eg = g.error_vector().reshape(-1, 2)
    assert np.all(eg[bad] == 0.0)
    keep = np.ones(len(eg), bool)
    keep[bad] = False
    assert rel(eg[keep].ravel(), o.error_vector()) < 1e-9
    g.build_system()
    o.build_system()
    assert rel(g.rhs(), o.rhs()) < 1e-9
    g.set_constant_conditioner(10.0)
    o.set_constant_conditioner(10.0)
    gdx, gok = g.solve_system()
    odx, ook = o.solve_system()
    assert gok and ook and rel(gdx, odx) < 1e-7
    # the exported Jacobian has all-zero rows for the bailed-out terms, the oracle's rows elsewhere
    gp, gi, gv = g.jacobian_ccs()
    op, oi, ov = o.jacobian_ccs()
    rows_keep = np.repeat(keep, 2)
    lens = np.diff(gp)
    for r in np.flatnonzero(~rows_keep):
        assert np.all(gv[gp[r]:gp[r + 1]] == 0.0)
    sel = np.repeat(rows_keep, lens)
    assert np.array_equal(lens[rows_keep], np.diff(op)) and np.array_equal(gi[sel], oi)
    assert np.abs(gv[sel] - ov).max() <= 1e-9 * np.abs(ov).max()
    # block pattern and values of H with the BlockCholesky-capable part of the rig: same check on a pinhole + omni-radtan stereo pair
    # (4-dim distortion blocks, so the reference's BlockCholesky path is defined)
    g.close()


def test_bailed_out_term_in_the_block_pattern(capi, oracle_lib):
    """Same as above for H itself (block pattern bit-exact, values 1e-9) on a rig whose reference solver is BlockCholesky:
    omni-radtan + pinhole-radtan (cfg 6 without its omni-none camera would change the order; built directly)."""
    base = synthetic.make_problem([synthetic.OMNI_RADTAN, synthetic.PINHOLE_RADTAN], 8, 2, seed=4242)
    p, clean, bad = _behind_camera_problem(base, set_idx=2, cams=(0,))
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(clean)
    g.evaluate_error(); o.evaluate_error()
    assert g.num_invalid_terms() == 1
    g.build_system(); o.build_system()
    g.set_constant_conditioner(10.0); o.set_constant_conditioner(10.0)
    gdx, gok = g.solve_system()
    odx, ook = o.solve_system()
    assert gok and ook and rel(gdx, odx) < 1e-7
    gc, gr, gvp, gv = g.hessian_blocks()
    oc, orow, ovp, ov = o.hessian_blocks()
    assert np.array_equal(gc, oc) and np.array_equal(gr, orow) and np.array_equal(gvp, ovp)
    assert np.abs(gv - ov).max() <= 1e-9 * np.abs(ov).max()


def _on_axis_problem():
    """pinhole-equi, one synced set re-posed so that target corner 0 (the target's origin) lies EXACTLY on the optical axis:
    T_target_cam = (identity rotation, t = (0, 0, -0.6)) gives p_c = p_t - t = (0, 0, 0.6) for p_t = 0."""
    p = synthetic.make_problem([synthetic.PINHOLE_EQUI], 6, 0, seed=515)
    poses = p.set_poses.copy()
    poses[3] = [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, -0.6]
    q = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, poses, p.target_points, p.view_set, p.view_cam, p.view_begin, p.y_u, p.y_v,
                p.corner_id)
    w = int(np.flatnonzero(q.view_set == 3)[0])
    b, e = int(q.view_begin[w]), int(q.view_begin[w + 1])
    on_axis = b + int(np.flatnonzero(q.corner_id[b:e] == 0)[0])
    return q, on_axis


def test_equidistant_jacobian_on_the_axis_is_nan_like_the_reference(capi, oracle_lib):
    """Q5: EquidistantDistortion::distort(y, J) divides by r without the guard its value path has (EquidistantDistortion.hpp:54-85 vs
    :180): at r = 0 the residual is finite and the Jacobian is NaN.  The device reproduces that, entry for entry."""
    p, term = _on_axis_problem()
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    Jg, Jo = g.evaluate_error(), o.evaluate_error()
    assert np.isfinite(Jo) and abs(Jg - Jo) <= 1e-11 * abs(Jo)  # the value path is guarded
    eg, eo = g.error_vector(), o.error_vector()
    assert np.all(np.isfinite(eo)) and rel(eg, eo) < 1e-9
    gp, gi, gv = g.jacobian_ccs()
    op, oi, ov = o.jacobian_ccs()
    assert np.array_equal(gp, op) and np.array_equal(gi, oi)
    nan_o = np.isnan(ov)
    assert nan_o.any(), "the oracle (restating the reference) must produce NaN here"
    rows_with_nan = np.unique(np.searchsorted(op, np.flatnonzero(nan_o), side="right") - 1)
    assert set(rows_with_nan) <= {2 * term, 2 * term + 1}
    assert np.array_equal(np.isnan(gv), nan_o)
    fin = ~nan_o
    assert np.abs(gv[fin] - ov[fin]).max() <= 1e-9 * np.abs(ov[fin]).max()
    assert g.num_invalid_terms() == 0  # not a validity bail-out: the reference's pinhole always writes y_hat
    # the normal equations inherit the NaN in the same places, and the solve reports failure on both sides
    g.build_system(); o.build_system()
    rg, ro = g.rhs(), o.rhs()
    assert np.array_equal(np.isnan(rg), np.isnan(ro)) and np.isnan(ro).any()
    f = ~np.isnan(ro)
    assert np.abs(rg[f] - ro[f]).max() <= 1e-9 * np.abs(ro[f]).max()
    g.set_constant_conditioner(10.0); o.set_constant_conditioner(10.0)
    _, gok = g.solve_system()
    _, ook = o.solve_system()
    assert gok == ook == False  # noqa: E712


@pytest.mark.parametrize("device_loop", [1, 0])
def test_failed_solves_are_sticky_like_optimizer2(capi, oracle_lib, device_loop):
    """A solve that is not positive definite (here: through the NaN of the on-axis equidistant term) is a FAILED iteration:
    no update, lambda grows, linearSolverFailure stays set and keeps the loop alive until failedIterations reaches maxIterations
    (Optimizer2.cpp:208-229) - same counts on the device-resident loop, the host mirror and the oracle."""
    p, _ = _on_axis_problem()
    opt = KbOptimizerOptions.kalibr2_default(device_loop=device_loop)
    opt.max_iterations = 9
    g = capi.B200SchurLinearSystemSolver(p)
    o = oracle_lib.OracleProblem(p)
    gs, gt = g.optimize(opt)
    os_, ot = o.optimize(opt)
    assert os_.linear_solver_failure == 1 and os_.iterations == 0 and os_.failed_iterations == 9
    assert (gs.iterations, gs.failed_iterations, gs.linear_solver_failure) == (os_.iterations, os_.failed_iterations, os_.linear_solver_failure)
    assert gs.j_final == gs.j_start and abs(gs.j_final - os_.j_final) <= 1e-11 * os_.j_final  # nothing was applied
    assert np.array_equal(g.camera_params(), p.cam_params) and np.array_equal(g.set_poses(), p.set_poses)


@pytest.mark.parametrize("cfg", [4, 5])
def test_full_size_against_the_stored_oracle_outputs(capi, cfg):
    """BASELINE configs[3] (8 cameras, 20 000 sets, 19.2 M terms) and configs[4] (16 cameras, 6 250 sets, 12 M terms; n_c = 218 splits
    schur_kernel's tile pairs over gridDim.y) at FULL size: cost, e(), rhs, dx(lambda = 10), the block pattern (sha256) and a weighted
    checksum of the H values against what the oracle produced for the same seeded problem (tools/full_size_oracle.py, minutes of CPU,
    stored in tests/golden/full_size_cfgN.npz)."""
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import full_size_oracle as fso

    ref = np.load(os.path.join(root, "tests", "golden", f"full_size_cfg{cfg}.npz"))
    p = synthetic.make_config(cfg)
    assert p.n_terms == int(ref["n_terms"])
    g = capi.B200SchurLinearSystemSolver(p)
    J = g.evaluate_error()
    e = g.error_vector()
    g.build_system()
    rhs = g.rhs()
    g.set_constant_conditioner(10.0)
    dx, ok = g.solve_system()
    s = fso.summarise(J, e, rhs, dx, ok, g.hessian_blocks())
    assert abs(s["J"] - ref["J"]) <= 1e-11 * ref["J"]
    assert rel(s["e_sample"], ref["e_sample"]) < 1e-9
    assert abs(s["e_checksum"] - ref["e_checksum"]) <= 1e-9 * ref["e_abs_sum"]
    assert rel(s["rhs"], ref["rhs"]) < 1e-9
    assert int(s["pos_def"]) == int(ref["pos_def"]) == 1
    assert rel(s["dx"], ref["dx"]) < 1e-7
    assert int(s["n_blocks"]) == int(ref["n_blocks"]) and int(s["n_values"]) == int(ref["n_values"])
    assert np.array_equal(s["pattern_sha256"], ref["pattern_sha256"]), "block pattern differs from the oracle's"
    assert abs(s["h_checksum"] - ref["h_checksum"]) <= 1e-9 * ref["h_abs_sum"]
    assert g.num_invalid_terms() == 0
