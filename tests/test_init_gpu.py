"""GPU parity of the initial-guess stage (kb_estimate_transformations, kb_initialize_set_poses, kb_estimate_stereo_baseline)
through the C ABI, against cv::solvePnP outputs (tests/golden/pnp_cv2.npz, generated with cv2 4.13) and the numpy oracle.

Bar: poses within 1e-6 (OpenCV itself stops its Levenberg-Marquardt at a relative parameter change of FLT_EPSILON; the
device and the oracle iterate to convergence and agree to ~1e-10 with each other).
"""
import os

import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions, Problem
from oracle import ko_init as ki

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pnp_cv2.npz")
CASES = {"cfg1_S6": (1, 6, {}), "cfg2_S9": (2, 9, {}), "cfg3_S5": (3, 5, {}), "cfg4_S2": (4, 2, {}), "cfg6_S4": (6, 4, {}),
         "cfg7_S7": (7, 7, {}), "cfg3_S4_ragged": (3, 4, {"dropout": 0.3})}
TOL_CV = 1e-6
TOL_ORACLE = 1e-8


def pose_err(a, b):
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    return max((np.abs(ki.pose_to_T(x) - ki.pose_to_T(y)).max() for x, y in zip(a, b)), default=0.0)


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()
    return m


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", list(CASES))
def test_view_transformations_match_opencv_and_oracle(capi, golden, name):
    cfg, S, kw = CASES[name]
    p = synthetic.make_config(cfg, n_sets=S, **kw)
    g = capi.B200SchurLinearSystemSolver(p)
    T, ok = g.estimate_transformations()
    assert np.array_equal(ok, golden[name + "/ok"])
    assert pose_err(T, golden[name + "/T_views"]) < TOL_CV
    assert np.all(T[:, 3] >= 0) and np.abs(np.linalg.norm(T[:, :4], axis=1) - 1).max() < 1e-12  # r2quat: unit, w >= 0
    To, oko = ki.view_transformations(p)
    assert np.array_equal(ok, oko) and pose_err(T, To) < TOL_ORACLE


@pytest.mark.parametrize("name", list(CASES))
def test_set_pose_guesses_match_opencv_and_oracle(capi, golden, name):
    cfg, S, kw = CASES[name]
    p = synthetic.make_config(cfg, n_sets=S, **kw)
    g = capi.B200SchurLinearSystemSolver(p)
    assert g.initialize_set_poses() == int((~golden[name + "/set_ok"]).sum())
    sp = g.set_poses()
    assert pose_err(sp, golden[name + "/set_poses"]) < TOL_CV
    so, _ = ki.target_pose_guesses(p)
    assert pose_err(sp, so) < TOL_ORACLE
    g.reset_state()  # the guesses are the new initial state
    assert np.array_equal(g.set_poses(), sp)
    if p.n_cams >= 2:
        b, n = g.estimate_stereo_baseline(0, 1)
        assert n == S
        assert pose_err(b, golden[name + "/baseline01"]) < TOL_CV
        assert pose_err(b, ki.stereo_baseline_guess(p, 0, 1)) < TOL_ORACLE


def test_reference_like_start_then_optimize(capi, oracle_lib):
    """The drivers' flow: focal-length guess and zero distortion -> PnP per set -> batch optimisation (CalibrationTools.hpp:93-144).
    The PnP poses come from the device, the optimisation from the same handle; the oracle optimises from the same start."""
    p = synthetic.make_config(1, n_sets=30)
    truth = np.asarray(synthetic.TRUTH_PARAMS[p.cam_model[0]][0], float)
    cam0 = np.zeros((1, 10))
    cam0[0, :4] = [390.0, 390.0, 319.5, 239.5]  # f0 from the vanishing points, image centre, distortion cleared
    q = Problem(p.driver_order, p.cam_model, cam0, p.baselines, np.tile([0, 0, 0, 1.0, 0, 0, 0], (p.n_sets, 1)), p.target_points, p.view_set,
                p.view_cam, p.view_begin, p.y_u, p.y_v, p.corner_id)
    g = capi.B200SchurLinearSystemSolver(q)
    assert g.initialize_set_poses(resolution=[[640, 480]]) == 0
    sp = g.set_poses()
    so, good = ki.target_pose_guesses(q, resolution=[(640, 480)])
    assert good.all() and pose_err(sp, so) < TOL_ORACLE
    gs, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
    q2 = Problem(q.driver_order, q.cam_model, cam0, q.baselines, sp, q.target_points, q.view_set, q.view_cam, q.view_begin, q.y_u, q.y_v, q.corner_id)
    os_, _ = oracle_lib.OracleProblem(q2).optimize(KbOptimizerOptions.kalibr2_default())
    assert gs.iterations == os_.iterations and abs(gs.j_final - os_.j_final) <= 1e-9 * os_.j_final
    assert np.abs(g.camera_params()[0, :4] / truth[:4] - 1).max() < 5e-3  # and it calibrates the camera


def test_edge_cases(capi):
    p = synthetic.make_config(2, n_sets=4)
    # view 0 keeps 3 corners (PnP refuses), view 1 keeps 4 (minimum), set 2 loses camera 0 entirely, set 3 loses both cameras
    lens = np.diff(p.view_begin).copy()
    keep = np.ones(p.n_terms, bool)
    w00 = int(np.flatnonzero((p.view_set == 0) & (p.view_cam == 0))[0])
    w01 = int(np.flatnonzero((p.view_set == 0) & (p.view_cam == 1))[0])
    keep[p.view_begin[w00] + 3:p.view_begin[w00 + 1]] = False
    keep[p.view_begin[w01] + 2:p.view_begin[w01 + 1]] = False   # 2 corners: set 0 has no usable view at all
    w10 = int(np.flatnonzero((p.view_set == 1) & (p.view_cam == 0))[0])
    sel = np.arange(p.view_begin[w10], p.view_begin[w10 + 1])
    keep[sel] = False
    keep[sel[[0, 11, 60, 119]]] = True                          # four spread-out corners
    view_keep = ~(((p.view_set == 2) & (p.view_cam == 0)) | (p.view_set == 3))
    for w in np.flatnonzero(~view_keep):
        keep[p.view_begin[w]:p.view_begin[w + 1]] = False
    lens = np.array([keep[p.view_begin[w]:p.view_begin[w + 1]].sum() for w in range(len(p.view_set))])
    q = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, p.target_points, p.view_set[view_keep], p.view_cam[view_keep],
                np.concatenate([[0], np.cumsum(lens[view_keep])]).astype(np.int64), p.y_u[keep], p.y_v[keep], p.corner_id[keep])
    g = capi.B200SchurLinearSystemSolver(q)
    T, ok = g.estimate_transformations()
    To, oko = ki.view_transformations(q)
    assert np.array_equal(ok, oko) and (~ok).sum() == 2          # the views with 3 and with 2 corners
    assert pose_err(T[ok], To[ok]) < 1e-7
    assert np.array_equal(T[~ok], np.tile([0, 0, 0, 1.0, 0, 0, 0], ((~ok).sum(), 1)))
    before = g.set_poses()
    n_failed = g.initialize_set_poses()
    so, good = ki.target_pose_guesses(q)
    assert n_failed == int((~good).sum()) == 2
    sp = g.set_poses()
    assert pose_err(sp[good], so[good]) < 1e-7
    assert np.array_equal(sp[3], before[3])                      # nobody saw set 3: its pose is left alone
    assert pose_err(sp[0], so[0]) < 1e-12                        # failed PnP: identity (chained), as the reference
    b, n = g.estimate_stereo_baseline(0, 1)
    assert n == 1 and pose_err(b, ki.stereo_baseline_guess(q, 0, 1)) < 1e-7


# ---- initializeIntrinsics --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CASES))
def test_initialize_intrinsics_matches_golden_and_oracle(capi, golden, name):
    cfg, S, kw = CASES[name]
    p = synthetic.make_config(cfg, n_sets=S, **kw)
    g = capi.B200SchurLinearSystemSolver(p)
    res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
    before = g.camera_params()
    for k in range(p.n_cams):
        prm, ok = g.initialize_intrinsics(k, 10, 12, res)
        assert ok == bool(golden[name + f"/init_ok{k}"])
        gold = golden[name + f"/init_params{k}"]
        if ok:
            assert np.abs(prm - gold).max() <= 1e-8 * max(np.abs(gold).max(), 1.0), (k, prm, gold)
            assert np.array_equal(g.camera_params()[k], prm)              # the guess is the camera's state now
        else:
            assert np.array_equal(g.camera_params()[k], before[k])
    g.reset_state()
    for k in range(p.n_cams):
        if bool(golden[name + f"/init_ok{k}"]):
            assert np.abs(g.camera_params()[k] - golden[name + f"/init_params{k}"]).max() <= 1e-8 * 1e3   # ... and the reset point


def test_initialize_intrinsics_fallbacks_and_errors(capi):
    p = synthetic.make_config(3, n_sets=2, dropout=0.7)
    g = capi.B200SchurLinearSystemSolver(p)
    res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
    for k in range(p.n_cams):
        prm_o, ok_o = ki.initialize_intrinsics(p, k, 10, 12, res[k])
        prm, ok = g.initialize_intrinsics(k, 10, 12, res)
        assert ok == ok_o
        if ok:
            assert np.abs(prm - prm_o).max() <= 1e-8 * max(np.abs(prm_o).max(), 1.0)
        prm_o, ok_o = ki.initialize_intrinsics(p, k, 10, 12, res[k], fallback=450.0)
        prm, ok = g.initialize_intrinsics(k, 10, 12, res, fallback_focal_length=450.0)
        assert ok == ok_o
        if ok or p.cam_model[k] in (2, 6):  # omni sets the fallback but still reports failure
            assert np.abs(prm - prm_o).max() <= 1e-8 * max(np.abs(prm_o).max(), 1.0)
    with pytest.raises(capi.KalibrB200Error):
        g.initialize_intrinsics(0, 9, 12, res)      # rows x cols must match the target
    with pytest.raises(capi.KalibrB200Error):
        g.initialize_intrinsics(7, 10, 12, res)


def test_full_initialisation_then_calibration(capi, oracle_lib):
    """CalibrateSingleCamera end to end on the device (CalibrationTools.hpp:93-144): initializeIntrinsics -> estimateTransformation per
    view -> Optimizer2; the oracle optimises from the device's initial guesses."""
    for model, S in ((1, 30), (6, 30)):  # pinhole-equi, omni-none
        p = synthetic.make_problem([model], S, 0, seed=77 + model)
        res = [synthetic.TRUTH_PARAMS[model][1]]
        q = Problem(p.driver_order, p.cam_model, np.zeros((1, 10)) + 1.0, p.baselines, np.tile([0, 0, 0, 1.0, 0, 0, 0], (p.n_sets, 1)), p.target_points,
                    p.view_set, p.view_cam, p.view_begin, p.y_u, p.y_v, p.corner_id)
        g = capi.B200SchurLinearSystemSolver(q)
        prm, ok = g.initialize_intrinsics(0, 10, 12, res)
        assert ok
        assert g.initialize_set_poses(res) == 0
        sp = g.set_poses()
        gs, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
        q2 = Problem(q.driver_order, q.cam_model, prm[None], q.baselines, sp, q.target_points, q.view_set, q.view_cam, q.view_begin, q.y_u, q.y_v, q.corner_id)
        os_, _ = oracle_lib.OracleProblem(q2).optimize(KbOptimizerOptions.kalibr2_default())
        assert gs.iterations == os_.iterations and abs(gs.j_final - os_.j_final) <= 1e-9 * os_.j_final
        truth = np.asarray(synthetic.TRUTH_PARAMS[model][0], float)
        P = 4 if model == 1 else 5
        assert np.abs(g.camera_params()[0, P - 4:P] / truth[P - 4:P] - 1).max() < 5e-2, (model, g.camera_params()[0], truth)  # xi and f trade off
