"""CPU tests of the initial-guess stage's oracle (oracle/ko_init.py).

PINNED against the reference's third-party dependency: cv::solvePnP outputs generated with cv2 4.13
(tests/golden/pnp_cv2.npz, tests/golden/make_pnp_golden.py), and live cv2.solvePnP when cv2 is importable.
  CAM/.../implementation/PinholeProjection.hpp:831-891 (+ omni / eucm / ds)   estimateTransformation
  K2/include/kalibr2/CalibrationTools.hpp:195-234, 316-356                    stereo baseline guess, getTargetPoseGuess
"""
import os

import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import MODEL_D, MODEL_P
from oracle import ko_init as ki

EUCM, DS = 3, 4

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pnp_cv2.npz")
CASES = {"cfg1_S6": (1, 6, {}), "cfg2_S9": (2, 9, {}), "cfg3_S5": (3, 5, {}), "cfg4_S2": (4, 2, {}), "cfg6_S4": (6, 4, {}),
         "cfg7_S7": (7, 7, {}), "cfg3_S4_ragged": (3, 4, {"dropout": 0.3})}
TOL = 1e-6  # OpenCV stops its LM at a relative parameter change of FLT_EPSILON


def pose_err(a, b):
    """max abs difference between the 4x4 matrices of two (q, t) poses (q and -q are the same rotation)."""
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    return max(np.abs(ki.pose_to_T(x) - ki.pose_to_T(y)).max() for x, y in zip(a, b))


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", list(CASES))
def test_solve_pnp_matches_opencv_golden(golden, name):
    r, t = ki.solve_pnp(golden[name + "/Ps"], golden[name + "/Ms"])
    assert np.abs(r - golden[name + "/rvec"]).max() < TOL and np.abs(t - golden[name + "/tvec"]).max() < TOL


def test_solve_pnp_matches_opencv_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    pts = synthetic.aprilgrid_points()
    for trial in range(12):
        rvec = rng.normal(0, 0.35, 3)
        tvec = np.array([rng.uniform(-0.3, 0.1), rng.uniform(-0.3, 0.1), rng.uniform(0.4, 1.2)])
        keep = rng.random(len(pts)) > (0.0 if trial % 3 else 0.5)
        P = np.float32(pts[keep]).astype(float)
        pc = P @ ki.rodrigues(rvec).T + tvec
        M = np.float32(pc[:, :2] / pc[:, 2:3] + rng.normal(0, 1e-3, (len(P), 2))).astype(float)
        r, t = ki.solve_pnp(P, M)
        ok, r2, t2 = cv2.solvePnP(np.float32(P), np.float32(M), np.eye(3), np.zeros(4))
        assert ok and np.abs(r - r2.ravel()).max() < TOL and np.abs(t - t2.ravel()).max() < TOL


@pytest.mark.parametrize("model", range(7))
def test_keypoint_to_euclidean_inverts_the_projection(oracle_lib, model):
    """CameraGeometryTestHarness.hpp: keypointToEuclidean(euclideanToKeypoint(p)) is parallel to p."""
    params = np.asarray(synthetic.TRUTH_PARAMS[model][0], float)
    rng = np.random.default_rng(40 + model)
    n = 0
    for _ in range(200):
        p = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.3, 0.3), rng.uniform(0.5, 1.5), 1.0])
        y, _, _, _, ok = oracle_lib.camera_project(model, params, p)
        if not ok:
            continue
        bp, ok2 = ki.keypoint_to_euclidean(model, np.pad(params, (0, 10 - len(params))), y)
        assert ok2
        d = bp / np.linalg.norm(bp) - p[:3] / np.linalg.norm(p[:3])
        assert np.abs(d).max() < (1e-6 if MODEL_D[model] == 4 and model != 1 else 1e-9)  # radtan: 5 Gauss-Newton steps only
        n += 1
    assert n > 100


def test_equidistant_undistort_is_nan_at_the_centre():
    """EquidistantDistortion.hpp:186-211 divides by thetad = 0: the corner drops out of the PnP (NaN fails the cone test)."""
    params = np.pad(np.asarray(synthetic.TRUTH_PARAMS[1][0], float), (0, 2))
    bp, ok = ki.keypoint_to_euclidean(1, params, (params[2], params[3]))
    assert ok and np.isnan(bp[:2]).all()
    Ps, Ms = ki.pnp_inputs(1, params, [params[2], 100.0], [params[3], 120.0], np.zeros((2, 3)))
    assert len(Ps) == 1


@pytest.mark.parametrize("name", list(CASES))
def test_view_transformations_match_opencv_golden(golden, name):
    cfg, S, kw = CASES[name]
    p = synthetic.make_config(cfg, n_sets=S, **kw)
    T, ok = ki.view_transformations(p)
    assert np.array_equal(ok, golden[name + "/ok"])
    assert pose_err(T, golden[name + "/T_views"]) < TOL
    sp, good = ki.target_pose_guesses(p)
    assert np.array_equal(good, golden[name + "/set_ok"])
    assert pose_err(sp, golden[name + "/set_poses"]) < TOL
    # the PnP poses are good initial guesses: close to the generator's (perturbed) set poses for camera 0's views
    for w in np.flatnonzero(p.view_cam == 0):
        assert np.abs(ki.pose_to_T(T[w]) - ki.pose_to_T(p.set_poses[p.view_set[w]])).max() < 0.1
    if p.n_cams >= 2:
        b = ki.stereo_baseline_guess(p, 0, 1)
        assert pose_err(b, golden[name + "/baseline01"]) < TOL
        assert np.abs(ki.pose_to_T(b) - ki.pose_to_T(p.baselines[0])).max() < 0.1


def test_target_pose_guess_chains_baselines_in_accumulate_order():
    """CalibrationTools.hpp:352-353: std::accumulate multiplies T_t_cN by baseline 0 first, then 1, ... (not the reverse)."""
    p = synthetic.make_config(4, n_sets=2)
    # make camera 3 the one with most corners in set 0 by dropping corners of the other views
    keep = np.ones(p.n_terms, bool)
    for w in np.flatnonzero((p.view_set == 0) & (p.view_cam != 3)):
        keep[p.view_begin[w] + 100:p.view_begin[w + 1]] = False
    from kalibr_b200.problem import Problem

    lens = np.array([keep[p.view_begin[w]:p.view_begin[w + 1]].sum() for w in range(len(p.view_set))])
    q = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, p.target_points, p.view_set, p.view_cam,
                np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), p.y_u[keep], p.y_v[keep], p.corner_id[keep])
    sp, good = ki.target_pose_guesses(q)
    T_views, _ = ki.view_transformations(q)
    w3 = int(np.flatnonzero((q.view_set == 0) & (q.view_cam == 3))[0])
    expect = ki.pose_to_T(T_views[w3]) @ ki.pose_to_T(q.baselines[0]) @ ki.pose_to_T(q.baselines[1]) @ ki.pose_to_T(q.baselines[2])
    assert np.abs(ki.pose_to_T(sp[0]) - expect).max() < 1e-12
    assert good.all()


def test_fewer_than_four_corners_fails():
    p = synthetic.make_config(1, n_sets=1)
    pose, ok = ki.estimate_transformation(p.cam_model[0], p.cam_params[0], p.y_u[:3], p.y_v[:3], p.target_points[p.corner_id[:3]])
    assert not ok and np.array_equal(pose, [0, 0, 0, 1, 0, 0, 0])


# ---- initializeIntrinsics --------------------------------------------------------------------------------------------------
def test_circle_helpers():
    """PinholeHelpers::fitCircle / intersectCircles / medianOfVectorElements (PinholeProjection.hpp:612-707)."""
    a = np.linspace(0.3, 1.4, 12)
    cx, cy, r = ki.fit_circle(np.stack([5.0 + 3.0 * np.cos(a), -2.0 + 3.0 * np.sin(a)], 1))
    assert abs(cx - 5.0) < 1e-9 and abs(cy + 2.0) < 1e-9 and abs(r - 3.0) < 1e-9
    ip = ki.intersect_circles(0, 0, 5, 6, 0, 5)
    assert len(ip) == 2 and np.allclose(sorted(ip), [(3, -4), (3, 4)])
    assert ki.intersect_circles(0, 0, 1, 5, 0, 1) == [] and ki.intersect_circles(0, 0, 5, 1, 0, 1) == []
    assert len(ki.intersect_circles(0, 0, 1, 2, 0, 1)) == 1
    assert ki.median_of_vector_elements([3, 1, 2]) == 2 and ki.median_of_vector_elements([4, 1, 3, 2]) == 2.5
    assert ki.median([4, 1, 3, 2]) == 3  # kalibr2::math::median is the upper median: the two helpers differ


def test_omni_line_image_recovers_the_focal_length():
    """A straight line seen by a unified camera with xi = 1 gives gamma exactly (the camodocal construction)."""
    rng = np.random.default_rng(2)
    gamma, cu, cv = 420.0, 319.5, 239.5
    for _ in range(5):
        p0 = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(0.5, 1.0)])
        d = rng.normal(size=3)
        pts = p0 + np.linspace(-0.4, 0.4, 12)[:, None] * d / np.linalg.norm(d)
        rz = 1.0 / (pts[:, 2] + np.linalg.norm(pts, axis=1))
        u, v = gamma * pts[:, 0] * rz, gamma * pts[:, 1] * rz
        g = ki.omni_row_candidate(u, v)
        assert g is None or abs(g - gamma) < 1e-6 * gamma


@pytest.mark.parametrize("name", list(CASES))
def test_initialize_intrinsics_matches_golden(golden, name):
    cfg, S, kw = CASES[name]
    p = synthetic.make_config(cfg, n_sets=S, **kw)
    for k in range(p.n_cams):
        m = p.cam_model[k]
        res = synthetic.TRUTH_PARAMS[m][1]
        prm, ok = ki.initialize_intrinsics(p, k, 10, 12, res)
        assert ok == bool(golden[name + f"/init_ok{k}"])
        assert np.abs(prm - golden[name + f"/init_params{k}"]).max() <= 1e-9 * max(np.abs(prm).max(), 1.0)
        if ok:
            P = MODEL_P[m]
            assert np.all(prm[P:] == 0.0)                                   # distortion cleared
            assert prm[P - 2] == (res[0] - 1) / 2 and prm[P - 1] == (res[1] - 1) / 2   # image centre
            truth_f = synthetic.TRUTH_PARAMS[m][0][P - 4]
            f_equiv = prm[P - 4] * (2.0 if m in (EUCM, DS) else 1.0)  # EUCM / DS store half the unified-model focal length
            assert 0.5 * truth_f < f_equiv < 4.0 * truth_f                # a usable start, not an estimate


def test_initialize_intrinsics_fallbacks():
    p = synthetic.make_config(3, n_sets=2, dropout=0.7)  # no complete view, hardly a row with more than 4 corners
    res = synthetic.TRUTH_PARAMS
    prm, ok = ki.initialize_intrinsics(p, 3, 10, 12, res[1][1])            # pinhole-equi: no complete view
    assert not ok and np.all(prm == 0)
    prm, ok = ki.initialize_intrinsics(p, 3, 10, 12, res[1][1], fallback=450.0)
    assert ok and prm[0] == prm[1] == 450.0                                # PinholeProjection.hpp:781-791 returns true with the fallback
