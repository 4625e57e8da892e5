"""The C++ mirror of kalibr2's drivers (CalibrateSingleCamera, CalibrateStereoPair, CalibrateMultiCameraRig over the C ABI) on the
GPU: each driver's result is compared with the same flow done step by step through the Python binding, and the optimisation
inside with the CPU oracle started from the same initial guesses (same iteration count, cost within 1e-9, parameters within 1e-6).
  K2/include/kalibr2/CalibrationTools.hpp:93-152, 183-300, 316-428
"""
import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions, Problem
from oracle import ko_init as ki

from driver_util import run_adapter, run_driver, write_problem

pytestmark = pytest.mark.gpu
IDENT = np.array([0, 0, 0, 1.0, 0, 0, 0])


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()
    return m


def with_state(p, cam_params=None, baselines=None, set_poses=None):
    return Problem(p.driver_order, p.cam_model, p.cam_params if cam_params is None else cam_params, p.baselines if baselines is None else baselines,
                   p.set_poses if set_poses is None else set_poses, p.target_points, p.view_set, p.view_cam, p.view_begin, p.y_u, p.y_v, p.corner_id)


def check_against_oracle(oracle_lib, start, sol, cams, baselines=None):
    o = oracle_lib.OracleProblem(start)
    os_, _ = o.optimize(KbOptimizerOptions.kalibr2_default())
    assert int(sol[0]) == os_.iterations and int(sol[1]) == os_.failed_iterations
    assert abs(sol[4] - os_.j_final) <= 1e-9 * os_.j_final
    oc = o.camera_params()
    assert (np.abs(cams - oc) / np.maximum(np.abs(oc), 1e-3)).max() < 1e-6
    if baselines is not None:
        assert np.abs(baselines - o.baselines()).max() < 1e-6


@pytest.mark.parametrize("model", [1, 6, 0])
def test_calibrate_single_camera(capi, oracle_lib, tmp_path, model):
    p = synthetic.make_problem([model], 24, 0, seed=300 + model)
    res = [synthetic.TRUTH_PARAMS[model][1]]
    path = str(tmp_path / "single.bin")
    write_problem(path, with_state(p, cam_params=np.ones((1, 10))), res)
    code, out = run_driver("single", path)
    assert code == 0 and out["ok"][0] == 1
    # the same flow through the Python binding
    g = capi.B200SchurLinearSystemSolver(with_state(p, cam_params=np.ones((1, 10)), set_poses=np.tile(IDENT, (p.n_sets, 1))))
    prm, ok = g.initialize_intrinsics(0, 10, 12, res)
    T, okv = g.estimate_transformations(res)
    assert ok and okv.all()
    start = with_state(p, cam_params=prm[None], set_poses=T)
    check_against_oracle(oracle_lib, start, out["solution"], out["camera0"][None])
    g2 = capi.B200SchurLinearSystemSolver(start)
    g2.optimize()
    assert np.array_equal(g2.camera_params()[0], out["camera0"])
    assert np.abs(g2.reprojection_statistics()[0] - out["stats"]).max() < 1e-12
    assert out["stats"][0] == p.n_terms and 0.2 < out["stats"][3] < 0.45  # the generator's 0.3 px noise


def test_calibrate_stereo_pair(capi, oracle_lib, tmp_path):
    p = synthetic.make_config(2, n_sets=14)
    res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
    path = str(tmp_path / "stereo.bin")
    write_problem(path, p, res)
    code, out = run_driver("stereo", path)
    assert code == 0
    g = capi.B200SchurLinearSystemSolver(with_state(p, baselines=IDENT[None], set_poses=np.tile(IDENT, (p.n_sets, 1))))
    b, n = g.estimate_stereo_baseline(0, 1, res)
    T, okv = g.estimate_transformations(res)
    assert n == p.n_sets and okv.all()
    poses = T[:p.n_sets]  # stereo order: camera L's views come first, one per set
    start = with_state(p, baselines=b[None], set_poses=poses)
    check_against_oracle(oracle_lib, start, out["solution"], np.stack([out["camera0"], out["camera1"]]), out["baseline0"][None])
    assert np.abs(ki.pose_to_T(out["baseline0"]) - ki.pose_to_T(p.truth["baselines"][0])).max() < 5e-3  # and it finds the rig


def test_calibrate_stereo_pair_with_missing_images(capi, oracle_lib, tmp_path):
    """Sets seen by one camera only: L missing -> the pose comes from H chained through the baseline guess (:252-255); a set nobody
    saw gets no pose design variable."""
    p = synthetic.make_config(2, n_sets=10)
    res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
    drop = ((p.view_set == 2) & (p.view_cam == 0)) | ((p.view_set == 5) & (p.view_cam == 1)) | (p.view_set == 7)
    keep_v = ~drop
    keep_t = np.repeat(keep_v, np.diff(p.view_begin))
    vb = np.concatenate([[0], np.cumsum(np.diff(p.view_begin)[keep_v])]).astype(np.int64)
    q = Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses, p.target_points, p.view_set[keep_v], p.view_cam[keep_v], vb,
                p.y_u[keep_t], p.y_v[keep_t], p.corner_id[keep_t])
    path = str(tmp_path / "stereo_missing.bin")
    write_problem(path, q, res)
    code, out = run_driver("stereo", path)
    assert code == 0
    # expected start, from the oracle's PnP: sets renumbered without set 7
    remap = {s: i for i, s in enumerate([s for s in range(10) if s != 7])}
    vs = np.array([remap[s] for s in q.view_set], np.int32)
    q9 = Problem(q.driver_order, q.cam_model, q.cam_params, q.baselines, np.tile(IDENT, (9, 1)), q.target_points, vs, q.view_cam, q.view_begin, q.y_u, q.y_v, q.corner_id)
    Tv, okv = ki.view_transformations(q9)
    assert okv.all()
    b = ki.stereo_baseline_guess(q9, 0, 1)
    poses = np.zeros((9, 7))
    for s in range(9):
        wl = np.flatnonzero((vs == s) & (q.view_cam == 0))
        wh = np.flatnonzero((vs == s) & (q.view_cam == 1))
        T = ki.pose_to_T(Tv[wl[0]]) if len(wl) else ki.pose_to_T(Tv[wh[0]]) @ np.linalg.inv(ki.pose_to_T(b))
        poses[s] = ki.T_to_pose(T)
    start = Problem(q.driver_order, q.cam_model, q.cam_params, b[None], poses, q.target_points, vs, q.view_cam, q.view_begin, q.y_u, q.y_v, q.corner_id)
    check_against_oracle(oracle_lib, start, out["solution"], np.stack([out["camera0"], out["camera1"]]), out["baseline0"][None])


def test_calibrate_multi_camera_rig(capi, oracle_lib, tmp_path):
    p = synthetic.make_config(3, n_sets=10)
    res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
    path = str(tmp_path / "rig.bin")
    write_problem(path, p, res)
    code, out = run_driver("rig", path)
    assert code == 0
    g = capi.B200SchurLinearSystemSolver(with_state(p, set_poses=np.tile(IDENT, (p.n_sets, 1))))
    assert g.initialize_set_poses(res) == 0
    start = with_state(p, set_poses=g.set_poses())
    cams = np.stack([out[f"camera{k}"] for k in range(p.n_cams)])
    base = np.stack([out[f"baseline{j}"] for j in range(p.n_cams - 1)])
    check_against_oracle(oracle_lib, start, out["solution"], cams, base)


@pytest.mark.parametrize("cfg,n_sets,delta", [(2, 10, 0.2), (2, 12, 2.5), (3, 8, 1.0)])
def test_incremental_estimator_matches_oracle(oracle_lib, tmp_path, cfg, n_sets, delta):
    """IncrementalEstimator::addBatch over a sequence of synced sets (kalibr2_ros CalibrateCameras.cpp:279-304): the same accept /
    reject decisions, information gains, ranks, iteration counts and costs as the dense numpy restatement, and the same calibration."""
    from oracle import ko_estimator as ke

    p = synthetic.make_config(cfg, n_sets=n_sets)
    res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
    path = str(tmp_path / "estimator.bin")
    write_problem(path, p, res)
    code, out = run_driver("estimator", path, delta, "check")
    assert code == 0, out
    est = ke.OracleIncrementalEstimator(oracle_lib, p.cam_model, p.cam_params, p.baselines, p.target_points, info_gain_delta=delta, check_validity=True)
    decisions = []
    for s in range(p.n_sets):
        batch = {}
        for w in np.flatnonzero(p.view_set == s):
            b, e = p.view_begin[w], p.view_begin[w + 1]
            batch[int(p.view_cam[w])] = (p.corner_id[b:e], p.y_u[b:e], p.y_v[b:e])
        r = est.add_batch(batch, p.set_poses[s])
        acc, gain, rank, iters, j0, j1 = out[f"batch{s}"]
        # the sequences are chosen (on the CPU, with the oracle) away from the rank tolerance and the acceptance threshold: no skip
        assert r["rank_margin"] >= 1e-3 and abs(r["information_gain"] - delta) >= 1e-3, "pick another sequence instead of skipping"
        assert bool(acc) == r["batch_accepted"], (s, out[f"batch{s}"], r)
        assert int(rank) == r["rank_theta"] and int(iters) == r["num_iterations"]
        assert abs(gain - r["information_gain"]) <= 1e-6 * max(1.0, abs(r["information_gain"]))
        assert abs(j1 - r["j_final"]) <= 1e-9 * r["j_final"] and abs(j0 - r["j_start"]) <= 1e-9 * r["j_start"]
        decisions.append(bool(acc))
        # ReturnValue's bases and covariance (IncrementalEstimator.cpp:395-417): dimensions follow the rank, [obs | nobs] is orthonormal
        n, n_obs, n_nobs, orth, cov, n_scaled, n_obs_scaled = out[f"spaces{s}"]
        assert int(n) == p.n_c and int(n_obs) == r["rank_theta"] and int(n_obs) + int(n_nobs) == p.n_c
        assert orth < 1e-10 and cov < 1e-10
        assert int(n_scaled) == p.n_c and int(n_obs_scaled) == r["rank_theta"]  # column scaling is on: the scaled system's spaces are reported too
    assert int(out["accepted"][0]) == len(est.batches) == sum(decisions)
    if delta > 1.0:
        assert not all(decisions)  # the threshold really rejects something in this sequence
    cams = np.stack([out[f"camera{k}"] for k in range(p.n_cams)])
    assert (np.abs(cams - est.cam_params) / np.maximum(np.abs(est.cam_params), 1e-3)).max() < 1e-6
    base = np.stack([out[f"baseline{j}"] for j in range(p.n_cams - 1)])
    assert np.abs(base - est.baselines).max() < 1e-6


# ---- the reference-side adapter under a host optimiser that owns the design variables ------------------------------------------
# include/kalibr_b200/reference_adapter.hpp compiled against the stand-in of aslam::backend::{LinearSystemSolver, DesignVariable,
# ErrorTerm} (exact virtual signatures: BE/include/aslam/backend/LinearSystemSolver.hpp:16-109), driven by a host loop in the role
# of Optimizer2 (state updates / reverts on the HOST, Optimizer2.cpp:290-318).  It must reproduce kb_optimize - same iterations,
# failed iterations, cost and parameters - in both integration variants of INTEGRATION.md §2.
@pytest.mark.parametrize("virtual_evaluate", [False, True])
@pytest.mark.parametrize("cfg,n_sets", [(1, 60), (2, 40), (3, 30), (8, 24)])
def test_reference_adapter_reproduces_kb_optimize(capi, oracle_lib, tmp_path, cfg, n_sets, virtual_evaluate):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    res = [synthetic.TRUTH_PARAMS[int(m)][1] for m in p.cam_model]
    path = str(tmp_path / "adapter.bin")
    write_problem(path, p, res)
    code, out = run_adapter(path, p.driver_order, virtual_evaluate)
    assert code == 0, out.get("error")
    g = capi.B200SchurLinearSystemSolver(p)
    sol, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
    assert int(out["solution"][0]) == sol.iterations and int(out["solution"][1]) == sol.failed_iterations
    assert int(out["solution"][2]) == sol.linear_solver_failure
    assert abs(out["solution"][3] - sol.j_start) <= 1e-12 * sol.j_start
    assert abs(out["solution"][4] - sol.j_final) <= 1e-9 * sol.j_final
    gc = g.camera_params()
    for k in range(p.n_cams):
        n = len(out[f"camera{k}"])
        assert (np.abs(out[f"camera{k}"] - gc[k, :n]) / np.maximum(np.abs(gc[k, :n]), 1e-3)).max() < 1e-6
    gb = g.baselines()
    for j in range(p.n_cams - 1):
        assert np.abs(out[f"baseline{j}"] - gb[j]).max() < 1e-6
    # and the oracle agrees on the iteration count (the chain reference -> oracle -> device -> adapter)
    os_, _ = oracle_lib.OracleProblem(p).optimize(KbOptimizerOptions.kalibr2_default())
    assert int(out["solution"][0]) == os_.iterations and int(out["solution"][1]) == os_.failed_iterations
    assert out["launches"][0] > 0


def test_reference_adapter_rejects_a_wrong_design_variable_order(tmp_path):
    p = synthetic.make_config(3, n_sets=6)
    res = [synthetic.TRUTH_PARAMS[int(m)][1] for m in p.cam_model]
    path = str(tmp_path / "adapter.bin")
    write_problem(path, p, res)
    code, out = run_adapter(path, 3, False)  # the host optimiser orders the variables as CreateBatchProblem does, the recorder says so too: fine
    assert code == 0
    # stereo order announced for a four-camera problem: kb_create refuses
    code, out = run_adapter(path, 1, False)
    assert code == 1 and "stereo" in out["error"]
