"""Host-side logic that needs no GPU: design-variable ordering of the three kalibr2 drivers, sharding, generator."""
import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import ORDER_BATCH, ORDER_RIG, ORDER_SINGLE, ORDER_STEREO


@pytest.mark.parametrize("cfg,n_sets", [(1, 5), (2, 4), (3, 3), (4, 2), (5, 2), (8, 3)])
def test_dv_layout_matches_the_oracle_problem_construction(oracle_lib, cfg, n_sets):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    col, dims, labels = p.dv_layout()
    oc, od = oracle_lib.OracleProblem(p).dv_layout()
    assert np.array_equal(col, oc) and np.array_equal(dims, od)
    assert int(col[-1] + dims[-1]) == p.n_c + 6 * p.n_sets


def test_driver_orders():
    p1 = synthetic.make_config(1, n_sets=2)
    assert [l[0] for l in p1.dv_layout()[2]] == ["proj", "dist", "set_q", "set_t", "set_q", "set_t"]
    p2 = synthetic.make_config(2, n_sets=1)
    assert [l[0] for l in p2.dv_layout()[2]] == ["baseline_q", "baseline_t", "set_q", "set_t", "proj", "dist", "proj", "dist"]
    p3 = synthetic.make_config(3, n_sets=1)
    assert [l[0] for l in p3.dv_layout()[2]][:8] == ["proj", "dist"] * 4
    assert p1.driver_order == ORDER_SINGLE and p2.driver_order == ORDER_STEREO and p3.driver_order == ORDER_RIG
    # the incremental estimator's merged problem (CalibrationTools.hpp:460-491 + IncrementalEstimator.cpp:550-565): poses, baselines, intrinsics
    p8 = synthetic.make_config(8, n_sets=2)
    assert p8.driver_order == ORDER_BATCH
    assert [l[0] for l in p8.dv_layout()[2]] == ["set_q", "set_t"] * 2 + ["baseline_q", "baseline_t"] * 2 + ["proj", "dist"] * 3
    col8, dims8, _ = p8.dv_layout()
    assert int(col8[4]) == 12 and int(col8[-1] + dims8[-1]) - 12 == p8.n_c  # the calibration block is the last n_c columns
    # term order: stereo lists all camera-0 views first, the rig driver interleaves cameras per set
    assert list(p2.view_cam) == [0, 1]
    assert list(synthetic.make_config(2, n_sets=3).view_cam) == [0, 0, 0, 1, 1, 1]
    assert list(synthetic.make_config(3, n_sets=2).view_cam) == [0, 1, 2, 3, 0, 1, 2, 3]


def test_config_sizes_match_the_survey_table():
    for cfg, (n_c, per_set) in {1: (8, 120), 2: (22, 240), 3: (47, 480), 4: (106, 960), 5: (218, 1920)}.items():
        p = synthetic.make_config(cfg, n_sets=2)
        assert p.n_c == n_c and p.n_terms == 2 * per_set
    assert synthetic.CONFIGS[4][2] * 960 == 19_200_000 and synthetic.CONFIGS[5][2] * 1920 == 12_000_000


def test_shard_sets_is_a_partition():
    for n in (0, 1, 7, 20000):
        for r in (1, 2, 3, 8):
            ranges = [synthetic.shard_sets(n, r, k) for k in range(r)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[k][1] == ranges[k + 1][0] for k in range(r - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_generator_is_deterministic_and_valid():
    a = synthetic.make_config(3, n_sets=6)
    b = synthetic.make_config(3, n_sets=6)
    assert np.array_equal(a.y_u, b.y_u) and np.array_equal(a.set_poses, b.set_poses)
    # different set seed, same rig
    c = synthetic.make_config(3, n_sets=6, set_seed=123)
    assert np.array_equal(a.cam_params, c.cam_params) and np.array_equal(a.baselines, c.baselines)
    assert not np.array_equal(a.set_poses, c.set_poses)
    # every observation lies inside its image
    for k, (w, h) in enumerate(a.truth["resolution"]):
        sel = np.repeat(a.view_cam, np.diff(a.view_begin)) == k
        assert a.y_u[sel].min() > 0 and a.y_u[sel].max() < w and a.y_v[sel].min() > 0 and a.y_v[sel].max() < h
    assert np.allclose(np.linalg.norm(a.set_poses[:, :4], axis=1), 1.0)


def test_aprilgrid_geometry():
    pts = synthetic.aprilgrid_points()
    assert pts.shape == (120, 3) and np.all(pts[:, 2] == 0)
    assert np.isclose(pts[1, 0], 0.088) and np.isclose(pts[2, 0], 0.088 * 1.2954) and np.isclose(pts[12, 1], 0.088)
