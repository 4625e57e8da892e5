"""CPU tests of the oracle's term weighting (inverse measurement covariance + M-estimator policies) and of the
reprojection statistics.

  BE/src/MEstimatorPolicies.cpp:16-125                     the policies' weights; Blake-Zisserman's epsilon from the chi-squared quantile
  BE/include/aslam/backend/implementation/ErrorTerm.hpp:97-109, 118-127, 170-192   setInvR, weighted error / Jacobians / Hessian
  BE/src/ErrorTerm.cpp:19-24                               the cost is always policy-weighted
  Schweizer-Messer/sm_eigen/include/sm/eigen/matrix_sqrt.hpp:21-40   S = P^T L sqrt(D)
  K2/include/kalibr2/CameraCalibrator.hpp:368-405          PrintReprojectionErrorStatistics
"""
import numpy as np
import pytest

from kalibr_b200 import synthetic

from test_oracle_cpu import _dense_from_blocks, _dense_from_ccs

NONE, HUBER, CAUCHY, GEMAN, BLAKE = range(5)
POLICIES = [(HUBER, 1.2, 0.0, 0.0), (CAUCHY, 0.8, 0.0, 0.0), (GEMAN, 2.0, 0.0, 0.0), (BLAKE, 2, 0.999, 0.1)]
INV_R = np.array([[3.0, 0.4], [0.4, 5.0]])  # second diagonal entry larger: the pivoted LDL^T swaps


def numpy_weight(kind, s, p0, p1=0.999, p2=0.1):
    from scipy.stats import chi2

    s = np.asarray(s, float)
    if kind == HUBER:
        return np.where(s < p0 * p0, 1.0, p0 / np.sqrt(np.maximum(s, 1e-300)))
    if kind == CAUCHY:
        return 1.0 / (1.0 + s / p0)
    if kind == GEMAN:
        return p0 / (p0 + s) ** 2
    if kind == BLAKE:
        eps = (1 - p2) / p2 * np.exp(-chi2.ppf(p1, int(p0)))
        return np.exp(-s) / (np.exp(-s) + eps)
    return np.ones_like(s)


def test_chi2_quantile_matches_scipy(oracle_lib):
    from scipy.stats import chi2

    for df in (1, 2, 3, 4, 6, 10, 30):
        for p in (0.01, 0.5, 0.9, 0.99, 0.999, 0.99999):
            assert abs(oracle_lib.chi2_inv_cdf(p, df) - chi2.ppf(p, df)) <= 1e-10 * chi2.ppf(p, df), (df, p)
    # df = 2 has the closed form -2 ln(1 - p)
    assert abs(oracle_lib.chi2_inv_cdf(0.999, 2) + 2 * np.log(1e-3)) < 1e-11


@pytest.mark.parametrize("kind,p0,p1,p2", POLICIES)
def test_policy_weights(oracle_lib, kind, p0, p1, p2):
    for s in (0.0, 1e-6, 0.3, 1.0, 1.44, 1.45, 3.0, 25.0, 400.0):
        w = oracle_lib.m_estimator_weight(kind, s, p0, p1, p2)
        assert abs(w - numpy_weight(kind, s, p0, p1, p2)) <= 1e-13 * max(1.0, w), (kind, s)
    assert oracle_lib.m_estimator_weight(NONE, 7.0) == 1.0


def test_matrix_sqrt_follows_the_pivoted_ldlt(oracle_lib):
    assert np.array_equal(oracle_lib.matrix_sqrt2(np.eye(2)), np.eye(2))
    # no swap (first diagonal entry is the larger one, or a tie): lower triangular
    for A in (np.array([[5.0, 0.4], [0.4, 3.0]]), np.array([[2.0, 0.5], [0.5, 2.0]]), np.eye(2) / 0.09):
        S = oracle_lib.matrix_sqrt2(A)
        assert S[0, 1] == 0.0 and np.abs(S @ S.T - A).max() < 1e-15 * np.abs(A).max()
        assert np.allclose(S, np.linalg.cholesky(A), rtol=1e-15)
    # swap: S = P^T L sqrt(D) is the row-swapped factor of the permuted matrix
    S = oracle_lib.matrix_sqrt2(INV_R)
    assert S[1, 1] == 0.0 and np.abs(S @ S.T - INV_R).max() < 1e-15
    Lp = np.linalg.cholesky(INV_R[::-1, ::-1])
    assert np.allclose(S, Lp[::-1, :], rtol=1e-15)


def _raw_errors(oracle_lib, p):
    o = oracle_lib.OracleProblem(p, n_threads=2)
    o.evaluate_error()
    return -o.error_vector().reshape(-1, 2)  # y - y_hat


@pytest.mark.parametrize("kind,p0,p1,p2", [(NONE, 0.0, 0.0, 0.0)] + POLICIES)
@pytest.mark.parametrize("cfg,n_sets", [(1, 4), (3, 2)])
def test_weighted_cost_error_and_normal_equations(oracle_lib, cfg, n_sets, kind, p0, p1, p2):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    e_raw = _raw_errors(oracle_lib, p)
    S = oracle_lib.matrix_sqrt2(INV_R)
    raw = np.einsum("ni,ij,nj->n", e_raw, INV_R, e_raw)
    w = numpy_weight(kind, raw, p0, p1, p2)

    o = oracle_lib.OracleProblem(p, n_threads=2)
    o.set_inv_r(INV_R)
    o.set_m_estimator(kind, p0, p1, p2)
    J = o.evaluate_error()
    assert abs(J - np.sum(w * raw)) <= 1e-12 * J  # BE/src/ErrorTerm.cpp:19-24
    e_w = -o.error_vector().reshape(-1, 2)
    assert np.abs(e_w - np.sqrt(w)[:, None] * (e_raw @ S)).max() <= 1e-12 * np.abs(e_w).max()  # S^T e per term

    # H == J^T J and rhs == -J^T e with the weighted J and e (BE/test/TestOptimizer.cpp:101-120)
    cp, ri, jv = o.jacobian_ccs()
    Jw = _dense_from_ccs(cp, ri, jv, o.jcols)
    o.build_system()
    col, dims = o.dv_layout()
    H = _dense_from_blocks(*o.hessian_blocks(), col, dims)
    assert np.abs(H - Jw.T @ Jw).max() <= 1e-12 * np.abs(H).max()
    assert np.abs(o.rhs() + Jw.T @ e_w.ravel()).max() <= 1e-12 * np.abs(o.rhs()).max()

    # the weighted Jacobian is sqrt(w) S^T times the unweighted one, term by term
    o0 = oracle_lib.OracleProblem(p, n_threads=2)
    o0.evaluate_error()
    J0 = _dense_from_ccs(*o0.jacobian_ccs(), o0.jcols).reshape(-1, 2, o0.jcols)
    expect = np.sqrt(w)[:, None, None] * np.einsum("ji,njc->nic", S, J0)
    assert np.abs(Jw.reshape(-1, 2, o.jcols) - expect).max() <= 1e-12 * np.abs(expect).max()

    # useMEstimator = false: the cost keeps the policy weight, e() and J lose it (ErrorTerm.hpp:170-192)
    o.set_use_m_estimator(False)
    assert abs(o.evaluate_error() - J) <= 1e-15 * J
    e_n = -o.error_vector().reshape(-1, 2)
    assert np.abs(e_n - e_raw @ S).max() <= 1e-12 * np.abs(e_n).max()


def test_weighted_lm_rejects_outliers(oracle_lib):
    """A Cauchy policy makes the calibration robust against gross outliers (what the policies are for)."""
    from kalibr_b200.problem import KbOptimizerOptions

    p = synthetic.make_config(1, n_sets=25)
    rng = np.random.default_rng(5)
    bad = rng.choice(p.n_terms, p.n_terms // 25, replace=False)
    p.y_u[bad] += rng.normal(0, 40.0, bad.size)
    p.y_v[bad] += rng.normal(0, 40.0, bad.size)
    truth = np.asarray(synthetic.TRUTH_PARAMS[p.cam_model[0]][0], float)

    def run(kind):
        o = oracle_lib.OracleProblem(p, n_threads=4)
        if kind:
            o.set_m_estimator(kind, 1.0)
        o.optimize(KbOptimizerOptions.kalibr2_default())
        return np.abs(o.camera_params()[0, :4] / truth[:4] - 1.0).max()

    plain, robust = run(NONE), run(CAUCHY)
    assert robust < 5e-3 and robust < 0.25 * plain, (plain, robust)


@pytest.mark.parametrize("cfg,n_sets", [(1, 6), (3, 3), (2, 1)])
def test_reprojection_statistics(oracle_lib, cfg, n_sets):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    e = _raw_errors(oracle_lib, p)
    o = oracle_lib.OracleProblem(p)
    o.set_inv_r(INV_R)  # statistics use the raw errors whatever the weighting
    st = o.reprojection_statistics()
    cam = np.repeat(p.view_cam, np.diff(p.view_begin))
    for k in range(p.n_cams):
        ek = e[cam == k]
        n = len(ek)
        assert st[k, 0] == n
        assert np.abs(st[k, 1:3] - ek.mean(0)).max() < 1e-13
        assert np.abs(st[k, 3:5] - ek.std(0, ddof=1)).max() < 1e-12
        assert abs(st[k, 5] - np.linalg.norm(ek.sum(0)) / np.sqrt(n)) < 1e-12  # "RMSE" as the reference prints it
