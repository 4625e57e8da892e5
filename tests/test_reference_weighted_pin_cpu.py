"""The oracle's term weighting (rows a2 / a19 of SURVEY.md §8: invR, M-estimator policies) inside the solvers and the optimiser loop against
the REFERENCE's own code.

tests/golden/reference_weighted_golden.npz (generator: tests/golden/make_reference_weighted_golden.py) holds what the reference's
ErrorTermFs<2>::setInvR / ErrorTerm::setMEstimatorPolicy / getWeightedJacobians / getWeightedError / evaluateError, its M-estimator policies,
both its solvers and Optimizer2 return for six small problems with gross outliers (oracle/ref_pin_optimizer.cpp: ref_set_weighting): the
weighted compressed-column J^T, weighted e, rhs, the policy-weighted cost, a damped step, whole optimisations over the BlockCholesky and the
SparseCholesky solver.  Stand-ins in that build: the CHOLMOD factorisation, Eigen::LDLT behind the matrix square root of invR (exact for the
invR = c I Kalibr2 passes; case "general_huber" uses a general matrix and says so), Boost.Math's chi-squared quantile (case "blake").
The GPU path is held to the same fixture in tests/test_zzz_a_reference_weighted_pin_gpu.py."""
import os

import numpy as np
import pytest

from oracle import oracle_api as oa
from test_reference_pin_cpu import check_against_reference_optimizer

WEIGHTED_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_weighted_golden.npz")
TAGS = ["iso", "huber", "cauchy_iso", "geman", "blake", "general_huber"]
SOLVERS = [("block", 0), ("sparse", 1)]  # (fixture key, oracle solver kind = kb solver semantic)


def weighted_problem(g, tag):
    from kalibr_b200.problem import Problem

    f = lambda n: g[f"{tag}_{n}"]  # noqa: E731
    p = Problem(driver_order=int(f("order")), cam_model=f("cam_model"), cam_params=f("cam_params"), baselines=f("baselines"), set_poses=f("set_poses"),
                target_points=f("target_points"), view_set=f("view_set"), view_cam=f("view_cam"), view_begin=f("view_begin"), y_u=f("y_u"), y_v=f("y_v"),
                corner_id=f("corner_id"))
    pol = f("policy")
    return p, f("inv_r"), (int(pol[0]), float(pol[1]), float(pol[2]), float(pol[3]))


def check_weighted_system(g, tag, system, value_rtol, dx_rtol):
    """system(problem, invR, policy, lambda) -> (cost, e, col_ptr, row_idx, values, rhs, dx, ok): the reference's weighted system"""
    f = lambda n: g[f"{tag}_{n}"]  # noqa: E731
    p, inv_r, policy = weighted_problem(g, tag)
    cost, e, col_ptr, row_idx, values, rhs, dx, ok = system(p, inv_r, policy, float(g["lambda"]))
    assert abs(cost - float(f("cost"))) <= 1e-11 * float(f("cost"))
    assert np.abs(e - f("e")).max() <= value_rtol * np.abs(f("e")).max()
    assert np.array_equal(col_ptr, f("col_ptr")) and np.array_equal(row_idx, f("row_ind"))
    ref = f("values")
    scale = np.ones_like(ref)
    for c in range(len(col_ptr) - 1):
        scale[col_ptr[c]:col_ptr[c + 1]] = max(np.abs(ref[col_ptr[c]:col_ptr[c + 1]]).max(initial=0.0), 1.0)
    assert (np.abs(values - ref) / scale).max() <= value_rtol
    assert np.abs(rhs - f("rhs")).max() <= value_rtol * np.abs(f("rhs")).max()
    assert ok and np.abs(dx - f("dx")).max() <= dx_rtol * np.abs(f("dx")).max()


def optimizer_view(g, tag, solver):
    """the fixture of one (case, solver) under the keys check_against_reference_optimizer reads (problem 0)"""
    v = {f"opt0_{k}": g[f"{tag}_{k}"] for k in ("order", "cam_model", "cam_params", "baselines", "set_poses", "target_points", "view_set", "view_cam",
                                                "view_begin", "y_u", "y_v", "corner_id")}
    v["opt0_lambda_init"] = np.array(10.0)
    for k in ("result", "final_cam_params", "final_baselines", "final_set_poses", "truncated"):
        v[f"opt0_{k}"] = g[f"{tag}_{solver}_{k}"]
    return v


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_weighted_system_reproduces_the_reference(oracle_lib, tag):
    def system(p, inv_r, policy, lam):
        o = oa.OracleProblem(p, oa.SPARSE_CHOLESKY, n_threads=1)
        o.set_inv_r(inv_r)
        o.set_m_estimator(*policy)
        cost = o.evaluate_error()
        col_ptr, row_idx, values = o.jacobian_ccs()
        o.build_system()
        o.set_constant_conditioner(lam)
        dx, ok = o.solve_system()
        return cost, o.error_vector(), col_ptr, row_idx, values, o.rhs(), dx, ok

    check_weighted_system(np.load(WEIGHTED_GOLD), tag, system, 1e-12, 1e-9)


@pytest.mark.parametrize("solver,kind", SOLVERS)
@pytest.mark.parametrize("tag", TAGS)
def test_oracle_weighted_optimizer_walks_the_reference_optimizer(oracle_lib, tag, solver, kind):
    g = np.load(WEIGHTED_GOLD)
    _, inv_r, policy = weighted_problem(g, tag)

    def solve(p, opt):
        o = oa.OracleProblem(p, kind, n_threads=1)
        o.set_inv_r(inv_r)
        o.set_m_estimator(*policy)
        sol, _ = o.optimize(opt)
        return sol, o.camera_params(), o.baselines(), o.set_poses()

    check_against_reference_optimizer(optimizer_view(g, tag, solver), 0, solve)


def test_weighted_fixture_is_what_the_reference_returns_now(oracle_lib):
    """build container only: the reference's compiled weighting, run again, returns the committed numbers"""
    if oa.build_reference_cameras() is None:
        pytest.skip("no reference tree and no prebuilt oracle/_ref here")
    from kalibr_b200.problem import KbOptimizerOptions

    g = np.load(WEIGHTED_GOLD)
    try:
        for tag in TAGS:
            p, inv_r, policy = weighted_problem(g, tag)
            oa.reference_set_weighting(inv_r, policy)
            r = oa.reference_sparse_system(p, float(g["lambda"]), 1)
            for k in ("col_ptr", "row_ind", "values", "e", "rhs", "dx"):
                assert np.array_equal(r[k], g[f"{tag}_{k}"]), (tag, k)
            res = oa.reference_optimize(p, KbOptimizerOptions.kalibr2_default(), oa.SPARSE_CHOLESKY_KIND, 1)[0]
            assert [res["iterations"], res["failed_iterations"], res["j_start"], res["j_final"], res["linear_solver_failure"]] == list(g[f"{tag}_sparse_result"])
    finally:
        oa.reference_set_weighting()
