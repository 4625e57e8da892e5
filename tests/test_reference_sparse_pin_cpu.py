"""The oracle's SparseCholesky regime (Kalibr2's DEFAULT solver: BE/src/Optimizer2.cpp:83-86) against the REFERENCE's own code - row a26 of
SURVEY.md §8.

tests/golden/reference_sparse_golden.npz (generator: tests/golden/make_reference_sparse_golden.py) holds what the reference's
SparseCholeskyLinearSystemSolver.cpp, CompressedColumnJacobianTransposeBuilder, CompressedColumnMatrix and Cholmod wrapper return, compiled
from their sources by oracle/ref_pin_optimizer.cpp together with Optimizer2, the LM policy, ErrorTerm, JacobianContainer, the expression tree
and the camera models (stand-in headers: oracle/ref_shim/): J^T in compressed-column form - column pointers, row indices and values as the
reference's builder lays them out -, the error vector, rhs = J^T e, dx of one damped solve, and whole optimisations over that solver on the
eleven problems of the BlockCholesky fixture.  Only the factorisation behind the reference's Cholmod wrapper is a stand-in
(oracle/ref_shim/cholmod.h: dense Cholesky of A A^T; SuiteSparse is not in the image), i.e. rounding of dx.  The GPU path is held to the
same fixture in tests/test_zz_reference_sparse_pin_gpu.py."""
import os

import numpy as np
import pytest

from oracle import oracle_api as oa
from test_reference_pin_cpu import GOLD, N_OPT, check_against_reference_optimizer, opt_problem

SPARSE_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sparse_golden.npz")
SYSTEM_TAGS = ["rig", "batch", "stereo", "single"]
# cost per iteration: 1e-9, except problem 3 (pinhole-equi + double sphere started at lambda = 1e-4): its first systems have
# cond(J^T J + lambda^2 I) = 5.6e12, so a double-precision factorisation moves dx by ~1e-4 of its size and the cost after the step by
# 1e-9 of its value (measured: oracle vs the extended-precision stand-in 1.0e-9, vs a double-precision stand-in 1.8e-9); the counts,
# accept / reject decisions and final parameters are held as everywhere else
COST_RTOL = {3: 2e-8}


def system_problem(g, tag):
    from kalibr_b200.problem import Problem

    f = lambda n: g[f"sys_{tag}_{n}"]  # noqa: E731
    return Problem(driver_order=int(f("order")), cam_model=f("cam_model"), cam_params=f("cam_params"), baselines=f("baselines"), set_poses=f("set_poses"),
                   target_points=f("target_points"), view_set=f("view_set"), view_cam=f("view_cam"), view_begin=f("view_begin"), y_u=f("y_u"), y_v=f("y_v"),
                   corner_id=f("corner_id"))


def check_sparse_system(g, tag, system, value_rtol, dx_rtol):
    """system(problem, lambda) -> (cost, e, col_ptr, row_idx, values, rhs, dx, ok) must be the reference's: the compressed-column layout of
    J^T bit for bit, its values, e, rhs = J^T e and the damped step"""
    f = lambda n: g[f"sys_{tag}_{n}"]  # noqa: E731
    cost, e, col_ptr, row_idx, values, rhs, dx, ok = system(system_problem(g, tag), float(g["sys_lambda"]))
    assert abs(cost - float(f("cost"))) <= 1e-11 * float(f("cost"))
    assert np.abs(e - f("e")).max() <= value_rtol * np.abs(f("e")).max()
    assert np.array_equal(col_ptr, f("col_ptr")) and np.array_equal(row_idx, f("row_ind"))  # the sparsity / index pattern: bit-exact
    ref = f("values")
    scale = np.ones_like(ref)
    for c in range(len(col_ptr) - 1):  # relative to the largest entry of the column (one residual row of J)
        scale[col_ptr[c]:col_ptr[c + 1]] = max(np.abs(ref[col_ptr[c]:col_ptr[c + 1]]).max(initial=0.0), 1.0)
    assert (np.abs(values - ref) / scale).max() <= value_rtol
    assert np.abs(rhs - f("rhs")).max() <= value_rtol * np.abs(f("rhs")).max()
    assert ok
    assert np.abs(dx - f("dx")).max() <= dx_rtol * np.abs(f("dx")).max()


def sparse_golden():
    """the inputs of the optimiser problems live in reference_golden.npz; the SparseCholesky results overwrite the BlockCholesky ones"""
    g = dict(np.load(GOLD))
    g.update(dict(np.load(SPARSE_GOLD)))
    return g


@pytest.mark.parametrize("tag", SYSTEM_TAGS)
def test_oracle_sparse_system_reproduces_the_reference(oracle_lib, tag):
    def system(p, lam):
        o = oa.OracleProblem(p, oa.SPARSE_CHOLESKY, n_threads=1)
        cost = o.evaluate_error()
        col_ptr, row_idx, values = o.jacobian_ccs()
        o.build_system()
        o.set_constant_conditioner(lam)
        dx, ok = o.solve_system()
        return cost, o.error_vector(), col_ptr, row_idx, values, o.rhs(), dx, ok

    check_sparse_system(np.load(SPARSE_GOLD), tag, system, 1e-13, 1e-9)


@pytest.mark.parametrize("n", range(N_OPT))
def test_oracle_optimizer_walks_the_reference_sparse_optimizer(oracle_lib, n):
    """Optimizer2::optimize with the LM policy over SparseCholeskyLinearSystemSolver (damping appended as columns of J^T: no Q2 residual, so
    the runs with rejected steps take other paths than the BlockCholesky ones) - the reference's compiled loop against the oracle's"""
    g = sparse_golden()
    assert int(g["opt_count"]) == N_OPT

    def solve(p, opt):
        o = oa.OracleProblem(p, oa.SPARSE_CHOLESKY, n_threads=1)
        sol, _ = o.optimize(opt)
        return sol, o.camera_params(), o.baselines(), o.set_poses()

    check_against_reference_optimizer(g, n, solve, COST_RTOL.get(n, 1e-9))
    assert sum(int(g[f"opt{i}_result"][1]) > 0 for i in range(N_OPT)) >= 4
    b = np.load(GOLD)  # the two regimes really differ on this fixture
    assert any(tuple(b[f"opt{i}_result"][:2]) != tuple(g[f"opt{i}_result"][:2]) for i in range(N_OPT))


def test_sparse_fixture_is_what_the_reference_returns_now(oracle_lib):
    """build container only: the reference's compiled SparseCholesky regime, run again (serially and on four threads), returns the committed numbers"""
    if oa.build_reference_cameras() is None:
        pytest.skip("no reference tree and no prebuilt oracle/_ref here")
    g = sparse_golden()
    for tag in SYSTEM_TAGS:
        p = system_problem(g, tag)
        for threads in (1, 4):
            r = oa.reference_sparse_system(p, float(g["sys_lambda"]), threads)
            for k in ("col_ptr", "row_ind", "values", "e", "rhs", "dx"):
                assert np.array_equal(r[k], g[f"sys_{tag}_{k}"]), (tag, threads, k)
    for n in (0, 3, 8, 10):
        p, opt = opt_problem(g, n)
        r, cp, bl, sp = oa.reference_optimize(p, opt, oa.SPARSE_CHOLESKY_KIND, 1)
        assert [r["iterations"], r["failed_iterations"], r["j_start"], r["j_final"], r["linear_solver_failure"]] == list(g[f"opt{n}_result"])
        assert np.array_equal(cp, g[f"opt{n}_final_cam_params"]) and np.array_equal(sp, g[f"opt{n}_final_set_poses"])
