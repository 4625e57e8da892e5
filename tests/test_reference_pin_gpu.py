"""The CUDA path against values computed by the REFERENCE's own code (tests/golden/reference_golden.npz, see tests/test_reference_pin_cpu.py
and tests/golden/make_reference_golden.py): residuals and the complete Jacobian rows of every reprojection term of two small problems - the
pose chain through the reference's expression nodes and its camera models - straight against the kernels, without the oracle in between."""
import numpy as np
import pytest

from test_reference_pin_cpu import GOLD, dense_from_ccs, term_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()  # fails loudly if the extension is missing
    return m


@pytest.mark.parametrize("tag", ["rig", "batch"])
def test_kernels_reproduce_the_reference_terms(capi, tag):
    g = np.load(GOLD)
    p = term_problem(g, tag)
    s = capi.B200SchurLinearSystemSolver(p)
    res, rows = g[f"term_{tag}_residuals"], g[f"term_{tag}_jacobian"]
    J0 = s.evaluate_error()
    assert abs(J0 - res @ res) <= 1e-11 * (res @ res)                              # cost (identity invR)
    assert np.abs(-s.error_vector() - res).max() <= 1e-9 * np.abs(res).max()       # e() = -(y - y_hat): the 1e-9 bar of the north star
    J = dense_from_ccs(*s.jacobian_ccs(), s.jcols)                                  # the materialising kernel (CCS J^T export)
    assert J.shape == rows.shape
    assert np.array_equal(J != 0, rows != 0)
    assert (np.abs(J - rows) / np.maximum(np.abs(rows).max(axis=1, keepdims=True), 1.0)).max() <= 1e-9
    # the fused path: rhs = -J^T e with the reference's J = d(y - y_hat)/dx (rows) and e = y - y_hat (res)
    s.build_system()
    rhs_ref = -(rows.T @ res)
    assert np.abs(s.rhs() - rhs_ref).max() <= 1e-9 * np.abs(rhs_ref).max()


@pytest.mark.parametrize("device_loop", [1, 0], ids=["device-loop", "host-loop"])
@pytest.mark.parametrize("n", range(11))
def test_kb_optimize_walks_the_reference_optimizer(capi, n, device_loop):
    """kb_optimize (the LM loop resident on the device, and the host-side Optimizer2 mirror over the call-by-call entry points) against what the
    REFERENCE's own compiled Optimizer2 / LevenbergMarquardtTrustRegionPolicy / BlockCholeskyLinearSystemSolver returned for the same eleven
    problems (tests/golden/make_reference_golden.py, oracle/ref_pin_optimizer.cpp): iteration and failed-iteration counts, the cost after every
    iteration, the final design variables, the sticky linear-solver failure of problem 3 - without the oracle in between"""
    from test_reference_pin_cpu import check_against_reference_optimizer

    g = np.load(GOLD)

    def solve(p, opt):
        opt.device_loop = device_loop
        s = capi.B200SchurLinearSystemSolver(p)
        sol, _ = s.optimize(opt)
        return sol, s.camera_params(), s.baselines(), s.set_poses()

    check_against_reference_optimizer(g, n, solve)
