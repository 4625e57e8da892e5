"""The CUDA path in the SparseCholesky semantic (kb_set_solver_semantic(1); Kalibr2's DEFAULT solver, BE/src/Optimizer2.cpp:83-86) against
values computed by the REFERENCE's own code (tests/golden/reference_sparse_golden.npz; tests/golden/make_reference_sparse_golden.py,
oracle/ref_pin_optimizer.cpp): the compressed-column J^T of CompressedColumnJacobianTransposeBuilder - column pointers and row indices bit for
bit, values 1e-9 -, e, rhs = J^T e, one damped step, and whole optimisations over SparseCholeskyLinearSystemSolver - without the oracle in
between.  Row a26 of SURVEY.md §8.

(File name sorts last on purpose: written at the end of round 2 with a few GPU-minutes left.)"""
import numpy as np
import pytest

from test_reference_pin_cpu import N_OPT, check_against_reference_optimizer
from test_reference_sparse_pin_cpu import SPARSE_GOLD, SYSTEM_TAGS, check_sparse_system, sparse_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()  # fails loudly if the extension is missing
    return m


@pytest.mark.parametrize("tag", SYSTEM_TAGS)
def test_kernels_reproduce_the_reference_sparse_system(capi, tag):
    def system(p, lam):
        s = capi.B200SchurLinearSystemSolver(p)
        s.set_solver_semantic(1)
        cost = s.evaluate_error()
        e = s.error_vector()
        col_ptr, row_idx, values = s.jacobian_ccs()  # the materialising kernel
        s.build_system()                             # the fused kernel
        rhs = s.rhs()
        s.set_constant_conditioner(lam)
        dx, ok = s.solve_system()
        return cost, e, col_ptr, row_idx, values, rhs, dx, ok

    check_sparse_system(np.load(SPARSE_GOLD), tag, system, 1e-9, 1e-7)  # the bars of the north star: e, J 1e-9; the step 1e-7


# problem 3: cond(J^T J + lambda^2 I) = 5.6e12 in its first iterations (see tests/test_reference_sparse_pin_cpu.py)
COST_RTOL = {3: 1e-7}


@pytest.mark.parametrize("device_loop", [1, 0], ids=["device-loop", "host-loop"])
@pytest.mark.parametrize("n", range(N_OPT))
def test_kb_optimize_walks_the_reference_sparse_optimizer(capi, n, device_loop):
    """kb_optimize in the SparseCholesky semantic (device-resident loop and host mirror) against what the REFERENCE's own compiled
    Optimizer2 / LevenbergMarquardtTrustRegionPolicy / SparseCholeskyLinearSystemSolver returned for the eleven problems: counts, the cost
    after every iteration, the final design variables"""
    g = sparse_golden()

    def solve(p, opt):
        opt.device_loop = device_loop
        s = capi.B200SchurLinearSystemSolver(p)
        s.set_solver_semantic(1)
        sol, _ = s.optimize(opt)
        return sol, s.camera_params(), s.baselines(), s.set_poses()

    check_against_reference_optimizer(g, n, solve, COST_RTOL.get(n, 1e-9))
