"""The CUDA path with term weighting (kb_set_inv_r, kb_set_m_estimator) against values computed by the REFERENCE's own code
(tests/golden/reference_weighted_golden.npz; tests/golden/make_reference_weighted_golden.py, oracle/ref_pin_optimizer.cpp): the weighted
compressed-column J^T, weighted e, rhs, the policy-weighted cost, one damped step, and whole optimisations over both solver semantics -
without the oracle in between.  Rows a2 / a19 of SURVEY.md §8.

(File name sorts after every test that HAS run on a GPU, on purpose: written at the end of round 2 after the GPU budget of the round was spent; the oracle reproduces this
fixture to 1e-13 on the CPU and the kernels reproduce the oracle's weighted path in tests/test_weighting_gpu.py.)"""
import numpy as np
import pytest

from test_reference_pin_cpu import check_against_reference_optimizer
from test_reference_weighted_pin_cpu import SOLVERS, TAGS, WEIGHTED_GOLD, check_weighted_system, optimizer_view, weighted_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()  # fails loudly if the extension is missing
    return m


@pytest.mark.parametrize("tag", TAGS)
def test_kernels_reproduce_the_reference_weighted_system(capi, tag):
    def system(p, inv_r, policy, lam):
        s = capi.B200SchurLinearSystemSolver(p)
        s.set_solver_semantic(1)
        s.set_inv_r(inv_r)
        s.set_m_estimator(*policy)
        cost = s.evaluate_error()
        e = s.error_vector()
        col_ptr, row_idx, values = s.jacobian_ccs()  # the materialising kernel, WEIGHTED instantiation
        s.build_system()                             # the fused kernel, WEIGHTED instantiation
        rhs = s.rhs()
        s.set_constant_conditioner(lam)
        dx, ok = s.solve_system()
        return cost, e, col_ptr, row_idx, values, rhs, dx, ok

    check_weighted_system(np.load(WEIGHTED_GOLD), tag, system, 1e-9, 1e-7)


# the policies tests/test_weighting_gpu.py optimises with from a perturbed state (Huber, Cauchy) and the plain invR cases
OPTIMIZE_TAGS = ["iso", "huber", "cauchy_iso", "general_huber"]


@pytest.mark.parametrize("device_loop", [1, 0], ids=["device-loop", "host-loop"])
@pytest.mark.parametrize("solver,semantic", SOLVERS)
@pytest.mark.parametrize("tag", OPTIMIZE_TAGS)
def test_kb_optimize_walks_the_reference_weighted_optimizer(capi, tag, solver, semantic, device_loop):
    g = np.load(WEIGHTED_GOLD)
    _, inv_r, policy = weighted_problem(g, tag)

    def solve(p, opt):
        opt.device_loop = device_loop
        s = capi.B200SchurLinearSystemSolver(p)
        s.set_solver_semantic(semantic)
        s.set_inv_r(inv_r)
        s.set_m_estimator(*policy)
        sol, _ = s.optimize(opt)
        return sol, s.camera_params(), s.baselines(), s.set_poses()

    check_against_reference_optimizer(optimizer_view(g, tag, solver), 0, solve)
