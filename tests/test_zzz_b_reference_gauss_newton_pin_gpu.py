"""kb_optimize_gauss_newton (Optimizer2 with the Gauss-Newton policy over the truncated-SVD solver, resident on the device) against what the
REFERENCE's own compiled Optimizer2 / GaussNewtonTrustRegionPolicy returned (tests/golden/reference_gauss_newton_golden.npz; see
tests/test_reference_gauss_newton_pin_cpu.py for what that fixture does and does not pin): iteration counts, the cost after every iteration,
the final design variables - on the problems where the solver's options truncate nothing, so that its step is the least-squares step the
fixture's Cholesky solve takes.

(File name sorts after every test that HAS run on a GPU, on purpose: written at the end of round 2 after the GPU budget of the round was spent; the oracle reproduces this
fixture on the CPU and the device loop reproduces the oracle's in tests/test_estimator_gpu.py.)"""
import numpy as np
import pytest

from test_reference_gauss_newton_pin_cpu import GN_GOLD, HELD, check_against_reference_gauss_newton

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()  # fails loudly if the extension is missing
    return m


def solver_options(name):
    from kalibr_b200.problem import KbSvdSolverOptions

    if name == "kalibr2":
        return KbSvdSolverOptions.kalibr2()
    o = KbSvdSolverOptions.default()
    o.column_scaling = 1
    return o


# problem 2 stays on the CPU: even the float64 numpy solve of its scaled system sits 5e-10 from the fixture's cost after the first step
CASES = [(n, "column-scaling") for n in HELD["column-scaling"] if n != 2] + [(n, "kalibr2") for n in HELD["kalibr2"]]


@pytest.mark.parametrize("n,options", CASES)
def test_kb_optimize_gauss_newton_walks_the_reference_optimizer(capi, n, options):
    from kalibr_b200.problem import KbOptimizerOptions

    def solve(p, max_iterations):
        opt = KbOptimizerOptions.estimator_default()
        opt.max_iterations = max_iterations
        s = capi.B200SchurLinearSystemSolver(p)
        sol, _ = s.optimize_gauss_newton(opt, solver_options(options))
        assert not sol.linear_solver_failure
        return sol.iterations, sol.failed_iterations, sol.j_start, sol.j_final, s.camera_params(), s.baselines(), s.set_poses()

    check_against_reference_gauss_newton(np.load(GN_GOLD), n, solve)
