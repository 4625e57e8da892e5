"""GPU parity of the SparseCholesky semantic (kb_set_solver_semantic(1)): Kalibr2's default solver appends the damping as extra
columns of J^T, so repeated solves on one build leave no lambda^2 - lambda residual on the diagonal (quirk Q2 is BlockCholesky's only).
  BE/src/SparseCholeskyLinearSystemSolver.cpp:14-89 against BE/src/BlockCholeskyLinearSystemSolver.cpp:74-106

(File name sorts last on purpose: this semantic switch had no GPU test of its own before the end of round 1 and could not be run
on a GPU when it was written; everything else in the suite was.)
"""
import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max(initial=0.0) / max(np.abs(b).max(initial=0.0), 1e-300)


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()
    return m


@pytest.mark.parametrize("cfg,n_sets", [(1, 20), (2, 12), (3, 8)])
def test_repeated_solves_match_the_sparse_oracle(capi, oracle_lib, cfg, n_sets):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    g.set_solver_semantic(1)
    o = oracle_lib.OracleProblem(p, oracle_lib.SPARSE_CHOLESKY)
    b = oracle_lib.OracleProblem(p, oracle_lib.BLOCK_CHOLESKY)
    for s in (g, o, b):
        s.evaluate_error()
        s.build_system()
    dxs = []
    for lam in (10.0, 20.0, 5.0):  # three solves on one build
        for s in (g, o, b):
            s.set_constant_conditioner(lam)
        gdx, gok = g.solve_system()
        odx, ook = o.solve_system()
        bdx, _ = b.solve_system()
        assert gok and ook
        assert rel_err(gdx, odx) < 1e-7
        dxs.append((gdx, bdx))
    # from the second solve on the two semantics differ (the block solver carries the residual): the switch does something
    assert rel_err(dxs[1][0], dxs[1][1]) > 1e-6


@pytest.mark.parametrize("device_loop", [1, 0])
@pytest.mark.parametrize("cfg,n_sets", [(1, 25), (2, 14)])
def test_optimize_matches_the_sparse_oracle(capi, oracle_lib, cfg, n_sets, device_loop):
    p = synthetic.make_config(cfg, n_sets=n_sets)
    g = capi.B200SchurLinearSystemSolver(p)
    g.set_solver_semantic(1)
    o = oracle_lib.OracleProblem(p, oracle_lib.SPARSE_CHOLESKY)
    opt = KbOptimizerOptions.kalibr2_default()
    opt.device_loop = device_loop
    gs, gtr = g.optimize(opt)
    os_, otr = o.optimize(KbOptimizerOptions.kalibr2_default())
    assert gs.iterations == os_.iterations and gs.failed_iterations == os_.failed_iterations
    assert abs(gs.j_final - os_.j_final) <= 1e-9 * os_.j_final
    assert rel_err(gtr[:, 0], otr[:, 0]) < 1e-8
    assert rel_err(g.camera_params(), o.camera_params()) < 1e-6
