"""CUDA path against the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", ["cfg1_S3", "cfg2_S2", "cfg3_S2", "cfg4_S1", "cfg6_S2", "cfg7_S2"])
def test_cuda_path_reproduces_golden(name):
    from kalibr_b200 import capi

    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = synthetic.make_config(int(g["cfg"]), n_sets=int(g["n_sets"]))
    assert np.array_equal(p.y_u, g["y_u"]), "generator changed: regenerate the fixtures"
    s = capi.B200SchurLinearSystemSolver(p)
    assert abs(s.evaluate_error() - float(g["J0"])) <= 1e-11 * float(g["J0"])
    assert rel(s.error_vector(), g["e"]) < 1e-9
    cp, ri, jv = s.jacobian_ccs()
    assert np.array_equal(cp, g["jt_col_ptr"]) and np.array_equal(ri, g["jt_row_idx"])
    assert rel(jv, g["jt_values"]) < 1e-9
    s.build_system()
    assert rel(s.rhs(), g["rhs"]) < 1e-9
    s.set_constant_conditioner(10.0)
    dx, ok = s.solve_system()
    assert ok == bool(g["pos_def"]) and rel(dx, g["dx"]) < 1e-7
    hcp, hbr, hvp, hval = s.hessian_blocks()
    assert np.array_equal(hcp, g["h_col_ptr"]) and np.array_equal(hbr, g["h_block_row"]) and np.array_equal(hvp, g["h_value_ptr"])
    assert rel(hval, g["h_values"]) < 1e-9
    s.reset_state()
    sol, tr = s.optimize(KbOptimizerOptions.kalibr2_default())
    assert sol.iterations == int(g["iterations"]) and sol.failed_iterations == int(g["failed_iterations"])
    assert abs(sol.j_final - float(g["j_final"])) <= 1e-9 * float(g["j_final"])
    assert rel(s.camera_params(), g["cam_params"]) < 1e-6
    assert rel(s.set_poses(), g["set_poses"]) < 1e-6
    if p.n_cams > 1:
        assert rel(s.baselines(), g["baselines"]) < 1e-6
