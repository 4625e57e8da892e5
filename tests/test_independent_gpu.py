"""The CUDA path against the INDEPENDENT derivation of tests/independent_model.py (published model formulas in numpy, finite-difference
Jacobians over the design variables' manifold, scipy's trust-region least squares): residuals, every Jacobian column and the converged
calibration, without the oracle in between.  The one deviation is the reference's own (EUCM quirk Q4, reproduced on purpose)."""
import numpy as np
import pytest

import independent_model as im
from kalibr_b200 import synthetic
from test_independent_cpu import CASES, SCIPY_CASES, apply_reference_quirks, converge, dense_from_ccs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import capi as m

    m.load_library()
    return m


@pytest.mark.parametrize("models,order,n_sets", CASES)
def test_device_residuals_and_jacobians_match_the_independent_model(capi, models, order, n_sets):
    p = synthetic.make_problem(models, n_sets, order, seed=900 + sum(models), dropout=0.6)
    st = im.State(p)
    g = capi.B200SchurLinearSystemSolver(p)
    J0 = g.evaluate_error()
    e_ind = im.residuals(p, st)
    assert np.abs(g.error_vector() - e_ind).max() <= 1e-10 * np.abs(e_ind).max()
    assert abs(J0 - e_ind @ e_ind) <= 1e-12 * J0
    Jg = dense_from_ccs(*g.jacobian_ccs(), g.jcols)
    Jfd = apply_reference_quirks(p, -im.fd_jacobian(p, st))
    scale = np.abs(Jfd).max(axis=1, keepdims=True)
    assert (np.abs(Jg - Jfd) / scale).max() < 2e-6
    # and the normal equations are J^T J / -J^T e of that Jacobian
    g.build_system()
    rhs = g.rhs()
    ref = -Jfd.T @ (-e_ind)  # rhs = -J^T (y - y_hat) with e_ind = -(y - y_hat)
    assert np.abs(rhs - (-Jfd.T @ -e_ind)).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("models,order,n_sets,dropout", SCIPY_CASES)
def test_device_calibration_matches_scipy_least_squares(capi, models, order, n_sets, dropout):
    p = synthetic.make_problem(models, n_sets, order, seed=77 + sum(models), dropout=dropout)
    g = capi.B200SchurLinearSystemSolver(p)
    sol = converge(g)
    assert sol.iterations < 200
    cost, st = im.least_squares_calibration(p, im.State(p))
    assert abs(sol.j_final - cost) <= 1e-9 * cost
    gc = g.camera_params()
    assert (np.abs(gc - st.cam) / np.maximum(np.abs(st.cam), 1e-3)).max() < 1e-6
    for a, b in zip(g.baselines(), st.base):
        assert np.abs(im.rot_from_quat(a[:4]) - im.rot_from_quat(b[:4])).max() < 1e-6 and np.abs(a[4:] - b[4:]).max() < 1e-6
