"""The oracle against an INDEPENDENT derivation (tests/independent_model.py: the camera models from their published formulas in plain
numpy, every Jacobian column by central finite differences over the design variables' manifold, the calibration by scipy's
trust-region least squares) - a check that does not share the oracle's analytic Jacobians, expression-tree chain rule or LM policy.
The GPU twin is tests/test_independent_gpu.py."""
import numpy as np
import pytest

import independent_model as im
from kalibr_b200 import synthetic
from kalibr_b200.problem import KbOptimizerOptions

# (models, driver order, sets): every model of the table, every design-variable order
CASES = [([0], 0, 3), ([1], 0, 3), ([5, 6], 1, 3), ([2, 3, 4], 2, 3), ([0, 2, 1], 3, 3)]


def dense_from_ccs(col_ptr, row_idx, values, jcols):
    J = np.zeros((len(col_ptr) - 1, jcols))
    for r in range(len(col_ptr) - 1):
        J[r, row_idx[col_ptr[r]:col_ptr[r + 1]]] = values[col_ptr[r]:col_ptr[r + 1]]
    return J


def apply_reference_quirks(p, J):
    """The one place where the reference's ANALYTIC Jacobian is not the derivative of its own projection: the EUCM
    intrinsics Jacobian scales the v-row of d/d(alpha, beta) with fu where fv is meant (ExtendedUnifiedProjection.hpp:438-439).  The
    independent finite differences find exactly this deviation - and nothing else; the oracle and the kernels reproduce the reference."""
    J = J.copy()
    col, dims, labels = p.dv_layout()
    term_cam = np.repeat(p.view_cam, np.diff(p.view_begin))
    for b, (kind, k) in enumerate(labels):
        if kind == "proj" and int(p.cam_model[k]) == im.EUCM_NONE:
            fu, fv = p.cam_params[k][2], p.cam_params[k][3]
            v_rows = 2 * np.flatnonzero(term_cam == k) + 1
            J[np.ix_(v_rows, [int(col[b]), int(col[b]) + 1])] *= fu / fv
    return J


@pytest.mark.parametrize("models,order,n_sets", CASES)
def test_oracle_residuals_and_jacobians_match_the_independent_model(oracle_lib, models, order, n_sets):
    p = synthetic.make_problem(models, n_sets, order, seed=900 + sum(models), dropout=0.6)
    st = im.State(p)
    o = oracle_lib.OracleProblem(p)
    J0 = o.evaluate_error()
    e_ind = im.residuals(p, st)
    assert np.abs(o.error_vector() - e_ind).max() <= 1e-10 * np.abs(e_ind).max()
    assert abs(J0 - e_ind @ e_ind) <= 1e-12 * J0
    Jo = dense_from_ccs(*o.jacobian_ccs(), o.jcols)
    # the exported rows are d(y - y_hat)/dx (the reference's J); e() = -(y - y_hat)
    Jfd = apply_reference_quirks(p, -im.fd_jacobian(p, st))
    scale = np.abs(Jfd).max(axis=1, keepdims=True)
    assert (np.abs(Jo - Jfd) / scale).max() < 2e-6  # central differences with h = 1e-6


# well-conditioned cases (a handful of sets, half of the corners): the reference's LM loop and scipy's trust-region solver must land on
# the same minimum.  (Rigs with an omni / equidistant camera and this little data have a nearly flat valley - xi against the focal
# length, the higher equidistant coefficients - in which the LM loop crawls for its 200 iterations: not a parity question.)
SCIPY_CASES = [([0], 0, 8, 0.5), ([0, 5], 3, 6, 0.6)]


def converge(problem_like):
    opt = KbOptimizerOptions.kalibr2_default()
    opt.convergence_delta_x, opt.convergence_delta_j = 1e-9, 1e-12  # run the LM loop to the minimum, not to kalibr2's early stop
    return problem_like.optimize(opt)[0]


@pytest.mark.parametrize("models,order,n_sets,dropout", SCIPY_CASES)
def test_oracle_calibration_matches_scipy_least_squares(oracle_lib, models, order, n_sets, dropout):
    p = synthetic.make_problem(models, n_sets, order, seed=77 + sum(models), dropout=dropout)
    o = oracle_lib.OracleProblem(p)
    sol = converge(o)
    assert sol.iterations < 200
    cost, st = im.least_squares_calibration(p, im.State(p))
    assert abs(sol.j_final - cost) <= 1e-9 * cost
    oc = o.camera_params()
    assert (np.abs(oc - st.cam) / np.maximum(np.abs(st.cam), 1e-3)).max() < 1e-6
    for a, b in zip(o.baselines(), st.base):
        assert np.abs(im.rot_from_quat(a[:4]) - im.rot_from_quat(b[:4])).max() < 1e-6 and np.abs(a[4:] - b[4:]).max() < 1e-6
