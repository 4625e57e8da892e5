"""Golden vectors of the REFERENCE's SparseCholesky regime (Kalibr2's default solver: BE/src/Optimizer2.cpp:83-86) ->
tests/golden/reference_sparse_golden.npz.

Needs /root/reference (read-only) in the build container.  oracle/ref_pin_optimizer.cpp compiles the reference's own
SparseCholeskyLinearSystemSolver.cpp, CompressedColumnJacobianTransposeBuilder, CompressedColumnMatrix and Cholmod wrapper (with
Optimizer2, the LM policy, ErrorTerm, JacobianContainer, the expression tree and the camera models, as make_reference_golden.py describes)
against the stand-in headers of oracle/ref_shim/; only the factorisation behind the Cholmod wrapper is a stand-in (ref_shim/cholmod.h: dense
Cholesky of A A^T).  Recorded here:
  * for the two term problems of make_reference_golden.py and two more (stereo order, single-camera order): J^T in compressed-column form
    exactly as the reference's builder lays it out (column pointers, row indices, values), the error vector, rhs = J^T e as
    SparseCholeskyLinearSystemSolver::buildSystem forms it, the cost, and dx of one solveSystem under the constant conditioner 10;
  * for the eleven optimiser problems of make_reference_golden.py (all four design-variable orders, all seven models, damping seeds from
    1e-8 to 1e3, runs with rejected steps): what Optimizer2::optimize returns over the SparseCholesky solver - iteration and
    failed-iteration counts, JStart / JFinal, linearSolverFailure, the final design variables - and the same scalars for the runs cut after
    1, 2, ... iterations (the cost per iteration).  The inputs are those stored under the same opt<n>_* keys of reference_golden.npz.
    python tests/golden/make_reference_sparse_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kalibr_b200 import synthetic  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402
from make_reference_golden import MAX_TRUNCATED, OPT_PROBLEMS, TERM_PROBLEMS  # noqa: E402

# (tag, models, driver order, sets, seed): the term problems of make_reference_golden.py plus a stereo-order and a single-camera problem
SYSTEM_PROBLEMS = list(TERM_PROBLEMS) + [("stereo", [0, 1], 1, 4, 2), ("single", [2], 0, 5, 1)]
SYSTEM_INPUTS = ("cam_model", "cam_params", "baselines", "set_poses", "target_points", "view_set", "view_cam", "view_begin", "y_u", "y_v", "corner_id")
LAMBDA = 10.0


def main():
    from kalibr_b200.problem import KbOptimizerOptions

    assert oa.build_reference_cameras() is not None, "needs the reference tree"
    out = {}
    for tag, models, order, n_sets, seed in SYSTEM_PROBLEMS:
        p = synthetic.make_problem(models, n_sets, order, seed=seed, dropout=0.75)
        r = oa.reference_sparse_system(p, LAMBDA, 1)
        r4 = oa.reference_sparse_system(p, LAMBDA, 4)  # the threaded materialisation writes the same matrix
        assert all(np.array_equal(r[k], r4[k]) for k in ("col_ptr", "row_ind", "values", "e", "rhs", "dx"))
        for name in SYSTEM_INPUTS:
            out[f"sys_{tag}_{name}"] = getattr(p, name)
        out[f"sys_{tag}_order"] = np.array(order)
        for k in ("col_ptr", "row_ind", "values", "e", "rhs", "dx"):
            out[f"sys_{tag}_{k}"] = r[k]
        out[f"sys_{tag}_cost"] = np.array(r["cost"])
        print("system", tag, models, "order", order, "J^T", r["jcols"], "x", len(r["e"]), "nnz", len(r["values"]), "cost", r["cost"])
    out["sys_tags"] = np.array([t[0] for t in SYSTEM_PROBLEMS])
    out["sys_lambda"] = np.array(LAMBDA)
    for n, (models, n_sets, order, seed, lam0, dropout) in enumerate(OPT_PROBLEMS):
        p = synthetic.make_problem(models, n_sets, order, seed=seed, dropout=dropout)
        opt = KbOptimizerOptions.kalibr2_default()  # K2/CalibrationTools.hpp:57-66
        opt.lm_lambda_init = lam0
        r, cp, bl, sp = oa.reference_optimize(p, opt, oa.SPARSE_CHOLESKY_KIND, 1)
        out[f"opt{n}_result"] = np.array([r["iterations"], r["failed_iterations"], r["j_start"], r["j_final"], r["linear_solver_failure"]])
        out[f"opt{n}_final_cam_params"], out[f"opt{n}_final_baselines"], out[f"opt{n}_final_set_poses"] = cp, bl, sp
        rows = []
        for k in range(1, min(r["iterations"], MAX_TRUNCATED) + 1):
            opt.max_iterations = k
            rk = oa.reference_optimize(p, opt, oa.SPARSE_CHOLESKY_KIND, 1)[0]
            rows.append([rk["iterations"], rk["failed_iterations"], rk["j_start"], rk["j_final"], rk["linear_solver_failure"]])
        out[f"opt{n}_truncated"] = np.array(rows)
        print("optimiser problem", n, models, "order", order, "lambda0", lam0, r)
    out["opt_count"] = np.array(len(OPT_PROBLEMS))
    path = os.path.join(ROOT, "tests", "golden", "reference_sparse_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
