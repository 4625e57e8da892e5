"""Golden vectors of the REFERENCE's Optimizer2 under GaussNewtonTrustRegionPolicy -> tests/golden/reference_gauss_newton_golden.npz.

This is the optimiser loop the incremental estimator runs (IC/src/core/IncrementalEstimator.cpp:343-377: Optimizer2 with the Gauss-Newton
policy: build + UNDAMPED solve every iteration, never revert; BE/src/GaussNewtonTrustRegionPolicy.cpp, BE/src/Optimizer2.cpp:183-273), compiled
from the reference's sources (oracle/ref_pin_optimizer.cpp, ref_set_trust_region_policy) and run with the estimator's options (deltas 1e-3,
20 iterations) on ten full-rank problems in all four design-variable orders.  The linear solver in this build is the reference's
SparseCholeskyLinearSystemSolver (over the dense stand-in for CHOLMOD), NOT the estimator's own aslam::calibration::LinearSolver (SuiteSparseQR
+ SVD, not buildable here): at full rank both return the least-squares step, so what this fixture pins is the POLICY and the loop - iteration
counts, the cost after every iteration, the stopping rule, the final design variables - not the truncated-SVD solver.
For the batch-order problems the same optimisation also runs over the estimator's MERGED problem built from the reference's own containers
(aslam::calibration::OptimizationProblem per synced set, filled as kalibr2's CreateBatchProblem fills it; IncrementalOptimizationProblem::add;
the groups ordering of IncrementalEstimator::orderMarginalizedDesignVariables): recorded is the order - kind, index, column base, dimension -
in which Optimizer2 enumerates its active design variables (oracle/ref_pin_optimizer.cpp: ref_estimator_problem).
    python tests/golden/make_reference_gauss_newton_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kalibr_b200 import synthetic  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402

# (models, sets, driver order, seed): problems on which the undamped iteration converges from the generator's perturbed start
GN_PROBLEMS = [([0], 6, 0, 1), ([0, 0], 5, 1, 3), ([0, 2], 5, 2, 5), ([5, 3, 6], 4, 3, 6), ([6], 5, 0, 4), ([3], 6, 0, 2), ([4], 6, 0, 11), ([1], 6, 0, 12),
               ([0, 5], 4, 3, 13), ([0, 0, 0], 4, 2, 14)]
INPUTS = ("cam_model", "cam_params", "baselines", "set_poses", "target_points", "view_set", "view_cam", "view_begin", "y_u", "y_v", "corner_id")


def main():
    from kalibr_b200.problem import KbOptimizerOptions

    assert oa.build_reference_cameras() is not None, "needs the reference tree"
    out = {}
    oa.reference_set_trust_region_policy(True)
    try:
        for n, (models, n_sets, order, seed) in enumerate(GN_PROBLEMS):
            p = synthetic.make_problem(models, n_sets, order, seed=seed, dropout=0.5)
            for name in INPUTS:
                out[f"gn{n}_{name}"] = getattr(p, name)
            out[f"gn{n}_order"] = np.array(order)
            opt = KbOptimizerOptions.estimator_default()
            r, cp, bl, sp = oa.reference_optimize(p, opt, oa.SPARSE_CHOLESKY_KIND, 1)
            rb = oa.reference_optimize(p, opt, oa.BLOCK_CHOLESKY_KIND, 1)[0]  # no damping: both solvers walk the same iterations
            assert (r["iterations"], r["failed_iterations"]) == (rb["iterations"], rb["failed_iterations"]) and abs(r["j_final"] - rb["j_final"]) <= 1e-10 * r["j_final"]
            assert not r["linear_solver_failure"]
            out[f"gn{n}_result"] = np.array([r["iterations"], r["failed_iterations"], r["j_start"], r["j_final"], r["linear_solver_failure"]])
            out[f"gn{n}_final_cam_params"], out[f"gn{n}_final_baselines"], out[f"gn{n}_final_set_poses"] = cp, bl, sp
            rows = []
            for k in range(1, r["iterations"] + 1):
                opt.max_iterations = k
                rk = oa.reference_optimize(p, opt, oa.SPARSE_CHOLESKY_KIND, 1)[0]
                rows.append([rk["iterations"], rk["failed_iterations"], rk["j_start"], rk["j_final"], rk["linear_solver_failure"]])
            out[f"gn{n}_truncated"] = np.array(rows)
            print("Gauss-Newton problem", n, models, "order", order, r)
    finally:
        oa.reference_set_trust_region_policy(False)
    # the estimator's MERGED problem through the reference's own containers (ref_estimator_problem): the order of the active design variables
    # as Optimizer2 enumerates them over IncrementalOptimizationProblem, for the batch-order problems above
    for n, (models, n_sets, order, seed) in enumerate(GN_PROBLEMS):
        if order != 3:
            continue
        p = synthetic.make_problem(models, n_sets, order, seed=seed, dropout=0.5)
        dv_order, groups, r, cp, bl, sp = oa.reference_estimator_problem(p, KbOptimizerOptions.estimator_default())
        assert [r["iterations"], r["failed_iterations"], r["j_start"], r["j_final"], r["linear_solver_failure"]] == list(out[f"gn{n}_result"])  # the same run
        out[f"gn{n}_estimator_order"], out[f"gn{n}_estimator_groups"] = dv_order, np.array(groups)
        print("merged estimator problem", n, "groups ordering", groups, "active design variables", len(dv_order))
    out["gn_count"] = np.array(len(GN_PROBLEMS))
    path = os.path.join(ROOT, "tests", "golden", "reference_gauss_newton_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
