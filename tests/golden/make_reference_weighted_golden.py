"""Golden vectors of the REFERENCE's term weighting inside its solvers and its optimiser loop -> tests/golden/reference_weighted_golden.npz.

Needs /root/reference (read-only) in the build container.  Every reprojection term of the problems below gets invR through the reference's
ErrorTermFs<2>::setInvR and an M-estimator policy through ErrorTerm::setMEstimatorPolicy (BE/include/aslam/backend/implementation/ErrorTerm.hpp,
BE/src/ErrorTerm.cpp, BE/src/MEstimatorPolicies.cpp, compiled from their sources: oracle/ref_pin_optimizer.cpp, ref_set_weighting); a few gross
outliers give the policies something to act on.  Recorded per case: the weighted compressed-column J^T, the weighted error vector, rhs, the
cost (policy-weighted: ErrorTerm.cpp:19-24), one damped step, and what Optimizer2::optimize returns over the BlockCholesky and the
SparseCholesky solver (counts, JStart / JFinal, final design variables, the runs cut after 1, 2, ... iterations).
Stand-ins in that build, besides those make_reference_sparse_golden.py names: the matrix square root behind setInvR (Eigen::LDLT; exact
for the invR = c I Kalibr2 passes, a restated pivoted LDL^T for the one general matrix of case "general_huber") and Boost.Math's
chi-squared quantile behind Blake-Zisserman's epsilon.
    python tests/golden/make_reference_weighted_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kalibr_b200 import synthetic  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402
from make_reference_golden import MAX_TRUNCATED  # noqa: E402

INV_R_ISO = np.eye(2) / (0.3 * 0.3)  # I / sigma^2 as CreateBatchProblem passes (K2/CalibrationTools.hpp:495-496)
INV_R = np.array([[3.0, 0.4], [0.4, 5.0]])
IDENTITY = np.eye(2)
NO_POLICY = (0, 0.0, 0.999, 0.1)
# (tag, models, driver order, sets, seed, invR, (kind, p0, p1, p2))
CASES = [("iso", [0, 1], 1, 5, 2, INV_R_ISO, NO_POLICY), ("huber", [0, 2, 4], 2, 4, 5, IDENTITY, (1, 1.5, 0.0, 0.0)),
         ("cauchy_iso", [2], 0, 6, 1, INV_R_ISO, (2, 4.0, 0.0, 0.0)), ("geman", [5, 3], 3, 4, 6, IDENTITY, (3, 9.0, 0.0, 0.0)),
         ("blake", [1], 0, 5, 3, IDENTITY, (4, 2.0, 0.999, 0.1)), ("general_huber", [0, 6], 1, 4, 8, INV_R, (1, 1.5, 0.0, 0.0))]
INPUTS = ("cam_model", "cam_params", "baselines", "set_poses", "target_points", "view_set", "view_cam", "view_begin", "y_u", "y_v", "corner_id")
LAMBDA = 10.0


def make_case(models, order, n_sets, seed):
    p = synthetic.make_problem(models, n_sets, order, seed=seed, dropout=0.6)
    rng = np.random.default_rng(100 + seed)
    bad = rng.choice(p.n_terms, max(p.n_terms // 40, 1), replace=False)  # a few gross outliers
    p.y_u[bad] += rng.normal(0, 15.0, bad.size)
    p.y_v[bad] += rng.normal(0, 15.0, bad.size)
    return p


def main():
    from kalibr_b200.problem import KbOptimizerOptions

    assert oa.build_reference_cameras() is not None, "needs the reference tree"
    out = {}
    for tag, models, order, n_sets, seed, inv_r, policy in CASES:
        p = make_case(models, order, n_sets, seed)
        oa.reference_set_weighting(inv_r, policy)
        for name in INPUTS:
            out[f"{tag}_{name}"] = getattr(p, name)
        out[f"{tag}_order"], out[f"{tag}_inv_r"], out[f"{tag}_policy"] = np.array(order), np.array(inv_r), np.array(policy, float)
        r = oa.reference_sparse_system(p, LAMBDA, 1)
        for k in ("col_ptr", "row_ind", "values", "e", "rhs", "dx"):
            out[f"{tag}_{k}"] = r[k]
        out[f"{tag}_cost"] = np.array(r["cost"])
        for kind, name in ((oa.BLOCK_CHOLESKY_KIND, "block"), (oa.SPARSE_CHOLESKY_KIND, "sparse")):
            opt = KbOptimizerOptions.kalibr2_default()
            res, cp, bl, sp = oa.reference_optimize(p, opt, kind, 1)
            out[f"{tag}_{name}_result"] = np.array([res["iterations"], res["failed_iterations"], res["j_start"], res["j_final"], res["linear_solver_failure"]])
            out[f"{tag}_{name}_final_cam_params"], out[f"{tag}_{name}_final_baselines"], out[f"{tag}_{name}_final_set_poses"] = cp, bl, sp
            rows = []
            for k in range(1, min(res["iterations"], MAX_TRUNCATED) + 1):
                opt.max_iterations = k
                rk = oa.reference_optimize(p, opt, kind, 1)[0]
                rows.append([rk["iterations"], rk["failed_iterations"], rk["j_start"], rk["j_final"], rk["linear_solver_failure"]])
            out[f"{tag}_{name}_truncated"] = np.array(rows)
            print(tag, name, res)
    oa.reference_set_weighting()
    out["tags"] = np.array([c[0] for c in CASES])
    out["lambda"] = np.array(LAMBDA)
    path = os.path.join(ROOT, "tests", "golden", "reference_weighted_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
