"""Golden vectors of the REFERENCE's camera models and SE(3) helpers -> tests/golden/reference_golden.npz.

Needs /root/reference (read-only) in the build container: oracle/ref_pin.cpp compiles the reference's own PinholeProjection /
OmniProjection / ExtendedUnifiedProjection / DoubleSphereProjection and *Distortion code where it lies (against the stand-in headers of
oracle/ref_shim/) into the git-ignored oracle/_ref/, and this script records what THAT code returns for seeded inputs: keypoint, point
Jacobian (2x4), intrinsics Jacobian, distortion Jacobian per model - on ordinary points, points behind the camera / outside the validity
cone (where some models return before writing), negative and zero homogeneous scale, and a point on the optical axis; and sm_kinematics'
quat2r, updateQuat (all its small-angle branches), boxMinus and boxTimes from the reference's quaternion_algebra.cpp / transformations.cpp;
and, for two small calibration problems, the residual and the whole Jacobian row pair of every reprojection term as the reference's expression
tree (RotationQuaternion, EuclideanPoint, TransformationBasic, the Transformation / Homogeneous expression nodes, JacobianContainer) and camera
models produce them; and, for eleven small calibration problems (all four design-variable orders, all seven models, damping seeds from
1e-8 to 1e3, three runs with a rejected step, one that ends in the sticky linear-solver failure), what the reference's own Optimizer2 / LevenbergMarquardtTrustRegionPolicy /
BlockCholeskyLinearSystemSolver loop returns (oracle/ref_pin_optimizer.cpp): iteration and failed-iteration counts, JStart / JFinal, the
final design variables, and the same for runs truncated after 1, 2, ... iterations (the cost per iteration).
    python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kalibr_b200 import synthetic  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402


def inputs(model, n=200, seed=0):
    """(params [k, 10], points [k, 4]): the generator's ground-truth parameters and two perturbed sets, n points each"""
    rng = np.random.default_rng(1000 + 17 * model + seed)
    base = np.asarray(synthetic.TRUTH_PARAMS[model][0], float)
    P, H = [], []
    for variant in range(3):
        prm = base * (1.0 + (0.05 * rng.standard_normal(base.shape) if variant else 0.0))
        for t in range(n):
            ph = np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.6, 0.6), rng.uniform(0.3, 3.0), rng.choice([1.0, 1.0, -1.0, 0.5, 0.0])])
            if t % 8 == 0:
                ph[2] = -abs(ph[2]) * rng.uniform(0.1, 3.0)  # behind the camera
            if t % 50 == 1:
                ph[:2] = 0.0                                  # on the optical axis
            if t % 50 == 2:
                ph[:2] *= 40.0                                # far off axis
            P.append(np.pad(prm, (0, 10 - len(prm))))
            H.append(ph)
    return np.array(P), np.array(H)


N_P = [4, 4, 5, 6, 6, 4, 5]
N_D = [4, 4, 4, 0, 0, 1, 0]


def reference_rows(p):
    """residuals y - y_hat and the dense Jacobian rows d(y - y_hat)/d(design variables) of every term of problem p, computed by the
    REFERENCE's code only: its expression tree (pose chain) and its camera models, combined as ReprojectionError does
    (CVE/.../implementation/ReprojectionError.hpp:50-77: point.evaluateJacobians(container, -J), camera.evaluateJacobians(container, p));
    the column of a design variable is p.dv_layout()'s"""
    col, dims, labels = p.dv_layout()
    off = {lab: int(c) for lab, c in zip(labels, col)}
    jcols = int(col[-1] + dims[-1])
    res, rows = np.zeros(2 * p.n_terms), np.zeros((2 * p.n_terms, jcols))
    for w in range(len(p.view_set)):
        v, k = int(p.view_set[w]), int(p.view_cam[w])
        model = int(p.cam_model[k])
        for i in range(int(p.view_begin[w]), int(p.view_begin[w + 1])):
            p4 = np.append(p.target_points[p.corner_id[i]], 1.0)
            pc, _ = oa.reference_point_chain(p.set_poses[v], p.baselines[:k], p4)
            y, Jp, Ji, Jd, _ = oa.reference_camera_project(model, p.cam_params[k], pc)
            _, Jpose = oa.reference_point_chain(p.set_poses[v], p.baselines[:k], p4, chain=-Jp)
            r = rows[2 * i:2 * i + 2]
            r[:, off[("set_q", v)]:off[("set_q", v)] + 3] = Jpose[0]
            r[:, off[("set_t", v)]:off[("set_t", v)] + 3] = Jpose[1]
            for j in range(k):
                r[:, off[("baseline_q", j)]:off[("baseline_q", j)] + 3] = Jpose[2 + 2 * j]
                r[:, off[("baseline_t", j)]:off[("baseline_t", j)] + 3] = Jpose[3 + 2 * j]
            r[:, off[("proj", k)]:off[("proj", k)] + N_P[model]] = -Ji[:, :N_P[model]]
            if N_D[model]:
                r[:, off[("dist", k)]:off[("dist", k)] + N_D[model]] = -Jd[:, :N_D[model]]
            res[2 * i:2 * i + 2] = np.array([p.y_u[i], p.y_v[i]]) - y
    return res, rows


# (tag, models, driver order, sets, seed): a rig with four different models and a batch-order problem with the remaining three
TERM_PROBLEMS = [("rig", [0, 2, 1, 4], 2, 3, 5), ("batch", [5, 3, 6], 3, 3, 6)]


# (models, sets, driver order, seed, lambda_init, dropout): the reference's optimiser loop on these (see reference_optimizations); the
# last three take a rejected step (revertLastStateUpdate, the lambda^2 / lambda damping of the solve that follows without a rebuild)
OPT_PROBLEMS = [([0], 6, 0, 1, 10.0, 0.5), ([0, 2], 5, 1, 5, 10.0, 0.5), ([5, 3, 6], 4, 3, 6, 10.0, 0.5), ([1, 4], 5, 2, 7, 1e-4, 0.5),
                ([0], 6, 0, 1, 1e-6, 0.5), ([2, 6], 4, 1, 9, 1e3, 0.5), ([0, 0], 5, 1, 3, 1e-8, 0.5), ([4], 6, 0, 11, 1e-7, 0.5),
                ([2], 4, 0, 3, 1e-2, 0.6), ([4, 4], 4, 1, 4, 1e-2, 0.6), ([5, 2], 4, 3, 5, 1e-2, 0.6)]
OPT_INPUTS = ("cam_model", "cam_params", "baselines", "set_poses", "target_points", "view_set", "view_cam", "view_begin", "y_u", "y_v", "corner_id")
MAX_TRUNCATED = 10


def reference_optimizations(out):
    from kalibr_b200.problem import KbOptimizerOptions

    for n, (models, n_sets, order, seed, lam0, dropout) in enumerate(OPT_PROBLEMS):
        p = synthetic.make_problem(models, n_sets, order, seed=seed, dropout=dropout)
        for name in OPT_INPUTS:
            out[f"opt{n}_{name}"] = getattr(p, name)
        out[f"opt{n}_order"], out[f"opt{n}_lambda_init"] = np.array(order), np.array(lam0)
        opt = KbOptimizerOptions.kalibr2_default()  # K2/CalibrationTools.hpp:57-66
        opt.lm_lambda_init = lam0
        r, cp, bl, sp = oa.reference_optimize(p, opt)
        out[f"opt{n}_result"] = np.array([r["iterations"], r["failed_iterations"], r["j_start"], r["j_final"], r["linear_solver_failure"]])
        out[f"opt{n}_final_cam_params"], out[f"opt{n}_final_baselines"], out[f"opt{n}_final_set_poses"] = cp, bl, sp
        rows = []
        for k in range(1, min(r["iterations"], MAX_TRUNCATED) + 1):  # the run cut after k iterations: cost and counts per iteration
            opt.max_iterations = k
            rk = oa.reference_optimize(p, opt)[0]
            rows.append([rk["iterations"], rk["failed_iterations"], rk["j_start"], rk["j_final"], rk["linear_solver_failure"]])
        out[f"opt{n}_truncated"] = np.array(rows)
        print("optimiser problem", n, models, "order", order, "lambda0", lam0, r)
    out["opt_count"] = np.array(len(OPT_PROBLEMS))


def main():
    out = {}
    reference_optimizations(out)
    for tag, models, order, n_sets, seed in TERM_PROBLEMS:
        p = synthetic.make_problem(models, n_sets, order, seed=seed, dropout=0.75)
        res, rows = reference_rows(p)
        for name in ("cam_model", "cam_params", "baselines", "set_poses", "target_points", "view_set", "view_cam", "view_begin", "y_u", "y_v", "corner_id"):
            out[f"term_{tag}_{name}"] = getattr(p, name)
        out[f"term_{tag}_order"] = np.array(order)
        out[f"term_{tag}_residuals"], out[f"term_{tag}_jacobian"] = res, rows
    for model in range(7):
        P, H = inputs(model)
        Y, JP, JI, JD, OK = [], [], [], [], []
        for prm, ph in zip(P, H):
            y, Jp, Ji, Jd, ok = oa.reference_camera_project(model, prm, ph)
            Y.append(y); JP.append(Jp); JI.append(Ji); JD.append(Jd); OK.append(ok)
        out[f"m{model}_params"], out[f"m{model}_ph"] = P, H
        out[f"m{model}_y"], out[f"m{model}_Jp"], out[f"m{model}_Ji"], out[f"m{model}_Jd"] = np.array(Y), np.array(JP), np.array(JI), np.array(JD)
        out[f"m{model}_ok"] = np.array(OK, np.int32)
    rng = np.random.default_rng(4242)
    Q, DQ, P4, T = [], [], [], []
    for t in range(300):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        dq = rng.standard_normal(3) * [1.0, 0.3, 1e-3, 1e-5, 1e-9, 0.0][t % 6]
        Tm = np.eye(4)
        Tm[:3, :3] = oa.reference_kinematics("quat2r", q)
        Tm[:3, 3] = 3.0 * rng.standard_normal(3)
        Q.append(q); DQ.append(dq); P4.append(rng.standard_normal(4) * [1.0, 10.0, 0.1][t % 3]); T.append(Tm)
    out["kin_q"], out["kin_dq"], out["kin_p4"], out["kin_T"] = np.array(Q), np.array(DQ), np.array(P4), np.array(T)
    out["kin_quat2r"] = np.array([oa.reference_kinematics("quat2r", q) for q in Q])
    out["kin_update_quat"] = np.array([oa.reference_kinematics("update_quat", q, dq) for q, dq in zip(Q, DQ)])
    out["kin_box_minus"] = np.array([oa.reference_kinematics("box_minus", p) for p in P4])
    out["kin_box_times"] = np.array([oa.reference_kinematics("box_times", t) for t in T])
    # M-estimator weights (BE/src/MEstimatorPolicies.cpp): kind, (p0, p1, p2), squared error -> weight
    K, PR, S, W = [], [], [], []
    for kind, prm in ((0, (0.0, 0.0, 0.0)), (1, (1.5, 0.0, 0.0)), (1, (0.3, 0.0, 0.0)), (2, (4.0, 0.0, 0.0)), (3, (2.5, 0.0, 0.0)), (4, (2.0, 0.999, 0.1)), (4, (2.0, 0.95, 0.3))):
        for sq in list(rng.uniform(0.0, 30.0, 60)) + [0.0, 1e-12, prm[0] ** 2, prm[0] ** 2 * (1 + 1e-9), 700.0]:
            K.append(kind); PR.append(prm); S.append(sq); W.append(oa.reference_m_estimator_weight(kind, sq, *prm))
    out["mest_kind"], out["mest_params"], out["mest_s"], out["mest_w"] = np.array(K, np.int32), np.array(PR), np.array(S), np.array(W)
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
