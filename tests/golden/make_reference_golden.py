"""Golden vectors of the REFERENCE's camera models and SE(3) helpers -> tests/golden/reference_golden.npz.

Needs /root/reference (read-only) in the build container: oracle/ref_pin.cpp compiles the reference's own PinholeProjection /
OmniProjection / ExtendedUnifiedProjection / DoubleSphereProjection and *Distortion code where it lies (against the stand-in headers of
oracle/ref_shim/) into the git-ignored oracle/_ref/, and this script records what THAT code returns for seeded inputs: keypoint, point
Jacobian (2x4), intrinsics Jacobian, distortion Jacobian per model - on ordinary points, points behind the camera / outside the validity
cone (where some models return before writing), negative and zero homogeneous scale, and a point on the optical axis; and sm_kinematics'
quat2r, updateQuat (all its small-angle branches), boxMinus and boxTimes from the reference's quaternion_algebra.cpp / transformations.cpp.
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


def main():
    out = {}
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
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
