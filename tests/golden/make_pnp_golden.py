"""Generates tests/golden/pnp_cv2.npz: outputs of OpenCV's own cv::solvePnP (cv2 4.13, the Python build of the library the
reference links: CAM/include/aslam/cameras/implementation/PinholeProjection.hpp:876) on the (Ps, Ms) the reference's
estimateTransformation would hand it, for small seeded problems.

    python tests/golden/make_pnp_golden.py

These are REAL third-party outputs: they pin oracle/ko_init.py's restatement of solvePnP and, through it and directly, the
CUDA path (tests/test_init_cpu.py, tests/test_init_gpu.py).  The back-projection in front of the PnP is the oracle's.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from kalibr_b200 import synthetic  # noqa: E402
from oracle import ko_init as ki  # noqa: E402

# name -> (config, synced sets, generator options)
CASES = {
    "cfg1_S6": (1, 6, {}),
    "cfg2_S9": (2, 9, {}),
    "cfg3_S5": (3, 5, {}),
    "cfg4_S2": (4, 2, {}),
    "cfg6_S4": (6, 4, {}),
    "cfg7_S7": (7, 7, {}),
    "cfg3_S4_ragged": (3, 4, {"dropout": 0.3}),
}


def main():
    import cv2

    out = {"cv2_version": np.array(cv2.__version__)}
    for name, (cfg, S, kw) in CASES.items():
        p = synthetic.make_config(cfg, n_sets=S, **kw)
        T, ok = ki.view_transformations(p, pnp=ki.cv2_pnp)
        out[name + "/T_views"] = T
        out[name + "/ok"] = ok
        sp, good = ki.target_pose_guesses(p, pnp=ki.cv2_pnp)
        out[name + "/set_poses"] = sp
        out[name + "/set_ok"] = good
        if p.n_cams >= 2:
            out[name + "/baseline01"] = ki.stereo_baseline_guess(p, 0, 1, pnp=ki.cv2_pnp)
        # initializeIntrinsics per camera with cv2's PnP inside the omni family's candidate test
        res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
        for k in range(p.n_cams):
            prm, ok_k = ki.initialize_intrinsics(p, k, 10, 12, res[k], pnp=ki.cv2_pnp)
            out[name + f"/init_params{k}"] = prm
            out[name + f"/init_ok{k}"] = np.array(ok_k)
        # the raw PnP problem of the first view, for a test of solve_pnp alone
        b, e = p.view_begin[0], p.view_begin[1]
        k = p.view_cam[0]
        Ps, Ms = ki.pnp_inputs(p.cam_model[k], p.cam_params[k], p.y_u[b:e], p.y_v[b:e], p.target_points[p.corner_id[b:e]])
        r, t = ki.cv2_pnp(Ps, Ms)
        out[name + "/Ps"], out[name + "/Ms"], out[name + "/rvec"], out[name + "/tvec"] = Ps, Ms, r, t
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pnp_cv2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
