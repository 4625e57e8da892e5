"""Full-size (BASELINE.json sizes) property checks of the widened rows on the GPU — no oracle runs at these sizes, so the checks are
size-independent identities:
  weighting   invR = c I scales the cost by c and leaves dx unchanged; a Huber threshold above every residual equals no policy;
              reprojection statistics agree with numpy on the downloaded error vector
  init stage  every PnP succeeds, the target-pose guesses sit next to the generator's (perturbed) poses, the initialised problem converges to
              the same cost as the generator's start
python tools/full_size_properties.py [cfg ...]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kalibr_b200 import capi, synthetic
from kalibr_b200.problem import KbOptimizerOptions
from oracle import ko_init as ki  # pose helpers only

for cfg in [int(a) for a in sys.argv[1:]] or [2, 3, 4]:
    p = synthetic.make_config(cfg)
    g = capi.B200SchurLinearSystemSolver(p)
    out = {"cfg": cfg, "terms": int(p.n_terms), "views": int(p.n_views)}
    # ---- weighting identities
    J0 = g.evaluate_error(); g.build_system(); g.set_constant_conditioner(10.0); dx0, ok0 = g.solve_system()
    e0 = g.error_vector()
    g.set_inv_r(np.eye(2) * 4.0)
    J1 = g.evaluate_error(); g.build_system(); g.set_constant_conditioner(20.0); dx1, ok1 = g.solve_system()  # damping scales with the weights too
    out["invR_cost_ratio_err"] = abs(J1 / J0 - 4.0)
    out["invR_dx_rel_err"] = float(np.abs(dx1 - dx0).max() / np.abs(dx0).max())
    out["invR_e_err"] = float(np.abs(g.error_vector() - 2.0 * e0).max() / np.abs(e0).max())
    g.set_inv_r(np.eye(2))
    g.set_m_estimator(capi.MEST_HUBER, 1e6)
    J2 = g.evaluate_error()
    out["huber_inactive_cost_err"] = abs(J2 - J0) / J0
    g.set_m_estimator(capi.MEST_CAUCHY, 25.0)
    J3 = g.evaluate_error()
    raw = (e0.reshape(-1, 2) ** 2).sum(1)
    out["cauchy_cost_rel_err"] = abs(J3 - float((raw / (1.0 + raw / 25.0)).sum())) / J3
    g.set_m_estimator(capi.MEST_NONE)
    st = g.reprojection_statistics()
    cam_of_term = np.repeat(p.view_cam, np.diff(p.view_begin))
    err = 0.0
    for k in range(p.n_cams):
        ek = -e0.reshape(-1, 2)[cam_of_term == k]
        ref = np.array([len(ek), *ek.mean(0), *ek.std(0, ddof=1), np.linalg.norm(ek.sum(0)) / np.sqrt(len(ek))])
        err = max(err, float(np.abs(st[k] - ref).max()))
    out["stats_abs_err"] = err
    # ---- initial-guess stage
    sol_ref, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
    g.reset_state()
    t = time.time(); n_failed = g.initialize_set_poses(); out["initialize_set_poses_s"] = round(time.time() - t, 4)
    sp = g.set_poses()
    d = max(np.abs(ki.pose_to_T(a) - ki.pose_to_T(b)).max() for a, b in zip(sp[::max(1, p.n_sets // 500)], p.set_poses[::max(1, p.n_sets // 500)]))
    out["init_failed_sets"] = int(n_failed)
    out["init_pose_vs_generator_guess"] = float(d)
    sol, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
    out["cost_after_init_vs_generator_start"] = abs(sol.j_final - sol_ref.j_final) / sol_ref.j_final
    out["iterations"] = [sol_ref.iterations, sol.iterations]
    T, okv = g.estimate_transformations()
    out["pnp_ok_fraction"] = float(okv.mean())
    good = (out["invR_cost_ratio_err"] < 1e-10 and out["invR_dx_rel_err"] < 1e-8 and out["invR_e_err"] < 1e-12 and out["huber_inactive_cost_err"] < 1e-12
            and out["cauchy_cost_rel_err"] < 1e-10 and out["stats_abs_err"] < 1e-9 and n_failed == 0 and d < 0.15 and out["cost_after_init_vs_generator_start"] < 1e-6
            and out["pnp_ok_fraction"] == 1.0)
    out["PASS"] = bool(good)
    print(json.dumps(out), flush=True)
    g.close()
