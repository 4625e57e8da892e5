"""Full-size runs of the BASELINE configs on the GPU: LM convergence properties (no oracle at these sizes) + timings."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from kalibr_b200 import synthetic, capi
from kalibr_b200.problem import KbOptimizerOptions, MODEL_P

cfgs = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5]
for cfg in cfgs:
    t = time.time(); p = synthetic.make_config(cfg); tg = time.time() - t
    t = time.time(); g = capi.B200SchurLinearSystemSolver(p); tc = time.time() - t
    g.enable_stage_timing(True)
    t = time.time(); sol, tr = g.optimize(KbOptimizerOptions.kalibr2_default()); to = time.time() - t
    dof = 2 * p.n_terms - g.jcols
    cam = g.camera_params(); truth = p.truth["cam_params"]
    err_f = max(np.abs(cam[k, MODEL_P[m] - 4:MODEL_P[m]] - truth[k, MODEL_P[m] - 4:MODEL_P[m]]).max() for k, m in enumerate(p.cam_model))
    tot = g.stage_totals()
    print(json.dumps({"cfg": cfg, "terms": p.n_terms, "sets": p.n_sets, "n_c": p.n_c, "gen_s": round(tg, 2), "create_s": round(tc, 2),
                      "optimize_s": round(to, 4), "iterations": sol.iterations, "failed": sol.failed_iterations, "solver_failure": sol.linear_solver_failure,
                      "j_start": sol.j_start, "j_final": sol.j_final, "chi2_over_dof_sigma2": sol.j_final / (dof * 0.09),
                      "max_focal_centre_err_px": float(err_f), "ms_per_iteration": 1e3 * to / max(sol.iterations, 1),
                      "stage_ms_avg": {k: round(v[0] / max(v[1], 1), 4) for k, v in tot.items() if v[1]}, "invalid_terms": g.num_invalid_terms()}), flush=True)
    g.close()
