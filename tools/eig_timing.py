"""Device time of the dense eigen stage (sym_eig_kernel: Householder + QL; jacobi_polish_kernel) behind kb_solve_system_svd (column
scaled: QL only) and kb_analyze_marginal (QL + polish), CUDA events on the library's stream.  python tools/eig_timing.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalibr_b200 import capi, synthetic  # noqa: E402
from kalibr_b200.problem import KbSvdSolverOptions  # noqa: E402

for cfg, S in [(2, 40), (3, 30), (4, 12), (4, 200), (5, 12), (5, 100)]:
    p = synthetic.make_config(cfg, n_sets=S)
    g = capi.B200SchurLinearSystemSolver(p)
    g.evaluate_error(); g.build_system()
    g.solve_system_svd(KbSvdSolverOptions.kalibr2()); g.analyze_marginal(last_build=True)  # warm-up
    g.enable_stage_timing(True)
    t_solve, t_marg = [], []
    for _ in range(5):
        g.solve_system_svd(KbSvdSolverOptions.kalibr2())
        t_solve.append(g.stage_ms()["reduced_solve"])
        res, sv, V, cols = g.analyze_marginal(last_build=True)
        t_marg.append(g.stage_ms()["reduced_solve"])
    print(json.dumps({"cfg": cfg, "sets": S, "n_c": p.n_c, "svd_solve_scaled_ms": round(min(t_solve), 4), "analyze_marginal_eig_ms": round(min(t_marg), 4),
                      "rank": res.rank, "orthogonality": float(np.abs(V.T @ V - np.eye(p.n_c)).max())}), flush=True)
    g.close()
