"""Where one addBatch of the incremental estimator spends its time (wall clock through the Python binding): handle creation,
build, truncated-SVD solve, marginal analysis.  python tools/estimator_breakdown.py [cfg] [n_sets]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalibr_b200 import capi, synthetic  # noqa: E402
from kalibr_b200.problem import KbSvdSolverOptions  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = int(sys.argv[2]) if len(sys.argv) > 2 else 12
p = synthetic.make_config(cfg, n_sets=S)
capi.B200SchurLinearSystemSolver(p).close()


def timed(fn, reps=5):
    fn()
    t = time.time()
    for _ in range(reps):
        r = fn()
    return 1e3 * (time.time() - t) / reps, r


t_create, _ = timed(lambda: capi.B200SchurLinearSystemSolver(p).close())
g = capi.B200SchurLinearSystemSolver(p)
t_eval, _ = timed(g.evaluate_error)
t_build, _ = timed(g.build_system)
t_svd, _ = timed(lambda: g.solve_system_svd(KbSvdSolverOptions.kalibr2(), fetch_dx=False))
t_marg, _ = timed(lambda: g.analyze_marginal(last_build=True))
g.set_constant_conditioner(10.0)
t_chol, _ = timed(lambda: g.solve_system(fetch_dx=False))
print(json.dumps({"cfg": cfg, "sets": S, "n_c": int(p.n_c), "create_destroy_ms": round(t_create, 3), "evaluate_ms": round(t_eval, 3), "build_ms": round(t_build, 3),
                  "solve_svd_ms": round(t_svd, 3), "analyze_marginal_last_build_ms": round(t_marg, 3), "solve_cholesky_ms": round(t_chol, 3)}), flush=True)
