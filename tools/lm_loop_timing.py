"""Wall time of kb_optimize per LM iteration, device-resident loop vs host-driven loop, after a warm-up run (full-size configs)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalibr_b200 import synthetic, capi
from kalibr_b200.problem import KbOptimizerOptions

cfgs = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5]
for cfg in cfgs:
    p = synthetic.make_config(cfg)
    g = capi.B200SchurLinearSystemSolver(p)
    row = {"cfg": cfg, "terms": p.n_terms}
    for name, dl in (("device_loop", 1), ("host_loop", 0)):
        opt = KbOptimizerOptions.kalibr2_default(device_loop=dl)
        best = None
        for rep in range(4):
            g.reset_state()
            t0 = time.perf_counter()
            sol, _ = g.optimize(opt)
            dt = time.perf_counter() - t0
            if rep > 0:
                best = dt if best is None else min(best, dt)
        row[name] = {"iterations": sol.iterations, "optimize_ms": round(1e3 * best, 4), "ms_per_iteration": round(1e3 * best / max(sol.iterations, 1), 4)}
    print(json.dumps(row), flush=True)
    g.close()
