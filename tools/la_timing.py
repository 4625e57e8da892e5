"""Times the fused linearise+assemble stage (kb_evaluate_error in its default speculative mode = set_prep + fused kernel + finalize_gram)
and a whole kb_iterate step for the default library and every experimental build under kalibr_b200/_exp/ (tools/build_variants.sh),
each in its own process.   python tools/la_timing.py [cfg] [sets]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch

    from kalibr_b200 import capi, synthetic

    cfg, sets = int(sys.argv[2]), int(sys.argv[3])
    p = synthetic.make_config(cfg, n_sets=sets)
    g = capi.B200SchurLinearSystemSolver(p, device=0)
    J0 = g.evaluate_error()
    g.enable_stage_timing(True)
    for _ in range(3):
        g.iterate(10.0, revert=True)
    g.enable_stage_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    torch.cuda.synchronize()
    ev0.record(torch.cuda.ExternalStream(g.cuda_stream()))
    for _ in range(n):
        g.iterate(10.0, revert=True)
    ev1.record(torch.cuda.ExternalStream(g.cuda_stream()))
    torch.cuda.synchronize()
    t = g.stage_totals()
    print(json.dumps({"lib": os.environ.get("KB_LIB_PATH", "default").split("/")[-1], "cfg": cfg, "sets": sets, "J0": J0, "step_ms": round(ev0.elapsed_time(ev1) / n, 4),
                      "stage_ms": {k: round(v[0] / v[1], 4) for k, v in t.items() if v[1] > 0}}), flush=True)
    sys.exit(0)

cfg = sys.argv[1] if len(sys.argv) > 1 else "4"
sets = sys.argv[2] if len(sys.argv) > 2 else "20000"
exp = os.path.join(ROOT, "kalibr_b200", "_exp")
libs = [None] + sorted(os.path.join(exp, f) for f in (os.listdir(exp) if os.path.isdir(exp) else []) if f.endswith(".so"))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["KB_LIB_PATH"] = lib
    else:
        env.pop("KB_LIB_PATH", None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", cfg, sets], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else f"{lib}: FAILED rc={r.returncode} {r.stderr[-400:]}", flush=True)
