"""ms per kb_iterate step (enqueue-only, one kb_wait at the end, no stage timing) and per kb_optimize iteration.
  python tools/iterate_timing.py cfg sets [steps]      (KB_NO_PDL=1: without programmatic dependent launch)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kalibr_b200 import capi, synthetic  # noqa: E402
from kalibr_b200.problem import KbOptimizerOptions  # noqa: E402

cfg, sets = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
p = synthetic.make_config(cfg, n_sets=sets)
g = capi.B200SchurLinearSystemSolver(p, device=0)
st = torch.cuda.ExternalStream(g.cuda_stream())
for _ in range(5):
    g.iterate(10.0, revert=True)
out = {"cfg": cfg, "sets": sets, "pdl": os.environ.get("KB_NO_PDL") is None}
for name, wait in (("iterate_async_ms", False), ("iterate_sync_ms", True)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(steps):
        g.iterate(10.0, revert=True, wait=wait)
    e1.record(st)
    g.wait_iterations()
    torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / steps, 4)
g.reset_state()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g.optimize(KbOptimizerOptions.kalibr2_default())  # warm: graph instantiation
g.reset_state()
torch.cuda.synchronize()
e0.record(st)
s, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
e1.record(st)
torch.cuda.synchronize()
out["optimize_ms"] = round(e0.elapsed_time(e1), 4)
out["optimize_iterations"] = s.iterations + s.failed_iterations
print(json.dumps(out), flush=True)
