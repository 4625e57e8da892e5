"""Times the initial-guess stage (kb_estimate_transformations, kb_initialize_set_poses) on full-size configs: device time with CUDA
events on the library's stream.  python tools/pnp_timing.py [cfg ...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalibr_b200 import capi, synthetic  # noqa: E402

for cfg in [int(a) for a in sys.argv[1:]] or [3, 4]:
    p = synthetic.make_config(cfg)
    g = capi.B200SchurLinearSystemSolver(p)
    stream = torch.cuda.ExternalStream(g.cuda_stream())
    g.estimate_transformations()
    out = {"cfg": cfg, "views": int(p.n_views), "sets": int(p.n_sets)}
    for name, fn in (("estimate_transformations", g.estimate_transformations), ("initialize_set_poses", g.initialize_set_poses)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        out[name + "_ms"] = e0.elapsed_time(e1) / 3
    out["views_per_s"] = p.n_views / (out["estimate_transformations_ms"] * 1e-3)
    print(json.dumps(out), flush=True)
    g.close()
