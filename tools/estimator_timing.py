"""Wall time of the incremental estimator loop on the device (C++ mirror through tests/cpp/driver_main) and of the dense numpy
oracle on the host for the same batch sequence.  python tools/estimator_timing.py [cfg] [n_sets]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from driver_util import run_driver, write_problem  # noqa: E402
from kalibr_b200 import synthetic  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 40
p = synthetic.make_config(cfg, n_sets=S)
res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "p.bin")
    write_problem(path, p, res)
    run_driver("estimator", path, 0.2)  # warm-up (context creation, module load)
    t = time.time()
    code, out = run_driver("estimator", path, 0.2)
    t_dev = time.time() - t
acc = int(out["accepted"][0])
bm = np.asarray(out["batch_ms"])
line = {"cfg": cfg, "batches_offered": S, "accepted": acc, "process_wall_s": round(t_dev, 3), "addBatch_loop_ms": float(out["loop_ms"][0]),
        "first_batch_ms_incl_cuda_context": round(float(bm[0]), 2), "median_ms_per_batch": round(float(np.median(bm[1:])), 3),
        "ms_per_batch_at": {str(i): round(float(bm[i]), 3) for i in sorted({1, len(bm) // 4, len(bm) // 2, len(bm) - 1})}}
if "--oracle" in sys.argv:
    from oracle import ko_estimator as ke
    from oracle import oracle_api as oa

    est = ke.OracleIncrementalEstimator(oa, p.cam_model, p.cam_params, p.baselines, p.target_points, check_validity=True)
    t = time.time()
    n = min(S, 12)
    for s in range(n):
        batch = {}
        for w in np.flatnonzero(p.view_set == s):
            b, e = p.view_begin[w], p.view_begin[w + 1]
            batch[int(p.view_cam[w])] = (p.corner_id[b:e], p.y_u[b:e], p.y_v[b:e])
        est.add_batch(batch, p.set_poses[s])
    line["oracle_dense_numpy_ms_per_batch_first_%d" % n] = round(1e3 * (time.time() - t) / n, 1)
print(json.dumps(line), flush=True)
