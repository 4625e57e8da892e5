"""Jacobi sweep counts / QL usage per eigen-decomposition inside the incremental estimator loop (KB_SVD_TRACE of the library).
python tools/estimator_sweeps.py cfg n_sets"""
import collections, os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from driver_util import build_driver, write_problem
from kalibr_b200 import synthetic
cfg, S = int(sys.argv[1]), int(sys.argv[2])
p = synthetic.make_config(cfg, n_sets=S)
res = [synthetic.TRUTH_PARAMS[m][1] for m in p.cam_model]
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "p.bin")
    write_problem(path, p, res)
    env = dict(os.environ, KB_SVD_TRACE="1", **({"KB_EIG_NO_WARM": "1"} if "--no-warm" in sys.argv else {}))
    r = subprocess.run([build_driver(), "estimator", path, "0.2"], capture_output=True, text=True, env=env)
solve, marg = collections.Counter(), collections.Counter()
for ln in r.stderr.splitlines():
    m = re.search(r"(truncated-SVD solve|marginal analysis): n = (\d+), Jacobi polish sweeps = (\d+), QL failed = (\d+).*?(\d+) QL steps", ln)
    if m:
        (solve if m.group(1).startswith("trunc") else marg)[(int(m.group(3)), "QL" if int(m.group(5)) else "warm")] += 1
bm = [float(x) for x in next(l for l in r.stdout.splitlines() if l.startswith("batch_ms")).split()[1:]]
print({"cfg": cfg, "sets": S, "warm": "--no-warm" not in sys.argv, "median_ms_per_batch": sorted(bm[1:])[len(bm[1:]) // 2],
       "solve (sweeps, start) -> count": dict(sorted(solve.items())), "marginal": dict(sorted(marg.items()))})
