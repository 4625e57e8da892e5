"""Cross-check of the tridiagonal-QL prototype (tools/sym_eig_tridiag.cu, serial reference) against numpy on REAL reduced camera
systems: the column-scaled Omega of the estimator's solver for several rigs.  What matters downstream: the singular values relative
to the largest one (rank decision at sv[0] * 1e-6 * n) and the log2 sum over the retained ones.
  python tools/sym_eig_check.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kalibr_b200 import synthetic  # noqa: E402
from oracle import ko_estimator as ke  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402

tool = os.path.join(ROOT, "tools", "sym_eig_tridiag")
if not os.path.exists(tool):
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tool, tool + ".cu"], check=True)
worst = 0.0
for cfg, S in [(1, 10), (2, 9), (3, 6), (3, 30), (4, 3), (4, 12), (6, 5), (7, 7)]:
    p = synthetic.make_config(cfg, n_sets=S)
    o = oa.OracleProblem(p)
    J, b = ke.system_of(o, p)
    cal, rest = ke.calibration_columns(p)
    for scaled in (True, False):
        A_r = J[:, cal] * (ke.column_scaling(J[:, cal], ke.EPS) if scaled else 1.0)
        Q, _ = np.linalg.qr(J[:, rest])
        ArtQ = A_r.T @ Q
        Om = A_r.T @ A_r - ArtQ @ ArtQ.T
        Om = 0.5 * (Om + Om.T)
        sv = np.linalg.svd(Om, compute_uv=False)
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            Om.astype(np.float64).tofile(f.name)
            out = subprocess.run([tool, f.name, str(len(sv))], capture_output=True, text=True, check=True)
        ev = np.sort(np.abs(np.array([float(x) for x in out.stdout.split()])))[::-1]
        err = np.abs(ev - sv).max() / sv[0]
        tol = sv[0] * (1e-6 if scaled else ke.EPS) * len(sv)
        rank = lambda v: len(v) - next((i for i, x in enumerate(v[::-1]) if x > tol), len(v) - 1) if v[-1] <= tol else len(v)
        r_np, r_tq = int((sv > tol).sum()), int((ev > tol).sum())
        l_np, l_tq = np.log2(sv[:r_np]).sum(), np.log2(ev[:r_np]).sum()
        worst = max(worst, err)
        print(f"cfg{cfg} S={S:2d} scaled={int(scaled)} n={len(sv):3d} max|ev - sv|/sv0 = {err:.2e}  rank numpy/tridiag = {r_np}/{r_tq}  "
              f"log2sum diff = {abs(l_np - l_tq):.2e}  {out.stderr.strip()}")
print("worst relative-to-largest deviation:", worst)
