"""Numerical experiment for DESIGN.md §7c (CPU, numpy): how many one-sided Jacobi sweeps does the marginal analysis need when it starts
from the eigenvectors of an absolutely-accurate method (tridiagonal QL / numpy eigh) instead of from the identity?
Reference values: one-sided Jacobi from the identity, run to convergence (the current device algorithm).
  python tools/jacobi_polish_experiment.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalibr_b200 import synthetic  # noqa: E402
from oracle import ko_estimator as ke  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402


def jacobi_sweeps(G, max_sweeps=40):
    """One-sided Jacobi on the columns of G (in place); returns (singular values sorted descending, sweeps used)."""
    n = G.shape[1]
    tol = n * np.finfo(float).eps
    for sweep in range(max_sweeps):
        rotated = False
        for a in range(n - 1):
            for b in range(a + 1, n):
                al, be, ga = G[:, a] @ G[:, a], G[:, b] @ G[:, b], G[:, a] @ G[:, b]
                if ga * ga > tol * tol * al * be and ga != 0.0:
                    d, g2 = be - al, 2.0 * ga
                    t = np.copysign(1.0, d) * g2 / (abs(d) + np.hypot(d, g2))
                    c = 1.0 / np.sqrt(1.0 + t * t)
                    s = c * t
                    x, y = G[:, a].copy(), G[:, b].copy()
                    G[:, a], G[:, b] = c * x - s * y, s * x + c * y
                    rotated = True
        if not rotated:
            return np.sort(np.linalg.norm(G, axis=0))[::-1], sweep
    return np.sort(np.linalg.norm(G, axis=0))[::-1], max_sweeps


for cfg, S in [(2, 9), (3, 6), (3, 30), (4, 3)]:
    p = synthetic.make_config(cfg, n_sets=S)
    o = oa.OracleProblem(p)
    J, _ = ke.system_of(o, p)
    cal, rest = ke.calibration_columns(p)
    Q, _ = np.linalg.qr(J[:, rest])
    A_r = J[:, cal]
    ArtQ = A_r.T @ Q
    Om = A_r.T @ A_r - ArtQ @ ArtQ.T
    Om = 0.5 * (Om + Om.T)
    ref, sweeps_ref = jacobi_sweeps(Om.copy())
    w, V = np.linalg.eigh(Om)
    sv_abs = np.sort(np.abs(w))[::-1]
    pol, sweeps_pol = jacobi_sweeps(Om @ V)
    keep = ref > ref[0] * 1e-13
    print(f"cfg{cfg} S={S:2d} n={len(ref):3d}: Jacobi from I: {sweeps_ref} sweeps | eigh alone: max rel err on sv > 1e-13 sv0 = "
          f"{np.abs(sv_abs[keep] / ref[keep] - 1).max():.1e} | eigh + Jacobi polish: {sweeps_pol} sweeps, max rel err = "
          f"{np.abs(pol[keep] / ref[keep] - 1).max():.1e}")
