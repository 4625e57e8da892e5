"""Golden vectors of the decision numerics of the incremental estimator's solver -> tests/golden/reference_linalg_golden.npz.

Needs /root/reference (read-only) in the build container: oracle/ref_pin_linalg.cpp compiles IC/src/algorithms/linalg.cpp from its source
(stand-ins: oracle/ref_shim_linalg/, oracle/ref_shim/Eigen) and this script records what ITS rankTol / estimateNumericalRank / svGap,
colNorm / columnScalingMatrix, qrTol and analyzeSVD + solveSVD return: on spectra with clear and with marginal gaps, all-tiny and all-zero
spectra, fixed tolerances; on matrices with tiny and zero columns; on well-conditioned and rank-deficient symmetric systems.  The SVD inside
analyzeSVD is a stand-in (one-sided Jacobi for Eigen::JacobiSVD) - the singular values agree with LAPACK's to rounding; what the fixture
pins is everything decided FROM them.
    python tests/golden/make_reference_linalg_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_api as oa  # noqa: E402

EPS = float(np.finfo(float).eps)


def main():
    assert oa.reference_linalg() is not None, "needs the reference tree"
    rng = np.random.default_rng(77)
    out = {}
    # 1. rank decisions: (spectrum, eps, fixed tolerance or -1)
    spectra = []
    for t in range(60):
        n = int(rng.integers(1, 24))
        sv = np.sort(10.0 ** rng.uniform(-17, 6, n))[::-1]
        if t % 7 == 0:
            sv[n // 2:] = 0.0
        if t % 11 == 0:
            sv[:] = 0.0
        if t % 13 == 0:
            sv[:] = sv[0] * 1e-18
        spectra.append((sv, [EPS, 1e-6, 1e-9][t % 3], [-1.0, -1.0, 1e-3][t % 3] if t % 5 == 0 else -1.0))
    out["rank_n"] = np.array([len(s[0]) for s in spectra])
    out["rank_sv"] = np.concatenate([s[0] for s in spectra])
    out["rank_eps"] = np.array([s[1] for s in spectra])
    out["rank_tol_in"] = np.array([s[2] for s in spectra])
    out["rank_out"] = np.array([oa.reference_linalg_rank(*s) for s in spectra])  # tolerance, rank, gap
    # 2. column scaling and qrTol
    for t in range(8):
        m, n = int(rng.integers(5, 60)), int(rng.integers(1, 12))
        A = rng.standard_normal((m, n)) * 10.0 ** rng.uniform(-3, 4, n)
        if t % 2 == 0:
            A[:, rng.integers(0, n)] = 0.0
        if t % 3 == 0:
            A[:, rng.integers(0, n)] *= 1e-12
        eps = [EPS, 1e-8][t % 2]
        G, qr = oa.reference_linalg_column_scaling(A, eps, EPS)
        out[f"scale{t}_A"], out[f"scale{t}_eps"], out[f"scale{t}_G"], out[f"scale{t}_qr_tol"] = A, np.array(eps), G, np.array(qr)
    out["scale_count"] = np.array(8)
    # 3. analyzeSVD + rank decision + solveSVD on symmetric positive semi-definite systems
    for t in range(10):
        n = int(rng.integers(3, 20))
        r = n if t % 2 == 0 else int(rng.integers(1, n))  # exact rank; the kept part of the spectrum is well conditioned, so that x is
        M = rng.standard_normal((n, r)) * 10.0 ** rng.uniform(-0.5, 0.5, r)  # determined to rounding whichever SVD algorithm runs
        Omega = M @ M.T + (0.5 * np.eye(n) if r == n else 0.0)
        b = Omega @ rng.standard_normal(n) if t % 4 else rng.standard_normal(n)
        eps, tol = (EPS if r == n else 1e-6), -1.0  # a rank-deficient Gram matrix has trailing singular values of a few eps: cut well above
        x, sv, tolerance, rank, gap = oa.reference_linalg_svd_solve(Omega, b, eps, tol)
        out[f"solve{t}_Omega"], out[f"solve{t}_b"], out[f"solve{t}_eps"] = Omega, b, np.array(eps)
        out[f"solve{t}_x"], out[f"solve{t}_sv"], out[f"solve{t}_out"] = x, sv, np.array([tolerance, rank, gap])
        print("system", t, "n", n, "exact rank", r, "numerical rank", rank)
    out["solve_count"] = np.array(10)
    path = os.path.join(ROOT, "tests", "golden", "reference_linalg_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
