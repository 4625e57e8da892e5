"""The decision numerics of the incremental estimator's solver against the REFERENCE's own code.

tests/golden/reference_linalg_golden.npz (generator: tests/golden/make_reference_linalg_golden.py) holds what the functions of
IC/src/algorithms/linalg.cpp return - compiled from that source (oracle/ref_pin_linalg.cpp; stand-ins: oracle/ref_shim_linalg/, the shim
Eigen): rankTol / estimateNumericalRank / svGap (numerical rank, its tolerance, its gap), colNorm / columnScalingMatrix (the column
scaling), qrTol, analyzeSVD + solveSVD (the truncated solve; the SVD inside analyzeSVD is a stand-in, everything decided from the singular
values is reference code).  The oracle's restatements in oracle/ko_estimator.py (numpy) and oracle/ko_marginal.hpp (C++) are held to them.
Still unpinned: the sparse QR elimination that produces Omega (SuiteSparseQR is not in the image)."""
import os

import numpy as np
import pytest

from oracle import ko_estimator as ke
from oracle import oracle_api as oa

LINALG_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_linalg_golden.npz")


def spectra(g):
    at = 0
    for n, eps, tol_in, ref in zip(g["rank_n"], g["rank_eps"], g["rank_tol_in"], g["rank_out"]):
        yield g["rank_sv"][at:at + n], float(eps), float(tol_in), ref
        at += n


def same_or_both_nan(a, b, rtol=0.0):
    return (np.isnan(a) and np.isnan(b)) or a == b or abs(a - b) <= rtol * abs(b)


def test_numerical_rank_tolerance_and_gap_are_the_references():
    g = np.load(LINALG_GOLD)
    n_deficient = 0
    for sv, eps, tol_in, (tol, rank, gap) in spectra(g):
        t, r, gp = ke.numerical_rank(sv, eps, tol_in)
        assert r == int(rank) and same_or_both_nan(t, tol, 1e-15) and same_or_both_nan(gp, gap, 1e-15), (sv, eps, tol_in)
        n_deficient += int(rank) < len(sv)
    assert n_deficient >= 20  # the fixture really cuts
    # the two edge cases the reference's loop has: a spectrum entirely below the tolerance still has rank 1, and its gap can be 0 / 0
    assert ke.numerical_rank(np.zeros(3))[1] == 1 and np.isnan(ke.numerical_rank(np.zeros(3))[2])


def test_column_scaling_is_the_references():
    g = np.load(LINALG_GOLD)
    zeros = 0
    for t in range(int(g["scale_count"])):
        A, eps, G = g[f"scale{t}_A"], float(g[f"scale{t}_eps"]), g[f"scale{t}_G"]
        mine = ke.column_scaling(A, eps)
        assert np.array_equal(mine == 0.0, G == 0.0)  # the same columns are switched off
        assert np.abs(mine - G).max() <= 1e-14 * np.abs(G).max()
        zeros += int((G == 0.0).sum())
        # qrTol = 20 (m + n) eps max column norm (linalg.cpp:263-272)
        assert abs(20.0 * sum(A.shape) * np.finfo(float).eps * np.sqrt((A * A).sum(0)).max() - float(g[f"scale{t}_qr_tol"])) <= 1e-14 * float(g[f"scale{t}_qr_tol"])
    assert zeros >= 4


def test_truncated_svd_solve_is_the_references():
    g = np.load(LINALG_GOLD)
    for t in range(int(g["solve_count"])):
        Omega, b, eps = g[f"solve{t}_Omega"], g[f"solve{t}_b"], float(g[f"solve{t}_eps"])
        x, sv, tol, rank, gap = ke.svd_truncated_solve(Omega, b, eps)
        rtol, rrank, rgap = g[f"solve{t}_out"]
        assert rank == int(rrank) and abs(tol - rtol) <= 1e-12 * rtol
        assert np.abs(sv[:rank] - g[f"solve{t}_sv"][:rank]).max() <= 1e-10 * sv[0]  # LAPACK vs the one-sided Jacobi stand-in
        assert np.abs(x - g[f"solve{t}_x"]).max() <= 1e-9 * np.abs(g[f"solve{t}_x"]).max()  # the kept spectrum is well conditioned by construction
        if rank < len(sv):
            assert np.isfinite(gap) and gap > 1.0


@pytest.mark.parametrize("cfg,n_sets", [(2, 6), (8, 5)])
def test_marginal_analysis_decides_rank_as_the_reference(oracle_lib, cfg, n_sets):
    """oracle/ko_marginal.hpp (C++): rank, tolerance and gap of analyze_marginal follow from its singular values exactly as the pinned
    numerical_rank says - on a full-rank problem and on one with a camera that has no observations (rank deficient)"""
    from kalibr_b200 import synthetic
    from test_estimator_cpu import without_camera

    p = synthetic.make_config(cfg, n_sets=n_sets)
    for problem in (p, without_camera(p, p.n_cams - 1)):
        o = oa.OracleProblem(problem)
        res, sv = o.analyze_marginal()[:2]
        tol, rank, gap = ke.numerical_rank(sv, float(np.finfo(float).eps), -1.0)
        assert res.rank == rank and abs(res.tolerance - tol) <= 1e-15 * tol and same_or_both_nan(res.sv_gap, gap, 1e-12)


def test_linalg_fixture_is_what_the_reference_returns_now():
    """build container only"""
    if oa.reference_linalg() is None:
        pytest.skip("no reference tree and no prebuilt oracle/_ref here")
    g = np.load(LINALG_GOLD)
    for sv, eps, tol_in, ref in spectra(g):
        assert np.array_equal(np.array(oa.reference_linalg_rank(sv, eps, tol_in)), ref, equal_nan=True)
    for t in range(int(g["scale_count"])):
        G, qr = oa.reference_linalg_column_scaling(g[f"scale{t}_A"], float(g[f"scale{t}_eps"]), float(np.finfo(float).eps))
        assert np.array_equal(G, g[f"scale{t}_G"]) and qr == float(g[f"scale{t}_qr_tol"])
    for t in range(int(g["solve_count"])):
        x, sv, tol, rank, gap = oa.reference_linalg_svd_solve(g[f"solve{t}_Omega"], g[f"solve{t}_b"], float(g[f"solve{t}_eps"]))
        assert np.array_equal(x, g[f"solve{t}_x"]) and np.array_equal(sv, g[f"solve{t}_sv"])
