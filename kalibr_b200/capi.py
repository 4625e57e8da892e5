"""ctypes binding of the C ABI (include/kalibr_b200.h) exported by kalibr_b200/libkalibr_b200.so.

The library is the product: if it is missing this module raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB_PATH
from .problem import KbOptimizerOptions, KbProblemDesc, KbSolution, Problem

KB_OK = 0
KB_NUM_STAGES = 8
STAGE_NAMES = ["evaluate", "linearise_assemble", "expand", "schur", "reduced_solve", "backsub", "update", "linearise_materialise"]

# every symbol include/kalibr_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "kb_create", "kb_destroy", "kb_last_error", "kb_nccl_unique_id", "kb_jrows", "kb_jcols", "kb_local_jrows",
    "kb_num_design_variables", "kb_get_dv_layout", "kb_evaluate_error", "kb_build_system", "kb_set_constant_conditioner",
    "kb_solve_system", "kb_lm_rho_denominator", "kb_apply_state_update", "kb_revert_last_state_update",
    "kb_default_optimizer_options", "kb_optimize", "kb_get_trace", "kb_set_solver_semantic", "kb_set_speculative_linearise", "kb_get_error_vector", "kb_get_rhs",
    "kb_linearise", "kb_jacobian_nnz", "kb_get_jacobian_ccs", "kb_get_hessian_blocks", "kb_get_camera_params", "kb_get_baselines",
    "kb_get_set_poses", "kb_set_observations", "kb_evaluate_error_streamed", "kb_prefetch_observations", "kb_commit_observations", "kb_peer_exchange_handle", "kb_attach_peers", "kb_default_marginal_options", "kb_analyze_marginal", "kb_num_invalid_terms", "kb_reset_state", "kb_kernel_launches", "kb_get_stage_ms",
    "kb_enable_stage_timing", "kb_get_stage_totals", "kb_cuda_stream",
    "kb_set_inv_r", "kb_get_sqrt_inv_r", "kb_set_m_estimator", "kb_m_estimator_parameter", "kb_reprojection_statistics",
    "kb_iterate", "kb_wait", "kb_set_observations_f32", "kb_evaluate_error_streamed_f32", "kb_prefetch_observations_f32",
    "kb_append_set", "kb_remove_last_set", "kb_save_design_variables", "kb_restore_design_variables",
    "kb_set_state", "kb_set_camera_params", "kb_set_baselines", "kb_set_set_poses", "kb_set_conditioner",
    "kb_estimate_transformations", "kb_initialize_set_poses", "kb_estimate_stereo_baseline", "kb_initialize_intrinsics", "kb_default_svd_solver_options", "kb_solve_system_svd", "kb_optimize_gauss_newton", "kb_analyze_marginal_last_build", "kb_get_last_svd_solve", "kb_get_last_svd_decomposition",
]

MEST_NONE, MEST_HUBER, MEST_CAUCHY, MEST_GEMAN_MCCLURE, MEST_BLAKE_ZISSERMAN = range(5)  # = kb_m_estimator
REPROJ_STAT_STRIDE = 6

_lib = None


class KalibrB200Error(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """Load the CUDA extension; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    lib_path = os.environ.get("KB_LIB_PATH", LIB_PATH)  # kernel-variant experiments (tools/) point this at another build
    if not os.path.exists(lib_path):
        raise KalibrB200Error(
            f"{lib_path} is missing: build it with `python -m kalibr_b200.build` (nvcc, sm_100a). "
            "The B200 hot path has no CPU fallback."
        )
    L = C.CDLL(lib_path)
    vp = C.c_void_p
    L.kb_create.argtypes = [C.POINTER(KbProblemDesc), C.POINTER(vp)]
    L.kb_create.restype = C.c_int32
    L.kb_destroy.argtypes = [vp]
    L.kb_destroy.restype = None
    L.kb_last_error.argtypes = [vp]
    L.kb_last_error.restype = C.c_char_p
    L.kb_nccl_unique_id.argtypes = [C.c_char_p]
    L.kb_nccl_unique_id.restype = C.c_int32
    for name in ("kb_jrows", "kb_jcols", "kb_local_jrows", "kb_jacobian_nnz", "kb_kernel_launches", "kb_num_invalid_terms"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int64
    L.kb_num_design_variables.argtypes = [vp]
    L.kb_num_design_variables.restype = C.c_int32
    L.kb_get_dv_layout.argtypes = [vp, vp, vp]
    L.kb_evaluate_error.argtypes = [vp, C.c_int32, C.POINTER(C.c_double)]
    L.kb_build_system.argtypes = [vp, C.c_int32]
    L.kb_set_constant_conditioner.argtypes = [vp, C.c_double]
    L.kb_solve_system.argtypes = [vp, vp, C.c_int32, C.POINTER(C.c_int32)]
    L.kb_lm_rho_denominator.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
    L.kb_apply_state_update.argtypes = [vp, C.POINTER(C.c_double)]
    L.kb_revert_last_state_update.argtypes = [vp]
    L.kb_default_optimizer_options.argtypes = [C.POINTER(KbOptimizerOptions)]
    L.kb_default_optimizer_options.restype = None
    L.kb_optimize.argtypes = [vp, C.POINTER(KbOptimizerOptions), C.POINTER(KbSolution)]
    L.kb_get_trace.argtypes = [vp, vp, C.c_int32]
    L.kb_get_trace.restype = C.c_int32
    L.kb_set_solver_semantic.argtypes = [vp, C.c_int32]
    L.kb_set_speculative_linearise.argtypes = [vp, C.c_int32]
    L.kb_get_error_vector.argtypes = [vp, vp]
    L.kb_get_rhs.argtypes = [vp, vp]
    L.kb_linearise.argtypes = [vp]
    L.kb_get_jacobian_ccs.argtypes = [vp, vp, vp, vp]
    L.kb_get_hessian_blocks.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, vp, vp, vp]
    L.kb_get_camera_params.argtypes = [vp, vp]
    L.kb_get_baselines.argtypes = [vp, vp]
    L.kb_get_set_poses.argtypes = [vp, vp]
    L.kb_iterate.argtypes = [vp, C.c_double, C.c_int32, C.c_int32, vp]
    L.kb_wait.argtypes = [vp, vp]
    L.kb_append_set.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.kb_remove_last_set.argtypes = [vp]
    L.kb_save_design_variables.argtypes = [vp]
    L.kb_restore_design_variables.argtypes = [vp]
    L.kb_set_state.argtypes = [vp, vp, vp, vp]
    L.kb_set_camera_params.argtypes = [vp, vp]
    L.kb_set_baselines.argtypes = [vp, vp]
    L.kb_set_set_poses.argtypes = [vp, vp]
    L.kb_set_conditioner.argtypes = [vp, vp]
    L.kb_set_observations.argtypes = [vp, vp, vp]
    L.kb_evaluate_error_streamed.argtypes = [vp, vp, vp, C.c_int32, vp]
    L.kb_prefetch_observations.argtypes = [vp, vp, vp]
    L.kb_commit_observations.argtypes = [vp]
    L.kb_set_observations_f32.argtypes = [vp, vp, vp]
    L.kb_evaluate_error_streamed_f32.argtypes = [vp, vp, vp, C.c_int32, vp]
    L.kb_prefetch_observations_f32.argtypes = [vp, vp, vp]
    L.kb_default_marginal_options.argtypes = [vp]
    L.kb_default_marginal_options.restype = None
    L.kb_analyze_marginal.argtypes = [vp, vp, vp, vp, vp, vp]
    L.kb_peer_exchange_handle.argtypes = [vp, C.c_char_p]
    L.kb_attach_peers.argtypes = [vp, C.c_char_p]
    L.kb_reset_state.argtypes = [vp]
    L.kb_get_stage_ms.argtypes = [vp, vp]
    L.kb_enable_stage_timing.argtypes = [vp, C.c_int32]
    L.kb_get_stage_totals.argtypes = [vp, vp, vp]
    L.kb_cuda_stream.argtypes = [vp]
    L.kb_cuda_stream.restype = vp
    L.kb_set_inv_r.argtypes = [vp, vp]
    L.kb_get_sqrt_inv_r.argtypes = [vp, vp]
    L.kb_set_m_estimator.argtypes = [vp, C.c_int32, C.c_double, C.c_double, C.c_double]
    L.kb_m_estimator_parameter.argtypes = [vp]
    L.kb_m_estimator_parameter.restype = C.c_double
    L.kb_reprojection_statistics.argtypes = [vp, vp]
    L.kb_estimate_transformations.argtypes = [vp, vp, vp, vp]
    L.kb_initialize_set_poses.argtypes = [vp, vp, C.POINTER(C.c_int32)]
    L.kb_estimate_stereo_baseline.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, C.POINTER(C.c_int32)]
    L.kb_default_svd_solver_options.argtypes = [vp]
    L.kb_default_svd_solver_options.restype = None
    L.kb_solve_system_svd.argtypes = [vp, vp, vp, C.c_int32, vp, vp]
    L.kb_optimize_gauss_newton.argtypes = [vp, vp, vp, vp]
    L.kb_analyze_marginal_last_build.argtypes = [vp, vp, vp, vp, vp, vp]
    L.kb_get_last_svd_solve.argtypes = [vp, vp]
    L.kb_get_last_svd_decomposition.argtypes = [vp, vp, vp, vp]
    L.kb_initialize_intrinsics.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_double, vp, C.POINTER(C.c_int32)]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:  # default: status-returning entry points
            fn.restype = C.c_int32
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    st = load_library().kb_nccl_unique_id(buf)
    if st != KB_OK:
        raise KalibrB200Error("kb_nccl_unique_id: " + load_library().kb_last_error(None).decode())
    return buf.raw


class _IterationResult(C.Structure):
    _fields_ = [("cost", C.c_double), ("rho_denominator", C.c_double), ("max_abs_dx", C.c_double), ("pos_def", C.c_int32)]


class B200SchurLinearSystemSolver:
    """Python face of the reference-facing solver plugin (≙ aslam::backend::LinearSystemSolver,
    aslam_optimizer/aslam_backend/include/aslam/backend/LinearSystemSolver.hpp:16-109) over the C ABI.

    Method names follow the reference (camelCase in C++, snake_case here); all compute runs in the CUDA library.
    """

    def __init__(self, problem: Problem, n_ranks: int = 1, rank: int = 0, nccl_id: bytes | None = None, device: int = 0,
                 n_sets_total: int = 0, set_offset: int = 0, n_terms_total: int = 0):
        self._L = load_library()
        self.problem = problem
        self._desc = problem.desc(n_ranks=n_ranks, rank=rank, nccl_id=nccl_id, device=device,
                                  n_sets_total=n_sets_total, set_offset=set_offset, n_terms_total=n_terms_total)
        h = C.c_void_p()
        st = self._L.kb_create(C.byref(self._desc), C.byref(h))
        if st != KB_OK:
            raise KalibrB200Error(f"kb_create failed ({st}): " + self._L.kb_last_error(None).decode())
        self._h = h
        self._local_views_known = n_ranks == 1 or n_sets_total > 0  # the description's views are exactly this rank's views
        self.jrows = self._L.kb_jrows(h)
        self.local_jrows = self._L.kb_local_jrows(h)
        self.jcols = self._L.kb_jcols(h)
        self.n_dv = self._L.kb_num_design_variables(h)

    def name(self) -> str:
        return "b200_schur"

    # -- life cycle
    def close(self):
        if getattr(self, "_h", None):
            self._L.kb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int, what: str):
        if st != KB_OK:
            raise KalibrB200Error(f"{what} failed ({st}): " + self._L.kb_last_error(self._h).decode())

    # -- LinearSystemSolver surface
    def dv_layout(self):
        col = np.zeros(self.n_dv, np.int32)
        dims = np.zeros(self.n_dv, np.int32)
        self._check(self._L.kb_get_dv_layout(self._h, _p(col), _p(dims)), "kb_get_dv_layout")
        return col, dims

    def evaluate_error(self, use_m_estimator: bool = True) -> float:
        out = C.c_double()
        self._check(self._L.kb_evaluate_error(self._h, 1 if use_m_estimator else 0, C.byref(out)), "kb_evaluate_error")
        return out.value

    def build_system(self, use_m_estimator: bool = True):
        self._check(self._L.kb_build_system(self._h, 1 if use_m_estimator else 0), "kb_build_system")

    def set_constant_conditioner(self, lam: float):
        self._check(self._L.kb_set_constant_conditioner(self._h, lam), "kb_set_constant_conditioner")

    def solve_system(self, fetch_dx: bool = True, gather: bool = True):
        pd = C.c_int32()
        dx = np.zeros(self.jcols) if fetch_dx else None
        self._check(self._L.kb_solve_system(self._h, _p(dx), 1 if gather else 0, C.byref(pd)), "kb_solve_system")
        return dx, bool(pd.value)

    def solve_system_svd(self, options=None, fetch_dx: bool = True, gather: bool = True):
        """≙ aslam::calibration::LinearSolver::solveSystem: undamped step, calibration block by truncated SVD.
        Returns (dx, KbSvdSolveResult, singular values of the (scaled) reduced system)."""
        from .problem import KbSvdSolveResult, KbSvdSolverOptions

        o = options or KbSvdSolverOptions.default()
        res = KbSvdSolveResult()
        dx = np.zeros(self.jcols) if fetch_dx else None
        sv = np.zeros(self.problem.n_c)
        self._check(self._L.kb_solve_system_svd(self._h, C.byref(o), _p(dx), 1 if gather else 0, C.byref(res), _p(sv)), "kb_solve_system_svd")
        return dx, res, sv

    def iterate(self, lam: float, use_m_estimator: bool = True, revert: bool = False, wait: bool = True):
        """evaluate + build + solve(lam) + apply [+ revert] with one host synchronisation: (cost, rho denominator, max|dx|, pos_def).
        wait=False only enqueues the iteration (several can be in flight); wait_iterations() then returns the last one's scalars."""
        if not wait:
            self._check(self._L.kb_iterate(self._h, lam, 1 if use_m_estimator else 0, 1 if revert else 0, None), "kb_iterate")
            return None
        out = _IterationResult()
        self._check(self._L.kb_iterate(self._h, lam, 1 if use_m_estimator else 0, 1 if revert else 0, C.byref(out)), "kb_iterate")
        return out.cost, out.rho_denominator, out.max_abs_dx, bool(out.pos_def)

    def wait_iterations(self):
        out = _IterationResult()
        self._check(self._L.kb_wait(self._h, C.byref(out)), "kb_wait")
        return out.cost, out.rho_denominator, out.max_abs_dx, bool(out.pos_def)

    def lm_rho_denominator(self, lam: float) -> float:
        out = C.c_double()
        self._check(self._L.kb_lm_rho_denominator(self._h, lam, C.byref(out)), "kb_lm_rho_denominator")
        return out.value

    def apply_state_update(self) -> float:
        out = C.c_double()
        self._check(self._L.kb_apply_state_update(self._h, C.byref(out)), "kb_apply_state_update")
        return out.value

    def revert_last_state_update(self):
        self._check(self._L.kb_revert_last_state_update(self._h), "kb_revert_last_state_update")

    # -- weighting of the terms (≙ ErrorTermFs::setInvR, ErrorTerm::setMEstimatorPolicy on every term)
    def set_inv_r(self, inv_r):
        a = np.ascontiguousarray(inv_r, np.float64).reshape(2, 2)
        self._check(self._L.kb_set_inv_r(self._h, _p(a)), "kb_set_inv_r")

    def sqrt_inv_r(self) -> np.ndarray:
        out = np.zeros((2, 2))
        self._check(self._L.kb_get_sqrt_inv_r(self._h, _p(out)), "kb_get_sqrt_inv_r")
        return out

    def set_m_estimator(self, kind: int, p0: float = 0.0, p1: float = 0.999, p2: float = 0.1) -> float:
        """Installs the policy; returns its parameter (k, sigma^2, or the derived epsilon of Blake-Zisserman)."""
        self._check(self._L.kb_set_m_estimator(self._h, kind, p0, p1, p2), "kb_set_m_estimator")
        return float(self._L.kb_m_estimator_parameter(self._h))

    def reprojection_statistics(self) -> np.ndarray:
        """≙ CameraCalibrator::PrintReprojectionErrorStatistics per camera: rows [n, mean_u, mean_v, std_u, std_v, rmse]."""
        out = np.zeros((self.problem.n_cams, REPROJ_STAT_STRIDE))
        self._check(self._L.kb_reprojection_statistics(self._h, _p(out)), "kb_reprojection_statistics")
        return out

    # -- initial-guess stage (≙ CameraGeometry::estimateTransformation, getTargetPoseGuess, CalibrateStereoPair's baseline guess)
    @staticmethod
    def _res(resolution):
        return None if resolution is None else np.ascontiguousarray(resolution, np.int32)

    def estimate_transformations(self, resolution=None):
        """PnP per local view: (T_target_camera poses [n_views, 7], ok [n_views])."""
        if not self._local_views_known:
            raise KalibrB200Error("estimate_transformations: use one rank or a pre-sharded problem from Python")
        nv = len(self.problem.view_set)
        T = np.zeros((nv, 7))
        ok = np.zeros(nv, np.int32)
        r = self._res(resolution)
        self._check(self._L.kb_estimate_transformations(self._h, _p(r), _p(T), _p(ok)), "kb_estimate_transformations")
        return T, ok.astype(bool)

    def initialize_set_poses(self, resolution=None) -> int:
        """getTargetPoseGuess for every set; returns the number of sets without a usable PnP."""
        nf = C.c_int32()
        r = self._res(resolution)
        self._check(self._L.kb_initialize_set_poses(self._h, _p(r), C.byref(nf)), "kb_initialize_set_poses")
        return int(nf.value)

    def initialize_intrinsics(self, cam: int, rows: int, cols: int, resolution, fallback_focal_length: float = 0.0):
        """≙ CameraGeometry::initializeIntrinsics for one camera: (params[10], success); the guess becomes the camera's state."""
        out = np.zeros(10)
        ok = C.c_int32()
        r = self._res(resolution)
        self._check(self._L.kb_initialize_intrinsics(self._h, cam, rows, cols, _p(r), fallback_focal_length, _p(out), C.byref(ok)),
                    "kb_initialize_intrinsics")
        return out, bool(ok.value)

    def estimate_stereo_baseline(self, cam_l: int = 0, cam_h: int = 1, resolution=None):
        out = np.zeros(7)
        n = C.c_int32()
        r = self._res(resolution)
        self._check(self._L.kb_estimate_stereo_baseline(self._h, _p(r), cam_l, cam_h, _p(out), C.byref(n)), "kb_estimate_stereo_baseline")
        return out, int(n.value)

    def set_solver_semantic(self, semantic: int):
        self._check(self._L.kb_set_solver_semantic(self._h, semantic), "kb_set_solver_semantic")

    def set_speculative_linearise(self, on: bool):
        self._check(self._L.kb_set_speculative_linearise(self._h, 1 if on else 0), "kb_set_speculative_linearise")

    def optimize(self, options: KbOptimizerOptions | None = None):
        """One Optimizer2::optimize() with the LM policy, state resident on the device."""
        options = options or KbOptimizerOptions.kalibr2_default()
        sol = KbSolution()
        self._check(self._L.kb_optimize(self._h, C.byref(options), C.byref(sol)), "kb_optimize")
        n = self._L.kb_get_trace(self._h, None, 0)
        tr = np.zeros((n, 3))
        if n:
            self._L.kb_get_trace(self._h, _p(tr), n)
        return sol, tr

    def optimize_gauss_newton(self, options: KbOptimizerOptions | None = None, solver_options=None):
        """Optimizer2 with the Gauss-Newton policy over the truncated-SVD solver (the incremental estimator's optimisation)."""
        from .problem import KbSvdSolverOptions

        options = options or KbOptimizerOptions.estimator_default()
        so = solver_options or KbSvdSolverOptions.kalibr2()
        sol = KbSolution()
        self._check(self._L.kb_optimize_gauss_newton(self._h, C.byref(options), C.byref(so), C.byref(sol)), "kb_optimize_gauss_newton")
        n = self._L.kb_get_trace(self._h, None, 0)
        tr = np.zeros((n, 3))
        if n:
            self._L.kb_get_trace(self._h, _p(tr), n)
        return sol, tr

    # -- read back
    def error_vector(self) -> np.ndarray:
        e = np.zeros(self.local_jrows)
        self._check(self._L.kb_get_error_vector(self._h, _p(e)), "kb_get_error_vector")
        return e

    def rhs(self) -> np.ndarray:
        r = np.zeros(self.jcols)
        self._check(self._L.kb_get_rhs(self._h, _p(r)), "kb_get_rhs")
        return r

    def linearise(self):
        self._check(self._L.kb_linearise(self._h), "kb_linearise")

    def jacobian_ccs(self):
        self.linearise()
        nnz = self._L.kb_jacobian_nnz(self._h)
        col_ptr = np.zeros(self.local_jrows + 1, np.int64)
        row_idx = np.zeros(nnz, np.int32)
        vals = np.zeros(nnz)
        self._check(self._L.kb_get_jacobian_ccs(self._h, _p(col_ptr), _p(row_idx), _p(vals)), "kb_get_jacobian_ccs")
        return col_ptr, row_idx, vals

    def hessian_blocks(self):
        nb, nv = C.c_int64(), C.c_int64()
        self._check(self._L.kb_get_hessian_blocks(self._h, C.byref(nb), C.byref(nv), None, None, None, None), "kb_get_hessian_blocks")
        col_ptr = np.zeros(self.n_dv + 1, np.int64)
        block_row = np.zeros(nb.value, np.int32)
        value_ptr = np.zeros(nb.value, np.int64)
        values = np.zeros(nv.value)
        self._check(
            self._L.kb_get_hessian_blocks(self._h, C.byref(nb), C.byref(nv), _p(col_ptr), _p(block_row), _p(value_ptr), _p(values)),
            "kb_get_hessian_blocks",
        )
        return col_ptr, block_row, value_ptr, values

    def camera_params(self) -> np.ndarray:
        out = np.zeros((self.problem.n_cams, 10))
        self._check(self._L.kb_get_camera_params(self._h, _p(out)), "kb_get_camera_params")
        return out

    def baselines(self) -> np.ndarray:
        out = np.zeros((max(self.problem.n_cams - 1, 0), 7))
        if out.size:
            self._check(self._L.kb_get_baselines(self._h, _p(out)), "kb_get_baselines")
        return out

    def set_poses(self) -> np.ndarray:
        out = np.zeros((getattr(self, "_n_sets_live", self.problem.n_sets), 7))
        self._check(self._L.kb_get_set_poses(self._h, _p(out)), "kb_get_set_poses")
        return out

    def _refresh_sizes(self):
        self.jrows = self._L.kb_jrows(self._h)
        self.local_jrows = self._L.kb_local_jrows(self._h)
        self.jcols = self._L.kb_jcols(self._h)
        self.n_dv = self._L.kb_num_design_variables(self._h)

    def append_set(self, view_cam, view_begin, y_u, y_v, corner_id, set_pose):
        """≙ IncrementalOptimizationProblem::add of one batch (one synced set): the live handle grows by this set and its views."""
        vc = np.ascontiguousarray(view_cam, np.int32)
        vb = np.ascontiguousarray(view_begin, np.int64)
        yu, yv = np.ascontiguousarray(y_u, np.float64), np.ascontiguousarray(y_v, np.float64)
        cid = np.ascontiguousarray(corner_id, np.int32)
        pose = np.ascontiguousarray(set_pose, np.float64)
        self._check(self._L.kb_append_set(self._h, len(vc), _p(vc), _p(vb), _p(yu), _p(yv), _p(cid), _p(pose)), "kb_append_set")
        self._refresh_sizes()
        self._n_sets_live = getattr(self, "_n_sets_live", self.problem.n_sets) + 1

    def remove_last_set(self):
        self._check(self._L.kb_remove_last_set(self._h), "kb_remove_last_set")
        self._refresh_sizes()
        self._n_sets_live = getattr(self, "_n_sets_live", self.problem.n_sets) - 1

    def save_design_variables(self):
        self._check(self._L.kb_save_design_variables(self._h), "kb_save_design_variables")

    def restore_design_variables(self):
        self._check(self._L.kb_restore_design_variables(self._h), "kb_restore_design_variables")

    def set_state(self, cam_params=None, baselines=None, set_poses=None):
        """≙ DesignVariable::setParameters for every design variable: the host's values replace the device state."""
        a = [None if x is None else np.ascontiguousarray(x, np.float64) for x in (cam_params, baselines, set_poses)]
        self._check(self._L.kb_set_state(self._h, _p(a[0]), _p(a[1]), _p(a[2])), "kb_set_state")

    def set_conditioner(self, diag):
        d = np.ascontiguousarray(diag, np.float64)
        self._check(self._L.kb_set_conditioner(self._h, _p(d)), "kb_set_conditioner")

    @staticmethod
    def _obs_suffix(y_u, y_v):
        """float32 arrays (the detector's precision) go through the *_f32 entry points: half the bytes over PCIe, widened on the device"""
        if y_u.dtype == np.float32 and y_v.dtype == np.float32:
            return "_f32"
        if y_u.dtype == np.float64 and y_v.dtype == np.float64:
            return ""
        raise KalibrB200Error("observations must be two float64 or two float32 arrays")

    def set_observations(self, y_u: np.ndarray, y_v: np.ndarray):
        fn = getattr(self._L, "kb_set_observations" + self._obs_suffix(y_u, y_v))
        self._check(fn(self._h, _p(y_u), _p(y_v)), "kb_set_observations")

    def evaluate_error_streamed(self, y_u: np.ndarray, y_v: np.ndarray, use_m_estimator: bool = True) -> float:
        """set_observations + evaluate_error with the upload pipelined against the evaluation (pass pinned host arrays)."""
        J = C.c_double(0.0)
        fn = getattr(self._L, "kb_evaluate_error_streamed" + self._obs_suffix(y_u, y_v))
        self._check(fn(self._h, _p(y_u), _p(y_v), 1 if use_m_estimator else 0, C.byref(J)), "kb_evaluate_error_streamed")
        return J.value

    def prefetch_observations(self, y_u: np.ndarray, y_v: np.ndarray):
        """Start the upload of the NEXT batch of measurements into the back buffers (pinned host arrays, kept alive by the caller)."""
        fn = getattr(self._L, "kb_prefetch_observations" + self._obs_suffix(y_u, y_v))
        self._check(fn(self._h, _p(y_u), _p(y_v)), "kb_prefetch_observations")

    def commit_observations(self):
        self._check(self._L.kb_commit_observations(self._h), "kb_commit_observations")

    def analyze_marginal(self, options=None, last_build: bool = False):
        """≙ LinearSolver::analyzeMarginal: (KbMarginalResult, singular values, V, DV column of each row of V).
        last_build: analyse the system of the last build_system (what addBatch sees after optimize()) instead of re-linearising."""
        from .problem import KbMarginalOptions, KbMarginalResult

        o = options or KbMarginalOptions.default()
        n = self.problem.n_c
        res = KbMarginalResult()
        sv, V, cols = np.zeros(n), np.zeros((n, n)), np.zeros(n, np.int32)
        fn = self._L.kb_analyze_marginal_last_build if last_build else self._L.kb_analyze_marginal
        self._check(fn(self._h, C.byref(o), C.byref(res), _p(sv), _p(V), _p(cols)), "kb_analyze_marginal")
        return res, sv, V, cols

    def last_svd_decomposition(self):
        """(|eigenvalues| descending, V, DV columns of V's rows) of the last solve_system_svd (the column-scaled system when scaling is on)."""
        n = self.problem.n_c
        sv, V, cols = np.zeros(n), np.zeros((n, n)), np.zeros(n, np.int32)
        self._check(self._L.kb_get_last_svd_decomposition(self._h, _p(sv), _p(V), _p(cols)), "kb_get_last_svd_decomposition")
        return sv, V, cols

    def last_svd_solve(self):
        from .problem import KbSvdSolveResult

        res = KbSvdSolveResult()
        self._check(self._L.kb_get_last_svd_solve(self._h, C.byref(res)), "kb_get_last_svd_solve")
        return res

    def peer_exchange_handle(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's exchange buffer (all-gather them, then attach_peers on every rank)."""
        buf = C.create_string_buffer(64)
        self._check(self._L.kb_peer_exchange_handle(self._h, buf), "kb_peer_exchange_handle")
        return buf.raw

    def attach_peers(self, handles: bytes):
        """handles: n_ranks x 64 bytes in rank order.  Switches the three exchange steps of an iteration from NCCL to NVLink stores."""
        self._check(self._L.kb_attach_peers(self._h, handles), "kb_attach_peers")

    def reset_state(self):
        self._check(self._L.kb_reset_state(self._h), "kb_reset_state")

    def num_invalid_terms(self) -> int:
        return int(self._L.kb_num_invalid_terms(self._h))

    # -- instrumentation
    def kernel_launches(self) -> int:
        return int(self._L.kb_kernel_launches(self._h))

    def enable_stage_timing(self, on: bool = True, stages=None):
        """stages: names from STAGE_NAMES to time (default: all of them)"""
        mask = 1 if on else 0
        if on and stages is not None:
            mask = sum(2 << STAGE_NAMES.index(n) for n in stages)
        self._L.kb_enable_stage_timing(self._h, mask)

    def stage_ms(self) -> dict:
        ms = np.zeros(KB_NUM_STAGES)
        self._L.kb_get_stage_ms(self._h, _p(ms))
        return dict(zip(STAGE_NAMES, ms.tolist()))

    def stage_totals(self) -> dict:
        """{stage: (total_ms, calls)} accumulated since enable_stage_timing(True)."""
        ms = np.zeros(KB_NUM_STAGES)
        calls = np.zeros(KB_NUM_STAGES, np.int64)
        self._L.kb_get_stage_totals(self._h, _p(ms), _p(calls))
        return {n: (float(m), int(k)) for n, m, k in zip(STAGE_NAMES, ms, calls)}

    def cuda_stream(self) -> int:
        return int(self._L.kb_cuda_stream(self._h) or 0)
