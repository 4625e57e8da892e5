"""Flat batch-calibration problem description shared by the C ABI (include/kalibr_b200.h: kb_problem_desc).

It is the flattened form of what kalibr2's drivers hand to ``Optimizer2`` as an ``OptimizationProblem``
(reference: aslam_offline_calibration/kalibr2/include/kalibr2/CalibrationTools.hpp:93-144, 183-300, 376-428).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

# kb_camera_model (reference model strings: kalibr2/CameraCalibrator.hpp:421-441)
PINHOLE_RADTAN, PINHOLE_EQUI, OMNI_RADTAN, EUCM_NONE, DS_NONE, PINHOLE_FOV, OMNI_NONE = range(7)
MODEL_NAMES = ["pinhole-radtan", "pinhole-equi", "omni-radtan", "eucm-none", "ds-none", "pinhole-fov", "omni-none"]
MODEL_P = [4, 4, 5, 6, 6, 4, 5]  # projection parameters
MODEL_D = [4, 4, 4, 0, 0, 1, 0]  # distortion parameters (0-dim distortion DV stays active: SURVEY.md Q7)

# kb_driver_order
ORDER_SINGLE, ORDER_STEREO, ORDER_RIG, ORDER_BATCH = range(4)

CAM_PARAM_STRIDE = 10
POSE_STRIDE = 7


class KbProblemDesc(C.Structure):
    _fields_ = [
        ("driver_order", C.c_int32),
        ("n_cams", C.c_int32),
        ("cam_model", C.POINTER(C.c_int32)),
        ("cam_params", C.POINTER(C.c_double)),
        ("baselines", C.POINTER(C.c_double)),
        ("n_sets", C.c_int32),
        ("set_poses", C.POINTER(C.c_double)),
        ("n_target_points", C.c_int32),
        ("target_points", C.POINTER(C.c_double)),
        ("n_views", C.c_int32),
        ("view_set", C.POINTER(C.c_int32)),
        ("view_cam", C.POINTER(C.c_int32)),
        ("view_begin", C.POINTER(C.c_int64)),
        ("n_terms", C.c_int64),
        ("y_u", C.POINTER(C.c_double)),
        ("y_v", C.POINTER(C.c_double)),
        ("corner_id", C.POINTER(C.c_int32)),
        ("n_ranks", C.c_int32),
        ("rank", C.c_int32),
        ("nccl_id", C.c_char_p),
        ("device", C.c_int32),
        ("n_sets_total", C.c_int32),
        ("set_offset", C.c_int32),
        ("n_terms_total", C.c_int64),
    ]


class KbOptimizerOptions(C.Structure):
    _fields_ = [
        ("convergence_delta_x", C.c_double),
        ("convergence_delta_j", C.c_double),
        ("max_iterations", C.c_int32),
        ("lm_lambda_init", C.c_double),
        ("verbose", C.c_int32),
        ("device_loop", C.c_int32),
    ]

    @classmethod
    def kalibr2_default(cls, verbose: int = 0, device_loop: int = 1) -> "KbOptimizerOptions":
        """kalibr2::tools::CreateDefaultOptimizer (CalibrationTools.hpp:57-66)."""
        return cls(1e-3, 1.0, 200, 10.0, verbose, device_loop)

    @classmethod
    def estimator_default(cls, verbose: int = 0) -> "KbOptimizerOptions":
        """Optimizer2Options defaults (BE/include/aslam/backend/Optimizer2Options.hpp: deltas 1e-3) with kalibr2_ros' maxIterations = 20
        (CalibrateCameras.cpp:269-272): what IncrementalEstimator's optimiser runs with."""
        return cls(1e-3, 1e-3, 20, 0.0, verbose, 0)


class KbMarginalOptions(C.Structure):
    _fields_ = [("eps_svd", C.c_double), ("svd_tol", C.c_double)]

    @classmethod
    def default(cls) -> "KbMarginalOptions":
        """aslam::calibration::LinearSolverOptions (src/core/LinearSolverOptions.cpp:31-36)."""
        return cls(float(np.finfo(np.float64).eps), -1.0)


class KbMarginalResult(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("rank", C.c_int32),
        ("rank_deficiency", C.c_int32),
        ("tolerance", C.c_double),
        ("sv_log2_sum", C.c_double),
        ("sv_gap", C.c_double),
    ]


class KbSolution(C.Structure):
    _fields_ = [
        ("j_start", C.c_double),
        ("j_final", C.c_double),
        ("dx_final", C.c_double),
        ("dj_final", C.c_double),
        ("iterations", C.c_int32),
        ("failed_iterations", C.c_int32),
        ("linear_solver_failure", C.c_int32),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


@dataclass
class Problem:
    """Host-side problem (numpy, C-contiguous).  Terms are in the reference's error-term insertion order."""

    driver_order: int
    cam_model: np.ndarray  # [C] int32
    cam_params: np.ndarray  # [C, 10] float64
    baselines: np.ndarray  # [C-1, 7] float64  (q xyzw, t)
    set_poses: np.ndarray  # [S, 7]
    target_points: np.ndarray  # [T, 3]
    view_set: np.ndarray  # [V] int32
    view_cam: np.ndarray  # [V] int32
    view_begin: np.ndarray  # [V+1] int64
    y_u: np.ndarray  # [N]
    y_v: np.ndarray  # [N]
    corner_id: np.ndarray  # [N] int32
    truth: dict = field(default_factory=dict)  # ground truth used by the generator (not part of the ABI)
    name: str = ""

    def __post_init__(self):
        self.cam_model = np.ascontiguousarray(self.cam_model, np.int32)
        self.cam_params = np.ascontiguousarray(self.cam_params, np.float64).reshape(-1, CAM_PARAM_STRIDE)
        self.baselines = np.ascontiguousarray(self.baselines, np.float64).reshape(-1, POSE_STRIDE)
        self.set_poses = np.ascontiguousarray(self.set_poses, np.float64).reshape(-1, POSE_STRIDE)
        self.target_points = np.ascontiguousarray(self.target_points, np.float64).reshape(-1, 3)
        self.view_set = np.ascontiguousarray(self.view_set, np.int32)
        self.view_cam = np.ascontiguousarray(self.view_cam, np.int32)
        self.view_begin = np.ascontiguousarray(self.view_begin, np.int64)
        self.y_u = np.ascontiguousarray(self.y_u, np.float64)
        self.y_v = np.ascontiguousarray(self.y_v, np.float64)
        self.corner_id = np.ascontiguousarray(self.corner_id, np.int32)

    @property
    def n_cams(self) -> int:
        return int(self.cam_model.shape[0])

    @property
    def n_sets(self) -> int:
        return int(self.set_poses.shape[0])

    @property
    def n_views(self) -> int:
        return int(self.view_set.shape[0])

    @property
    def n_terms(self) -> int:
        return int(self.y_u.shape[0])

    @property
    def n_c(self) -> int:
        """Dimension of the reduced camera system (intrinsics + baselines)."""
        return int(sum(MODEL_P[m] + MODEL_D[m] for m in self.cam_model) + 6 * (self.n_cams - 1))

    def desc(self, n_ranks: int = 1, rank: int = 0, nccl_id: bytes | None = None, device: int = 0,
             n_sets_total: int = 0, set_offset: int = 0, n_terms_total: int = 0) -> KbProblemDesc:
        """ctypes view; keeps `self` arrays alive only as long as `self` lives."""
        d = KbProblemDesc()
        d.driver_order = self.driver_order
        d.n_cams = self.n_cams
        d.cam_model = _ptr(self.cam_model, C.c_int32)
        d.cam_params = _ptr(self.cam_params, C.c_double)
        d.baselines = _ptr(self.baselines, C.c_double)
        d.n_sets = self.n_sets
        d.set_poses = _ptr(self.set_poses, C.c_double)
        d.n_target_points = int(self.target_points.shape[0])
        d.target_points = _ptr(self.target_points, C.c_double)
        d.n_views = self.n_views
        d.view_set = _ptr(self.view_set, C.c_int32)
        d.view_cam = _ptr(self.view_cam, C.c_int32)
        d.view_begin = _ptr(self.view_begin, C.c_int64)
        d.n_terms = self.n_terms
        d.y_u = _ptr(self.y_u, C.c_double)
        d.y_v = _ptr(self.y_v, C.c_double)
        d.corner_id = _ptr(self.corner_id, C.c_int32)
        d.n_ranks = n_ranks
        d.rank = rank
        d.nccl_id = nccl_id
        d.device = device
        d.n_sets_total = n_sets_total
        d.set_offset = set_offset
        d.n_terms_total = n_terms_total
        return d

    def dv_layout(self):
        """(column_base, dims, labels) of the active design variables in the driver's insertion order
        (reference: BE/src/Optimizer2.cpp:110-124 over the orders of SURVEY.md §3.2)."""
        dims, labels = [], []

        def intr(k):
            m = int(self.cam_model[k])
            dims.extend([MODEL_P[m], MODEL_D[m]])
            labels.extend([("proj", k), ("dist", k)])

        def base():
            for k in range(self.n_cams - 1):
                dims.extend([3, 3])
                labels.extend([("baseline_q", k), ("baseline_t", k)])

        def sets():
            for v in range(self.n_sets):
                dims.extend([3, 3])
                labels.extend([("set_q", v), ("set_t", v)])

        if self.driver_order == ORDER_SINGLE:
            intr(0)
            sets()
        elif self.driver_order == ORDER_STEREO:
            base()
            sets()
            intr(0)
            intr(1)
        elif self.driver_order == ORDER_BATCH:
            sets()
            base()
            for k in range(self.n_cams):
                intr(k)
        else:
            for k in range(self.n_cams):
                intr(k)
            base()
            sets()
        dims = np.asarray(dims, np.int32)
        col = np.concatenate([[0], np.cumsum(dims)[:-1]]).astype(np.int32)
        return col, dims, labels


class KbSvdSolverOptions(C.Structure):  # kb_svd_solver_options
    _fields_ = [("column_scaling", C.c_int32), ("eps_norm", C.c_double), ("eps_svd", C.c_double), ("svd_tol", C.c_double)]

    @classmethod
    def default(cls):
        eps = float(np.finfo(float).eps)
        return cls(0, eps, eps, -1.0)

    @classmethod
    def kalibr2(cls):
        """kalibr2_ros CalibrateCameras.cpp:264-267: column scaling on, epsSVD = 1e-6."""
        return cls(1, float(np.finfo(float).eps), 1e-6, -1.0)


class KbSvdSolveResult(C.Structure):  # kb_svd_solve_result
    _fields_ = [("n", C.c_int32), ("rank", C.c_int32), ("rank_deficiency", C.c_int32), ("tolerance", C.c_double), ("sv_gap", C.c_double)]
