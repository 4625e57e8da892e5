"""Synthetic aprilgrid calibration problems.  The reference ships no generator; this one
defines the five BASELINE.json configs: 6x5 aprilgrid (120 corners), ground-truth intrinsics = the
reference's getTestProjection() parameter sets, noisy observations, perturbed initial guess.

Only forward projection (numpy, vectorised) is needed here; it is generator code, not the hot path.
"""
from __future__ import annotations

import numpy as np

from .problem import (
    CAM_PARAM_STRIDE,
    DS_NONE,
    EUCM_NONE,
    MODEL_D,
    MODEL_P,
    OMNI_NONE,
    OMNI_RADTAN,
    ORDER_BATCH,
    ORDER_RIG,
    ORDER_SINGLE,
    ORDER_STEREO,
    PINHOLE_EQUI,
    PINHOLE_FOV,
    PINHOLE_RADTAN,
    Problem,
)

# ground truth: reference getTestProjection()/getTestDistortion()
#   pinhole  PinholeProjection.hpp:566-569          (400,400,320,240) 640x480
#   omni     OmniProjection.hpp:687-690             xi 0.9 + same K
#   eucm     ExtendedUnifiedProjection.hpp:714-717  (0.63,1.04,380,380,640,512) 1280x1024
#   ds       DoubleSphereProjection.hpp:765-768     (-0.18,0.59,313,313,640,512) 1280x1024
#   radtan   RadialTangentialDistortion.cpp:74-76   (-0.2,0.13,0.0005,0.0005)
#   equi     CAM/test/EquidistantDistortion.cpp:11  (-0.00185,0.01901,-0.02422,0.01429)
TRUTH_PARAMS = {
    PINHOLE_RADTAN: ([400.0, 400.0, 320.0, 240.0, -0.2, 0.13, 0.0005, 0.0005], (640, 480)),
    PINHOLE_EQUI: ([400.0, 400.0, 320.0, 240.0, -0.00185, 0.01901, -0.02422, 0.01429], (640, 480)),
    OMNI_RADTAN: ([0.9, 400.0, 400.0, 320.0, 240.0, -0.2, 0.13, 0.0005, 0.0005], (640, 480)),
    EUCM_NONE: ([0.63, 1.04, 380.0, 380.0, 640.0, 512.0], (1280, 1024)),
    DS_NONE: ([-0.18, 0.59, 313.0, 313.0, 640.0, 512.0], (1280, 1024)),
    PINHOLE_FOV: ([400.0, 400.0, 320.0, 240.0, 1.0], (640, 480)),  # FovDistortion::getTestDistortion: w = 1 (src/FovDistortion.cpp:52-54)
    OMNI_NONE: ([0.9, 400.0, 400.0, 320.0, 240.0], (640, 480)),
}


def aprilgrid_points(tag_rows: int = 5, tag_cols: int = 6, tag_size: float = 0.088, tag_spacing: float = 0.2954) -> np.ndarray:
    """Corner coordinates in the target frame, row-major over the (2*tag_rows) x (2*tag_cols) corner grid
    (reference: aslam_cv/aslam_cameras_april/src/GridCalibrationTargetAprilgrid.cpp:78-98)."""
    rows, cols = 2 * tag_rows, 2 * tag_cols
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    x = (c // 2) * (1 + tag_spacing) * tag_size + (c % 2) * tag_size
    y = (r // 2) * (1 + tag_spacing) * tag_size + (r % 2) * tag_size
    return np.stack([x.ravel(), y.ravel(), np.zeros(rows * cols)], axis=1).astype(np.float64)


# ---- rotations: sm_kinematics convention (quat2r of a scalar-last quaternion, quaternion_algebra.cpp:77-101) ----
def quat2r(q: np.ndarray) -> np.ndarray:
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = x * x - y * y - z * z + w * w
    R[..., 0, 1] = 2 * (x * y + z * w)
    R[..., 0, 2] = 2 * (x * z - y * w)
    R[..., 1, 0] = 2 * (x * y - z * w)
    R[..., 1, 1] = -x * x + y * y - z * z + w * w
    R[..., 1, 2] = 2 * (x * w + y * z)
    R[..., 2, 0] = 2 * (x * z + y * w)
    R[..., 2, 1] = 2 * (y * z - x * w)
    R[..., 2, 2] = -x * x - y * y + z * z + w * w
    return R


def axis_angle_to_quat(a: np.ndarray) -> np.ndarray:
    """Quaternion q (x,y,z,w) such that quat2r(q) = exp([a]x) (the usual rotation by |a| about a)."""
    a = np.asarray(a, np.float64)
    th = np.linalg.norm(a, axis=-1, keepdims=True)
    safe = np.where(th > 1e-12, th, 1.0)
    n = a / safe
    # quat2r is the transpose of the Hamilton matrix, hence the sign on the vector part
    return np.concatenate([-n * np.sin(0.5 * th), np.cos(0.5 * th)], axis=-1)


def quat_mul_update(q: np.ndarray, dq: np.ndarray) -> np.ndarray:
    """sm::kinematics::updateQuat (quaternion_algebra.cpp:302-317), vectorised; dq is a 3-vector."""
    th = np.linalg.norm(dq, axis=-1, keepdims=True)
    na = np.where(th < np.finfo(np.float64).eps ** 0.25, 0.5 + th * th / 48.0, np.sin(0.5 * th) / np.where(th > 0, th, 1.0))
    d = np.concatenate([dq * na, np.cos(0.5 * th)], axis=-1)
    ca = d[..., 3]
    out = np.empty_like(q)
    out[..., 0] = q[..., 0] * ca + d[..., 0] * q[..., 3] - d[..., 1] * q[..., 2] + d[..., 2] * q[..., 1]
    out[..., 1] = q[..., 1] * ca + d[..., 0] * q[..., 2] + d[..., 1] * q[..., 3] - d[..., 2] * q[..., 0]
    out[..., 2] = q[..., 2] * ca - d[..., 0] * q[..., 1] + d[..., 1] * q[..., 0] + d[..., 2] * q[..., 3]
    out[..., 3] = q[..., 3] * ca - d[..., 0] * q[..., 0] - d[..., 1] * q[..., 1] - d[..., 2] * q[..., 2]
    return out


def r_to_quat(R: np.ndarray) -> np.ndarray:
    """Inverse of quat2r for proper rotations (batched, w >= 0)."""
    Rt = np.swapaxes(R, -1, -2)  # Hamilton matrix of q
    m00, m11, m22 = Rt[..., 0, 0], Rt[..., 1, 1], Rt[..., 2, 2]
    w = 0.5 * np.sqrt(np.maximum(1 + m00 + m11 + m22, 1e-300))
    x = (Rt[..., 2, 1] - Rt[..., 1, 2]) / (4 * w)
    y = (Rt[..., 0, 2] - Rt[..., 2, 0]) / (4 * w)
    z = (Rt[..., 1, 0] - Rt[..., 0, 1]) / (4 * w)
    q = np.stack([x, y, z, w], axis=-1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


# ---- forward projection (generator only) ---------------------------------------------------------------
def _radtan(mx, my, k):
    k1, k2, p1, p2 = k
    r2 = mx * mx + my * my
    rad = k1 * r2 + k2 * r2 * r2
    return (mx + mx * rad + 2 * p1 * mx * my + p2 * (r2 + 2 * mx * mx), my + my * rad + 2 * p2 * mx * my + p1 * (r2 + 2 * my * my))


def _equi(mx, my, k):
    r = np.sqrt(mx * mx + my * my)
    th = np.arctan(r)
    th2 = th * th
    thd = th * (1 + k[0] * th2 + k[1] * th2**2 + k[2] * th2**3 + k[3] * th2**4)
    s = np.where(r > 1e-8, thd / np.where(r > 1e-8, r, 1.0), 1.0)
    return mx * s, my * s


def _fov(mx, my, k):
    w = k[0]
    r = np.sqrt(mx * mx + my * my)
    t = np.tan(0.5 * w)
    if w * w < 1e-5:
        s = np.ones_like(r)
    else:
        small = r * r < 1e-5
        s = np.where(small, 2 * t / w, np.arctan(2 * t * r) / (np.where(small, 1.0, r) * w))
    return mx * s, my * s


def project(model: int, params, p: np.ndarray):
    """p: [..., 3] points in the camera frame -> (u, v, valid)."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    params = list(params)
    if model in (PINHOLE_RADTAN, PINHOLE_EQUI, PINHOLE_FOV):
        fu, fv, cu, cv = params[:4]
        zs = np.where(z > 1e-9, z, 1.0)
        mx, my = x / zs, y / zs
        mx, my = {PINHOLE_RADTAN: _radtan, PINHOLE_EQUI: _equi, PINHOLE_FOV: _fov}[model](mx, my, params[4:8])
        return fu * mx + cu, fv * my + cv, z > 1e-3
    if model in (OMNI_RADTAN, OMNI_NONE):
        xi, fu, fv, cu, cv = params[:5]
        d = np.sqrt(x * x + y * y + z * z)
        fov = xi if xi <= 1.0 else 1.0 / xi
        den = z + xi * d
        ok = z > -(fov * d) * 0.9
        den = np.where(ok, den, 1.0)
        mx, my = x / den, y / den
        if model == OMNI_RADTAN:
            mx, my = _radtan(mx, my, params[5:9])
        return fu * mx + cu, fv * my + cv, ok
    if model == EUCM_NONE:
        al, be, fu, fv, cu, cv = params[:6]
        d = np.sqrt(be * (x * x + y * y) + z * z)
        fov = al / (1 - al) if al <= 0.5 else (1 - al) / al
        ok = z > -(fov * d) * 0.9
        norm = np.where(ok, al * d + (1 - al) * z, 1.0)
        return fu * x / norm + cu, fv * y / norm + cv, ok
    if model == DS_NONE:
        xi, al, fu, fv, cu, cv = params[:6]
        r2 = x * x + y * y
        d1 = np.sqrt(r2 + z * z)
        t = al / (1 - al) if al <= 0.5 else (1 - al) / al
        fov = (t + xi) / np.sqrt(2 * t * xi + xi * xi + 1)
        ok = z > -(fov * d1) * 0.9
        k = xi * d1 + z
        d2 = np.sqrt(r2 + k * k)
        norm = np.where(ok, al * d2 + (1 - al) * k, 1.0)
        return fu * x / norm + cu, fv * y / norm + cv, ok
    raise ValueError(f"unknown model {model}")


# ---- configs ------------------------------------------------------------------------------------------
CONFIGS = {
    # cfg: (driver order, models, S)   — SURVEY.md §8d table
    1: (ORDER_SINGLE, [PINHOLE_RADTAN], 300),
    2: (ORDER_STEREO, [PINHOLE_RADTAN] * 2, 2000),
    3: (ORDER_RIG, [OMNI_RADTAN, EUCM_NONE, DS_NONE, PINHOLE_EQUI], 5000),
    4: (ORDER_RIG, [PINHOLE_RADTAN] * 8, 20000),
    5: (ORDER_RIG, [PINHOLE_RADTAN] * 16, 6250),
    # not a BASELINE config: the two remaining rows of kalibr2::CreateCalibrator's model table next to a pinhole-radtan camera
    6: (ORDER_RIG, [PINHOLE_FOV, OMNI_NONE, PINHOLE_RADTAN], 1000),
    7: (ORDER_STEREO, [OMNI_NONE, PINHOLE_FOV], 1000),
    # the design-variable order of the incremental estimator's merged problem (CreateBatchProblem: set poses, baselines, intrinsics)
    8: (ORDER_BATCH, [PINHOLE_RADTAN, OMNI_RADTAN, PINHOLE_EQUI], 1000),
}


def make_problem(
    models,
    n_sets: int,
    driver_order: int,
    seed: int,
    noise_px: float = 0.3,
    float_corners: bool = False,
    dropout: float = 0.0,
    perturb: bool = True,
    name: str = "",
    set_seed: int | None = None,
) -> Problem:
    """Generate one problem: dense observations of the 120-corner aprilgrid by every camera in every synced set
    (optionally with per-corner dropout), measurement noise, and a perturbed initial guess.

    The rig (baselines, intrinsics and their initial guess) is drawn from `seed`; the synced sets (poses, noise)
    from `set_seed` (default: seed + 7919), so that ranks of a pre-sharded multi-GPU problem can generate their own
    sets while sharing the same cameras."""
    rng = np.random.default_rng(seed)
    rng_cam_guess = np.random.default_rng([seed, 1])
    set_rng = np.random.default_rng(seed + 7919 if set_seed is None else set_seed)
    models = [int(m) for m in models]
    Cn = len(models)
    pts = aprilgrid_points()
    T = pts.shape[0]
    centre = pts.mean(axis=0)

    truth_params = np.zeros((Cn, CAM_PARAM_STRIDE))
    res = []
    for k, m in enumerate(models):
        p, r = TRUTH_PARAMS[m]
        truth_params[k, : len(p)] = p
        res.append(r)

    # rig: chained baselines T_cam(k+1)_cam(k); spread scaled so that a long chain keeps a common field of view
    scale = 1.0 / max(1, Cn - 1)
    base_aa = np.deg2rad(rng.uniform(-10, 10, size=(max(Cn - 1, 0), 3))) * scale
    base_t = rng.uniform(-0.15, 0.15, size=(max(Cn - 1, 0), 3)) * scale
    base_q = axis_angle_to_quat(base_aa) if Cn > 1 else np.zeros((0, 4))
    base_R = quat2r(base_q) if Cn > 1 else np.zeros((0, 3, 3))
    # cumulative T_cam(k)_cam(0)
    cumR = [np.eye(3)]
    cumt = [np.zeros(3)]
    for k in range(Cn - 1):
        cumR.append(base_R[k] @ cumR[-1])
        cumt.append(base_R[k] @ cumt[-1] + base_t[k])

    # set poses: accept a set only if every corner is valid and inside the image in every camera
    acc_q, acc_t = [], []
    need = n_sets
    margin = 4.0
    while need > 0:
        B = max(256, int(need * 2.5))
        depth = set_rng.uniform(0.4, 0.9, size=B)
        aa = np.deg2rad(set_rng.uniform(-25, 25, size=(B, 3)))
        lat = set_rng.uniform(-0.35, 0.35, size=(B, 2)) * depth[:, None]
        q_ct = axis_angle_to_quat(aa)  # rotation target -> cam0
        R_ct = quat2r(q_ct)
        # place the target centre at (lat, depth) in cam0
        t_ct = np.concatenate([lat, depth[:, None]], axis=1) - np.einsum("bij,j->bi", R_ct, centre)
        ok = np.ones(B, bool)
        p0 = np.einsum("bij,tj->bti", R_ct, pts) + t_ct[:, None, :]
        for k, m in enumerate(models):
            pk = np.einsum("ij,btj->bti", cumR[k], p0) + cumt[k]
            u, v, valid = project(m, truth_params[k], pk)
            inside = valid & (u > margin) & (u < res[k][0] - margin) & (v > margin) & (v < res[k][1] - margin)
            ok &= inside.all(axis=1)
        idx = np.nonzero(ok)[0][:need]
        acc_q.append(q_ct[idx])
        acc_t.append(t_ct[idx])
        need -= idx.size
    q_ct = np.concatenate(acc_q)
    t_ct = np.concatenate(acc_t)
    R_ct = quat2r(q_ct)
    # pose design variable = T_target_cam0 (the error term applies its inverse, CalibrationTools.hpp:405)
    R_tc = np.swapaxes(R_ct, 1, 2)
    t_tc = -np.einsum("bij,bj->bi", R_tc, t_ct)
    q_tc = r_to_quat(R_tc)

    # observations, in the driver's error-term insertion order
    if driver_order == ORDER_STEREO:
        view_cam = np.repeat(np.arange(Cn, dtype=np.int32), n_sets)
        view_set = np.tile(np.arange(n_sets, dtype=np.int32), Cn)
    else:
        view_set = np.repeat(np.arange(n_sets, dtype=np.int32), Cn)
        view_cam = np.tile(np.arange(Cn, dtype=np.int32), n_sets)
    V = view_set.size
    u_all = np.empty((V, T))
    v_all = np.empty((V, T))
    chunk = 4096
    for k, m in enumerate(models):
        sel = np.nonzero(view_cam == k)[0]
        for a in range(0, sel.size, chunk):
            w = sel[a : a + chunk]
            s = view_set[w]
            p0 = np.einsum("bij,tj->bti", R_ct[s], pts) + t_ct[s][:, None, :]
            pk = np.einsum("ij,btj->bti", cumR[k], p0) + cumt[k]
            u, v, _ = project(m, truth_params[k], pk)
            u_all[w] = u
            v_all[w] = v
    u_all += set_rng.normal(0.0, noise_px, size=u_all.shape)
    v_all += set_rng.normal(0.0, noise_px, size=v_all.shape)
    if dropout > 0:
        keep = set_rng.uniform(size=(V, T)) >= dropout
        keep[:, 0] = True  # never an empty view unless asked for explicitly
    else:
        keep = np.ones((V, T), bool)
    counts = keep.sum(axis=1)
    view_begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    corner_id = np.tile(np.arange(T, dtype=np.int32), V).reshape(V, T)[keep]
    if float_corners:
        # corner detectors deliver single precision (the reference keeps image points as cv::Point2f): the measurements are
        # rounded to float32 and stored as doubles, so that they can also travel as 8 B / term (kb_*_observations_f32)
        u_all = u_all.astype(np.float32).astype(np.float64)
        v_all = v_all.astype(np.float32).astype(np.float64)
    y_u = u_all[keep]
    y_v = v_all[keep]

    # initial guess
    cam0 = truth_params.copy()
    base_q0, base_t0 = base_q.copy(), base_t.copy()
    q0, t0 = q_tc.copy(), t_tc.copy()
    if perturb:
        for k, m in enumerate(models):
            P, D = MODEL_P[m], MODEL_D[m]
            n_shape = {OMNI_RADTAN: 1, OMNI_NONE: 1, EUCM_NONE: 2, DS_NONE: 2}.get(m, 0)  # xi / alpha,beta / xi,alpha lead the vector
            cam0[k, :n_shape] += rng_cam_guess.normal(0, 0.01, size=n_shape)
            cam0[k, n_shape:P] *= 1 + rng_cam_guess.uniform(-0.02, 0.02, size=P - n_shape)
            cam0[k, P : P + D] += rng_cam_guess.normal(0, 0.01, size=D)
        if Cn > 1:
            base_q0 = quat_mul_update(base_q0, rng_cam_guess.normal(0, np.deg2rad(0.5), size=base_q0[:, :3].shape))
            base_t0 = base_t0 + rng_cam_guess.normal(0, 0.005, size=base_t0.shape)
        q0 = quat_mul_update(q0, set_rng.normal(0, np.deg2rad(0.5), size=q0[:, :3].shape))
        t0 = t0 + set_rng.normal(0, 0.005, size=t0.shape)

    truth = {
        "cam_params": truth_params,
        "baselines": np.concatenate([base_q, base_t], axis=1) if Cn > 1 else np.zeros((0, 7)),
        "set_poses": np.concatenate([q_tc, t_tc], axis=1),
        "resolution": res,
    }
    return Problem(
        driver_order=driver_order,
        cam_model=np.asarray(models, np.int32),
        cam_params=cam0,
        baselines=np.concatenate([base_q0, base_t0], axis=1) if Cn > 1 else np.zeros((0, 7)),
        set_poses=np.concatenate([q0, t0], axis=1),
        target_points=pts,
        view_set=view_set,
        view_cam=view_cam,
        view_begin=view_begin,
        y_u=y_u,
        y_v=y_v,
        corner_id=corner_id,
        truth=truth,
        name=name,
    )


def make_config(cfg: int, n_sets: int | None = None, **kw) -> Problem:
    """BASELINE.json config `cfg` (1..5); `n_sets` overrides S for scaled-down parity cases."""
    order, models, S = CONFIGS[cfg]
    S = S if n_sets is None else n_sets
    kw.setdefault("seed", 20260000 + cfg)
    return make_problem(models, S, order, name=f"cfg{cfg}_S{S}", **kw)


def shard_sets(n_sets: int, n_ranks: int, rank: int) -> tuple[int, int]:
    """Contiguous range of synced sets owned by `rank`; mirrors kb_shard_range in the C ABI."""
    base, rem = divmod(n_sets, n_ranks)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
