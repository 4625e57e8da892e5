"""An INDEPENDENT numpy statement of the reprojection residual of a calibration problem, written from the camera models' published
formulas - not from the oracle's or the kernels' code - and differentiated by finite differences.  It shares nothing with oracle/ or
kalibr_b200/csrc but the conventions a design variable's VALUE is stored in (scalar-last quaternion of sm_kinematics,
T_target_cam0 per synced set, T_cam(k+1)_cam(k) per baseline): a second derivation of residuals, of every Jacobian column (including
the pose / baseline chain rule) and, through scipy's trust-region least squares, of the converged calibration.

Models (x, y, z: point in the camera frame):
  pinhole           m = (x/z, y/z)
  omni (Mei)        m = (x, y) / (z + xi |p|)
  EUCM (Khomutenko) m = (x, y) / (alpha rho + (1 - alpha) z),  rho = sqrt(beta (x^2 + y^2) + z^2)
  double sphere (Usenko)  d1 = |p|, d2 = sqrt(x^2 + y^2 + (xi d1 + z)^2),  m = (x, y) / (alpha d2 + (1 - alpha)(xi d1 + z))
  radtan (Brown)    m' = m (1 + k1 r^2 + k2 r^4) + (2 p1 mx my + p2 (r^2 + 2 mx^2), p1 (r^2 + 2 my^2) + 2 p2 mx my)
  equidistant (Kannala-Brandt)  theta = atan(r),  m' = m theta (1 + k1 th^2 + k2 th^4 + k3 th^6 + k4 th^8) / r
  FOV (Devernay-Faugeras)       m' = m atan(2 r tan(w/2)) / (w r)
  keypoint          (fu mx' + cu, fv my' + cv)
"""
import numpy as np

PINHOLE_RADTAN, PINHOLE_EQUI, OMNI_RADTAN, EUCM_NONE, DS_NONE, PINHOLE_FOV, OMNI_NONE = range(7)
N_P = [4, 4, 5, 6, 6, 4, 5]
N_D = [4, 4, 4, 0, 0, 1, 0]


def rot_from_quat(q):
    """rotation matrix of a unit quaternion (x, y, z, w) in the convention sm_kinematics stores poses in (JPL):
    C = (w^2 - v.v) I + 2 v v^T - 2 w [v]x"""
    x, y, z, w = q
    v = np.array([x, y, z])
    vx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return (w * w - v @ v) * np.eye(3) + 2.0 * np.outer(v, v) - 2.0 * w * vx


def small_rotation_update(q, d):
    """q <- exp(d) (x) q: the manifold update of a rotation design variable (left perturbation by the rotation vector d)"""
    th = np.linalg.norm(d)
    if th < 1e-12:
        dq = np.array([0.5 * d[0], 0.5 * d[1], 0.5 * d[2], 1.0])
    else:
        dq = np.concatenate([np.sin(0.5 * th) * d / th, [np.cos(0.5 * th)]])
    # quaternion product dq (+) q in the same convention: plus-matrix of dq applied to q
    x, y, z, w = dq
    M = np.array([[w, z, -y, x], [-z, w, x, y], [y, -x, w, z], [-x, -y, -z, w]])
    out = M @ q
    return out / np.linalg.norm(out)


def distort(model, d, m):
    mx, my = m
    r2 = mx * mx + my * my
    if model in (PINHOLE_RADTAN, OMNI_RADTAN):
        k1, k2, p1, p2 = d
        rad = 1.0 + k1 * r2 + k2 * r2 * r2
        return np.array([mx * rad + 2.0 * p1 * mx * my + p2 * (r2 + 2.0 * mx * mx), my * rad + p1 * (r2 + 2.0 * my * my) + 2.0 * p2 * mx * my])
    if model == PINHOLE_EQUI:
        r = np.sqrt(r2)
        if r < 1e-8:
            return np.array([mx, my])
        th = np.arctan(r)
        k1, k2, k3, k4 = d
        thd = th * (1.0 + k1 * th**2 + k2 * th**4 + k3 * th**6 + k4 * th**8)
        return np.array([mx, my]) * thd / r
    if model == PINHOLE_FOV:
        w = d[0]
        r = np.sqrt(r2)
        if r < 1e-8 or abs(w) < 1e-8:
            return np.array([mx, my])
        return np.array([mx, my]) * np.arctan(2.0 * r * np.tan(0.5 * w)) / (w * r)
    return np.array([mx, my])


def project(model, prm, p):
    x, y, z = p
    P = N_P[model]
    proj, d = prm[:P], prm[P:P + N_D[model]]
    if model in (PINHOLE_RADTAN, PINHOLE_EQUI, PINHOLE_FOV):
        m = np.array([x / z, y / z])
        fu, fv, cu, cv = proj
    elif model in (OMNI_RADTAN, OMNI_NONE):
        xi, fu, fv, cu, cv = proj
        m = np.array([x, y]) / (z + xi * np.sqrt(x * x + y * y + z * z))
    elif model == EUCM_NONE:
        alpha, beta, fu, fv, cu, cv = proj
        rho = np.sqrt(beta * (x * x + y * y) + z * z)
        m = np.array([x, y]) / (alpha * rho + (1.0 - alpha) * z)
    else:
        xi, alpha, fu, fv, cu, cv = proj
        d1 = np.sqrt(x * x + y * y + z * z)
        d2 = np.sqrt(x * x + y * y + (xi * d1 + z) ** 2)
        m = np.array([x, y]) / (alpha * d2 + (1.0 - alpha) * (xi * d1 + z))
    md = distort(model, d, m)
    return np.array([fu * md[0] + cu, fv * md[1] + cv])


class State:
    """values of every design variable of a problem, with the manifold update of the reference's design variables"""

    def __init__(self, problem):
        self.cam = np.array(problem.cam_params, float)
        self.base = np.array(problem.baselines, float).reshape(-1, 7)
        self.sets = np.array(problem.set_poses, float).reshape(-1, 7)

    def copy(self):
        s = State.__new__(State)
        s.cam, s.base, s.sets = self.cam.copy(), self.base.copy(), self.sets.copy()
        return s


def residuals(problem, st):
    """-(y - y_hat) per term, in term order: what LinearSystemSolver::e() holds"""
    C = len(problem.cam_model)
    # T_cam(k)_cam(0) = B_(k-1) ... B_0
    Rk, tk = [np.eye(3)], [np.zeros(3)]
    for j in range(C - 1):
        R, t = rot_from_quat(st.base[j, :4]), st.base[j, 4:]
        Rk.append(R @ Rk[-1])
        tk.append(R @ tk[-1] + t)
    out = np.zeros(2 * problem.n_terms)
    for w in range(len(problem.view_set)):
        v, k = int(problem.view_set[w]), int(problem.view_cam[w])
        Rt, tt = rot_from_quat(st.sets[v, :4]), st.sets[v, 4:]  # T_target_cam0
        R = Rk[k] @ Rt.T                                         # T_cam(k)_target = T_cam(k)_cam(0) inverse(T_target_cam0)
        t = tk[k] - R @ tt
        model = int(problem.cam_model[k])
        for i in range(int(problem.view_begin[w]), int(problem.view_begin[w + 1])):
            p = R @ problem.target_points[problem.corner_id[i]] + t
            yh = project(model, st.cam[k], p)
            out[2 * i] = -(problem.y_u[i] - yh[0])
            out[2 * i + 1] = -(problem.y_v[i] - yh[1])
    return out


def perturbed(problem, st, column, h):
    """the state after a step of size h along design-variable column `column` (the problem's own layout)"""
    col, dims, labels = problem.dv_layout()
    b = int(np.searchsorted(col, column, side="right") - 1)
    i = column - int(col[b])
    kind, idx = labels[b]
    s = st.copy()
    if kind == "proj":
        s.cam[idx, i] += h
    elif kind == "dist":
        s.cam[idx, N_P[int(problem.cam_model[idx])] + i] += h
    elif kind in ("baseline_q", "set_q"):
        arr = s.base if kind == "baseline_q" else s.sets
        d = np.zeros(3)
        d[i] = h
        arr[idx, :4] = small_rotation_update(arr[idx, :4], d)
    else:
        arr = s.base if kind == "baseline_t" else s.sets
        arr[idx, 4 + i] += h
    return s


def fd_jacobian(problem, st, h=1e-6):
    """dense d e / d (design variables) by central differences: [2 n_terms, jcols]"""
    col, dims, _ = problem.dv_layout()
    jcols = int(col[-1] + dims[-1])
    J = np.zeros((2 * problem.n_terms, jcols))
    for c in range(jcols):
        J[:, c] = (residuals(problem, perturbed(problem, st, c, h)) - residuals(problem, perturbed(problem, st, c, -h))) / (2.0 * h)
    return J


def least_squares_calibration(problem, st, max_nfev=200):
    """scipy's trust-region reflective least squares on the same cost, over the same manifold (increments applied with perturbed()):
    returns (final cost = sum e^2, state)"""
    from scipy.optimize import least_squares

    col, dims, _ = problem.dv_layout()
    jcols = int(col[-1] + dims[-1])
    cur = st.copy()
    cost = float(np.sum(residuals(problem, cur) ** 2))
    for _ in range(8):  # re-centre the manifold a few times: x = 0 at the current state
        def fun(x):
            s = cur
            for c in np.flatnonzero(x):
                s = perturbed(problem, s, int(c), float(x[c]))
            return residuals(problem, s)

        def jac(x):
            s = cur
            for c in np.flatnonzero(x):
                s = perturbed(problem, s, int(c), float(x[c]))
            return fd_jacobian(problem, s)

        r = least_squares(fun, np.zeros(jcols), jac=jac, method="trf", x_scale="jac", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=max_nfev)
        s = cur
        for c in np.flatnonzero(r.x):
            s = perturbed(problem, s, int(c), float(r.x[c]))
        cur = s
        new_cost = float(np.sum(residuals(problem, cur) ** 2))
        if abs(cost - new_cost) <= 1e-13 * new_cost:
            cost = new_cost
            break
        cost = new_cost
    return cost, cur
