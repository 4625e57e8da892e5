"""world_size-2 gloo test of the multi-GPU design on the CPU: synced sets are sharded by contiguous
ranges, every rank forms the partial reduced camera system of ITS sets, one all-reduce(sum) gives the reduced system of
the whole problem, and every rank then solves it redundantly and back-substitutes its own poses.

The per-rank arithmetic is done with the CPU oracle here (the CUDA path needs a B200); what is under test is the
sharding rule (kalibr_b200.synthetic.shard_sets / pre-sharded problem descriptions) and the additivity the NCCL
all-reduce relies on.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kalibr_b200 import synthetic
from kalibr_b200.problem import Problem


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def sub_problem(p: Problem, lo: int, hi: int) -> Problem:
    """The pre-sharded description rank owning sets [lo, hi) hands to kb_create (local 0-based view_set)."""
    keep_v = (p.view_set >= lo) & (p.view_set < hi)
    counts = np.diff(p.view_begin)
    keep_t = np.repeat(keep_v, counts)
    vb = np.concatenate([[0], np.cumsum(counts[keep_v])])
    return Problem(p.driver_order, p.cam_model, p.cam_params, p.baselines, p.set_poses[lo:hi], p.target_points,
                   p.view_set[keep_v] - lo, p.view_cam[keep_v], vb, p.y_u[keep_t], p.y_v[keep_t], p.corner_id[keep_t])


def reduced_system(o, p: Problem, damping_pose: float):
    """Augmented reduced camera system [[S, b],[b^T, .]] of the oracle's H, rhs (no damping on the camera block)."""
    col, dims = o.dv_layout()
    cp, br, vp, vals = o.hessian_blocks()
    n = int(col[-1] + dims[-1])
    H = np.zeros((n, n))
    for c in range(dims.size):
        for b in range(cp[c], cp[c + 1]):
            r = br[b]
            blk = vals[vp[b]:vp[b] + dims[r] * dims[c]].reshape(dims[c], dims[r]).T
            H[col[r]:col[r] + dims[r], col[c]:col[c] + dims[c]] = blk
            H[col[c]:col[c] + dims[c], col[r]:col[r] + dims[r]] = blk.T
    rhs = o.rhs()
    _, _, labels = p.dv_layout()
    cam_idx = np.concatenate([np.arange(col[i], col[i] + dims[i]) for i, l in enumerate(labels) if not l[0].startswith("set_")]).astype(int)
    S = H[np.ix_(cam_idx, cam_idx)].copy()
    b = rhs[cam_idx].copy()
    for v in range(p.n_sets):
        idx = np.concatenate([np.arange(col[i], col[i] + 3) for i, l in enumerate(labels) if l in (("set_q", v), ("set_t", v))])
        V = H[np.ix_(idx, idx)] + damping_pose * np.eye(6)
        W = H[np.ix_(cam_idx, idx)]
        S -= W @ np.linalg.solve(V, W.T)
        b -= W @ np.linalg.solve(V, rhs[idx])
    out = np.zeros((cam_idx.size + 1, cam_idx.size + 1))
    out[:-1, :-1] = S
    out[:-1, -1] = b
    out[-1, :-1] = b
    return out, cam_idx


def _worker(rank, world, port, cfg, n_sets, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_api as oa

    lam = 10.0
    p = synthetic.make_config(cfg, n_sets=n_sets)
    lo, hi = synthetic.shard_sets(n_sets, world, rank)
    mine = sub_problem(p, lo, hi)
    o = oa.OracleProblem(mine, n_threads=1)
    cost = torch.tensor([o.evaluate_error()], dtype=torch.float64)
    o.build_system()
    S_part, cam_idx = reduced_system(o, mine, lam * lam)
    t = torch.from_numpy(S_part)
    dist.all_reduce(t)      # the one exchange step of the path (ncclAllReduce on the GPUs)
    dist.all_reduce(cost)
    S = t.numpy()
    n_c = S.shape[0] - 1
    dx_c = np.linalg.solve(S[:-1, :-1] + lam * lam * np.eye(n_c), S[:-1, -1])  # damping added once, after the reduction
    q.put((rank, float(cost.item()), dx_c, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg,n_sets", [(2, 7), (4, 5)])
def test_sharded_reduced_system_matches_single_process(oracle_lib, cfg, n_sets):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, n_sets, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # single-process reference
    p = synthetic.make_config(cfg, n_sets=n_sets)
    o = oracle_lib.OracleProblem(p, n_threads=1)
    J = o.evaluate_error()
    o.build_system()
    o.set_constant_conditioner(10.0)
    dx, ok = o.solve_system()
    assert ok
    col, dims, labels = p.dv_layout()
    cam_idx = np.concatenate([np.arange(col[i], col[i] + dims[i]) for i, l in enumerate(labels) if not l[0].startswith("set_")]).astype(int)
    ranges = sorted((lo, hi) for _, _, _, lo, hi in results)
    assert ranges[0][0] == 0 and ranges[-1][1] == n_sets and ranges[0][1] == ranges[1][0]
    for rank, cost, dx_c, lo, hi in results:
        assert abs(cost - J) <= 1e-12 * J
        assert np.abs(dx_c - dx[cam_idx]).max() <= 1e-8 * np.abs(dx[cam_idx]).max()
    # both ranks hold the identical replicated solution
    assert np.array_equal(results[0][2], results[1][2])


# ---- the widened rows: what their multi-rank paths rely on -------------------------------------------------------------------
def _worker_widened(rank, world, port, cfg, n_sets, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ko_estimator as ke
    from oracle import ko_init as ki
    from oracle import oracle_api as oa

    p = synthetic.make_config(cfg, n_sets=n_sets)
    lo, hi = synthetic.shard_sets(n_sets, world, rank)
    mine = sub_problem(p, lo, hi)
    o = oa.OracleProblem(mine, n_threads=1)
    o.evaluate_error()
    # (1) kb_reprojection_statistics: pass 0 sums (n, sum e_u, sum e_v) per camera, all-reduce, pass 1 squared deviations from the
    #     GLOBAL mean, all-reduce
    e = -o.error_vector().reshape(-1, 2)
    cam = np.repeat(mine.view_cam, np.diff(mine.view_begin))
    acc = torch.zeros((p.n_cams, 3), dtype=torch.float64)
    for k in range(p.n_cams):
        acc[k] = torch.tensor([np.sum(cam == k), e[cam == k, 0].sum(), e[cam == k, 1].sum()])
    dist.all_reduce(acc)
    mean = (acc[:, 1:] / acc[:, :1]).numpy()
    ssd = torch.zeros((p.n_cams, 2), dtype=torch.float64)
    for k in range(p.n_cams):
        ssd[k] = torch.from_numpy(((e[cam == k] - mean[k]) ** 2).sum(0))
    dist.all_reduce(ssd)
    stats = np.concatenate([acc[:, :1].numpy(), mean, np.sqrt(ssd.numpy() / (acc[:, :1].numpy() - 1)),
                            (np.linalg.norm(acc[:, 1:].numpy(), axis=1) / np.sqrt(acc[:, 0].numpy()))[:, None]], 1)
    # (2) kb_solve_system_svd: the undamped reduced system and the diagonal of the camera block are sums over the ranks; every rank
    #     then scales, decomposes and truncates the identical matrix
    o.build_system()
    S_part, cam_idx = reduced_system(o, mine, 0.0)
    col, dims = o.dv_layout()
    J = ke.dense_from_ccs(*o.jacobian_ccs(), o.jcols)
    diag = torch.from_numpy((J[:, cam_idx] ** 2).sum(0))
    t = torch.from_numpy(S_part)
    dist.all_reduce(t)
    dist.all_reduce(diag)
    S = t.numpy()
    rows = 2 * p.n_terms
    g = np.where(np.sqrt(diag.numpy()) < np.sqrt(rows * ke.EPS), 0.0, 1.0 / np.sqrt(diag.numpy()))
    Ss, bs = S[:-1, :-1] * np.outer(g, g), S[:-1, -1] * g
    U, sv, Vt = np.linalg.svd(Ss)
    tol = sv[0] * 1e-6 * len(sv)
    rank_svd = len(sv)
    for i in range(len(sv) - 1, 0, -1):
        if sv[i] > tol:
            break
        rank_svd -= 1
    x_c = g * (Vt[:rank_svd].T @ ((U[:, :rank_svd].T @ bs) / sv[:rank_svd]))
    # (3) kb_initialize_set_poses: every rank initialises its own sets
    guesses, ok = ki.target_pose_guesses(mine)
    q.put((rank, stats, x_c, rank_svd, guesses, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_svd_solve_and_pose_guesses_match_single_process(oracle_lib):
    from oracle import ko_estimator as ke
    from oracle import ko_init as ki

    cfg, n_sets, world = 2, 7, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_widened, args=(r, world, port, cfg, n_sets, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = sorted((q.get(timeout=180) for _ in range(world)), key=lambda r: r[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = synthetic.make_config(cfg, n_sets=n_sets)
    o = oracle_lib.OracleProblem(p, n_threads=1)
    st = o.reprojection_statistics()
    J, b = ke.system_of(o, p)
    cal, rest = ke.calibration_columns(p)
    x, info = ke.linear_solver_solve(J, b, cal, rest, column_scaling_on=True, eps_svd=1e-6)
    guesses, _ = ki.target_pose_guesses(p)
    for rank, stats, x_c, rank_svd, g_loc, lo, hi in results:
        assert np.abs(stats - st).max() < 1e-10
        assert rank_svd == info["rank"]
        assert np.abs(x_c - x[cal]).max() <= 1e-7 * np.abs(x[cal]).max()
        assert np.abs(g_loc - guesses[lo:hi]).max() < 1e-12
    assert np.array_equal(results[0][2], results[1][2])  # the replicated solution is identical on both ranks
