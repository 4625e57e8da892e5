"""N-rank parity check (torchrun): the sharded CUDA path must match the single-process CPU oracle.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalibr_b200 import synthetic, capi
from kalibr_b200.problem import KbOptimizerOptions

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

def fresh_nccl_id():
    """every kb_create with n_ranks > 1 needs its own id (one NCCL communicator per handle)"""
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().tolist())

def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))

def attach_peers(g):
    """all-gather the CUDA IPC handles of the exchange buffers and attach (NVLink peer exchange instead of NCCL)"""
    mine = torch.tensor(list(g.peer_exchange_handle()), dtype=torch.uint8, device="cuda")
    allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine)
    g.attach_peers(b"".join(bytes(t.cpu().tolist()) for t in allh))

ok_all = True
use_px = os.environ.get("KB_PEER_EXCHANGE", "1") != "0" and world <= 8


def camera_major(p):
    """the same observations with the views (and their terms) ordered by camera, then set: a rank's terms are then scattered over the
    caller's arrays (the gather branch of kb_create instead of the straight upload)"""
    import dataclasses
    order = np.lexsort((p.view_set, p.view_cam))
    cnt = np.diff(p.view_begin)[order]
    vb = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    idx = np.concatenate([np.arange(p.view_begin[w], p.view_begin[w + 1]) for w in order]) if len(order) else np.zeros(0, np.int64)
    return dataclasses.replace(p, view_set=p.view_set[order].copy(), view_cam=p.view_cam[order].copy(), view_begin=vb, y_u=p.y_u[idx].copy(),
                               y_v=p.y_v[idx].copy(), corner_id=p.corner_id[idx].copy())


for cfg, S in [(2, 9), (3, 7), (4, 10), (5, 5), (6, 8), (-3, 7)]:
    p = synthetic.make_config(abs(cfg), n_sets=S)
    if cfg < 0:
        p = camera_major(p)
    g = capi.B200SchurLinearSystemSolver(p, n_ranks=world, rank=rank, nccl_id=fresh_nccl_id(), device=lr)
    if use_px:
        attach_peers(g)
    J = g.evaluate_error()
    g.build_system()
    g.set_constant_conditioner(10.0)
    dx, ok = g.solve_system(fetch_dx=True, gather=True)
    rhs = g.rhs()
    rho = g.lm_rho_denominator(10.0)
    g.reset_state()
    sol, tr = g.optimize(KbOptimizerOptions.kalibr2_default())
    cam = g.camera_params()
    if rank == 0:
        from oracle import oracle_api as oa
        o = oa.OracleProblem(p)
        Jo = o.evaluate_error(); o.build_system(); o.set_constant_conditioner(10.0)
        odx, ook = o.solve_system()
        o2 = oa.OracleProblem(p)
        osol, otr = o2.optimize(KbOptimizerOptions.kalibr2_default())
        r = dict(J=abs(J - Jo) / Jo, dx=rel(dx, odx), rhs=rel(rhs, o.rhs()), rho=abs(rho - float(odx @ (10 * odx + o.rhs()))) / abs(rho),
                 iters=(sol.iterations, osol.iterations), failed=(sol.failed_iterations, osol.failed_iterations),
                 jfinal=abs(sol.j_final - osol.j_final) / osol.j_final, cam=rel(cam, o2.camera_params()))
        good = r["J"] < 1e-11 and r["dx"] < 1e-7 and r["rhs"] < 1e-9 and r["rho"] < 1e-7 and sol.iterations == osol.iterations and r["jfinal"] < 1e-9 and r["cam"] < 1e-6 and ok == ook
        ok_all &= good
        print(f"cfg{cfg} S={S} world={world} peer_exchange={use_px}: {'OK' if good else 'MISMATCH'} {r}", flush=True)
    g.close()
# ---- weighting, reprojection statistics and the initial-guess stage on sharded problems ----
INV_R = np.array([[3.0, 0.4], [0.4, 5.0]])
for cfg, S, policy in [(2, 9, (1, 1.5, 0.0, 0.0)), (3, 7, (2, 4.0, 0.0, 0.0)), (4, 10, None)]:
    p = synthetic.make_config(cfg, n_sets=S)
    g = capi.B200SchurLinearSystemSolver(p, n_ranks=world, rank=rank, nccl_id=fresh_nccl_id(), device=lr)
    if use_px:
        attach_peers(g)
    g.set_inv_r(INV_R)
    if policy:
        g.set_m_estimator(*policy)
    J = g.evaluate_error()
    g.build_system()
    g.set_constant_conditioner(10.0)
    dx, ok = g.solve_system(fetch_dx=True, gather=True)
    stats = g.reprojection_statistics()
    sol, tr = g.optimize(KbOptimizerOptions.kalibr2_default())
    stats_final = g.reprojection_statistics()
    # initial-guess stage: every rank initialises its own sets
    g.reset_state()
    n_failed = g.initialize_set_poses()
    lo, hi = synthetic.shard_sets(S, world, rank)
    mine = g.set_poses()[lo:hi]
    from oracle import ko_init as ki
    so, _ = ki.target_pose_guesses(p)
    init_err = max((np.abs(ki.pose_to_T(a) - ki.pose_to_T(b)).max() for a, b in zip(mine, so[lo:hi])), default=0.0)
    t = torch.tensor([init_err, float(n_failed)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        from oracle import oracle_api as oa
        o = oa.OracleProblem(p)
        o.set_inv_r(INV_R)
        if policy:
            o.set_m_estimator(*policy)
        Jo = o.evaluate_error(); o.build_system(); o.set_constant_conditioner(10.0)
        odx, ook = o.solve_system()
        ostats = o.reprojection_statistics()
        osol, otr = o.optimize(KbOptimizerOptions.kalibr2_default())
        ostats_final = o.reprojection_statistics()
        r = dict(J=abs(J - Jo) / Jo, dx=rel(dx, odx), stats=float(np.abs(stats - ostats).max()), iters=(sol.iterations, osol.iterations),
                 jfinal=abs(sol.j_final - osol.j_final) / osol.j_final, stats_final=float(np.abs(stats_final - ostats_final).max()),
                 init_pose_err=float(t[0].item()), init_failed=int(t[1].item()))
        good = (r["J"] < 1e-11 and r["dx"] < 1e-7 and r["stats"] < 1e-10 and sol.iterations == osol.iterations and r["jfinal"] < 1e-9
                and r["stats_final"] < 1e-6 and r["init_pose_err"] < 1e-8 and r["init_failed"] == 0)
        ok_all &= good
        print(f"weighted cfg{cfg} S={S} policy={policy} world={world} peer_exchange={use_px}: {'OK' if good else 'MISMATCH'} {r}", flush=True)
    g.close()
dist.barrier()
if rank == 0:
    print("MULTI_GPU_PARITY", "PASS" if ok_all else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
