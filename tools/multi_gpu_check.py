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
for cfg, S in [(2, 9), (3, 7), (4, 10), (5, 5), (6, 8)]:
    p = synthetic.make_config(cfg, n_sets=S)
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
dist.barrier()
if rank == 0:
    print("MULTI_GPU_PARITY", "PASS" if ok_all else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
