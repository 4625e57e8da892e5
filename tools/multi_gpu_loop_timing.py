"""N-rank kb_optimize wall time per LM iteration with the NVLink peer exchange (graph replay possible) vs NCCL (torchrun)."""
import os, sys, time, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kalibr_b200 import synthetic, capi
from kalibr_b200.problem import KbOptimizerOptions

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

def bcast_id():
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().tolist())

for cfg in [int(a) for a in sys.argv[1:]] or [2, 3, 4]:
    p = synthetic.make_config(cfg)
    row = {"cfg": cfg, "world": world, "terms": p.n_terms}
    for name, px in (("nccl", False), ("peer_exchange", True)):
        g = capi.B200SchurLinearSystemSolver(p, n_ranks=world, rank=rank, nccl_id=bcast_id(), device=lr)
        if px:
            mine = torch.tensor(list(g.peer_exchange_handle()), dtype=torch.uint8, device="cuda")
            allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            g.attach_peers(b"".join(bytes(t.cpu().tolist()) for t in allh))
        best = None
        for rep in range(4):
            g.reset_state()
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            sol, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
            dt = time.perf_counter() - t0
            if rep > 0:
                best = dt if best is None else min(best, dt)
        row[name] = {"iterations": sol.iterations, "ms_per_iteration": round(1e3 * best / max(sol.iterations, 1), 4), "j_final": sol.j_final}
        g.close()
    if rank == 0:
        print(json.dumps(row), flush=True)
dist.barrier()
dist.destroy_process_group()
