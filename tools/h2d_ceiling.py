"""Aggregate host-to-device bandwidth of k concurrent ranks (k = 1, 2, 4, ... world), each copying the per-step measurement buffers of
the end-to-end bench arm (2 x 76.8 MB float32, pinned) to its own GPU: the ceiling of `e2e` at N GPUs, whatever the kernels do.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/h2d_ceiling.py"""
import json
import os

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 19_200_000
host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
for h in host:
    h.fill_(1.0)
dev = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]
reps = 20
out = []
k = 1
while k <= world:
    active = rank < k
    for _ in range(2 if active else 0):
        for h, d in zip(host, dev):
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if active:
        for _ in range(reps):
            for h, d in zip(host, dev):
                d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) if active else 0.0], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gb = k * reps * 2 * n * 4 / 1e9
    out.append({"ranks": k, "ms_per_step_upload": round(ms.item() / reps, 3), "aggregate_GB_per_s": round(gb / (ms.item() * 1e-3), 1),
                "e2e_ceiling_G_terms_per_s": round(k * n / (ms.item() / reps * 1e-3) / 1e9, 2)})
    k *= 2
if rank == 0:
    for o in out:
        print(json.dumps(o), flush=True)
dist.destroy_process_group()
