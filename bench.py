#!/usr/bin/env python
"""Benchmark of the B200 batch-calibration hot path (contract: see the task statement / DESIGN.md §6).

One "step" = one Levenberg-Marquardt iteration's worth of hot path on one batch of synthetic observations:
  evaluate error -> linearise + assemble -> set conditioner -> Schur-reduced solve -> rho denominator ->
  apply state update -> revert (so every step does identical work from the same state).
Workload (N=1): BASELINE.json configs[3] "8-camera rig, 20k views": 8 x pinhole-radtan, 20 000 synced sets,
19.2 M reprojection terms.  N>1: every rank holds 20 000 sets of the same rig (weak scaling, sets sharded by
rank, one NCCL all-reduce of the reduced camera system per solve); `--scaling strong` shards the single
20k-set problem instead.

  python bench.py --gpus N --steps K --warmup W                 (torchrun launches it for N>1)
  python bench.py --impl reference ...                            CPU oracle (reference semantics) on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reproj terms linearised/s (full LM iteration: evaluate+linearise+assemble+Schur solve+update)"
UNIT = "terms/s"
# algorithmic work of the fused linearise+assemble kernel per term
FLOP_PER_TERM = 750.0
BYTES_IN_PER_TERM = 18.0        # y (16 B) + corner id (2 B)
BYTES_E_PER_TERM = 16.0         # e() written by the evaluate-time launch of the fused kernel
BYTES_OUT_PER_VIEW = 1024.0     # two 8x8 FP64 tiles of the per-view Gram block (pose x pose|intrinsics|e)
BYTES_LINEARISE_PER_TERM = lambda W: 18.0 + 16.0 + 16.0 * W  # materialising linearise: ids+y in, e and J (2 x W) out


def read_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


def peaks():
    mp = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    hbm = (mp or {}).get("hbm_gbs")
    src_hbm = "MEASURED_PEAKS.json (measured)" if hbm else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp = read_json(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")) or {}
    fp64 = fp.get("fp64_dmma_tflops", 37.14)
    return (hbm or 6650.0), src_hbm, fp64, "profiles/r01_fp64_peak.json (tools/fp64_peak.cu, FP64 DMMA measured on this pool; MEASURED_PEAKS.json has no FP64 figure)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (the same counters nvidia-smi's
    clocks.sm / clocks_event_reasons.* print), every 2 ms from a thread."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as ex:
            self.err = f"NVML unavailable: {ex}"
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        bits = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception as ex:
                self.err = str(ex)
                break
            self._stop.wait(0.002)

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


def lm_step(g, lam=10.0, fetch_dx=False, host_obs=None):
    # host_obs: (y_u, y_v) pinned host arrays uploaded as part of the evaluation (end-to-end leg)
    J = g.evaluate_error_streamed(*host_obs) if host_obs is not None else g.evaluate_error()
    g.build_system()
    g.set_constant_conditioner(lam)
    dx, ok = g.solve_system(fetch_dx=fetch_dx, gather=False)
    rho = g.lm_rho_denominator(lam)
    m = g.apply_state_update()
    g.revert_last_state_update()
    return J, ok, rho, m, dx


def run_reference(args):
    """CPU arm: the oracle (restatement of the reference's aslam_backend path; the reference itself cannot be compiled
    here) on the host cores.  Each step = evaluate + build + solve on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from kalibr_b200 import synthetic
    from oracle import oracle_api as oa

    cores = os.cpu_count() or 1
    S = args.cpu_sets
    p = synthetic.make_config(args.config, n_sets=S)
    t0 = time.time()
    o = oa.OracleProblem(p, oa.BLOCK_CHOLESKY if args.cpu_regime == "block" else oa.SPARSE_CHOLESKY, n_threads=cores)
    build_s = time.time() - t0
    o.evaluate_error()
    for _ in range(args.warmup if args.warmup < 2 else 1):
        o.time_iteration(10.0)
    stage = np.zeros(3)
    t0 = time.time()
    for _ in range(args.steps):
        t, _ok = o.time_iteration(10.0)
        stage += t
    el = time.time() - t0
    val = p.n_terms * args.steps / el
    sample = (f"cfg{args.config} restricted to {S} synced sets ({p.n_terms} terms), {args.steps} LM iterations "
              f"(evaluate+build+solve), {'BlockCholesky (serial assemble)' if args.cpu_regime == 'block' else 'SparseCholesky (threaded J)'} semantic, "
              f"{cores} threads for evaluate")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, S, p.n_terms, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "stage_s_per_iteration": {"evaluate": stage[0] / args.steps, "build": stage[1] / args.steps, "solve": stage[2] / args.steps},
                         "problem_construction_s": build_s},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sets_per_rank, terms_per_rank, n_ranks):
    from kalibr_b200 import synthetic
    from kalibr_b200.problem import MODEL_NAMES

    order, models, _ = synthetic.CONFIGS[args.config]
    return {
        "workload": f"BASELINE.json configs[{args.config - 1}]: {len(models)}-camera rig ({', '.join(sorted(set(MODEL_NAMES[m] for m in models)))}), "
                    f"6x5 aprilgrid (120 corners), {sets_per_rank} synced sets per rank",
        "cameras": len(models), "synced_sets_per_rank": sets_per_rank, "terms_per_rank": terms_per_rank, "ranks": n_ranks,
        "parallelism": (f"sets sharded over {n_ranks} rank(s); per solve one exchange of the reduced camera system "
                        + ("(NCCL all-reduce)" if args.no_peer_exchange or n_ranks > 8 else "(NVLink peer stores fused into the producing kernel, summed by the consumer)"))
        if n_ranks > 1 else "single GPU",
        "l2": "inputs larger than L2 (observations alone exceed 126 MB)" if terms_per_rank * 18 > 126e6 else "L2 flushed between steps (128 MiB+ scratch write)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, help="BASELINE config 1..5 (default 4: 8-camera rig, 20k views)")
    ap.add_argument("--sets", type=int, default=None, help="override the number of synced sets per rank")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sets", type=int, default=None, help="synced sets of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-regime", default="block", choices=["block", "sparse"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-init-stage", action="store_true")
    ap.add_argument("--no-peer-exchange", action="store_true", help="N > 1: keep the exchange steps on NCCL instead of NVLink peer stores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    from kalibr_b200 import synthetic

    order, models, S_cfg = synthetic.CONFIGS[args.config]
    if args.cpu_sets is None:
        # about 10-30 s of CPU work: the oracle runs ~0.1-0.2 M terms/s per LM iteration
        args.cpu_sets = max(8, min(S_cfg, int(400_000 / (120 * len(models)))))
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from kalibr_b200 import capi

    capi.load_library()
    S_total_cfg = args.sets if args.sets is not None else S_cfg
    if args.scaling == "weak":
        S_rank = S_total_cfg
        set_offset = rank * S_rank
        S_total = S_rank * world
    else:
        lo, hi = synthetic.shard_sets(S_total_cfg, world, rank)
        S_rank, set_offset, S_total = hi - lo, lo, S_total_cfg
    p = synthetic.make_config(args.config, n_sets=S_rank, set_seed=20260000 + 100 * args.config + 7919 + set_offset)
    terms_rank = p.n_terms
    terms_total = terms_rank
    nccl_id = None
    if world > 1:
        t = torch.tensor([terms_rank], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        terms_total = int(t.item())
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())
    g = capi.B200SchurLinearSystemSolver(p, n_ranks=world, rank=rank, nccl_id=nccl_id, device=local_rank,
                                         n_sets_total=S_total if world > 1 else 0, set_offset=set_offset, n_terms_total=terms_total)
    peer_exchange = False
    if world > 1 and world <= 8 and not args.no_peer_exchange:
        # NVLink peer exchange: all-gather the CUDA IPC handles of the ranks' exchange buffers, then attach
        mine = torch.tensor(list(g.peer_exchange_handle()), dtype=torch.uint8, device="cuda")
        allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        g.attach_peers(b"".join(bytes(t.cpu().tolist()) for t in allh))
        peer_exchange = True
    stream = torch.cuda.ExternalStream(g.cuda_stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        lm_step(g)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    g.enable_stage_timing(True)
    l0 = g.kernel_launches()
    ms_total = timed(lambda: lm_step(g), args.steps)
    launches = g.kernel_launches() - l0
    totals = g.stage_totals()
    g.enable_stage_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    value = terms_total * args.steps / (ms_total * 1e-3)
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())

    # ---- end-to-end arm: host buffers in, host results out, every step ----
    e2e = None
    if not args.no_e2e:
        yu_pin = torch.from_numpy(p.y_u).pin_memory()
        yv_pin = torch.from_numpy(p.y_v).pin_memory()
        yu_np, yv_np = yu_pin.numpy(), yv_pin.numpy()

        def e2e_step_streamed():
            # H2D of the observations from pinned host memory, pipelined in chunks against the fused kernel inside the
            # library; D2H: cost, pos-def flag, dx (jcols doubles), rho, max|dx|
            lm_step(g, fetch_dx=True, host_obs=(yu_np, yv_np))

        def e2e_step_pipelined():
            # double buffering across steps: this step's batch was uploaded while the previous step computed; the next
            # step's batch starts travelling now (one 16 B/term upload per step inside the timed region either way)
            g.commit_observations()
            g.prefetch_observations(yu_np, yv_np)
            lm_step(g, fetch_dx=True)

        k_e2e = max(3, args.steps // 2)
        for _ in range(2):
            e2e_step_streamed()
        ms_streamed = timed(e2e_step_streamed, k_e2e)
        g.prefetch_observations(yu_np, yv_np)
        for _ in range(2):
            e2e_step_pipelined()

        def pipelined_run():
            for _ in range(k_e2e):
                e2e_step_pipelined()
            g.commit_observations()               # the last upload is waited for inside the timed region
            g.prefetch_observations(yu_np, yv_np)

        ms_e2e = timed(pipelined_run, 1)
        g.commit_observations()
        e2e = {"value": terms_total * k_e2e / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(16 * terms_rank), "d2h_bytes_per_step": int(8 * g.jcols + 8 * 4 + 4),
               "ms_per_step": ms_e2e / k_e2e, "steps": k_e2e,
               "single_step_latency_ms": ms_streamed / k_e2e,
               "note": "per rank; through B200SchurLinearSystemSolver (C ABI) with pinned HOST observation buffers uploaded every step and dx fetched "
                       "every step. value: double-buffered (kb_prefetch_observations / kb_commit_observations: the next step's upload overlaps this "
                       "step's kernels); single_step_latency_ms: one isolated step with kb_evaluate_error_streamed (chunked upload overlapped with "
                       "the fused kernel). Both are bounded by the 16 B/term PCIe transfer"}

    # ---- materialising linearise (HBM-bound variant), timed alone ----
    lin = None
    if rank == 0:
        try:
            g.linearise()
            g.enable_stage_timing(True)
            for _ in range(5):
                g.linearise()
            tot = g.stage_totals()["linearise_materialise"]
            g.enable_stage_timing(False)
            nnz = capi.load_library().kb_jacobian_nnz(g._h)
            lin_bytes = 18.0 * terms_rank + 16.0 * terms_rank + 8.0 * nnz
            lin_ms = tot[0] / max(tot[1], 1)
            hbm_peak, hbm_src, _, _ = peaks()
            lin = {"kernel": "linearise_materialise", "ms": lin_ms, "terms_per_s": terms_rank / (lin_ms * 1e-3),
                   "bound": "hbm", "achieved": lin_bytes / (lin_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": lin_bytes / (lin_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_term": lin_bytes / terms_rank}
        except Exception as ex:  # e.g. out of memory for the J buffer on huge problems
            lin = {"error": str(ex)}

    # ---- initial-guess stage (the step before the path): PnP for every view, timed alone ----
    init = None
    if rank == 0 and not args.no_init_stage:
        try:
            g.estimate_transformations()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record(stream)
            for _ in range(reps):
                T_views, ok_views = g.estimate_transformations()
            e1.record(stream)
            torch.cuda.synchronize()
            pnp_ms = e0.elapsed_time(e1) / reps
            init = {"kernel": "pnp_kernel (estimateTransformation, one warp per view)", "views": int(p.n_views), "ms": pnp_ms,
                    "views_per_s": p.n_views / (pnp_ms * 1e-3), "ok_fraction": float(ok_views.mean()),
                    "note": "kb_estimate_transformations through the C ABI incl. the device-to-host copy of the poses"}
            try:
                import cv2
                from oracle import ko_init as ki

                n_s = min(p.n_views, 400)
                inputs = []
                for w in range(n_s):
                    b, e = p.view_begin[w], p.view_begin[w + 1]
                    k = p.view_cam[w]
                    Ps, Ms = ki.pnp_inputs(p.cam_model[k], p.cam_params[k], p.y_u[b:e], p.y_v[b:e], p.target_points[p.corner_id[b:e]])
                    inputs.append((np.float32(Ps), np.float32(Ms)))
                K, dz = np.eye(3), np.zeros(4)
                t0 = time.time()
                n_done = 0
                while time.time() - t0 < 3.0:
                    for Ps, Ms in inputs:
                        cv2.solvePnP(Ps, Ms, K, dz)
                    n_done += len(inputs)
                el = time.time() - t0
                init["cpu_baseline"] = {"value": n_done / el, "unit": "views/s", "cores": 1, "kind": "reference",
                                        "sample": f"cv2.solvePnP {cv2.__version__} (the library call inside estimateTransformation) on the "
                                                  f"prepared corners of {n_s} views, repeated for 3 s, one thread as in the reference's drivers"}
            except Exception as ex:
                init["cpu_baseline"] = {"error": str(ex)}
        except Exception as ex:
            init = {"error": str(ex)}

    if rank == 0:
        hbm_peak, hbm_src, fp64_peak, fp64_src = peaks()
        la_ms, la_calls = totals["linearise_assemble"]
        la_ms = la_ms / max(la_calls, 1)
        n_views_rank = p.n_views
        flops = FLOP_PER_TERM * terms_rank
        byts = (BYTES_IN_PER_TERM + BYTES_E_PER_TERM) * terms_rank + BYTES_OUT_PER_VIEW * n_views_rank
        traffic = (read_json(os.path.join(ROOT, "profiles", "r01_traffic.json")) or {}).get("linearise_assemble_dram_bytes_per_launch")
        roofline = {"kernel": "linearise_assemble_kernel<pinhole-radtan>", "bound": "tensor",
                    "achieved": flops / (la_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": flops / (la_ms * 1e-3) / 1e12 / fp64_peak, "traffic": traffic,
                    "peak_source": fp64_src, "avg_launch_ms": la_ms, "launches_timed": la_calls,
                    "note": "FP64 pipe bound (about 24 flop/B): DMMA m8n8k4 Gram accumulation + FP64 projection/Jacobian math"}
        roofline_hbm = {"kernel": "linearise_assemble_kernel<pinhole-radtan>", "bound": "hbm", "achieved": byts / (la_ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": byts / (la_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "algorithmic_bytes_per_launch": byts}
        stages = {k: (v[0] / max(v[1], 1)) for k, v in totals.items() if v[1] > 0}
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle_api as oa

            cores = os.cpu_count() or 1
            ps = synthetic.make_config(args.config, n_sets=args.cpu_sets)
            o = oa.OracleProblem(ps, oa.BLOCK_CHOLESKY, n_threads=cores)
            o.evaluate_error()
            n_it, t0, st = 0, time.time(), np.zeros(3)
            while n_it < 2 or (time.time() - t0 < 10.0 and n_it < 50):
                t, _ = o.time_iteration(10.0)
                st += t
                n_it += 1
            el = time.time() - t0
            cpu = {"value": ps.n_terms * n_it / el, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"cfg{args.config} restricted to {args.cpu_sets} synced sets ({ps.n_terms} terms), {n_it} LM iterations (evaluate+build+solve), "
                             f"BlockCholesky semantic (threaded evaluate on {cores} threads, serial assemble as in the reference)",
                   "stage_s_per_iteration": {"evaluate": st[0] / n_it, "build": st[1] / n_it, "solve": st[2] / n_it}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, S_rank, terms_rank, world),
            "per_gpu_value": value / world, "lm_iteration_ms": ms_total / args.steps, "terms_total": terms_total,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_hbm": roofline_hbm,
            "linearise_materialised": lin, "initial_guess": init, "stage_ms": stages, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        g.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
