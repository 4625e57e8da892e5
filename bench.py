#!/usr/bin/env python
"""Benchmark of the B200 batch-calibration hot path (contract: see the task statement / DESIGN.md §6).

One "step" = one Levenberg-Marquardt iteration's worth of hot path on one batch of synthetic observations:
  evaluate error -> linearise + assemble -> set conditioner -> Schur-reduced solve -> rho denominator ->
  apply state update -> revert (so every step does identical work from the same state).
Workload: BASELINE.json configs[3] "8-camera rig, 20k views": 8 x pinhole-radtan, 20 000 synced sets, 19.2 M
reprojection terms.  N>1 (torchrun): the headline is STRONG scaling - the one 20k-set problem sharded by synced set
over the ranks, as BASELINE configs[3] / configs[4] describe it; the same line also carries `weak_scaling` (20 000
sets per rank) and `cfg5_strong` (configs[4]: 16 cameras, 6 250 sets, 12 M terms, sharded the same way).
`--scaling weak` makes the weak run the headline instead.

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


def lm_step(g, lam=10.0, fetch_dx=False, host_obs=None, wait=True):
    if host_obs is None and not fetch_dx:
        # the device-resident arm: the same six calls through kb_iterate (enqueued back to back; wait=False: no host synchronisation
        # per step - the K steps are enqueued ahead of the device and waited for once, kb_wait)
        return g.iterate(lam, revert=True, wait=wait)
    # host_obs: (y_u, y_v) pinned host arrays uploaded as part of the evaluation (end-to-end leg)
    J = g.evaluate_error_streamed(*host_obs) if host_obs is not None else g.evaluate_error()
    g.build_system()
    g.set_constant_conditioner(lam)
    dx, ok = g.solve_system(fetch_dx=fetch_dx, gather=False)
    rho = g.lm_rho_denominator(lam)
    m = g.apply_state_update()
    g.revert_last_state_update()
    return J, ok, rho, m, dx


CPU_REGIMES = [("sparse", "all"), ("sparse", 4), ("block", "all"), ("block", 4)]  # (solver semantic, threads); headline = the FASTEST of them


def cpu_regimes(config, cpu_sets, steps_each, warmup=1):
    """The oracle (restatement of the reference's aslam_backend path; the reference as a whole cannot be built here: no Eigen / Boost /
    SuiteSparse in the image - its evaluate + build compiled against stand-in headers is the extra "reference" regime below) on the host cores, in the reference's two solver regimes - SparseCholesky (Kalibr2's DEFAULT:
    Optimizer2.cpp:83-86; Jacobian materialisation threaded) and BlockCholesky (serial assembly) - at T = 4 threads (the reference's
    default nThreads, Optimizer2Options.hpp:16) and T = all host cores.  One step = evaluate + build + solve at lambda = 10 on a bounded
    sample of the workload.  The HEADLINE CPU number is the fastest of all regimes (these four and the reference-code ones below), so that the GPU / CPU ratio is never inflated by a slow
    regime: in this port the SparseCholesky solve assembles J^T J entry by entry (the reference hands J^T to CHOLMOD's supernodal
    factorisation, which is not in the image), which makes the BlockCholesky regime the faster one here.
    Returns (per-regime results sorted as CPU_REGIMES, index of the fastest, problem of the sample)."""
    from kalibr_b200 import synthetic
    from oracle import oracle_api as oa

    cores = os.cpu_count() or 1
    ps = synthetic.make_config(config, n_sets=cpu_sets)
    out = []
    for i, (regime, T) in enumerate(CPU_REGIMES):
        threads = cores if T == "all" else min(int(T), cores)
        t0 = time.time()
        o = oa.OracleProblem(ps, oa.BLOCK_CHOLESKY if regime == "block" else oa.SPARSE_CHOLESKY, n_threads=threads)
        build_s = time.time() - t0
        o.evaluate_error()
        for _ in range(warmup):
            o.time_iteration(10.0)
        steps = steps_each
        st = np.zeros(3)
        t0 = time.time()
        for _ in range(steps):
            t, _ok = o.time_iteration(10.0)
            st += t
        el = time.time() - t0
        out.append({"regime": "SparseCholesky (reference default: threaded Jacobian materialisation)" if regime == "sparse" else "BlockCholesky (threaded evaluate, serial assembly)",
                    "solver": regime, "threads": threads, "value": ps.n_terms * steps / el, "unit": UNIT, "ms_per_step": 1e3 * el / steps, "steps": steps,
                    "stage_s_per_iteration": {"evaluate": st[0] / steps, "build": st[1] / steps, "solve": st[2] / steps},
                    "problem_construction_s": build_s})
        o.close()
    # the REFERENCE's own evaluate + build (Optimizer2::evaluateError, then BlockCholeskyLinearSystemSolver::buildSystem or
    # SparseCholeskyLinearSystemSolver::buildSystem with its CompressedColumnJacobianTransposeBuilder; the expression tree, JacobianContainer,
    # SparseBlockMatrix and CompressedColumnMatrix compiled from the reference's sources into oracle/_ref, against stand-in Eigen / Boost
    # headers: oracle/ref_pin_optimizer.cpp) on the same sample, when that prebuilt library travelled here.  Its solve is not
    # reference code in this image (no CHOLMOD), so the step takes the port's solve time of the same regime.  The same code that pins the
    # oracle's numbers (tests/test_reference_pin_cpu.py) here shows the port is not a slow stand-in for it.
    if oa.build_reference_cameras() is not None:
        for solver, kind_id, what in (("block", oa.BLOCK_CHOLESKY_KIND, "BlockCholesky (serial Hessian assembly)"),
                                      ("sparse", oa.SPARSE_CHOLESKY_KIND, "SparseCholesky (Kalibr2's default: threaded materialisation of the compressed-column J^T)")):
            port_solve = min(r["stage_s_per_iteration"]["solve"] for r in out if r["solver"] == solver and r.get("kind") != "reference")
            try:
                t = reference_evaluate_build_isolated(config, cpu_sets, cores, 1, kind_id)  # one timed repeat after the child's serial first pass
                step_s = t["evaluate_s"] + t["build_s"] + port_solve
                out.append({"regime": f"{what} through the reference's own compiled evaluate + build (oracle/_ref; stand-in Eigen / Boost headers; solve time from the port)",
                            "solver": solver, "kind": "reference", "config": config, "threads": cores, "value": ps.n_terms / step_s, "unit": UNIT, "ms_per_step": 1e3 * step_s,
                            "steps": 1, "stage_s_per_iteration": {"evaluate": t["evaluate_s"], "build": t["build_s"], "solve": port_solve},
                            "problem_construction_s": t["setup_s"]})
            except Exception as e:  # the checker's library is optional on the box; the port regimes above always run
                out.append({"regime": f"{what} through the reference's own compiled evaluate + build", "kind": "reference", "unavailable": repr(e), "value": 0.0,
                            "solver": solver, "threads": cores})
    best = max(range(len(out)), key=lambda i: out[i]["value"])
    return out, best, ps


def reference_evaluate_build_isolated(config, cpu_sets, threads, repeats, solver_kind):
    """oracle/reference_timing.py in a process of its own (the reference's code over stand-in headers, threaded: a failure there must not
    take the bench with it): dict(setup_s, evaluate_s, build_s, cost); raises when the child fails or the library is not there"""
    import subprocess

    r = subprocess.run([sys.executable, "-m", "oracle.reference_timing", str(config), str(cpu_sets), str(threads), str(repeats), str(solver_kind)],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError(f"oracle.reference_timing exited with {r.returncode}: {r.stderr.strip()[-200:]}")
    t = json.loads(r.stdout.strip().splitlines()[-1])
    if "unavailable" in t:
        raise RuntimeError(t["unavailable"])
    return t


def time_cpu_regime(head, ps, steps, warmup):
    """EXACTLY `steps` timed steps of one regime of cpu_regimes (after `warmup`): (terms/s, ms per step, stage seconds per iteration)"""
    from oracle import oracle_api as oa

    K = max(steps, 1)
    if head.get("kind") == "reference":
        try:
            t = reference_evaluate_build_isolated(head["config"], int(ps.n_sets), head["threads"], K, oa.SPARSE_CHOLESKY_KIND if head["solver"] == "sparse" else oa.BLOCK_CHOLESKY_KIND)
        except Exception:  # the child failed on the re-timing: keep what the survey measured for this regime
            return head["value"], head["ms_per_step"], head["stage_s_per_iteration"]
        solve = head["stage_s_per_iteration"]["solve"]
        step_s = t["evaluate_s"] + t["build_s"] + solve
        return ps.n_terms / step_s, 1e3 * step_s, {"evaluate": t["evaluate_s"], "build": t["build_s"], "solve": solve}
    o = oa.OracleProblem(ps, oa.BLOCK_CHOLESKY if head["solver"] == "block" else oa.SPARSE_CHOLESKY, n_threads=head["threads"])
    o.evaluate_error()
    for _ in range(warmup):
        o.time_iteration(10.0)
    st = np.zeros(3)
    t0 = time.time()
    for _ in range(K):
        t, _ok = o.time_iteration(10.0)
        st += t
    el = time.time() - t0
    o.close()
    return ps.n_terms * K / el, 1e3 * el / K, {"evaluate": st[0] / K, "build": st[1] / K, "solve": st[2] / K}


def cpu_sample_text(config, ps, cores):
    return (f"bounded sample of the workload: cfg{config} restricted to {ps.n_sets} of its synced sets ({ps.n_terms} terms); one step = evaluate + "
            f"build + solve(lambda = 10); regimes: SparseCholesky (Kalibr2's default) / BlockCholesky semantic x 4 threads (the reference's default) / "
            f"{cores} threads (all host cores), each through the oracle port, plus the evaluate + build of both regimes through the reference's own compiled code "
            f"(oracle/_ref, kind 'reference') when that library is present; headline = the fastest of them; throughput metric, so the sample size does not enter the unit")


def run_reference(args):
    """CPU arm of the driver's ratio: same metric, unit and `config` (the workload) as the b200 arm; every step is a bounded sample of
    that workload (cpu_baseline.sample says which)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    from kalibr_b200 import synthetic

    _, _, S_cfg = synthetic.CONFIGS[args.config]
    # survey: two steps of each regime; then EXACTLY K timed steps of the fastest one (after W warm-up steps, at most one)
    regimes, best, ps = cpu_regimes(args.config, args.cpu_sets, 2, warmup=0)
    from oracle import oracle_api as oa

    head = dict(regimes[best])
    K = max(args.steps, 1)
    value, ms, stages = time_cpu_regime(head, ps, K, min(args.warmup, 1))
    head.update({"value": value, "ms_per_step": ms, "steps": K, "stage_s_per_iteration": stages})
    line = {
        "impl": "reference", "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling or ("strong" if args.gpus > 1 else "weak"),
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus, args.scaling or ("strong" if args.gpus > 1 else "weak")),
        "cpu_baseline": {"value": head["value"], "unit": UNIT, "cores": head["threads"], "kind": head.get("kind", "port"), "cpu_sets": int(ps.n_sets),
                         "sample": cpu_sample_text(args.config, ps, cores), "regimes": regimes,
                         "headline_regime": f"{head['solver']} / {head['threads']} threads / {head.get('kind', 'port')} (fastest of the {len(regimes)})"},
        "e2e": {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


L2_BYTES = 126e6


def workload_config(args, n_ranks, scaling):
    """`config` of the JSON line: a function of the command line alone, so that the b200 arm and the reference arm (which runs a bounded
    sample of the same workload on the host) print the identical object."""
    from kalibr_b200 import synthetic
    from kalibr_b200.problem import MODEL_NAMES

    order, models, S_cfg = synthetic.CONFIGS[args.config]
    S_total = args.sets if args.sets is not None else S_cfg
    per_set = 120 * len(models)
    if n_ranks > 1 and scaling == "strong":
        lo, hi = synthetic.shard_sets(S_total, n_ranks, 0)
        sets_rank0 = hi - lo
    else:
        sets_rank0 = S_total
    flush = sets_rank0 * per_set * 18 <= L2_BYTES * 1.5
    cfg = {
        "workload": f"BASELINE.json configs[{args.config - 1}]: {len(models)}-camera rig ({', '.join(sorted(set(MODEL_NAMES[m] for m in models)))}), "
                    f"6x5 aprilgrid (120 corners), {S_total} synced sets" + (" per rank" if n_ranks > 1 and scaling == "weak" else "")
                    + f", {S_total * per_set} reprojection terms" + (" per rank" if n_ranks > 1 and scaling == "weak" else ""),
        "cameras": len(models), "synced_sets": S_total,
        "parallelism": (f"{scaling} scaling: synced sets sharded over {n_ranks} ranks ({sets_rank0} sets on rank 0); per solve one exchange of the "
                        "reduced camera system " + ("(NCCL all-reduce)" if args.no_peer_exchange or n_ranks > 8 else "(NVLink peer stores fused into the producing kernel, summed by the consumer)"))
        if n_ranks > 1 else "single GPU",
        "l2": "L2 flushed between steps (a 256 MiB scratch buffer is overwritten before every step; steps timed one by one)" if flush
              else "inputs larger than L2 (the observations of a rank alone exceed 126 MB)",
    }
    return cfg


def flush_needed(terms_rank):
    return terms_rank * 18 <= L2_BYTES * 1.5





class Harness:
    """torch.distributed plumbing of one bench process (one rank per GPU): barrier, device-side timing, NCCL id, peer handles."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.flush_buf = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.int64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def timed(self, stream, fn, steps, flush_l2=False, finish=None):
        """K steps timed on the device (CUDA events on the library's stream) between barrier + synchronize on both sides, MAX over
        ranks.  flush_l2: the per-rank working set fits the 126 MB L2, so a 256 MiB buffer is overwritten before every step and
        every step is timed on its own (the flush stays outside the timed intervals)."""
        torch = self.torch
        self.barrier()
        if not flush_l2:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            if finish:
                finish()
            self.barrier()
            return self.max_over_ranks(e0.elapsed_time(e1))
        if self.flush_buf is None:
            self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        evs = []
        for _ in range(steps):
            with torch.cuda.stream(stream):
                self.flush_buf.add_(1)  # read + write of 256 MiB: evicts the L2
            if self.world > 1:
                self.dist.barrier()      # every rank starts the step together (the exchange steps wait for the slowest rank anyway)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            if finish:
                finish()
            evs.append((e0, e1))
        self.barrier()
        # a step ends when its slowest rank ends: max over ranks per step, then the sum
        per_step = torch.tensor([a.elapsed_time(b) for a, b in evs], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(per_step, op=self.dist.ReduceOp.MAX)
        return float(per_step.sum().item())

    def make_solver(self, config, scaling, sets_total):
        """Problem + handle of this rank for `config` under weak (sets_total per rank) or strong (sets_total sharded) scaling."""
        from kalibr_b200 import capi, synthetic

        torch, dist, world, rank = self.torch, self.dist, self.world, self.rank
        if scaling == "weak":
            S_rank, set_offset, S_total = sets_total, rank * sets_total, sets_total * world
        else:
            lo, hi = synthetic.shard_sets(sets_total, world, rank)
            S_rank, set_offset, S_total = hi - lo, lo, sets_total
        # measurements as a corner detector delivers them: float32-representable (the reference's cv::Point2f), held as doubles
        p = synthetic.make_config(config, n_sets=S_rank, set_seed=20260000 + 100 * config + 7919 + set_offset, float_corners=True)
        terms_total = self.sum_over_ranks(p.n_terms)
        nccl_id = None
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(idt, 0)
            nccl_id = bytes(idt.cpu().tolist())
        g = capi.B200SchurLinearSystemSolver(p, n_ranks=world, rank=rank, nccl_id=nccl_id, device=self.local_rank,
                                             n_sets_total=S_total if world > 1 else 0, set_offset=set_offset, n_terms_total=terms_total)
        if world > 1 and world <= 8 and not self.args.no_peer_exchange:
            # NVLink peer exchange: all-gather the CUDA IPC handles of the ranks' exchange buffers, then attach
            mine = torch.tensor(list(g.peer_exchange_handle()), dtype=torch.uint8, device="cuda")
            allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            g.attach_peers(b"".join(bytes(t.cpu().tolist()) for t in allh))
        stream = torch.cuda.ExternalStream(g.cuda_stream(), device=torch.device("cuda", self.local_rank))
        return p, g, stream, terms_total

    def run_steps(self, config, scaling, sets_total, steps, warmup, sample_clocks=False):
        """The device-resident arm for one (config, scaling): W warm-up steps, K timed steps.  Returns a dict and keeps (p, g, stream)."""
        p, g, stream, terms_total = self.make_solver(config, scaling, sets_total)
        for _ in range(warmup):
            lm_step(g)
        sampler = ClockSampler(self.local_rank) if sample_clocks and self.rank == 0 else None
        if sampler:
            sampler.start()
        flush = bool(self.max_over_ranks(1.0 if flush_needed(p.n_terms) else 0.0))  # the same decision on every rank
        # inside the timed region only the roofline kernel is bracketed by a CUDA event pair (the events of the other stages would sit
        # between kernels that are launched programmatically dependent); the full per-stage breakdown comes from 5 steps right after it
        g.enable_stage_timing(True, stages=["linearise_assemble"])
        l0 = g.kernel_launches()
        ms_total = self.timed(stream, lambda: lm_step(g, wait=False), steps, flush_l2=flush, finish=g.wait_iterations)
        launches = self.sum_over_ranks(g.kernel_launches() - l0)
        totals = g.stage_totals()
        g.enable_stage_timing(True)
        for _ in range(5):
            lm_step(g)
        totals_all = g.stage_totals()
        g.enable_stage_timing(False)
        totals = {k: (totals[k] if k == "linearise_assemble" else totals_all[k]) for k in totals_all}
        clocks = sampler.stop() if sampler else None
        return {"p": p, "g": g, "stream": stream, "terms_total": terms_total, "ms_total": ms_total, "launches": launches, "totals": totals,
                "clocks": clocks, "l2_flushed": flush, "value": terms_total * steps / (ms_total * 1e-3), "ms_per_step": ms_total / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, help="BASELINE config 1..5 (default 4: 8-camera rig, 20k views)")
    ap.add_argument("--sets", type=int, default=None, help="override the number of synced sets of the workload")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="N > 1: which scaling the headline value is (default strong)")
    ap.add_argument("--cpu-sets", type=int, default=None, help="synced sets of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-init-stage", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the weak-scaling and cfg-5 extra measurements")
    ap.add_argument("--no-peer-exchange", action="store_true", help="N > 1: keep the exchange steps on NCCL instead of NVLink peer stores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    from kalibr_b200 import synthetic

    order, models, S_cfg = synthetic.CONFIGS[args.config]
    if args.cpu_sets is None:
        # about 10-30 s of CPU work for the four port regimes (the two reference-code regimes add one timed repeat each): the oracle runs ~0.1-0.2 M terms/s per LM iteration
        per_set = 120 * len(models)
        args.cpu_sets = max(8, min(S_cfg, int((400_000 if args.impl == "reference" else 200_000) / per_set)))
    if args.impl == "reference":
        return run_reference(args)

    from kalibr_b200 import capi

    capi.load_library()
    H = Harness(args)
    torch, world, rank = H.torch, H.world, H.rank
    if world != args.gpus and world > 1:
        args.gpus = world
    scaling = args.scaling or ("strong" if world > 1 else "weak")
    S_workload = args.sets if args.sets is not None else S_cfg

    # ---- headline: device-resident arm ----
    R = H.run_steps(args.config, scaling, S_workload, args.steps, args.warmup, sample_clocks=True)
    p, g, stream = R["p"], R["g"], R["stream"]
    terms_rank, terms_total = p.n_terms, R["terms_total"]
    ms_total, launches, totals, clocks, value = R["ms_total"], R["launches"], R["totals"], R["clocks"], R["value"]

    # ---- a whole LM run on the (sharded) workload: kb_optimize = the device-resident loop, iterations replayed from a CUDA graph ----
    lm_loop = None
    try:
        from kalibr_b200.problem import KbOptimizerOptions

        g.reset_state()
        g.optimize(KbOptimizerOptions.kalibr2_default())  # first run: plain launches + graph capture
        g.reset_state()
        box = {}
        ms_opt = H.timed(stream, lambda: box.update(sol=g.optimize(KbOptimizerOptions.kalibr2_default())[0]), 1)
        sol = box["sol"]
        n_it = sol.iterations + sol.failed_iterations
        lm_loop = {"ms_total": ms_opt, "iterations": sol.iterations, "failed_iterations": sol.failed_iterations,
                   "ms_per_iteration": ms_opt / max(n_it, 1), "j_start": sol.j_start, "j_final": sol.j_final,
                   "note": "kb_optimize (Optimizer2::optimize + LM policy) to convergence on the sharded workload, incl. the initial evaluation; device events, max over ranks"}
        g.reset_state()
    except Exception as ex:
        lm_loop = {"error": str(ex)}

    # ---- end-to-end arm: host buffers in, host results out, every step ----
    e2e = None
    if not args.no_e2e:
        def e2e_arm(yu_host, yv_host, bytes_per_term):
            """K steps through the public entry points with the observations in pinned HOST memory, uploaded every step, and dx fetched
            every step.  Returns (double-buffered ms per step, isolated streamed-step ms)."""
            yu_pin, yv_pin = torch.from_numpy(yu_host).pin_memory(), torch.from_numpy(yv_host).pin_memory()
            yu_np, yv_np = yu_pin.numpy(), yv_pin.numpy()

            def step_streamed():
                # H2D of the observations, pipelined in chunks against the fused kernel inside the library; D2H: cost, pos-def flag,
                # dx (jcols doubles), rho, max|dx|
                lm_step(g, fetch_dx=True, host_obs=(yu_np, yv_np))

            def step_pipelined():
                # double buffering across steps: this step's batch was uploaded while the previous step computed; the next step's
                # batch starts travelling now (one upload per step inside the timed region either way)
                g.commit_observations()
                g.prefetch_observations(yu_np, yv_np)
                lm_step(g, fetch_dx=True)

            for _ in range(2):
                step_streamed()
            ms_streamed = H.timed(stream, step_streamed, k_e2e)
            g.prefetch_observations(yu_np, yv_np)
            for _ in range(2):
                step_pipelined()

            def pipelined_run():
                for _ in range(k_e2e):
                    step_pipelined()
                g.commit_observations()               # the last upload is waited for inside the timed region
                g.prefetch_observations(yu_np, yv_np)

            ms = H.timed(stream, pipelined_run, 1)
            g.commit_observations()
            return ms / k_e2e, ms_streamed / k_e2e, int(bytes_per_term * terms_rank)

        k_e2e = max(3, args.steps // 2)
        yu32, yv32 = p.y_u.astype(np.float32), p.y_v.astype(np.float32)
        assert np.array_equal(yu32.astype(np.float64), p.y_u) and np.array_equal(yv32.astype(np.float64), p.y_v)  # exact: same values, half the bytes
        ms64, lat64, b64 = e2e_arm(p.y_u, p.y_v, 16)
        ms32, lat32, b32 = e2e_arm(yu32, yv32, 8)
        e2e = {"value": terms_total / (ms32 * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": b32, "d2h_bytes_per_step": int(8 * g.jcols + 8 * 4 + 4),
               "ms_per_step": ms32, "steps": k_e2e, "single_step_latency_ms": lat32,
               "host_buffers": "float32 (the corner detector's type, cv::Point2f in the reference; widened exactly on the device)",
               "f64_host_buffers": {"value": terms_total / (ms64 * 1e-3), "ms_per_step": ms64, "single_step_latency_ms": lat64, "h2d_bytes_per_step": b64},
               "note": "bytes per rank; through B200SchurLinearSystemSolver (C ABI) with pinned HOST observation buffers uploaded every step and dx fetched "
                       "every step. value: double-buffered (kb_prefetch_observations_f32 / kb_commit_observations: the next step's upload overlaps this "
                       "step's kernels); single_step_latency_ms: one isolated step with kb_evaluate_error_streamed_f32 (chunked upload overlapped with "
                       "the fused kernel). Both are bounded by the PCIe transfer of the measurements (8 B/term as float32, 16 B/term as float64)"}

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

    g.close()
    del g

    # ---- N > 1: the other scaling mode and the cfg-5 sweep point, same harness ----
    extras = {}
    if world > 1 and not args.no_extras:
        other = "weak" if scaling == "strong" else "strong"
        k_x, w_x = max(5, args.steps // 2), 3
        X = H.run_steps(args.config, other, S_workload, k_x, w_x)
        extras[f"{other}_scaling"] = {"value": X["value"], "unit": UNIT, "ms_per_step": X["ms_per_step"], "steps": k_x, "terms_total": X["terms_total"],
                                      "synced_sets_rank0": int(X["p"].n_sets), "l2_flushed_between_steps": X["l2_flushed"],
                                      "note": "weak: every rank holds the full workload's number of synced sets; strong: the one workload sharded by synced set"}
        X["g"].close()
        del X
    if not args.no_extras and args.config == 4:
        k_x, w_x = max(5, args.steps // 2), 3
        X = H.run_steps(5, "strong", synthetic.CONFIGS[5][2], k_x, w_x)
        extras["cfg5_strong"] = {"workload": "BASELINE.json configs[4]: 16-camera rig (pinhole-radtan), 6250 synced sets, 12000000 terms (n_c = 218)",
                                 "value": X["value"], "unit": UNIT, "ms_per_step": X["ms_per_step"], "steps": k_x, "terms_total": X["terms_total"],
                                 "l2_flushed_between_steps": X["l2_flushed"],
                                 "stage_ms": {k: (v[0] / max(v[1], 1)) for k, v in X["totals"].items() if v[1] > 0}}
        X["g"].close()
        del X

    # ---- N = 1: calibration-level end to end (create -> optimize to convergence -> fetch), next to the oracle's optimize ----
    calib = None
    if rank == 0 and world == 1 and not args.no_e2e:
        calib = calibration_e2e(args, capi, synthetic, S_workload)

    if rank == 0:
        hbm_peak, hbm_src, fp64_peak, fp64_src = peaks()
        la_ms, la_calls = totals["linearise_assemble"]
        la_ms = la_ms / max(la_calls, 1)
        n_views_rank = p.n_views
        flops = FLOP_PER_TERM * terms_rank
        byts = (BYTES_IN_PER_TERM + BYTES_E_PER_TERM) * terms_rank + BYTES_OUT_PER_VIEW * n_views_rank
        traffic = (read_json(os.path.join(ROOT, "profiles", "r02_traffic.json")) or {}).get("linearise_assemble_dram_bytes_per_launch") if world == 1 and args.config == 4 and args.sets is None else None
        roofline = {"kernel": "linearise_assemble_kernel<pinhole-radtan>", "bound": "tensor",
                    "achieved": flops / (la_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": flops / (la_ms * 1e-3) / 1e12 / fp64_peak, "traffic": traffic,
                    "peak_source": fp64_src, "avg_launch_ms": la_ms, "launches_timed": la_calls,
                    "note": "rank 0's launch; FP64 pipe bound (about 24 flop/B): DMMA m8n8k4 Gram accumulation + FP64 projection/Jacobian math"}
        roofline_hbm = {"kernel": "linearise_assemble_kernel<pinhole-radtan>", "bound": "hbm", "achieved": byts / (la_ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": byts / (la_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "algorithmic_bytes_per_launch": byts}
        stages = {k: (v[0] / max(v[1], 1)) for k, v in totals.items() if v[1] > 0}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            regimes, best, ps = cpu_regimes(args.config, args.cpu_sets, 2, warmup=0)
            head = regimes[best]
            cpu = {"value": head["value"], "unit": UNIT, "cores": head["threads"], "kind": head.get("kind", "port"), "cpu_sets": int(ps.n_sets),
                   "sample": cpu_sample_text(args.config, ps, cores), "regimes": regimes,
                   "headline_regime": f"{head['solver']} / {head['threads']} threads / {head.get('kind', 'port')} (fastest of the {len(regimes)})"}
        cfg = workload_config(args, world, scaling)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "per_gpu_value": value / world, "lm_iteration_ms": ms_total / args.steps, "terms_total": terms_total,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_hbm": roofline_hbm,
            "lm_loop": lm_loop, "linearise_materialised": lin, "initial_guess": init, "stage_ms": stages, "cpu_baseline": cpu, "calibration_e2e": calib,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        H.dist.barrier()
        H.dist.destroy_process_group()


def calibration_e2e(args, capi, synthetic, S_workload):
    """A whole calibration through the public entry points, host arrays in, host parameters out: kb_create (uploads the problem once)
    -> kb_optimize (device-resident LM loop to convergence) -> kb_get_camera_params / kb_get_baselines.  At the workload's full size,
    and on the bounded CPU sample next to the oracle's Optimizer2::optimize restatement (same problem, same options)."""
    from kalibr_b200.problem import KbOptimizerOptions

    def gpu_run(p):
        t0 = time.time()
        g = capi.B200SchurLinearSystemSolver(p)
        t1 = time.time()
        sol, _ = g.optimize(KbOptimizerOptions.kalibr2_default())
        cam, base = g.camera_params(), g.baselines()
        t2 = time.time()
        g.close()
        return {"create_ms": 1e3 * (t1 - t0), "optimize_and_fetch_ms": 1e3 * (t2 - t1), "total_ms": 1e3 * (t2 - t0), "iterations": sol.iterations,
                "failed_iterations": sol.failed_iterations, "j_final": sol.j_final, "terms": int(p.n_terms),
                "terms_iterations_per_s": p.n_terms * (sol.iterations + sol.failed_iterations) / max(t2 - t0, 1e-9)}, cam, base

    out = {"note": "wall clock of the calling process (time.time), not device events: this is what a caller of the C ABI sees"}
    try:
        p_full = synthetic.make_config(args.config, n_sets=S_workload)
        gpu_run(p_full)  # first call pays module load / graph capture
        out["gpu_full_size"], _, _ = gpu_run(p_full)
        if not args.no_cpu_baseline:
            from oracle import oracle_api as oa

            ps = synthetic.make_config(args.config, n_sets=max(8, args.cpu_sets // 2))
            out["gpu_cpu_sample"], cam, base = gpu_run(ps)
            cores = os.cpu_count() or 1
            t0 = time.time()
            o = oa.OracleProblem(ps, oa.SPARSE_CHOLESKY if args.config == 3 else oa.BLOCK_CHOLESKY, n_threads=cores)
            t1 = time.time()
            osol, _ = o.optimize(KbOptimizerOptions.kalibr2_default())
            t2 = time.time()
            oc = o.camera_params()
            out["cpu_sample"] = {"kind": "port", "cores": cores, "sets": int(ps.n_sets), "terms": int(ps.n_terms), "construct_ms": 1e3 * (t1 - t0),
                                 "optimize_ms": 1e3 * (t2 - t1), "total_ms": 1e3 * (t2 - t0), "iterations": osol.iterations,
                                 "failed_iterations": osol.failed_iterations, "j_final": osol.j_final,
                                 "max_rel_parameter_difference_vs_gpu": float((np.abs(cam - oc) / np.maximum(np.abs(oc), 1e-3)).max())}
            o.close()
    except Exception as ex:
        out["error"] = str(ex)
    return out


if __name__ == "__main__":
    main()
