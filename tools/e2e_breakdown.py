"""Where the end-to-end step goes: upload alone, streamed evaluate, resident evaluate, the rest of the LM step (cfg 4)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kalibr_b200 import capi, synthetic

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
p = synthetic.make_config(cfg)
g = capi.B200SchurLinearSystemSolver(p)
yu = torch.from_numpy(p.y_u).pin_memory().numpy()
yv = torch.from_numpy(p.y_v).pin_memory().numpy()
stream = torch.cuda.ExternalStream(g.cuda_stream())

def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def upload():
    g.set_observations(yu, yv)
    torch.cuda.synchronize()

def rest(fetch):
    g.build_system(); g.set_constant_conditioner(10.0)
    g.solve_system(fetch_dx=fetch, gather=False); g.lm_rho_denominator(10.0); g.apply_state_update(); g.revert_last_state_update()

print("upload only            %.3f ms  (%.1f GB/s)" % ((t := timed(upload)), 16 * p.n_terms / t / 1e6))
print("evaluate resident      %.3f ms" % timed(lambda: g.evaluate_error()))
print("evaluate streamed      %.3f ms" % timed(lambda: g.evaluate_error_streamed(yu, yv)))
print("upload + evaluate      %.3f ms" % timed(lambda: (g.set_observations(yu, yv), g.evaluate_error())))
print("rest of step (no dx)   %.3f ms" % timed(lambda: rest(False)))
print("rest of step (dx D2H)  %.3f ms" % timed(lambda: rest(True)))
