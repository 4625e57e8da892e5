"""Step-by-step GPU-vs-oracle check with progress output (debug helper; the real gates are tests/test_parity_gpu.py)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from kalibr_b200 import synthetic, capi
from kalibr_b200.problem import KbOptimizerOptions
from oracle import oracle_api as oa

def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))

def say(*a):
    print(*a, flush=True)

cases = [(1, 40), (2, 30), (3, 24), (4, 12), (5, 6)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]]
for cfg, S in cases:
    p = synthetic.make_config(cfg, n_sets=S)
    say(f"== cfg {cfg} S={S} terms={p.n_terms} n_c={p.n_c}")
    g = capi.B200SchurLinearSystemSolver(p)
    o = oa.OracleProblem(p)
    Jg = g.evaluate_error(); Jo = o.evaluate_error()
    say("  cost", Jg, Jo, "e rel", rel(g.error_vector(), o.error_vector()))
    gp, gi, gv = g.jacobian_ccs(); op, oi, ov = o.jacobian_ccs()
    say("  J ccs ptr/idx equal", np.array_equal(gp, op), np.array_equal(gi, oi), "val rel", rel(gv, ov))
    g.build_system(); o.build_system()
    say("  rhs rel", rel(g.rhs(), o.rhs()))
    g.set_constant_conditioner(10.0); o.set_constant_conditioner(10.0)
    gdx, gok = g.solve_system(); odx, ook = o.solve_system()
    say("  solve ok", gok, ook, "dx rel", rel(gdx, odx))
    gh = g.hessian_blocks(); oh = o.hessian_blocks()
    say("  H pattern equal", all(np.array_equal(a, b) for a, b in zip(gh[:3], oh[:3])), "val rel", rel(gh[3], oh[3]) if gh[3].shape == oh[3].shape else "shape!")
    say("  rho", g.lm_rho_denominator(10.0), float(odx @ (10.0 * odx + o.rhs())))
    g.reset_state()
    o2 = oa.OracleProblem(p)
    t = time.time(); gs, gtr = g.optimize(KbOptimizerOptions.kalibr2_default()); tg = time.time() - t
    t = time.time(); os_, otr = o2.optimize(KbOptimizerOptions.kalibr2_default()); to = time.time() - t
    say("  optimize gpu", gs.as_dict(), f"{tg:.3f}s")
    say("  optimize cpu", os_.as_dict(), f"{to:.3f}s")
    say("  params rel", rel(g.camera_params(), o2.camera_params()), "poses rel", rel(g.set_poses(), o2.set_poses()))
    say("  launches", g.kernel_launches(), "invalid", g.num_invalid_terms())
say("DONE")
