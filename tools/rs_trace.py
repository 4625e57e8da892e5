import sys; sys.path.insert(0,'.')
from kalibr_b200 import capi, synthetic
for cfg,S in [(4,400),(5,200),(3,400)]:
    p=synthetic.make_config(cfg,n_sets=S); g=capi.B200SchurLinearSystemSolver(p)
    g.evaluate_error(); g.build_system(); g.set_constant_conditioner(10.0)
    for _ in range(3): g.solve_system()
    g.close()
