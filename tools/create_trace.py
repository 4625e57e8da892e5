import sys; sys.path.insert(0,'.')
from kalibr_b200 import capi, synthetic
for cfg,S in ((4,12),(2,40)):
    p = synthetic.make_config(cfg, n_sets=S)
    capi.B200SchurLinearSystemSolver(p).close()
    print("cfg",cfg, file=sys.stderr, flush=True)
    capi.B200SchurLinearSystemSolver(p).close()
