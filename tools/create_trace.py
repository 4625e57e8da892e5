"""KB_CREATE_TRACE=1 python tools/create_trace.py [cfg sets]: wall-clock checkpoints of kb_create / kb_destroy (stderr), second create of
each problem (the first one of a process pays the CUDA context and module load)."""
import sys
sys.path.insert(0, '.')
from kalibr_b200 import capi, synthetic

cases = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(4, 12), (2, 40)]
for cfg, S in cases:
    p = synthetic.make_config(cfg, n_sets=S)
    capi.B200SchurLinearSystemSolver(p).close()
    print("cfg", cfg, "sets", S, file=sys.stderr, flush=True)
    capi.B200SchurLinearSystemSolver(p).close()
