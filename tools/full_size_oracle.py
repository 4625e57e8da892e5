"""CPU side of the full-size parity check of BASELINE configs[3] / configs[4] (cfg 4: 8 cameras, 20 000 sets, 19.2 M terms; cfg 5:
16 cameras, 6 250 sets, 12 M terms): runs the oracle ONCE (minutes of CPU, done in the build container, where it costs no GPU time)
and stores compact, exactly reproducible summaries under tests/golden/full_size_cfgN.npz:
    cost J; rhs and dx(lambda = 10) in full; every 997th term of e(); sha256 of the block pattern arrays; a weighted checksum of the H
    values (fixed pseudo-random weights) and their absolute sum.
tests/test_full_size_gpu.py::test_full_size_against_the_stored_oracle_outputs recomputes the same quantities on the device.
Usage: python tools/full_size_oracle.py 4 5 [--threads 8]"""
import hashlib
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from kalibr_b200 import synthetic  # noqa: E402
from oracle import oracle_api  # noqa: E402

E_STRIDE = 997


def weights(n):
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(2654435761)) % np.uint64(1 << 32)).astype(np.float64) / float(1 << 32) + 0.5


def summarise(J, e, rhs, dx, ok, blocks):
    col_ptr, block_row, value_ptr, values = blocks
    h = hashlib.sha256()
    for a in (col_ptr.astype(np.int64), block_row.astype(np.int32), value_ptr.astype(np.int64)):
        h.update(np.ascontiguousarray(a).tobytes())
    return dict(J=np.float64(J), rhs=rhs, dx=dx, pos_def=np.int32(ok), e_sample=e.reshape(-1, 2)[::E_STRIDE].copy(), n_terms=np.int64(e.size // 2),
                pattern_sha256=np.frombuffer(h.digest(), np.uint8).copy(), n_blocks=np.int64(block_row.size), n_values=np.int64(values.size),
                h_checksum=np.float64(np.dot(values, weights(values.size))), h_abs_sum=np.float64(np.abs(values).sum()),
                e_checksum=np.float64(np.dot(e, weights(e.size))), e_abs_sum=np.float64(np.abs(e).sum()))


def main():
    threads = 8
    args = [a for a in sys.argv[1:]]
    if "--threads" in args:
        threads = int(args[args.index("--threads") + 1])
        del args[args.index("--threads"):args.index("--threads") + 2]
    for cfg in [int(a) for a in args] or [4, 5]:
        t0 = time.time()
        p = synthetic.make_config(cfg)
        o = oracle_api.OracleProblem(p, n_threads=threads)  # BlockCholesky: all cameras pinhole-radtan
        t1 = time.time()
        J = o.evaluate_error()
        e = o.error_vector()
        t2 = time.time()
        o.build_system()
        rhs = o.rhs()
        t3 = time.time()
        o.set_constant_conditioner(10.0)
        dx, ok = o.solve_system()
        t4 = time.time()
        blocks = o.hessian_blocks()  # after the solve: every diagonal block present (BlockCholeskyLinearSystemSolver.cpp:80)
        s = summarise(J, e, rhs, dx, ok, blocks)
        np.savez_compressed(f"tests/golden/full_size_cfg{cfg}.npz", **s)
        print(json.dumps({"cfg": cfg, "terms": int(p.n_terms), "sets": int(p.n_sets), "threads": threads, "create_s": round(t1 - t0, 1),
                          "evaluate_s": round(t2 - t1, 1), "build_s": round(t3 - t2, 1), "solve_s": round(t4 - t3, 1), "J": float(J),
                          "n_blocks": int(s["n_blocks"]), "pos_def": bool(ok)}), flush=True)
        o.close()


if __name__ == "__main__":
    main()
