"""Generates tests/golden/*.npz from the CPU oracle on small seeded problems.

    python tests/golden/make_golden.py

The reference stores no golden vectors for this path and cannot be built here, so these fixtures
are outputs of the oracle restatement ("parity unpinned"): they pin the oracle and the CUDA path against
regressions and let the GPU box check parity without /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from kalibr_b200 import synthetic  # noqa: E402
from kalibr_b200.problem import KbOptimizerOptions  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402

CASES = {"cfg1_S3": (1, 3), "cfg2_S2": (2, 2), "cfg3_S2": (3, 2), "cfg4_S1": (4, 1), "cfg6_S2": (6, 2), "cfg7_S2": (7, 2)}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for name, (cfg, S) in CASES.items():
        if only and name not in only:
            continue
        p = synthetic.make_config(cfg, n_sets=S)
        o = oa.OracleProblem(p, oa.BLOCK_CHOLESKY, n_threads=1)
        J0 = o.evaluate_error()
        e = o.error_vector()
        cp, ri, jv = o.jacobian_ccs()
        o.build_system()
        rhs = o.rhs()
        o.set_constant_conditioner(10.0)
        dx, ok = o.solve_system()
        hcp, hbr, hvp, hval = o.hessian_blocks()
        o2 = oa.OracleProblem(p, oa.BLOCK_CHOLESKY, n_threads=1)
        sol, tr = o2.optimize(KbOptimizerOptions.kalibr2_default())
        np.savez_compressed(
            os.path.join(here, name + ".npz"),
            cfg=cfg, n_sets=S, J0=J0, e=e, jt_col_ptr=cp, jt_row_idx=ri, jt_values=jv, rhs=rhs, dx=dx, pos_def=ok,
            h_col_ptr=hcp, h_block_row=hbr, h_value_ptr=hvp, h_values=hval,
            iterations=sol.iterations, failed_iterations=sol.failed_iterations, j_final=sol.j_final, trace=tr,
            cam_params=o2.camera_params(), baselines=o2.baselines(), set_poses=o2.set_poses(),
            # the inputs, so that a change of the generator is detected instead of silently re-basing the fixture
            y_u=p.y_u, y_v=p.y_v, init_cam=p.cam_params, init_sets=p.set_poses,
        )
        print(name, "terms", p.n_terms, "J0", J0, "iters", sol.iterations)


if __name__ == "__main__":
    main()
