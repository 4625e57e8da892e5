"""TEST INFRASTRUCTURE (bench.py's cpu_baseline only): times the REFERENCE's own compiled evaluate + build (oracle/_ref, see
oracle/ref_pin_optimizer.cpp) on a sample of a bench configuration, in a process of its own.

    python -m oracle.reference_timing <config> <synced sets> <threads> <repeats> <solver kind: 0 block, 1 sparse>

prints one JSON line: {"setup_s", "evaluate_s", "build_s", "cost"}.  A separate process because that library is the reference's code
over STAND-IN Eigen / Boost headers running the reference's threaded passes: if it ever fails there, the bench - a CUDA process that
must print its line - records the regime as unavailable instead of dying with it.
"""
import json
import sys


def main(argv):
    from kalibr_b200 import synthetic
    from oracle import oracle_api as oa

    config, n_sets, threads, repeats, kind = (int(a) for a in argv[:5])
    if oa.build_reference_cameras() is None:
        print(json.dumps({"unavailable": "oracle/_ref/libkalibr_ref.so is not here"}))
        return 0
    ps = synthetic.make_config(config, n_sets=n_sets)
    print(json.dumps(oa.reference_time_evaluate_build(ps, threads, repeats, kind)))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
