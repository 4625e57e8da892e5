"""Host logic of bench.py that needs no GPU: the reference arm (`--impl reference`, the CPU leg of the contract) prints ONE line with the keys
the driver reads, its `config` is the b200 arm's, the CPU regimes are all there, and the reference-code regimes - timed in a process of
their own - degrade to "unavailable" instead of taking the bench with them."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def reference_line(oracle_lib):
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--config", "2", "--cpu-sets", "6", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1  # ONE JSON line
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line(reference_line):
    d = reference_line
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["gpu_launches"] == 0
    assert d["unit"] == "terms/s" and d["value"] > 0 and d["ms_per_step"] > 0
    assert set(d["config"]) >= {"workload", "cameras", "synced_sets", "parallelism", "l2"} and "configs[1]" in d["config"]["workload"]
    assert "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["value"] == d["value"] and c["kind"] in ("port", "reference") and c["cores"] >= 1 and "6 of its synced sets" in c["sample"]


def test_reference_arm_config_is_the_b200_arms(reference_line):
    """`config` is a function of the command line alone (bench.workload_config): both arms print the identical object"""
    import argparse

    import bench

    args = argparse.Namespace(config=2, sets=None, no_peer_exchange=False, gpus=1)
    assert reference_line["config"] == bench.workload_config(args, 1, "weak")


def test_all_cpu_regimes_are_reported(reference_line):
    from oracle import oracle_api as oa

    regimes = reference_line["cpu_baseline"]["regimes"]
    port = [(r["solver"], r["threads"]) for r in regimes if r.get("kind", "port") == "port"]
    cores = os.cpu_count() or 1
    assert port == [("sparse", cores), ("sparse", min(4, cores)), ("block", cores), ("block", min(4, cores))]
    ref = [r for r in regimes if r.get("kind") == "reference"]
    if oa.build_reference_cameras() is not None:  # the reference's own compiled evaluate + build of both solvers (oracle/_ref)
        assert [r["solver"] for r in ref] == ["block", "sparse"] and all(r["value"] > 0 and "unavailable" not in r for r in ref)
        for r in ref:
            assert r["stage_s_per_iteration"]["evaluate"] > 0 and r["stage_s_per_iteration"]["build"] > 0
    assert reference_line["value"] >= 0.5 * max(r["value"] for r in regimes)  # the headline is the fastest regime, re-timed


def test_a_failing_reference_timing_child_is_an_exception_not_a_crash(oracle_lib):
    import bench
    from oracle import oracle_api as oa

    if oa.build_reference_cameras() is None:
        pytest.skip("no oracle/_ref here")
    t = bench.reference_evaluate_build_isolated(1, 3, 2, 1, oa.SPARSE_CHOLESKY_KIND)
    assert t["evaluate_s"] > 0 and t["build_s"] > 0 and t["cost"] > 0
    with pytest.raises(RuntimeError):
        bench.reference_evaluate_build_isolated(99, 3, 2, 1, oa.SPARSE_CHOLESKY_KIND)  # no such configuration: the child exits non-zero
