"""The C-ABI library loads on a CPU-only box, exports every symbol include/kalibr_b200.h declares, validates
arguments, and refuses to compute without a B200 (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from kalibr_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from kalibr_b200 import build, capi as m

    build.build_extension()  # nvcc cross-compiles sm_100a without a GPU
    return m


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kalibr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"KB_API\s+[\w\s\*]+?\b(kb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in kalibr_b200.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == names


def test_header_cites_the_reference_interface_for_each_entry_point():
    text = open(os.path.join(ROOT, "include", "kalibr_b200.h")).read()
    for needle in ("LinearSystemSolver.cpp", "BlockCholeskyLinearSystemSolver.cpp", "Optimizer2.cpp", "LevenbergMarquardtTrustRegionPolicy.cpp",
                   "CompressedColumnJacobianTransposeBuilder.hpp", "CalibrationTools.hpp"):
        assert needle in text


def test_struct_layout_matches_the_header(capi):
    from kalibr_b200.problem import KbOptimizerOptions, KbProblemDesc, KbSolution

    # field order of the ctypes mirror vs the header
    text = open(os.path.join(ROOT, "include", "kalibr_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} kb_problem_desc;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(\w+)\s*;", body)
    assert fields == [f for f, _ in KbProblemDesc._fields_]
    assert C.sizeof(KbOptimizerOptions) == 40 and C.sizeof(KbSolution) == 48


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_cpu_fallback(capi):
    p = synthetic.make_config(1, n_sets=2)
    with pytest.raises(capi.KalibrB200Error, match="no CUDA device|no CPU fallback"):
        capi.B200SchurLinearSystemSolver(p)


def test_argument_validation_happens_before_any_device_work(capi):
    p = synthetic.make_config(2, n_sets=2)
    lib = capi.load_library()
    d = p.desc()
    d.driver_order = 7
    h = C.c_void_p()
    assert lib.kb_create(C.byref(d), C.byref(h)) == -1
    assert b"driver order" in lib.kb_last_error(None)
    d = p.desc()
    d.driver_order = 0  # single-camera order with two cameras
    assert lib.kb_create(C.byref(d), C.byref(h)) == -1
    d = p.desc(n_ranks=2, rank=2)
    assert lib.kb_create(C.byref(d), C.byref(h)) == -1


def test_missing_library_fails_loudly(capi, monkeypatch, tmp_path):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(capi.KalibrB200Error, match="no CPU fallback"):
        capi.load_library()


@pytest.mark.parametrize("virtual_evaluate", [False, True])
def test_reference_adapter_compiles_against_the_reference_interface(capi, virtual_evaluate, tmp_path):
    """include/kalibr_b200/reference_adapter.hpp (the LinearSystemSolver subclass a Kalibr2 maintainer adds) builds warning-free against
    the stand-in of the reference's plugin headers, in both variants of INTEGRATION.md §2, and refuses to run without a device."""
    import subprocess
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from driver_util import build_adapter, write_problem

    exe = build_adapter(virtual_evaluate)
    assert os.path.exists(exe)
    if os.path.exists("/dev/nvidia0"):
        return
    p = synthetic.make_config(1, n_sets=2)
    path = str(tmp_path / "p.bin")
    write_problem(path, p, [synthetic.TRUTH_PARAMS[0][1]])
    r = subprocess.run([exe, path, "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stdout  # kb_create inside initMatrixStructureImplementation: no CPU fallback


def test_state_machine_is_the_single_policy_implementation():
    """One LM / Gauss-Newton state machine (include/kalibr_b200/lm_state_machine.h) is used by the device control kernels and the host
    mirror alike: neither kb_kernels.cu nor optimizer.hpp carries its own copy of the lambda schedule."""
    kern = open(os.path.join(ROOT, "kalibr_b200", "csrc", "kb_kernels.cu")).read()
    host = open(os.path.join(ROOT, "include", "kalibr_b200", "optimizer.hpp")).read()
    for text in (kern, host):
        assert "lm_before_solve" in text and "lm_after_solve" in text and "lm_after_eval" in text
        assert "mu *= 10" not in text.replace("_mu", "mu") and "pow(" not in text
