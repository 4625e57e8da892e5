"""Helpers shared by the C++ driver tests: build tests/cpp/driver_main.cpp against the in-tree library, write a problem file."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "driver_main")


def build_driver():
    from kalibr_b200 import build as kb_build

    if not os.path.exists(kb_build.LIB_PATH):  # never rebuild behind a running test session: only when there is nothing to link
        kb_build.build_extension()
    src = os.path.join(ROOT, "tests", "cpp", "driver_main.cpp")
    lib_dir = os.path.join(ROOT, "kalibr_b200")
    hdrs = [os.path.join(ROOT, "include", "kalibr_b200", h) for h in ("calibration_tools.hpp", "incremental_estimator.hpp")] + [os.path.join(ROOT, "include", "kalibr_b200.h")]
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(f) for f in [src, kb_build.LIB_PATH] + hdrs):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", lib_dir,
                        "-lkalibr_b200", f"-Wl,-rpath,{lib_dir}", "-o", BIN], check=True)
    return BIN


def build_adapter(virtual_evaluate: bool):
    """tests/cpp/reference_adapter_main.cpp against the stand-in of the reference's plugin headers (tests/cpp/aslam_mock)."""
    from kalibr_b200 import build as kb_build

    if not os.path.exists(kb_build.LIB_PATH):
        kb_build.build_extension()
    out = os.path.join(ROOT, "tests", "cpp", "reference_adapter_virtual" if virtual_evaluate else "reference_adapter_main")
    src = os.path.join(ROOT, "tests", "cpp", "reference_adapter_main.cpp")
    lib_dir = os.path.join(ROOT, "kalibr_b200")
    deps = [src, kb_build.LIB_PATH, os.path.join(ROOT, "tests", "cpp", "aslam_mock", "aslam_backend_mock.hpp"), os.path.join(ROOT, "include", "kalibr_b200.h")]
    deps += [os.path.join(ROOT, "include", "kalibr_b200", h) for h in ("reference_adapter.hpp", "lm_state_machine.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in deps):
        flags = ["-DKB_REFERENCE_HAS_VIRTUAL_EVALUATE_ERROR"] if virtual_evaluate else []
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", *flags, "-I", os.path.join(ROOT, "include"), "-I",
                        os.path.join(ROOT, "tests", "cpp", "aslam_mock"), src, "-L", lib_dir, "-lkalibr_b200", f"-Wl,-rpath,{lib_dir}", "-o", out], check=True)
    return out


def run_adapter(path, order, virtual_evaluate: bool):
    r = subprocess.run([build_adapter(virtual_evaluate), path, str(order)], capture_output=True, text=True)
    out = {}
    for line in r.stdout.splitlines():
        k, *v = line.split(" ")
        out[k] = " ".join(v) if k == "error" else np.array([float(x) for x in v])
    return r.returncode, out


def write_problem(path, p, resolution, rows=10, cols=12):
    def arr(f, a, dt):
        a = np.ascontiguousarray(a, dt).ravel()
        f.write(np.int64(a.size).tobytes())
        f.write(a.tobytes())

    with open(path, "wb") as f:
        arr(f, [p.n_cams, p.n_sets, rows, cols], np.int32)
        arr(f, p.cam_model, np.int32)
        arr(f, resolution, np.int32)
        arr(f, p.cam_params, np.float64)
        arr(f, p.baselines, np.float64)
        arr(f, p.target_points, np.float64)
        arr(f, p.view_set, np.int32)
        arr(f, p.view_cam, np.int32)
        arr(f, p.view_begin, np.int64)
        arr(f, p.corner_id, np.int32)
        arr(f, p.y_u, np.float64)
        arr(f, p.y_v, np.float64)
        arr(f, p.set_poses, np.float64)


def run_driver(mode, path, *extra):
    r = subprocess.run([build_driver(), mode, path, *[str(e) for e in extra]], capture_output=True, text=True)
    out = {}
    for line in r.stdout.splitlines():
        k, *v = line.split(" ")
        out[k] = " ".join(v) if k == "error" else np.array([float(x) for x in v])
    return r.returncode, out
