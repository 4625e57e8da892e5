"""The C++ mirror of kalibr2's drivers (include/kalibr_b200/calibration_tools.hpp) builds against the C ABI, and without a GPU it
fails loudly instead of computing anything on the CPU."""
import numpy as np
import pytest

from kalibr_b200 import synthetic

from driver_util import build_driver, run_driver, write_problem


def test_driver_mirror_builds_and_fails_loudly_without_a_device(tmp_path):
    import torch

    build_driver()
    p = synthetic.make_config(1, n_sets=3)
    path = str(tmp_path / "p.bin")
    write_problem(path, p, [[640, 480]])
    code, out = run_driver("single", path)
    if torch.cuda.is_available():
        assert code == 0 and "camera0" in out
    else:
        assert code == 1 and "error" in out and "kb_create" in out["error"]
