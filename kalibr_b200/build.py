"""Builds the in-tree CUDA extension kalibr_b200/libkalibr_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libkalibr_b200.so")
SOURCES = ["kb_kernels.cu", "kb_init.cu", "kb_host.cpp"]
HEADERS = ["kb_device.cuh", "kb_models.cuh", "../../include/kalibr_b200.h", "../../include/kalibr_b200/optimizer.hpp", "../../include/kalibr_b200/lm_state_machine.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in SOURCES + HEADERS:
        p = os.path.join(_CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the package into one shared library.  Raises on failure."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES + ["-ldl"]
    subprocess.run(cmd, cwd=_CSRC, check=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build_extension(force="--force" in sys.argv, verbose="-v" in sys.argv))
