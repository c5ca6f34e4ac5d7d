"""Build libmpcb.so (the C-ABI CUDA library) in-tree for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libmpcb.so")
SOURCES = ["mpcb.cu"]
HEADERS = ["mpcb_device.cuh", "mpcb_solver.cuh", "mpcb_sim.cuh", os.path.join("..", "..", "include", "mpcb.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmpcb.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", SO] + SOURCES
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as fh:
        fh.write(r.stderr)
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
