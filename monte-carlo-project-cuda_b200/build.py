"""Build recipe for libmcb200.so (sm_100a only, in-tree so the .so travels with gpurun)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libmcb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    # static cudart (nvcc's default): the library is self-contained and does not care which
    # libcudart.so.12 the host process (e.g. torch) has already mapped.
    "-cudart", "static",
]


def sources():
    out = [os.path.join(ROOT, "include", "mcb200.h")]
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, name))
    return out


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/mcb200.cu -> libmcb200.so with nvcc for sm_100a.  Returns the library path."""
    if not force and not stale():
        return LIB
    nvcc = nvcc_path()
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libmcb200.so cannot be built (and there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB + ".tmp", os.path.join(CSRC, "mcb200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
