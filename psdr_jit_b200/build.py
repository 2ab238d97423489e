"""Builds libpsdr_b200.so (C ABI + host scene code + sm_100a kernels) in-tree with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpsdr_b200.so")
SOURCES = ["capi.cpp", "scene.cpp", "scene_grad.cpp", "device_upload.cu", "kernels.cu", "kernels_vjp.cu"]
HEADERS = ["pmath.h", "dscene.h", "scene.h", "kernels.h", "device_path.cuh", "adjoint.cuh", "grad_layout.h", os.path.join("..", "..", "include", "psdr_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           # no implicit fused multiply-adds on either side: the kernels spell out fmaf() where they
           # want one, so results do not depend on contraction choices (see pmath.h)
           "-fmad=false", "-ccbin", host_cxx, "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
           "-x", "cu", "-shared", "-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_native(force=True, verbose="-v" in sys.argv))
