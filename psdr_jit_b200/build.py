"""Builds libpsdr_b200.so (C ABI + host scene code + sm_100a kernels) in-tree with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpsdr_b200.so")
# Variant "refarith": the same sources compiled with Dr.Jit-like approximate fp32 division / reciprocal / square root
# (Dr.Jit emits div.approx.ftz, rcp.approx.ftz, sqrt.approx.ftz for CUDA floats: drjit-core cuda_eval.cpp:441-475,
# 638-647).  Not bit-comparable with the CPU oracle; used by tools/ref_parity.py to measure how much of the residual
# against the running reference is rounding mode.  Selected with PSDR_REFERENCE_ARITHMETIC=1 (psdr_jit_b200/_lib.py).
LIB_REFARITH = os.path.join(HERE, "libpsdr_b200_refarith.so")
SOURCES = ["kern_cfg11a.cu", "kern_cfg10a.cu", "vjp_cfg11a.cu", "vjp_cfg10a.cu", "vjp_cfg11.cu", "vjp_cfg10.cu", "kern_cfg11.cu", "kern_cfg10.cu", "vjp_cfg3.cu", "vjp_cfg2.cu", "vjp_cfg1.cu", "vjp_cfg0.cu", "kern_cfg3.cu", "kern_cfg2.cu", "kern_cfg1.cu", "kern_cfg0.cu",
           "capi.cpp", "scene.cpp", "scene_grad.cpp", "device_upload.cu", "kernels.cu", "edge_sort.cu"]
HEADERS = ["pmath.h", "dscene.h", "scene.h", "kernels.h", "device_path.cuh", "adjoint.cuh", "grad_layout.h", "texture.h", "kernels_impl.cuh", "kernels_vjp_impl.cuh", "launch_decl.h", os.path.join("..", "..", "include", "psdr_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


STAMP = LIB + ".stamp"


def source_digest() -> str:
    """Content hash of every source and header (file times do not survive a copy to another machine)."""
    import hashlib
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def is_stale(lib: str = None) -> bool:
    lib = lib or LIB
    if not os.path.exists(lib) or not os.path.exists(lib + ".stamp"):
        return True
    with open(lib + ".stamp") as fh:
        return fh.read().strip() != source_digest()


def build_native(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    """Every source is compiled to an object in parallel (nvcc -c), then linked; ptxas dominates the time.
    Objects are keyed by a content hash of their source, the headers and the flags (file times do not survive a copy)."""
    lib = LIB_REFARITH if variant == "refarith" else LIB
    if not force and not is_stale(lib):
        return lib
    from concurrent.futures import ThreadPoolExecutor
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    objdir = os.path.join(HERE, "build" + ("_" + variant if variant else ""))
    os.makedirs(objdir, exist_ok=True)
    common = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # no implicit fused multiply-adds on either side: the kernels spell out fmaf() where they
              # want one, so results do not depend on contraction choices (see pmath.h)
              "-fmad=false", "-ccbin", host_cxx, "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-x", "cu"]
    if variant == "refarith":
        common[1:1] = ["-prec-div=false", "-prec-sqrt=false", "-ftz=true"]
    if verbose:
        common.insert(1, "-Xptxas=-v")
    import hashlib
    hh = hashlib.sha256(" ".join(common).encode())
    for f in HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            hh.update(fh.read())
    hdr_digest = hh.hexdigest()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        spath = os.path.join(CSRC, src)
        with open(spath, "rb") as fh:
            key = hashlib.sha256(hdr_digest.encode() + fh.read()).hexdigest()
        keyfile = obj + ".key"
        if not force and os.path.exists(obj) and os.path.exists(keyfile) and open(keyfile).read().strip() == key:
            return obj
        cmd = common + ["-c", "-o", obj, spath]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        with open(keyfile, "w") as fh:
            fh.write(key + "\n")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc_path(), "-shared", "-o", lib, "-ccbin", host_cxx] + objs)
    with open(lib + ".stamp", "w") as fh:
        fh.write(source_digest() + "\n")
    return lib


if __name__ == "__main__":
    print(build_native(force=True, verbose="-v" in sys.argv))
