#!/usr/bin/env python3
"""Builds a VARIANT of libpsdr_b200.so here (no GPU needed): the named translation units recompiled with extra -D flags,
everything else taken from the objects of the regular build.  The variant travels to the GPU box with the snapshot and is
selected with PSDR_B200_LIB=<path>, so an A/B costs only its bench runs there (tools/gpu_*_sweep.sh compile on the box).
    python tools/build_variant.py <name> "<flags>" kern_cfg2.cu [more.cu ...]   ->  psdr_jit_b200/libpsdr_b200_<name>.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psdr_jit_b200 import build  # noqa: E402

name, flags, units = sys.argv[1], sys.argv[2].split(), sys.argv[3:]
build.build_native()
objdir, vardir = os.path.join(build.HERE, "build"), os.path.join(build.HERE, "build_var_" + name)
os.makedirs(vardir, exist_ok=True)
common = [build.nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-ccbin", "/usr/bin/g++",
          "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-x", "cu"]
objs = []
for src in build.SOURCES:
    base = os.path.splitext(src)[0] + ".o"
    if src in units:
        out = os.path.join(vardir, base)
        subprocess.check_call(common + flags + ["-c", "-o", out, os.path.join(build.CSRC, src)])
        objs.append(out)
    else:
        objs.append(os.path.join(objdir, base))
lib = os.path.join(build.HERE, "libpsdr_b200_%s.so" % name)
subprocess.check_call([build.nvcc_path(), "-shared", "-o", lib, "-ccbin", "/usr/bin/g++"] + objs)
print(lib)
