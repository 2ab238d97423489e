#!/bin/bash
# Variant sweep of the cfg-0 ADJOINT kernels on the GPU box: recompiles vjp_cfg0.cu per variant, relinks, runs bench.py (with the reverse step).
# bash tools/gpu_vjp_sweep.sh "-DPSDR_VJP_GEO_NOINLINE=1" "-DPSDR_TRACE_NOINLINE=1" ...
cd "$(dirname "$0")/.."
C=psdr_jit_b200/csrc; B=psdr_jit_b200/build
for V in "$@" ""; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-O2 -x cu $V -c -o $B/vjp_cfg0.o $C/vjp_cfg0.cu 2>/dev/null
  nvcc -shared -o psdr_jit_b200/libpsdr_b200.so -ccbin /usr/bin/g++ $B/*.o
  echo "variant [$V]"
  python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['vjp']['ms_per_step'], d['vjp']['kernel_ms'])"
done
