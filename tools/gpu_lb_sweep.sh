#!/bin/bash
# Launch-bounds sweep of the cfg-0 forward kernels ON the GPU box: recompiles kern_cfg0.cu per variant, relinks, runs bench.py.
# bash tools/gpu_lb_sweep.sh "PRIMARY=6" "PRIMARY=5 SECONDARY=6" "-DPSDR_SEC_FILL_ROUNDS=12" ...
cd "$(dirname "$0")/.."
C=psdr_jit_b200/csrc; B=psdr_jit_b200/build
for V in "$@" ""; do
  D=""; for kv in $V; do case $kv in -D*) D="$D $kv";; *) D="$D -DPSDR_LB_${kv}";; esac; done
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-O2 -x cu $D -c -o $B/kern_cfg0.o $C/kern_cfg0.cu 2>/dev/null
  nvcc -shared -o psdr_jit_b200/libpsdr_b200.so -ccbin /usr/bin/g++ $B/*.o
  echo "variant [$V]"
  python bench.py --no-cpu-baseline --no-vjp 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'])"
done
