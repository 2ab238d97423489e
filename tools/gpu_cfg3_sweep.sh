#!/bin/bash
# Block-size sweep of the full-feature family's kernels (kern_cfg2.cu) on BASELINE config 3.   bash tools/gpu_cfg3_sweep.sh "-DPSDR_BLOCK_I=512" ...
cd "$(dirname "$0")/.."
C=psdr_jit_b200/csrc; B=psdr_jit_b200/build
for V in "$@" ""; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-O2 -x cu $V -c -o $B/kern_cfg2.o $C/kern_cfg2.cu 2>/dev/null
  nvcc -shared -o psdr_jit_b200/libpsdr_b200.so -ccbin /usr/bin/g++ $B/*.o
  echo "variant [$V]"
  python bench.py --config 3 --no-cpu-baseline --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'])"
done
