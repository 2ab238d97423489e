#!/bin/bash
# r05b: adjoint with warp-merged edge adds under the sorted lane order; GPU tests, bench, CTA-shape variants of the sorted
# primary-edge kernel (prebuilt with tools/build_variant.py), ncu full capture of the sorted kernel
O=gpurun_out/r05b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $O/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>$O/bench.err | tee $O/bench_ours.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('default', d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['vjp']['ms_per_step'], d['vjp']['kernel_ms'])"
for v in sync2 sync0 b1024 b640; do
  L=psdr_jit_b200/libpsdr_b200_$v.so
  [ -f $L ] || continue
  PSDR_B200_LIB=$PWD/$L timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-vjp 2>$O/$v.err | tee $O/$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$v', d['ms_per_step'], d['kernel_ms'])"
done
REP=/tmp/prof_r05b
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'primary_edge_kernel' -s 4 -c 1 -f -o $REP python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-vjp > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ncu -i $REP.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --print-source sass > $O/sass_primary_edge_kernel.csv 2>/dev/null
ls -la $O | tail -12
