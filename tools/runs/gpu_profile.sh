#!/bin/bash
# ncu --set full capture of one timed renderD step (forward + adjoint kernels) of bench.py; the .ncu-rep stays on
# the box (it exceeds gpurun's 64 MiB return limit), only CSV pages come back.  Usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
REP=/tmp/prof_$TAG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'interior|primary_edge|secondary_edge' -s 18 -c 7 -f -o $REP \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ncu -i $REP.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
for k in interior_kernel primary_edge_kernel secondary_edge_kernel interior_vjp_kernel; do
  ncu -i $REP.ncu-rep --page source --csv --print-source sass -k regex:$k > $OUT/sass_$k.csv 2>/dev/null
done
ls -la $OUT
