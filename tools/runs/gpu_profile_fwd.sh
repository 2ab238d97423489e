#!/bin/bash
# ncu --set full of the three forward (JVP) kernels of one timed bench step; CSV pages only.  bash tools/gpu_profile_fwd.sh <tag>
TAG=${1:-fwd}
OUT=gpurun_out/$TAG
mkdir -p $OUT
REP=/tmp/prof_$TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'interior_kernel|primary_edge_kernel|secondary_edge_kernel' -s 9 -c 3 -f -o $REP \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-vjp > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ncu -i $REP.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
for k in interior_kernel primary_edge_kernel secondary_edge_kernel; do
  ncu -i $REP.ncu-rep --page source --csv --print-source sass -k regex:$k > $OUT/sass_$k.csv 2>/dev/null
done
ls -la $OUT
