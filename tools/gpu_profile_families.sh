#!/bin/bash
# ncu --set full of the other kernel families: cfg 3 (kCfg = 2: Microfacet + environment map, brute-force scan) and
# cfg 4 (kCfg = 1: BVH traversal, guided secondary edges).   bash tools/gpu_profile_families.sh <tag>
TAG=${1:-r02w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in "3 4 1" "4 8 2"; do
  set -- $spec; CFG=$1; SKIP=$2; CNT=$3
  REP=/tmp/prof_${TAG}_cfg$CFG
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'interior_kernel|primary_edge_kernel|secondary_edge_kernel' -s $SKIP -c $CNT -f -o $REP \
      python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_cfg$CFG.log 2>&1
  tail -2 $OUT/ncu_cfg$CFG.log
  ncu -i $REP.ncu-rep --page raw --csv > $OUT/raw_cfg$CFG.csv 2>/dev/null
  ncu -i $REP.ncu-rep --page source --csv --print-source sass > $OUT/sass_cfg$CFG.csv 2>/dev/null
done
ls -la $OUT
