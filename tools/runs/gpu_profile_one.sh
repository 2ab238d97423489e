#!/bin/bash
# ncu --set full of ONE kernel family (regex $1) from bench.py; CSV pages only come back.  bash tools/gpu_profile_one.sh <regex> <tag> [skip]
K=$1; TAG=${2:-one}; SKIP=${3:-3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
REP=/tmp/prof_$TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c 1 -f -o $REP \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ncu -i $REP.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --print-source sass > $OUT/sass.csv 2>/dev/null
ls -la $OUT
