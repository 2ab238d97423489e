#!/bin/bash
# ncu --set full of ONE kernel (regex $1) while tools/run_configs.py renders BASELINE config $2 at spp scale $3.
# bash tools/gpu_profile_cfg.sh <regex> <config> <scale> <tag> [skip]
K=$1; CFG=$2; SCALE=$3; TAG=${4:-cfg}; SKIP=${5:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
REP=/tmp/prof_$TAG
timeout 800 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c 1 -f -o $REP \
    python tools/run_configs.py --configs $CFG --reps 1 --scale $SCALE > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ncu -i $REP.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --print-source sass > $OUT/sass.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --print-source cuda > $OUT/cuda.csv 2>/dev/null
ls -la $OUT
