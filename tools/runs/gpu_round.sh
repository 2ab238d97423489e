#!/bin/bash
# One GPU-box visit: GPU parity tests, smoke, both bench arms, ncu launch list + full capture (CSV pages only:
# the .ncu-rep exceeds gpurun's return limit).   bash tools/gpu_round.sh <tag>     (SKIP_REF=1 / SKIP_NCU=1)
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee -a $OUT/pytest_gpu.log
echo "== smoke" | tee $OUT/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a $OUT/smoke.log
echo "== bench ours"
timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/bench_ours.err | tee $OUT/bench_ours.json
tail -3 $OUT/bench_ours.err
if [ "$SKIP_REF" != "1" ]; then
  echo "== bench reference"
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json
  tail -3 $OUT/bench_ref.err
fi
if [ "$SKIP_NCU" != "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
  grep -c . $OUT/launches.csv
  echo "== ncu full capture (one timed step: 3 forward + 4 reverse-step kernels)"
  REP=/tmp/prof_$TAG
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'interior|primary_edge|secondary_edge' -s 18 -c 7 -f -o $REP \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
  tail -2 $OUT/ncu_full.log
  ncu -i $REP.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
  for k in interior_kernel primary_edge_kernel secondary_edge_kernel interior_vjp_kernel; do
    ncu -i $REP.ncu-rep --page source --csv --print-source sass -k regex:"$k" > $OUT/sass_$k.csv 2>/dev/null
  done
  ls -la $OUT
fi
