#!/bin/bash
# Round-2 GPU-box visit: GPU tests, bench (both arms), A/B of the box cull, parity experiments vs the running reference,
# ncu launch list + full capture of the forward kernels.   bash tools/gpu_round2.sh <tag> [steps: tests bench ab ref parity ncu]
TAG=${1:-r02a}; shift
STEPS=${*:-tests bench ab ref parity ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
for S in $STEPS; do case $S in
tests)
  echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee -a $OUT/pytest_gpu.log
  echo "== smoke" | tee $OUT/smoke.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a $OUT/smoke.log ;;
bench)
  echo "== bench ours"
  timeout 600 python bench.py --steps 20 --warmup 3 2>$OUT/bench_ours.err | tee $OUT/bench_ours.json
  tail -3 $OUT/bench_ours.err ;;
ab)
  echo "== A/B: box cull off"
  bash tools/gpu_lb_sweep.sh "-DPSDR_BRUTE_CULL=0" 2>&1 | tee $OUT/ab_cull.log ;;
ref)
  echo "== bench reference"
  timeout 900 python bench.py --impl reference --steps 10 --warmup 2 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json
  tail -3 $OUT/bench_ref.err ;;
parity)
  echo "== parity vs the running reference"
  timeout 1500 python tools/ref_parity.py all > $OUT/parity.log 2>&1
  tail -30 $OUT/parity.log
  cp -r gpurun_out/parity $OUT/ 2>/dev/null ;;
ncu)
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
  grep -c . $OUT/launches.csv
  echo "== ncu full capture (one timed step: 3 forward kernels)"
  REP=/tmp/prof_$TAG
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'interior_kernel|primary_edge_kernel|secondary_edge_kernel' -s 12 -c 3 -f -o $REP \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-vjp > $OUT/ncu_full.log 2>&1
  tail -2 $OUT/ncu_full.log
  ncu -i $REP.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
  for k in interior_kernel primary_edge_kernel secondary_edge_kernel; do
    ncu -i $REP.ncu-rep --page source --csv --print-source sass -k regex:"$k" > $OUT/sass_$k.csv 2>/dev/null
  done
  ls -la $OUT ;;
ncuvjp)
  echo "== ncu full capture of the adjoint kernels"
  REP=/tmp/profv_$TAG
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'vjp_kernel' -s 9 -c 3 -f -o $REP \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_vjp.log 2>&1
  tail -2 $OUT/ncu_vjp.log
  ncu -i $REP.ncu-rep --page raw --csv > $OUT/raw_vjp.csv 2>/dev/null
  for k in interior_vjp_kernel primary_edge_vjp_kernel secondary_edge_vjp_kernel; do
    ncu -i $REP.ncu-rep --page source --csv --print-source sass -k regex:"$k" > $OUT/sass_$k.csv 2>/dev/null
  done
  ls -la $OUT ;;
esac; done
