#!/bin/bash
# r04b: reverse mode of the dielectric / per-vertex BSDFs, full GPU suite + smoke, ncu launch list + full capture of the
# cfg 2 kernels in their final form, ncu of the (smaller) full-feature family on cfg 3
mkdir -p gpurun_out/r04b
timeout 900 python -m pytest tests/test_gpu_ext_bsdfs.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r04b/pytest_ext.log
bash tools/gpu_round2.sh r04b tests ncu 2>&1 | tail -40
sed -i 's/for spec in "3 4 1" "4 8 2"; do/for spec in ${SPECS:-"3 4 1" "4 8 2"}; do/' tools/gpu_profile_families.sh
timeout 900 bash -c 'REP=/tmp/prof_r04b_cfg3; ncu --set full --clock-control none --import-source on -k regex:"interior_kernel" -s 4 -c 1 -f -o $REP python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r04b/ncu_cfg3.log 2>&1; tail -2 gpurun_out/r04b/ncu_cfg3.log; ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/r04b/raw_cfg3.csv 2>/dev/null; ncu -i $REP.ncu-rep --page source --csv --print-source sass > gpurun_out/r04b/sass_cfg3.csv 2>/dev/null'
ls -la gpurun_out/r04b
