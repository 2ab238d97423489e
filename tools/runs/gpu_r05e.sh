#!/bin/bash
# r05e: ncu launch list of the default bench command in its final form (edge-ordering kernels included)
O=gpurun_out/r05e; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
grep -c . $O/launches.csv; tail -n 2 $O/ncu_launches.log | cut -c1-200
