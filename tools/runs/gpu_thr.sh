#!/bin/bash
# Lane utilisation (threads per warp instruction) and duration of the kernels matching $1 during one bench step.
K=${1:-interior_vjp}
timeout 600 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"$K" -s 2 -c 2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "^\s+(void|smsp__|gpu__)" | cut -c1-150
