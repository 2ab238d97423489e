#!/bin/bash
mkdir -p gpurun_out/r02r
python tools/exp_vjp_scaling.py 2>&1 | grep -v Warn | tee gpurun_out/r02r/scaling.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02r/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02r/bench.err | tee gpurun_out/r02r/bench_ours.json | cut -c1-300
