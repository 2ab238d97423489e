#!/bin/bash
mkdir -p gpurun_out/r02z
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02z/pytest.log
python tools/exp_vjp_scaling.py 2>&1 | grep -v Warn | tee gpurun_out/r02z/scaling.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02z/bench.err | tee gpurun_out/r02z/bench_ours.json | cut -c1-300
for c in 4 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02z/cfg$c.err | tee gpurun_out/r02z/cfg$c.json | cut -c1-330; done
