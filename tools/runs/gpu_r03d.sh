#!/bin/bash
mkdir -p gpurun_out/r03d
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r03d/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r03d/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r03d/bench.err | tee gpurun_out/r03d/bench_ours.json | cut -c1-300
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r03d/cfg3.err | tee gpurun_out/r03d/cfg3.json | cut -c1-330
timeout 600 python bench.py --config 1 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r03d/cfg1.err | tee gpurun_out/r03d/cfg1.json | cut -c1-330
