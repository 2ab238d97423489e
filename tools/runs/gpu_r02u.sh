#!/bin/bash
mkdir -p gpurun_out/r02u
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r02u/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r02u/bench.err | tee gpurun_out/r02u/bench_ours.json | cut -c1-300
