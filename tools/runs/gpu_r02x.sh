#!/bin/bash
# 1 GPU: full tests + bench with overlapped term streams
mkdir -p gpurun_out/r02x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02x/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02x/bench.err | tee gpurun_out/r02x/bench_ours.json | cut -c1-300
tail -3 gpurun_out/r02x/bench.err
