#!/bin/bash
# r04g: NormalMap in reverse mode; full GPU suite; default bench (checks that the adjoint refactor left cfg 2 where it was)
mkdir -p gpurun_out/r04g
timeout 900 python -m pytest tests/test_gpu_ext_bsdfs.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r04g/pytest_ext.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r04g/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r04g/bench.err | tee gpurun_out/r04g/bench_ours.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['vjp']['ms_per_step'], d['vjp']['kernel_ms'])"
