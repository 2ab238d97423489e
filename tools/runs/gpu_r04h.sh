#!/bin/bash
# r04h: CollocatedIntegrator -- CUDA vs oracle, golden of the running reference, reverse mode; full GPU suite; default bench
mkdir -p gpurun_out/r04h
timeout 600 python -m pytest tests/test_gpu_collocated.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r04h/pytest_a.log
timeout 600 python tools/ref_golden12.py > gpurun_out/r04h/golden12.log 2>&1; tail -6 gpurun_out/r04h/golden12.log
if [ -f gpurun_out/ref_golden12/collocated.npz ]; then cp gpurun_out/ref_golden12/collocated.npz tests/golden/collocated.npz; fi
timeout 600 python -m pytest tests/test_gpu_collocated.py -m gpu -q -k golden 2>&1 | tail -12 | tee gpurun_out/r04h/pytest_b.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r04h/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r04h/bench.err | tee gpurun_out/r04h/bench_ours.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['vjp']['ms_per_step'], d['vjp']['kernel_ms'])"
