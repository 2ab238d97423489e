#!/bin/bash
# r04k: OrthographicCamera -- CUDA vs oracle, reverse mode, golden of the running reference; full GPU suite; default bench
mkdir -p gpurun_out/r04k
timeout 600 python -m pytest tests/test_gpu_ortho.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r04k/pytest_a.log
timeout 600 python tools/ref_golden13.py > gpurun_out/r04k/golden13.log 2>&1; tail -6 gpurun_out/r04k/golden13.log
if [ -f gpurun_out/ref_golden13/ortho.npz ]; then cp gpurun_out/ref_golden13/ortho.npz tests/golden/ortho.npz; fi
timeout 600 python -m pytest tests/test_gpu_ortho.py -m gpu -q -k golden 2>&1 | tail -12 | tee gpurun_out/r04k/pytest_b.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r04k/pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r04k/bench.err | tee gpurun_out/r04k/bench_ours.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['vjp']['ms_per_step'], d['vjp']['kernel_ms'])"
