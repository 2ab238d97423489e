#!/bin/bash
# r05a: primary-edge lane ordering (edge_sort.cu): GPU tests, then cfg 2 with 0 / 128 / 512 / 2048 buckets
mkdir -p gpurun_out/r05a
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r05a/pytest_gpu.log
for b in 0 128 512 2048; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --edge-sort $b 2>gpurun_out/r05a/bins$b.err | tee gpurun_out/r05a/bins$b.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('bins $b', d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['vjp']['ms_per_step'], d['vjp']['kernel_ms'], d['gpu_launches'])"
done
