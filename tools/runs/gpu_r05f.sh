#!/bin/bash
# r05f: dynamic chunk hand-out in the large-CTA primary-edge kernels (default) against static slices (PSDR_EDGE_DYNAMIC=0)
O=gpurun_out/r05f; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest_gpu.log
PSDR_EDGE_DYNAMIC=0 timeout 120 python -m pytest tests/test_gpu_edge_sort.py -m gpu -q -x 2>&1 | tail -2 | tee $O/pytest_static.log
for d in 1 0; do
  PSDR_EDGE_DYNAMIC=$d timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>$O/dyn$d.err | tee $O/dyn$d.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('dynamic $d', d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['vjp']['ms_per_step'], d['vjp']['kernel_ms'])"
done
