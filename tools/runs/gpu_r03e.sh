#!/bin/bash
mkdir -p gpurun_out/r03e
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adjoint.py tests/test_gpu_fields.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r03e/pytest.log
for pol in 0 1; do echo "cta policy $pol"; timeout 600 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --cta-policy $pol 2>gpurun_out/r03e/cfg4_p$pol.err | tee gpurun_out/r03e/cfg4_p$pol.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms'], d['e2e']['ms_per_step'])"; done
