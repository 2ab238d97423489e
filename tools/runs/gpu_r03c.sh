#!/bin/bash
mkdir -p gpurun_out/r03c
timeout 600 python tools/ref_golden9.py > gpurun_out/r03c/golden9.log 2>&1; tail -14 gpurun_out/r03c/golden9.log
timeout 900 python -m pytest tests/test_gpu_fields.py tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r03c/pytest.log
