#!/bin/bash
mkdir -p gpurun_out/r02w
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02w/pytest_configs.log
bash tools/gpu_round2.sh r02w ncu ncuvjp
bash tools/gpu_profile_families.sh r02w
