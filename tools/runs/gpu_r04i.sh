#!/bin/bash
mkdir -p gpurun_out/r04i
timeout 900 python -m pytest tests/test_gpu_collocated.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r04i/pytest_colloc.log
