#!/bin/bash
# r04p: reverse mode of FieldExtractionIntegrator; full GPU suite
mkdir -p gpurun_out/r04p
timeout 900 python -m pytest tests/test_gpu_fields.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r04p/pytest_fields.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r04p/pytest_all.log
