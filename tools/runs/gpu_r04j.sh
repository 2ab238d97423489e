#!/bin/bash
# r04j: FieldExtractionIntegrator("bsdf"), reverse-mode gradients of the bitmaps' uv transforms; full GPU suite
mkdir -p gpurun_out/r04j
timeout 900 python -m pytest tests/test_gpu_collocated.py tests/test_gpu_adjoint.py -m gpu -q -k "bsdf_field or slots" 2>&1 | tail -25 | tee gpurun_out/r04j/pytest_a.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r04j/pytest_all.log
