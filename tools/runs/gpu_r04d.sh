#!/bin/bash
# r04d: PerspectiveCamera(fx, fy, cx, cy, near, far) -- CUDA vs oracle, golden of the running reference, then the golden test
mkdir -p gpurun_out/r04d
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k intrinsics 2>&1 | tail -8 | tee gpurun_out/r04d/pytest_a.log
timeout 600 python tools/ref_golden11.py > gpurun_out/r04d/golden11.log 2>&1; tail -8 gpurun_out/r04d/golden11.log
if [ -f gpurun_out/ref_golden11/intrinsics.npz ]; then cp gpurun_out/ref_golden11/intrinsics.npz tests/golden/intrinsics.npz; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k intrinsics 2>&1 | tail -8 | tee gpurun_out/r04d/pytest_b.log
timeout 300 python -m pytest tests/test_cpu_oracle.py -q -k intrinsics 2>&1 | tail -4 | tee gpurun_out/r04d/pytest_c.log
