#!/bin/bash
mkdir -p gpurun_out/r02t
timeout 600 python tools/ref_golden8.py > gpurun_out/r02t/golden8.log 2>&1; tail -8 gpurun_out/r02t/golden8.log
mkdir -p tests/golden; cp gpurun_out/ref_golden8/conductor.npz tests/golden/conductor.npz 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_conductor.py tests/test_gpu_adjoint.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r02t/pytest.log
