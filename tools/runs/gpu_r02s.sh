#!/bin/bash
mkdir -p gpurun_out/r02s
python tools/exp_vjp_scaling.py 2>&1 | grep -v Warn | tee gpurun_out/r02s/scaling.log
