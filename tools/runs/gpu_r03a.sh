#!/bin/bash
mkdir -p gpurun_out/r03a
bash tools/gpu_cfg3_sweep.sh "-DPSDR_BLOCK_I=512" "-DPSDR_BLOCK_I=384" "-DPSDR_BLOCK_I=768" "-DPSDR_BLOCK_I=256 -DPSDR_LB_INTERIOR=2 -DPSDR_LB_INTERIOR_DUAL=2" 2>&1 | grep -v "nvcc warning" | tee gpurun_out/r03a/cfg3_sweep.log
