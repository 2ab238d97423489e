#!/bin/bash
mkdir -p gpurun_out/r03b
bash tools/gpu_cfg3_sweep.sh "-DPSDR_FULL_OUTLINE=1" "-DPSDR_FULL_OUTLINE=1 -DPSDR_BLOCK_I=768" "-DPSDR_BLOCK_I=896" 2>&1 | grep -v "nvcc warning" | tee gpurun_out/r03b/cfg3_sweep.log
