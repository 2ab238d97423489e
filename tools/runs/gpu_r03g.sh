#!/bin/bash
mkdir -p gpurun_out/r03g
bash tools/gpu_cfg3_sweep.sh "-DPSDR_TEXCOND=0" 2>&1 | grep -v "nvcc warning" | tee gpurun_out/r03g/cfg3_texcond.log
