#!/bin/bash
# interior-adjoint sweep: CTA size x phase barrier (instruction-cache locality experiment)
bash tools/gpu_vjp_sweep.sh "-DPSDR_VJP_PHASE_SYNC=1" "-DPSDR_VJP_PHASE_SYNC=1 -DPSDR_BLOCK_IVJP=256 -DPSDR_LB_IVJP=2" "-DPSDR_VJP_PHASE_SYNC=1 -DPSDR_BLOCK_IVJP=640 -DPSDR_LB_IVJP=1" "-DPSDR_BLOCK_IVJP=256 -DPSDR_LB_IVJP=2" "-DPSDR_BLOCK_IVJP=64 -DPSDR_LB_IVJP=10" "-DPSDR_LB_IVJP=4" "-DPSDR_LB_IVJP=6" 2>&1 | tee gpurun_out/r02l_sweep.log
