#!/bin/bash
# new default CTA shapes: full GPU test suite + bench, then a fine block-size sweep around the new defaults
mkdir -p gpurun_out/r02o
bash tools/gpu_round2.sh r02o tests bench
for c in 1 3 4; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02o/cfg$c.err | tee gpurun_out/r02o/cfg$c.json | cut -c1-400; done
S1=""
bash tools/gpu_lb_sweep.sh \
  "-DPSDR_BLOCK_I=512 -DPSDR_BLOCK_P=768 -DPSDR_BLOCK_S=896" \
  "-DPSDR_BLOCK_I=576 -DPSDR_BLOCK_P=832 -DPSDR_BLOCK_S=768" \
  "-DPSDR_BLOCK_I=704 -DPSDR_BLOCK_P=960 -DPSDR_BLOCK_S=640" \
  "-DPSDR_BLOCK_I=448 -DPSDR_BLOCK_P=640 -DPSDR_BLOCK_S=512" \
  2>&1 | grep -v "nvcc warning" | tee gpurun_out/r02o/fwd_sweep.log
