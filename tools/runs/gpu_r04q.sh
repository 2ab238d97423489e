#!/bin/bash
# r04q: A/B of the BVH traversal loop (one unit of work per iteration + distance-tagged stack) on cfg 4, with the BVH parity tests
mkdir -p gpurun_out/r04q
for v in "" _bvh2; do
  L=$PWD/psdr_jit_b200/libpsdr_b200$v.so
  echo "variant [$v]" | tee -a gpurun_out/r04q/cfg4_variants.log
  PSDR_B200_LIB=$L timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q -k "bvh or cfg4 or bunny or sphere" 2>&1 | tail -2 | tee -a gpurun_out/r04q/cfg4_variants.log
  PSDR_B200_LIB=$L timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline --no-vjp 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['value'], d.get('kernel_ms'))" | tee -a gpurun_out/r04q/cfg4_variants.log
done
