#!/bin/bash
# r04c: bucket table for the envmap cell CDF (tests + cfg 3), A/B of cfg-3 kernel variants prebuilt with tools/build_variant.py
mkdir -p gpurun_out/r04c
timeout 900 python -m pytest tests -m gpu -q -k "env or config or adjoint" 2>&1 | tail -6 | tee gpurun_out/r04c/pytest_env.log
for v in "" _i1024 _tni _outl _i1024tni; do
  L=psdr_jit_b200/libpsdr_b200$v.so
  echo "variant [$v]" | tee -a gpurun_out/r04c/cfg3_variants.log
  PSDR_B200_LIB=$PWD/$L timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --no-vjp 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('kernel_ms'))" | tee -a gpurun_out/r04c/cfg3_variants.log
done
