#!/bin/bash
# r04a: extended material set (RoughDielectric / MicrofacetPerVertex / NormalMap) -- GPU tests vs the oracle, goldens of the
# running reference, full GPU suite, cfg 3 with the split kernel families, default bench
mkdir -p gpurun_out/r04a
timeout 900 python -m pytest tests/test_gpu_ext_bsdfs.py -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/r04a/pytest_ext.log
timeout 700 python tools/ref_golden10.py > gpurun_out/r04a/golden10.log 2>&1; tail -25 gpurun_out/r04a/golden10.log
if [ -f gpurun_out/ref_golden10/ext_bsdfs.npz ]; then cp gpurun_out/ref_golden10/ext_bsdfs.npz tests/golden/ext_bsdfs.npz; fi
timeout 600 python -m pytest tests/test_gpu_ext_bsdfs.py -m gpu -q -k "golden" 2>&1 | tail -15 | tee gpurun_out/r04a/pytest_golden.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r04a/pytest_all.log
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r04a/cfg3.err | tee gpurun_out/r04a/cfg3.json | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r04a/bench.err | tee gpurun_out/r04a/bench_ours.json | cut -c1-600
