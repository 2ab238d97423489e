#!/bin/bash
# r04o: final validation of the round: GPU tests, smoke, both bench arms at N = 1, the other BASELINE configs
mkdir -p gpurun_out/r04o
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r04o/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r04o/smoke.log
timeout 600 python bench.py 2>gpurun_out/r04o/bench_ours.err | tee gpurun_out/r04o/bench_ours.json | cut -c1-400
timeout 900 python bench.py --impl reference 2>gpurun_out/r04o/bench_ref.err | tee gpurun_out/r04o/bench_ref.json | cut -c1-400
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r04o/cfg$c.err | tee gpurun_out/r04o/cfg$c.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('cfg', d['config']['workload'][:50], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
done
