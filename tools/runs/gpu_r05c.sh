#!/bin/bash
# r05c: final single-GPU validation with the ordered primary-edge lanes: GPU tests, smoke, both bench arms, configs 3 4 5
O=gpurun_out/r05c; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
timeout 600 python bench.py 2>$O/bench_ours.err | tee $O/bench_ours.json | cut -c1-300
timeout 600 python bench.py --impl reference 2>$O/bench_ref.err | tee $O/bench_ref.json | cut -c1-300
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline 2>$O/cfg$c.err | tee $O/cfg$c.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('cfg', d['config']['workload'][:50], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('kernel_ms'))"
done
