#!/bin/bash
# r05g: secondary-edge samples ordered along the edge list (last GPU visit of the round): GPU tests, cfg 2 and cfg 4 with the
# ordering, then cfg 4 without (PSDR_SEC_EDGE_SORT=0)
O=gpurun_out/r05g; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu.log
P="import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('kernel_ms'), (d.get('vjp') or {}).get('ms_per_step'), (d.get('vjp') or {}).get('kernel_ms'))"
timeout 100 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>$O/cfg2_sec1.err | tee $O/cfg2_sec1.json | python -c "$P"
timeout 100 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-vjp 2>$O/cfg4_sec1.err | tee $O/cfg4_sec1.json | python -c "$P"
PSDR_SEC_EDGE_SORT=0 timeout 100 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-vjp 2>$O/cfg4_sec0.err | tee $O/cfg4_sec0.json | python -c "$P"
