#!/bin/bash
# r05h: the last 30 s of the round's GPU budget: cfg 2 forward step with and without the secondary-edge sample ordering, same box
O=gpurun_out/r05h; mkdir -p $O
P="import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('kernel_ms'))"
PSDR_SEC_EDGE_SORT=0 timeout 14 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-vjp 2>$O/sec0.err | tee $O/sec0.json | python -c "$P"
timeout 14 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-vjp 2>$O/sec1.err | tee $O/sec1.json | python -c "$P"
