#!/bin/bash
# r05d (2 GPUs): multi-GPU tests; bench at N = 2 with the primary-edge lanes ordered (default) and in lane order
O=gpurun_out/r05d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 | tee $O/pytest_multi.log
for b in 512 0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --edge-sort $b 2>$O/bench_n2_bins$b.err | tee $O/bench_n2_bins$b.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('bins $b', d['ms_per_step'], d['e2e']['ms_per_step'], d['vjp']['ms_per_step'])"
done
for f in $O/*.err; do tail -n 2 $f; done
