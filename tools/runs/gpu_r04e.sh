#!/bin/bash
# r04e (2 GPUs): multi-GPU tests incl. the shared host buffer, bench at N = 2 with and without the split D2H
mkdir -p gpurun_out/r04e
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -8 | tee gpurun_out/r04e/pytest_multi.log
for flag in "" "--no-shared-d2h"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline $flag 2>gpurun_out/r04e/bench_n2$flag.err | tee gpurun_out/r04e/bench_n2$flag.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"
done
tail -3 gpurun_out/r04e/*.err
