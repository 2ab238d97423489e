#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out/r02y
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02y/pytest_multi.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/r02y/n$N.err | tee gpurun_out/r02y/bench_n$N.json | cut -c1-300
tail -3 gpurun_out/r02y/n$N.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-peer 2>gpurun_out/r02y/n${N}_nccl.err | tee gpurun_out/r02y/bench_n${N}_nccl.json | cut -c1-300
