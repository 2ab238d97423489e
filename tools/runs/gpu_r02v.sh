#!/bin/bash
# N GPUs (default 8): fused multicast reduction vs NCCL, configs 4 and 5, the 2-GPU tests
N=${1:-8}
mkdir -p gpurun_out/r02v
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/r02v/n$N.err | tee gpurun_out/r02v/bench_n$N.json | cut -c1-400
tail -3 gpurun_out/r02v/n$N.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-peer 2>gpurun_out/r02v/n${N}_nccl.err | tee gpurun_out/r02v/bench_n${N}_nccl.json | cut -c1-400
timeout 600 $TR bench.py --gpus $N --config 4 --steps 10 --warmup 3 2>gpurun_out/r02v/cfg4_n$N.err | tee gpurun_out/r02v/cfg4_n$N.json | cut -c1-400
timeout 600 $TR bench.py --gpus $N --config 5 --steps 5 --warmup 3 2>gpurun_out/r02v/cfg5_n$N.err | tee gpurun_out/r02v/cfg5_n$N.json | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02v/pytest_multi.log
