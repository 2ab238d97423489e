#!/bin/bash
# final multi-GPU record: N GPUs
N=${1:-8}
mkdir -p gpurun_out/r03f
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/r03f/n$N.err | tee gpurun_out/r03f/bench_n$N.json | cut -c1-300
tail -3 gpurun_out/r03f/n$N.err
timeout 600 $TR bench.py --gpus $N --config 4 --steps 10 --warmup 3 2>gpurun_out/r03f/cfg4_n$N.err | tee gpurun_out/r03f/cfg4_n$N.json | cut -c1-300
timeout 600 $TR bench.py --gpus $N --config 5 --steps 5 --warmup 3 2>gpurun_out/r03f/cfg5_n$N.err | tee gpurun_out/r03f/cfg5_n$N.json | cut -c1-300
timeout 600 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 2>gpurun_out/r03f/ref_n$N.err | tee gpurun_out/r03f/ref_n$N.json | cut -c1-300
