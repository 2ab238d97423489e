#!/bin/bash
# 2-GPU: symmetric-memory probe, fused multicast reduction tests, bench N=2 fused vs NCCL
mkdir -p gpurun_out/r02p
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/probe_symm.py 2>&1 | grep -v "^W\|warn" | tail -12 | tee gpurun_out/r02p/probe.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | tail -15 | tee gpurun_out/r02p/pytest_multi.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/r02p/n2.err | tee gpurun_out/r02p/bench_n2.json | cut -c1-600
tail -5 gpurun_out/r02p/n2.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-peer 2>gpurun_out/r02p/n2_nccl.err | tee gpurun_out/r02p/bench_n2_nccl.json | cut -c1-600
