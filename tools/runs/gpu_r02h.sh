mkdir -p gpurun_out/r02h
nvidia-smi --query-gpu=name --format=csv > gpurun_out/r02h/gpu.txt
echo "== pytest multi + bvh"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu or bvh or brute" -s 2>&1 | tail -12 | tee gpurun_out/r02h/pytest.log
for n in 1 2; do
echo "== bench cfg2 N=$n"
if [ $n = 1 ]; then timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02h/b$n.err | tee gpurun_out/r02h/bench_n$n.json | cut -c1-300; else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/r02h/b$n.err | tee gpurun_out/r02h/bench_n$n.json | cut -c1-300; fi
tail -2 gpurun_out/r02h/b$n.err
done
echo "== cfg4 N=1"; timeout 600 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02h/c4.err | tee gpurun_out/r02h/cfg4_n1.json | cut -c1-200
echo "== cfg4 N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --config 4 --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r02h/c4b.err | tee gpurun_out/r02h/cfg4_n2.json | cut -c1-200
python - <<'PY'
import json
for f in ("bench_n1","bench_n2","cfg4_n1","cfg4_n2"):
    try:
        d=json.loads(open("gpurun_out/r02h/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["ms_per_step"], d.get("vjp",{}) and d["vjp"].get("ms_per_step"), d.get("configure_ms"))
    except Exception as e: print(f, "ERR", e)
PY
