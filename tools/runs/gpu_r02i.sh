mkdir -p gpurun_out/r02i
nvidia-smi --query-gpu=name --format=csv | head -3 > gpurun_out/r02i/gpu.txt
for n in 8 4; do
echo "== bench cfg2 N=$n"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/r02i/b$n.err | tee gpurun_out/r02i/bench_n$n.json | cut -c1-200
tail -2 gpurun_out/r02i/b$n.err
done
echo "== cfg4 N=8"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --config 4 --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r02i/c4.err | tee gpurun_out/r02i/cfg4_n8.json | cut -c1-200
echo "== cfg5 N=8"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --config 5 --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r02i/c5.err | tee gpurun_out/r02i/cfg5_n8.json | cut -c1-200
python - <<'PY'
import json
for f in ("bench_n8","bench_n4","cfg4_n8","cfg5_n8"):
    try:
        d=json.loads(open("gpurun_out/r02i/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["kernel_ms"], "e2e", d["e2e"]["ms_per_step"], "vjp", d.get("vjp") and d["vjp"].get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
