mkdir -p gpurun_out/r02k
timeout 900 python -m pytest tests/test_gpu_adjoint.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02k/pytest_adjoint.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02k/b.err | tee gpurun_out/r02k/bench_ours.json | cut -c1-100
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02k/bench_ours.json").read().strip().splitlines()[-1]); print("cfg2", d["value"], d["ms_per_step"], d["kernel_ms"], "e2e", d["e2e"]["ms_per_step"], "vjp", d["vjp"]["ms_per_step"], d["vjp"]["kernel_ms"])
PY
bash tools/gpu_round2.sh r02k ncuvjp > gpurun_out/r02k/ncuvjp.log 2>&1; tail -3 gpurun_out/r02k/ncuvjp.log
