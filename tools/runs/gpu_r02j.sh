bash tools/gpu_round2.sh r02j tests bench
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02j/bench_ours.json").read().strip().splitlines()[-1]); print("cfg2", d["value"], d["ms_per_step"], d["kernel_ms"], "e2e", d["e2e"]["ms_per_step"], "vjp", d["vjp"]["ms_per_step"], d["vjp"]["kernel_ms"])
PY
