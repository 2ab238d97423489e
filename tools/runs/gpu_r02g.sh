bash tools/gpu_round2.sh r02g tests bench
timeout 600 python tools/ref_golden7.py > gpurun_out/r02g/ref_golden7.log 2>&1; tail -4 gpurun_out/r02g/ref_golden7.log
echo "== config 4 reference"; timeout 900 python bench.py --config 4 --impl reference --steps 3 --warmup 1 2>gpurun_out/r02g/cfg4_ref.err | tee gpurun_out/r02g/cfg4_ref.json | cut -c1-700; tail -2 gpurun_out/r02g/cfg4_ref.err
echo "== config 3 ours"; timeout 900 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02g/cfg3_ours.err | tee gpurun_out/r02g/cfg3_ours.json | cut -c1-200; python -c "
import json;d=json.load(open('gpurun_out/r02g/cfg3_ours.json'));print(d['ms_per_step'],d['e2e'],d['configure_ms'])"
echo "== vjp sweep"
bash tools/gpu_vjp_sweep.sh "-DPSDR_VJP_GEO_NOINLINE=1" "-DPSDR_TRACE_NOINLINE=1" "-DPSDR_AGG_PARTIAL=0" "-DPSDR_LB_IVJP=4" "-DPSDR_LB_IVJP=3" "-DPSDR_LB_IVJP=4 -DPSDR_VJP_GEO_NOINLINE=1 -DPSDR_AGG_PARTIAL=0" 2>&1 | tee gpurun_out/r02g/vjp_sweep.log
