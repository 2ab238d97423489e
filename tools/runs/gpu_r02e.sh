bash tools/gpu_round2.sh r02e tests ncuvjp
timeout 600 python tools/ref_golden6.py > gpurun_out/r02e/ref_golden6.log 2>&1; tail -4 gpurun_out/r02e/ref_golden6.log
