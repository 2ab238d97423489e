bash tools/gpu_round2.sh r02c tests
timeout 600 python tools/ref_golden5.py > gpurun_out/r02c/ref_golden5.log 2>&1; tail -3 gpurun_out/r02c/ref_golden5.log
PARITY_SKIP_CENSUS=1 timeout 1500 python tools/ref_parity.py all > gpurun_out/r02c/parity.log 2>&1; tail -25 gpurun_out/r02c/parity.log | cut -c1-1500
