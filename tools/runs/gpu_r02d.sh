bash tools/gpu_round2.sh r02d tests bench
PARITY_SKIP_CENSUS=1 timeout 1500 python tools/ref_parity.py all > gpurun_out/r02d/parity.log 2>&1; tail -14 gpurun_out/r02d/parity.log | cut -c1-2500
cp gpurun_out/parity/summary.json gpurun_out/r02d/parity_summary.json
