#!/bin/bash
mkdir -p gpurun_out/r04n
timeout 900 python tools/ref_parity_scaled.py all 2>&1 | grep -v Warning | tail -70 | tee gpurun_out/r04n/parity_scaled.log
cp gpurun_out/parity_scaled/summary.json gpurun_out/r04n/ 2>/dev/null
