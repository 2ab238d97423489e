#!/bin/bash
mkdir -p gpurun_out/r04r
timeout 1200 python tools/ref_probe6.py all 2>&1 | grep -v Warning | tail -30 | tee gpurun_out/r04r/probe6.log
cp gpurun_out/ref_probe6/summary.json gpurun_out/r04r/ 2>/dev/null
