#!/bin/bash
mkdir -p gpurun_out/r04f
timeout 300 python tools/probe_d2h.py 2>&1 | tail -8 | tee gpurun_out/r04f/probe_d2h.log
