#!/bin/bash
mkdir -p gpurun_out/r04l
timeout 600 python tools/ref_golden14.py > gpurun_out/r04l/golden14.log 2>&1; tail -8 gpurun_out/r04l/golden14.log
