#!/bin/bash
mkdir -p gpurun_out/r04m
timeout 900 python tools/ref_probe5.py 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/r04m/probe5.log
