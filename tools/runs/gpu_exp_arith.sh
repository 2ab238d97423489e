#!/bin/bash
# A/B on the GPU box: the cfg-0 forward kernels compiled with Dr.Jit-like approximate division / sqrt (-prec-div=false
# -prec-sqrt=false -ftz=true) against the default IEEE build, both compared with the reference goldens.
cd "$(dirname "$0")/.."
C=psdr_jit_b200/csrc; B=psdr_jit_b200/build
mkdir -p gpurun_out/exp_arith
python tools/exp_arith.py ieee 2>&1 | grep '^\[' | tee gpurun_out/exp_arith/result.txt
for V in "-prec-div=false -prec-sqrt=false -ftz=true" "-prec-div=false -ftz=true" "-prec-div=false" ; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false $V -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-O2 -x cu -c -o $B/kern_cfg0.o $C/kern_cfg0.cu 2>/dev/null
  nvcc -shared -o psdr_jit_b200/libpsdr_b200.so -ccbin /usr/bin/g++ $B/*.o
  python tools/exp_arith.py "$V" 2>&1 | grep '^\[' | tee -a gpurun_out/exp_arith/result.txt
done
# restore
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-O2 -x cu -c -o $B/kern_cfg0.o $C/kern_cfg0.cu 2>/dev/null
nvcc -shared -o psdr_jit_b200/libpsdr_b200.so -ccbin /usr/bin/g++ $B/*.o
