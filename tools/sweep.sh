#!/bin/bash
# Build compile-time variants of the library here (CPU box):  bash tools/sweep.sh "name:-Dflags" ...
# then run them on one box in one call:  gpurun -- bash tools/sweep_run.sh [repeats]
cd "$(dirname "$0")/.."
mkdir -p wrach_b200/lib/sweep
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  (cd wrach_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC $flags -Xptxas -v -shared -o ../lib/sweep/lib_$name.so wrach_worker.cu wrach_host.cpp -ldl 2>&1 | grep -A2 "Function properties for _ZN5wrach6k_physILi1\|Function properties for _ZN5wrach7k_rebin" | grep -E "Used" | sed 's/ptxas info    : //' | tr '\n' '|'; echo " <- $name")
done
