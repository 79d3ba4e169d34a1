#!/bin/bash
# Build compile-time variants of the library here (CPU box), run with:  gpurun -- bash tools/sweep.sh run
# Each variant: name + nvcc -D flags.
cd "$(dirname "$0")/.."
VARIANTS=(
 "base:"
 "pm5:-DWRACH_PHYS_MINBLOCKS=5"
 "pm7:-DWRACH_PHYS_MINBLOCKS=7"
)
if [ "$1" = "build" ]; then
  mkdir -p wrach_b200/lib/sweep
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}; flags=${v#*:}
    (cd wrach_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC $flags -Xptxas -v -shared -o ../lib/sweep/lib_$name.so wrach_worker.cu wrach_host.cpp -ldl 2>&1 | grep -A2 "Function properties for _ZN5wrach6k_physILi1\|Function properties for _ZN5wrach7k_rebin" | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name")
  done
else
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}
    WRACH_CUDA_LIB=$PWD/wrach_b200/lib/sweep/lib_$name.so python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', 'step %.4f ms' % d['ms_per_step'], 'phys %.4f rebin+scan %.4f' % (k['k_phys']['ms'], k['k_rebin']['ms']))"
  done
fi
