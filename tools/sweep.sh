#!/bin/bash
# Build compile-time variants of the library here (CPU box), run with:  gpurun -- bash tools/sweep.sh run
# Each variant: name + nvcc -D flags.
cd "$(dirname "$0")/.."
VARIANTS=(
 "v1p5:-DWRACH_PHYS_STAGE_VEL=1 -DWRACH_PHYS_MINBLOCKS=5"
 "v0p5:-DWRACH_PHYS_STAGE_VEL=0 -DWRACH_PHYS_MINBLOCKS=5"
 "v0p6:-DWRACH_PHYS_STAGE_VEL=0 -DWRACH_PHYS_MINBLOCKS=6"
 "v0p7:-DWRACH_PHYS_STAGE_VEL=0 -DWRACH_PHYS_MINBLOCKS=7"
 "v0p8:-DWRACH_PHYS_STAGE_VEL=0 -DWRACH_PHYS_MINBLOCKS=8"
)
if [ "$1" = "build" ]; then
  mkdir -p wrach_b200/lib/sweep
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}; flags=${v#*:}
    (cd wrach_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC $flags -Xptxas -v -shared -o ../lib/sweep/lib_$name.so wrach_worker.cu wrach_host.cpp 2>&1 | grep -A1 "k_rebin\|k_physILi1" | grep Used | tr '\n' ' '; echo " <- $name")
  done
else
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}
    WRACH_CUDA_LIB=$PWD/wrach_b200/lib/sweep/lib_$name.so python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', 'step %.4f ms' % d['ms_per_step'], 'phys %.4f rebin+scan %.4f' % (k['k_phys']['ms'], k['k_rebin']['ms']))"
  done
fi
