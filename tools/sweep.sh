#!/bin/bash
# Build compile-time variants of the library here (CPU box), run with:  gpurun -- bash tools/sweep.sh run
# Each variant: name + nvcc -D flags.
cd "$(dirname "$0")/.."
VARIANTS=(
 "np:-DWRACH_REBIN_PERSISTENT=0"
 "p4b2:-DWRACH_REBIN_PERSISTENT=1 -DWRACH_REBIN_PBLOCKS=4 -DWRACH_REBIN_BATCH=2"
 "p4b4:-DWRACH_REBIN_PERSISTENT=1 -DWRACH_REBIN_PBLOCKS=4 -DWRACH_REBIN_BATCH=4"
 "p4b7:-DWRACH_REBIN_PERSISTENT=1 -DWRACH_REBIN_PBLOCKS=4 -DWRACH_REBIN_BATCH=7"
 "p3b7:-DWRACH_REBIN_PERSISTENT=1 -DWRACH_REBIN_PBLOCKS=3 -DWRACH_REBIN_BATCH=7"
)
if [ "$1" = "build" ]; then
  mkdir -p wrach_b200/lib/sweep
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}; flags=${v#*:}
    (cd wrach_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC $flags -Xptxas -v -shared -o ../lib/sweep/lib_$name.so wrach_worker.cu wrach_host.cpp -ldl 2>&1 | grep -A1 "k_rebin\|k_physILi1" | grep Used | tr '\n' ' '; echo " <- $name")
  done
else
  for v in "${VARIANTS[@]}"; do
    name=${v%%:*}
    WRACH_CUDA_LIB=$PWD/wrach_b200/lib/sweep/lib_$name.so python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', 'step %.4f ms' % d['ms_per_step'], 'phys %.4f rebin+scan %.4f' % (k['k_phys']['ms'], k['k_rebin']['ms']))"
  done
fi
