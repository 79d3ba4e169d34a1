#!/bin/bash
# Run bench.py once per prebuilt library variant in wrach_b200/lib/sweep/ (same box, same call):
#   gpurun -- bash tools/sweep_run.sh [repeats] [workload] [steps]
cd "$(dirname "$0")/.."
for rep in $(seq 1 ${1:-1}); do
for so in wrach_b200/lib/sweep/lib_*.so; do
  name=$(basename $so .so)
  WRACH_CUDA_LIB=$PWD/$so python bench.py --workload ${2:-16m} --steps ${3:-100} --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('$name', 'step %.4f ms' % d['ms_per_step'], 'phys %.4f rebin+scan %.4f' % (k['k_phys']['ms'], k['k_rebin']['ms']))"
done
done
