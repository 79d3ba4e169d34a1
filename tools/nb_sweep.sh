#!/bin/bash
# GPU box: frame time of the opt-in neighbour mode for every library variant in wrach_b200/lib/sweep/ (and the product)
cd "$(dirname "$0")/.."
for so in "" wrach_b200/lib/sweep/lib_*.so; do
  name=${so:-product}
  ms=$(WRACH_CUDA_LIB=${so:-wrach_b200/lib/libwrach_cuda.so} python bench.py --neighbours --steps 50 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernels']['k_phys']['ms'], d['config'].get('state_checksum'))")
  echo "$name $ms"
done
